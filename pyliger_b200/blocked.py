"""Blocked device copies of X for the two streaming passes (see csrc/blocked.cuh).

Built once per factorisation from the device CSR with torch sort / bincount / cumsum (plumbing only; the
arithmetic passes over these structures are libinmf_b200.so kernels).

pass 1  "gene-blocked CSR":  pairs ordered by (dataset, gene block, cell, gene);  tile = rows of Mt
pass 2  "cell-blocked CSC":  pairs ordered by (dataset, cell block, gene, cell);  tile = rows of H
"""
import numpy as np
import torch

from . import _lib

WORK_FIELDS = 8
SMEM_BUDGET = 227 * 1024 - 128 - 2048
USE_DEVICE_BUILD = True      # count / scan / fill kernels (csrc/blockfmt.cuh); False = generic torch sort path
BUILD2_SMEM = 200 * 1024     # uint16 counters [W][genes] of the format-2 builder


def _sfx(vals):
    return "f32" if vals.dtype == torch.float32 else "f64"


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _empty_pairs(nnz, vals):
    return torch.empty(nnz, 2, dtype=torch.int32 if vals.dtype == torch.float32 else torch.int64, device=vals.device)


def _device_format1(X, kp, esize, NB, gbs):
    """ptr, pairs of the gene-blocked CSR over the owned rows through inmf_blockfmt1_* (same arrays as the sort path)."""
    dev = X.rowptr.device
    seg1 = [[int(X.own_off[s]), X.n_own[s], int(X.row_base[s]) + (X.own[s][0] - X.bounds[s][0]), 0]
            for s in range(X.N)]
    seg1_t = torch.tensor(seg1, dtype=torch.int64, device=dev)
    n_tot = X.n_own_total
    ptr = torch.zeros(NB * n_tot + 1, dtype=torch.int64, device=dev)
    sfx = _sfx(X.vals)
    _lib.call("inmf_blockfmt1", sfx, 0, X.rowptr, X.colidx, X.vals, seg1_t, X.N, n_tot, NB, gbs, row_bytes(kp, esize),
              ptr[1:], None, _stream(dev))
    torch.cumsum(ptr[1:], 0, out=ptr[1:])
    pairs = _empty_pairs(int(ptr[-1].item()), X.vals)
    _lib.call("inmf_blockfmt1", sfx, 1, X.rowptr, X.colidx, X.vals, seg1_t, X.N, n_tot, NB, gbs, row_bytes(kp, esize),
              ptr, pairs, _stream(dev))
    return ptr, pairs


def _device_format2(X, kp, esize, nb, cbs, blk_off):
    """ptr, pairs of the cell-blocked CSC through inmf_blockfmt2_*; None if the gene counters do not fit."""
    dev = X.rowptr.device
    g = X.g
    W = min(16, BUILD2_SMEM // (2 * g))
    if W < 1 or max(cbs) > 60000 * W:
        return None
    blk2 = []
    for s in range(X.N):
        r0 = int(X.row_base[s]) + (X.own[s][0] - X.bounds[s][0])
        for b in range(nb[s]):
            blk2.append([r0 + b * cbs[s], min(cbs[s], X.n_own[s] - b * cbs[s]), g * (int(blk_off[s]) + b), 0])
    total_blocks = len(blk2)
    ptr = torch.zeros(g * total_blocks + 1, dtype=torch.int64, device=dev)
    if total_blocks == 0:
        return ptr, _empty_pairs(0, X.vals)
    blk2_t = torch.tensor(blk2, dtype=torch.int64, device=dev)
    offs = torch.empty(total_blocks * W * g, dtype=torch.int16, device=dev)
    sfx = _sfx(X.vals)
    _lib.call("inmf_blockfmt2", sfx, 0, X.rowptr, X.colidx, X.vals, blk2_t, total_blocks, g, row_bytes(kp, esize), W,
              ptr[1:], None, offs, None, _stream(dev))
    torch.cumsum(ptr[1:], 0, out=ptr[1:])
    pairs = _empty_pairs(int(ptr[-1].item()), X.vals)
    _lib.call("inmf_blockfmt2", sfx, 1, X.rowptr, X.colidx, X.vals, blk2_t, total_blocks, g, row_bytes(kp, esize), W,
              None, ptr, offs, pairs, _stream(dev))
    return ptr, pairs


# Kernel variants of the row-ordered format (module attributes, not environment switches; tests flip them to compare
# the variants).  Split-tile kernel (csrc/blocked.cuh, inmf_blocked_split_f32) for fp32 and 32 < KP <= 64: [rows x 32] +
# [rows x rs] tile, pairs carry row * 128.  Wide kernel for fp32, KP == 32 (inmf_blocked_wide_f32): same formats as the
# plain kernel, chosen at call time.  The full passes of fp32 runs use the sliced format below (build_sell) instead.
SPLIT_TILE = True
WIDE_TILE = True
SELL = True  # sliced (SELL-8) passes for fp32 full passes (csrc/sell.cuh)


def split_applies(kp, esize):
    return SPLIT_TILE and esize == 4 and 32 < kp <= 64


def row_bytes(kp, esize):
    """Stride that the pairs' byte offsets are premultiplied with (esize = 2: the bf16 tile of the mixed mode)."""
    return 128 if split_applies(kp, esize) else kp * esize


def smem_row_bytes(kp, esize):
    if split_applies(kp, esize):
        rem = kp - 32
        return (32 + (8 if rem <= 8 else 16 if rem <= 16 else 32)) * 4
    return kp * esize


def tile_cap_rows(kp, esize):
    return SMEM_BUDGET // smem_row_bytes(kp, esize)


def _pack_pairs(idx_local, vals):
    """(int byte offset, T val) -> packed pairs: f32 {int32, float32} = int32[nnz, 2]; f64 {int64, float64}."""
    if vals.dtype == torch.float32:
        out = torch.empty(idx_local.numel(), 2, dtype=torch.int32, device=vals.device)
        out[:, 0] = idx_local.to(torch.int32)
        out[:, 1] = vals.view(torch.int32)
    else:
        out = torch.empty(idx_local.numel(), 2, dtype=torch.int64, device=vals.device)
        out[:, 0] = idx_local.to(torch.int64)
        out[:, 1] = vals.view(torch.int64)
    return out


class BlockedPass(object):
    __slots__ = ("work", "nwork", "ptr", "pairs", "max_tile_rows", "split", "work_host", "blocks")


class SellPass(object):
    """Sliced form of a BlockedPass (csrc/sell.cuh): rows of every work item sorted by length, 8 per slice."""
    __slots__ = ("work", "nwork", "slice_ptr", "rowid", "pairs4", "max_tile_rows", "nnz", "padded", "height", "nparts",
                 "work_host", "row_lo", "row_hi")


SELL_HEIGHT_KP32 = 32  # slice height at kp == 32: 32 (one lane per row) or 8 (four lanes per row)
SELL_HEIGHT_KP40 = 33  # 32 < kp <= 40: 33 (one lane per row, residue-ordered non-zeros) or 8 (four lanes per row)


def build_sell(bp, kp, height=None):
    """Sliced copy of the row-ordered blocked format ``bp`` (fp32).  Index plumbing (sort of the rows of every work item
    by length, slice table, prefix sums) is torch; the data movement is inmf_sell_fill_f32."""
    dev = bp.ptr.device
    W = bp.work_host[:bp.nwork]
    sp = SellPass()
    # -32: parity-ordered rows (mixed precision); 33: 32 residue-ordered rows (32 < kp <= 40)
    sp.height = int(height or (SELL_HEIGHT_KP32 if kp == 32 else (SELL_HEIGHT_KP40 if 32 < kp <= 40 else 8)))
    S = 32 if sp.height == 33 else abs(sp.height)
    sp.max_tile_rows = bp.max_tile_rows
    sp.nwork = bp.nwork
    sp.nnz = int(bp.pairs.shape[0])
    if bp.nwork == 0:
        sp.work = torch.zeros(1, WORK_FIELDS, dtype=torch.int64, device=dev)
        sp.slice_ptr = torch.zeros(1, dtype=torch.int64, device=dev)
        sp.rowid = torch.zeros(S, dtype=torch.int32, device=dev)
        sp.pairs4 = torch.zeros(1, 4, dtype=torch.int32, device=dev)
        sp.padded = 0
        sp.nparts = 1
        return sp
    lo = W[:, 0] + W[:, 3]
    nrows = W[:, 4] - W[:, 3]
    ns = (nrows + S - 1) // S
    start = np.concatenate([[0], np.cumsum(nrows)])
    slice_off = np.concatenate([[0], np.cumsum(ns)])
    total_rows, total_slices = int(start[-1]), int(slice_off[-1])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(dev)
    wid = torch.repeat_interleave(torch.arange(bp.nwork, device=dev), t(nrows))
    q = torch.arange(total_rows, device=dev) - t(start[:-1])[wid]          # position inside the work item
    ridx = q + t(lo)[wid]                                                   # row of the blocked format (ptr index)
    length = bp.ptr[ridx + 1] - bp.ptr[ridx]
    big = 1 << 22
    if total_rows and int(length.max().item()) >= big:
        raise ValueError("sparse row too long for the sliced format")
    order = torch.sort(wid * big + (big - 1 - length), stable=True)[1]     # by work item, longest row first
    slot = t(slice_off[:-1])[wid] * S + q
    slice_rows = torch.full((total_slices * S,), -1, dtype=torch.int64, device=dev)
    slice_rows[slot] = ridx[order]
    slice_len = torch.zeros(total_slices * S, dtype=torch.int64, device=dev)
    slice_len[slot] = length[order]
    L2 = (slice_len.view(total_slices, S).max(dim=1)[0] + 1) // 2           # 128-bit units (two steps) per row
    sp.slice_ptr = torch.zeros(total_slices + 1, dtype=torch.int64, device=dev)
    torch.cumsum(S * L2, 0, out=sp.slice_ptr[1:])
    n4 = int(sp.slice_ptr[-1].item())
    sp.padded = 2 * n4
    swid = torch.repeat_interleave(torch.arange(bp.nwork, device=dev), t(ns))
    rel = slice_rows.view(total_slices, S) - t(W[:, 0])[swid][:, None]
    sp.rowid = torch.where(slice_rows.view(total_slices, S) >= 0, rel, torch.full_like(rel, -1)).to(torch.int32).contiguous()
    sp.pairs4 = torch.empty(max(n4, 1), 4, dtype=torch.int32, device=dev)
    _lib.call("inmf_sell_fill_f32", "", bp.ptr, bp.pairs, slice_rows, sp.slice_ptr, total_slices, sp.height, sp.pairs4, _stream(dev))
    work = np.zeros((bp.nwork, WORK_FIELDS), dtype=np.int64)
    work[:, 0] = slice_off[:-1]
    work[:, 1] = slice_off[1:]
    work[:, 2] = W[:, 1]
    work[:, 3] = W[:, 2]
    work[:, 4] = W[:, 5]
    work[:, 5] = W[:, 6]
    work[:, 6] = W[:, 7]
    sp.nparts = int(W[:, 7].max()) + 1
    sp.row_lo, sp.row_hi = W[:, 3].copy(), W[:, 4].copy()  # output rows of every item, relative to its out_row_base
    sp.work_host = work
    sp.work = torch.from_numpy(work).to(dev)
    return sp


def mc_work(sp, g):
    """Work array of the fused pass 2 + all-reduce (inmf_sell_pass_mc_f32) for the sliced pass ``sp`` over datasets of
    ``g`` genes: the items that add into one (dataset, gene range) share an arrival counter -- field 6 = counter id,
    field 7 = (first row << 32) | end row of the range, rows relative to out_row_base.  Returns (work, expected) with
    expected[c] = number of items of counter c."""
    W = sp.work_host.copy()
    _, cid, expected = np.unique(W[:, 4] * (g + 1) + sp.row_lo, return_inverse=True, return_counts=True)
    W[:, 6] = cid
    W[:, 7] = (sp.row_lo.astype(np.int64) << 32) | sp.row_hi.astype(np.int64)
    return W, expected.astype(np.int32)


def sell_applies(kp, esize):
    return SELL and esize == 4 and 32 <= kp <= 64


def _own_nnz(X):
    """Per-nnz (own_row, col, val) of the rows this rank owns, in CSR order. own_row indexes the
    concatenated owned rows (the H buffer)."""
    dev = X.rowptr.device
    rows, cols, vals = [], [], []
    for s in range(X.N):
        r0 = int(X.row_base[s]) + (X.own[s][0] - X.bounds[s][0])
        n = X.n_own[s]
        if n == 0:
            continue
        rp = X.rowptr[r0:r0 + n + 1]
        p0, p1 = int(rp[0].item()), int(rp[-1].item())
        counts = rp[1:] - rp[:-1]
        rows.append(torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64) + int(X.own_off[s]), counts))
        cols.append(X.colidx[p0:p1])
        vals.append(X.vals[p0:p1])
    if not rows:
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        return z, torch.zeros(0, dtype=torch.int32, device=dev), X.vals[:0]
    return torch.cat(rows), torch.cat(cols), torch.cat(vals)


MIN_ITEM_ROWS = 1024  # a CTA needs a few slices per warp to balance its warps


def _wave_chunks(n_sms, pairs, rows):
    """Row chunks per (dataset, gene block) pair so that the grid fills whole waves of CTAs: with c chunks the grid is
    pairs * c CTAs on n_sms SMs; pick the c (up to 4 waves) with the best occupancy of its last wave.  C4 on one of
    8 GPUs: 32 pairs -> 4 chunks = 128 CTAs (86 % of one wave) before, 9 chunks = 288 CTAs (97 % of two waves) now."""
    best_c, best_eff = 1, 0.0
    for waves in (1, 2, 3, 4):
        c = (waves * n_sms) // pairs
        if c < 1 or (c > 1 and rows // c < MIN_ITEM_ROWS):
            continue
        eff = pairs * c / float(-(-pairs * c // n_sms) * n_sms)
        if eff >= 0.95:  # good enough: fewer, larger items balance better inside a CTA
            return c
        if eff > best_eff + 1e-9:
            best_c, best_eff = c, eff
    return max(1, best_c)


def _wave_blocks(n_sms, n_own, cap):
    """Cell blocks per dataset for pass 2: at least ceil(n / cap) each (the tile must fit), in total a whole number of
    waves of CTAs (C4 on one of 8 GPUs: 8 x 22 = 176 blocks = 1.19 waves before, 8 x 37 = 296 = 2 waves now)."""
    need = [0 if n == 0 else -(-n // cap) for n in n_own]
    n_tot = max(1, sum(n_own))
    waves = max(1, -(-sum(need) // n_sms))
    target = waves * n_sms
    nb = [0 if n == 0 else max(nd, int(n / n_tot * target)) for n, nd in zip(n_own, need)]
    # hand the remaining CTAs of the last wave to the datasets with the largest blocks
    while sum(nb) < target:
        i = max(range(len(nb)), key=lambda j: (n_own[j] / nb[j]) if nb[j] else -1.0)
        if nb[i] == 0 or n_own[i] // (nb[i] + 1) < 64:
            break
        nb[i] += 1
    return nb


def build_pass1(X, kp, esize, n_sms):
    """Gene-blocked CSR over the owned cells."""
    dev = X.rowptr.device
    g, N = X.g, X.N
    cap = tile_cap_rows(kp, esize)
    NB = max(1, -(-g // cap))
    gbs = -(-g // NB)
    bp = BlockedPass()
    bp.split = split_applies(kp, esize)
    if USE_DEVICE_BUILD and dev.type == "cuda" and X.n_own_total > 0:
        bp.ptr, bp.pairs = _device_format1(X, kp, esize, NB, gbs)
    else:
        own_row, col, val = _own_nnz(X)
        own_off = torch.from_numpy(np.asarray(X.own_off, dtype=np.int64)).to(dev)
        n_own = torch.from_numpy(np.asarray(X.n_own, dtype=np.int64)).to(dev)
        seg = torch.bucketize(own_row, own_off[1:], right=True)
        gb = (col.to(torch.int64) // gbs)
        local_row = own_row - own_off[seg]
        cidx = NB * own_off[seg] + gb * n_own[seg] + local_row
        order = torch.sort(cidx, stable=True)[1]
        n_tot = X.n_own_total
        ptr = torch.zeros(NB * n_tot + 1, dtype=torch.int64, device=dev)
        if cidx.numel():
            ptr[1:] = torch.cumsum(torch.bincount(cidx, minlength=NB * n_tot), 0)
        idx_local = (col.to(torch.int64) - gb * gbs)[order]
        bp.pairs = _pack_pairs(idx_local * row_bytes(kp, esize), val[order])  # byte offset of the gathered tile row
        bp.ptr = ptr
    chunks = _wave_chunks(n_sms, max(1, N * NB), max(X.n_own) if N else 0)
    work = []
    for s in range(N):
        n = X.n_own[s]
        if n == 0:
            continue
        per = -(-n // chunks)
        for b in range(NB):
            rows_b = min(gbs, g - b * gbs)
            if rows_b <= 0:
                continue
            for lo in range(0, n, per):
                work.append([NB * int(X.own_off[s]) + b * n, s * g + b * gbs, rows_b, lo, min(n, lo + per),
                             int(X.own_off[s]), -1, b])  # last field: partial buffer of the deterministic mode
    bp.nwork = len(work)
    bp.work_host = np.asarray(work if work else [[0] * WORK_FIELDS], dtype=np.int64)
    bp.work = torch.from_numpy(bp.work_host).to(dev)
    bp.max_tile_rows = min(gbs, g)
    return bp


def build_pass2(X, kp, esize, n_sms, cbs=None, chunks=None):
    """Cell-blocked CSC over the owned cells.  ``cbs``: cells per block of every dataset (default: blocks sized to fill
    whole waves of CTAs); ``chunks``: gene ranges per block (default: what fills the SMs once)."""
    dev = X.rowptr.device
    g, N = X.g, X.N
    cap = tile_cap_rows(kp, esize)
    n_tot = max(1, X.n_own_total)
    if cbs is None:
        nb = _wave_blocks(n_sms, X.n_own, cap)
        cbs = [1 if b == 0 else -(-n // b) for n, b in zip(X.n_own, nb)]
    else:
        cbs = [max(1, min(int(c), cap)) for c in cbs]
    nb = [0 if n == 0 else -(-n // c) for n, c in zip(X.n_own, cbs)]
    blk_off = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
    total_blocks = int(blk_off[-1])
    bp = BlockedPass()
    bp.split = split_applies(kp, esize)
    built = None
    if USE_DEVICE_BUILD and dev.type == "cuda" and X.n_own_total > 0:
        built = _device_format2(X, kp, esize, nb, cbs, blk_off)
    if built is not None:
        bp.ptr, bp.pairs = built
    else:
        own_row, col, val = _own_nnz(X)
        own_off = torch.from_numpy(np.asarray(X.own_off, dtype=np.int64)).to(dev)
        seg = torch.bucketize(own_row, own_off[1:], right=True)
        local_row = own_row - own_off[seg]
        cbs_t = torch.tensor(cbs, dtype=torch.int64, device=dev)
        blk_off_t = torch.from_numpy(blk_off).to(dev)
        cb = local_row // cbs_t[seg]
        cidx = (blk_off_t[seg] + cb) * g + col.to(torch.int64)
        order = torch.sort(cidx, stable=True)[1]
        ptr = torch.zeros(g * total_blocks + 1, dtype=torch.int64, device=dev)
        if cidx.numel():
            ptr[1:] = torch.cumsum(torch.bincount(cidx, minlength=g * total_blocks), 0)
        idx_local = (local_row - cb * cbs_t[seg])[order]
        bp.pairs = _pack_pairs(idx_local * row_bytes(kp, esize), val[order])  # byte offset of the gathered tile row
        bp.ptr = ptr
    if chunks is None:
        chunks = max(1, n_sms // max(1, total_blocks))
    per = -(-g // max(1, int(chunks)))
    glo = np.arange(0, g, per, dtype=np.int64)
    nch = len(glo)
    # items in (dataset, block, gene range) order: item index = (blk_off[s] + b) * nch + ci
    blk_s = np.repeat(np.arange(N, dtype=np.int64), nb)
    blk_b = np.arange(total_blocks, dtype=np.int64) - blk_off[:-1][blk_s] if total_blocks else np.zeros(0, np.int64)
    n_own = np.asarray(X.n_own, dtype=np.int64)
    cbs_a = np.asarray(cbs, dtype=np.int64)
    own_off = np.asarray(X.own_off, dtype=np.int64)
    work = np.zeros((total_blocks, nch, WORK_FIELDS), dtype=np.int64)
    if total_blocks:
        work[:, :, 0] = (g * np.arange(total_blocks, dtype=np.int64))[:, None]
        work[:, :, 1] = (own_off[blk_s] + blk_b * cbs_a[blk_s])[:, None]
        work[:, :, 2] = np.minimum(cbs_a[blk_s], n_own[blk_s] - blk_b * cbs_a[blk_s])[:, None]
        work[:, :, 3] = glo[None, :]
        work[:, :, 4] = np.minimum(g, glo + per)[None, :]
        work[:, :, 5] = (blk_s * g)[:, None]
        work[:, :, 6] = -1
        work[:, 0, 6] = blk_s
        work[:, :, 7] = blk_b[:, None]
    work = work.reshape(total_blocks * nch, WORK_FIELDS)
    bp.nwork = len(work)
    bp.work_host = np.ascontiguousarray(work) if len(work) else np.zeros((1, WORK_FIELDS), dtype=np.int64)
    bp.blocks = (list(nb), list(cbs), blk_off, nch)
    bp.work = torch.from_numpy(bp.work_host).to(dev)
    bp.max_tile_rows = max(cbs) if cbs else 1
    return bp


# ------------------------------------------------------------------------------------------------------
# Dense-tile (tcgen05) formats: pairs grouped by (128-row tile, 64-wide contraction chunk, 16-row band)
# ------------------------------------------------------------------------------------------------------
TC_M, TC_KC, TC_BANDS = 128, 64, 8


def _tc_a_offset(m, kk):
    """Byte offset of element (m, kk) in the 128 x 64 fp32 UMMA A tile (K-major, no swizzle)."""
    return (m >> 3) * ((TC_KC // 4) * 128) + (m & 7) * 16 + (kk >> 2) * 128 + (kk & 3) * 4


class TcPass(object):
    __slots__ = ("work", "nwork", "ptr", "pairs", "segd", "nseg", "total_chunks", "max_seg_chunks", "npad")


def _tc_finish(tp, key, n_keys, m, kk, val, work, segd, total_chunks, max_seg_chunks, npad, dev):
    order = torch.sort(key, stable=True)[1]
    ptr = torch.zeros(n_keys + 1, dtype=torch.int64, device=dev)
    if key.numel():
        ptr[1:] = torch.cumsum(torch.bincount(key, minlength=n_keys), 0)
    tp.ptr = ptr
    tp.pairs = _pack_pairs(_tc_a_offset(m, kk)[order], val[order])
    tp.nwork = len(work)
    tp.work = torch.tensor(work if work else [[0] * WORK_FIELDS], dtype=torch.int64, device=dev)
    tp.segd = torch.tensor(segd, dtype=torch.int64, device=dev)
    tp.nseg = len(segd)
    tp.total_chunks = int(total_chunks)
    tp.max_seg_chunks = int(max_seg_chunks)
    tp.npad = npad
    return tp


def build_tc_pass1(X, k):
    """rows = cells (tiles of 128 per dataset), contraction = genes (chunks of 64), dense operand = Mt."""
    dev = X.rowptr.device
    g, N = X.g, X.N
    npad = max(16, -(-k // 16) * 16)
    nch = -(-g // TC_KC)
    nt = [-(-n // TC_M) for n in X.n_own]
    tile_off = np.concatenate([[0], np.cumsum(nt)]).astype(np.int64)
    own_row, col, val = _own_nnz(X)
    own_off = torch.from_numpy(np.asarray(X.own_off, dtype=np.int64)).to(dev)
    seg = torch.bucketize(own_row, own_off[1:], right=True)
    local_row = own_row - own_off[seg]
    col = col.to(torch.int64)
    tile = torch.from_numpy(tile_off).to(dev)[seg] + local_row // TC_M
    m, chunk, kk = local_row % TC_M, col // TC_KC, col % TC_KC
    key = (tile * nch + chunk) * TC_BANDS + m // 16
    work = []
    for s in range(N):
        for tl in range(nt[s]):
            work.append([(int(tile_off[s]) + tl) * nch, s * nch, nch, int(X.own_off[s]) + tl * TC_M,
                         min(TC_M, X.n_own[s] - tl * TC_M), 0, 0, 0])
    segd = [[s * g, g, s * nch] for s in range(N)]
    return _tc_finish(TcPass(), key, int(tile_off[-1]) * nch * TC_BANDS, m, kk, val, work, segd, N * nch, nch, npad, dev)


def build_tc_pass2(X, k, n_sms):
    """rows = genes (tiles of 128), contraction = the dataset's cells (chunks of 64), dense operand = H."""
    dev = X.rowptr.device
    g, N = X.g, X.N
    npad = max(16, -(-k // 16) * 16)
    ntg = -(-g // TC_M)
    nch = [-(-n // TC_KC) for n in X.n_own]
    chunk_off = np.concatenate([[0], np.cumsum(nch)]).astype(np.int64)
    pbase = np.concatenate([[0], np.cumsum([ntg * c for c in nch])]).astype(np.int64)
    own_row, col, val = _own_nnz(X)
    own_off = torch.from_numpy(np.asarray(X.own_off, dtype=np.int64)).to(dev)
    seg = torch.bucketize(own_row, own_off[1:], right=True)
    local_row = own_row - own_off[seg]
    col = col.to(torch.int64)
    nch_t = torch.tensor(nch, dtype=torch.int64, device=dev)
    pbase_t = torch.from_numpy(pbase).to(dev)
    m, kk = col % TC_M, local_row % TC_KC
    pidx = pbase_t[seg] + (col // TC_M) * nch_t[seg] + local_row // TC_KC
    key = pidx * TC_BANDS + m // 16
    n_items = max(1, N * ntg)
    split = max(1, -(-3 * n_sms // n_items))
    work = []
    for s in range(N):
        if nch[s] == 0:
            continue
        parts = min(split, nch[s])
        per = -(-nch[s] // parts)
        for tl in range(ntg):
            for c0 in range(0, nch[s], per):
                work.append([int(pbase[s]) + tl * nch[s] + c0, int(chunk_off[s]) + c0, min(per, nch[s] - c0),
                             s * g + tl * TC_M, min(TC_M, g - tl * TC_M), 1 if parts > 1 else 0, 0, 0])
    segd = [[int(X.own_off[s]), X.n_own[s], int(chunk_off[s])] for s in range(N)]
    return _tc_finish(TcPass(), key, int(pbase[-1]) * TC_BANDS, m, kk, val, work, segd, int(chunk_off[-1]),
                      max(nch) if nch else 0, npad, dev)


# ------------------------------------------------------------------------------------------------------
# Pass-1 format over ALL resident cells + per-step work lists for mini-batch row windows (online iNMF)
# ------------------------------------------------------------------------------------------------------
def build_pass1_resident(X, kp, esize):
    """Gene-blocked CSR over every resident cell (replicated datasets of online_iNMF).  Row windows of a
    mini-batch are then just row ranges of the (dataset, gene block) sub-matrices: see window_work."""
    dev = X.rowptr.device
    g, N = X.g, X.N
    cap = tile_cap_rows(kp, esize)
    NB = max(1, -(-g // cap))
    gbs = -(-g // NB)
    rows, cols, vals = [], [], []
    for s in range(N):
        r0, n = int(X.row_base[s]), X.n_local[s]
        if n == 0:
            continue
        rp = X.rowptr[r0:r0 + n + 1]
        p0, p1 = int(rp[0].item()), int(rp[-1].item())
        rows.append(torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64) + r0, rp[1:] - rp[:-1]))
        cols.append(X.colidx[p0:p1])
        vals.append(X.vals[p0:p1])
    row = torch.cat(rows) if rows else torch.zeros(0, dtype=torch.int64, device=dev)
    col = (torch.cat(cols) if cols else torch.zeros(0, dtype=torch.int32, device=dev)).to(torch.int64)
    val = torch.cat(vals) if vals else X.vals[:0]
    base = torch.from_numpy(np.asarray(X.row_base, dtype=np.int64)).to(dev)
    nloc = torch.from_numpy(np.asarray(X.n_local, dtype=np.int64)).to(dev)
    seg = torch.bucketize(row, base[1:], right=True)
    gb = col // gbs
    cidx = NB * base[seg] + gb * nloc[seg] + (row - base[seg])
    order = torch.sort(cidx, stable=True)[1]
    ptr = torch.zeros(NB * X.n_total + 1, dtype=torch.int64, device=dev)
    if cidx.numel():
        ptr[1:] = torch.cumsum(torch.bincount(cidx, minlength=NB * X.n_total), 0)
    bp = BlockedPass()
    bp.split = split_applies(kp, esize)
    bp.pairs = _pack_pairs((col - gb * gbs)[order] * row_bytes(kp, esize), val[order])
    bp.ptr = ptr
    bp.work, bp.nwork, bp.work_host = None, 0, None
    bp.max_tile_rows = min(gbs, g)
    return bp, NB, gbs


def window_work(X, NB, gbs, plan, pieces):
    """Work lists [steps][items][8] (+ valid items per step) for the row windows of ``plan`` (the seg_desc array of
    online_windows): window position q -> row (win_start + q) % n of dataset s lands in output row out_off + q.
    Every window is cut into ``pieces`` row ranges (x 2 where it wraps around the end of the dataset) per gene block;
    built with array operations for all steps at once (it sits on the critical path of online_iNMF's set-up)."""
    g, N = X.g, X.N
    steps = plan.shape[0]
    P = max(1, int(pieces))
    out_off = plan[:, :N, 0].astype(np.int64)                       # [T, N]
    win_start = plan[:, :N, 2].astype(np.int64)
    n = np.maximum(plan[:, :N, 3].astype(np.int64), 1)
    count = plan[:, 1:N + 1, 0].astype(np.int64) - out_off          # [T, N]
    per = np.maximum(-(-count // P), 1)
    j = np.arange(P, dtype=np.int64)[None, None, :]
    q0 = j * per[:, :, None]                                        # [T, N, P]
    q1 = np.minimum(count[:, :, None], q0 + per[:, :, None])
    live = q0 < count[:, :, None]
    lo = (win_start[:, :, None] + q0) % n[:, :, None]
    hi = lo + (q1 - q0)
    nn = n[:, :, None]
    # part 0: [lo, min(hi, n)) ; part 1 (wrapped): [0, hi - n)
    a = np.stack([lo, np.zeros_like(lo)], axis=-1)                  # [T, N, P, 2]
    e = np.stack([np.minimum(hi, nn), np.maximum(hi - nn, 0)], axis=-1)
    ob = np.stack([out_off[:, :, None] + q0 - lo, out_off[:, :, None] + q0 + (nn - lo)], axis=-1)
    valid = (e > a) & live[..., None]
    b = np.arange(NB, dtype=np.int64)
    rows_b = np.minimum(gbs, g - b * gbs)                           # [NB]
    row_base = np.asarray(X.row_base[:N], dtype=np.int64)
    n_local = np.asarray(X.n_local, dtype=np.int64)
    pbase = NB * row_base[:, None] + b[None, :] * n_local[:, None]  # [N, NB]
    drow0 = np.arange(N, dtype=np.int64)[:, None] * g + b[None, :] * gbs
    shape = (steps, N, NB, P, 2)
    work = np.zeros(shape + (WORK_FIELDS,), dtype=np.int64)
    work[..., 0] = pbase[None, :, :, None, None]
    work[..., 1] = drow0[None, :, :, None, None]
    work[..., 2] = rows_b[None, None, :, None, None]
    work[..., 3] = a[:, :, None, :, :]
    work[..., 4] = e[:, :, None, :, :]
    work[..., 5] = ob[:, :, None, :, :]
    work[..., 6] = -1
    ok = np.broadcast_to(valid[:, :, None, :, :] & (rows_b > 0)[None, None, :, None, None], shape)
    items = N * NB * P * 2
    work = work.reshape(steps, items, WORK_FIELDS)
    ok = ok.reshape(steps, items)
    order = np.argsort(~ok, axis=1, kind="stable")                  # valid items first, original order kept
    work = np.take_along_axis(work, order[:, :, None], axis=1)
    counts = ok.sum(axis=1).astype(np.int64)
    pad = np.arange(items)[None, :] >= counts[:, None]
    work[pad] = 0
    work[pad, 6] = -1
    return work, counts


# ------------------------------------------------------------------------------------------------------
# Mini-batch statistics of online_iNMF as a pass 2 over block-aligned windows
# ------------------------------------------------------------------------------------------------------
class _ResidentView(object):
    pass


def resident_view(X):
    """``X`` (a DeviceCSR) with every resident row counted as owned: what build_pass2 needs to block the replicated
    datasets of online_iNMF."""
    v = _ResidentView()
    v.N, v.g = X.N, X.g
    v.rowptr, v.colidx, v.vals = X.rowptr, X.colidx, X.vals
    v.bounds = list(X.bounds)
    v.own = list(X.bounds)
    v.n_local = list(X.n_local)
    v.n_own = list(X.n_local)
    v.row_base = X.row_base
    v.own_off = X.row_base
    v.n_total = X.n_total
    v.n_own_total = X.n_total
    return v


def window_blocks(cnt, kp, esize):
    """Cells per block for windows of ``cnt`` cells: the window size itself when its tile fits in shared memory
    (a window then touches one block when it is aligned and two otherwise), else the smallest equal split that fits."""
    cap = tile_cap_rows(kp, esize)
    out = []
    for c in cnt:
        c = max(1, int(c))
        m = -(-c // cap)
        out.append(-(-c // m))
    return out


def window_stats_work(sell_work_host, blocks, plan, n_global):
    """Per-step work lists of the sliced pass 2 over the cell blocks that the windows of ``plan`` touch.

    ``blocks`` = BlockedPass.blocks of a build_pass2(..., cbs=...) over all resident cells; ``sell_work_host`` the work
    array of its sliced form (one row per (block, gene range)).  A window of dataset i is the rows (win_start + q) % n_i,
    q < count_i: a tail run [start, ...) and, when it wraps, a head run [0, ...).  Every touched block gets a slot of
    cbs[i] rows in the tile buffer (tail blocks first, then head blocks; dataset i owns slots_i consecutive slots from
    row region[i]); window row q goes to row ``place[t, out_off_i + q]`` of the buffer = the position of its cell inside
    its block's slot; the other rows of the slots stay zero and so add nothing.  Built with array operations for all
    steps at once.  Returns (work [T, items, 8], counts [T], place [T, total], region [N + 1]) or None when a window is
    longer than its dataset or wraps onto its own first block (not worth a special case: the row kernels take those)."""
    nb, cbs, blk_off, nch = blocks
    T, N = plan.shape[0], plan.shape[1] - 1
    out_off = plan[:, :, 0].astype(np.int64)
    count = out_off[:, 1:] - out_off[:, :-1]                       # [T, N]
    total = int(out_off[0, N]) if T else 0
    cmax = count.max(axis=0) if T else np.zeros(N, np.int64)
    slots = [(-(-int(cmax[i]) // cbs[i]) + 2) if (nb[i] and cmax[i]) else 0 for i in range(N)]
    region = np.concatenate([[0], np.cumsum([slots[i] * cbs[i] for i in range(N)])]).astype(np.int64)
    place = np.zeros((T, max(total, 1)), dtype=np.int64)
    W3 = sell_work_host[:sum(nb) * nch].reshape(sum(nb), nch, WORK_FIELDS)
    works, oks = [], []
    for i in range(N):
        if slots[i] == 0:
            continue
        n, cb, S = int(n_global[i]), int(cbs[i]), slots[i]
        c = count[:, i]
        if (c > n).any():
            return None
        start = plan[:, i, 2].astype(np.int64) % n
        e1 = np.minimum(start + c, n)
        tail = e1 - start
        head = c - tail
        b0 = start // cb
        ntail = np.where(tail > 0, (e1 - 1) // cb - b0 + 1, 0)
        nhead = np.where(head > 0, (head - 1) // cb + 1, 0)
        if ((head > 0) & (nhead - 1 >= b0)).any() or (ntail + nhead > S).any():
            return None
        q = np.arange(int(cmax[i]), dtype=np.int64)[None, :]
        pos = np.where(q < tail[:, None], (start - b0 * cb)[:, None] + q, (ntail * cb - tail)[:, None] + q) + region[i]
        for t in np.unique(c):  # (all steps of a dataset have the same count in practice)
            sel = np.nonzero(c == t)[0]
            cols = out_off[sel, i][:, None] + np.arange(int(t), dtype=np.int64)[None, :]
            place[sel[:, None], cols] = pos[sel, :int(t)]
        j = np.arange(S, dtype=np.int64)[None, :]
        blk = np.where(j < ntail[:, None], b0[:, None] + j, j - ntail[:, None])       # [T, S]
        ok = (j < (ntail + nhead)[:, None]) & (c > 0)[:, None]
        blk = np.where(ok, blk, 0)
        w = W3[int(blk_off[i]) + blk].copy()                                          # [T, S, nch, 8]
        w[..., 2] = (region[i] + j * cb)[:, :, None]     # dense_row0: the block's slot in the tile buffer
        works.append(w.reshape(T, S * nch, WORK_FIELDS))
        oks.append(np.repeat(ok, nch, axis=1))
    if not works:
        return (np.zeros((T, 1, WORK_FIELDS), dtype=np.int64), np.zeros(T, dtype=np.int64), place, region)
    work = np.concatenate(works, axis=1)
    ok = np.concatenate(oks, axis=1)
    order = np.argsort(~ok, axis=1, kind="stable")
    work = np.take_along_axis(work, order[:, :, None], axis=1)
    counts = ok.sum(axis=1).astype(np.int64)
    keep = int(counts.max()) if T else 0
    work = np.ascontiguousarray(work[:, :max(keep, 1)])
    pad = np.arange(work.shape[1])[None, :] >= counts[:, None]
    work[pad] = 0
    work[pad, 5] = -1
    return work, counts, place, region
