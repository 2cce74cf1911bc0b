"""Device-side state and step functions of the iNMF hot path.

Host orchestration only: torch tensors own the device buffers, every arithmetic step is a call into
libinmf_b200.so (``_lib.call``).  Layouts are the entity-major ones of include/inmf_b200.h:
``Wt``/``Vt[s]`` are genes x k, ``H`` is cells x k over all datasets concatenated.

Multi-GPU: every rank owns a contiguous row slice of EVERY dataset; W, V and the
reduced statistics are replicated; one all-reduce(SUM) of ``St`` (N x g x k) and ``HtH`` (N x k x k)
per ALS iteration / online mini-batch.  H is only gathered for the final write-back.
"""
import numpy as np
import scipy.sparse as sps
import torch

from . import _lib, blocked, devcache, staging

_SUFFIX = {torch.float32: "f32", torch.float64: "f64"}


def is_mixed(dtype):
    """dtype="mixed": fp32 everywhere except the dense operand of the two X-streaming passes, which is rounded to
    bf16 (csrc/sell.cuh, sell32h_pass_kernel).  A named reduced-precision mode, k <= 32."""
    return isinstance(dtype, str) and dtype in ("mixed", "bf16-tile")


def resolve_dtype(dtype):
    if dtype in (torch.float32, torch.float64):
        return dtype
    if is_mixed(dtype):
        return torch.float32
    name = np.dtype(dtype).name if not isinstance(dtype, str) else dtype
    if name in ("float32", "f32", "fp32"):
        return torch.float32
    if name in ("float64", "f64", "fp64"):
        return torch.float64
    raise ValueError("dtype must be float32, float64 or 'mixed', got %r" % (dtype,))


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise _lib.InmfError("pyliger_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    _lib.load()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def shard_bounds(n, rank, world):
    """Contiguous row slice [lo, hi) of a dataset with n cells owned by ``rank``."""
    return (n * rank) // world, (n * (rank + 1)) // world


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _Range(object):
    """NVTX range (visible in nsys / ncu --nvtx timelines): ``with _Range("pass1"): ...``."""
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        torch.cuda.nvtx.range_pop()
        return False


def all_gather_rows(local, counts, group=None):
    """Concatenate per-rank row blocks of unequal height (all_gather needs equal shapes: pad)."""
    cap = max(counts)
    padded = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in counts]
    torch.distributed.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)


def online_windows(total_iters, mbs, sub, n_global, row_base):
    """seg_desc [total_iters, N+1, 4] of the online mini-batches.

    Step t of dataset i covers global rows [t*mb_i, (t+1)*mb_i) mod n_i -- the reference takes its
    h5 chunks in order without shuffling and simply continues into the next epoch
    (_online_iNMF.py:511-552) -- and a rank handles the sub-range ``sub[i]`` = [lo, hi) of it."""
    N = len(mbs)
    plan = np.zeros((total_iters, N + 1, 4), dtype=np.int64)
    sub_n = [hi - lo for lo, hi in sub]
    plan[:, :, 0] = np.concatenate([[0], np.cumsum(sub_n)])[None, :]
    t = np.arange(total_iters, dtype=np.int64)
    for i in range(N):
        n = int(n_global[i])
        plan[:, i, 1] = row_base[i]
        plan[:, i, 2] = (t * int(mbs[i]) + sub[i][0]) % n
        plan[:, i, 3] = n
    return plan


class DeviceMatrix(object):
    """One dataset's scale_data already resident on the device as CSR (rowptr int64[n+1], colidx int32[nnz] sorted
    within rows, vals): accepted wherever a scipy CSR matrix is (InmfEngine, DeviceCSR) and never uploaded."""

    def __init__(self, rowptr, colidx, vals, shape):
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.shape = (int(shape[0]), int(shape[1]))
        self.nnz = int(colidx.numel())
        self._indptr_host = None

    @property
    def indptr(self):
        if self._indptr_host is None:
            self._indptr_host = self.rowptr.cpu().numpy()
        return self._indptr_host


class DeviceCSR(object):
    """All datasets' (local row slices of) scale_data concatenated by rows on the device."""

    def __init__(self, mats, dtype, device, rank=0, world=1, replicate=False, presharded=False):
        self.N = len(mats)
        self.g = int(mats[0].shape[1])
        self.n_global = [int(m.shape[0]) for m in mats]
        self.presharded = bool(presharded)
        # own[i]    = rows (of the matrices passed in) whose H this rank computes
        # bounds[i] = rows resident on this rank (== own, or everything when replicated)
        if presharded:  # every rank was handed only its own cells: take them all
            self.own = [(0, n) for n in self.n_global]
        else:
            self.own = [shard_bounds(n, rank, world) for n in self.n_global]
        self.bounds = [(0, n) for n in self.n_global] if replicate else list(self.own)
        self.n_local = [hi - lo for lo, hi in self.bounds]
        self.n_own = [hi - lo for lo, hi in self.own]
        self.row_base = np.concatenate([[0], np.cumsum(self.n_local)]).astype(np.int64)
        self.own_off = np.concatenate([[0], np.cumsum(self.n_own)]).astype(np.int64)
        self.n_total = int(self.row_base[-1])
        self.n_own_total = int(self.own_off[-1])
        # The index / value arrays go to the device straight from scipy's own buffers (no host-side
        # concatenation or dtype conversion pass); casts happen on the device.
        indptr = [np.zeros(1, dtype=np.int64)]
        pieces = []
        off = 0
        for m, (lo, hi) in zip(mats, self.bounds):
            if m.shape[1] != self.g:
                raise ValueError("all datasets must share the same genes (columns)")
            if isinstance(m, DeviceMatrix):
                twin = (m.rowptr, m.colidx, m.vals)
            else:
                twin = devcache.lookup(m, device)  # left on the device by scale_not_center: no upload
                if not sps.isspmatrix_csr(m):
                    m = sps.csr_matrix(m)
                if not m.has_sorted_indices:
                    m = m.sorted_indices()
            p0, p1 = int(m.indptr[lo]), int(m.indptr[hi])
            indptr.append(m.indptr[lo + 1:hi + 1].astype(np.int64) - p0 + off)
            if twin is not None:
                pieces.append((off, twin[1][p0:p1], twin[2][p0:p1]))
            else:
                pieces.append((off, m.indices[p0:p1], m.data[p0:p1]))
            off += p1 - p0
        self.nnz = off
        self.rowptr = torch.from_numpy(np.concatenate(indptr)).to(device)
        self.colidx = torch.empty(off, dtype=torch.int32, device=device)
        self.vals = torch.empty(off, dtype=dtype, device=device)
        self.h2d_bytes = self.rowptr.numel() * 8
        jobs = []
        for o, idx, val in pieces:
            n = idx.shape[0]
            if n == 0:
                continue
            if torch.is_tensor(idx):  # device twin: device-to-device copies (with casts)
                self.colidx[o:o + n].copy_(idx)
                self.vals[o:o + n].copy_(val)
            elif device.type == "cuda":
                # pinned staging ring; int64 indices / float64 values are narrowed on the host side of the bus
                jobs.append((self.colidx[o:o + n], idx))
                jobs.append((self.vals[o:o + n], val))
            else:  # CPU tensors (host-logic tests)
                self.colidx[o:o + n].copy_(torch.from_numpy(np.ascontiguousarray(idx)))
                self.vals[o:o + n].copy_(torch.from_numpy(np.ascontiguousarray(val)))
        if jobs:
            self.h2d_bytes += staging.upload(device, jobs)


_draw_streams = {}


def launch_als_draw(device, ns, g, k, seed, h_stream=None):
    """Start the seeded draw of (W, V_1..V_N, H_1..H_N) on a side stream; returns (flat double buffer, event).
    ``h_stream``: pre-sharded multi-GPU runs draw their local H from a per-rank stream (see _als_init_host)."""
    device = torch.device(device)
    side = _draw_streams.get(device)
    if side is None:
        side = _draw_streams[device] = torch.cuda.Stream(device)
    n_wv = (len(ns) + 1) * k * g
    n_h = int(sum(ns)) * k
    with torch.cuda.device(device), torch.cuda.stream(side):
        buf = torch.empty(n_wv + n_h, dtype=torch.float64, device=device)
        if h_stream is None:
            _lib.call("inmf_mt19937_uniform", "", int(seed) % (2 ** 32), n_wv + n_h, 2.0, buf, side.cuda_stream)
        else:
            _lib.call("inmf_mt19937_uniform", "", int(seed) % (2 ** 32), n_wv, 2.0, buf, side.cuda_stream)
            _lib.call("inmf_mt19937_uniform", "", (int(seed) + 7919 * (h_stream + 1)) % (2 ** 32), n_h, 2.0,
                      buf[n_wv:], side.cuda_stream)
        done = torch.cuda.Event()
        done.record(side)
    return buf, done


class InmfEngine(object):
    """Buffers + step functions shared by optimize_ALS and online_iNMF."""

    def __init__(self, mats, k, value_lambda, dtype=torch.float64, device=None, group=None,
                 replicate=False, n_prev=0, presharded=False, use_blocked=True, need_stats=True, use_tc=False,
                 tc_three_pass=True, deterministic=False):
        # deterministic: bit-reproducible ALS iterations in fp32 (sliced passes store per-tile partials that are summed
        # in a fixed order; one CTA per reduction in prep / objective) -- slower, opt-in
        self.deterministic = bool(deterministic)
        # use_tc: dense-tile tcgen05 form of the two passes (fp32 only): 3xTF32 (fp32-grade) by default
        self.mixed = is_mixed(dtype)
        if self.mixed and not 1 <= int(k) <= 32:
            raise ValueError("dtype='mixed' supports k <= 32 factors, got %d" % int(k))
        self.use_tc = bool(use_tc) and resolve_dtype(dtype) == torch.float32 and not self.mixed
        self.tc_three_pass = bool(tc_three_pass)
        self.use_blocked = bool(use_blocked)
        self.need_stats = bool(need_stats)
        self.device = require_cuda(device)
        self.dtype = resolve_dtype(dtype)
        self.sfx = _SUFFIX[self.dtype]
        self.group = group
        self.world = 1
        self.rank = 0
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(group)
            self.rank = torch.distributed.get_rank(group)
        self.k = int(k)
        if not 1 <= self.k <= 64:
            raise ValueError("pyliger_b200 supports 1 <= k <= 64 factors, got %d" % self.k)
        self.lam = float(value_lambda)
        self.n_prev = int(n_prev)
        self._chunks = {}
        self._parts = {}
        self._guards = []
        if self.deterministic and (resolve_dtype(dtype) != torch.float32 or is_mixed(dtype)):
            raise ValueError("deterministic=True is implemented for dtype='float32' (the sliced passes)")
        with torch.cuda.device(self.device):
            self.X = DeviceCSR(mats, self.dtype, self.device, self.rank, self.world,
                               replicate=replicate, presharded=presharded)
            self._alloc()

    # ------------------------------------------------------------------ buffers
    # Debug aid (compute-sanitizer is not available on every pool): guard_elems > 0 brackets every buffer the kernels
    # write with bands of a sentinel value; check_guards() reports the buffers whose bands were touched.
    guard_elems = 0
    _GUARD = {torch.float32: 7.25e33, torch.float64: 7.25e33, torch.int64: 0x5A5A5A5A5A5A5A5A, torch.int16: 0x5A5A}

    def _zeros(self, *shape, dtype=None, name=None):
        dtype = self.dtype if dtype is None else dtype
        ge = int(self.guard_elems)
        if not ge:
            return torch.zeros(*shape, dtype=dtype, device=self.device)
        n = int(np.prod(shape))
        raw = torch.full((n + 2 * ge,), self._GUARD[dtype], dtype=dtype, device=self.device)
        raw[ge:ge + n] = 0
        self._guards.append((name or "buffer%d" % len(self._guards), raw, ge, n))
        return raw[ge:ge + n].view(*shape)

    def check_guards(self):
        """Names of the guarded buffers with a modified band (empty list = no out-of-bounds write seen)."""
        bad = []
        for name, raw, ge, n in self._guards:
            ref = torch.full((ge,), self._GUARD[raw.dtype], dtype=raw.dtype, device=raw.device)
            if not (torch.equal(raw[:ge], ref) and torch.equal(raw[ge + n:], ref)):
                bad.append(name)
        return bad

    def _alloc(self):
        X, k, dev, dt = self.X, self.k, self.device, self.dtype
        N, g = X.N, X.g
        self.N, self.g = N, g
        # leading dimension of H / Mt: k rounded up to 4, and at least 32 in fp32 (128-byte tile rows: the wide /
        # sliced kernels of csrc/blocked.cuh, sell.cuh); pad columns stay zero
        self.ldm = max(4, (k + 3) // 4 * 4)
        if dt == torch.float32 and self.use_blocked:
            self.ldm = max(32, self.ldm)
        f64 = torch.float64
        z = self._zeros
        self.Wt = z(g, k, name="Wt")
        # stacked [previous datasets (online scenario 2) | current datasets]
        self.Vt_all = z(self.n_prev + N, g, k, name="Vt")
        self.Vt = self.Vt_all[self.n_prev:]
        self.H = z(max(X.n_own_total, 1), self.ldm, name="H")  # leading dimension ldm: pad columns stay zero
        self.Mt = z(N, g, self.ldm, name="Mt")
        self.G = z(N, k, k, dtype=f64, name="G")
        self.Gwv = z(N, k, k, dtype=f64, name="Gwv")
        self.Gvv = z(N, k, k, dtype=f64, name="Gvv")
        self.St = z(N, g, k, name="St")
        self.HtH = z(N, k, k, dtype=f64, name="HtH")
        self.Rv = z(N, g, k, name="Rv")
        self.Rw = z(g, k, name="Rw")
        self.Gv = z(N, k, k, dtype=f64, name="Gv")
        self.Gw = z(1, k, k, dtype=f64, name="Gw")
        # BPP workspaces (G^-1 + the half-warp solver's deferral lists), one per solve shape
        self.ws_H = _lib.bpp_workspace(max(max(X.n_own) if N else 0, 1), max(N, 1), k, dev)
        self.ws_V = _lib.bpp_workspace(g, max(N, 1), k, dev)
        self.ws_W = _lib.bpp_workspace(g, 1, k, dev)
        self.xnorm2 = z(N, dtype=f64, name="xnorm2")
        self.obj = z(1, dtype=f64, name="obj")
        self.bpp_stats = z(3, 8, dtype=torch.int64, name="bpp_stats")  # rows: H, V, W solves
        self.mask_H = z(max(X.n_own_total, 1), dtype=torch.int64, name="mask_H")
        self.mask_V = z(N * g, dtype=torch.int64, name="mask_V")
        self.mask_W = z(g, dtype=torch.int64, name="mask_W")
        full = np.zeros((N + 1, 4), dtype=np.int64)
        full[:, 0] = X.own_off
        full[:N, 1] = X.row_base[:N]
        full[:N, 2] = [o[0] - b[0] for o, b in zip(X.own, X.bounds)]
        full[:N, 3] = np.maximum(X.n_local, 1)
        self.seg_full = torch.from_numpy(full).to(dev)
        self.seg_full_off = self.seg_full[:, 0].contiguous()
        self.max_rows_full = int(max(X.n_own)) if N else 0
        whole = full.copy()  # every resident row exactly once (sqnorm)
        whole[:N, 2] = 0
        whole[:N, 3] = X.n_local
        self.seg_genes = torch.arange(0, (N + 1) * g, g, dtype=torch.int64, device=dev)
        self.seg_w = torch.tensor([0, g], dtype=torch.int64, device=dev)
        prev = _lib.load().inmf_set_deterministic(1) if self.deterministic else 0
        _lib.call("inmf_sqnorm", self.sfx, X.rowptr, X.vals, torch.from_numpy(whole).to(dev), N, self.xnorm2,
                  _stream())
        if self.deterministic:
            _lib.load().inmf_set_deterministic(prev)
        if X.bounds == X.own and (self.world > 1):
            self._allreduce(self.xnorm2)
        # blocked copies of X for the two streaming passes over whole datasets (csrc/blocked.cuh)
        self.pass1 = self.pass2 = None
        if self.use_blocked and X.n_own_total > 0:
            esize = 4 if dt == torch.float32 else 8
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            if self.mixed:  # bf16 tiles: 64-byte rows, parity-ordered slices of 32 rows
                self.pass1 = blocked.build_sell(blocked.build_pass1(X, 32, 2, sms), 32, height=-32)
                self.Mt16 = z(N * g, 32, dtype=torch.int16, name="Mt16")
                if self.need_stats:
                    self.pass2 = blocked.build_sell(blocked.build_pass2(X, 32, 2, sms), 32, height=-32)
                    self.H16 = z(max(X.n_own_total, 1), 32, dtype=torch.int16, name="H16")
            elif not self.use_tc:
                sell = blocked.sell_applies(self.ldm, esize)
                self.pass1 = blocked.build_pass1(X, self.ldm, esize, sms)
                if sell:
                    self.pass1 = blocked.build_sell(self.pass1, self.ldm)
                if self.need_stats:
                    self.pass2 = blocked.build_pass2(X, self.ldm, esize, sms)
                    if sell:
                        self.pass2 = blocked.build_sell(self.pass2, self.ldm)
        if self.world > 1 and self.use_blocked and self.need_stats and dt == torch.float32 and not self.use_tc:
            self._setup_fused_reduce()
        self.tc1 = self.tc2 = None
        if self.use_tc and X.n_own_total > 0:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            self.tc1 = blocked.build_tc_pass1(X, k)
            per = self.tc1.npad * blocked.TC_KC
            self.dc1_hi = torch.empty(self.tc1.total_chunks * per, dtype=dt, device=dev)
            self.dc1_lo = torch.empty_like(self.dc1_hi)
            if self.need_stats:
                self.tc2 = blocked.build_tc_pass2(X, k, sms)
                self.dc2_hi = torch.empty(self.tc2.total_chunks * per, dtype=dt, device=dev)
                self.dc2_lo = torch.empty_like(self.dc2_hi)

    coalesce_allreduce = True

    # Multi-GPU: pass 2 fused with the all-reduce of its outputs (inmf_sell_pass_mc_f32).  The sums over the ranks,
    # self.St / self.HtH, live in symmetric memory bound to a multicast object of the NVLink switch; the pass accumulates
    # this rank's partial sums in St_part / HtH_part and the CTA that completes a (dataset, gene range) adds that range
    # into every rank's copy with multimem.red while the rest of the pass is still running.  No NCCL collective
    # follows.  Opt-in: measured on 2 GPUs it only matches pass 2 + NCCL all-reduce (C2 0.870 vs 0.861 ms per iteration:
    # two cross-rank barriers and the extra clear cost what the collective's launch did), and with more than two ranks the
    # switch applies the ranks' adds to each copy in arrival order, so the copies can differ in the last bit -- the
    # objective is then taken from rank 0 (one 8-byte broadcast) to keep the ranks' stopping decisions identical.
    fused_reduce = False
    mc = None
    fused_reduce_error = None

    def _setup_fused_reduce(self):
        """Collective over the ranks (every rank of a multi-GPU engine with blocked statistics passes calls it)."""
        self.mc = None
        N, g, k, dev = self.N, self.g, self.k, self.device
        p = self.pass2
        eligible = int(self.fused_reduce and isinstance(p, blocked.SellPass) and p.height in (8, 32, 33)
                       and not self.deterministic and not self.mixed and self.guard_elems == 0)
        flag = torch.tensor([eligible], dtype=torch.int32, device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            return
        ok = 0
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else torch.distributed.group.WORLD
            n_st = (N * g * k * 4 + 255) // 256 * 256
            buf = symm.empty(n_st + N * k * k * 8, dtype=torch.uint8, device=dev)
            hdl = symm.rendezvous(buf, group)
            mcp = int(getattr(hdl, "multicast_ptr", 0) or 0)
            if mcp:
                buf.zero_()
                ok = 1
            else:
                self.fused_reduce_error = "no multicast support"
        except Exception as e:  # symmetric memory not available in this process / on this fabric
            self.fused_reduce_error = repr(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:  # all ranks or none
            return
        part = torch.zeros(n_st + N * k * k * 8, dtype=torch.uint8, device=dev)  # one clear for both partial buffers
        self.St_part = part[:N * g * k * 4].view(torch.float32).view(N, g, k)
        self.HtH_part = part[n_st:n_st + N * k * k * 8].view(torch.float64).view(N, k, k)
        self.St = buf[:N * g * k * 4].view(torch.float32).view(N, g, k)
        self.HtH = buf[n_st:n_st + N * k * k * 8].view(torch.float64).view(N, k, k)
        W, expected = blocked.mc_work(p, g)
        work_mc = torch.from_numpy(W).to(dev)
        expected_t = torch.from_numpy(expected).to(dev)
        counters = torch.zeros(len(expected), dtype=torch.int32, device=dev)
        args = torch.tensor([mcp, mcp + n_st, counters.data_ptr(), expected_t.data_ptr()], dtype=torch.int64, device=dev)
        self.mc = (hdl, buf, work_mc, args, part, (expected_t, counters))

    def _allreduce(self, *tensors):
        """SUM over the ranks.  Tensors of one dtype go out as ONE coalesced NCCL launch; different dtypes (fp32 St +
        fp64 H'H) as one launch per dtype (c10d's coalescing of mixed dtypes is not relied upon)."""
        if self.world <= 1:
            return
        by_dtype = {}
        for t in tensors:
            by_dtype.setdefault(t.dtype, []).append(t)
        for group_t in by_dtype.values():
            if len(group_t) > 1 and self.coalesce_allreduce:
                try:
                    from torch.distributed.distributed_c10d import _coalescing_manager
                    with _coalescing_manager(group=self.group, device=self.device, async_ops=False):
                        for t in group_t:
                            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=self.group)
                    continue
                except (ImportError, TypeError, RuntimeError, NotImplementedError):  # backend without coalescing
                    pass
            for t in group_t:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=self.group)

    def reduce_stats(self):
        """All-reduce of the statistics of the last stats(reduce=False)."""
        self._allreduce(self.St, self.HtH)

    # ------------------------------------------------------------------ kernels
    def prep(self, with_V=True):
        prev = _lib.load().inmf_set_deterministic(1) if self.deterministic else 0
        try:
            _lib.call("inmf_prep", self.sfx, self.Wt, self.Vt if with_V else None, self.Mt, self.ldm, self.G,
                      self.Gwv, self.Gvv, self.lam, self.N, self.g, self.k, _stream())
        finally:
            if self.deterministic:
                _lib.load().inmf_set_deterministic(prev)

    def rhs_H(self, seg_desc=None, max_rows=None, out=None, nseg=None):
        """out[q, :k] = X[row(q), :] (W+V)  for the windows in seg_desc (default: all owned cells)."""
        full = seg_desc is None
        seg_desc = self.seg_full if full else seg_desc
        max_rows = self.max_rows_full if max_rows is None else max_rows
        out = self.H if out is None else out
        nseg = self.N if nseg is None else nseg
        if max_rows <= 0:
            return
        X, k = self.X, self.k
        if full and out is self.H and self.tc1 is not None:
            t = self.tc1
            _lib.call("inmf_tc_relayout_f32", "", self.Mt, self.ldm, t.segd, t.nseg, t.max_seg_chunks, self.dc1_hi,
                      self.dc1_lo, k, t.npad, _stream())
            _lib.call("inmf_tc_pass_f32", "", t.work, t.nwork, t.ptr, t.pairs, self.dc1_hi, self.dc1_lo, out, self.ldm,
                      k, t.npad, 1 if self.tc_three_pass else 0, 0, _stream())
        elif full and self.pass1 is not None:
            if out.data_ptr() != self._clean_ptr:  # not cleared by the previous H solve (bpp_H(clear=...)): do it now
                out.zero_()
            else:
                self._clean_ptr = None
            p = self.pass1
            self._blocked(p, p.work, p.nwork, self.Mt, out, self.ldm, None)
        else:
            _lib.call("inmf_spmm", self.sfx, X.rowptr, X.colidx, X.vals, self.Mt, self.ldm, self.g, out, self.ldm,
                      seg_desc, nseg, max_rows, k, _stream())

    def _blocked(self, p, work, nwork, dense, out, ldo, gram):
        """One tiled pass (out += X_blocked * dense tile): sliced, plain or split-tile kernel, as the format was built."""
        if isinstance(p, blocked.SellPass) and p.height == -32:  # mixed precision: the operand goes through bf16
            rows = dense.numel() // self.ldm
            buf = self.Mt16 if p is self.pass1 else self.H16
            if gram is not None:  # pass 2: H -> bf16 and H'H of the fp32 H (not of the rounded tile) in one pass
                _lib.call("inmf_to_bf16_gram", "", dense, self.ldm, self.seg_full_off, self.N, self.max_rows_full, self.k,
                          buf, gram, _stream())
            else:
                _lib.call("inmf_to_bf16_rows", "", dense, self.ldm, rows, self.k, buf, _stream())
            _lib.call("inmf_sell_pass_mixed", "", work, nwork, p.slice_ptr, p.rowid, p.pairs4, buf, out, ldo, self.k,
                      p.max_tile_rows, _stream())
            return
        if isinstance(p, blocked.SellPass):
            if self.deterministic and (p.nparts > 2 or gram is not None):
                # ordered sums: every work item stores into its own partial buffer, then part[0] + part[1] + ...
                n = out.numel()
                key = (id(p), n)
                if key not in self._parts:
                    self._parts[key] = (torch.zeros(p.nparts, n, dtype=out.dtype, device=out.device),
                                        None if gram is None else
                                        torch.zeros(p.nparts, gram.numel(), dtype=gram.dtype, device=gram.device))
                part, gpart = self._parts[key]
                if gpart is not None:
                    gpart.zero_()  # (the Gram partials are accumulated by their owner threads)
                _lib.call("inmf_sell_pass_f32", "", work, nwork, p.slice_ptr, p.rowid, p.pairs4, p.height, dense,
                          self.ldm, part, ldo, self.k, p.max_tile_rows, gpart, n, 0 if gram is None else gram.numel(),
                          _stream())
                _lib.call("inmf_reduce_parts", "f32", part, p.nparts, n, n, out, _stream())
                if gram is not None:
                    _lib.call("inmf_reduce_parts", "f64", gpart, p.nparts, gram.numel(), gram.numel(), gram, _stream())
                return
            _lib.call("inmf_sell_pass_f32", "", work, nwork, p.slice_ptr, p.rowid, p.pairs4, p.height, dense, self.ldm, out, ldo,
                      self.k, p.max_tile_rows, gram, 0, 0, _stream())
            return
        name, sfx = ("inmf_blocked_split_f32", "") if p.split else ("inmf_blocked_spmm", self.sfx)
        if blocked.WIDE_TILE and not p.split and self.sfx == "f32" and self.ldm == 32:
            name, sfx = "inmf_blocked_wide_f32", ""  # same formats, 8 factors per lane
        _lib.call(name, sfx, work, nwork, p.ptr, p.pairs, dense, self.ldm, out, ldo, self.k, p.max_tile_rows, gram,
                  _stream())

    # V / W solves: 2 = complement preference as for H.  Measured slower (C2: V 0.081 -> 0.116 ms, W 0.048 ->
    # 0.081 ms): gene loadings are sparse, so their passive sets are the small side already.
    vw_flags = 0

    def _vw(self, warm):
        """flags of the gene-side solves: warm-started ones at k <= 31 never need the inverse (|P| <= 31 is solved
        directly; the all-passive shortcut only pays on cold starts)."""
        return (1 if warm else 0) | self.vw_flags | (16 if (warm and self.k <= 31 and not self.vw_flags) else 0)

    _clean_ptr = None  # data_ptr of the H-shaped buffer known to be all zero (cleared by the last bpp_H(clear=...))

    def bpp_H(self, seg_desc=None, max_rows=None, out=None, mask=None, warm=False, nseg=None, clear=None):
        """In-place batched BPP on the right-hand sides produced by rhs_H.  ``clear``: an H-shaped buffer whose
        solved columns are zeroed by the same kernel (the spare buffer the next pass 1 accumulates into)."""
        if seg_desc is None:
            seg_off = self.seg_full_off
        else:
            seg_off = seg_desc[:, 0].contiguous()  # column ranges = the out_off column of seg_desc
        max_rows = self.max_rows_full if max_rows is None else max_rows
        out = self.H if out is None else out
        nseg = self.N if nseg is None else nseg
        if max_rows <= 0:
            return
        k = self.k
        # bit 1: G = (W+V)(W+V)' + lambda VV' is well conditioned and H is mostly passive -> complement solves
        ws = self.ws_H
        if max_rows > self.max_rows_full or nseg > self.N:  # windows larger than the sized workspace (never in practice)
            ws = _lib.bpp_workspace(max_rows, nseg, k, self.device)
        _lib.call("inmf_bpp", self.sfx, self.G, out, self.ldm, out, self.ldm, clear, self.ldm if clear is not None else 0,
                  mask, (1 if warm else 0) | 2 | (4 if clear is not None else 0), seg_off, nseg, max_rows, k,
                  self.bpp_stats[0], *_lib.ws_args(ws), _stream())
        if clear is not None:
            self._clean_ptr = clear.data_ptr()

    def solve_H(self, seg_desc=None, max_rows=None, out=None, mask=None, warm=False, nseg=None):
        """out[q,:] = argmin_{h>=0} for every row of the windows in seg_desc (default: all owned cells)."""
        self.rhs_H(seg_desc, max_rows, out, nseg)
        self.bpp_H(seg_desc, max_rows, out, mask, warm, nseg)

    def stats(self, seg_desc=None, max_rows=None, H=None, St=None, HtH=None, reduce=True, nseg=None):
        full = seg_desc is None and H is None
        seg_desc = self.seg_full if seg_desc is None else seg_desc
        max_rows = self.max_rows_full if max_rows is None else max_rows
        H = self.H if H is None else H
        St = self.St if St is None else St
        HtH = self.HtH if HtH is None else HtH
        nseg = self.N if nseg is None else nseg
        X, k = self.X, self.k
        if full and reduce and self.mc is not None and St is self.St and HtH is self.HtH:
            hdl, buf, work_mc, args, part = self.mc[:5]
            p = self.pass2
            part.zero_()
            buf.zero_()
            hdl.barrier(channel=0)  # every rank's sum buffer is zero before the first add arrives
            _lib.call("inmf_sell_pass_mc_f32", "", work_mc, p.nwork, p.slice_ptr, p.rowid, p.pairs4, p.height, H,
                      self.ldm, self.St_part, k, p.max_tile_rows, self.HtH_part, args, _stream())
            hdl.barrier(channel=1)  # every rank's adds have landed
            return
        St.zero_()
        HtH.zero_()
        if max_rows > 0:
            if full and self.tc2 is not None:
                t = self.tc2
                _lib.call("inmf_tc_relayout_f32", "", H, self.ldm, t.segd, t.nseg, t.max_seg_chunks, self.dc2_hi,
                          self.dc2_lo, k, t.npad, _stream())
                _lib.call("inmf_tc_pass_f32", "", t.work, t.nwork, t.ptr, t.pairs, self.dc2_hi, self.dc2_lo, St, k, k,
                          t.npad, 1 if self.tc_three_pass else 0, 0, _stream())
                _lib.call("inmf_gram_rows", self.sfx, H, self.ldm, self.seg_full, nseg, max_rows, HtH, k, _stream())
            elif full and self.pass2 is not None:
                p = self.pass2
                self._blocked(p, p.work, p.nwork, H, St, k, HtH)
            else:
                _lib.call("inmf_stats", self.sfx, X.rowptr, X.colidx, X.vals, H, self.ldm, St, HtH, self.g, seg_desc,
                          nseg, max_rows, k, _stream())
        if reduce:
            self._allreduce(St, HtH)

    def hals_H(self):
        """One HALS sweep over H in place (iNMF_HALS, _utilities.py:132-148): right-hand sides by pass 1 into a scratch
        buffer, then the per-cell Gauss-Seidel sweep.  Requires prep() of the current W, V."""
        if getattr(self, "R_hals", None) is None:
            self.R_hals = self._zeros(*self.H.shape, name="R_hals")
        self.rhs_H(out=self.R_hals)
        if self.max_rows_full > 0:
            _lib.call("inmf_hals_h", self.sfx, self.G, self.R_hals, self.ldm, self.H, self.ldm, self.seg_full_off, self.N,
                      self.max_rows_full, self.k, _stream())

    def hals_WV(self, n_inner=1):
        """HALS sweeps of W then V (_utilities.py:102-129) with A = H'H, B = X'H of the current statistics."""
        _lib.call("inmf_hals", self.sfx, self.HtH, self.St, self.Wt, self.Vt_all, self.lam, self.n_prev,
                  self.n_prev + self.N, self.g, self.k, int(n_inner), _stream())

    # Multi-GPU: the gene-side solves are SHARDED over the ranks (every rank solves a contiguous 1/world chunk of the
    # N*g columns of V, and of the g columns of W, and the chunks are all-gathered) instead of being repeated on every
    # rank: at C4 on 8 GPUs the replicated solves were 0.43 ms of a 4.6 ms iteration.  Same kernels per column, so the
    # iterates do not depend on the number of ranks.
    shard_solves = True
    shard_min_columns = 32768  # smaller solves are latency-bound: a shard takes as long as the whole + the gather

    def _solve_chunk(self, total, per_seg):
        """This rank's chunk [c0, c1) of ``total`` columns organised in segments of ``per_seg`` (None if the solves are
        not sharded): (c0, c1, first segment, device seg_off of the touched partial segments)."""
        if self.world == 1 or not self.shard_solves or total < self.shard_min_columns:
            return None
        key = (total, per_seg)
        if key not in self._chunks:
            per = -(-total // self.world)
            c0, c1 = min(total, self.rank * per), min(total, (self.rank + 1) * per)
            s0 = c0 // per_seg
            s1 = max(s0, -(-c1 // per_seg))
            offs = [c0] + [min(c1, (s + 1) * per_seg) for s in range(s0, s1)]
            self._chunks[key] = (c0, c1, s0, torch.tensor(offs, dtype=torch.int64, device=self.device), per)
        return self._chunks[key]

    def _gather_chunks(self, flat, chunk, total):
        """All ranks' solved chunks of ``flat`` ([total, k]) -> the full array on every rank."""
        c0, c1, _, _, per = chunk
        k = self.k
        if per * self.world == total:  # in-place all-gather (input = output + rank * count)
            torch.distributed.all_gather_into_tensor(flat.view(-1), flat[c0:c1].reshape(-1), group=self.group)
            return
        pad = torch.zeros(per * k, dtype=flat.dtype, device=flat.device)
        pad[:(c1 - c0) * k] = flat[c0:c1].reshape(-1)
        out = torch.empty(self.world * per * k, dtype=flat.dtype, device=flat.device)
        torch.distributed.all_gather_into_tensor(out, pad, group=self.group)
        flat.view(-1).copy_(out[:total * k])

    def solve_V(self, warm=False):
        k, g, N = self.k, self.g, self.N
        _lib.call("inmf_rhs_v", self.sfx, self.St, self.Wt, self.HtH, self.Rv, self.Gv, self.lam, N, g, k,
                  _stream())
        chunk = self._solve_chunk(N * g, g)
        if chunk is None:
            _lib.call("inmf_bpp", self.sfx, self.Gv, self.Rv, k, self.Vt, k, None, 0, self.mask_V,
                      self._vw(warm), self.seg_genes, N, g, k, self.bpp_stats[1],
                      *_lib.ws_args(self.ws_V), _stream())
            return
        c0, c1, s0, offs, _ = chunk
        if c1 > c0:
            _lib.call("inmf_bpp", self.sfx, self.Gv[s0:], self.Rv, k, self.Vt, k, None, 0, self.mask_V,
                      self._vw(warm), offs, offs.numel() - 1, g, k, self.bpp_stats[1],
                      *_lib.ws_args(self.ws_V), _stream())
        self._gather_chunks(self.Vt.view(N * g, k), chunk, N * g)

    def solve_W(self, warm=False):
        k, g, N = self.k, self.g, self.N
        _lib.call("inmf_rhs_w", self.sfx, self.St, self.Vt, self.HtH, self.Rw, self.Gw, N, g, k, _stream())
        chunk = self._solve_chunk(g, g)
        if chunk is None:
            _lib.call("inmf_bpp", self.sfx, self.Gw, self.Rw, k, self.Wt, k, None, 0, self.mask_W,
                      self._vw(warm), self.seg_w, 1, g, k, self.bpp_stats[2],
                      *_lib.ws_args(self.ws_W), _stream())
            return
        c0, c1, _, offs, _ = chunk
        if c1 > c0:
            _lib.call("inmf_bpp", self.sfx, self.Gw, self.Rw, k, self.Wt, k, None, 0, self.mask_W,
                      self._vw(warm), offs, 1, g, k, self.bpp_stats[2], *_lib.ws_args(self.ws_W), _stream())
        self._gather_chunks(self.Wt, chunk, g)

    def objective(self):
        """Objective from the current St/HtH and the Mt/Gwv/Gvv of the last prep(). Device scalar."""
        prev = _lib.load().inmf_set_deterministic(1) if self.deterministic else 0
        try:
            _lib.call("inmf_objective", self.sfx, self.St, self.Mt, self.ldm, self.HtH, self.Gwv, self.Gvv,
                      self.xnorm2, self.lam, self.N, self.g, self.k, self.obj, _stream())
        finally:
            if self.deterministic:
                _lib.load().inmf_set_deterministic(prev)
        if self.mc is not None and self.world > 2:  # the ranks' copies of the sums may differ in the last bit
            torch.distributed.broadcast(self.obj, src=torch.distributed.get_global_rank(self.group, 0)
                                        if self.group is not None else 0, group=self.group)
        return self.obj

    # ------------------------------------------------------------------ ALS
    def set_als_state(self, W, V, H=None):
        """W: k x g, V: list of k x g, H: list of n_i x k (GLOBAL rows; the local slice is taken)."""
        dev, dt = self.device, self.dtype
        self.h2d_state_bytes = 8 * (np.size(W) + sum(np.size(v) for v in V))
        self.Wt.copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(W, dtype=np.float64).T)).to(dev, dt))
        for i, Vi in enumerate(V):
            self.Vt[i].copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(Vi, dtype=np.float64).T)).to(dev, dt))
        if H is not None:
            for i, Hi in enumerate(H):
                lo, hi = self.X.own[i]
                b0, b1 = int(self.X.own_off[i]), int(self.X.own_off[i + 1])
                if b1 > b0:
                    self.h2d_state_bytes += 8 * (b1 - b0) * self.k
                    self.H[b0:b1, :self.k].copy_(torch.from_numpy(np.ascontiguousarray(
                        np.asarray(Hi, dtype=np.float64)[lo:hi])).to(dev, dt))

    def draw_als_state(self, seed, h_stream=None, pending=None):
        """Seeded W, V_i, H_i in the reference's draw order (_iNMF_ANLS.py:103-108: W, then every V_i, then every
        H_i, all abs(uniform(0, 2)) from numpy's global RNG) generated on the device -- the same bits as
        ``_als_init_host`` without the host RNG pass and the upload.  ``pending``: a draw started earlier with
        ``launch_als_draw`` (so that it overlapped the upload of X)."""
        k, g, N = self.k, self.g, self.N
        ns = self.X.n_global
        buf, done = pending if pending is not None else launch_als_draw(self.device, ns, g, k, seed, h_stream)
        torch.cuda.current_stream(self.device).wait_event(done)
        n_wv = (N + 1) * k * g
        self.Wt.copy_(buf[:k * g].view(k, g).t())
        for i in range(N):
            self.Vt[i].copy_(buf[(i + 1) * k * g:(i + 2) * k * g].view(k, g).t())
        off = n_wv
        for i, n in enumerate(ns):
            lo, hi = self.X.own[i]
            b0, b1 = int(self.X.own_off[i]), int(self.X.own_off[i + 1])
            if b1 > b0:
                self.H[b0:b1, :k].copy_(buf[off + lo * k:off + hi * k].view(hi - lo, k))
            off += n * k
        buf.record_stream(torch.cuda.current_stream(self.device))
        self.h2d_state_bytes = 0

    def initial_objective(self):
        """obj0 of _iNMF_ANLS.py:122-129 from the H currently in self.H (one statistics pass)."""
        self.prep()
        self.stats()
        return float(self.objective().item())

    def als_iteration(self, warm=False):
        """One (H, V, W) block update + objective; returns the device scalar objective.

        Requires prep() of the current W, V (done by initial_objective() / the previous iteration)."""
        if not self._begun:
            self.als_begin()
        self._begun = False
        # the buffer that held the previous iteration's H is dead from here on: the solve clears it for the next pass 1
        with _Range("inmf.bpp_H"):
            self.bpp_H(mask=self.mask_H, warm=warm, clear=self.H_alt)
        with _Range("inmf.pass2_stats"):
            self.stats()
        with _Range("inmf.solve_V"):
            self.solve_V(warm=warm)
        with _Range("inmf.solve_W"):
            self.solve_W(warm=warm)
        with _Range("inmf.prep_objective"):
            self.prep()
            return self.objective()

    # The reference decides after every iteration whether to go on (delta > thresh), which costs one host read-back
    # of the objective per iteration.  To keep the GPU busy across that read-back, the first kernel of the NEXT
    # iteration (pass 1, which only reads X and the already prepared W + V_i) is enqueued before the host waits; it
    # writes into a spare H buffer, so stopping just drops it.
    _begun = False
    H_alt = None
    _obj_host = None
    _obj_event = None

    def als_begin(self):
        """Enqueue pass 1 of the next iteration into the spare H buffer (the current H stays intact in H_alt).
        The two buffers ping-pong: pass 1 accumulates into the one the previous H solve cleared."""
        if self._begun:
            return
        if self.H_alt is None:
            self.H_alt = self._zeros(*self.H.shape, name="H_alt")
            self._clean_ptr = self.H_alt.data_ptr()
        self.H, self.H_alt = self.H_alt, self.H
        with _Range("inmf.pass1_rhs_H"):
            self.rhs_H()
        self._begun = True

    def als_cancel(self):
        """Drop a pass 1 started by als_begin(): self.H is the last completed iteration's H again."""
        if self._begun:
            self.H, self.H_alt = self.H_alt, self.H
            self._begun = False

    # ------------------------------------------------------------------ CUDA-graphed iterations
    # A warm-started iteration is the same ~20 launches (kernels, memsets, NCCL collectives) every time; only the two
    # ping-pong H buffers alternate.  After a few eager iterations it is captured as two CUDA graphs per buffer parity
    # -- "body" (H solve ... objective + its copy to pinned memory) and "pass 1" (right-hand sides into the other
    # buffer) -- and replayed: one launch call instead of ~20 per iteration and no gaps between dependent kernels.
    # At C4 on 8 GPUs the eager loop lost 0.48 ms of a 2.87 ms iteration to launch gaps and rank skew.
    # use_graph: None = policy (on when the cells are sharded over several GPUs, where the replay also removes the
    # launch skew between the ranks; off on one GPU, where the host keeps ahead of the GPU anyway and a 20 - 100
    # iteration call does not earn back the capture: measured 38.9 vs 40.9 ms for the 20-iteration public call at C2,
    # 100.3 vs 102 - 116 ms for 100 iterations), True = from the first warm iteration on, False = never.
    use_graph = None
    _graph_eager_warm = 0       # eager warm iterations seen (everything lazily allocated exists after the first)
    _graphs = None
    _capture_stream = None

    def _graphs_enabled(self):
        return (self.world > 1) if self.use_graph is None else bool(self.use_graph)

    def _graph_capture(self, key, fn):
        """Capture fn() (nothing runs); None if this stack cannot capture it -- the caller then runs fn() eagerly and
        graphs stay off (the results are the same either way).  Captured directly on a side stream: the
        torch.cuda.graph() context would also empty the caching allocator, which costs the NEXT call its cudaMallocs."""
        g = torch.cuda.CUDAGraph()
        if self._capture_stream is None:
            self._capture_stream = torch.cuda.Stream(self.device)
        side, cur = self._capture_stream, torch.cuda.current_stream(self.device)
        n0 = _lib.launch_count()
        try:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                g.capture_begin()
                try:
                    fn()
                finally:
                    g.capture_end()
            cur.wait_stream(side)
        except Exception as e:
            import warnings
            warnings.warn("pyliger_b200: CUDA-graph capture of the ALS iteration failed (%s); running eagerly" % (e,))
            self.use_graph = False
            torch.cuda.synchronize(self.device)
            return None
        n = _lib.launch_count() - n0   # kernels of the library inside this graph (counted at capture, run at replay)
        _lib.load().inmf_note_graph_launches(-n)
        self._graphs[key] = (g, n)
        return self._graphs[key]

    @staticmethod
    def _run(g, fn):
        if g is None:
            fn()
        else:
            g[0].replay()
            _lib.load().inmf_note_graph_launches(g[1])

    def _fn_pass1(self):
        self._clean_ptr = self.H.data_ptr()
        self.rhs_H()

    def _fn_body(self):
        self.bpp_H(mask=self.mask_H, warm=True, clear=self.H_alt)
        self.stats()
        self.solve_V(warm=True)
        self.solve_W(warm=True)
        self.prep()
        self._obj_host.copy_(self.objective().reshape(1), non_blocking=True)

    def _graph_capture_all(self):
        """Capture the body and pass-1 graphs of BOTH buffer parities at once (capturing runs nothing, so the buffers
        can be swapped on the host side for the second pair), so that no capture lands in a later iteration."""
        saved = (self.H, self.H_alt, self._clean_ptr)
        for _ in range(2):
            if self._graphs_enabled() and ("body", self.H.data_ptr()) not in self._graphs:
                self._graph_capture(("body", self.H.data_ptr()), self._fn_body)
            if self._graphs_enabled() and ("pass1", self.H.data_ptr()) not in self._graphs:
                self._graph_capture(("pass1", self.H.data_ptr()), self._fn_pass1)
            self.H, self.H_alt = self.H_alt, self.H
        self.H, self.H_alt, self._clean_ptr = saved

    def _graph_pass1(self):
        """als_begin() through the pass-1 graph of the buffer that becomes current."""
        if self._begun:
            return
        self.H, self.H_alt = self.H_alt, self.H
        if self.H.data_ptr() != self._clean_ptr:
            self.H.zero_()
        self._clean_ptr = self.H.data_ptr()
        self._run(self._graphs.get(("pass1", self.H.data_ptr())), self._fn_pass1)
        self._clean_ptr = None
        self._begun = True

    def _graph_iteration(self, speculate):
        if self._obj_host is None:
            self._obj_host = torch.empty(1, dtype=torch.float64, pin_memory=True)
            self._obj_event = torch.cuda.Event()
        if not self._graphs:
            self._graph_capture_all()
        if not self._begun:
            self._graph_pass1()
        self._begun = False
        self._run(self._graphs.get(("body", self.H.data_ptr())), self._fn_body)
        self._clean_ptr = self.H_alt.data_ptr()  # cleared by the H solve (inside the graph)
        self._obj_event.record()
        if speculate:
            self._graph_pass1()
        self._obj_event.synchronize()
        return float(self._obj_host[0])

    def als_iteration_value(self, warm=False, speculate=False):
        """als_iteration() + host value of the objective; with ``speculate`` the next pass 1 overlaps the read-back
        (call als_cancel() if no further iteration follows)."""
        if (self._graphs_enabled() and warm and self._graph_eager_warm >= 1 and self.pass1 is not None and self.pass2 is not None
                and self.H_alt is not None and not self.use_tc):
            if self._graphs is None:
                self._graphs = {}
            return self._graph_iteration(speculate)
        if warm:
            self._graph_eager_warm += 1
        obj = self.als_iteration(warm=warm)
        if not speculate:
            return float(obj.item())
        # the objective goes to pinned memory behind an event, so that the host does not wait for the pass 1
        # enqueued after it
        if self._obj_host is None:
            self._obj_host = torch.empty(1, dtype=obj.dtype, pin_memory=True)
            self._obj_event = torch.cuda.Event()
        self._obj_host.copy_(obj.reshape(1), non_blocking=True)
        self._obj_event.record()
        self.als_begin()
        self._obj_event.synchronize()
        return float(self._obj_host[0])

    def gather_H(self):
        """Full H per dataset as host float64 arrays (all-gather of the row slices)."""
        out = []
        for i in range(self.N):
            b0, b1 = int(self.X.own_off[i]), int(self.X.own_off[i + 1])
            local = self.H[b0:b1, :self.k].to(torch.float64).contiguous()
            if self.world == 1 or self.X.presharded:
                out.append(local.cpu().numpy())
                continue
            sizes = [shard_bounds(self.X.n_global[i], r, self.world) for r in range(self.world)]
            out.append(all_gather_rows(local, [hi - lo for lo, hi in sizes], self.group).cpu().numpy())
        return out

    def device_H_views(self, clone=False):
        """Per dataset, the device rows of H that host_factors() downloads (None when the rows of a dataset are spread
        over several ranks): registered as twins of the host arrays so that consumers of H skip the upload."""
        if self.world > 1 and not self.X.presharded:
            return None
        out = []
        for i in range(self.N):
            v = self.H[int(self.X.own_off[i]):int(self.X.own_off[i + 1]), :self.k]
            out.append(v.clone() if clone else v)
        return out

    def host_W(self):
        return self.Wt.to(torch.float64).cpu().numpy()

    def host_V(self):
        return [self.Vt[i].to(torch.float64).cpu().numpy() for i in range(self.N)]

    def host_factors(self):
        """(H list, W, V list) as host float64 arrays -- one staged download when no gather is needed."""
        if self.world > 1 and not self.X.presharded:
            return self.gather_H(), self.host_W(), self.host_V()
        hs = [self.H[int(self.X.own_off[i]):int(self.X.own_off[i + 1]), :self.k] for i in range(self.N)]
        outs = staging.to_host_f64(self.device, hs + [self.Wt] + [self.Vt[i] for i in range(self.N)])
        return outs[:self.N], outs[self.N], outs[self.N + 1:]

    def bpp_report(self):
        s = self.bpp_stats.cpu().numpy()
        names = ("not_converged", "factorizations", "backup_flips", "pivot_failures", "rounds_sum",
                 "rounds_max", "columns", "deferred")
        return {which: dict(zip(names, map(int, s[r]))) for r, which in enumerate(("H", "V", "W"))}

    # ------------------------------------------------------------------ online iNMF
    def alloc_online(self, miniBatch_sizes):
        """Buffers for the mini-batch loop (_online_iNMF.py:367-378): A/B statistics (+ the copies of
        information older than two epochs) stacked as [previous | current] datasets."""
        N, g, k, dev, dt = self.N, self.g, self.k, self.device, self.dtype
        nall = self.n_prev + N
        f64 = torch.float64
        z = self._zeros
        self.A_all = z(nall, k, k, dtype=f64, name="A")
        self.B_all = z(nall, g, k, name="B")
        self.A = self.A_all[self.n_prev:]
        self.B = self.B_all[self.n_prev:]
        self.A_old = z(N, k, k, dtype=f64, name="A_old")
        self.B_old = z(N, g, k, name="B_old")
        self.HHt = z(N, k, k, dtype=f64, name="HHt")
        self.XHt = z(N, g, k, name="XHt")
        self.mbs = [int(m) for m in miniBatch_sizes]
        self.sub = [shard_bounds(m, self.rank, self.world) for m in self.mbs]  # this rank's part of a batch
        self.max_sub = max(hi - lo for lo, hi in self.sub)
        self.H_mb = z(max(sum(hi - lo for lo, hi in self.sub), 1), self.ldm, name="H_mb")
        self.win = None  # blocked pass-1 copy over all resident cells for the mini-batch windows
        self.win2 = None  # sliced pass-2 copy in window-sized cell blocks for the mini-batch statistics
        self.inv_mb = torch.tensor([1.0 / m for m in self.mbs], dtype=f64, device=dev)

    def online_plan(self, total_iters):
        """Device seg_desc of every step (see ``online_windows``). Needs replicate=True.
        Also builds the per-step work lists of the tiled (blocked) right-hand-side kernel."""
        plan = online_windows(total_iters, self.mbs, self.sub, self.X.n_global, self.X.row_base)
        if self.use_blocked and total_iters > 0:
            esize = 4 if self.dtype == torch.float32 else 8
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            bp, NB, gbs = blocked.build_pass1_resident(self.X, self.ldm, esize)
            pieces = max(1, sms // max(1, self.N * NB))
            # window_work splits a piece into at most two row ranges (one wrap): no piece longer than its dataset
            pieces = max([pieces] + [-(-(hi - lo) // max(1, n)) for (lo, hi), n in zip(self.sub, self.X.n_global)])
            work, counts = blocked.window_work(self.X, NB, gbs, plan, pieces)
            self.win = (bp, torch.from_numpy(work).to(self.device), counts)
            self._plan_window_stats(plan, sms)
        return torch.from_numpy(plan).to(self.device)

    window_stats = True          # mini-batch statistics through the sliced pass 2 over block-aligned windows (fp32)
    window_stats_max_bytes = 1 << 30

    def _plan_window_stats(self, plan, sms):
        """Cell-blocked sliced copy of the resident cells (blocks = mini-batch windows) + per-step work lists, so that
        X_mb' H_mb and H_mb' H_mb of a mini-batch (_online_iNMF.py:563-601) are one gather pass with the window's H in
        shared memory instead of k atomics per non-zero."""
        self.win2 = None
        if not (self.window_stats and self.dtype == torch.float32 and blocked.sell_applies(self.ldm, 4)):
            return
        cnt = [hi - lo for lo, hi in self.sub]
        total = int(sum(cnt))
        if total == 0 or plan.shape[0] * total * 8 > self.window_stats_max_bytes:
            return
        # a third copy of the resident cells (8 bytes per non-zero, twice that while it is built): only with room to spare
        free_bytes = torch.cuda.mem_get_info(self.device)[0]
        if 3 * 8 * int(self.X.nnz) > free_bytes // 2:
            return
        cbs = blocked.window_blocks(cnt, self.ldm, 4)
        chunks = max(1, min(sms // max(1, self.N), self.g // 128))
        view = blocked.resident_view(self.X)
        bp2 = blocked.build_pass2(view, self.ldm, 4, sms, cbs=cbs, chunks=chunks)
        sp2 = blocked.build_sell(bp2, self.ldm)
        res = blocked.window_stats_work(sp2.work_host, bp2.blocks, plan, self.X.n_global)
        del bp2
        if res is None:
            return
        work, counts, place, region = res
        dev = self.device
        self.win2 = (sp2, torch.from_numpy(work).to(dev), counts, torch.from_numpy(place).to(dev), total,
                     self._zeros(max(int(region[-1]), 1), self.ldm, name="H_tiles"))

    def stats_window(self, step):
        """XHt, HHt of mini-batch ``step`` (all-reduced) from H_mb through the block-aligned window pass."""
        sp2, work, counts, place, total, Ht = self.win2
        Ht.zero_()
        Ht.index_copy_(0, place[step], self.H_mb[:total])
        self.XHt.zero_()
        self.HHt.zero_()
        if int(counts[step]) > 0:
            self._blocked(sp2, work[step], int(counts[step]), Ht, self.XHt, self.k, self.HHt)
        self._allreduce(self.XHt, self.HHt)

    def online_step(self, seg_desc_t, scales, rollover, n_inner, step=None):
        """One mini-batch: H_mb by BPP, A/B update, HALS on W and V (_online_iNMF.py:416-460).
        ``scales`` = per-dataset scale_param of _update_A_B (they only differ at iteration 1).
        With ``step`` given and the window work lists built, the right-hand sides come from the tiled kernel."""
        self.prep()
        if step is not None and self.win is not None:
            bp, work, counts = self.win
            self.H_mb.zero_()
            self._blocked(bp, work[step], int(counts[step]), self.Mt, self.H_mb, self.ldm, None)
            self.bpp_H(seg_desc=seg_desc_t, max_rows=self.max_sub, out=self.H_mb)
        else:
            self.solve_H(seg_desc=seg_desc_t, max_rows=self.max_sub, out=self.H_mb)
        if step is not None and getattr(self, "win2", None) is not None:
            self.stats_window(step)
        else:
            self.stats(seg_desc=seg_desc_t, max_rows=self.max_sub, H=self.H_mb, St=self.XHt, HtH=self.HHt)
        if len(set(scales)) == 1:
            _lib.call("inmf_update_ab", self.sfx, self.A, self.A_old, self.B, self.B_old, self.HHt, self.XHt,
                      self.inv_mb, float(scales[0]), int(rollover), self.N, self.g, self.k, _stream())
        else:
            for i, sc in enumerate(scales):
                _lib.call("inmf_update_ab", self.sfx, self.A[i], self.A_old[i], self.B[i], self.B_old[i],
                          self.HHt[i], self.XHt[i], self.inv_mb[i:i + 1], float(sc), int(rollover), 1, self.g,
                          self.k, _stream())
        _lib.call("inmf_hals", self.sfx, self.A_all, self.B_all, self.Wt, self.Vt_all, self.lam, self.n_prev,
                  self.n_prev + self.N, self.g, self.k, int(n_inner), _stream())

    def final_H(self, with_V=True):
        """H for every owned cell from the final W, V (_online_iNMF.py:465-508)."""
        self.prep(with_V=with_V)
        self.solve_H()
