"""Out-of-core online_iNMF: X is never resident on the GPU.

The reference's backed mode reads every mini-batch from an on-disk CSR store in row chunks
(/root/reference/src/pyliger/factorization/_online_iNMF.py:401-412) and the final H pass in miniBatch_size blocks
(:480-490).  Here the rows of step t+1 are gathered on the host (memory-mapped store -> pinned buffers, by a
prefetch thread) and copied to one of two device slots on a side stream while the kernels of step t run, so disk /
host reads, the PCIe copy and the GPU work overlap.  Every rank reads only ITS 1/world share of each mini-batch and of
the final pass: X is sharded over the ranks, not replicated.

The device side is the same kernel sequence as the resident path with the CSR row kernels (inmf_spmm / inmf_bpp /
inmf_stats / inmf_update_ab / inmf_hals) applied to the slot's CSR: a step's slot holds, dataset after dataset, the
rows (t*mb_i + q) mod n_i of its mini-batch window.
"""
import threading

import numpy as np
import scipy.sparse as sps
import torch

from .engine import InmfEngine, shard_bounds


def _rows_of(src, lo, hi):
    """(indptr rebased to 0, indices, data) of rows [lo, hi) of a CSRDataset / h5sparse dataset / scipy matrix."""
    if hasattr(src, "raw_rows"):
        return src.raw_rows(lo, hi)
    m = src[lo:hi]
    if not sps.isspmatrix_csr(m):
        m = sps.csr_matrix(m)
    if not m.has_sorted_indices:
        m = m.sorted_indices()
    return m.indptr, m.indices, m.data


def _window_ranges(start, count, n):
    """Rows (start + q) mod n, q in [0, count), as contiguous [lo, hi) ranges (wraps as often as needed)."""
    out = []
    pos = start % n
    left = count
    while left > 0:
        take = min(left, n - pos)
        out.append((pos, pos + take))
        left -= take
        pos = 0
    return out


class _Slot(object):
    """Pinned host staging + device CSR arrays of one step (grown on demand)."""

    def __init__(self, device, dtype, rows):
        self.device, self.dtype = device, dtype
        self.rows = rows
        self.h_rowptr = torch.zeros(rows + 1, dtype=torch.int64, pin_memory=True)
        self.d_rowptr = torch.zeros(rows + 1, dtype=torch.int64, device=device)
        self.cap = 0
        self.h_col = self.h_val = self.d_col = self.d_val = None
        self.ready = torch.cuda.Event()      # H2D of this slot finished (recorded on the copy stream)
        self.consumed = torch.cuda.Event()   # kernels reading this slot finished (recorded on the compute stream)
        self.consumed_valid = False
        self.nnz = 0

    def reserve(self, nnz):
        if nnz > self.cap:
            self.cap = int(nnz * 1.25) + 1024
            self.h_col = torch.empty(self.cap, dtype=torch.int32, pin_memory=True)
            self.h_val = torch.empty(self.cap, dtype=self.dtype, pin_memory=True)
            self.d_col = torch.empty(self.cap, dtype=torch.int32, device=self.device)
            self.d_val = torch.empty(self.cap, dtype=self.dtype, device=self.device)

    # what InmfEngine's CSR kernels read
    @property
    def rowptr(self):
        return self.d_rowptr

    @property
    def colidx(self):
        return self.d_col

    @property
    def vals(self):
        return self.d_val


class RowStreamer(object):
    """Double-buffered host -> device stream of row sets.  ``plan(t)`` returns, for request t, a list of
    (source, [(lo, hi), ...]) whose rows are concatenated in order into the slot."""

    def __init__(self, device, dtype, max_rows, plan, count):
        self.device, self.dtype = device, dtype
        self.plan, self.count = plan, count
        self.slots = [_Slot(device, dtype, max_rows), _Slot(device, dtype, max_rows)]
        self.copy_stream = torch.cuda.Stream(device)
        self.h2d_bytes = 0
        self._thread = None
        self._err = None
        if count > 0:
            self._start(0)

    def _fill(self, t):
        try:
            slot = self.slots[t % 2]
            parts = []
            nnz = 0
            for src, ranges in self.plan(t):
                for lo, hi in ranges:
                    ip, ix, dv = _rows_of(src, lo, hi)
                    parts.append((ip, ix, dv))
                    nnz += len(ix)
            if slot.consumed_valid:
                slot.consumed.synchronize()  # the kernels of request t-2 are done with this slot
            slot.reserve(nnz)
            rp = slot.h_rowptr.numpy()
            col = slot.h_col.numpy()
            val = slot.h_val.numpy()
            r = 0
            p = 0
            rp[0] = 0
            for ip, ix, dv in parts:
                nr = len(ip) - 1
                m = len(ix)
                np.add(ip[1:], p, out=rp[r + 1:r + 1 + nr], casting="unsafe")
                np.copyto(col[p:p + m], ix, casting="unsafe")
                np.copyto(val[p:p + m], dv, casting="unsafe")   # float64 stores are narrowed on the host side
                r += nr
                p += m
            rp[r + 1:] = p  # requests shorter than the slot: empty tail rows
            slot.nnz = nnz
            with torch.cuda.device(self.device), torch.cuda.stream(self.copy_stream):
                slot.d_rowptr.copy_(slot.h_rowptr, non_blocking=True)
                if nnz:
                    slot.d_col[:nnz].copy_(slot.h_col[:nnz], non_blocking=True)
                    slot.d_val[:nnz].copy_(slot.h_val[:nnz], non_blocking=True)
                slot.ready.record(self.copy_stream)
            self.h2d_bytes += slot.h_rowptr.numel() * 8 + nnz * (4 + slot.h_val.element_size())
        except BaseException as e:  # surfaced by get()
            self._err = e

    def _start(self, t):
        self._thread = threading.Thread(target=self._fill, args=(t,), daemon=True)
        self._thread.start()

    def get(self, t):
        """Slot of request t (its copy ordered before the current stream's next kernels); request t+1 starts loading."""
        self._thread.join()
        if self._err is not None:
            raise self._err
        slot = self.slots[t % 2]
        torch.cuda.current_stream(self.device).wait_event(slot.ready)
        if t + 1 < self.count:
            self._start(t + 1)
        return slot

    def release(self, slot):
        """The kernels reading ``slot`` have been enqueued on the current stream."""
        slot.consumed.record(torch.cuda.current_stream(self.device))
        slot.consumed_valid = True


class StreamingEngine(InmfEngine):
    """InmfEngine whose X lives in a store: W, V, A, B and the mini-batch buffers are resident, the cells stream."""

    def __init__(self, sources, k, value_lambda, dtype=torch.float64, device=None, group=None, n_prev=0):
        self.sources = list(sources)
        self.n_cells = [int(s.shape[0]) for s in self.sources]
        g = int(self.sources[0].shape[1])
        placeholders = [sps.csr_matrix((0, g), dtype=np.float64) for _ in self.sources]
        super().__init__(placeholders, k, value_lambda, dtype=dtype, device=device, group=group, n_prev=n_prev,
                         use_blocked=False, need_stats=False)
        self.stream_h2d_bytes = 0

    def _ensure_ws(self, rows):
        """BPP workspace (and its declared size) for streamed batches of ``rows`` cells per dataset."""
        if rows > self.max_rows_full:
            from . import _lib
            self.max_rows_full = int(rows)
            self.ws_H = _lib.bpp_workspace(self.max_rows_full, max(self.N, 1), self.k, self.device)

    # ------------------------------------------------------------------ mini-batch loop
    def stream_minibatches(self, total_iters, step_args, n_inner):
        """Runs ``total_iters`` mini-batch steps; ``step_args(it)`` -> (scales, rollover).  alloc_online() first."""
        N = self.N
        cnt = [hi - lo for lo, hi in self.sub]
        self._ensure_ws(max(cnt))
        seg = np.zeros((N + 1, 4), dtype=np.int64)
        seg[:, 0] = np.concatenate([[0], np.cumsum(cnt)])
        seg[:N, 1] = seg[:N, 0]          # the slot holds the datasets' rows back to back
        seg[:N, 3] = np.maximum(cnt, 1)
        seg_t = torch.from_numpy(seg).to(self.device)

        def plan(t):
            return [(self.sources[i], _window_ranges(t * self.mbs[i] + self.sub[i][0], cnt[i], self.n_cells[i]))
                    for i in range(N)]

        streamer = RowStreamer(self.device, self.dtype, int(sum(cnt)), plan, total_iters)
        resident = self.X
        try:
            for it in range(total_iters):
                slot = streamer.get(it)
                self.X = slot
                scales, rollover = step_args(it)
                self.online_step(seg_t, scales, rollover, n_inner, step=None)
                streamer.release(slot)
        finally:
            self.X = resident
        torch.cuda.current_stream(self.device).synchronize()
        self.stream_h2d_bytes += streamer.h2d_bytes

    # ------------------------------------------------------------------ final H pass
    def stream_final_H(self, with_V=True, chunk_rows=32768):
        """H of every cell this rank owns, streamed in row chunks (_online_iNMF_cal_H, _online_iNMF.py:465-508).
        Returns a list of host float64 arrays (this rank's rows of every dataset)."""
        N = self.N
        self._ensure_ws(chunk_rows)
        self.prep(with_V=with_V)
        own = [shard_bounds(n, self.rank, self.world) for n in self.n_cells]
        reqs = []
        for i, (lo, hi) in enumerate(own):
            for a in range(lo, hi, chunk_rows):
                reqs.append((i, a, min(hi, a + chunk_rows)))
        outs = [np.zeros((hi - lo, self.k), dtype=np.float64) for lo, hi in own]
        if not reqs:
            return outs
        streamer = RowStreamer(self.device, self.dtype, chunk_rows, lambda t: [(self.sources[reqs[t][0]],
                                                                               [(reqs[t][1], reqs[t][2])])], len(reqs))
        Hc = [torch.zeros(chunk_rows, self.ldm, dtype=self.dtype, device=self.device) for _ in range(2)]
        host = [torch.empty(chunk_rows, self.ldm, dtype=self.dtype, pin_memory=True) for _ in range(2)]
        done = [None, None]
        pending = [None, None]
        resident = self.X
        cur = torch.cuda.current_stream(self.device)

        def drain(b):
            if pending[b] is not None:
                done[b].synchronize()
                i, a, e = pending[b]
                outs[i][a - own[i][0]:e - own[i][0]] = host[b][:e - a, :self.k].numpy()
                pending[b] = None
        try:
            for t, (i, a, e) in enumerate(reqs):
                b = t % 2
                drain(b)
                slot = streamer.get(t)
                self.X = slot
                seg = np.zeros((N + 1, 4), dtype=np.int64)
                seg[i + 1:, 0] = e - a           # only segment i has rows
                seg[:, 3] = 1
                seg[i, 3] = max(e - a, 1)
                seg_t = torch.from_numpy(seg).to(self.device)
                self.solve_H(seg_desc=seg_t, max_rows=e - a, out=Hc[b])
                streamer.release(slot)
                host[b].copy_(Hc[b], non_blocking=True)
                done[b] = torch.cuda.Event()
                done[b].record(cur)
                pending[b] = (i, a, e)
            drain(0)
            drain(1)
        finally:
            self.X = resident
        self.stream_h2d_bytes += streamer.h2d_bytes
        return outs
