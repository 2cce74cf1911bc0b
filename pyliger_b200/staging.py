"""Host -> device upload of large numpy arrays through a small ring of pinned staging buffers.

A plain ``tensor.to(device)`` from pageable memory is a single-threaded driver memcpy (~10 GB/s); the X arrays
of a factorisation call are hundreds of MB, so this is most of the call's fixed cost.  Here a few worker threads
copy (and, if asked, cast float64 -> float32 on the fly, halving the PCIe bytes) chunks into pinned buffers and
queue asynchronous copies on a side stream, so the host-side pass and the DMA overlap.  The pinned ring is
allocated once per device and reused by later calls.
"""
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

import os

CHUNK_BYTES = 8 << 20


def _default_threads():
    """Host threads copying into the pinned ring (measured at C2 on a 16-core host: 8 -> 14-19 ms, 16 -> 12.7 ms):
    one per core up to 16, shared fairly between the processes of a multi-GPU launch (torchrun's LOCAL_WORLD_SIZE) so
    that 8 ranks do not run 128 copy threads on 32 cores."""
    cores = os.cpu_count() or 8
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(2, min(16, cores // local_world))


NBUF = _default_threads()
_TORCH_TO_NP = {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64}

_stagers = {}
_lock = threading.Lock()


class _Stager(object):
    def __init__(self, device):
        self.device = device
        self.bufs = [torch.empty(CHUNK_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(NBUF)]
        self.stream = torch.cuda.Stream(device)
        self.pool = ThreadPoolExecutor(NBUF)
        self.busy = threading.Lock()

    def _worker(self, b, chunks):
        buf = self.bufs[b]
        ev = None
        with torch.cuda.device(self.device):
            for dst, src in chunks[b::NBUF]:
                if ev is not None:
                    ev.synchronize()  # the previous DMA out of this buffer is done
                view = buf[:dst.numel() * dst.element_size()].view(dst.dtype)
                np.copyto(view.numpy(), src, casting="unsafe")  # releases the GIL; casts if the dtypes differ
                with torch.cuda.stream(self.stream):
                    dst.copy_(view, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.stream)
            if ev is not None:
                ev.synchronize()

    def _down_worker(self, b, chunks):
        buf = self.bufs[b]
        with torch.cuda.device(self.device):
            for dst, src in chunks[b::NBUF]:
                view = buf[:src.numel() * src.element_size()].view(src.dtype)
                with torch.cuda.stream(self.stream):
                    view.copy_(src, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self.stream)
                ev.synchronize()
                np.copyto(dst, view.numpy(), casting="unsafe")  # widens float32 -> float64 on the host side

    def download(self, jobs):
        """jobs: [(dst 1-D numpy array, src 1-D contiguous device tensor of the same length)]. Returns bytes."""
        chunks = []
        got = 0
        for dst, src in jobs:
            n = src.numel()
            if n == 0:
                continue
            if dst.shape[0] != n:
                raise ValueError("staged download: length mismatch")
            per = CHUNK_BYTES // src.element_size()
            for o in range(0, n, per):
                chunks.append((dst[o:o + per], src[o:o + per]))
            got += n * src.element_size()
        if not chunks:
            return 0
        with self.busy:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            futs = [self.pool.submit(self._down_worker, b, chunks) for b in range(min(NBUF, len(chunks)))]
            for f in futs:
                f.result()
        return got

    def upload(self, jobs):
        """jobs: [(dst 1-D device tensor, src 1-D numpy array of the same length)]. Returns the bytes sent."""
        chunks = []
        sent = 0
        for dst, src in jobs:
            n = dst.numel()
            if n == 0:
                continue
            if src.shape[0] != n:
                raise ValueError("staged upload: length mismatch")
            per = CHUNK_BYTES // dst.element_size()
            for o in range(0, n, per):
                chunks.append((dst[o:o + per], src[o:o + per]))
            sent += n * dst.element_size()
        if not chunks:
            return 0
        with self.busy:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))  # dst allocations are ordered before
            futs = [self.pool.submit(self._worker, b, chunks) for b in range(min(NBUF, len(chunks)))]
            for f in futs:
                f.result()
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return sent


def _get(device):
    device = torch.device(device)
    with _lock:
        st = _stagers.get(device)
        if st is None:
            st = _stagers[device] = _Stager(device)
    return st


def upload(device, jobs):
    return _get(device).upload(jobs)


def download(device, jobs):
    return _get(device).download(jobs)


def to_host_f64(device, tensors):
    """Device tensors (any float dtype, any strides) -> new host float64 arrays of the same shapes, in one
    overlapped staged download (the narrow dtype crosses the bus; widening happens on the host)."""
    outs, jobs = [], []
    for t in tensors:
        flat = t.contiguous().view(-1)
        out = np.empty(tuple(t.shape), dtype=np.float64)
        outs.append(out)
        jobs.append((out.reshape(-1), flat))
    download(device, jobs)
    return outs
