"""Device-resident twins of host CSR matrices produced by this package.

pyliger's API hands scipy matrices from stage to stage (``layers["norm_data"]`` -> ``layers["scale_data"]`` ->
optimize_ALS).  When a stage here has just computed such a matrix on the GPU it registers the device arrays
under the identity of the host matrix it returned; the next stage finds them and skips the host->device copy
("X uploaded once ... never round-trips").  Entries die with the host matrix
(weakref) or on ``clear()``.

Reuse is safe by construction: the ``data`` / ``indices`` / ``indptr`` arrays of a registered host matrix are made
READ-ONLY (``flags.writeable = False``), so an in-place edit raises instead of silently diverging from the device
copy; whoever wants to edit takes a copy (no twin) or re-enables ``writeable`` -- which ``lookup`` detects and
treats as "edited": the twin is dropped and the matrix is uploaded like any other.  A cheap sampled fingerprint
is kept as a second line of defence against arrays swapped for others of the same size (``forget`` drops an
entry explicitly).
"""
import weakref

import numpy as np

_cache = {}


def _fingerprint(m):
    """Cheap content check of a host CSR matrix (strided samples of data / indices + the ends of indptr): catches
    in-place edits made after the device twin was registered without touching all of the matrix."""
    step = max(1, m.nnz // 4096)
    return (int(m.nnz), tuple(m.shape), float(np.asarray(m.data[::step], dtype=np.float64).sum()),
            int(np.asarray(m.indices[::step], dtype=np.int64).sum()), int(m.indptr[-1]))


def _arrays(m):
    return [a for a in (getattr(m, "data", None), getattr(m, "indices", None), getattr(m, "indptr", None))
            if isinstance(a, np.ndarray)]


def _frozen(m):
    return all(not a.flags.writeable for a in _arrays(m))


def register(host_matrix, device, rowptr, colidx, vals):
    key = id(host_matrix)
    for a in _arrays(host_matrix):
        a.flags.writeable = False
    try:
        ref = weakref.ref(host_matrix, lambda _r, key=key: _cache.pop(key, None))
    except TypeError:  # pragma: no cover - scipy matrices are weak-referenceable
        return
    _cache[key] = (ref, str(device), rowptr, colidx, vals, _fingerprint(host_matrix))


def lookup(host_matrix, device):
    """(rowptr int64, colidx int32, vals) on ``device`` for this very matrix object, or None."""
    ent = _cache.get(id(host_matrix))
    if ent is None:
        return None
    ref, dev, rowptr, colidx, vals, fp = ent
    if ref() is not host_matrix or dev != str(device):
        return None
    if not _frozen(host_matrix) or fp != _fingerprint(host_matrix):  # made writable / arrays swapped: stale
        _cache.pop(id(host_matrix), None)
        return None
    return rowptr, colidx, vals


def forget(host_matrix):
    _cache.pop(id(host_matrix), None)


def clear():
    _cache.clear()
    _dense.clear()


# ---- dense arrays (obsm["H"]): device twins for the stages that consume the factors (quantile_norm) -------------
_dense = {}


def register_dense(host_array, device, tensor):
    """Remember the device tensor a host float64 array was downloaded from.  The host array is frozen (read-only), like
    the CSR twins above."""
    if not isinstance(host_array, np.ndarray):
        return
    key = id(host_array)
    try:
        ref = weakref.ref(host_array, lambda _r, key=key: _dense.pop(key, None))
    except TypeError:  # pragma: no cover
        return
    host_array.flags.writeable = False
    _dense[key] = (ref, str(device), tensor)


def lookup_dense(host_array, device):
    ent = _dense.get(id(host_array))
    if ent is None:
        return None
    ref, dev, tensor = ent
    if ref() is not host_array or dev != str(device) or host_array.flags.writeable or \
            tuple(tensor.shape) != tuple(host_array.shape):
        _dense.pop(id(host_array), None)
        return None
    return tensor
