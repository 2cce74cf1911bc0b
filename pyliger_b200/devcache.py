"""Device-resident twins of host CSR matrices produced by this package.

pyliger's API hands scipy matrices from stage to stage (``layers["norm_data"]`` -> ``layers["scale_data"]`` ->
optimize_ALS).  When a stage here has just computed such a matrix on the GPU it registers the device arrays
under the identity of the host matrix it returned; the next stage finds them and skips the host->device copy
("X uploaded once ... never round-trips").  Entries die with the host matrix
(weakref) or on ``clear()``; a matrix that was modified in place after registration must be dropped by the
caller (``forget``) -- the stages only ever register matrices they have just created.
"""
import weakref

_cache = {}


def register(host_matrix, device, rowptr, colidx, vals):
    key = id(host_matrix)
    try:
        ref = weakref.ref(host_matrix, lambda _r, key=key: _cache.pop(key, None))
    except TypeError:  # pragma: no cover - scipy matrices are weak-referenceable
        return
    _cache[key] = (ref, str(device), rowptr, colidx, vals, host_matrix.shape, host_matrix.nnz)


def lookup(host_matrix, device):
    """(rowptr int64, colidx int32, vals) on ``device`` for this very matrix object, or None."""
    ent = _cache.get(id(host_matrix))
    if ent is None:
        return None
    ref, dev, rowptr, colidx, vals, shape, nnz = ent
    if ref() is not host_matrix or dev != str(device) or shape != host_matrix.shape or nnz != host_matrix.nnz:
        return None
    return rowptr, colidx, vals


def forget(host_matrix):
    _cache.pop(id(host_matrix), None)


def clear():
    _cache.clear()
