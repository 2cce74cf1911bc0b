"""On-disk CSR store for the backed ("online learning") mode of pyliger.

The reference keeps the raw / normalised / scaled matrices of a backed AnnData in ``./results/<sample_name>.hdf5``
through ``h5sparse`` -- append-able CSR datasets read back in row chunks
(/root/reference/src/pyliger/_utilities.py:137-154 writer, preprocessing/_normalization.py:61-88 and _scale.py:62-89
chunked stages, factorization/_online_iNMF.py:410-412, 480-490, 555-560 readers).  h5py / h5sparse are not part of this
image, and HDF5's chunk cache is not what a GPU pipeline wants anyway: the mini-batch loop reads CONTIGUOUS ROW RANGES.
``CSRStore`` offers the same surface on raw files:

    with CSRStore("./results/ctrl.csr", "w") as f:
        f.create_dataset("raw_data", data=csr_chunk)        # like h5sparse.File.create_dataset
        f["raw_data"].append(next_chunk)                     # like h5sparse Dataset.append
    f = CSRStore("./results/ctrl.csr", "r")
    f["scale_data"][left:right]                              # scipy CSR of the rows, like h5sparse slicing
    f["scale_data"].raw_rows(left, right)                    # (indptr, indices, data) memmap views, zero copy

Layout: ``<path>/<dataset>/{indptr.i64, indices.i32, data.<f32|f64>, meta.json}``; three flat binary files, appended in
place, memory-mapped for reading (the page cache is the only buffer between the disk and the pinned staging ring that
feeds the GPU).  A real ``h5sparse.File`` works with the streaming code of this package too (it only needs
``file["scale_data"][l:r]`` -> scipy CSR), this store is what the package itself writes.
"""
import json
import os

import numpy as np
import scipy.sparse as sps

_DT = {"float32": np.float32, "float64": np.float64}


class CSRDataset(object):
    """One append-able CSR matrix (rows x cols) in a directory of flat files."""

    def __init__(self, path, mode):
        self.path = path
        self.mode = mode
        self._maps = None
        with open(os.path.join(path, "meta.json")) as fh:
            self.meta = json.load(fh)

    # -- writing ---------------------------------------------------------------------------------
    @classmethod
    def create(cls, path, data, dtype=None):
        os.makedirs(path, exist_ok=False)
        data = sps.csr_matrix(data)
        dt = np.dtype(dtype or (data.dtype if data.dtype in (np.float32, np.float64) else np.float64)).name
        with open(os.path.join(path, "meta.json"), "w") as fh:
            json.dump({"rows": 0, "cols": int(data.shape[1]), "nnz": 0, "dtype": dt}, fh)
        with open(os.path.join(path, "indptr.i64"), "wb") as fh:
            np.zeros(1, dtype=np.int64).tofile(fh)
        for name in ("indices.i32", "data." + ("f32" if dt == "float32" else "f64")):
            open(os.path.join(path, name), "wb").close()
        ds = cls(path, "r+")
        ds.append(data)
        return ds

    def _data_file(self):
        return os.path.join(self.path, "data." + ("f32" if self.meta["dtype"] == "float32" else "f64"))

    def append(self, m):
        if self.mode == "r":
            raise IOError("CSRStore opened read-only")
        m = sps.csr_matrix(m)
        if m.shape[1] != self.meta["cols"]:
            raise ValueError("appended rows have %d columns, the dataset has %d" % (m.shape[1], self.meta["cols"]))
        m.sort_indices()
        with open(os.path.join(self.path, "indptr.i64"), "ab") as fh:
            (m.indptr[1:].astype(np.int64) + self.meta["nnz"]).tofile(fh)
        with open(os.path.join(self.path, "indices.i32"), "ab") as fh:
            m.indices.astype(np.int32).tofile(fh)
        with open(self._data_file(), "ab") as fh:
            m.data.astype(_DT[self.meta["dtype"]]).tofile(fh)
        self.meta["rows"] += int(m.shape[0])
        self.meta["nnz"] += int(m.nnz)
        with open(os.path.join(self.path, "meta.json"), "w") as fh:
            json.dump(self.meta, fh)
        self._maps = None

    # -- reading ---------------------------------------------------------------------------------
    @property
    def shape(self):
        return (self.meta["rows"], self.meta["cols"])

    @property
    def nnz(self):
        return self.meta["nnz"]

    @property
    def dtype(self):
        return np.dtype(self.meta["dtype"])

    def _open(self):
        if self._maps is None:
            n, nnz = self.meta["rows"], self.meta["nnz"]
            indptr = np.memmap(os.path.join(self.path, "indptr.i64"), dtype=np.int64, mode="r", shape=(n + 1,))
            if nnz:
                indices = np.memmap(os.path.join(self.path, "indices.i32"), dtype=np.int32, mode="r", shape=(nnz,))
                data = np.memmap(self._data_file(), dtype=_DT[self.meta["dtype"]], mode="r", shape=(nnz,))
            else:
                indices, data = np.zeros(0, np.int32), np.zeros(0, _DT[self.meta["dtype"]])
            self._maps = (indptr, indices, data)
        return self._maps

    def raw_rows(self, lo, hi):
        """(indptr rebased to 0, indices, data) of rows [lo, hi): views of the memory maps, nothing copied."""
        indptr, indices, data = self._open()
        p0, p1 = int(indptr[lo]), int(indptr[hi])
        return indptr[lo:hi + 1] - p0, indices[p0:p1], data[p0:p1]

    def __getitem__(self, key):
        if isinstance(key, tuple):
            rows, cols = key
            return self[rows][:, cols]
        if isinstance(key, slice):
            lo, hi, step = key.indices(self.meta["rows"])
            if step != 1:
                raise IndexError("CSRStore datasets are sliced by contiguous row ranges")
            hi = max(hi, lo)
            ip, ix, dv = self.raw_rows(lo, hi)
            return sps.csr_matrix((np.asarray(dv, dtype=np.float64), np.asarray(ix), np.asarray(ip)),
                                  shape=(hi - lo, self.meta["cols"]))
        idx = np.asarray(key)
        if idx.dtype == bool:
            idx = np.nonzero(idx)[0]
        parts = [self[int(i):int(i) + 1] for i in np.atleast_1d(idx)]
        return sps.vstack(parts).tocsr() if parts else sps.csr_matrix((0, self.meta["cols"]))


class CSRStore(object):
    """Directory of CSRDatasets with the slice of h5sparse.File's API that pyliger's backed mode uses."""

    def __init__(self, path, mode="r"):
        if mode not in ("r", "r+", "w", "a"):
            raise ValueError("mode must be 'r', 'r+', 'a' or 'w'")
        self.path = path
        self.mode = mode
        if mode == "w":
            if os.path.isdir(path):
                import shutil
                shutil.rmtree(path)
            os.makedirs(path)
        elif not os.path.isdir(path):
            if mode == "a":
                os.makedirs(path)
            else:
                raise IOError("no CSR store at %s" % path)
        self._open = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        self._open.clear()

    def keys(self):
        return sorted(d for d in os.listdir(self.path) if os.path.isfile(os.path.join(self.path, d, "meta.json")))

    def __contains__(self, name):
        return name in self.keys()

    def __getitem__(self, name):
        if name not in self._open:
            p = os.path.join(self.path, name)
            if not os.path.isfile(os.path.join(p, "meta.json")):
                raise KeyError(name)
            self._open[name] = CSRDataset(p, "r" if self.mode == "r" else "r+")
        return self._open[name]

    def create_dataset(self, name, data=None, dtype=None, **_h5_options):
        """``chunks=`` / ``maxshape=`` of the h5sparse call are accepted and ignored (every dataset is append-able)."""
        if self.mode == "r":
            raise IOError("CSRStore opened read-only")
        ds = CSRDataset.create(os.path.join(self.path, name), data, dtype)
        self._open[name] = ds
        return ds


def store_path(adata):
    """Where the backed mode of this package keeps a sample: ./results/<sample_name>.csr (the reference:
    ./results/<sample_name>.hdf5, _online_iNMF.py:555-560)."""
    return os.path.join(".", "results", str(adata.uns["sample_name"]) + ".csr")


def open_backed(adata, mode="r"):
    """The reference's ``_get_scale_data`` for a backed AnnData: its on-disk store.  An h5sparse file of the reference's
    own layout is used when this package's store does not exist and h5sparse is importable."""
    p = adata.uns.get("backed_path") or store_path(adata)
    if os.path.isdir(p):
        return CSRStore(p, mode)
    h5 = os.path.join(".", "results", str(adata.uns["sample_name"]) + ".hdf5")
    if os.path.exists(h5):
        try:
            import h5sparse
        except ImportError:
            raise IOError("%s is an HDF5 store but h5sparse / h5py are not installed; write a CSRStore instead "
                          "(pyliger_b200.io.write_store)" % h5)
        return h5sparse.File(h5, "r" if mode == "r" else "r+")
    raise IOError("backed AnnData %r has no store at %s" % (adata.uns.get("sample_name"), p))


def h5_idx_generator(chunk_size, matrix_size):
    """Row chunks [0,c), [c,2c), ... clipped at matrix_size (pyliger/_utilities.py:100-116)."""
    lo = 0
    while lo < matrix_size:
        yield lo, min(lo + chunk_size, matrix_size)
        lo += chunk_size


def write_store(path, name, matrix, chunk_size=1000, dtype=None, mode="a"):
    """Write (or add) dataset ``name`` = ``matrix`` (scipy sparse, cells x genes) chunk by chunk, like
    ``_create_h5_using_adata`` (_utilities.py:137-154).  Returns the store path."""
    matrix = sps.csr_matrix(matrix)
    with CSRStore(path, mode) as f:
        for lo, hi in h5_idx_generator(chunk_size, matrix.shape[0]):
            if name not in f:
                f.create_dataset(name, data=matrix[lo:hi], dtype=dtype)
            else:
                f[name].append(matrix[lo:hi])
        if matrix.shape[0] == 0 and name not in f:
            f.create_dataset(name, data=matrix, dtype=dtype)
    return path
