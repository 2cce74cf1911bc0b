"""Minimal stand-ins for the two container types the factorisation path reads and writes.

pyliger keeps its data in ``anndata.AnnData`` objects held by a ``Liger`` object
(/root/reference/src/pyliger/pyliger.py:12-105).  ``optimize_ALS`` / ``online_iNMF`` only touch
``liger.adata_list``, ``liger.num_samples``, ``liger.var_genes`` and, per AnnData, ``.shape``,
``.layers["scale_data"]``, ``.obsm``, ``.varm``, ``.uns``, ``.isbacked``
(_iNMF_ANLS.py:79-84,182-185; _online_iNMF.py:131-137,171-176,192-198); normalize / scale_not_center
also use ``.X``, ``.var`` / ``.obs`` (pandas), ``.raw`` and ``adata[:, mask].copy()``
(preprocessing/_normalization.py:42-55, _scale.py:92-102).  A real pyliger ``Liger`` of
real AnnData objects works unchanged with this package; these classes exist so that the path can
be driven (tests, bench) where anndata is not installed.
"""


class _Raw(object):
    """What ``adata.raw = adata`` keeps (anndata freezes X and var at that point)."""

    def __init__(self, adata):
        self.X = adata.X
        self._shape = adata.shape
        self.var = adata.var.copy() if adata.var is not None else None

    @property
    def shape(self):
        return self._shape


class AnnDataLike(object):
    """Duck-typed AnnData: X, layers, obs / var (pandas), obsm / varm, uns, slicing by boolean masks and
    ``copy()`` -- the surface that normalize / scale_not_center / optimize_ALS / online_iNMF use."""

    def __init__(self, scale_data=None, sample_name=None, X=None, obs_names=None, var_names=None, shape=None):
        self.X = scale_data if X is None else X
        self.layers = {} if scale_data is None else {"scale_data": scale_data}
        self.obsm = {}
        self.varm = {}
        self.uns = {"sample_name": sample_name}
        self.isbacked = False
        self._raw = None
        self._shape = None if shape is None else (int(shape[0]), int(shape[1]))
        n, g = self.shape
        import pandas as pd
        self.obs = pd.DataFrame(index=pd.Index(obs_names if obs_names is not None
                                               else ["cell%d" % i for i in range(n)]).astype(str))
        self.var = pd.DataFrame(index=pd.Index(var_names if var_names is not None
                                               else ["g%d" % j for j in range(g)]).astype(str))

    @classmethod
    def from_raw(cls, X, sample_name=None, obs_names=None, var_names=None):
        """Dataset holding raw counts only (what create_liger starts from)."""
        return cls(None, sample_name, X=X, obs_names=obs_names, var_names=var_names)

    @classmethod
    def backed(cls, store_path, sample_name, dataset=None, obs_names=None, var_names=None):
        """Dataset whose matrices live in an on-disk ``pyliger_b200.io.CSRStore`` (pyliger's backed mode: nothing but
        the annotations is in memory).  ``dataset``: which matrix of the store gives the shape (default: raw_data,
        else the first one present)."""
        from .io import CSRStore
        with CSRStore(store_path, "r") as f:
            names = f.keys()
            name = dataset or ("raw_data" if "raw_data" in names else names[0])
            shape = f[name].shape
        out = cls(None, sample_name, X=None, obs_names=obs_names, var_names=var_names, shape=shape)
        out.isbacked = True
        out.uns["backed_path"] = store_path
        return out

    def uns_keys(self):
        return list(self.uns.keys())

    @property
    def shape(self):
        return self._shape if self.X is None else self.X.shape

    @property
    def raw(self):
        return self._raw

    @raw.setter
    def raw(self, adata):
        self._raw = None if adata is None else _Raw(adata)

    def copy(self):
        import copy as _copy
        out = AnnDataLike.__new__(AnnDataLike)
        out.X = None if self.X is None else self.X.copy()
        out._shape = self._shape
        out.isbacked = self.isbacked
        out.layers = {name: m.copy() for name, m in self.layers.items()}
        out.obsm = {name: m.copy() for name, m in self.obsm.items()}
        out.varm = {name: m.copy() for name, m in self.varm.items()}
        out.uns = _copy.copy(self.uns)
        out._raw = self._raw
        out.obs = self.obs.copy()
        out.var = self.var.copy()
        return out

    def __getitem__(self, key):
        """adata[rows, cols] with boolean masks, index arrays or slices (a new object, like a materialised view)."""
        import numpy as np
        rows, cols = key if isinstance(key, tuple) else (key, slice(None))

        def pos(sel, n):
            if isinstance(sel, slice):
                return np.arange(n)[sel]
            sel = np.asarray(sel)
            return np.nonzero(sel)[0] if sel.dtype == bool else sel

        n, g = self.shape
        r, c = pos(rows, n), pos(cols, g)
        all_rows, all_cols = len(r) == n and (r == np.arange(n)).all(), len(c) == g and (c == np.arange(g)).all()

        def cut(m):
            m = m if all_rows else m[r]
            return m if all_cols else m[:, c]

        out = AnnDataLike.__new__(AnnDataLike)
        out.X = None if self.X is None else cut(self.X)
        out._shape = None if self._shape is None else (len(r), len(c))
        out.layers = {name: cut(m) for name, m in self.layers.items()}
        out.obsm = {name: m[r] for name, m in self.obsm.items()}
        out.varm = {name: m[c] for name, m in self.varm.items()}
        out.uns = dict(self.uns)
        out.isbacked = self.isbacked
        out._raw = self._raw
        out.obs = self.obs.iloc[r]
        out.var = self.var.iloc[c]
        return out


class Liger(object):
    """Same attribute surface as pyliger.Liger for the factorisation path (pyliger.py:79-105)."""

    def __init__(self, adata_list=None, var_genes=None):
        self.adata_list = [] if adata_list is None else adata_list
        self.var_genes = var_genes

    @property
    def num_samples(self):
        return len(self.adata_list)

    @property
    def num_var_genes(self):
        return len(self.var_genes)

    @property
    def H(self):
        return [adata.obsm["H"] for adata in self.adata_list]

    @property
    def V(self):
        return [adata.varm["V"] for adata in self.adata_list]

    @property
    def W(self):
        return self.adata_list[0].varm["W"]


def liger_from_raw(mats, names=None, var_names=None):
    """``Liger`` of datasets that carry raw count matrices only (the input of normalize)."""
    names = names or ["sample%d" % i for i in range(len(mats))]
    return Liger([AnnDataLike.from_raw(m, nm, var_names=var_names) for m, nm in zip(mats, names)])


def liger_from_matrices(mats, names=None):
    """Build a ``Liger`` whose datasets carry ``mats`` (scipy CSR, cells x genes) as scale_data."""
    names = names or ["sample%d" % i for i in range(len(mats))]
    g = mats[0].shape[1]
    return Liger([AnnDataLike(m, nm) for m, nm in zip(mats, names)],
                 var_genes=["g%d" % j for j in range(g)])
