"""Minimal stand-ins for the two container types the factorisation path reads and writes.

pyliger keeps its data in ``anndata.AnnData`` objects held by a ``Liger`` object
(/root/reference/src/pyliger/pyliger.py:12-105).  ``optimize_ALS`` / ``online_iNMF`` only touch
``liger.adata_list``, ``liger.num_samples``, ``liger.var_genes`` and, per AnnData, ``.shape``,
``.layers["scale_data"]``, ``.obsm``, ``.varm``, ``.uns``, ``.isbacked``
(_iNMF_ANLS.py:79-84,182-185; _online_iNMF.py:131-137,171-176,192-198).  A real pyliger ``Liger`` of
real AnnData objects works unchanged with this package; these classes exist so that the path can
be driven (tests, bench) where anndata is not installed.
"""


class AnnDataLike(object):
    def __init__(self, scale_data, sample_name=None):
        self.X = scale_data
        self.layers = {"scale_data": scale_data}
        self.obsm = {}
        self.varm = {}
        self.uns = {"sample_name": sample_name}
        self.isbacked = False

    @property
    def shape(self):
        return self.X.shape


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


def liger_from_matrices(mats, names=None):
    """Build a ``Liger`` whose datasets carry ``mats`` (scipy CSR, cells x genes) as scale_data."""
    names = names or ["sample%d" % i for i in range(len(mats))]
    g = mats[0].shape[1]
    return Liger([AnnDataLike(m, nm) for m, nm in zip(mats, names)],
                 var_genes=["g%d" % j for j in range(g)])
