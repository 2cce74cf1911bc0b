"""ctypes binding of libinmf_b200.so (the C ABI in include/inmf_b200.h).

There is NO fallback: if the shared library is missing, or a call returns an error, this raises.
The library is built in-tree by ``pyliger_b200/build.py`` (``python -m pyliger_b200.build`` or
``__graft_entry__.build()``) and travels with the repository snapshot to the GPU box.
"""
import ctypes
import os

from .build import LIB_PATH

_c = ctypes
_I32, _I64, _F64, _PTR = _c.c_int32, _c.c_int64, _c.c_double, _c.c_void_p

# name -> argtypes (without the _f32/_f64 suffix; P = device pointer or stream)
_TYPED = {
    "inmf_bpp": [_PTR, _PTR, _I64, _PTR, _I64, _PTR, _I64, _PTR, _I32, _PTR, _I32, _I64, _I32, _PTR, _PTR, _I64, _PTR],
    "inmf_prep": [_PTR, _PTR, _PTR, _I64, _PTR, _PTR, _PTR, _F64, _I32, _I64, _I32, _PTR],
    "inmf_spmm": [_PTR, _PTR, _PTR, _PTR, _I64, _I64, _PTR, _I64, _PTR, _I32, _I64, _I32, _PTR],
    "inmf_stats": [_PTR, _PTR, _PTR, _PTR, _I64, _PTR, _PTR, _I64, _PTR, _I32, _I64, _I32, _PTR],
    "inmf_sqnorm": [_PTR, _PTR, _PTR, _I32, _PTR, _PTR],
    "inmf_rhs_v": [_PTR, _PTR, _PTR, _PTR, _PTR, _F64, _I32, _I64, _I32, _PTR],
    "inmf_rhs_w": [_PTR, _PTR, _PTR, _PTR, _PTR, _I32, _I64, _I32, _PTR],
    "inmf_objective": [_PTR, _PTR, _I64, _PTR, _PTR, _PTR, _PTR, _F64, _I32, _I64, _I32, _PTR, _PTR],
    "inmf_update_ab": [_PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR, _F64, _I32, _I32, _I64, _I32, _PTR],
    "inmf_hals": [_PTR, _PTR, _PTR, _PTR, _F64, _I32, _I32, _I64, _I32, _I32, _PTR],
    "inmf_normal_eq": [_PTR, _PTR, _I64, _I32, _I64, _PTR, _PTR, _PTR],
    "inmf_gram_rows": [_PTR, _I64, _PTR, _I32, _I64, _PTR, _I32, _PTR],
    "inmf_blocked_spmm": [_PTR, _I32, _PTR, _PTR, _PTR, _I32, _PTR, _I64, _I32, _I32, _PTR, _PTR],
    "inmf_normalize_rows": [_PTR, _PTR, _PTR, _I64, _PTR, _PTR, _PTR, _PTR, _PTR],
    "inmf_select_scale": [_PTR, _PTR, _PTR, _I64, _PTR, _PTR, _PTR, _PTR, _PTR, _PTR],
    "inmf_reduce_parts": [_PTR, _I32, _I64, _I64, _PTR, _PTR],
    "inmf_hals_h": [_PTR, _PTR, _I64, _PTR, _I64, _PTR, _I32, _I64, _I32, _PTR],
    "inmf_blockfmt1": [_I32, _PTR, _PTR, _PTR, _PTR, _I32, _I64, _I32, _I32, _I32, _PTR, _PTR, _PTR],
    "inmf_blockfmt2": [_I32, _PTR, _PTR, _PTR, _PTR, _I32, _I32, _I32, _I32, _PTR, _PTR, _PTR, _PTR, _PTR],
}

EXPORTED_SYMBOLS = ["inmf_version", "inmf_last_error", "inmf_launch_count", "inmf_note_graph_launches", "inmf_bpp_workspace_bytes", "inmf_tc_selftest",
                    "inmf_tc_relayout_f32", "inmf_tc_pass_f32", "inmf_tc_debug_buffer", "inmf_select_count", "inmf_mt19937_uniform", "inmf_blocked_split_f32", "inmf_blocked_wide_f32",
                    "inmf_sell_pass_f32", "inmf_sell_pass_mc_f32", "inmf_sell_fill_f32",
                    "inmf_qn_max_factor", "inmf_qn_knn", "inmf_qn_vote", "inmf_qn_align",
                    "inmf_sell_pass_mixed", "inmf_to_bf16_rows", "inmf_to_bf16_gram", "inmf_set_deterministic"] + [
    "%s_%s" % (n, t) for n in _TYPED for t in ("f32", "f64")
]

_lib = None


class InmfError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise InmfError(
            "libinmf_b200.so not found at %s: build it with `python -m pyliger_b200.build` "
            "(nvcc, sm_100a). pyliger_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.inmf_version.restype = _I32
    lib.inmf_version.argtypes = []
    lib.inmf_last_error.restype = _c.c_char_p
    lib.inmf_last_error.argtypes = []
    lib.inmf_launch_count.restype = _I64
    lib.inmf_launch_count.argtypes = []
    lib.inmf_note_graph_launches.restype = None
    lib.inmf_note_graph_launches.argtypes = [_I64]
    lib.inmf_bpp_workspace_bytes.restype = _I64
    lib.inmf_bpp_workspace_bytes.argtypes = [_I64, _I32, _I32]
    lib.inmf_tc_selftest.restype = _I32
    lib.inmf_tc_selftest.argtypes = [_PTR, _PTR, _PTR, _I32, _I32, _PTR]
    lib.inmf_tc_relayout_f32.restype = _I32
    lib.inmf_tc_relayout_f32.argtypes = [_PTR, _I64, _PTR, _I32, _I64, _PTR, _PTR, _I32, _I32, _PTR]
    lib.inmf_tc_pass_f32.restype = _I32
    lib.inmf_tc_pass_f32.argtypes = [_PTR, _I32, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _I32, _I32, _I32, _I32, _PTR]
    lib.inmf_mt19937_uniform.restype = _I32
    lib.inmf_mt19937_uniform.argtypes = [_c.c_uint32, _I64, _F64, _PTR, _PTR]
    for name in ("inmf_blocked_split_f32", "inmf_blocked_wide_f32"):
        getattr(lib, name).restype = _I32
        getattr(lib, name).argtypes = _TYPED["inmf_blocked_spmm"]
    lib.inmf_sell_pass_mc_f32.restype = _I32
    lib.inmf_sell_pass_mc_f32.argtypes = [_PTR, _I32, _PTR, _PTR, _PTR, _I32, _PTR, _I32, _PTR, _I32, _I32, _PTR, _PTR, _PTR]
    lib.inmf_sell_pass_f32.restype = _I32
    lib.inmf_sell_pass_f32.argtypes = [_PTR, _I32, _PTR, _PTR, _PTR, _I32, _PTR, _I32, _PTR, _I64, _I32, _I32, _PTR, _I64, _I64,
                                       _PTR]
    lib.inmf_sell_fill_f32.restype = _I32
    lib.inmf_sell_fill_f32.argtypes = [_PTR, _PTR, _PTR, _PTR, _I64, _I32, _PTR, _PTR]
    lib.inmf_sell_pass_mixed.restype = _I32
    lib.inmf_sell_pass_mixed.argtypes = [_PTR, _I32, _PTR, _PTR, _PTR, _PTR, _PTR, _I64, _I32, _I32, _PTR]
    lib.inmf_set_deterministic.restype = _I32
    lib.inmf_set_deterministic.argtypes = [_I32]
    lib.inmf_to_bf16_gram.restype = _I32
    lib.inmf_to_bf16_gram.argtypes = [_PTR, _I64, _PTR, _I32, _I64, _I32, _PTR, _PTR, _PTR]
    lib.inmf_to_bf16_rows.restype = _I32
    lib.inmf_to_bf16_rows.argtypes = [_PTR, _I64, _I64, _I32, _PTR, _PTR]
    lib.inmf_qn_max_factor.restype = _I32
    lib.inmf_qn_max_factor.argtypes = [_PTR, _I64, _I64, _PTR, _I32, _I32, _PTR, _PTR, _PTR]
    lib.inmf_qn_knn.restype = _I32
    lib.inmf_qn_knn.argtypes = [_PTR, _I64, _I64, _PTR, _I32, _I32, _PTR, _PTR]
    lib.inmf_qn_vote.restype = _I32
    lib.inmf_qn_vote.argtypes = [_PTR, _I32, _I64, _PTR, _PTR]
    lib.inmf_qn_align.restype = _I32
    lib.inmf_qn_align.argtypes = [_PTR, _I32, _PTR, _I64, _PTR, _I64, _PTR, _PTR, _PTR, _PTR, _PTR, _I32, _I32, _PTR]
    lib.inmf_select_count.restype = _I32
    lib.inmf_select_count.argtypes = [_PTR, _PTR, _I64, _PTR, _PTR, _PTR]
    for name, argtypes in _TYPED.items():
        for suffix in ("f32", "f64"):
            fn = getattr(lib, "%s_%s" % (name, suffix))
            fn.restype = _I32
            fn.argtypes = argtypes
    _lib = lib
    return lib


def _ptr(x):
    """torch tensor / int / None -> device address."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    return x.data_ptr()


def call(name, suffix, *args):
    """Invoke ``<name>_<suffix>``; tensors are passed by data pointer. Raises InmfError on failure."""
    lib = load()
    fn = getattr(lib, "%s_%s" % (name, suffix) if suffix else name)
    conv = [_ptr(a) if not isinstance(a, (float, bool)) else a for a in args]
    rc = fn(*conv)
    if rc != 0:
        raise InmfError("%s_%s failed (%d): %s" % (name, suffix, rc, lib.inmf_last_error().decode()))
    return rc


def launch_count():
    """Kernels launched by libinmf_b200.so so far in this process."""
    return int(load().inmf_launch_count())


def bpp_workspace(max_seg_cols, nseg, k, device):
    """Workspace tensor for inmf_bpp_* (G^-1 + the half-warp solver's deferral lists), sized by the library."""
    import torch
    nbytes = int(load().inmf_bpp_workspace_bytes(int(max_seg_cols), int(nseg), int(k)))
    if nbytes < 0:
        raise InmfError("inmf_bpp_workspace_bytes: bad arguments")
    return torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=device)


def ws_args(ws):
    """(pointer, byte count) pair of a workspace tensor (or (None, 0))."""
    if ws is None:
        return None, 0
    return ws, ws.numel() * ws.element_size()
