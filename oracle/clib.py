"""ctypes loader for oracle/_build/liboracle.so (test infrastructure only)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    """Compile oracle.c with gcc (make -C oracle)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None

_P = C.c_void_p
_I = C.c_int64


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_lev.restype = _I
        L.orc_lev.argtypes = [_P, _I, _P, _I]
        L.orc_lev_pairs.argtypes = [_P, _P, _P, _I, _P]
        for nm in ("orc_euclid_pairs_f32", "orc_euclid_pairs_f64", "orc_cosine_pairs_f32",
                   "orc_cosine_pairs_f64", "orc_w1_pairs_f64"):
            getattr(L, nm).argtypes = [_P, _I, _P, _I, _P]
        L.orc_ot_pairs_f64.argtypes = [_P, _I, _P, _P, _I, _P]
        L.orc_ot_pairs_f64.restype = C.c_int
        L.orc_bounds_ijs.argtypes = [_P, _I, _P, _I, _P]
        L.orc_dad_ijs.argtypes = [_P, _I, _P, _I, _I, _P]
        L.orc_update_bounds.argtypes = [_P, _I, _P, _P, _P, _P]
        L.orc_row_kth.argtypes = [_P, _P, _P, _I, _I, _P]
        L.orc_probs.argtypes = [_P, _P, _I, _P, _P, _P]
        L.orc_get_nn.argtypes = [_I, _I, _P, _P, _P, _P, _P, _P, _P]
        L.orc_f32_features.argtypes = [_P, _I, _P, _I, _P, _P, _P, _P]
        L.orc_f32_predict.argtypes = [_P, _P, _P, _I, _P, _P, _P, _P, _P]
        L.orc_f32_update_bounds.argtypes = [_P, _I, _P, _P, _P, _P, _P]
        L.orc_mt_state_size.restype = C.c_size_t
        L.orc_mt_seed.argtypes = [_P, C.c_uint32]
        L.orc_numba_shuffle.argtypes = [_P, _P, _I]
        _lib = L
    return _lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)
