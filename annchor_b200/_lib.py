"""ctypes binding of libannb.so (include/annb.h).  Fails loudly when the CUDA library is
missing or no B200 is present: there is no CPU fallback anywhere in this package."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ANNB_LIBRARY") or os.path.join(_HERE, "libannb.so")  # ANNB_LIBRARY: A/B builds

EUCLIDEAN, COSINE, LEVENSHTEIN, WASSERSTEIN1D, WASSERSTEIN = 0, 1, 2, 3, 4
F32, F64, U8 = 0, 1, 2
METRIC_IDS = {"euclidean": EUCLIDEAN, "cosine": COSINE, "levenshtein": LEVENSHTEIN,
              "wasserstein": WASSERSTEIN1D, "wasserstein1d": WASSERSTEIN1D}


class AnnbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libannb error %d: %s" % (code, msg))
        self.code = code


class IndexParams(C.Structure):
    _fields_ = [("n_anchors", C.c_int32), ("n_neighbors", C.c_int32), ("locality", C.c_int32),
                ("loc_thresh", C.c_int32), ("loc_min", C.c_int32), ("is_metric", C.c_int32),
                ("rank", C.c_int32), ("world", C.c_int32)]


_P, _I64, _I32, _U64 = C.c_void_p, C.c_int64, C.c_int, C.c_uint64
_PP = C.POINTER(C.c_void_p)

# name -> argtypes (all return int unless listed in _RESTYPES)
_SIGS = {
    "annb_ctx_create": [_I32, _PP],
    "annb_ctx_destroy": [_P],
    "annb_pool_trim": [],
    "annb_sync": [_P],
    "annb_ctx_stream": [_P, C.POINTER(_U64)],
    "annb_timer_start": [_P],
    "annb_timer_stop": [_P, C.POINTER(C.c_float)],
    "annb_dataset_dense": [_P, _P, _I64, _I64, _I32, _I32, _PP],
    "annb_dataset_strings": [_P, _P, _P, _I64, _PP],
    "annb_dataset_hist": [_P, _P, _I64, _I64, _I32, _PP],
    "annb_dataset_hist_cost": [_P, _P, _I64, _I64, _I32, _P, _PP],
    "annb_dataset_gather": [_P, _P, _P, _I64, _PP],
    "annb_dataset_free": [_P],
    "annb_pair_dists": [_P, _P, _I32, _P, _I64, _P],
    "annb_pair_dists_dev": [_P, _P, _I32, _P, _P, _I64, _P],
    "annb_maxmin_anchors": [_P, _P, _I32, _I64, _I64, _P, _P],
    "annb_anchor_dists": [_P, _P, _I32, _P, _I64, _P],
    "annb_bounds_ijs": [_P, _P, _I64, _P, _I64, _I64, _P],
    "annb_dad_ijs": [_P, _P, _I64, _P, _I64, _I64, _P],
    "annb_update_bounds": [_P, _P, _I64, _P, _P, _P, _I64, _P],
    "annb_predict_stratified": [_P, _P, _I64, _P, _P, _P, _I64, _P, _P],
    "annb_error_labels": [_P, _P, _I64, _P, _I64, _P],
    "annb_probs": [_P, _P, _P, _I64, _P, _P, _I64, _P],
    "annb_row_kth": [_P, _P, _I64, _P, _P, _I64, _I64, _P],
    "annb_get_nn": [_P, _I64, _I64, _P, _P, _I64, _P, _P, _P, _P, _P],
    "annb_index_create": [_P, _P, _I32, C.POINTER(IndexParams), _PP],
    "annb_index_destroy": [_P],
    "annb_index_reserve_pairs": [_P, _I64],
    "annb_index_maxmin": [_P, _I64, _P],
    "annb_index_set_anchors": [_P, _P, _I64, _P],
    "annb_index_spatial_order": [_P, _P],
    "annb_index_adopt_anchors": [_P, _P, _P],
    "annb_index_get_D": [_P, _P],
    "annb_index_locality": [_P, C.POINTER(_I64), C.POINTER(_I64)],
    "annb_index_sample_pool": [_P, _U64, _I64, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(C.c_int)],
    "annb_index_sample_pool_bins": [_P, _U64, _P, _P, _I64, _I64, C.POINTER(_I64)],
    "annb_index_get_pool": [_P, _P, _P],
    "annb_index_pair_features": [_P, _P, _I64, _P],
    "annb_index_pair_state": [_P, _P, _I64, _P, _P, _P],
    "annb_index_get_thresh": [_P, _P],
    "annb_index_set_lookahead": [_P, _P, _I64],
    "annb_index_add_known": [_P, _P, _P, _I64],
    "annb_index_eval_pairs": [_P, _P, _I64, _P],
    "annb_index_set_model": [_P, _P, _P, _P, _I64, _P, _P],
    "annb_index_row_thresh": [_P, _P],
    "annb_index_guarantee_nmin": [_P, _I64, C.POINTER(_I64)],
    "annb_index_select": [_P, _I64, _I64, C.POINTER(_I64), C.POINTER(_I64)],
    "annb_index_get_selected": [_P, _P, _P],
    "annb_index_refine_selected": [_P, C.POINTER(_I64)],
    "annb_index_update_bounds": [_P, C.POINTER(_I64)],
    "annb_index_neighbor_graph": [_P, _P, _P],
    "annb_index_stats": [_P, _P, _I64],
    "annb_index_last_sweep": [_P, C.POINTER(C.c_float), C.POINTER(_I64)],
    "annb_index_set_reducer": [_P, _P, _P],
    "annb_index_export_refined": [_P, _P, _P, _P, _I64, C.POINTER(_I64)],
    "annb_index_export_tightened": [_P, _P, _P, _P, _P, _I64, C.POINTER(_I64)],
    "annb_index_import_dev": [_P, _I32, _P, _P, _P, _P, _I64],
    "annb_bruteforce_knn": [_P, _P, _I32, _I64, _P, _P],
    "annb_nearest_enemies": [_P, _P, _I32, _P, _I64, _P, _P],
    "annb_pair_dists_query": [_P, _P, _I32, _I64, _P, _I64, _P],
    "annb_index_query": [_P, _P, _I64, _I64, C.c_double, _P, _P, C.POINTER(_I64)],
    "annb_numba_rng_new": [C.c_uint32, _PP],
    "annb_numba_rng_free": [_P],
    "annb_numba_rng_shuffle": [_P, _P, _I64],
}
_RESTYPES = {"annb_last_error": C.c_char_p, "annb_version": C.c_int, "annb_launch_count": _I64,
             "annb_dataset_len": _I64}

_lib = None


def declared_symbols():
    """Every function name declared in include/annb.h (parsed from the header)."""
    import re
    hdr = open(os.path.join(_HERE, "..", "include", "annb.h")).read()
    return sorted(set(re.findall(r"\b(annb_[A-Za-z0-9_]+)\s*\(", hdr)) - {"annb_reduce_fn"})


def load():
    """dlopen libannb.so.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "annchor_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or python annchor_b200/build.py).  There is no CPU fallback." % SO_PATH)
    L = C.CDLL(SO_PATH)
    missing = [n for n in list(_SIGS) + list(_RESTYPES) if not hasattr(L, n)]
    if missing:
        raise ImportError("libannb.so is stale: missing symbols %s -- rebuild it" % missing)
    for name, args in _SIGS.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    for name, rt in _RESTYPES.items():
        getattr(L, name).restype = rt
    L.annb_dataset_len.argtypes = [_P]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise AnnbError(rc, load().annb_last_error().decode(errors="replace"))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def as_c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)
