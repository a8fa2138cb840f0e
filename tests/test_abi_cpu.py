"""The C-ABI boundary (include/annb.h <-> annchor_b200/libannb.so <-> ctypes table) without a GPU:
the library loads, exports every symbol the header declares, and the ctypes binding covers all of
them.  No compute call is made here (those are the `-m gpu` tests)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from annchor_b200 import build
    return build.build()


def test_header_symbols_exported(lib_path):
    from annchor_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 50
    L = ctypes.CDLL(lib_path)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, "declared in include/annb.h but not exported: %s" % missing


def test_ctypes_table_matches_header(lib_path):
    from annchor_b200 import _lib
    declared = set(_lib.declared_symbols())
    bound = set(_lib._SIGS) | set(_lib._RESTYPES)
    assert declared - bound == set(), "declared but not bound in _lib.py: %s" % sorted(declared - bound)
    assert bound - declared == set(), "bound but not declared in annb.h: %s" % sorted(bound - declared)
    # argument counts agree with the prototypes
    hdr = open(os.path.join(ROOT, "include", "annb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for name, args in _lib._SIGS.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, hdr, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), "%s: header has %d parameters, ctypes table %d" % (name, n, len(args))


def test_load_binds_everything(lib_path):
    from annchor_b200 import _lib
    L = _lib.load()
    assert L.annb_version() >= 100
    assert L.annb_launch_count() >= 0


def test_product_path_has_no_oracle_import():
    """Nothing under annchor_b200/ may import the CPU oracle (oracle/ is test infrastructure)."""
    pkg = os.path.join(ROOT, "annchor_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_host_only_entry_point_numba_rng(lib_path):
    """annb_numba_rng_* is pure host code (numba's in-@njit MT19937 + Fisher-Yates, used by the
    exact-mode sampler, annchor/utils.py:555-557,572): callable without a GPU, and it must agree with
    the oracle's restatement (which tests/test_oracle.py pins against draws of the real numba)."""
    import numpy as np
    from annchor_b200.plugins import NumbaRNG
    from oracle.pipeline import NumbaRNG as OracleRNG
    for seed in (0, 42, 43, 2**31 + 5):
        a, b = NumbaRNG(seed), OracleRNG(seed)
        for n, size in ((10, 3), (1000, 714), (7, 7), (1, 1)):
            x = np.arange(n, dtype=np.int64) * 3 + 1
            np.testing.assert_array_equal(a.choice_no_replace(x, size), b.choice_no_replace(x, size))
