"""annchor_b200 -- B200 (sm_100a) implementation of the Annchor.fit() hot path of gchq/annchor.

Python host code over hand-written CUDA behind a C ABI (include/annb.h, libannb.so).  There
is no CPU fallback: importing works anywhere, but every compute call needs the built
library and a B200.
"""
from ._lib import AnnbError, load as load_library  # noqa: F401
from .core import Context, Dataset, GpuExactIJs, default_context, launch_count  # noqa: F401
from . import ops  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # heavy host-side modules are imported lazily
    if name in ("Annchor", "BruteForce", "compare_neighbor_graphs"):
        from . import annchor as _a
        return getattr(_a, name)
    raise AttributeError(name)
