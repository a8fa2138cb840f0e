"""Import shim used ONLY by tests/golden/make_golden.py (see Levenshtein.py here).
The reference wraps `kantorovich(x, y, cost=M)` in an @njit closure
(annchor/utils.py:82-84), so the stand-in must itself be numba-jittable.  It
implements the 1-D ground-cost case |a-b| (closed form on CDFs); `cost` is
accepted and ignored, so only use it with that cost matrix."""
import numpy as np
from numba import njit


@njit()
def kantorovich(x, y, cost=None):
    sx = 0.0
    sy = 0.0
    for k in range(x.shape[0]):
        sx += x[k]
        sy += y[k]
    cx = 0.0
    cy = 0.0
    w = 0.0
    for k in range(x.shape[0]):
        cx += x[k] / sx
        cy += y[k] / sy
        w += abs(cx - cy)
    return w
