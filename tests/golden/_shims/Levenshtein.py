"""Import shim used ONLY by tests/golden/make_golden.py, in the build container where
/root/reference exists: the reference does `import Levenshtein as lev`
(annchor/distances.py:5) and that wheel is not installed here.  Backed by the
repo's CPU oracle (textbook unit-cost DP)."""
from oracle.metrics import levenshtein as distance  # noqa: F401
