"""Label gate shared by the GPU parity tests: see oracle/parity.py (margin 1e-5, hard mismatches must be 0)."""
from oracle.parity import MARGIN, label_parity  # noqa: F401
