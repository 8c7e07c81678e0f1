"""Global settings (counterpart of /root/reference/src/fqe/settings.py:20-51).

The reference switches between a C and a pure-Python code path with
``use_accelerated_code``.  This package has exactly one code path, CUDA on
sm_100a, so the switch is a constant; it is kept so that code written against the
reference's settings module keeps working.
"""
from enum import Enum


class CodePath(Enum):
    """Available code paths: CUDA only."""
    CUDA = "cuda"


available_code_paths = (CodePath.CUDA,)
use_accelerated_code = True

global_max_norb = 63
"""Largest number of spatial orbitals (strings are uint64; the reference's C path
handles 63, settings.py:47-51)."""

workspace_fraction = 0.70
"""Fraction of the free device memory the sigma build may take for its D/E chunk
workspace when the whole D tensor does not fit."""

max_workspace_bytes = None
"""Optional hard cap (bytes) on the sigma workspace; None = no cap."""

symmetry_tolerance = 0.0
"""The contraction runs in the compressed i>=j pair space when the folded two-body tensor
satisfies h2'[ij,kl] == h2'[ji,kl] == h2'[ij,lk] EXACTLY (real-orbital integrals built by a
symmetric procedure).  Integrals that are symmetric only up to rounding can opt in by setting
this to a relative tolerance (e.g. 1e-13): tensors whose asymmetry is below
tolerance * max|h2'| are symmetrised (averaged) on the host first, which perturbs the
operator by at most that amount."""

import os as _os

allreduce_slices = int(_os.environ.get("FQEB_ALLREDUCE_SLICES", "6"))
"""Multi-GPU sigma (fqe_b200.distributed.sharded_apply): the scatter of a rank's last chunk is
issued in this many slices of target rows and each finished slice is all-reduced on a side
stream while the next one is scattered; 1 = one all-reduce after the whole build."""
