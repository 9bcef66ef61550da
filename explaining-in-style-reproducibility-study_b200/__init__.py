"""B200-native AttFind hot path of StylEx (drop-in for the reference's Python surface).

See DESIGN.md.  Host-side mirror of the reference interface lives in ``modules`` / ``attfind`` /
``classifiers``; the CUDA kernels and the C-ABI library are under ``csrc`` (``include/stylex_b200.h``).
"""
from . import synthetic  # noqa: F401
