"""stryke_b200 — B200-native engine for stryke's individual-based Monte Carlo fish-passage simulation.

``from stryke_b200 import simulation`` mirrors ``Stryke.stryke.simulation`` (reference: Stryke/stryke.py:249).
The per-fish work exists only as sm_100a CUDA behind the C ABI in include/stryke_b200.h; importing the package
does not need a GPU, running a simulation does (there is no CPU fallback).
"""
from .epri import epri  # noqa: F401
from .simulation import simulation  # noqa: F401
from .summary import bootstrap_mean_ci, summarize_ci  # noqa: F401  (module-level helpers of stryke.py:136-162)

__all__ = ["simulation", "epri", "bootstrap_mean_ci", "summarize_ci"]
__version__ = "0.1.0"
