"""climt_b200 -- B200-native column radiative transfer behind climt's component surface.

Only the hot path named in BASELINE.json is here: RRTMG longwave / shortwave (and, later, CORK and
Gray) as drop-in components whose numerics run in hand-written sm_100a CUDA kernels.
"""
from .constants import get_constant, set_constant, reset_constants  # noqa: F401

__all__ = ["get_constant", "set_constant", "reset_constants"]
