"""climt_b200 -- B200-native column radiative transfer behind climt's component surface.

The hot path named in BASELINE.json -- RRTMG longwave / shortwave (with McICA), the CORK correlated-k / picket-fence schemes and
the grey scheme -- plus the column steps SURVEY.md 8(f) lists next to it, as drop-in components whose numerics run in
hand-written sm_100a CUDA kernels behind the C ABI of include/climt_b200.h.  The names below are the ones `climt` itself exports
for these components (climt/__init__.py:4-29,56-83): `import climt_b200 as climt` covers this path and nothing else.
"""
from .constants import get_constant, set_constant, reset_constants  # noqa: F401
from .state import get_interface_values, mass_to_volume_mixing_ratio  # noqa: F401  (climt/_core/util.py:47-142)
from .rrtmg_lw import RRTMGLongwave  # noqa: F401
from .rrtmg_sw import RRTMGShortwave  # noqa: F401
from .cork import CorkLongwaveRadiation, CorkShortwaveRadiation  # noqa: F401
from .gray import GrayLongwaveRadiation  # noqa: F401
from .emanuel import EmanuelConvection, EmanuelConvectionPython  # noqa: F401
from .instellation import Instellation  # noqa: F401
from .berger_solar_insolation import BergerSolarInsolation  # noqa: F401
from .slab_surface import SlabSurface  # noqa: F401
from .simple_physics import SimplePhysics  # noqa: F401

__all__ = ["get_constant", "set_constant", "reset_constants", "get_interface_values", "mass_to_volume_mixing_ratio",
           "RRTMGLongwave", "RRTMGShortwave", "CorkLongwaveRadiation", "CorkShortwaveRadiation", "GrayLongwaveRadiation",
           "EmanuelConvection", "EmanuelConvectionPython", "Instellation", "BergerSolarInsolation", "SlabSurface",
           "SimplePhysics"]
