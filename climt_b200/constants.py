"""Physical constants used by the radiation components.

climt reads these from sympl's global constant registry (third-party, `sympl>=0.5.0`,
setup.py:49 -- not in the reference tree).  Values are sympl's defaults as recorded in
SURVEY.md section 8(c); the ones that enter the flux arithmetic (avogadro, gravity, seconds per
day, cp of dry air, stellar irradiance) are pinned at the 1e-8 level by the RRTMG goldens.
Unlike the reference (process-global registry + Fortran module variables) every engine
instance takes an explicit copy.
"""

DEFAULTS = {
    # name: (value, units)
    "gravitational_acceleration": (9.80665, "m s^-2"),
    "planck_constant": (6.62607004e-34, "J s"),
    "boltzmann_constant": (1.38064852e-23, "J K^-1"),
    "speed_of_light": (299792458.0, "m s^-1"),
    "avogadro_constant": (6.022140857e23, "mole^-1"),
    "loschmidt_constant": (2.6516467e25, "m^-3"),
    "universal_gas_constant": (8.3144598, "J mole^-1 K^-1"),
    "stefan_boltzmann_constant": (5.670367e-8, "W m^-2 K^-4"),
    "seconds_per_day": (86400.0, "dimensionless"),
    "heat_capacity_of_dry_air_at_constant_pressure": (1004.64, "J kg^-1 K^-1"),
    "gas_constant_of_dry_air": (287.0, "J kg^-1 K^-1"),
    "reference_air_pressure": (1.0132e5, "Pa"),
    "top_of_model_pressure": (20.0, "Pa"),       # climt/__init__.py:51
    "stellar_irradiance": (1367.0, "W m^-2"),
    # water as the condensible (sympl's defaults; read by EmanuelConvection, climt/_components/emanuel/component.py:236-246).
    # No golden of the reference pins these five: its Emanuel caches hold only zeros.
    "heat_capacity_of_vapor_phase": (1846.0, "J kg^-1 K^-1"),
    "gas_constant_of_vapor_phase": (461.5, "J kg^-1 K^-1"),
    "latent_heat_of_condensation": (2.5e6, "J kg^-1"),
    "density_of_liquid_phase": (1e3, "kg m^-3"),
    "specific_enthalpy_of_vapor_phase": (2500.0, "J kg^-1"),
    # read by SimplePhysics (climt/_components/simple_physics/component.py:196-207).  Radius and rotation rate enter only its
    # internal-SST branch (simulate_cyclone=True without an external surface temperature); values as in
    # climt/_data/atmospheric_properties/earth.toml:5-6.  No golden of the reference pins these three.
    "planetary_radius": (6.371e6, "m"),
    "planetary_rotation_rate": (7.292e-5, "s^-1"),
    "density_of_liquid_water": (1e3, "kg m^-3"),
    # read by SlabSurface(include_ekman=True) (climt/_components/slab_surface.py:343); sympl's default, no golden pins it
    "heat_capacity_of_sea_water": (3.985e3, "J kg^-1 K^-1"),
}

_registry = {k: v[0] for k, v in DEFAULTS.items()}


def get_constant(name, units=None):
    """sympl.get_constant(name, units): the registry keeps each constant in the SI unit sympl's defaults are given in, which is
    the unit every call site on this path asks for (DEFAULTS); `units` is accepted for signature compatibility."""
    return _registry[name]


def set_constant(name, value):
    _registry[name] = float(value)


def reset_constants():
    _registry.clear()
    _registry.update({k: v[0] for k, v in DEFAULTS.items()})


def rrtmg_constants():
    """The 10 values climt passes to rrtmg_set_constants, in RRTMG's cgs-ish units
    (rrtmg/lw/component.py:298-328) plus cp of dry air."""
    import math
    return {
        "pi": math.pi,
        "grav": get_constant("gravitational_acceleration"),
        "planck": get_constant("planck_constant") * 1e7,            # erg s
        "boltz": get_constant("boltzmann_constant") * 1e7,          # erg K^-1
        "clight": get_constant("speed_of_light") * 1e2,             # cm s^-1
        "avogad": get_constant("avogadro_constant"),
        "alosmt": get_constant("loschmidt_constant") * 1e-6,        # cm^-3
        "gascon": get_constant("universal_gas_constant") * 1e7,     # erg mol^-1 K^-1
        "sbcnst": get_constant("stefan_boltzmann_constant") * 1e-4,  # W cm^-2 K^-4
        "secdy": get_constant("seconds_per_day"),
        "cpdair": get_constant("heat_capacity_of_dry_air_at_constant_pressure"),
    }
