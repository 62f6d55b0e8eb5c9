"""Physical constants for the snapshot drivers: `units()` exposes the same attribute names as the reference's
`units_library.units` (library/units_library.py), e.g. `units().rho_crit`."""

# attribute -> (value, unit)
_CONSTANTS = {
    "rho_crit": (2.77536627e11, "critical density, h^2 Msun / Mpc^3"),
    "c_kms": (3e5, "speed of light, km / s"),
    "Mpc_cm": (3.0856e24, "cm per Mpc"),
    "kpc_cm": (3.0856e21, "cm per kpc"),
    "Msun_g": (1.989e33, "g per solar mass"),
    "Ymass": (0.24, "helium mass fraction"),
    "mH_g": (1.6726e-24, "proton mass, g"),
    "yr_s": (3.15576e7, "s per year"),
    "km_cm": (1e5, "cm per km"),
    "kB": (1.3806e-26, "Boltzmann constant, g (km/s)^2 / K"),
    "nu0_MHz": (1420.0, "rest frequency of the 21-cm line, MHz"),
}


class units(object):
    def __init__(self):
        for name, (value, _unit) in _CONSTANTS.items():
            setattr(self, name, value)

    @staticmethod
    def describe(name):
        """Unit / meaning of one constant."""
        return _CONSTANTS[name][1]
