"""ORACLE (test infrastructure only): numpy restatement of climt's grey longwave scheme.

Follows `_gray_lw_kernel_np` (climt/_components/radiation.py:162-190) and the tendency arithmetic of
`GrayLongwaveRadiation.array_call` (:89-109).  Pinned by tests/golden (TestGrayLongwaveRadiation-{column,3d}).
"""
import numpy as np


def gray_lw(T, p_interface, T_surface, tau, sigma, g, cpd):
    nlev, ncol = T.shape
    upward_flux = np.zeros((nlev + 1, ncol))
    downward_flux = np.zeros((nlev + 1, ncol))
    T4 = sigma * T ** 4
    upward_flux[0] = sigma * T_surface ** 4
    for k in range(1, nlev + 1):
        dtau = tau[k] - tau[k - 1]
        trans = np.exp(-dtau)
        upward_flux[k] = upward_flux[k - 1] * trans + T4[k - 1] * (1.0 - trans)
    downward_flux[nlev] = 0.0
    for k in range(nlev - 1, -1, -1):
        dtau = tau[k + 1] - tau[k]
        trans = np.exp(-dtau)
        downward_flux[k] = downward_flux[k + 1] * trans + T4[k] * (1.0 - trans)
    net = upward_flux - downward_flux
    tend = g / cpd * (net[1:] - net[:-1]) / (p_interface[1:] - p_interface[:-1])
    return {"lw_down": downward_flux, "lw_up": upward_flux, "tendency": tend, "tendency_per_day": tend * 86400.0}
