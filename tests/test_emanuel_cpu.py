"""Emanuel convection without a GPU:
 * the oracle (oracle/emanuel_oracle.cpp, a restatement of the Fortran CONVECT 4.3c) against golden vectors produced by running
   the reference's numba port of the same routine on convecting soundings (tests/golden/make_emanuel_golden.py);
 * the CUDA engine's per-thread code, compiled for the host over a NaN-poisoned workspace (tests/emul/emanuel_emul.cpp), against
   the golden vectors and against the oracle with the Fortran component's constants and saturation humidity.
Tolerance: BASELINE.json asks 1e-6 relative; asserted at 1e-10 of each field's scale, convective_state exactly."""
import numpy as np
import pytest

import helpers as H
from oracle import emanuel as OE

CASES = ("tropics60", "tropics30_longstep", "levels72")
SYMPL = dict(cpd=1004.64, cpv=1846.0, cl=2500.0, rv=461.5, rd=287.0, lv0=2.5e6, g=9.80665, rowl=1e3)


def _ref(z, case):
    return {k: z[f"{case}/out/{n}"] for k, n in H.EMANUEL_OUT.items()}


def _flat(tend, diag):
    return {"ft": tend["air_temperature"], "fq": tend["specific_humidity"], "fu": tend["eastward_wind"], "fv": tend["northward_wind"],
            "iflag": diag["convective_state"], "precip": diag["convective_precipitation_rate"],
            "wd": diag["convective_downdraft_velocity_scale"], "tprime": diag["convective_downdraft_temperature_scale"],
            "qprime": diag["convective_downdraft_specific_humidity_scale"], "cbmf": diag["cloud_base_mass_flux"],
            "cape": diag["atmosphere_convective_available_potential_energy"]}


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_the_reference_numba_port(case):
    z = np.load(H.EMANUEL_GOLDEN)
    st = H.emanuel_case(z, case)
    tend, diag = OE.python_component_call(st, float(st["timestep"]))
    ref = _ref(z, case)
    H.emanuel_compare(_flat(tend, diag), ref, 1e-12, case)
    assert (ref["iflag"] == 1).sum() > 20 and (ref["iflag"] == 4).sum() > 0 and (ref["iflag"] == 0).sum() > 0
    np.testing.assert_allclose(diag["air_temperature_tendency_from_convection"], z[f"{case}/out/air_temperature_tendency_from_convection"],
                               rtol=0, atol=1e-10 * np.abs(ref["ft"]).max() * 86400)
    # the two saturation-humidity formulas, against the reference's own evaluations
    np.testing.assert_allclose(OE.python_qs(st["air_temperature"], st["air_pressure"], 287.04, 461.5), st["qs_python"], rtol=1e-14)
    np.testing.assert_allclose(OE.bolton_q_sat(st["air_temperature"], st["air_pressure"] * 100, 287.0, 461.5), st["qs_bolton"], rtol=1e-14)


@pytest.mark.parametrize("case", CASES)
def test_kernel_code_on_host_matches_the_reference_numba_port(case):
    z = np.load(H.EMANUEL_GOLDEN)
    st = H.emanuel_case(z, case)
    got = H.run_emanuel_emul(OE.PYTHON_DEFAULTS, H.emanuel_arrays(st), float(st["timestep"]), qs_mode=2)
    H.emanuel_compare(got, _ref(z, case), 1e-10, case)
    for k, v in got.items():
        assert np.all(np.isfinite(v)), k     # nothing was read from the poisoned workspace
    # saturation humidity handed in (qs_mode 0) == computed in the kernel (qs_mode 2)
    got0 = H.run_emanuel_emul(OE.PYTHON_DEFAULTS, H.emanuel_arrays(st, qs=st["qs_python"]), float(st["timestep"]), qs_mode=0)
    H.emanuel_compare(got0, got, 1e-12, case + " qs given")
    # the radiation engines' (level, column) layout, read in place through strides: bit-identical
    a0 = {k: (np.ascontiguousarray(v.T) if v.ndim == 2 else v) for k, v in H.emanuel_arrays(st).items()}
    got_t = H.run_emanuel_emul(OE.PYTHON_DEFAULTS, a0, float(st["timestep"]), qs_mode=2, layout=0)
    for k, v in got.items():
        assert np.array_equal(got_t[k].T if v.ndim == 2 else got_t[k], v), k


@pytest.mark.parametrize("case", CASES[:2])
def test_kernel_code_on_host_matches_oracle_with_fortran_component_settings(case):
    """climt.EmanuelConvection: sympl's constants, bolton_q_sat, rain/snow switch at 273.0 K, and non-default parameters."""
    z = np.load(H.EMANUEL_GOLDEN)
    st = H.emanuel_case(z, case)
    tweak = dict(minorig=2, entp=1.2, sigd=0.08, sigs=0.2, cu=0.5, dtmax=0.6, alpha=0.2, damp=0.15, elcrit=0.0009)
    tend, diag = OE.fortran_component_call(st, float(st["timestep"]), SYMPL, par=tweak)
    par = dict(OE.FORTRAN_DEFAULTS, **SYMPL, **tweak)
    got = H.run_emanuel_emul(par, H.emanuel_arrays(st), float(st["timestep"]), qs_mode=1)
    H.emanuel_compare(got, _flat(tend, diag), 1e-10, case)
    assert (got["iflag"] >= 1).sum() > 10


def test_early_exits_and_conservation():
    """Columns built to leave through each early return (convect43c.f90:480-519), and the scheme's own invariants on the rest:
    column enthalpy and momentum tendencies integrate to zero (:1118-1136)."""
    from climt_b200 import synthetic as SY
    st = SY.make_emanuel_state(64, 40, seed=9)
    st["specific_humidity"][0, :] = 0.0              # Q(NK) <= 0 -> state 0
    st["air_temperature"][1, :] = 200.0              # T(NK) < 250 -> state 0
    st["specific_humidity"][2, :] = 1e-9             # bone dry -> LCL below 200 mbar -> state 2
    st["cloud_base_mass_flux"][3] = 0.0
    par = OE.PYTHON_DEFAULTS
    qs = OE.python_qs(st["air_temperature"], st["air_pressure"], par["rd"], par["rv"])
    ref = OE.convect(par, st["air_temperature"], st["specific_humidity"], qs, st["eastward_wind"], st["northward_wind"],
                     st["air_pressure"], st["air_pressure_on_interface_levels"], st["cloud_base_mass_flux"], 900.0)
    got = H.run_emanuel_emul(par, H.emanuel_arrays(st), 900.0, qs_mode=2)
    H.emanuel_compare(got, ref, 1e-10)
    assert got["iflag"][0] == 0 and got["iflag"][1] == 0 and got["iflag"][2] == 2
    for c in (0, 1, 2):
        assert got["cbmf"][c] == 0.0 and not got["ft"][c].any() and got["cape"][c] == 0.0
    conv = got["iflag"] >= 1
    dp = st["air_pressure_on_interface_levels"][:, :-1] - st["air_pressure_on_interface_levels"][:, 1:]
    q, T = st["specific_humidity"], st["air_temperature"]
    cpn = par["cpd"] * (1 - q) + par["cpv"] * q
    lv = par["lv0"] - (par["cl"] - par["cpv"]) * (T - 273.15)
    ents = ((cpn * got["ft"] + lv * got["fq"]) * dp).sum(axis=1)
    scale = (np.abs(cpn * got["ft"]) * dp).sum(axis=1) + 1e-30
    assert np.all(np.abs(ents[conv]) / scale[conv] < 1e-9)
    mom = (got["fu"] * dp).sum(axis=1)
    assert np.all(np.abs(mom[conv]) / ((np.abs(got["fu"]) * dp).sum(axis=1)[conv] + 1e-30) < 1e-9)
