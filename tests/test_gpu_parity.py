"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle's
converged ("tight") evaluation of the reference's integrands on the same node
grids and splines.  Bar: 1e-5 relative on P(k) and w(theta) (BASELINE.json
north_star); intermediate tables are held tighter so that a failure points at
the stage that caused it."""
import numpy as np
import pytest

from chomp_b200 import _lib, design, engine
from oracle import chomp_oracle as O

from common import (C_DICT, C_DICT_2, H_DICT, H_DICT_2, HOD_DICT, HOD_DICT_2, oracle_wtheta,
                    rel_err, w_err)

pytestmark = pytest.mark.gpu

TOL_FINAL = 1e-5      # north_star bar, P(k) and w(theta)
TOL_TABLE = 2e-6      # intermediate tables


def _survey(power_spec="power_gg", **kw):
    dist = engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1)
    return engine.Survey(dist, bins_per_decade=10.0, power_spec=power_spec, **kw)


def _run_point(eng, survey, cosmo, halo, hod, which):
    c = engine.pack_params([cosmo], _lib.COSMO_KEYS)
    h = engine.pack_params([halo], _lib.HALO_KEYS)
    g = np.zeros((1, _lib.N_HOD))
    keys = _lib.HOD_ZHENG_KEYS if survey.hod_kind == _lib.HOD_ZHENG else _lib.HOD_MANDELBAUM_KEYS
    g[0, :len(keys)] = [hod[k] for k in keys]
    import torch
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    w = eng.wtheta(c, h, g, survey.theta, which, status=status)
    torch.cuda.synchronize()
    return w.cpu().numpy()[0], int(status.cpu()[0])


# concentrations below 1 at the high-mass end (c = 2.2 (M / M*)^-0.25): nodes that cannot use the
# small-argument series of the profile (its recurrence runs forward, c >= 1) and the both-arguments-
# small branch of the table-driven profile
H_DICT_LOW_C = dict(H_DICT, c0=2.2, beta=-0.25)

CASES = [
    ("base", C_DICT, H_DICT, HOD_DICT),
    ("low_concentration", C_DICT, H_DICT_LOW_C, HOD_DICT),
    ("cosmo2", C_DICT_2, H_DICT, HOD_DICT),
    ("halo2", C_DICT, H_DICT_2, HOD_DICT),
    ("hod2", C_DICT, H_DICT, HOD_DICT_2),
]


@pytest.mark.parametrize("name,cosmo,halo,hod", CASES, ids=[c[0] for c in CASES])
def test_stage_tables_against_oracle(name, cosmo, halo, hod):
    survey = _survey()
    eng = engine.Engine(survey)
    w, status = _run_point(eng, survey, cosmo, halo, hod, _lib.P_GG)
    assert status == 0
    ref = oracle_wtheta(cosmo, halo, hod, ("gaussian", (0.0, 2.0, 0.5, 0.1)))
    nz, nw = survey.precision["cosmo_npoints"], survey.precision["window_npoints"]
    chi = eng.table(_lib.T_CHI_NODES, 1).cpu().numpy()[0].reshape(3, nz)
    assert rel_err(chi[0][1:], ref["chi_nodes"][1:]) < 1e-9
    win = eng.table(_lib.T_WINDOW_NODES, 1).cpu().numpy()[0].reshape(2, nw)
    peak = np.max(np.abs(ref["wa_nodes"]))
    assert np.max(np.abs(win[0] - ref["wa_nodes"]))/peak < 1e-8
    assert abs(eng.table(_lib.T_ZBAR, 1).cpu().numpy()[0, 0] - ref["z_bar"]) < 1e-12
    assert rel_err(eng.table(_lib.T_DBAR, 1).cpu().numpy()[0, 0], ref["D_z"]) < 1e-10
    kn = eng.table(_lib.T_KERNEL_NODES, 1).cpu().numpy()[0]
    assert np.max(np.abs(kn - ref["kernel_nodes"]))/np.max(np.abs(ref["kernel_nodes"])) < 1e-7
    ep = dict(zip(_lib.EPOCH_FIELDS, eng.table(_lib.T_EPOCH, 1).cpu().numpy()[0]))
    assert rel_err(ep["sigma_norm"], ref["sigma_norm"]) < 1e-8
    assert rel_err(ep["growth"], ref["growth"]) < 1e-12
    assert rel_err(eng.table(_lib.T_LNM_NODES, 1).cpu().numpy()[0], ref["ln_mass_nodes"]) < 1e-12
    assert rel_err(eng.table(_lib.T_NU_NODES, 1).cpu().numpy()[0], ref["nu_nodes"]) < 1e-7
    assert rel_err(ep["f_norm"], ref["f_norm"]) < 1e-7
    assert rel_err(ep["bias_norm"], ref["bias_norm"]) < 1e-7
    assert rel_err(ep["ln_m_star"], ref["ln_m_star"]) < 1e-8
    assert rel_err(eng.table(_lib.T_NBAR, 1).cpu().numpy()[0, 0], ref["n_bar_over_rho_bar"]) < TOL_TABLE
    nk = survey.precision["halo_npoints"]
    tabs = eng.table(_lib.T_HALO_NODES, 1).cpu().numpy()[0].reshape(5, nk)
    for i, nm in enumerate(("h_m", "pp_mm", "h_g", "pp_gm", "pp_gg")):
        assert rel_err(tabs[i], ref[nm]) < TOL_TABLE, nm
    assert w_err(w, ref["w"]) < TOL_FINAL


@pytest.mark.parametrize("spec", ["linear_power", "power_mm", "power_gm", "power_gg"])
def test_power_spectra_and_wtheta(spec):
    survey = _survey(spec)
    eng = engine.Engine(survey)
    which = _lib.POWER_SPEC[spec]
    w, status = _run_point(eng, survey, C_DICT, H_DICT, HOD_DICT, which)
    assert status == 0
    ref = oracle_wtheta(C_DICT, H_DICT, HOD_DICT, ("gaussian", (0.0, 2.0, 0.5, 0.1)), power_spec=spec)
    k = np.logspace(-3.5, 2.2, 200)      # includes the k < k_min and k > k_max branches
    P = eng.power(1, which, k).cpu().numpy()[0]
    Pref = ref["halo"].power(spec, k)
    assert np.all((P == 0) == (Pref == 0))
    nz = Pref != 0
    assert rel_err(P[nz], Pref[nz]) < TOL_FINAL
    assert w_err(w, ref["w"]) < TOL_FINAL


def test_named_shape_200_k_nodes():
    """halo_npoints = 200 (the named benchmark shape) on both sides."""
    prec = dict(O.DEFAULT_PRECISION, halo_npoints=200)
    survey = _survey(precision=prec)
    eng = engine.Engine(survey)
    w, status = _run_point(eng, survey, C_DICT, H_DICT, HOD_DICT, _lib.P_GG)
    assert status == 0
    ref = oracle_wtheta(C_DICT, H_DICT, HOD_DICT, ("gaussian", (0.0, 2.0, 0.5, 0.1)), prec=prec)
    assert w_err(w, ref["w"]) < TOL_FINAL


def test_synthetic_batch_points_and_batch_invariance():
    """Latin-hypercube points: each against the oracle, and the same point gives
    bit-identical results whatever batch it is evaluated in."""
    import torch
    survey = _survey()
    eng = engine.Engine(survey)
    cosmo, halo, hod = design.synthetic_batch(64)
    w = eng.wtheta(cosmo, halo, hod, survey.theta, _lib.P_GG).cpu().numpy()
    assert np.all(np.isfinite(w))
    for i, (cd, hd, gd) in enumerate(design.as_dicts(cosmo, halo, hod)[:4]):
        ref = oracle_wtheta(cd, hd, gd, ("gaussian", (0.0, 2.0, 0.5, 0.1)))
        assert w_err(w[i], ref["w"]) < TOL_FINAL, i
    sub = slice(16, 48)
    w2 = eng.wtheta(cosmo[sub], halo[sub], hod[sub], survey.theta, _lib.P_GG).cpu().numpy()
    assert np.array_equal(w2, w[sub])
    wh, status = eng.wtheta_host(cosmo, halo, hod, survey.theta, _lib.P_GG)
    assert np.array_equal(wh, w) and not status.any()
    torch.cuda.synchronize()


# ------------------------------------------------------------------ the other BASELINE configs
MAND = {"log_M_0": 12.3, "w": 1.2}
CONFIGS = {
    # config 1: matter w(theta), Gaussian dN/dz at z0 = 1
    "cfg1_mm": dict(dist_a=("gaussian", (0.0, 2.0, 1.0, 0.2)), power_spec="power_mm", bins_per_decade=5.0),
    # config 3: galaxy-galaxy lensing gamma_t: HODMandelbaum power_gm, galaxy x convergence, J2
    "cfg3_gammat": dict(dist_a=("gaussian", (0.0, 2.0, 0.4, 0.1)), dist_b=("gaussian", (0.0, 2.0, 1.0, 0.2)),
                        window_a="galaxy", window_b="convergence", bessel_order=2, hod_kind="mandelbaum",
                        power_spec="power_gm", bins_per_decade=5.0),
    # the unit tests' magnitude-limited lens sample (int b: Python-2 z_max cap) x convergence
    "maglim_conv": dict(dist_a=("maglim", (0.0, 2.0, 2, 0.3, 2)), dist_b=("gaussian", (0.0, 2.0, 1.0, 0.2)),
                        window_a="galaxy", window_b="convergence", power_spec="power_gm", bins_per_decade=5.0),
    # float-b magnitude-limited sample, convergence x convergence (config 4's windows)
    "maglim_shear": dict(dist_a=("maglim", (0.0, 2.0, 2.0, 0.5, 2.0)), window_a="convergence",
                         power_spec="power_mm", bins_per_decade=5.0),
}


def _dist(spec):
    kind, args = spec
    mk = engine.RedshiftDistribution.gaussian if kind == "gaussian" else engine.RedshiftDistribution.maglim
    return mk(*args)


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_baseline_configs_against_oracle(name):
    import json
    import os
    kw = dict(CONFIGS[name])
    hod = MAND if kw.get("hod_kind") == "mandelbaum" else HOD_DICT
    survey = engine.Survey(_dist(kw["dist_a"]), _dist(kw["dist_b"]) if "dist_b" in kw else None,
                           kw.get("window_a", "galaxy"), kw.get("window_b"),
                           bins_per_decade=kw["bins_per_decade"], bessel_order=kw.get("bessel_order", 0),
                           power_spec=kw["power_spec"], hod=kw.get("hod_kind", "zheng"))
    eng = engine.Engine(survey)
    which = _lib.POWER_SPEC[kw["power_spec"]]
    w, status = _run_point(eng, survey, C_DICT, H_DICT, hod, which)
    assert status == 0
    ref = oracle_wtheta(C_DICT, H_DICT, hod, **kw)
    nw = survey.precision["window_npoints"]
    win = eng.table(_lib.T_WINDOW_NODES, 1).cpu().numpy()[0].reshape(2, nw)
    for got, want in ((win[0], ref["wa_nodes"]), (win[1], ref["wb_nodes"])):
        assert np.max(np.abs(got - want))/np.max(np.abs(want)) < 1e-7
    assert abs(eng.table(_lib.T_ZBAR, 1).cpu().numpy()[0, 0] - ref["z_bar"]) < 1e-12
    kn = eng.table(_lib.T_KERNEL_NODES, 1).cpu().numpy()[0]
    assert np.max(np.abs(kn - ref["kernel_nodes"]))/np.max(np.abs(ref["kernel_nodes"])) < 1e-6
    assert np.array_equal(survey.theta, ref["theta"])
    assert w_err(w, ref["w"]) < TOL_FINAL
    # and against the reference itself at its default tolerances (committed run), within the
    # reference's own quadrature error
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.json")))["corr"]
    if name in gold:
        assert w_err(w, gold[name]["w"]) < 2e-3


def test_halo_exclusion_against_oracle_and_reference_run():
    """HaloExclusion (halo.py:1201-1233): the mass window multiplies the h_m and h_g integrands."""
    import json
    import os
    from oracle.quadrature import Tight
    survey = _survey("power_gm", exclusion=True)
    eng = engine.Engine(survey)
    c = engine.pack_params([C_DICT], _lib.COSMO_KEYS)
    h = engine.pack_params([H_DICT], _lib.HALO_KEYS)
    g = engine.pack_params([HOD_DICT], _lib.HOD_ZHENG_KEYS)
    eng.mass_tables(c, h, [0.0])
    eng.halo_tables(h, g)
    se = O.SingleEpoch(0.0, C_DICT, O.precision(), Tight(40))
    ref = O.HaloExclusion(se, O.MassFunction(se, H_DICT), O.HODZheng(HOD_DICT), H_DICT)
    tabs = eng.table(_lib.T_HALO_NODES, 1).cpu().numpy()[0].reshape(5, -1)
    for i, nm in enumerate(("h_m", "pp_mm", "h_g", "pp_gm", "pp_gg")):
        want = ref.table(nm)[0]        # h_m, h_g oscillate through zero at high k: relative to the peak
        assert np.max(np.abs(tabs[i] - want))/np.max(np.abs(want)) < TOL_TABLE, nm
    k = np.logspace(-3, 2, 200)
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.json")))["halo"]["exclusion"]
    for spec in ("power_mm", "power_gm", "power_gg"):
        P = eng.power(1, _lib.POWER_SPEC[spec], k).cpu().numpy()[0]
        assert rel_err(P, ref.power(spec, k)) < TOL_FINAL, spec
        assert rel_err(P, gold[spec]) < 1e-3, spec        # the reference's own Romberg error
    from chomp_b200 import cosmology, halo, hod
    hx = halo.HaloExclusion(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cosmology.SingleEpoch(0.0, C_DICT),
                            halo_dict=H_DICT)
    assert rel_err(hx.power_gm(k), ref.power_gm(k)) < TOL_FINAL
