"""Pins the oracle on the round-2 boundary cases against committed runs of the reference itself
(tests/golden/reference_r2.json, made by tests/golden/make_golden_r2.py): Correlation(k_min=, k_max=)
including the Python-2 ``None < x`` outcome of correlation.py:104-107."""
import json
import os

import pytest

from oracle.quadrature import Romberg

import numpy as np

from common import C_DICT, H_DICT, HOD_DICT, oracle_covariance_cross, oracle_wtheta, w_err

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_r2.json")))


@pytest.mark.parametrize("key", ["wide_power_gg", "kmax_only_power_mm"])
def test_oracle_k_limits_match_reference_run(key):
    g = GOLD["k_limits"][key]
    spec = "power_gg" if key.endswith("gg") else "power_mm"
    r = oracle_wtheta(C_DICT, H_DICT, HOD_DICT, ("gaussian", (0.0, 2.0, 0.5, 0.1)), power_spec=spec, bins_per_decade=3.0,
                      theta_deg=(0.01, 1.0), integ=Romberg(), **g["args"])
    assert bool(r["halo"].extrapolate) == g["extrapolate"]
    assert w_err(r["w"], g["w"]) < 1e-11


def test_oracle_cross_covariance_matches_reference_run():
    """covariance.Covariance between two different correlations (covariance.py:60-63, 421-453, 455-591): the oracle
    with the reference's Romberg rule against the committed run of the reference."""
    g = GOLD["cross_cov"]
    cfg = g["config"]
    oc = oracle_covariance_cross(C_DICT, H_DICT, HOD_DICT, cfg["hod_b"], tuple(cfg["dist_a"]), tuple(cfg["dist_b"]),
                                 theta_deg=tuple(cfg["theta_deg"]), tri_z=cfg["tri_z"], area_deg2=cfg["area_deg2"], n_a=cfg["n_a"],
                                 n_b=cfg["n_b"], variance=cfg["variance"], integ=Romberg())
    assert oc.kernel.z_bar_NG == pytest.approx(g["z_bar_NG"], rel=1e-13)
    assert oc.equal_windows == g["equal_windows"] and oc.cosmic_shear == g["cosmic_shear"]
    oc.projected_table()
    for key, got in (("a", oc.proj_nodes), ("b", oc.proj_nodes_b), ("ab", oc.proj_nodes_ab), ("ba", oc.proj_nodes_ba)):
        ref = np.array(g["proj"][key])
        assert np.max(np.abs(got - ref)) < 1e-11*np.max(np.abs(ref)), key
    n = len(g["bins_center"])
    # three bin pairs of each term (the whole matrix takes the Romberg rule a minute)
    for i, j in ((0, 0), (2, 6), (9, 9)):
        a, b = oc.bins[i][2], oc.bins[j][2]
        assert oc.covariance_G(a, b) == pytest.approx(np.array(g["cov_G"]).reshape(n, n)[i, j], rel=1e-10)
        assert oc.covariance_NG(a, b) == pytest.approx(np.array(g["cov_NG"]).reshape(n, n)[i, j], rel=1e-10)


def test_oracle_bao_transfer_matches_reference_run():
    """SingleEpoch(with_bao=True), cosmology.py:474-538."""
    from oracle import chomp_oracle as O
    for z in (0.0, 0.5):
        g = GOLD["bao"]["z%.1f" % z]
        se = O.SingleEpoch(z, C_DICT, with_bao=True)
        assert np.max(np.abs(se.linear_power(np.array(g["k"]))/np.array(g["linear_power"]) - 1.0)) < 1e-13
        assert se.sigma_norm == pytest.approx(g["sigma_norm"], rel=1e-13)


def test_oracle_tinker_matches_reference_run():
    """TinkerMassFunction, mass_function.py:436-564."""
    from oracle import chomp_oracle as O
    from common import H_DICT_2
    for z, hd, key in ((0.0, H_DICT, "z0.0"), (0.5, H_DICT_2, "z0.5_delta_v_200")):
        g = GOLD["tinker"][key]
        mf = O.TinkerMassFunction(O.SingleEpoch(z, C_DICT), hd)
        nu = np.array(g["nu"])
        assert np.max(np.abs(mf.f_nu(nu)/np.array(g["f_nu"]) - 1.0)) < 1e-13
        assert np.max(np.abs(mf.bias_nu(nu)/np.array(g["bias_nu"]) - 1.0)) < 1e-12
        assert mf.bias_norm == pytest.approx(g["bias_norm"], rel=1e-12)


def test_oracle_ssc_halo_and_xi3d_match_reference_run():
    """HaloSuperSampleCovariance (halo.py:1089-1199) and Correlation3d (correlation.py:408-510)."""
    from oracle import chomp_oracle as O
    g = GOLD["ssc_halo"]["z0.5"]
    k = np.array(g["k"])
    se = O.SingleEpoch(0.5, C_DICT)
    h = O.HaloSuperSampleCovariance(se, O.MassFunction(se, H_DICT), O.HODZheng(HOD_DICT), H_DICT, delta_b=0.02)
    ref = np.array(g["dln_power_ddelta_b"])
    ok = ref != 0
    assert np.max(np.abs(h.dln_power_ddelta_b(k)[ok]/ref[ok] - 1.0)) < 1e-12 and np.all(h.dln_power_ddelta_b(k)[~ok] == 0)
    assert np.max(np.abs(h.power_mm_ssc(k)[ok]/np.array(g["power_mm_ssc"])[ok] - 1.0)) < 1e-12
    g3 = GOLD["xi3d"]["power_mm"]
    s3 = O.SingleEpoch(0.3, C_DICT)
    h3 = O.Halo(s3, O.MassFunction(s3, H_DICT), O.HODZheng(HOD_DICT), H_DICT)
    sel = [0, 17, 49]
    xi = O.Correlation3d(0.05, 60.0, h3, "power_mm").raw_correlation(np.array(g3["r"])[sel])
    assert np.max(np.abs(xi/np.array(g3["xi"])[sel] - 1.0)) < 1e-11
