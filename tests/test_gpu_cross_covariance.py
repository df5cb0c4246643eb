"""Covariance between two DIFFERENT correlations (reference covariance.py:60-63 ``matching_corrs`` False) and
CovarianceMulti (covariance.py:794-871) through the C ABI: K_NG with four windows, the four projected spectra, Gaussian
and non-Gaussian terms against the oracle's converged integrals (bar 1e-5) and against a committed run of the reference
(tests/golden/reference_r2.json, section cross_cov)."""
import json
import os

import numpy as np
import pytest

from chomp_b200 import _lib, defaults, engine
from oracle.quadrature import Tight

from common import C_DICT, D2R, H_DICT, HOD_DICT, oracle_covariance_cross, rel_err

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_r2.json"))).get("cross_cov")
needs_gold = pytest.mark.skipif(GOLD is None, reason="reference run of the cross-covariance not committed")


def block_err(got, ref, scale):
    return float(np.max(np.abs(np.asarray(got) - np.asarray(ref)))/scale)


def _engines(cfg):
    theta = tuple(cfg["theta_deg"])

    def make(dist, tri_moment=-1):
        survey = engine.Survey(engine.RedshiftDistribution.gaussian(*dist), theta_deg=theta, bins_per_decade=5.0, power_spec="power_gg")
        c = survey.config()
        c.tri_moment = tri_moment
        return survey, engine.Engine(c)
    sa, ea = make(cfg["dist_a"])
    sb, eb = make(cfg["dist_b"])
    st, et = make(cfg["dist_a"], _lib.TRISPECTRUM_MOMENT["power_gggg"])
    setup = engine.CovarianceSetup(sa, theta_deg=theta, bins_per_decade=5.0, survey_area_deg2=cfg["area_deg2"], n_a=cfg["n_a"],
                                   n_b=cfg["n_b"], variance=cfg["variance"], power_spec="power_gg")
    for i in range(6):
        setup.params.poisson[i] = 0.0
    return ea, eb, et, setup


@needs_gold
def test_cross_covariance_against_oracle_and_reference_run():
    import torch
    cfg = GOLD["config"]
    ea, eb, et, setup = _engines(cfg)
    pack = engine.pack_params
    c, h = pack([C_DICT], _lib.COSMO_KEYS), pack([H_DICT], _lib.HALO_KEYS)
    ga, gb = pack([HOD_DICT], _lib.HOD_ZHENG_KEYS), pack([cfg["hod_b"]], _lib.HOD_ZHENG_KEYS)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    out, parts = ea.covariance_cross(eb, c, h, ga, h, gb, setup, tri_engine=et, halo_t=h, hod_t=ga, tri_z=cfg["tri_z"],
                                     status=status, parts=True)
    out, parts = out.cpu().numpy()[0], parts.cpu().numpy()[0]
    assert int(status[0]) == 0
    n = out.shape[0]
    ocov = oracle_covariance_cross(C_DICT, H_DICT, HOD_DICT, cfg["hod_b"], tuple(cfg["dist_a"]), tuple(cfg["dist_b"]),
                                   theta_deg=tuple(cfg["theta_deg"]), tri_z=cfg["tri_z"], area_deg2=cfg["area_deg2"],
                                   n_a=cfg["n_a"], n_b=cfg["n_b"], variance=cfg["variance"], integ=Tight(16))
    total, P, G, NG = ocov.get_covariance(parts=True)
    assert not P.any() and not parts[0].any()                       # no Poisson term between different correlations
    scale = float(np.max(np.abs(total)))
    g_scale, ng_scale = float(np.max(np.abs(G))), float(np.max(np.abs(NG)))
    # K_NG with the four windows, z_bar_NG on the common range
    K = ea.table(_lib.T_KNG, 1).cpu().numpy().reshape(50, 50)
    ref_K = ocov.kernel.table()
    assert np.max(np.abs(K - ref_K)) < 1e-8*np.max(np.abs(ref_K))
    assert float(ea.table(_lib.T_ZBAR_NG, 1)[0, 0]) == pytest.approx(ocov.kernel.z_bar_NG, rel=1e-12)
    assert block_err(parts[1], G, g_scale) < 1e-5                    # the parity bar, every term on its own scale
    assert block_err(parts[2], NG, ng_scale) < 1e-5
    assert block_err(out, total, scale) < 1e-5
    assert np.array_equal(out, out.T)
    # the reference's own run (Romberg error as in the matching case, tests/test_gpu_covariance.py)
    gold_K = np.array(GOLD["kernel_NG_table"]).reshape(50, 50)
    # the reference's Romberg leaves -3.9e-20 in element (27, 49) of this table (converged: -4e-32; table maximum
    # 1.6e-16): 2.4e-4 of the maximum, measured with the oracle's Tight strategy
    assert np.max(np.abs(K - gold_K)) < 5e-4*np.max(np.abs(gold_K))
    assert float(ea.table(_lib.T_ZBAR_NG, 1)[0, 0]) == pytest.approx(GOLD["z_bar_NG"], rel=1e-12)
    gscale = float(np.max(np.abs(GOLD["cov"])))
    assert block_err(parts[1], np.array(GOLD["cov_G"]).reshape(n, n), float(np.max(np.abs(GOLD["cov_G"])))) < 5e-5
    assert block_err(parts[2], np.array(GOLD["cov_NG"]).reshape(n, n), float(np.max(np.abs(GOLD["cov_NG"])))) < 3e-4
    assert block_err(out, np.array(GOLD["cov"]).reshape(n, n), gscale) < 3e-4
    # identical correlations on two handles: the cross path reproduces the matching path's G and NG terms
    ea2, eb2, et2, setup2 = _engines(dict(cfg, dist_b=cfg["dist_a"]))
    same = ea2.covariance_cross(eb2, c, h, ga, h, ga, setup2, tri_engine=et2, tri_z=cfg["tri_z"], parts=True)[1].cpu().numpy()[0]
    survey = engine.Survey(engine.RedshiftDistribution.gaussian(*cfg["dist_a"]), theta_deg=tuple(cfg["theta_deg"]),
                           bins_per_decade=5.0, power_spec="power_gg")
    c1 = survey.config()
    c1.tri_moment = _lib.TRISPECTRUM_MOMENT["power_gggg"]
    e1 = engine.Engine(c1)
    setup1 = engine.CovarianceSetup(survey, theta_deg=tuple(cfg["theta_deg"]), bins_per_decade=5.0, survey_area_deg2=cfg["area_deg2"],
                                    n_a=cfg["n_a"], n_b=cfg["n_b"], variance=cfg["variance"], power_spec="power_gg")
    for i in range(6):
        setup1.params.poisson[i] = 0.0
    match = e1.covariance(c, h, ga, setup1, tri_z=cfg["tri_z"], parts=True)[1].cpu().numpy()[0]
    assert block_err(same[1], match[1], float(np.max(np.abs(match[1])))) < 1e-9
    assert block_err(same[2], match[2], float(np.max(np.abs(match[2])))) < 1e-9


@needs_gold
def test_drop_in_cross_covariance_and_covariance_multi():
    """covariance.Covariance(corr_a, corr_b) and CovarianceMulti built as the golden generator builds the reference's."""
    from chomp_b200 import correlation, cosmology, covariance, halo, halo_trispectrum, hod, kernel, mass_function
    cfg = GOLD["config"]

    def make_corr(dist, hod_dict):
        cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
        d = kernel.dNdzGaussian(*dist)
        kern = kernel.Kernel(1e-6*D2R, 100.0*D2R, kernel.WindowFunctionGalaxy(d, cm), kernel.WindowFunctionGalaxy(d, cm), cm)
        h = halo.Halo(input_hod=hod.HODZheng(hod_dict), cosmo_single_epoch=cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT), halo_dict=H_DICT)
        return correlation.Correlation(cfg["theta_deg"][0], cfg["theta_deg"][1], kern, bins_per_decade=5.0, input_halo=h,
                                       power_spec="power_gg")
    corr_a, corr_b = make_corr(cfg["dist_a"], HOD_DICT), make_corr(cfg["dist_b"], cfg["hod_b"])
    cs_t = cosmology.SingleEpoch(cfg["tri_z"], cosmo_dict=C_DICT)
    tri = halo_trispectrum.HaloTrispectrumOneHalo(cfg["tri_z"], cs_t, mass_function.MassFunction(cfg["tri_z"], cs_t, H_DICT), None,
                                                  H_DICT, hod.HODZheng(HOD_DICT), "power_gggg")
    cov = covariance.Covariance(corr_a, corr_b, bins_per_decade=5.0, survey_area_deg2=cfg["area_deg2"], n_a=cfg["n_a"],
                                n_b=cfg["n_b"], variance=cfg["variance"], nongaussian_cov=True, input_halo_trispectrum=tri,
                                power_spec="power_gg")
    assert not cov.matching_corrs
    total = cov.get_covariance()
    n = total.shape[0]
    assert cov.equal_windows == GOLD["equal_windows"] and cov.cosmic_shear == GOLD["cosmic_shear"]
    assert rel_err([b.center for b in cov.annular_bins], GOLD["bins_center"]) < 1e-14
    gscale = float(np.max(np.abs(GOLD["cov"])))
    assert block_err(total, np.array(GOLD["cov"]).reshape(n, n), gscale) < 3e-4
    multi = covariance.CovarianceMulti([corr_a, corr_b], bins_per_decade=5.0, survey_area_deg2=cfg["area_deg2"], n_a=cfg["n_a"],
                                       n_b=cfg["n_b"], variance=cfg["variance"], nongaussian_cov=False, input_halo_trispectrum=tri)
    w = multi.get_covariance()
    gold = np.array(GOLD["multi_gaussian"]).reshape(2*n, 2*n)
    assert w.shape == gold.shape and np.array_equal(w, w.T)
    assert block_err(w, gold, float(np.max(np.abs(gold)))) < 1e-4
