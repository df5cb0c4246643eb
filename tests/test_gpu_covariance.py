"""Covariance of w(theta) (BASELINE config 5) through the C ABI:
K_NG table, projected spectra, Poisson + Gaussian + non-Gaussian terms against the oracle's converged
integrals (bar 1e-5 relative to the diagonal) and against the committed run of the reference."""
import json
import os

import numpy as np
import pytest

from chomp_b200 import _lib, defaults, design, engine
from oracle.quadrature import Tight

from common import C_DICT, H_DICT, HOD_DICT, cov_err, oracle_covariance, rel_err

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_covariance.json")))
THETA = tuple(GOLD["theta_deg"])


def _setup(tri_spec="power_gggg", cov_spec="power_gg", n_halo=50, theta=THETA, bpd=5.0):
    prec = dict(defaults.default_precision, halo_npoints=n_halo)
    survey = engine.Survey(engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1), theta_deg=theta,
                           bins_per_decade=bpd, power_spec=cov_spec, precision=prec)
    cfg = survey.config()
    cfg.tri_moment = _lib.TRISPECTRUM_MOMENT[tri_spec]
    eng = engine.Engine(cfg)
    setup = engine.CovarianceSetup(survey, theta_deg=theta, bins_per_decade=bpd, survey_area_deg2=GOLD["area_deg2"],
                                   n_a=GOLD["n_a"], n_b=GOLD["n_b"], variance=GOLD["variance"], power_spec=cov_spec)
    return survey, eng, setup


def _params(cd=C_DICT, hd=H_DICT, gd=HOD_DICT):
    return (engine.pack_params([cd], _lib.COSMO_KEYS), engine.pack_params([hd], _lib.HALO_KEYS),
            engine.pack_params([gd], _lib.HOD_ZHENG_KEYS))


@pytest.fixture(scope="module")
def oracle_gggg():
    cov = oracle_covariance(C_DICT, H_DICT, HOD_DICT, theta_deg=THETA, tri_z=GOLD["tri_z"], integ=Tight(16))
    return cov, cov.get_covariance(parts=True)


def test_kernel_ng_table(oracle_gggg):
    cov, _ = oracle_gggg
    survey, eng, setup = _setup()
    c, h, g = _params()
    eng.limber_tables(c)
    K, zb, dz = eng.cov_kernel_ng(1, setup)
    K = K.cpu().numpy()[0]
    ref = cov.kernel.table()
    assert np.array_equal(K, K.T)
    assert np.max(np.abs(K - ref)) < 1e-8*np.max(np.abs(ref))
    assert float(zb[0]) == pytest.approx(cov.kernel.z_bar_NG, rel=1e-12)
    assert float(dz[0]) == pytest.approx(cov.D_z_NG, rel=1e-9)
    assert float(eng.table(_lib.T_KNG_MIN, 1)[0, 0]) == pytest.approx(cov.kernel.kernel_min, rel=1e-7)
    gold = np.array(GOLD["cases"]["power_gggg"]["kernel_NG_table"]).reshape(50, 50)
    assert np.max(np.abs(K - gold)) < 1e-7*np.max(np.abs(gold))


def test_covariance_terms_against_oracle(oracle_gggg):
    cov, (total, P, G, NG) = oracle_gggg
    survey, eng, setup = _setup()
    c, h, g = _params()
    import torch
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    out, parts = eng.covariance(c, h, g, setup, tri_z=GOLD["tri_z"], status=status, parts=True)
    out, parts = out.cpu().numpy()[0], parts.cpu().numpy()[0]
    assert int(status[0]) == 0
    assert np.allclose(setup.bins, cov.bins, rtol=1e-14, atol=0)
    proj = eng.table(_lib.T_PROJECTED, 1).cpu().numpy()[0]
    assert np.max(np.abs(proj - cov.proj_nodes)) < 1e-6*np.max(np.abs(cov.proj_nodes))
    assert rel_err(np.diag(parts[0]), np.diag(P)) < 1e-12 and np.count_nonzero(parts[0] - np.diag(np.diag(parts[0]))) == 0
    assert cov_err(parts[1], G) < 1e-5            # the parity bar
    assert cov_err(parts[2], NG) < 1e-5
    assert cov_err(out, total) < 1e-5
    assert np.array_equal(out, out.T)
    # the state left behind is that of the w(theta) path: w(theta) follows without recomputation
    w = eng.wtheta_stage(1, _lib.P_GG, survey.theta).cpu().numpy()[0]
    w_ref = eng.wtheta(c, h, g, survey.theta, _lib.P_GG).cpu().numpy()[0]
    assert np.array_equal(w, w_ref)


def test_covariance_against_reference_run():
    for spec, cov_spec in (("power_gggg", "power_gg"), ("power_mmmm", "power_mm")):
        gold = GOLD["cases"][spec]
        n = len(gold["bins_center"])
        survey, eng, setup = _setup(spec, cov_spec)
        c, h, g = _params()
        out, parts = eng.covariance(c, h, g, setup, tri_z=GOLD["tri_z"], parts=True)
        out, parts = out.cpu().numpy()[0], parts.cpu().numpy()[0]
        # within the reference's own Romberg error (measured with the oracle, tests/test_oracle_covariance.py)
        assert rel_err(np.diag(parts[0]), np.diag(np.array(gold["cov_P"]).reshape(n, n))) < 1e-12
        assert cov_err(parts[1], np.array(gold["cov_G"]).reshape(n, n)) < 5e-5
        assert cov_err(parts[2], np.array(gold["cov_NG"]).reshape(n, n)) < 3e-4
        assert cov_err(out, np.array(gold["cov"]).reshape(n, n)) < 3e-4


def test_design_point_finer_halo_table_and_default_trispectrum_redshift():
    """A synthetic-batch point, halo_npoints = 100 (the trispectrum grid no longer coincides with the
    ln k_a nodes), trispectrum at z_bar_NG (Covariance.set_cosmology, covariance.py:258), 10 bins/decade."""
    cosmo, halo, hod = design.synthetic_batch(3)
    cd, hd, gd = design.as_dicts(cosmo, halo, hod)[1]
    theta = (0.05, 1.0)
    from oracle import chomp_oracle as O
    ocov = oracle_covariance(cd, hd, gd, theta_deg=theta, bins_per_decade=10.0, tri_spec="power_ggmm",
                             prec=O.precision(halo_npoints=100), integ=Tight(16))
    total, P, G, NG = ocov.get_covariance(parts=True)
    survey, eng, setup = _setup("power_ggmm", "power_gg", n_halo=100, theta=theta, bpd=10.0)
    out, parts = eng.covariance(cosmo, halo, hod, setup, parts=True)
    out, parts = out.cpu().numpy(), parts.cpu().numpy()
    assert out.shape == (3, ocov.bins.shape[0], ocov.bins.shape[0])
    assert cov_err(parts[1, 1], G) < 1e-5
    assert cov_err(parts[1, 2], NG) < 1e-5
    assert cov_err(out[1], total) < 1e-5
    # a point's result does not depend on the batch it sits in
    one = eng.covariance(cosmo[1:2], halo[1:2], hod[1:2], setup).cpu().numpy()[0]
    assert np.array_equal(one, out[1])


def test_flags_and_errors():
    survey, eng, setup = _setup()
    c, h, g = _params()
    setup.params.poisson_only = 1
    out = eng.covariance(c, h, g, setup).cpu().numpy()[0]
    assert np.count_nonzero(out - np.diag(np.diag(out))) == 0 and np.all(np.diag(out) > 0)
    setup.params.poisson_only = 0
    setup.params.nongaussian = 0
    out_g, parts = eng.covariance(c, h, g, setup, parts=True)
    parts = parts.cpu().numpy()[0]
    assert np.all(parts[2] == 0) and np.allclose(out_g.cpu().numpy()[0], parts[0] + parts[1], rtol=1e-15)
    cfg = survey.config()                      # tri_moment = -1
    eng2 = engine.Engine(cfg)
    setup.params.nongaussian = 1
    from chomp_b200 import ChompError
    with pytest.raises(ChompError, match="trispectrum"):
        eng2.covariance(c, h, g, setup)


def test_drop_in_covariance_class():
    """covariance.Covariance built exactly as the golden generator builds the reference's
    (tests/golden/make_golden_cov.py), compared with that run."""
    from chomp_b200 import correlation, cosmology, covariance, halo, halo_trispectrum, hod, kernel, mass_function
    from common import D2R
    gold = GOLD["cases"]["power_gggg"]
    n = len(gold["bins_center"])
    cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    dist = kernel.dNdzGaussian(0.0, 2.0, 0.5, 0.1)
    wa, wb = kernel.WindowFunctionGalaxy(dist, cm), kernel.WindowFunctionGalaxy(dist, cm)
    kern = kernel.Kernel(1e-6*D2R, 100.0*D2R, wa, wb, cm)
    cs = cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT)
    h = halo.Halo(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cs, halo_dict=H_DICT)
    corr = correlation.Correlation(THETA[0], THETA[1], kern, bins_per_decade=5.0, input_halo=h, power_spec="power_gg")
    cs_t = cosmology.SingleEpoch(GOLD["tri_z"], cosmo_dict=C_DICT)
    tri = halo_trispectrum.HaloTrispectrumOneHalo(GOLD["tri_z"], cs_t, mass_function.MassFunction(GOLD["tri_z"], cs_t, H_DICT),
                                                  None, H_DICT, hod.HODZheng(HOD_DICT), "power_gggg")
    cov = covariance.Covariance(corr, corr, bins_per_decade=5.0, survey_area_deg2=GOLD["area_deg2"], n_a=GOLD["n_a"],
                                n_b=GOLD["n_b"], variance=GOLD["variance"], nongaussian_cov=True,
                                input_halo_trispectrum=tri, power_spec="power_gg")
    assert len(cov.annular_bins) == n
    assert rel_err([b.center for b in cov.annular_bins], gold["bins_center"]) < 1e-14
    assert rel_err([b.delta for b in cov.annular_bins], gold["bins_delta"]) < 1e-13
    total = cov.get_covariance()
    assert cov.equal_windows == gold["equal_windows"] and cov.cosmic_shear == gold["cosmic_shear"]
    assert cov_err(total, np.array(gold["cov"]).reshape(n, n)) < 3e-4
    a, b = cov.annular_bins[2], cov.annular_bins[6]
    ref_ng = np.array(gold["cov_NG"]).reshape(n, n)
    assert cov.covariance(a, b) == total[2, 6]
    assert cov.covariance_NG(a.center, b.center) == pytest.approx(ref_ng[2, 6], rel=2e-3)
    assert cov.covariance_P(a.delta, a.center) == pytest.approx(np.array(gold["cov_P"]).reshape(n, n)[2, 2], rel=1e-12)
    # two different correlations: tests/test_gpu_cross_covariance.py
    other = covariance.Covariance(corr, correlation.Correlation(THETA[0], THETA[1], kern, input_halo=h), input_halo_trispectrum=tri)
    assert not other.matching_corrs
