"""BASELINE config 4: shear-shear spectrum via HaloFit (Takahashi 12) Limber projection with a
magnitude-limited dN/dz."""
import json
import os

import numpy as np
import pytest

from chomp_b200 import _lib, engine
from oracle import chomp_oracle as O
from oracle.quadrature import Tight

from common import C_DICT, D2R, H_DICT, HOD_DICT, rel_err

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.json")))["cfg4_halofit"]


def _oracle(integ, fit_z=0.0, n_halo=50):
    prec = O.precision(halo_npoints=n_halo)
    cm = O.MultiEpoch(0.0, 5.0, C_DICT, prec, integ)
    win = O.WindowFunctionConvergence(O.dNdzMagLim(0.0, 2.0, 2.0, 0.5, 2.0, prec, integ), cm)
    kern = O.Kernel(1e-6*D2R, 100.0*D2R, win, win, cm)
    fit_epoch = O.SingleEpoch(fit_z, C_DICT, prec, integ)

    def factory(z):
        se = O.SingleEpoch(z, C_DICT, prec, integ)
        return O.HaloFit(se, O.MassFunction(se, H_DICT), O.HODZheng(HOD_DICT), H_DICT, fit_epoch=fit_epoch)
    return O.CorrelationFourier(kern, factory, "power_mm")


@pytest.mark.parametrize("n_halo", [50, 200])
def test_halofit_parameters_spectra_and_cl_against_oracle(n_halo):
    from chomp_b200 import defaults
    ref = _oracle(Tight(40), n_halo=n_halo)
    dist = engine.RedshiftDistribution.maglim(0.0, 2.0, 2.0, 0.5, 2.0)
    survey = engine.Survey(dist, window_a="convergence", power_spec="power_mm",
                           precision=dict(defaults.default_precision, halo_npoints=n_halo))
    cfg = survey.config()
    cfg.use_halofit = 1
    eng = engine.Engine(cfg)
    c = engine.pack_params([C_DICT], _lib.COSMO_KEYS)
    h = engine.pack_params([H_DICT], _lib.HALO_KEYS)
    g = engine.pack_params([HOD_DICT], _lib.HOD_ZHENG_KEYS)
    eng.limber_tables(c)
    eng.mass_tables(c, h)                     # at z_bar
    eng.halo_tables(h, g)
    fit = dict(zip(_lib.HALOFIT_FIELDS, eng.halofit(1, fit_z=0.0).cpu().numpy()[0]))
    rfit = ref.halo._fit_params()
    assert fit["k_s"] == pytest.approx(rfit["k_s"], rel=1e-9)
    assert fit["n_eff"] == pytest.approx(rfit["n_eff"], rel=1e-8)       # quintic-spline derivative
    assert fit["C"] == pytest.approx(rfit["C"], rel=1e-7)               # quintic-spline second derivative
    for name in ("a_n", "b_n", "c_n", "gamma_n", "alpha_n", "beta_n", "nu_n"):
        assert fit[name] == pytest.approx(rfit[name], rel=1e-7), name
    assert fit["f_1"] == pytest.approx(ref.halo.f1, rel=1e-13)
    k = np.logspace(-4, 3, 120)               # HaloFit power_mm has no k-range guards
    for spec in ("power_mm", "power_gm", "power_gg"):
        got = eng.power(1, _lib.POWER_SPEC[spec], k).cpu().numpy()[0]
        want = getattr(ref.halo, spec)(k)
        assert np.all((got == 0) == (want == 0)), spec
        nz = want != 0
        assert rel_err(got[nz], want[nz]) < 1e-5, spec
    ell = np.array(GOLD["ell"])
    cl = eng.cl(1, _lib.P_MM, ell).cpu().numpy()[0]
    assert rel_err(cl, ref.correlation(ell)) < 1e-5
    ref.power_spec = "linear_power"
    assert rel_err(eng.cl(1, _lib.P_LINEAR, ell).cpu().numpy()[0], ref.correlation(ell)) < 1e-5
    # w(theta) with the HALOFIT spectrum in the Hankel stage
    if n_halo == 50:
        theta = engine.theta_bins(0.001, 1.0, 5.0)
        w = eng.wtheta_stage(1, _lib.P_MM, theta).cpu().numpy()[0]
        oc = O.Correlation(0.001, 1.0, ref.kernel, lambda z: ref.halo, "power_mm", bins_per_decade=5.0, integ=Tight(40))
        assert rel_err(w, oc.compute_correlation()) < 1e-5


def test_facade_shear_shear_like_the_example_script():
    """examples/shear_shear_spectrum.py with a magnitude-limited dN/dz, against the committed run
    of the reference (its Romberg error on sigma^2(R) moves C by 2e-5, C(l) by ~1e-5)."""
    from chomp_b200 import correlation, cosmology, halo, hod, kernel
    cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    win = kernel.WindowFunctionConvergence(kernel.dNdzMagLim(0.0, 2.0, 2.0, 0.5, 2.0), cm)
    kern = kernel.Kernel(1e-6*D2R, 100.0*D2R, win, win, cm)
    hf = halo.HaloFit(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cosmology.SingleEpoch(0.0, C_DICT),
                      halo_dict=H_DICT)
    cf = correlation.CorrelationFourier(10, 1e5, kern, input_halo=hf, powSpec="power_mm")
    assert kern.z_bar == pytest.approx(GOLD["z_bar"], abs=1e-12)
    ell = np.array(GOLD["ell"])
    assert rel_err(cf.correlation(ell), GOLD["cl"]) < 1e-4
    k = np.array(GOLD["k"])
    assert rel_err(hf.power_mm(k), GOLD["power_mm"]) < 1e-4
    assert hf._k_s == pytest.approx(GOLD["fit"]["k_s"], rel=1e-6)
    assert hf._n_eff == pytest.approx(GOLD["fit"]["n_eff"], rel=1e-5)
    cf.set_power_spectrum("linear_power")
    assert rel_err(cf.correlation(ell), GOLD["cl_linear"]) < 1e-5
    cf.compute_correlation()
    assert cf.power_array.shape == cf.l_array.shape == (50,)
