"""MassFunctionSecondOrder and CorrelationFourier with the halo-model
(table-based) spectra (row a30) through the C ABI, against the oracle's converged values and the
committed runs of the reference."""
import json
import os

import numpy as np
import pytest

from chomp_b200 import _lib, engine
from oracle import chomp_oracle as O
from oracle.quadrature import Tight

from common import C_DICT, H_DICT, HOD_DICT, rel_err, w_err
from test_oracle_extra import cl_oracle

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_extra.json")))


@pytest.mark.parametrize("z", [0.0, 0.5])
def test_second_order_bias(z):
    g = GOLD["second_order"]["z%.1f" % z]
    survey = engine.Survey(engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1))
    eng = engine.Engine(survey)
    c = engine.pack_params([C_DICT], _lib.COSMO_KEYS)
    h = engine.pack_params([H_DICT], _lib.HALO_KEYS)
    eng.mass_tables(c, h, [z])
    b2n = float(eng.mass_second_order(1)[0])
    se = O.SingleEpoch(z, C_DICT, O.precision(), Tight(40))
    mf = O.MassFunctionSecondOrder(se, H_DICT)
    assert b2n == pytest.approx(mf.bias_2_norm, rel=1e-7)
    assert b2n == pytest.approx(g["bias_2_norm"], rel=1e-6)
    nu = np.array(g["nu"])
    got = eng.evaluate(_lib.EVAL_BIAS_2_NU, nu).cpu().numpy()
    assert rel_err(got, mf.bias_2_nu(nu)) < 1e-6
    assert rel_err(got, g["bias_2_nu"]) < 1e-6
    sig = eng.evaluate(_lib.EVAL_SIGMA_OF_NU, mf.nu_nodes[1:-1]).cpu().numpy()
    assert rel_err(sig, mf.sigma_nodes[1:-1]) < 1e-7


def test_second_order_drop_in_class():
    from chomp_b200 import cosmology, mass_function
    g = GOLD["second_order"]["z0.5"]
    cs = cosmology.SingleEpoch(0.5, cosmo_dict=C_DICT)
    mf = mass_function.MassFunctionSecondOrder(0.5, cs, H_DICT)
    assert mf.bias_2_norm == pytest.approx(g["bias_2_norm"], rel=1e-6)
    assert rel_err(mf.bias_2_mass(np.array(g["masses"])), g["bias_2_nu"]) < 1e-6
    assert np.ndim(mf.bias_2_nu(1.3)) == 0


def _against_gold(got, gold, key):
    gold = np.asarray(gold, dtype=float)
    ok = gold != 0.0
    assert w_err(got[ok], gold[ok]) < 3e-4, key
    assert np.all(np.abs(got[~ok]) < 1e-5*np.max(np.abs(gold))), key


@pytest.mark.parametrize("extrapolate", [False, True])
def test_cl_with_halo_model_spectra(extrapolate):
    ell = np.array(GOLD["cl_tables"]["ell"])
    survey = engine.Survey(engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1), extrapolate=extrapolate)
    eng = engine.Engine(survey)
    c = engine.pack_params([C_DICT], _lib.COSMO_KEYS)
    h = engine.pack_params([H_DICT], _lib.HALO_KEYS)
    g = engine.pack_params([HOD_DICT], _lib.HOD_ZHENG_KEYS)
    eng.wtheta(c, h, g, survey.theta, _lib.P_GG)              # stages 1-3 at z_bar
    for spec in ("power_mm", "power_gm", "power_gg"):
        key = spec + ("_extrapolated" if extrapolate else "")
        got = eng.cl(1, _lib.POWER_SPEC[spec], ell).cpu().numpy()[0]
        ref = cl_oracle(spec, extrapolate, Tight(24)).correlation(ell)
        assert w_err(got, ref, floor=1e-7) < 1e-5, key          # the parity bar
        # the reference's Romberg error; at l / k_max beyond the bulk of the window its Romberg stops at the
        # all-zero first levels and returns exactly 0 where the integral is ~1e-7 of the peak
        _against_gold(got, GOLD["cl_tables"][key], key)


def test_cl_drop_in_with_halo_spectrum():
    from chomp_b200 import correlation, cosmology, halo, hod, kernel
    from common import D2R
    cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    dist = kernel.dNdzGaussian(0.0, 2.0, 0.5, 0.1)
    kern = kernel.Kernel(1e-6*D2R, 100.0*D2R, kernel.WindowFunctionGalaxy(dist, cm), kernel.WindowFunctionGalaxy(dist, cm), cm)
    h = halo.Halo(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT),
                  halo_dict=H_DICT)
    cf = correlation.CorrelationFourier(10, 1e5, kern, input_halo=h, powSpec="power_gg")
    ell = np.array(GOLD["cl_tables"]["ell"])
    got = cf.correlation(ell)
    _against_gold(got, GOLD["cl_tables"]["power_gg"], "drop-in")
