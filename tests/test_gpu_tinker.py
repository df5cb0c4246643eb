"""TinkerMassFunction (reference mass_function.py:436-564) through the C ABI: f(nu), b(nu), the bias
normalisation, and the halo-model spectra / w_gg(theta) of a Halo built on it -- against the oracle's converged values and
a committed run of the reference (tests/golden/reference_r2.json, section tinker)."""
import json
import os

import numpy as np
import pytest

from oracle import chomp_oracle as O
from oracle.quadrature import Tight

from common import C_DICT, D2R, H_DICT, H_DICT_2, HOD_DICT, rel_err, w_err

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_r2.json")))["tinker"]


@pytest.mark.parametrize("z,hd,key", [(0.0, H_DICT, "z0.0"), (0.5, H_DICT_2, "z0.5_delta_v_200")])
def test_multiplicity_and_bias(z, hd, key):
    from chomp_b200 import cosmology, mass_function
    g = GOLD[key]
    mf = mass_function.TinkerMassFunction(z, cosmology.SingleEpoch(z, cosmo_dict=C_DICT), hd)
    nu = np.array(g["nu"])
    omf = O.TinkerMassFunction(O.SingleEpoch(z, C_DICT, O.precision(), Tight(40)), hd)
    assert mf.delta_v == pytest.approx(g["delta_v"], rel=1e-12)
    assert rel_err(mf.f_nu(nu), omf.f_nu(nu)) < 1e-12 and rel_err(mf.f_nu(nu), g["f_nu"]) < 1e-12     # closed form
    assert mf.bias_norm == pytest.approx(omf.bias_norm, rel=1e-7) and mf.f_norm == 1.0
    assert rel_err(mf.bias_nu(nu), omf.bias_nu(nu)) < 1e-7
    assert mf.bias_norm == pytest.approx(g["bias_norm"], rel=1e-6)        # the reference: Romberg, mass_precision 1.48e-8
    assert rel_err(mf.bias_nu(nu), g["bias_nu"]) < 1e-6
    assert rel_err(mf._nu_array, g["nu_nodes"]) < 1e-6 and mf.m_star == pytest.approx(g["m_star"], rel=1e-6)


def test_halo_model_on_the_tinker_mass_function():
    from chomp_b200 import correlation, cosmology, halo, hod, kernel, mass_function
    cs = cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT)
    mf = mass_function.TinkerMassFunction(0.0, cs, H_DICT)
    h = halo.Halo(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cs, mass_func=mf, halo_dict=H_DICT)
    se = O.SingleEpoch(0.0, C_DICT, O.precision(), Tight(40))
    oh = O.Halo(se, O.TinkerMassFunction(se, H_DICT), O.HODZheng(HOD_DICT, O.precision()["halo_precision"]), H_DICT)
    k = np.array(GOLD["halo"]["k"])
    for spec, tol in (("power_mm", 5e-5), ("power_gm", 1e-3), ("power_gg", 1e-3)):
        got = getattr(h, spec)(k)
        assert rel_err(got, oh.power(spec, k)) < 1e-5, spec                # the parity bar
        assert rel_err(got, GOLD["halo"][spec]) < tol, spec                # the reference's own Romberg error
    cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    dist = kernel.dNdzGaussian(0.0, 2.0, 0.5, 0.1)
    kern = kernel.Kernel(1e-6*D2R, 100.0*D2R, kernel.WindowFunctionGalaxy(dist, cm), kernel.WindowFunctionGalaxy(dist, cm), cm)
    corr = correlation.Correlation(0.01, 1.0, kern, bins_per_decade=3.0, input_halo=h, power_spec="power_gg")
    corr.compute_correlation()
    assert w_err(corr.wtheta_array, GOLD["wtheta"]["w"]) < 5e-4
    # the oracle at z_bar
    sz = O.SingleEpoch(float(kern.z_bar), C_DICT, O.precision(), Tight(40))
    ohz = O.Halo(sz, O.TinkerMassFunction(sz, H_DICT), O.HODZheng(HOD_DICT, O.precision()["halo_precision"]), H_DICT)
    kz = np.logspace(-2.5, 1.5, 40)
    assert rel_err(h.power_gg(kz), ohz.power("power_gg", kz)) < 1e-5
    # correlation_batch carries the mass-function kind
    w, st = corr.correlation_batch([C_DICT], None, [HOD_DICT])
    assert not st.any() and np.array_equal(w[0], corr.wtheta_array)
