"""SingleEpoch(with_bao=True): the Eisenstein & Hu (1998) transfer function with baryon wiggles (reference
cosmology.py:474-538) through the C ABI -- linear power, sigma(R), the mass function and halo-model
spectra built on it, w_gg(theta) -- against the oracle's converged values and a committed run of the reference
(tests/golden/reference_r2.json, section bao), including the reference's reset of the flag whenever a cosmology object
is re-initialised (Q16: Correlation moving the halo to z_bar drops the wiggles)."""
import json
import os

import numpy as np
import pytest

from oracle import chomp_oracle as O
from oracle.quadrature import Tight

from common import C_DICT, D2R, H_DICT, HOD_DICT, oracle_wtheta, rel_err, w_err

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_r2.json")))["bao"]


@pytest.mark.parametrize("z", [0.0, 0.5])
def test_linear_power_and_sigma(z):
    from chomp_b200 import cosmology
    g = GOLD["z%.1f" % z]
    cs = cosmology.SingleEpoch(z, cosmo_dict=C_DICT, with_bao=True)
    k = np.array(g["k"])
    se = O.SingleEpoch(z, C_DICT, O.precision(), Tight(80), with_bao=True)     # order 40 is not converged on the wiggles (sigma(2 Mpc/h): 3.6e-7)
    assert rel_err(cs.linear_power(k), se.linear_power(k)) < 1e-8
    assert rel_err(cs.linear_power(k), g["linear_power"]) < 1e-7          # the reference: Romberg rtol 1.48e-8 in sigma_8
    assert cs._sigma_norm == pytest.approx(se.sigma_norm, rel=1e-8)
    assert rel_err([cs.sigma_r(r) for r in (0.5, 2.0, 8.0, 30.0)], g["sigma_r"]) < 1e-7
    plain = cosmology.SingleEpoch(z, cosmo_dict=C_DICT)
    assert rel_err(plain.linear_power(k), g["linear_power"]) > 1e-2          # the wiggles are really there
    cs.set_cosmology(C_DICT)                                              # cosmology.py:151: with_bao is not carried over
    assert rel_err(cs.linear_power(k), plain.linear_power(k)) == 0.0


def test_mass_function_and_halo_spectra():
    from chomp_b200 import cosmology, halo, hod, mass_function
    cs = cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT, with_bao=True)
    mf = mass_function.MassFunction(0.0, cs, H_DICT)
    g = GOLD["mass"]
    se = O.SingleEpoch(0.0, C_DICT, O.precision(), Tight(80), with_bao=True)
    omf = O.MassFunction(se, H_DICT)
    assert rel_err(mf._ln_mass_array, omf.ln_mass_nodes) < 1e-12 and rel_err(mf._nu_array, omf.nu_nodes) < 3e-7     # measured 1.1e-7
    assert rel_err(mf._nu_array, g["nu_nodes"]) < 1e-6 and rel_err(mf._ln_mass_array, g["ln_mass_nodes"]) < 1e-12
    assert mf.f_norm == pytest.approx(g["f_norm"], rel=1e-6) and mf.bias_norm == pytest.approx(g["bias_norm"], rel=1e-6)
    h = halo.Halo(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cs, halo_dict=H_DICT)
    oh = O.Halo(se, omf, O.HODZheng(HOD_DICT, O.precision()["halo_precision"]), H_DICT)
    k = np.array(GOLD["halo"]["k"])
    for spec, tol in (("power_mm", 2e-5), ("power_gm", 1e-3), ("power_gg", 1e-3)):
        got = getattr(h, spec)(k)
        assert rel_err(got, oh.power(spec, k)) < 1e-5, spec                # the parity bar
        assert rel_err(got, GOLD["halo"][spec]) < tol, spec                # the reference's own Romberg error


def test_wtheta_keeps_or_drops_the_wiggles_as_the_reference_does():
    from chomp_b200 import correlation, cosmology, halo, hod, kernel
    cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    dist = kernel.dNdzGaussian(0.0, 2.0, 0.5, 0.1)
    kern = kernel.Kernel(1e-6*D2R, 100.0*D2R, kernel.WindowFunctionGalaxy(dist, cm), kernel.WindowFunctionGalaxy(dist, cm), cm)
    zb = float(kern.z_bar)
    hz = halo.Halo(redshift=zb, input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cosmology.SingleEpoch(zb, cosmo_dict=C_DICT, with_bao=True),
                   halo_dict=H_DICT)
    corr = correlation.Correlation(0.01, 1.0, kern, bins_per_decade=3.0, input_halo=hz, power_spec="power_gg", keep_halo_z_bar=True)
    corr.compute_correlation()
    g = GOLD["wtheta_keep_z_bar"]
    assert hz.cosmo._with_bao == g["with_bao_after"]
    ref = oracle_wtheta(C_DICT, H_DICT, HOD_DICT, ("gaussian", (0.0, 2.0, 0.5, 0.1)), power_spec="power_gg", bins_per_decade=3.0,
                        theta_deg=(0.01, 1.0), with_bao=True, integ=Tight(80))
    assert w_err(corr.wtheta_array, ref["w"]) < 1e-5
    assert w_err(corr.wtheta_array, g["w"]) < 5e-4
    # the default: Correlation moves the halo to z_bar, which re-initialises its cosmology without the flag (Q16)
    h2 = halo.Halo(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT, with_bao=True),
                   halo_dict=H_DICT)
    corr2 = correlation.Correlation(0.01, 1.0, kern, bins_per_decade=3.0, input_halo=h2, power_spec="power_gg")
    corr2.compute_correlation()
    g2 = GOLD["wtheta_moved"]
    assert bool(h2.cosmo._with_bao) == g2["with_bao_after"]
    assert w_err(corr2.wtheta_array, g2["w"]) < 5e-4
    assert w_err(corr2.wtheta_array, g["w"]) > 1e-3
