"""Oracle restatements of MassFunctionSecondOrder and of CorrelationFourier with the halo-model
spectra, pinned to runs of the reference (tests/golden/reference_extra.json)."""
import json
import os

import numpy as np
import pytest

from oracle import chomp_oracle as O
from oracle.quadrature import Romberg, Tight

from common import C_DICT, D2R, H_DICT, HOD_DICT, rel_err

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_extra.json")))


@pytest.mark.parametrize("z", [0.0, 0.5])
def test_second_order_mass_function(z):
    g = GOLD["second_order"]["z%.1f" % z]
    for integ, tol in ((Romberg(), 1e-9), (Tight(40), 1e-7)):
        se = O.SingleEpoch(z, C_DICT, O.precision(), integ)
        mf = O.MassFunctionSecondOrder(se, H_DICT)
        assert mf.bias_2_norm == pytest.approx(g["bias_2_norm"], rel=tol)
        assert rel_err(mf.sigma_nodes, g["sigma_nodes"]) < 1e-7
        assert rel_err(mf.bias_2_nu(np.array(g["nu"])), g["bias_2_nu"]) < max(tol, 1e-7)


def cl_oracle(spec, extrapolate, integ):
    prec = O.precision()
    cm = O.MultiEpoch(0.0, 5.0, C_DICT, prec, integ)
    d = O.dNdzGaussian(0.0, 2.0, 0.5, 0.1, prec=prec, integ=integ)
    kern = O.Kernel(1e-6*D2R, 100*D2R, O.WindowFunctionGalaxy(d, cm), O.WindowFunctionGalaxy(d, cm), cm)

    def factory(z):
        se = O.SingleEpoch(z, C_DICT, prec, integ)
        return O.Halo(se, O.MassFunction(se, H_DICT), O.HODZheng(HOD_DICT, prec["halo_precision"]), H_DICT,
                      extrapolate=extrapolate)
    return O.CorrelationFourier(kern, factory, spec)


def test_cl_with_halo_model_spectra_matches_reference_run():
    ell = np.array(GOLD["cl_tables"]["ell"])
    for extrapolate in (False, True):
        for spec in ("power_mm", "power_gg"):
            key = spec + ("_extrapolated" if extrapolate else "")
            cf = cl_oracle(spec, extrapolate, Romberg())
            assert rel_err(cf.correlation(ell), GOLD["cl_tables"][key]) < 1e-9, key
