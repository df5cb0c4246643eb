"""GPU parity for dNdzInterpolation (kernel.py:181-208): the tabulated redshift distribution
evaluated on the device from the piecewise-polynomial form of the reference's FITPACK spline,
against the oracle's converged evaluation and against committed runs of the reference."""
import json
import os

import numpy as np
import pytest

from chomp_b200 import _lib, engine
from chomp_b200 import correlation as C, cosmology, halo as H, hod, kernel as K

from common import C_DICT, D2R, H_DICT, HOD_DICT, interp_table, oracle_wtheta, rel_err, w_err

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_interp.json")))


def _point():
    c = engine.pack_params([C_DICT], _lib.COSMO_KEYS)
    h = engine.pack_params([H_DICT], _lib.HALO_KEYS)
    g = np.zeros((1, _lib.N_HOD))
    g[0, :len(_lib.HOD_ZHENG_KEYS)] = [HOD_DICT[k] for k in _lib.HOD_ZHENG_KEYS]
    return c, h, g


@pytest.mark.parametrize("window,tol_win,tol_w", [("galaxy", 1e-9, 1e-5), ("convergence", 2e-6, 1e-5)])
def test_batched_wtheta_with_a_tabulated_dndz(window, tol_win, tol_w):
    import torch
    z, p = interp_table()
    dist = engine.RedshiftDistribution.table(z, p)
    survey = engine.Survey(dist, window_a=window, bins_per_decade=10.0, power_spec="power_mm")
    eng = engine.Engine(survey)
    c, h, g = _point()
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    w = eng.wtheta(c, h, g, survey.theta, _lib.P_MM, status=status).cpu().numpy()[0]
    assert int(status.cpu()[0]) == 0
    ref = oracle_wtheta(C_DICT, H_DICT, HOD_DICT, ("table", (z, p)), window_a=window, power_spec="power_mm")
    nw = survey.precision["window_npoints"]
    win = eng.table(_lib.T_WINDOW_NODES, 1).cpu().numpy()[0].reshape(2, nw)
    assert np.max(np.abs(win[0] - ref["wa_nodes"]))/np.max(np.abs(ref["wa_nodes"])) < tol_win
    assert abs(eng.table(_lib.T_ZBAR, 1).cpu().numpy()[0, 0] - ref["z_bar"]) < 1e-12
    kn = eng.table(_lib.T_KERNEL_NODES, 1).cpu().numpy()[0]
    assert np.max(np.abs(kn - ref["kernel_nodes"]))/np.max(np.abs(ref["kernel_nodes"])) < max(tol_win, 1e-7)
    assert w_err(w, ref["w"]) < tol_w
    # and the committed run of the reference itself, within its own quadrature tolerances
    gold = GOLD[window]
    assert abs(eng.table(_lib.T_ZBAR, 1).cpu().numpy()[0, 0] - gold["z_bar"]) < 1e-12
    assert w_err(w, gold["w"]) < 5e-4


@pytest.mark.parametrize("order", [1, 2, 3])
def test_dropin_class(order):
    z, p = interp_table()
    g = GOLD["order%d" % order]
    d = K.dNdzInterpolation(z, p, interpolation_order=order)
    assert d.z_min == z[0] and d.z_max == z[-1]
    assert d.norm == pytest.approx(g["norm"], rel=5e-7)          # exact integral vs the reference's Romberg
    zz = np.array(g["z"])
    assert np.allclose(d.raw_dndz(zz), g["raw"], rtol=1e-12, atol=1e-15)
    assert np.allclose(d.dndz(zz), g["dndz"], rtol=5e-7, atol=1e-12)
    assert isinstance(d.dndz(0.5), float)


def test_dropin_correlation_with_interpolated_dndz():
    z, p = interp_table()
    gold = GOLD["galaxy"]
    cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    d = K.dNdzInterpolation(z, p)
    wa, wb = K.WindowFunctionGalaxy(d, cm), K.WindowFunctionGalaxy(d, cm)
    kern = K.Kernel(1e-6*D2R, 100.0*D2R, wa, wb, cm)
    assert kern.z_bar == pytest.approx(gold["z_bar"], abs=1e-12)
    kn = np.array(gold["kernel_nodes"])
    assert np.max(np.abs(kern._kernel_array - kn))/np.max(np.abs(kn)) < 1e-5
    h = H.Halo(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT),
               halo_dict=H_DICT)
    corr = C.Correlation(0.001, 1.0, kern, bins_per_decade=10, input_halo=h, power_spec="power_mm")
    corr.compute_correlation()
    assert w_err(corr.wtheta_array, gold["w"]) < 5e-4
