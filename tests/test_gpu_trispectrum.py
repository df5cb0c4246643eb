"""1-halo trispectrum: the DMMA Gram kernel against the oracle's
converged integrals and against the committed run of the reference."""
import json
import os

import numpy as np
import pytest

from chomp_b200 import _lib, engine
from oracle import chomp_oracle as O
from oracle.quadrature import Tight

from common import C_DICT, H_DICT, HOD_DICT, rel_err

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.json")))


def _engine_table(spec, n_halo=50):
    from chomp_b200 import defaults
    prec = dict(defaults.default_precision, halo_npoints=n_halo)
    survey = engine.Survey(engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1), precision=prec)
    cfg = survey.config()
    cfg.tri_moment = _lib.TRISPECTRUM_MOMENT[spec]
    eng = engine.Engine(cfg)
    c = engine.pack_params([C_DICT], _lib.COSMO_KEYS)
    h = engine.pack_params([H_DICT], _lib.HALO_KEYS)
    g = engine.pack_params([HOD_DICT], _lib.HOD_ZHENG_KEYS)
    eng.mass_tables(c, h, [0.0])
    eng.halo_tables(h, g)
    return eng, eng.trispectrum_1h(1).cpu().numpy()[0]


@pytest.mark.parametrize("spec", ["power_mmmm", "power_gmmm", "power_ggmm", "power_gggg"])
def test_table_against_oracle(spec):
    eng, T = _engine_table(spec)
    assert np.array_equal(T, T.T)
    se = O.SingleEpoch(0.0, C_DICT, O.precision(), Tight(40))
    tri = O.HaloTrispectrumOneHalo(se, O.MassFunction(se, H_DICT), O.HODZheng(HOD_DICT), H_DICT, power_spec=spec)
    x = tri.ln_k_nodes
    for i in (0, 11, 25, 37, 44, 49):
        ref = np.array([tri.i_0_4(x[i], xj) for xj in x])
        assert rel_err(T[i], ref) < 1e-5, (spec, i)


def test_table_and_interpolation_against_reference_run():
    for spec in ("power_mmmm", "power_ggmm"):
        g = GOLD["trispectrum"][spec]
        eng, T = _engine_table(spec)
        ref = np.array(g["table"]).reshape(T.shape)
        # the reference's Romberg (halo_precision 1.48e-5) is converged to ~1e-5 on these smooth integrands
        assert rel_err(T, ref) < 2e-4
        got = eng.trispectrum_eval(g["k1"], g["k2"]).cpu().numpy()
        want = np.array(g["parallelogram"])
        assert np.all((got == 0) == (want == 0))
        nz = want != 0
        assert rel_err(got[nz], want[nz]) < 5e-4


def test_bicubic_interpolation_is_the_tensor_not_a_knot_spline():
    from scipy.interpolate import RectBivariateSpline
    eng, T = _engine_table("power_mmmm")
    x = np.linspace(np.log(1e-3), np.log(1e2), T.shape[0])
    sp = RectBivariateSpline(x, x, T, kx=3, ky=3, s=0)
    rng = np.random.default_rng(3)
    k1, k2 = np.exp(rng.uniform(x[0], x[-1], 64)), np.exp(rng.uniform(x[0], x[-1], 64))
    got = eng.trispectrum_eval(k1, k2).cpu().numpy()
    want = sp(np.log(k1), np.log(k2), grid=False)
    assert np.max(np.abs(got - want))/np.max(np.abs(want)) < 1e-11


def test_facade_and_named_shape():
    from chomp_b200 import cosmology, halo_trispectrum, hod, mass_function
    cs = cosmology.SingleEpoch(0.0, C_DICT)
    tri = halo_trispectrum.HaloTrispectrumOneHalo(0.0, cs, mass_function.MassFunction(0.0, cs, H_DICT), None,
                                                  H_DICT, hod.HODZheng(HOD_DICT), "power_mmmm")
    g = GOLD["trispectrum"]["power_mmmm"]
    got = tri.trispectrum_parallelogram(np.array(g["k1"]), np.array(g["k2"]))
    want = np.array(g["parallelogram"])
    nz = want != 0
    assert rel_err(got[nz], want[nz]) < 5e-4 and np.all(got[~nz] == 0)
    assert np.ndim(tri.trispectrum_parallelogram(0.1, 0.2)) == 0
    # halo_npoints = 200: 4 x 4 tiles of the Gram kernel, symmetric fill
    eng, T = _engine_table("power_mmmm", n_halo=200)
    assert T.shape == (200, 200) and np.array_equal(T, T.T) and np.all(T > 0)
    _, T50 = _engine_table("power_mmmm", n_halo=50)
    # ln k nodes 0 and 199 coincide with nodes 0 and 49 of the coarse grid
    assert T[0, 0] == pytest.approx(T50[0, 0], rel=1e-9)
    assert T[199, 0] == pytest.approx(T50[49, 0], rel=1e-9)
