"""Edge cases and error behaviour of the C ABI / batch interface: ragged batch sizes, single
points, per-point status flags instead of aborts (or the reference's infinite loops), call-level
errors for bad arguments, determinism."""
import ctypes

import numpy as np
import pytest

from chomp_b200 import ChompError, _lib, design, engine

pytestmark = pytest.mark.gpu


def _survey(**kw):
    return engine.Survey(engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1), bins_per_decade=10.0,
                         power_spec="power_gg", **kw)


def test_ragged_batch_sizes_and_determinism():
    import torch
    survey = _survey()
    eng = engine.Engine(survey)
    cosmo, halo, hod = design.synthetic_batch(37)
    full = eng.wtheta(cosmo, halo, hod, survey.theta, _lib.P_GG).cpu().numpy()
    assert full.shape == (37, 30) and np.all(np.isfinite(full))
    for n in (1, 2, 31, 33):
        part = eng.wtheta(cosmo[:n], halo[:n], hod[:n], survey.theta, _lib.P_GG).cpu().numpy()
        assert np.array_equal(part, full[:n]), n
    again = eng.wtheta(cosmo, halo, hod, survey.theta, _lib.P_GG).cpu().numpy()
    assert np.array_equal(again, full)                         # bit-reproducible
    one_theta = eng.wtheta(cosmo, halo, hod, survey.theta[7:8], _lib.P_GG).cpu().numpy()
    assert np.array_equal(one_theta[:, 0], full[:, 7])
    # a second handle (another "rank") gives the same bits
    eng2 = engine.Engine(survey)
    assert np.array_equal(eng2.wtheta(cosmo[20:], halo[20:], hod[20:], survey.theta, _lib.P_GG).cpu().numpy(), full[20:])
    torch.cuda.synchronize()


def test_status_flags_instead_of_aborts():
    import torch
    survey = _survey()
    eng = engine.Engine(survey)
    cosmo, halo, hod = design.synthetic_batch(6)
    cosmo[1, 8] = -0.9            # w0 != -1: outside the supported domain
    halo[2, 4] = -1.5             # non-NFW profile slope
    cosmo[3, 6] = np.nan          # sigma_8 = NaN
    status = torch.zeros(6, dtype=torch.int32, device="cuda")
    w = eng.wtheta(cosmo, halo, hod, survey.theta, _lib.P_GG, status=status).cpu().numpy()
    st = status.cpu().numpy()
    assert st[0] == 0 and st[4] == 0 and st[5] == 0
    assert st[1] & _lib.ST_DOMAIN and st[2] & _lib.ST_DOMAIN
    assert st[3] & _lib.ST_NONFINITE and not np.all(np.isfinite(w[3]))
    assert np.all(np.isfinite(w[[0, 4, 5]]))
    # the good points are unaffected by their neighbours
    cosmo2, halo2, hod2 = design.synthetic_batch(6)
    ref = eng.wtheta(cosmo2, halo2, hod2, survey.theta, _lib.P_GG).cpu().numpy()
    assert np.array_equal(w[[0, 4, 5]], ref[[0, 4, 5]])


def test_unreachable_mass_limit_is_flagged_not_an_infinite_loop():
    """At z >~ 1.9 nu(M) = 0.1 cannot be reached: the reference's while-loop never ends
    (mass_function.py:172-194); the kernel sets CHOMP_ST_MASS_WALK."""
    import torch
    eng = engine.Engine(_survey())
    cosmo, halo, _ = design.synthetic_batch(2)
    status = torch.zeros(2, dtype=torch.int32, device="cuda")
    eng.mass_tables(cosmo, halo, z=[0.5, 3.0], status=status)
    st = status.cpu().numpy()
    assert st[0] == 0 and st[1] & _lib.ST_MASS_WALK


def test_call_level_errors():
    survey = _survey()
    eng = engine.Engine(survey)
    cosmo, halo, hod = design.synthetic_batch(4)
    with pytest.raises(ValueError):
        eng.wtheta(cosmo[:, :9], halo, hod, survey.theta, _lib.P_GG)        # wrong column count
    with pytest.raises(ChompError):
        eng.power(4, 7, [0.1])                                             # unknown spectrum id
    bad = survey.config()
    bad.n_halo = 3
    with pytest.raises(ChompError):
        engine.Engine(bad)
    bad = survey.config()
    bad.nq_nu = 99
    with pytest.raises(ChompError):
        engine.Engine(bad)
    bad = survey.config()
    bad.corr_k_min, bad.corr_k_max = 1.0, 0.5                              # Correlation(k_min > k_max)
    with pytest.raises(ChompError, match="k_max must exceed k_min"):
        engine.Engine(bad)
    with pytest.raises(ChompError):
        eng.trispectrum_1h(4)                                              # tri_moment < 0: list not built
    lib = _lib.load()
    assert lib.chomp_b200_reserve(None, 4) != 0 and lib.chomp_b200_last_error()
    h = ctypes.c_void_p()
    assert lib.chomp_b200_create(ctypes.byref(h), 9999) != 0                # no such device


def test_large_batch_and_reserve_growth():
    survey = _survey()
    eng = engine.Engine(survey)
    cosmo, halo, hod = design.synthetic_batch(3000)
    small = eng.wtheta(cosmo[:8], halo[:8], hod[:8], survey.theta, _lib.P_GG).cpu().numpy()
    big = eng.wtheta(cosmo, halo, hod, survey.theta, _lib.P_GG).cpu().numpy()      # forces a re-reserve
    assert big.shape == (3000, 30) and np.all(np.isfinite(big))
    assert np.array_equal(big[:8], small)
    wh, st = eng.wtheta_host(cosmo, halo, hod, survey.theta, _lib.P_GG)
    assert np.array_equal(wh, big) and not st.any()


def test_interleaved_engines_with_different_surveys():
    """Several live handles with different set-ups on one device (a multi-probe MCMC): the dynamic
    shared-memory opt-in of a kernel is state of the function, shared by every handle, and must never be
    lowered by the handle configured last (two different windows need more than one shared window;
    halo_npoints = 200 needs more than 50)."""
    from chomp_b200 import defaults
    lens = engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.4, 0.1)
    src = engine.RedshiftDistribution.gaussian(0.0, 2.0, 1.0, 0.2)
    big = engine.Survey(lens, src, window_a="galaxy", window_b="convergence", bins_per_decade=10.0, power_spec="power_gm",
                        precision=dict(defaults.default_precision, halo_npoints=200))
    small = _survey()
    cosmo, halo, hod = design.synthetic_batch(5)
    eng_a = engine.Engine(big)
    first = eng_a.wtheta(cosmo, halo, hod, big.theta, _lib.P_GM).cpu().numpy()
    eng_b = engine.Engine(small)                                   # configured after A, smaller everywhere
    wb = eng_b.wtheta(cosmo, halo, hod, small.theta, _lib.P_GG).cpu().numpy()
    again = eng_a.wtheta(cosmo, halo, hod, big.theta, _lib.P_GM).cpu().numpy()
    assert np.all(np.isfinite(first)) and np.all(np.isfinite(wb))
    assert np.array_equal(first, again)


def test_stage_order_is_enforced():
    """A later stage refuses rows no earlier stage computed on this handle (instead of reading zeros)."""
    survey = _survey()
    eng = engine.Engine(survey)
    cosmo, halo, hod = design.synthetic_batch(4)
    with pytest.raises(ChompError):
        eng.halo_tables(halo, hod)                                  # no mass tables yet
    eng.limber_tables(cosmo)
    eng.mass_tables(cosmo, halo)
    with pytest.raises(ChompError):
        eng.wtheta_stage(4, _lib.P_GG, survey.theta)                # halo tables missing
    eng.halo_tables(halo, hod)
    w4 = eng.wtheta_stage(4, _lib.P_GG, survey.theta).cpu().numpy()
    assert np.all(np.isfinite(w4))
    with pytest.raises(ChompError):
        eng.wtheta_stage(9, _lib.P_GG, survey.theta)                # growing the scratch drops the tables
    with pytest.raises(ChompError):
        eng.evaluate(_lib.EVAL_KERNEL, [0.0], point=0)
