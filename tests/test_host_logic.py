"""CPU-only checks of the host side: the C-ABI library loads and exports every
symbol include/chomp_b200.h declares, the ctypes mirror of the config struct has
the C layout, and the pieces of the reference's host logic that the product
restates (angular bins, redshift-distribution clipping, Latin-hypercube design,
sharding)."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np
import pytest

import chomp_b200
from chomp_b200 import _lib, design, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_outputs.json")))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "chomp_b200.h")).read()
    declared = set(re.findall(r"\b(chomp_b200_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(chomp_b200.EXPORTED_SYMBOLS)
    assert lib.chomp_b200_version() == 100


def test_config_struct_matches_c_layout(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "chomp_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu", sizeof(chomp_b200_config), '
                   'offsetof(chomp_b200_config, halo_precision), offsetof(chomp_b200_config, dndz_p), '
                   'offsetof(chomp_b200_config, bessel_limit));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, o1, o2, o3 = map(int, subprocess.check_output([str(exe)]).split())
    assert ctypes.sizeof(_lib.Config) == size
    assert _lib.Config.halo_precision.offset == o1
    assert _lib.Config.dndz_p.offset == o2
    assert _lib.Config.bessel_limit.offset == o3


def test_compute_entry_points_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(chomp_b200.ChompError):
        engine.Engine()
    h = ctypes.c_void_p()
    assert _lib.load().chomp_b200_create(ctypes.byref(h), 0) != 0
    assert _lib.load().chomp_b200_last_error()


@pytest.mark.parametrize("name,bpd", [("cfg1_mm", 5.0), ("cfg2_gg", 10.0)])
def test_theta_bins_match_reference(name, bpd):
    ref = np.array(GOLD["corr"][name]["theta"])
    got = engine.theta_bins(0.001, 1.0, bpd)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)
    assert len(engine.theta_bins(0.001, 1.0, 10.0)) == 30
    assert np.array_equal(engine.theta_bins(0.5, 0.5), [0.5*np.pi/180.0])


def test_redshift_distribution_clipping():
    g = engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1)       # kernel.py:100-104
    assert (g.z_min, g.z_max) == (0.0, 0.5 + 8*0.1)
    g = engine.RedshiftDistribution.gaussian(0.0, 2.0, 1.0, 0.05)
    assert (g.z_min, g.z_max) == (1.0 - 0.4, 1.0 + 0.4)
    # Python-2 integer division of 1/b (kernel.py:164-168): int b -> z_max capped at z0
    m = engine.RedshiftDistribution.maglim(0.0, 2.0, 2, 0.3, 2)
    assert m.z_max == pytest.approx(0.3)
    m = engine.RedshiftDistribution.maglim(0.0, 2.0, 2.0, 0.5, 2.0)
    assert m.z_max == pytest.approx(min(2.0, (-np.log(1.48e-8))**0.5*0.5))


def test_bessel_limits():
    from scipy import special
    for n in (1, 8, 20):
        assert engine.bessel_limit(0, n) == pytest.approx(special.jn_zeros(0, n)[-1], rel=1e-15)
        assert engine.bessel_limit(2, n) == pytest.approx(special.jn_zeros(2, n)[-1], rel=1e-15)
    with pytest.raises(ValueError):
        engine.bessel_limit(0, 99)


def test_survey_config_round_trip():
    s = engine.Survey(engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.4, 0.1),
                      engine.RedshiftDistribution.gaussian(0.0, 2.0, 1.0, 0.2), "galaxy", "convergence",
                      bessel_order=2, hod="mandelbaum", power_spec="power_gm")
    c = s.config()
    assert (c.window_kind[0], c.window_kind[1]) == (_lib.WINDOW_GALAXY, _lib.WINDOW_CONVERGENCE)
    assert c.bessel_order == 2 and c.hod_kind == _lib.HOD_MANDELBAUM
    assert c.bessel_limit == pytest.approx(27.420573549984557)
    assert c.n_window == 100 and c.n_halo == 50
    assert c.dndz_zmax[0] == pytest.approx(1.2) and c.dndz_p[1][0] == 1.0


def test_pack_params_raises_keyerror_like_the_reference():
    with pytest.raises(KeyError):
        engine.pack_params([{"omega_m0": 0.3}], _lib.COSMO_KEYS)


def test_latin_hypercube_design():
    rng = np.random.default_rng(1)
    u = design.latin_hypercube(64, 5, rng)
    for d in range(5):                                   # one sample per stratum (simulation_design.py:17-33)
        assert sorted(np.floor(u[:, d]*64).astype(int)) == list(range(64))
    cosmo, halo, hod = design.synthetic_batch(128)
    assert cosmo.shape == (128, 10) and halo.shape == (128, 6) and hod.shape == (128, 5)
    assert np.allclose(cosmo[:, 0] + cosmo[:, 2] + cosmo[:, 3], 1.0)      # flat (simulation_design.py:238-239)
    assert np.array_equal(hod[:, 0], hod[:, 2])                           # log_M_0 = log_M_min (:291)
    c2, _, _ = design.synthetic_batch(128)
    assert np.array_equal(cosmo, c2)                                      # seeded
    d = design.as_dicts(cosmo, halo, hod)[3]
    assert set(d[0]) == set(_lib.COSMO_KEYS) and set(d[2]) == set(_lib.HOD_ZHENG_KEYS)


def test_shards_partition_the_batch():
    for n, w in ((4096, 8), (10, 3), (7, 8)):
        seen = []
        for r in range(w):
            s = design.shard(n, r, w)
            seen += list(range(n))[s]
        assert seen == list(range(n))


def test_covariance_setup_detects_log_spaced_bins():
    """The bins Covariance builds (covariance.py:53-74) are log-spaced to rounding: the set-up hands ln(center_0) and
    the spacing to the library (shift-aligned non-Gaussian kernel); irregular bins switch that off."""
    from chomp_b200 import engine
    survey = engine.Survey(engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1))
    setup = engine.CovarianceSetup(survey, (0.001, 1.0), 10.0, 25.0, [1e10, 1e10], [1e10, 1e10], 1.0, True, "power_gg")
    assert setup.bins.shape == (30, 4)
    assert setup.params.bin_dlog == pytest.approx(np.log(10.0)/10.0, rel=1e-12)
    assert setup.params.bin_log0 == pytest.approx(np.log(setup.bins[0, 2]), rel=1e-15)
    one = engine.CovarianceSetup(survey, (0.5, 0.6), 5.0, 25.0, 1e4, 1e4)
    assert one.bins.shape[0] <= 1 or one.params.bin_dlog > 0


def test_python2_comparison_rules_restated_for_the_generated_reference():
    """oracle/_py2compat.py: None orders below every number (correlation.py:106) and two different Correlation objects
    compare unequal before any array is reached (correlation.py:125-133 under CPython 2.7's dictionary order)."""
    from oracle import _py2compat as P

    class Corr(object):
        pass
    assert P.py2_lt(None, 1e-3) and not P.py2_gt(None, 1e2) and not P.py2_lt(1.0, None) and P.py2_gt(1.0, None)
    assert P.py2_lt(1.0, 2.0) and not P.py2_lt(2.0, 1.0)
    a, b = Corr(), Corr()
    for obj, dz in ((a, 0.77), (b, 0.81)):
        obj.__dict__.update(log_theta_min=-3.0, log_theta_max=-1.0, theta_array=np.arange(5.0), wtheta_array=np.zeros(5),
                            kernel=object(), D_z=dz, halo=object(), _ln_k_min=-6.9, _ln_k_max=4.6, power_spec=None)
    assert P.py2_corr_eq(a, a) and not P.py2_corr_eq(a, b)          # unequal at D_z: no ValueError from the arrays
    assert P._PY2_CORRELATION_KEY_ORDER[0] == "D_z" and P._PY2_CORRELATION_KEY_ORDER[1] == "kernel"
