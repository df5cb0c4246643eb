"""Batch parity (parity report): eight Latin-hypercube points of the bench
workload (halo_npoints = 200) against the oracle's converged evaluation; the full report
(tests/parity_report.py, 64+ points) is committed under profiles/."""
import pytest

pytestmark = pytest.mark.gpu


def test_eight_lhs_points_within_the_bar():
    import parity_report
    rep = parity_report.run(8, processes=8, seed_offset=7)
    assert rep["nonzero_status_points"] == 0
    e = rep["errors"]
    for name in ("h_m", "pp_mm", "h_g", "pp_gm", "pp_gg"):
        assert e[name]["max"] < 2e-6, (name, e[name])
    for name in ("P_mm", "P_gm", "P_gg", "w_theta"):
        assert e[name]["max"] < 1e-5, (name, e[name])
    assert e["z_bar_abs"]["max"] < 1e-12
