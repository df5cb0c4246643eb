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


def test_grouped_batch_equals_expanded_batch():
    """Fast / slow split (chomp_b200_wtheta_batch_grouped): 6 (cosmology, halo) rows shared by 48 HOD points
    in scrambled order give bit for bit the tables of the expanded batch; a bad index is flagged."""
    import numpy as np
    import torch
    from chomp_b200 import _lib, design, engine
    survey = engine.Survey(engine.RedshiftDistribution.gaussian(0.0, 2.0, 0.5, 0.1), bins_per_decade=10.0,
                           power_spec="power_gg")
    eng = engine.Engine(survey)
    G, B = 6, 48
    cosmo_g, halo_g, _ = design.synthetic_batch(G)
    _, _, hod = design.synthetic_batch(B, seed=5)
    rng = np.random.default_rng(3)
    idx = rng.integers(0, G, size=B).astype(np.int32)
    st_e = torch.zeros(B, dtype=torch.int32, device="cuda")
    expanded = eng.wtheta(cosmo_g[idx], halo_g[idx], hod, survey.theta, _lib.P_GG, status=st_e).cpu().numpy()
    st_g = torch.zeros(B, dtype=torch.int32, device="cuda")
    grouped = eng.wtheta_grouped(cosmo_g, halo_g, idx, hod, survey.theta, _lib.P_GG, status=st_g).cpu().numpy()
    assert np.all(np.isfinite(expanded)) and not st_e.any() and not st_g.any()
    assert np.array_equal(grouped, expanded)
    # the handle still serves ordinary batches afterwards, and group-level status reaches the points
    again = eng.wtheta(cosmo_g[idx], halo_g[idx], hod, survey.theta, _lib.P_GG).cpu().numpy()
    assert np.array_equal(again, expanded)
    bad = idx.copy()
    bad[5] = G + 3
    cosmo_bad = cosmo_g.copy()
    cosmo_bad[2, 8] = -0.9                      # w0 != -1 in group 2
    st = torch.zeros(B, dtype=torch.int32, device="cuda")
    eng.wtheta_grouped(cosmo_bad, halo_g, bad, hod, survey.theta, _lib.P_GG, status=st)
    st = st.cpu().numpy()
    assert st[5] & _lib.ST_DOMAIN
    assert all((st[i] & _lib.ST_DOMAIN) != 0 for i in range(B) if bad[i] == 2)
    assert all(st[i] == 0 for i in range(B) if bad[i] not in (2, G + 3))
