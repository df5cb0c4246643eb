#!/usr/bin/env python
"""Parity report over a batch: the CUDA path against the oracle's
converged evaluation on N Latin-hypercube parameter points of the bench workload (config 2,
halo_npoints = 200): max and 99th-percentile relative error of the five halo tables, of
P_mm / P_gm / P_gg at the 200 ln k nodes and of w(theta) at the 30 bins.  Bar: 1e-5.

    python tests/parity_report.py [N] [out.json] [config 2|3]     (needs a GPU; the oracle runs on all host cores)

Test infrastructure (it imports oracle/); tests/test_gpu_batch_parity.py runs a small N of it.
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DIST = ("gaussian", (0.0, 2.0, 0.5, 0.1))
N_HALO = 200
# config 3 of BASELINE.json: galaxy-galaxy lensing, lens x source, HODMandelbaum power_gm, J2
CONFIG3 = dict(dist_a=("gaussian", (0.0, 2.0, 0.4, 0.1)), dist_b=("gaussian", (0.0, 2.0, 1.0, 0.2)),
               window_a="galaxy", window_b="convergence", power_spec="power_gm", bessel_order=2, hod_kind="mandelbaum")


def _oracle_point(args):
    from oracle import chomp_oracle as O
    from common import oracle_wtheta
    cd, hd, gd, config = args
    if config == 3:
        ref = oracle_wtheta(cd, hd, gd, CONFIG3["dist_a"], CONFIG3["dist_b"], CONFIG3["window_a"], CONFIG3["window_b"],
                            CONFIG3["power_spec"], prec=O.precision(halo_npoints=N_HALO),
                            bessel_order=CONFIG3["bessel_order"], hod_kind=CONFIG3["hod_kind"])
    else:
        ref = oracle_wtheta(cd, hd, gd, DIST, prec=O.precision(halo_npoints=N_HALO))
    k = np.exp(np.linspace(np.log(1e-3), np.log(1e2), N_HALO))
    h = ref["halo"]
    out = {n: np.asarray(ref[n], dtype=float) for n in ("h_m", "pp_mm", "h_g", "pp_gm", "pp_gg", "w", "nu_nodes")}
    out.update(P_mm=h.power_mm(k), P_gm=h.power_gm(k), P_gg=h.power_gg(k), z_bar=ref["z_bar"])
    return out


def run(n_points, processes=None, seed_offset=0, config=2):
    import torch
    from chomp_b200 import _lib, defaults, design, engine
    from common import w_err
    prec = dict(defaults.default_precision, halo_npoints=N_HALO)
    if config == 3:
        G = engine.RedshiftDistribution.gaussian
        survey = engine.Survey(G(*CONFIG3["dist_a"][1]), G(*CONFIG3["dist_b"][1]), window_a="galaxy", window_b="convergence",
                               bins_per_decade=10.0, bessel_order=2, power_spec="power_gm", hod="mandelbaum", precision=prec)
        hod_name, which = "mandelbaum", _lib.P_GM
    else:
        survey = engine.Survey(engine.RedshiftDistribution.gaussian(*DIST[1]), bins_per_decade=10.0,
                               power_spec="power_gg", precision=prec)
        hod_name, which = "zheng", _lib.P_GG
    eng = engine.Engine(survey)
    cosmo, halo, hod = design.synthetic_batch(n_points, hod=hod_name, seed=design.SEED + seed_offset)
    status = torch.zeros(n_points, dtype=torch.int32, device="cuda")
    w = eng.wtheta(cosmo, halo, hod, survey.theta, which, status=status).cpu().numpy()
    tabs = eng.table(_lib.T_HALO_NODES, n_points).cpu().numpy().reshape(n_points, 5, N_HALO)
    nu = eng.table(_lib.T_NU_NODES, n_points).cpu().numpy()
    zbar = eng.table(_lib.T_ZBAR, n_points).cpu().numpy()[:, 0]
    k = np.exp(np.linspace(np.log(1e-3), np.log(1e2), N_HALO))
    P = {name: eng.power(n_points, which, k).cpu().numpy()
         for name, which in (("P_mm", _lib.P_MM), ("P_gm", _lib.P_GM), ("P_gg", _lib.P_GG))}
    with mp.get_context("spawn").Pool(processes or os.cpu_count()) as pool:
        refs = pool.map(_oracle_point, [d + (config,) for d in design.as_dicts(cosmo, halo, hod, hod_name)], chunksize=1)
    errs = {}
    for i, ref in enumerate(refs):
        for j, name in enumerate(("h_m", "pp_mm", "h_g", "pp_gm", "pp_gg")):
            errs.setdefault(name, []).append(float(np.max(np.abs(tabs[i, j]/ref[name] - 1.0))))
        for name in ("P_mm", "P_gm", "P_gg"):
            r = np.asarray(ref[name], dtype=float)
            nz = r != 0.0                                  # P = 0 beyond the table without extrapolation
            assert np.all(P[name][i][~nz] == 0.0)
            errs.setdefault(name, []).append(float(np.max(np.abs(P[name][i][nz]/r[nz] - 1.0))))
        errs.setdefault("nu_nodes", []).append(float(np.max(np.abs(nu[i]/ref["nu_nodes"] - 1.0))))
        errs.setdefault("w_theta", []).append(w_err(w[i], ref["w"]))
        errs.setdefault("z_bar_abs", []).append(abs(float(zbar[i] - ref["z_bar"])))
    report = {"config": config, "n_points": n_points, "halo_npoints": N_HALO, "nonzero_status_points": int((status != 0).sum().cpu()),
              "errors": {k2: {"max": float(np.max(v)), "p99": float(np.percentile(v, 99)), "median": float(np.median(v))}
                         for k2, v in errs.items()}}
    return report


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rep = run(n, config=int(sys.argv[3]) if len(sys.argv) > 3 else 2)
    text = json.dumps(rep, indent=1)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
