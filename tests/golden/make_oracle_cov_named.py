#!/usr/bin/env python
"""Generate tests/golden/oracle_covariance_named.json: the ORACLE's converged covariance (oracle.covariance_oracle
with the Tight(16) strategy) at BASELINE config 5's named shape -- 30 x 30 annular bins (0.001 .. 1 deg, 10 per
decade), halo_npoints = 200, 1-halo trispectrum gggg at z_bar_NG -- for the first N points of an 8-point Latin
hypercube of the bench workload (chomp_b200.design.synthetic_batch).  The oracle needs ~3 minutes per point at
this shape, too long for the GPU test run, so its outputs are committed as a fixture; the GPU test
(tests/test_gpu_covariance_named.py) compares the CUDA path with them.

    python tests/golden/make_oracle_cov_named.py [n_points]
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

THETA_DEG = (0.001, 1.0)
BPD = 10.0
N_HALO = 200
AREA, N_A, N_B, VARIANCE = 25.0, [1e10, 1e10], [1e10, 1e10], 1.0
N_LHS = 8


def one(args):
    from common import oracle_covariance
    from oracle import chomp_oracle as O
    from oracle.quadrature import Tight
    cd, hd, gd = args
    t0 = time.time()
    cov = oracle_covariance(cd, hd, gd, theta_deg=THETA_DEG, bins_per_decade=BPD, tri_spec="power_gggg",
                            cov_spec="power_gg", area_deg2=AREA, n_a=N_A, n_b=N_B, variance=VARIANCE,
                            prec=O.precision(halo_npoints=N_HALO), integ=Tight(16), corr_bins_per_decade=BPD)
    total, P, G, NG = cov.get_covariance(parts=True)
    return {"cov": np.asarray(total, dtype=float).ravel().tolist(), "P_diag": np.diag(np.asarray(P, dtype=float)).tolist(),
            "G": np.asarray(G, dtype=float).ravel().tolist(), "NG": np.asarray(NG, dtype=float).ravel().tolist(),
            "z_bar_NG": float(cov.kernel.z_bar_NG), "seconds": time.time() - t0}


def main():
    from chomp_b200 import design
    n = int(sys.argv[1]) if len(sys.argv) > 1 else N_LHS
    cosmo, halo, hod = design.synthetic_batch(N_LHS)
    dicts = design.as_dicts(cosmo, halo, hod)[:n]
    with mp.get_context("spawn").Pool(min(n, os.cpu_count() or 1)) as pool:
        rows = pool.map(one, dicts, chunksize=1)
    out = {"theta_deg": list(THETA_DEG), "bins_per_decade": BPD, "halo_npoints": N_HALO, "area_deg2": AREA, "n_a": N_A,
           "n_b": N_B, "variance": VARIANCE, "n_lhs": N_LHS, "strategy": "Tight(16)", "points": rows}
    path = os.path.join(HERE, "oracle_covariance_named.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes;", ["%.0f s" % r["seconds"] for r in rows])


if __name__ == "__main__":
    main()
