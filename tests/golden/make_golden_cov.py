#!/usr/bin/env python
"""Generate tests/golden/reference_covariance.json by running the REFERENCE's own
covariance.Covariance (generated copy oracle/_ref, see oracle/make_ref.py) in this container:
config 5 of BASELINE.json (config 2's galaxy-clustering set-up + HaloTrispectrumOneHalo +
Covariance) at the reference's default tolerances.  ~5 minutes of CPU (the 1275 Rombergs of
KernelCovariance._initialize_NG_spline, kernel.py:1016-1030, dominate).

    python oracle/make_ref.py && python tests/golden/make_golden_cov.py
"""
import importlib
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from common import C_DICT, D2R, H_DICT, HOD_DICT  # noqa: E402

THETA_DEG = (0.01, 1.0)
TRI_Z = 0.5
AREA, N_A, N_B, VARIANCE = 25.0, [1e10, 1e10], [1e10, 1e10], 1.0


def arr(x):
    return [float(v) for v in np.asarray(x, dtype=float).ravel()]


def main():
    os.chdir(tempfile.mkdtemp(prefix="chomp_golden_"))   # Kernel.__init__ writes debug files into the CWD
    R = oracle.import_ref()
    cosmology, mass_function, hod, halo, kernel, correlation, tri_mod = (
        R[k] for k in ("cosmology", "mass_function", "hod", "halo", "kernel", "correlation", "halo_trispectrum"))
    covariance = importlib.import_module("covariance")
    t0 = time.time()
    cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    dist = kernel.dNdzGaussian(0.0, 2.0, 0.5, 0.1)
    wa = kernel.WindowFunctionGalaxy(dist, cm)
    wb = kernel.WindowFunctionGalaxy(dist, cm)
    kern = kernel.Kernel(1e-6*D2R, 100.0*D2R, wa, wb, cm)
    cs = cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT)
    h = halo.Halo(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cs, halo_dict=H_DICT)
    corr = correlation.Correlation(THETA_DEG[0], THETA_DEG[1], kern, bins_per_decade=5.0, input_halo=h,
                                   power_spec="power_gg")
    out = {"theta_deg": list(THETA_DEG), "tri_z": TRI_Z, "area_deg2": AREA, "n_a": N_A, "n_b": N_B,
           "variance": VARIANCE, "cases": {}}
    for tri_spec, cov_spec in (("power_gggg", "power_gg"), ("power_mmmm", "power_mm")):
        cs_t = cosmology.SingleEpoch(TRI_Z, cosmo_dict=C_DICT)
        mf_t = mass_function.MassFunction(TRI_Z, cs_t, H_DICT)
        tri = tri_mod.HaloTrispectrumOneHalo(TRI_Z, cs_t, mf_t, None, H_DICT, hod.HODZheng(HOD_DICT), tri_spec)
        cov = covariance.Covariance(corr, corr, bins_per_decade=5.0, survey_area_deg2=AREA, n_a=N_A, n_b=N_B,
                                    variance=VARIANCE, nongaussian_cov=True, input_halo_trispectrum=tri,
                                    power_spec=cov_spec)
        if out["cases"]:
            # K_NG depends only on the cosmology and the windows: reuse the first case's table
            prev = out["_kernel_obj"]
            for name in ("_kernel_array", "_kernel_NG_min", "_kernel_NG_spline", "_initialized_NG_spline"):
                setattr(cov.kernel, name, getattr(prev, name))
        bins = cov.annular_bins
        n = len(bins)
        P = np.zeros((n, n))
        G = np.zeros((n, n))
        NG = np.zeros((n, n))
        for i in range(n):
            for j in range(i, n):
                a, b = bins[i], bins[j]
                if i == j:
                    P[i, j] = cov.covariance_P(a.delta, a.center)
                G[i, j] = G[j, i] = cov.covariance_G(a.center, b.center, a.delta, b.delta)
                NG[i, j] = NG[j, i] = cov.covariance_NG(a.center, b.center)
            print(tri_spec, "row", i, "of", n, "%.0f s" % (time.time() - t0), flush=True)
        total = np.asarray(cov.get_covariance(), dtype=float)
        kc = cov.kernel
        entry = {
            "bins_inner": arr([b.inner for b in bins]), "bins_outer": arr([b.outer for b in bins]),
            "bins_center": arr([b.center for b in bins]), "bins_delta": arr([b.delta for b in bins]),
            "equal_windows": [bool(x) for x in cov.equal_windows], "cosmic_shear": [bool(x) for x in cov.cosmic_shear],
            "z_bar_NG": float(kc.z_bar_NG), "D_z_NG": float(cov.D_z_NG),
            "kernel_chi_range": [float(kc.chi_min), float(kc.chi_max)],
            "ln_ktheta_range": [float(kc.ln_ktheta_min), float(kc.ln_ktheta_max)],
            "kernel_NG_table": arr(np.asarray(kc._kernel_array, dtype=float)),
            "kernel_NG_min": float(kc._kernel_NG_min),
            "ln_K_array": arr(cov._ln_K_array), "projected_a": arr(cov._halo_a_spline(cov._ln_K_array)),
            "z_bar_G": float(cov._z_bar_G_a), "D_z_G": float(cov._D_z_a),
            "chi_range_a": [float(cov._chi_min_a), float(cov._chi_max_a)],
            "tri_table": arr(tri._i_0_4_array),
            "cov_P": arr(P), "cov_G": arr(G), "cov_NG": arr(NG), "cov": arr(total)}
        out["cases"][tri_spec] = entry
        out["_kernel_obj"] = kc
    del out["_kernel_obj"]
    path = os.path.join(HERE, "reference_covariance.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes, %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
