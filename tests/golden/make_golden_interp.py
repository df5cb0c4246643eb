#!/usr/bin/env python
"""Generate tests/golden/reference_interp.json by running the REFERENCE ITSELF (oracle/_ref):
dNdzInterpolation (kernel.py:181-208) as the redshift distribution of a galaxy window and of a
lensing (convergence) window, through Kernel and Correlation.

    python oracle/make_ref.py && python tests/golden/make_golden_interp.py
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from common import C_DICT, D2R, H_DICT, HOD_DICT, interp_table  # noqa: E402


def arr(x):
    return [float(v) for v in np.asarray(x, dtype=float).ravel()]


def main():
    os.chdir(tempfile.mkdtemp(prefix="chomp_golden_"))
    R = oracle.import_ref()
    cosmology, hod, halo, kernel, correlation = (R[k] for k in ("cosmology", "hod", "halo", "kernel", "correlation"))
    z_arr, p_arr = interp_table()
    out = {"z_array": arr(z_arr), "p_array": arr(p_arr)}
    z_eval = np.linspace(0.0, 1.8, 37)
    for order in (1, 2, 3):
        d = kernel.dNdzInterpolation(z_arr, p_arr, interpolation_order=order)
        out["order%d" % order] = {"norm": float(d.norm), "z": arr(z_eval), "dndz": arr(d.dndz(z_eval)),
                                  "raw": arr(d.raw_dndz(z_eval))}
    cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    for name, wcls in (("galaxy", kernel.WindowFunctionGalaxy), ("convergence", kernel.WindowFunctionConvergence)):
        d = kernel.dNdzInterpolation(z_arr, p_arr)
        wa, wb = wcls(d, cm), wcls(d, cm)
        kern = kernel.Kernel(1e-6*D2R, 100.0*D2R, wa, wb, cm)
        h = halo.Halo(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT),
                      halo_dict=H_DICT)
        corr = correlation.Correlation(0.001, 1.0, kern, bins_per_decade=10, input_halo=h, power_spec="power_mm")
        corr.compute_correlation()
        out[name] = {"z_bar": float(kern.z_bar), "kernel_nodes": arr(kern._kernel_array),
                     "window_nodes": arr(kern.window_function_a._wf_array), "theta": arr(corr.theta_array),
                     "w": arr(corr.wtheta_array)}
    path = os.path.join(HERE, "reference_interp.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
