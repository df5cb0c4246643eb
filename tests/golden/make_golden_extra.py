#!/usr/bin/env python
"""Generate tests/golden/reference_extra.json by running the REFERENCE ITSELF (oracle/_ref):
MassFunctionSecondOrder (mass_function.py:365-433) and CorrelationFourier with the halo-model
(table-based) spectra (correlation.py:297-405), inputs = the dictionaries of unit_test.py:59-119.

    python oracle/make_ref.py && python tests/golden/make_golden_extra.py
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
from common import C_DICT, D2R, H_DICT, HOD_DICT  # noqa: E402


def arr(x):
    return [float(v) for v in np.asarray(x, dtype=float).ravel()]


def main():
    os.chdir(tempfile.mkdtemp(prefix="chomp_golden_"))
    R = oracle.import_ref()
    cosmology, mass_function, hod, halo, kernel, correlation = (
        R[k] for k in ("cosmology", "mass_function", "hod", "halo", "kernel", "correlation"))
    out = {}
    M = np.logspace(9, 16, 8)
    out["second_order"] = {}
    for z in (0.0, 0.5):
        cs = cosmology.SingleEpoch(z, cosmo_dict=C_DICT)
        mf = mass_function.MassFunctionSecondOrder(z, cs, H_DICT)
        nu = mf.nu(M)
        out["second_order"]["z%.1f" % z] = {
            "masses": arr(M), "nu": arr(nu), "bias_2_norm": float(mf.bias_2_norm), "bias_2_nu": arr(mf.bias_2_nu(nu)),
            "sigma_nodes": arr(mf._sigma_array), "bias_norm": float(mf.bias_norm), "f_norm": float(mf.f_norm)}
    cm = cosmology.MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    dist = kernel.dNdzGaussian(0.0, 2.0, 0.5, 0.1)
    wa, wb = kernel.WindowFunctionGalaxy(dist, cm), kernel.WindowFunctionGalaxy(dist, cm)
    kern = kernel.Kernel(1e-6*D2R, 100.0*D2R, wa, wb, cm)
    ell = np.logspace(0.5, 5.5, 16)
    out["cl_tables"] = {"ell": arr(ell), "z_bar": float(kern.z_bar)}
    for extrapolate in (False, True):
        for spec in ("power_mm", "power_gm", "power_gg"):
            h = halo.Halo(input_hod=hod.HODZheng(HOD_DICT), cosmo_single_epoch=cosmology.SingleEpoch(0.0, cosmo_dict=C_DICT),
                          halo_dict=H_DICT, extrapolate=extrapolate)
            cf = correlation.CorrelationFourier(10, 1e5, kern, input_halo=h, powSpec=spec)
            out["cl_tables"][spec + ("_extrapolated" if extrapolate else "")] = [float(cf.correlation(l)) for l in ell]
    path = os.path.join(HERE, "reference_extra.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
