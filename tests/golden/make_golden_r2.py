#!/usr/bin/env python
"""Generate tests/golden/reference_r2.json by running the REFERENCE ITSELF (oracle/_ref) on the round-2
boundary cases: Correlation(k_min=, k_max=) (correlation.py:104-112, 242-275) with limits wider and narrower
than the halo's k range, inputs = the dictionaries of unit_test.py:59-119.

    python oracle/make_ref.py && python tests/golden/make_golden_r2.py [section ...]

Sections already present in the JSON file are kept unless named on the command line."""
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

PATH = os.path.join(HERE, "reference_r2.json")


def arr(x):
    return [float(v) for v in np.asarray(x, dtype=float).ravel()]


def make_kernel(R, z0=0.5, sigma=0.1):
    cm = R["cosmology"].MultiEpoch(0.0, 5.0, cosmo_dict=C_DICT)
    dist = R["kernel"].dNdzGaussian(0.0, 2.0, z0, sigma)
    wa, wb = R["kernel"].WindowFunctionGalaxy(dist, cm), R["kernel"].WindowFunctionGalaxy(dist, cm)
    return R["kernel"].Kernel(1e-6*D2R, 100.0*D2R, wa, wb, cm)


def k_limits(R):
    """Correlation(k_min=, k_max=): (a) wider than the halo range on both sides -> extrapolation switched on;
    (b) narrower on both sides; (c) only k_max given (Python 2: None < x, extrapolation on)."""
    out = {}
    kern = make_kernel(R)
    for name, kw in (("wide", dict(k_min=1e-4, k_max=1e3)), ("narrow", dict(k_min=5e-3, k_max=30.0)),
                     ("kmax_only", dict(k_max=300.0))):
        for spec in ("power_gg", "power_mm"):
            h = R["halo"].Halo(input_hod=R["hod"].HODZheng(HOD_DICT),
                               cosmo_single_epoch=R["cosmology"].SingleEpoch(0.0, cosmo_dict=C_DICT), halo_dict=H_DICT)
            corr = R["correlation"].Correlation(0.01, 1.0, kern, bins_per_decade=3.0, input_halo=h, power_spec=spec, **kw)
            corr.compute_correlation()
            out[name + "_" + spec] = {"theta": arr(corr.theta_array), "w": arr(corr.wtheta_array),
                                      "extrapolate": bool(h.get_extrapolation()), "args": {k: float(v) for k, v in kw.items()}}
    return out


SECTIONS = {"k_limits": k_limits}


def main():
    os.chdir(tempfile.mkdtemp(prefix="chomp_golden_"))
    R = oracle.import_ref()
    out = json.load(open(PATH)) if os.path.exists(PATH) else {}
    want = sys.argv[1:] or [s for s in SECTIONS if s not in out]
    for s in want:
        print("running", s)
        out[s] = SECTIONS[s](R)
        with open(PATH, "w") as f:
            json.dump(out, f)
    print("wrote", PATH, os.path.getsize(PATH), "bytes")


if __name__ == "__main__":
    main()
