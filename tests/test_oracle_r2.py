"""Pins the oracle on the round-2 boundary cases against committed runs of the reference itself
(tests/golden/reference_r2.json, made by tests/golden/make_golden_r2.py): Correlation(k_min=, k_max=)
including the Python-2 ``None < x`` outcome of correlation.py:104-107."""
import json
import os

import pytest

from oracle.quadrature import Romberg

from common import C_DICT, H_DICT, HOD_DICT, oracle_wtheta, w_err

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_r2.json")))


@pytest.mark.parametrize("key", ["wide_power_gg", "kmax_only_power_mm"])
def test_oracle_k_limits_match_reference_run(key):
    g = GOLD["k_limits"][key]
    spec = "power_gg" if key.endswith("gg") else "power_mm"
    r = oracle_wtheta(C_DICT, H_DICT, HOD_DICT, ("gaussian", (0.0, 2.0, 0.5, 0.1)), power_spec=spec, bins_per_decade=3.0,
                      theta_deg=(0.01, 1.0), integ=Romberg(), **g["args"])
    assert bool(r["halo"].extrapolate) == g["extrapolate"]
    assert w_err(r["w"], g["w"]) < 1e-11
