"""The batch caller (reference simulation_design.py:36-293): design construction on the host,
and -- on the GPU -- the one-pass batched evaluation against the reference's point-by-point
recipe driven through the same object."""
import numpy as np
import pytest

from chomp_b200 import simulation_design as SD

PARAMS = {"omega_m0": [0.3, 0.25, 0.35], "sigma_8": [0.8, 0.7, 0.9], "log_M_min": [12.2, 11.8, 12.6],
          "c0": [9.0, 7.0, 11.0]}


class _Dummy(object):
    def __init__(self):
        self.calls = []

    def set_cosmology(self, d):
        self.calls.append(("cosmo", dict(d)))

    def set_halo(self, d):
        self.calls.append(("halo", dict(d)))

    def set_hod(self, d):
        self.calls.append(("hod", dict(d)))

    def model(self, x):
        c = [d for k, d in self.calls if k == "cosmo"][-1]
        return np.asarray(x)*c["omega_m0"]


def test_latin_hypercube_has_one_point_per_stratum():
    np.random.seed(3)
    P = SD.random_lhs(64, 5)
    assert P.shape == (64, 5) and P.min() >= 0.0 and P.max() < 1.0
    for j in range(5):
        assert sorted(np.floor(P[:, j]*64).astype(int)) == list(range(64))


def test_design_points_and_parameter_groups():
    np.random.seed(4)
    des = SD.SimulationDesignHODWakeAssumptions(_Dummy(), "model", PARAMS, n_design=16, independent_var=[1.0, 2.0])
    assert des._param_types == ["cosmo_dict", "cosmo_dict", "hod_dict", "halo_dict"]
    assert des._vary_cosmology and des._vary_halo and des._vary_hod
    des._init_design_points()
    assert list(des.points.columns) == list(PARAMS)
    for name, (_, lo, hi) in PARAMS.items():
        assert des.points[name].min() >= lo and des.points[name].max() <= hi
    p = des.points.iloc[0]
    c, g = des.cosmo_dict_for(p), des.hod_dict_for(p)
    assert c["omega_l0"] == pytest.approx(1.0 - c["omega_m0"] - c["omega_r0"])      # simulation_design.py:238-239
    assert g["log_M_0"] == g["log_M_min"] == pytest.approx(p["log_M_min"])            # simulation_design.py:291
    assert des.halo_dict_for(p)["c0"] == pytest.approx(p["c0"])


def test_generic_object_is_driven_point_by_point(tmp_path):
    np.random.seed(5)
    obj = _Dummy()
    des = SD.SimulationDesignFlatUniverse(obj, "model", {"omega_m0": [0.3, 0.2, 0.4]}, n_design=5,
                                          independent_var=[1.0, 2.0, 4.0])
    vals = des.run_design()
    assert vals.shape == (3, 5)                       # one column per design point (simulation_design.py:153)
    assert np.allclose(vals.values, np.outer([1.0, 2.0, 4.0], des.points["omega_m0"].values))
    assert [k for k, _ in obj.calls] == ["cosmo"]*5
    out = tmp_path/"design.csv"
    des.write(str(out))
    assert out.read_text().splitlines()[0] == "omega_m0,value_0,value_1,value_2"
    assert len(out.read_text().splitlines()) == 6


def test_hubble_normalised_densities():
    des = SD.SimulationDesignHubbleNormalizedDensities(_Dummy(), "model", {"omega_mh2": [0.14, 0.12, 0.16],
                                                                           "omega_bh2": [0.022, 0.02, 0.024]}, n_design=4)
    des._init_design_points()
    c = des.cosmo_dict_for(des.points.iloc[1])
    assert c["omega_m0"] == pytest.approx(des.points.iloc[1]["omega_mh2"]/c["h"]**2)
    assert c["omega_b0"] == pytest.approx(des.points.iloc[1]["omega_bh2"]/c["h"]**2)
    assert c["omega_m0"] + c["omega_l0"] + c["omega_r0"] == pytest.approx(1.0)


@pytest.mark.gpu
def test_batched_design_equals_the_point_by_point_recipe():
    from chomp_b200 import correlation as C, cosmology, halo as H, hod, kernel as K
    d2r = np.pi/180.0
    cm = cosmology.MultiEpoch(0.0, 5.0)
    dist = K.dNdzGaussian(0.0, 2.0, 0.5, 0.1)
    kern = K.Kernel(1e-6*d2r, 100.0*d2r, K.WindowFunctionGalaxy(dist, cm), K.WindowFunctionGalaxy(dist, cm), cm)
    corr = C.Correlation(0.001, 1.0, kern, bins_per_decade=10, input_halo=H.Halo(input_hod=hod.HODZheng()),
                         power_spec="power_gg")
    np.random.seed(6)
    # c0 / beta are left out of this comparison: Halo.set_halo does not rebuild the profile splines
    # (the reference's stale-spline behaviour, unit_test.py:388-398, kept by the drop-in class), so
    # point by point they act one design point late; the batched pass evaluates every point as a
    # freshly constructed model
    params = {k: v for k, v in PARAMS.items() if k != "c0"}
    params["stq"] = [0.3, 0.25, 0.35]
    des = SD.SimulationDesignHODWakeAssumptions(corr, "compute_correlation", params, n_design=5)
    batched = des.run_design()
    assert batched.shape == (corr.theta_array.size, 5) and int(np.sum(des.status)) == 0
    assert list(des.values_frame.columns[:4]) == list(params)

    class PointByPoint(SD.SimulationDesignHODWakeAssumptions):
        def _batched(self):
            return False
    ref = PointByPoint(corr, "compute_correlation", params, n_design=5)
    ref.points, ref.lhs, ref.params, ref._initialized_design = des.points, des.lhs, des.params, True
    serial = ref.run_design()
    scale = np.max(np.abs(serial.values), axis=0)
    assert np.max(np.abs(batched.values - serial.values)/scale) < 1e-12
