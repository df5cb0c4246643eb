"""The reference's batch caller (simulation_design.py:36-293) on top of the batched GPU path:
Latin-hypercube design over cosmology / halo / HOD parameters -> model prediction per design
point -> pandas frames.

Same classes, constructor arguments and attributes as the reference (`params` is a dictionary
name -> [center, min, max]; `points`, `lhs`, `design_values`; `run_design()`, `write()`).  When
the wrapped object is a `correlation.Correlation` and the method is `compute_correlation` /
`correlation`, the whole design is evaluated in ONE pass of the four GPU stages
(`Correlation.correlation_batch`); any other object / method is driven point by point through
its setters, as the reference does.  (The batched pass evaluates every design point as a freshly
constructed model; point by point, Halo.set_halo keeps the reference's behaviour of not
rebuilding the profile splines, so c0 / beta would act one design point late there.)  The
reference's module no longer runs under current pandas
(`DataFrame.append`) and its subclasses reference undefined names
(simulation_design.py:274-293); the behaviour they describe is what is implemented here.
"""
import copy

import numpy
import pandas

from . import defaults

default_parameter_dict = {"cosmo_dict": defaults.default_cosmo_dict,
                          "halo_dict": defaults.default_halo_dict,
                          "hod_dict": defaults.default_hod_dict}


def random_lhs(n, k):
    """Random Latin hypercube sample, simulation_design.py:17-33 (numpy's global generator)."""
    P = numpy.zeros((n, k), dtype="float64")
    for i in range(k):
        P[:, i] = numpy.random.permutation(range(n))
    P = P + numpy.random.uniform(size=(n, k))
    return P/n


class SimulationDesign(object):
    """simulation_design.py:36-225."""

    def __init__(self, input_chomp_object, method_name, params, n_design=100,
                 independent_var=None, default_param_dict=None):
        self._input_object = input_chomp_object
        self._method = method_name
        self.params = pandas.DataFrame(params, index=["center", "min", "max"])
        self.n_design = n_design
        self._ind_var = independent_var
        self._initialized_design = False
        if default_param_dict is None:
            default_param_dict = default_parameter_dict
        self._default_param_dict = copy.deepcopy(default_param_dict)
        self._vary_cosmology = self._vary_halo = self._vary_hod = False
        self._param_types = []
        for key in self.params.keys():                       # simulation_design.py:80-101
            for group, flag in (("cosmo_dict", "_vary_cosmology"), ("halo_dict", "_vary_halo"),
                                ("hod_dict", "_vary_hod")):
                if key in default_param_dict[group]:
                    self._param_types.append(group)
                    setattr(self, flag, True)
                    break

    def _init_design_points(self):
        """simulation_design.py:103-114."""
        diff = (self.params.loc["max"] - self.params.loc["min"]).rename("diff")
        self.params = pandas.concat([self.params, diff.to_frame().transpose()])
        points = pandas.DataFrame(random_lhs(self.n_design, self.params.shape[1]), columns=self.params.columns)
        self.lhs = points
        self.points = points*self.params.loc["diff"] + self.params.loc["min"]
        self._initialized_design = True

    # ---- dictionaries of one design point (the subclasses override these) ---------------------
    def cosmo_dict_for(self, point):
        d = dict(self._default_param_dict["cosmo_dict"])
        for key in self.params.keys():
            if key in d:
                d[key] = float(point[key])
        return d

    def halo_dict_for(self, point):
        d = dict(self._default_param_dict["halo_dict"])
        for key in self.params.keys():
            if key in d:
                d[key] = float(point[key])
        return d

    def hod_dict_for(self, point):
        d = dict(self._default_param_dict["hod_dict"])
        for key in self.params.keys():
            if key in d:
                d[key] = float(point[key])
        return d

    # ---- the reference's setters (simulation_design.py:158-214) -------------------------------
    def set_cosmology(self, cosmo_dict=None, values=None):
        self._input_object.set_cosmology(self.cosmo_dict_for(values))

    def set_halo(self, halo_dict=None, values=None):
        self._input_object.set_halo(self.halo_dict_for(values))

    def set_hod(self, hod_dict=None, values=None):
        self._input_object.set_hod(self.hod_dict_for(values))

    def _run_des_point(self, point):
        """simulation_design.py:116-138."""
        if self._vary_cosmology:
            self.set_cosmology(values=point)
        if self._vary_halo:
            self.set_halo(values=point)
        if self._vary_hod:
            self.set_hod(values=point)
        fn = getattr(self._input_object, self._method)
        values = fn() if self._ind_var is None else fn(self._ind_var)
        if values is None and hasattr(self._input_object, "wtheta_array"):
            values = self._input_object.wtheta_array           # compute_correlation() returns nothing
        return pandas.Series(numpy.asarray(values, dtype=float).flatten())

    def _batched(self):
        from . import correlation
        return (type(self._input_object) is correlation.Correlation and
                self._method in ("compute_correlation", "correlation") and
                not getattr(self._input_object.halo, "_use_halofit", False))

    def run_design(self):
        """simulation_design.py:140-156: a frame with one column per design point."""
        if not self._initialized_design:
            self._init_design_points()
        rows = [self.points.iloc[i] for i in range(self.n_design)]
        if self._batched():
            obj = self._input_object
            theta = obj.theta_array if (self._method == "compute_correlation" or self._ind_var is None) \
                else numpy.asarray(self._ind_var, dtype=float)
            cosmo = [self.cosmo_dict_for(p) if self._vary_cosmology else obj.get_cosmology() for p in rows]
            halo = [self.halo_dict_for(p) if self._vary_halo else None for p in rows]
            hod = [self.hod_dict_for(p) if self._vary_hod else None for p in rows]
            w, status = obj.correlation_batch(cosmo, halo, hod, theta)
            self.status = status
            self.design_values = pandas.DataFrame(w.T, columns=self.points.index)
        else:
            self.design_values = pandas.DataFrame({i: self._run_des_point(p) for i, p in enumerate(rows)})
        out = self.design_values.transpose()
        out.columns = ["value_%d" % j for j in range(out.shape[1])]
        self.values_frame = pandas.concat([self.points, out], axis=1)
        return self.design_values

    def write(self, output_name):
        """simulation_design.py:216-224."""
        self.values_frame.to_csv(output_name, index=False, sep=",")


class SimulationDesignFlatUniverse(SimulationDesign):
    """omega_l0 = 1 - omega_m0 - omega_r0 (simulation_design.py:227-241)."""

    def cosmo_dict_for(self, point):
        d = SimulationDesign.cosmo_dict_for(self, point)
        d["omega_l0"] = 1.0 - d["omega_m0"] - d["omega_r0"]
        return d


class SimulationDesignHubbleNormalizedDensities(SimulationDesign):
    """Design in omega_mh2 = Omega_m h^2, omega_bh2 = Omega_b h^2, flat (simulation_design.py:244-267)."""

    def cosmo_dict_for(self, point):
        d = SimulationDesign.cosmo_dict_for(self, point)
        d["omega_m0"] = float(point["omega_mh2"])/d["h"]**2
        d["omega_b0"] = float(point["omega_bh2"])/d["h"]**2
        d["omega_l0"] = 1.0 - d["omega_m0"] - d["omega_r0"]
        return d


class SimulationDesignHODWakeAssumptions(SimulationDesignFlatUniverse):
    """Flat universe and log_M_0 = log_M_min (Wake et al.; simulation_design.py:270-293)."""

    def hod_dict_for(self, point):
        d = SimulationDesign.hod_dict_for(self, point)
        d["log_M_0"] = d["log_M_min"]
        return d
