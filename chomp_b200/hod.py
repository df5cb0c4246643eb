"""Drop-in for the reference's hod.py: HODZheng (hod.py:141-230) and
HODMandelbaum (hod.py:232-299).  The moments are evaluated on the GPU."""
import numpy as np

from . import _facade, _lib, defaults


class HOD(object):
    """Base class (hod.py:17-129)."""
    _kind = _lib.HOD_ZHENG

    def __init__(self, hod_dict):
        self.hod_dict = hod_dict
        self._hod = {1: self.first_moment, 2: self.second_moment}
        self.first_moment_zero = -1
        self.second_moment_zero = -1
        self._safe_norm = -1

    def _device(self):
        gpu = getattr(self, "_gpu", None)
        if gpu is None:
            gpu = self._gpu = _facade.OnePoint()
        cfg = _facade.base_config(hod_kind=self._kind)
        cfg.halo_precision = self._halo_precision
        gpu.configure(cfg)
        gpu.eng.set_params(hod=_facade.hod_row(self._kind, self._params()))
        return gpu

    def first_moment(self, mass, z=None):
        return _facade.like_input(mass, self._device().ev(_lib.EVAL_FIRST_MOMENT, mass))

    def second_moment(self, mass, z=None):
        return _facade.like_input(mass, self._device().ev(_lib.EVAL_SECOND_MOMENT, mass))

    def nth_moment(self, mass, n=3, z=None):
        return _facade.like_input(mass, self._device().ev(_lib.EVAL_NTH_MOMENT, mass, aux=float(n)))

    def satellite_first_moment(self, mass, z=None):
        # <N(N-1)> = (2 + N_s) N_s  (hod.py:192-194)  =>  N_s = sqrt(1 + <N(N-1)>) - 1
        return np.sqrt(1.0 + self.second_moment(mass)) - 1.0

    def central_first_moment(self, mass, z=None):
        return self.first_moment(mass) - self.satellite_first_moment(mass)

    def get_hod(self):
        return self.hod_dict

    def set_hod(self, hod_dict):
        self.__init__(hod_dict)

    def set_halo(self, halo_dict):
        pass

    def write(self, output_file_name):
        dln = (np.log(1.0e16) - np.log(1.0e9))/200
        mass = np.exp(np.arange(np.log(1.0e9) - dln, np.log(1.0e16) + 2*dln, dln))
        a, b, c = self.first_moment(mass), self.second_moment(mass), self.nth_moment(mass, 3)
        with open(output_file_name, "w") as f:
            for row in zip(mass, a, b, c):
                f.write("%1.10f %1.10f %1.10f %1.10f\n" % row)


class HODZheng(HOD):
    """Zheng et al. 2007 five-parameter HOD (hod.py:141)."""
    _kind = _lib.HOD_ZHENG

    def __init__(self, hod_dict=None):
        src = defaults.default_hod_dict if hod_dict is None else hod_dict
        self.log_M_min = src["log_M_min"]
        self.sigma = src["sigma"]
        self.log_M_0 = src["log_M_0"]
        self.log_M_1p = src["log_M_1p"]
        self.alpha = src["alpha"]
        HOD.__init__(self, hod_dict)
        # first_moment_zero depends on the halo_precision in force now (hod.py:176-179)
        self._halo_precision = defaults.default_precision["halo_precision"]
        zeros = self._device().ev(_lib.EVAL_HOD_ZEROS, [0.0, 1.0, 2.0])
        self.first_moment_zero, self.second_moment_zero, self._safe_norm = (float(v) for v in zeros)

    def _params(self):
        return {"log_M_min": self.log_M_min, "sigma": self.sigma, "log_M_0": self.log_M_0,
                "log_M_1p": self.log_M_1p, "alpha": self.alpha}


class HODMandelbaum(HOD):
    """Mandelbaum et al. 2005 two-parameter HOD (hod.py:232)."""
    _kind = _lib.HOD_MANDELBAUM

    def __init__(self, hod_dict=None):
        self._halo_precision = defaults.default_precision["halo_precision"]
        if hod_dict is None:
            # the reference skips HOD.__init__ here (hod.py:251-254): such an object has no
            # first_moment_zero and cannot be handed to Halo
            self.log_M_0 = 12.14
            self.log_M_min = np.log10(3.0) + 12.14
            self.w = 1.0
        else:
            self.log_M_0 = hod_dict["log_M_0"]
            self.log_M_min = np.log10(3.0) + hod_dict["log_M_0"]
            self.w = hod_dict["w"]
            HOD.__init__(self, hod_dict)

    def _params(self):
        return {"log_M_0": self.log_M_0, "w": self.w}


HODMand = HODMandelbaum     # the name BASELINE.json uses
