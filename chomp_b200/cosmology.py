"""Drop-in for the reference's cosmology.py: SingleEpoch (cosmology.py:25-729)
and MultiEpoch (cosmology.py:731-1164), numerics on the GPU.

Supported: flat / open / closed LCDM with w0 = -1, wa = 0 and the no-wiggle
Eisenstein-Hu transfer function, without or with baryon wiggles (with_bao, cosmology.py:474-538; as in the
reference set_cosmology resets it, Q16).  Dynamical dark energy raises NotImplementedError.
Reference quirks kept: growth is the Carroll et al. closed form
(cosmology.py:215-231, 326), E0 has no curvature term (:175-178), the
Python-2 integer exponent in the transfer function (:464).
"""
import numpy as np

from . import _facade, _lib, defaults


def _check(cosmo_dict, with_bao):
    if cosmo_dict["w0"] != -1.0 or cosmo_dict["wa"] != 0.0:
        raise NotImplementedError("w0 != -1 / wa != 0 is not on the GPU path")


class SingleEpoch(object):
    """Cosmological quantities at one redshift (reference cosmology.py:25)."""

    def __init__(self, redshift, cosmo_dict=None, with_bao=False, **kws):
        if redshift < 0.0:
            redshift = 0.0
        if cosmo_dict is None:
            cosmo_dict = defaults.default_cosmo_dict
        _check(cosmo_dict, with_bao)
        self._redshift = redshift
        self.cosmo_dict = cosmo_dict
        self._omega_m0 = cosmo_dict["omega_m0"]
        self._omega_b0 = cosmo_dict["omega_b0"]
        self._omega_l0 = cosmo_dict["omega_l0"]
        self._omega_r0 = cosmo_dict["omega_r0"]
        self._cmb_temp = cosmo_dict["cmb_temp"]
        self._h = cosmo_dict["h"]
        self._sigma_8 = cosmo_dict["sigma_8"]
        self._n = cosmo_dict["n_scalar"]
        self._w0 = cosmo_dict["w0"]
        self._wa = cosmo_dict["wa"]
        self.H0 = 100.0/(2.998*10**5)
        self._with_bao = with_bao
        self._k_min = defaults.default_limits["k_min"]
        self._k_max = defaults.default_limits["k_max"]
        if not hasattr(self, "_gpu"):
            self._gpu = _facade.OnePoint()
        self._initialize_defaults()

    def _initialize_defaults(self):
        """cosmology.py:93-119: chi(z), growth, sigma_norm -- one launch of the
        mass-tables stage for this point."""
        self._gpu.configure(_facade.base_config(with_bao=int(bool(self._with_bao))))     # cosmology.py:556-571
        self._gpu.eng.mass_tables(_facade.cosmo_row(self.cosmo_dict),
                                  _facade.halo_row(defaults.default_halo_dict), [self._redshift])
        e = self._epoch = self._gpu.epoch()
        self._chi, self._growth, self._sigma_norm = e["chi"], e["growth"], e["sigma_norm"]
        self._flat, self._open = bool(e["flat"]), bool(e["open"])
        self._closed = not (self._flat or self._open)

    def set_redshift(self, redshift):
        if redshift != self._redshift:
            self._redshift = redshift
            self._initialize_defaults()

    def get_cosmology(self):
        return self.cosmo_dict

    def set_cosmology(self, cosmo_dict, redshift=None):
        if redshift is None:
            redshift = self._redshift
        self.__init__(redshift, cosmo_dict)          # cosmology.py:151 (with_bao is reset, Q16)

    # -- closed forms evaluated on the device ------------------------------------
    def E(self, redshift):
        return _facade.like_input(redshift, self._gpu.ev(_lib.EVAL_INV_HUBBLE, redshift))

    def E0(self, redshift):
        return _facade.like_input(redshift, self._gpu.ev(_lib.EVAL_E0, redshift))

    def w(self, redshift):
        return self._w0 + self._wa*(1 - 1.0/(1 + np.asarray(redshift, dtype=float)))

    def comoving_distance(self):
        return self._chi

    def luminosity_distance(self):
        return (1.0 + self._redshift)*self._chi

    def angular_diameter_distance(self):
        return self._chi/(1.0 + self._redshift)

    def redshift(self):
        return self._redshift

    def growth_factor(self):
        return self._growth

    def omega_m(self):
        return self._epoch["omega_m"]

    def omega_l(self):
        return self._epoch["omega_l"]

    def delta_c(self):
        return self._epoch["delta_c"]

    def delta_v(self):
        return self._epoch["delta_v_cosmo"]

    def rho_crit(self):
        return self._epoch["rho_crit"]

    def rho_bar(self):
        return self._epoch["rho_bar"]

    def linear_power(self, k):
        return _facade.like_input(k, self._gpu.ev(_lib.EVAL_LINEAR_POWER, k))

    def delta_k(self, k):
        k3 = np.asarray(k, dtype=float)**3
        return k3*self.linear_power(k)/(2.0*np.pi*np.pi)

    def sigma_r(self, scale):
        return _facade.like_input(scale, self._gpu.ev(_lib.EVAL_SIGMA_R, scale))

    def sigma_m(self, mass):
        scale = (3.0*np.asarray(mass, dtype=float)/(4.0*np.pi*self.rho_bar()))**(1.0/3.0)
        return self.sigma_r(scale)

    def nu_r(self, scale):
        s = self.delta_c()/self.sigma_r(scale)
        return s*s

    def nu_m(self, mass):
        s = self.delta_c()/self.sigma_m(mass)
        return s*s

    def write(self, output_power_file_name=None):
        print("z = %1.4f" % self._redshift)
        print("Comoving distance = %1.4f" % self._chi)
        print("Growth factor = %1.4f" % self._growth)
        print("Omega_m(z) = %1.4f" % self.omega_m())
        print("Omega_l(z) = %1.4f" % self.omega_l())
        print("DE w(z)    = %1.4f" % self.w(self._redshift))
        print("Delta_V(z) = %1.4f" % self.delta_v())
        print("delta_c(z) = %1.4f" % self.delta_c())
        print("sigma_8(z) = %1.4f" % self.sigma_r(8.0))
        if output_power_file_name is not None:
            dln_k = (np.log(self._k_max) - np.log(self._k_min))/200
            ln_k = np.arange(np.log(self._k_min) - dln_k, np.log(self._k_max) + 2*dln_k, dln_k)
            k = np.exp(ln_k)
            with open(output_power_file_name, "w") as f:
                f.write("#ttype1 = k [Mpc/h]\n#ttype2 = P(k) [(Mpc/h)^3]\n")
                for a, b in zip(k, self.linear_power(k)):
                    f.write("%1.10f %1.10f\n" % (a, b))


Cosmology = SingleEpoch      # the name BASELINE.json uses


class MultiEpoch(object):
    """Cosmological quantities over a redshift range (reference cosmology.py:731)."""

    def __init__(self, z_min, z_max, cosmo_dict=None, with_bao=False, **kws):
        if cosmo_dict is None:
            cosmo_dict = defaults.default_cosmo_dict
        _check(cosmo_dict, with_bao)
        self.epoch0 = SingleEpoch(0.0, cosmo_dict, with_bao, **kws)
        for name in ("_omega_m0", "_omega_b0", "_omega_l0", "_omega_r0", "_h", "H0", "_sigma_8", "_w0", "_wa",
                     "_flat", "_open", "_closed", "_k_min", "_k_max", "_n"):
            setattr(self, name, getattr(self.epoch0, name))
        self.growth_norm = 1.0
        if not hasattr(self, "_gpu"):
            self._gpu = _facade.OnePoint()
        self.set_redshift(z_min, z_max)

    def __copy__(self):
        # kernel.py:296 shallow-copies the cosmology and re-grids the copy; the copy needs
        # its own device tables
        other = MultiEpoch.__new__(MultiEpoch)
        other.__dict__.update(self.__dict__)
        other._gpu = _facade.OnePoint()
        other._initialize_splines()
        return other

    def set_redshift(self, z_min, z_max):
        self.z_max = z_max
        self.z_min = 0.0 if z_min < 0.0 else z_min
        self._z_array = np.linspace(self.z_min, self.z_max, defaults.default_precision["cosmo_npoints"])
        self._initialize_splines()

    def _initialize_splines(self):
        """cosmology.py:787-817 -- the chi(z), z(chi), D(z) tables come from the
        Limber stage's first tabulation."""
        cfg = _facade.base_config(zk_min=float(self.z_min), zk_max=float(self.z_max))
        hi = max(float(self.z_max), 1e-3)
        for i in range(2):
            cfg.dndz_zmin[i], cfg.dndz_zmax[i] = 0.0, hi
            cfg.dndz_p[i][0], cfg.dndz_p[i][1] = 0.5*hi, 0.25*hi
        self._gpu.configure(cfg)
        self._gpu.eng.limber_tables(_facade.cosmo_row(self.get_cosmology()))
        n = cfg.n_cosmo
        self._chi_array = self._gpu.table(_lib.T_CHI_NODES)[:n].copy()
        self._growth_array = self._gpu.ev(_lib.EVAL_GROWTH_APPROX, self._z_array)

    def get_cosmology(self):
        return self.epoch0.get_cosmology()

    def set_cosmology(self, cosmo_dict, z_min=None, z_max=None):
        if z_min is None:
            z_min = self.z_min
        if z_max is None:
            z_max = self.z_max
        self.__init__(z_min, z_max, cosmo_dict)

    def E(self, redshift):
        return self.epoch0.E(redshift)

    def comoving_distance(self, redshift):
        return _facade.like_input(redshift, self._gpu.ev(_lib.EVAL_CHI_OF_Z, redshift))

    def luminosity_distance(self, redshift):
        return (1.0 + np.asarray(redshift, dtype=float))*self.comoving_distance(redshift)

    def angular_diameter_distance(self, redshift):
        return self.comoving_distance(redshift)/(1.0 + np.asarray(redshift, dtype=float))

    def redshift(self, comoving_distance):
        return _facade.like_input(comoving_distance, self._gpu.ev(_lib.EVAL_Z_OF_CHI, comoving_distance))

    def growth_factor(self, redshift):
        return _facade.like_input(redshift, self._gpu.ev(_lib.EVAL_GROWTH_OF_Z, redshift))

    def omega_m(self, redshift=None):
        if redshift is None:
            redshift = 0.0
        return self._omega_m0*(1.0 + np.asarray(redshift, dtype=float))**3/self.epoch0.E0(redshift)

    def omega_l(self, redshift=None):
        if redshift is None:
            redshift = 0.0
        return self._omega_l0/self.epoch0.E0(redshift)

    def delta_c(self, redshift=None):
        d = 0.15*(12.0*np.pi)**(2.0/3.0)
        if self._open:
            d = d*self.omega_m(redshift)**0.0185
        if self._flat and self._omega_m0 < 1.0001:
            d = d*self.omega_m(redshift)**0.0055
        return d if redshift is None else d/self.growth_factor(redshift)

    def delta_v(self, redshift=None):
        d = 178.0
        if self._open:
            d = d/self.omega_m(redshift)**0.7
        if self._flat and self._omega_m0 < 1.0001:
            d = d/self.omega_m(redshift)**0.55
        return d if redshift is None else d/self.growth_factor(redshift)

    def rho_crit(self, redshift=None):
        if redshift is None:
            redshift = 0.0
        return 1.879/1.989*3.086**3*1e10*self.epoch0.E0(redshift)

    def rho_bar(self, redshift=None):
        return self.rho_crit(redshift)*self.omega_m(redshift)

    def delta_k(self, k, redshift=None):
        d = self.epoch0.delta_k(k)
        if redshift is not None:
            d = d*self.growth_factor(redshift)**2
        return d

    def linear_power(self, k, redshift=None):
        k = np.asarray(k, dtype=float)
        return 2.0*np.pi*np.pi*self.delta_k(k, redshift)/(k*k*k)

    def sigma_r(self, scale, redshift=None):
        s = self.epoch0.sigma_r(scale)
        if redshift is not None:
            s = s*self.growth_factor(redshift)
        return s

    def sigma_m(self, mass, redshift=None):
        scale = (3.0*np.asarray(mass, dtype=float)/(4.0*np.pi*self.rho_bar(redshift)))**(1.0/3.0)
        return self.sigma_r(scale, redshift)

    def nu_r(self, scale, redshift=None):
        s = self.delta_c(redshift)/self.sigma_r(scale, redshift)
        return s*s

    def nu_m(self, mass, redshift=None):
        s = self.delta_c(redshift)/self.sigma_m(mass, redshift)
        return s*s

    def write(self, output_file_name, output_power_file_name=None):
        with open(output_file_name, "w") as f:
            f.write("#ttype1 = z\n#ttype2 = chi [Mpc/h]\n#ttype3 = growth\n#ttype4 = omega_m\n"
                    "#ttype5 = omega_l\n#ttype6 = delta_c\n#ttype7 = delta_v\n#ttype8 = sigma_8\n")
            for z, chi, growth in zip(self._z_array, self._chi_array, self._growth_array):
                f.write("%1.10f %1.10f %1.10f %1.10f %1.10f %1.10f %1.10f %1.10f\n" % (
                    z, chi, growth, self.omega_m(z), self.omega_l(z), self.delta_c(z), self.delta_v(z),
                    self.sigma_r(8.0, z)))
        if output_power_file_name is not None:
            k = np.exp(np.linspace(np.log(self._k_min), np.log(self._k_max), 100))
            with open(output_power_file_name, "w") as f:
                for a, b in zip(k, self.linear_power(k)):
                    f.write("%1.10f %1.10f\n" % (a, b))
