"""Drop-in for the reference's halo.Halo and halo.HaloExclusion
(halo.py:23-1086, 1201-1233): Seljak (2000) halo-model power spectra from the
GPU tables.  HaloFit and HaloSuperSampleCovariance are listed as next in
SURVEY.md section 8(f).

The lazy-table and setter semantics of the reference are kept, including the
fact that ``set_halo`` refreshes the mass function but leaves the halo-profile
tables (c(M), r_vir(M)) as they were (halo.py:196-212) and that
``set_cosmology`` rebuilds them from the constructor's halo dictionary and the
current ``beta`` attribute (halo.py:135-173).
"""
import numpy as np

from . import _facade, _lib, cosmology, defaults, hod as hod_module, mass_function


class Halo(object):
    _exclusion = 0
    _halofit = False

    def __init__(self, redshift=0.0, input_hod=None, cosmo_single_epoch=None, mass_func=None,
                 halo_dict=None, extrapolate=False, **kws):
        self._k_min = defaults.default_limits["k_min"]
        self._k_max = defaults.default_limits["k_max"]
        self._ln_k_min, self._ln_k_max = np.log(self._k_min), np.log(self._k_max)
        self._ln_k_array = np.linspace(self._ln_k_min, self._ln_k_max, defaults.default_precision["halo_npoints"])
        self._redshift = redshift
        if cosmo_single_epoch is None:
            cosmo_single_epoch = cosmology.SingleEpoch(redshift)
        self.cosmo = cosmo_single_epoch
        if halo_dict is None:
            halo_dict = defaults.default_halo_dict
        self.halo_dict = halo_dict
        if mass_func is None:
            mass_func = mass_function.MassFunction(self._redshift, self.cosmo, self.halo_dict)
        self.mass = mass_func
        self.c0 = halo_dict["c0"]/(1.0 + self._redshift)
        self.beta = halo_dict["beta"]
        self.alpha = halo_dict["alpha"]
        if self.alpha != -1.0:
            raise NotImplementedError("non-NFW profiles (alpha != -1) are not on the GPU path")
        self.delta_v = self.halo_dict["delta_v"]
        if self.delta_v == -1:
            self.delta_v = self.cosmo.delta_v()
        self.rho_bar = self.cosmo.rho_bar()
        self._h = self.cosmo._h
        if input_hod is None:
            input_hod = hod_module.HODZheng()
        self.local_hod = input_hod
        self._extrapolate = extrapolate
        self._gpu = _facade.OnePoint()
        self._initialize_halo_splines()
        self._dirty = True

    # ---- state that the reference keeps in spline objects ---------------------------------
    def _initialize_halo_splines(self):
        """halo.py:839-855: freeze the profile parameters the tables are built from."""
        self._profile = {"c0": self.c0*(1.0 + self._redshift), "beta": self.beta,
                         "delta_v": self.halo_dict["delta_v"] if self.halo_dict["delta_v"] != -1 else -1.0}

    def _ensure(self):
        if not self._dirty:
            return
        cfg = _facade.base_config(hod_kind=self.local_hod._kind, exclusion=self._exclusion,
                                  extrapolate=int(bool(self._extrapolate)),
                                  tri_moment=int(getattr(self, "_tri_moment", -1)),
                                  use_halofit=int(self._halofit),
                                  with_bao=int(bool(getattr(self.cosmo, "_with_bao", False))),
                                  mass_function_kind=int(getattr(self.mass, "_mf_kind", 0)))
        cfg.halo_precision = getattr(self.local_hod, "_halo_precision", cfg.halo_precision)
        # first_moment_zero was fixed when the HOD object was built (hod.py:176-179)
        self._gpu.configure(cfg)
        eng = self._gpu.eng
        hrow = _facade.halo_row(self.mass.halo_dict, self._profile)
        eng.mass_tables(_facade.cosmo_row(self.cosmo.cosmo_dict), hrow, [self._redshift])
        eng.halo_tables(hrow, _facade.hod_row(self.local_hod._kind, self.local_hod._params()))
        if self._halofit:
            fit = eng.halofit(1, fit_z=self._fit_redshift).cpu().numpy()[0]
            for name, value in zip(_lib.HALOFIT_FIELDS, fit):
                setattr(self, "_" + name, float(value))
        self.n_bar_over_rho_bar = float(self._gpu.table(_lib.T_NBAR)[0])
        self.n_bar = self.n_bar_over_rho_bar*self.rho_bar
        self._dirty = False
        self._extrapolate_built = bool(self._extrapolate)

    def _power(self, which, k):
        if getattr(self, "_extrapolate_built", None) != bool(self._extrapolate):
            self._dirty = True
        self._ensure()
        return _facade.like_input(k, self._gpu.eng.power(1, which, _facade.flat(k)).cpu().numpy()[0])

    # ---- reference API ---------------------------------------------------------------------------
    def get_extrapolation(self):
        return self._extrapolate

    def set_extrapolation(self, boolean):
        self._extrapolate = boolean

    def get_cosmology(self):
        return self.cosmo.get_cosmology()

    def get_cosmology_object(self):
        return self.cosmo

    def set_cosmology(self, cosmo_dict, redshift=None):
        if redshift is None:
            redshift = self._redshift
        self.cosmo_dict = cosmo_dict
        self._redshift = redshift
        self.cosmo = cosmology.SingleEpoch(redshift, cosmo_dict)
        self.delta_v = self.halo_dict["delta_v"]
        if self.delta_v == -1:
            self.delta_v = self.cosmo.delta_v()
        self.rho_bar = self.cosmo.rho_bar()
        self._h = self.cosmo._h
        self.c0 = self.halo_dict["c0"]/(1.0 + redshift)
        self.mass.set_cosmology_object(self.cosmo)
        self._initialize_halo_splines()
        self._dirty = True

    def get_hod(self, return_object=False):
        return self.local_hod.get_hod()

    def get_hod_object(self):
        return self.local_hod

    def set_hod(self, hod_dict):
        self.local_hod.set_hod(hod_dict)
        self._dirty = True

    def set_hod_object(self, input_hod):
        self.local_hod = input_hod
        self._dirty = True

    def get_halo(self):
        return self.halo_dict

    def set_halo(self, halo_dict=None):
        self.c0 = halo_dict["c0"]/(1.0 + self._redshift)
        self.beta = halo_dict["beta"]
        self.alpha = -1.0
        self.mass.set_halo(halo_dict)
        self.set_hod_object(self.local_hod)

    def get_mass(self):
        return self.mass

    def get_redshift(self):
        return self._redshift

    def set_redshift(self, redshift):
        if redshift != self._redshift:
            self.set_cosmology(self.cosmo.cosmo_dict, redshift)

    def linear_power(self, k):
        return self.cosmo.linear_power(k)

    def power_mm(self, k):
        return self._power(_lib.P_MM, k)

    def power_gm(self, k):
        return self._power(_lib.P_GM, k)

    def power_mg(self, k):
        return self.power_gm(k)

    def power_gg(self, k):
        return self._power(_lib.P_GG, k)

    def virial_radius(self, mass):
        self._ensure()
        return _facade.like_input(mass, self._gpu.ev(_lib.EVAL_VIRIAL_RADIUS, mass))

    def concentration(self, mass):
        self._ensure()
        return _facade.like_input(mass, self._gpu.ev(_lib.EVAL_CONCENTRATION, mass))

    def y(self, ln_k, mass):
        self._ensure()
        return _facade.like_input(mass, self._gpu.ev(_lib.EVAL_Y_NFW, mass, aux=float(ln_k)))

    y_nfw = y

    def _table(self, index, k):
        self._ensure()
        k = np.asarray(k, dtype=float)
        nodes = self._gpu.table(_lib.T_HALO_NODES).reshape(5, -1)[index]
        return nodes if k.shape == self._ln_k_array.shape and np.allclose(np.log(k), self._ln_k_array) else None

    def halo_tables(self):
        """Node values of h_m, pp_mm, h_g, pp_gm, pp_gg on `_ln_k_array`
        (the arrays behind halo.py:919-1076's splines)."""
        self._ensure()
        t = self._gpu.table(_lib.T_HALO_NODES).reshape(5, -1)
        return dict(zip(("h_m", "pp_mm", "h_g", "pp_gm", "pp_gg"), t))

    def write(self, output_file_name):
        k = np.exp(self._ln_k_array)
        cols = (k, self.linear_power(k), self.power_mm(k), self.power_gg(k), self.power_gm(k))
        with open(output_file_name, "w") as f:
            f.write("#ttype1 = k [Mpc/h]\n#ttype2 = linear_power [(Mpc/h)^3]\n#ttype3 = power_mm\n"
                    "#ttype4 = power_gg\n#ttype5 = power_gm\n")
            for row in zip(*cols):
                f.write("%1.10f %1.10f %1.10f %1.10f %1.10f\n" % row)

    def write_power_components(self, output_file_name):
        t = self.halo_tables()
        with open(output_file_name, "w") as f:
            f.write("#ttype1 = k [Mpc/h]\n#ttype2 = 2 halo dark matter component\n"
                    "#ttype3 = dark matter poisson component\n#ttype4 = 2 halo galaxy component\n"
                    "#ttype5 = matter-galaxy poisson component\n#ttype6 = galaxy-galaxy poisson component\n")
            for row in zip(np.exp(self._ln_k_array), t["h_m"], t["pp_mm"], t["h_g"], t["pp_gm"], t["pp_gg"]):
                f.write("%1.10f %1.10f %1.10f %1.10f %1.10f %1.10f\n" % row)


class HaloExclusion(Halo):
    """Halo model with the halo-exclusion mass window (halo.py:1201-1233)."""
    _exclusion = 1

    def __init__(self, redshift=0.0, input_hod=None, cosmo_single_epoch=None, mass_func=None,
                 halo_dict=None, **kws):
        Halo.__init__(self, redshift, input_hod, cosmo_single_epoch, mass_func, halo_dict, **kws)


class HaloFit(Halo):
    """HALOFIT matter power spectrum (Smith et al. 2003, Takahashi et al. 2012 coefficients);
    power_gm / power_gg are the halo-model tables on top of it (halo.py:1236-1412).

    As in the reference, ``extrapolate`` is not forwarded (halo.py:1254-1259), power_mm has no
    k-range guards, and Omega_m(z), Omega_L(z) inside f_1..f_3 are those of the *construction*
    redshift (halo.py:1261-1266 runs once; Correlation* moving the halo to z_bar does not redo
    it).  Deviation: after ``set_cosmology`` the reference keeps the construction cosmology in
    f_1..f_3 as well; here they follow the new cosmology at the construction redshift."""
    _halofit = True

    def __init__(self, redshift=0.0, input_hod=None, cosmo_single_epoch=None, mass_func=None,
                 halo_dict=None, **kws):
        self._fit_redshift = float(redshift)
        Halo.__init__(self, redshift, input_hod, cosmo_single_epoch, mass_func, halo_dict)
        self._initialized_sigma_spline = False


class HaloSuperSampleCovariance(Halo):
    """halo.py:1089-1199 (Takada & Hu 2013): the response of the matter power spectrum to a super-survey
    over-density delta_b, through the extra mass integral I^1_2(k) = rho_bar^-1 int dln nu nu f b y^2 M."""

    def __init__(self, redshift=0.0, input_hod=None, cosmo_single_epoch=None, mass_func=None, halo_dict=None,
                 extrapolate=False, delta_b=0.0, **kws):
        Halo.__init__(self, redshift, input_hod, cosmo_single_epoch, mass_func, halo_dict, **kws)   # extrapolate not forwarded (halo.py:1105-1106)
        self._delta_b = delta_b

    @staticmethod
    def init_from_halo(input_halo, delta_b=0.0):
        """halo.py:1110-1136 (the tables of the donor are rebuilt here rather than copied)."""
        return HaloSuperSampleCovariance(input_halo.get_redshift(), input_halo.get_hod_object(), input_halo.get_cosmology_object(),
                                         input_halo.get_mass(), input_halo.get_halo(), input_halo.get_extrapolation(), delta_b)

    def _ssc(self, what, k):
        if getattr(self, "_extrapolate_built", None) != bool(self._extrapolate):
            self._dirty = True
        self._ensure()
        return _facade.like_input(k, self._gpu.eng.halo_ssc(1, _facade.flat(k), what).cpu().numpy()[0])

    def _i_1_2(self, k):
        return self._ssc(0, k)

    def dln_power_ddelta_b(self, k):
        return self._ssc(1, k)

    def power_mm_ssc(self, k):
        return self.power_mm(k)*(1.0 + self.dln_power_ddelta_b(k)*self._delta_b)
