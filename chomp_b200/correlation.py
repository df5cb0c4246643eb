"""Drop-in for the reference's correlation.Correlation (correlation.py:33-289):
w(theta) = int dln k k^2/(2 pi) P(k)/D(z_bar)^2 K(ln k theta), all four GPU
stages per call.  CorrelationFourier / Correlation3d are listed as next in
SURVEY.md section 8(f)."""
import numpy as np

from . import _facade, _lib, engine, halo as halo_module

deg_to_rad = np.pi/180.0
rad_to_deg = 180.0/np.pi


class Correlation(object):
    def __init__(self, theta_min_deg, theta_max_deg, input_kernel, bins_per_decade=5.0, input_halo=None,
                 power_spec=None, k_min=None, k_max=None, keep_halo_z_bar=False, **kws):
        self.log_theta_min = np.log10(theta_min_deg*deg_to_rad)
        self.log_theta_max = np.log10(theta_max_deg*deg_to_rad)
        self.theta_array = engine.theta_bins(theta_min_deg, theta_max_deg, bins_per_decade)
        self.wtheta_array = np.zeros(self.theta_array.size)
        self.kernel = input_kernel
        self.D_z = float(self.kernel.cosmo.growth_factor(self.kernel.z_bar))
        if input_halo is None:
            input_halo = halo_module.Halo(self.kernel.z_bar)
        self.halo = input_halo
        if not keep_halo_z_bar:
            self.halo.set_redshift(self.kernel.z_bar)            # correlation.py:103
        # correlation.py:104-112.  The reference's test runs under Python 2, where ``None < x`` is True
        # and ``None > x`` is False: giving only k_max switches the extrapolation on whatever its value.
        if ((k_min is not None or k_max is not None) and not self.halo.get_extrapolation() and
                ((k_min is None or k_min < self.halo._k_min) or
                 (k_max is not None and k_max > self.halo._k_max))):
            self.halo.set_extrapolation(True)
        if k_min is None:
            k_min = self.halo._k_min
        if k_max is None:
            k_max = self.halo._k_max
        self._k_limits = (float(k_min), float(k_max))
        self._ln_k_min = np.log(k_min)
        self._ln_k_max = np.log(k_max)
        if power_spec is None:
            power_spec = "linear_power"
        self.set_power_spectrum(power_spec)

    def get_redshift(self):
        return self.kernel.z_bar

    def set_redshift(self, redshift):
        self.kernel.z_bar = redshift
        self.D_z = float(self.kernel.cosmo.growth_factor(self.kernel.z_bar))
        self.halo.set_redshift(self.kernel.z_bar)

    def get_cosmology(self):
        return self.kernel.get_cosmology()

    def set_cosmology(self, cosmo_dict):
        self.kernel.set_cosmology(cosmo_dict)
        self.D_z = float(self.kernel.cosmo.growth_factor(self.kernel.z_bar))
        self.halo.set_cosmology(cosmo_dict, self.kernel.z_bar)

    def get_power_spectrum(self):
        return self._power_name

    def set_power_spectrum(self, powSpec):
        if powSpec not in _lib.POWER_SPEC or not hasattr(self.halo, powSpec):
            print("WARNING: Invalid input for power spectra variable,")
            print("\t setting to 'linear_power'")
            powSpec = "linear_power"
        self._power_name = powSpec
        self.power_spec = getattr(self.halo, powSpec)

    def get_halo(self):
        return self.halo.get_halo()

    def set_halo(self, halo_dict):
        self.halo.set_halo(halo_dict)

    def get_hod(self, return_object=False):
        return self.halo.get_hod(return_object)

    def set_hod(self, hod_dict):
        self.halo.set_hod(hod_dict)

    def set_hod_object(self, input_hod):
        self.halo.set_hod_object(input_hod)

    def compute_correlation(self):
        self.wtheta_array = np.array(self.correlation(self.theta_array), dtype=float)

    def correlation(self, theta_rad):
        """correlation.py:242-268.  The halo object holds its tables on its own handle, so the
        Limber stage is replayed there (same configuration as the kernel object) before the
        Hankel stage."""
        gpu = self._stage_on_halo_handle()                      # also covers Correlation.set_redshift
        w = gpu.eng.wtheta_stage(1, _lib.POWER_SPEC[self._power_name], _facade.flat(theta_rad))
        return _facade.like_input(theta_rad, w.cpu().numpy()[0])

    def correlation_batch(self, cosmo_dicts, halo_dicts=None, hod_dicts=None, theta_rad=None):
        """w(theta) for a whole batch of parameter points in one pass of the four GPU stages:
        what ``for d in dicts: corr.set_cosmology(d); corr.set_halo(..); corr.set_hod(..);
        corr.compute_correlation()`` (examples/example_script.py:141-143,
        simulation_design.py:116-138) computes point by point -- each point's z_bar comes from its
        own cosmology, as Correlation.set_cosmology does.  Entries of halo_dicts / hod_dicts may
        be None (or the lists omitted) to keep this object's current values.
        Returns (w [B, n_theta], status [B])."""
        h = self.halo
        B = len(cosmo_dicts)
        halo_dicts = halo_dicts or [None]*B
        hod_dicts = hod_dicts or [None]*B
        cur_halo, cur_hod = h.get_halo(), h.get_hod()
        kind = h.local_hod._kind
        cosmo = engine.pack_params(cosmo_dicts, _lib.COSMO_KEYS)
        halo = engine.pack_params([d if d is not None else cur_halo for d in halo_dicts], _lib.HALO_KEYS)
        keys = _lib.HOD_ZHENG_KEYS if kind == _lib.HOD_ZHENG else _lib.HOD_MANDELBAUM_KEYS
        hod = np.zeros((B, _lib.N_HOD))
        for i, d in enumerate(hod_dicts):
            d = d if d is not None else cur_hod
            hod[i, :len(keys)] = [d[k] for k in keys]
        h._ensure()
        cfg = self.kernel._config()
        hc = h._gpu.eng.cfg
        for name in ("hod_kind", "exclusion", "extrapolate", "halo_precision", "use_halofit", "tri_moment", "with_bao", "mass_function_kind"):
            setattr(cfg, name, getattr(hc, name))
        self._set_k_limits(cfg)
        gpu = getattr(self, "_batch_gpu", None)
        if gpu is None:
            gpu = self._batch_gpu = _facade.OnePoint()
        gpu.configure(cfg)
        theta = self.theta_array if theta_rad is None else _facade.flat(theta_rad)
        import torch
        status = torch.zeros(B, dtype=torch.int32, device="cuda:%d" % gpu.eng.device)
        which = _lib.POWER_SPEC[self._power_name]
        if cfg.use_halofit:
            # HaloFit (halo.py:1236-1412): the fit parameters follow each point's mass tables; f_1..f_3
            # stay at the halo object's construction redshift, as in Halo._ensure
            eng = gpu.eng
            eng.limber_tables(cosmo, status=status)
            eng.mass_tables(cosmo, halo, status=status)
            eng.halofit(B, fit_z=float(getattr(h, "_fit_redshift", -1.0)), status=status)
            if which in (_lib.P_GM, _lib.P_GG):
                eng.halo_tables(halo, hod, status=status)
            w = eng.wtheta_stage(B, which, theta, status=status)
        else:
            w = gpu.eng.wtheta(cosmo, halo, hod, theta, which, status=status)
        return w.cpu().numpy(), status.cpu().numpy()

    def _stage_on_halo_handle(self):
        """Replay the Limber stage on the handle that holds the halo's tables and pin z_bar."""
        h = self.halo
        h._ensure()
        cfg = self.kernel._config()
        hc = h._gpu.eng.cfg
        for name in ("hod_kind", "exclusion", "extrapolate", "halo_precision", "use_halofit", "tri_moment", "with_bao", "mass_function_kind"):
            setattr(cfg, name, getattr(hc, name))
        self._set_k_limits(cfg)
        gpu = h._gpu
        gpu.configure(cfg)
        gpu.eng.limber_tables(_facade.cosmo_row(self.kernel.cosmo.get_cosmology()))
        gpu.eng.set_params(cosmo=_facade.cosmo_row(h.cosmo.cosmo_dict))
        gpu.eng.set_zbar([self.kernel.z_bar])
        return gpu

    def _set_k_limits(self, cfg):
        """Correlation(k_min=, k_max=): integration limits of the Hankel stage (<= 0: the halo's)."""
        k_lo, k_hi = getattr(self, "_k_limits", (self.halo._k_min, self.halo._k_max))
        cfg.corr_k_min = k_lo if k_lo != self.halo._k_min else -1.0
        cfg.corr_k_max = k_hi if k_hi != self.halo._k_max else -1.0

    def write(self, output_file_name):
        with open(output_file_name, "w") as f:
            f.write("#ttype1 = theta [deg]\n#ttype2 = wtheta\n")
            for theta, w in zip(self.theta_array, self.wtheta_array):
                f.write("%1.10g %1.10g\n" % (theta/deg_to_rad, w))


class CorrelationFourier(Correlation):
    """C(l) = int dchi P(l/chi)/D(z_bar)^2 W_a W_b D^2 / chi^2 (correlation.py:297-405).
    Any of the spectra: closed-form (linear_power, HaloFit power_mm) or table-based."""

    def __init__(self, l_min, l_max, input_kernel, input_halo=None, powSpec=None, **kws):
        from . import defaults
        self.log_l_min = np.log10(l_min)
        self.log_l_max = np.log10(l_max)
        self.l_array = np.logspace(self.log_l_min, self.log_l_max, defaults.default_precision["corr_npoints"])
        if l_min == l_max:
            self.l_array = np.array([l_min])
        self.power_array = np.zeros(self.l_array.size, dtype="float64")
        self.kernel = input_kernel
        self.D_z = float(self.kernel.cosmo.growth_factor(self.kernel.z_bar))
        if input_halo is None:
            input_halo = halo_module.Halo(self.kernel.z_bar)
        self.halo = input_halo
        self.halo.set_redshift(self.kernel.z_bar)               # correlation.py:342
        if powSpec is None:
            powSpec = "linear_power"
        self.set_power_spectrum(powSpec)

    def compute_correlation(self):
        self.power_array = np.array(self.correlation(self.l_array), dtype=float)

    def correlation(self, l):
        gpu = self._stage_on_halo_handle()
        out = gpu.eng.cl(1, _lib.POWER_SPEC[self._power_name], _facade.flat(l))
        return _facade.like_input(l, out.cpu().numpy()[0])

    def write(self, output_file_name):
        with open(output_file_name, "w") as f:
            f.write("#ttype1 = l [deg]\n#ttype2 = power\n")
            for l, power in zip(self.l_array, self.power_array):
                f.write("%1.10f %1.10f\n" % (l, power))


class Correlation3d(Correlation):
    """correlation.py:408-510: xi(r) = int dln k k^2/(2 pi) P(k) J0(k r) on ``corr_npoints`` log-spaced radii, splined
    in r.  The integration limits are the halo model's k range."""

    def __init__(self, r_min, r_max, redshift=0.0, input_halo=None, powSpec=None, k_min=None, k_max=None):
        from . import defaults
        self.log_r_min = np.log10(r_min)
        self.log_r_max = np.log10(r_max)
        self.r_array = np.logspace(self.log_r_min, self.log_r_max, defaults.default_precision["corr_npoints"])
        if r_min == r_max:
            self.r_array = np.array([r_min])
        self.xi_array = np.zeros(self.r_array.size)
        if input_halo is None:
            input_halo = halo_module.Halo(redshift)
        self.halo = input_halo
        self.halo.set_redshift(redshift)
        if ((k_min is not None and k_min != self.halo._k_min) or (k_max is not None and k_max != self.halo._k_max)):
            raise NotImplementedError("Correlation3d(k_min/k_max) different from the halo limits")
        self._ln_k_min, self._ln_k_max = np.log(self.halo._k_min), np.log(self.halo._k_max)
        if powSpec is None:
            powSpec = "linear_power"
        self.set_power_spectrum(powSpec)
        self.initialized_spline = False

    def raw_correlation(self, r):
        h = self.halo
        h._ensure()
        out = h._gpu.eng.xi3d(1, _lib.POWER_SPEC[self._power_name], _facade.flat(r))
        return _facade.like_input(r, out.cpu().numpy()[0])

    def compute_correlation(self):
        self.xi_array = np.array(self.raw_correlation(self.r_array), dtype=float)
        if self.r_array.size > 3:
            self._xi_breaks, self._xi_coef = engine.interpolating_spline_piecewise(self.r_array, self.xi_array, 3)
        self.initialized_spline = True

    def correlation(self, r):
        """correlation.py:502-510: the spline inside (r_min, r_max], 0 outside."""
        if not self.initialized_spline:
            self.compute_correlation()
        rr = _facade.flat(r)
        i = np.clip(np.searchsorted(self._xi_breaks, rr, side="right") - 1, 0, len(self._xi_breaks) - 2)
        t = rr - self._xi_breaks[i]
        c = self._xi_coef[i]
        val = c[:, 0] + t*(c[:, 1] + t*(c[:, 2] + t*c[:, 3]))
        r_min, r_max = 10.0**self.log_r_min, 10.0**self.log_r_max
        return _facade.like_input(r, np.where((rr <= r_max) & (rr > r_min), val, 0.0))

    def write(self, output_file_name):
        with open(output_file_name, "w") as f:
            f.write("#ttype1 = r [Mpc/h]\n#ttype2 = xi\n")
            for r, xi in zip(self.r_array, self.xi_array):
                f.write("%1.10f %1.10f\n" % (r, xi))
