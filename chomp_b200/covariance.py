"""Drop-in for the reference's covariance.Covariance (covariance.py:23-683) and AnnulusBin
(:1085-1103): Poisson + Gaussian + 1-halo non-Gaussian covariance of w(theta) between annular bins,
computed for all bin pairs by one call of the batched C-ABI entry point (a batch of one point).

Supported: ``input_correlation_a is input_correlation_b`` (the reference's ``matching_corrs``, the case of its
examples and of BASELINE config 5) and two DIFFERENT correlations (four windows, two halo models, the projected
cross spectra of covariance.py:555-591; the blocks of CovarianceMulti, :794-871), J0 kernels, Halo / HaloExclusion /
HaloFit halos.  In the matching case the trispectrum object shares the correlation's halo / HOD parameters (its
redshift and moment kind are its own); between different correlations it carries its own.
Super-sample covariance (``ssc_cov``) is listed as next in SURVEY.md 8(f)."""
import numpy as np

from . import _facade, _lib, defaults, engine, halo_trispectrum

deg_to_rad = np.pi/180.0
rad_to_deg = 180.0/np.pi
deg2_to_strad = deg_to_rad*deg_to_rad
strad_to_deg2 = rad_to_deg*rad_to_deg


class AnnulusBin(object):
    """covariance.py:1085-1103."""

    def __init__(self, inner, outer):
        self.inner = inner
        self.outer = outer
        self.center = np.power(10.0, 0.5*(np.log10(inner) + np.log10(outer)))
        self.delta = outer - inner


class _Setup(object):
    pass


class Covariance(object):
    def __init__(self, input_correlation_a, input_correlation_b, bins_per_decade=5.0, survey_area_deg2=20,
                 n_a=1.0e4, n_b=1.0e4, variance=1.0, nongaussian_cov=True, input_halo_trispectrum=None,
                 power_spec="power_mm", poisson_noise_only=False, ssc_cov=False, **kws):
        if ssc_cov:
            raise NotImplementedError("super-sample covariance is not on the GPU path yet")
        self.corr_a, self.corr_b = input_correlation_a, input_correlation_b
        # covariance.py:60-63 compares the objects' dictionaries; two separately built correlations never compare
        # equal there (their kernels hold distinct window / cosmology copies)
        self.matching_corrs = input_correlation_a is input_correlation_b
        self.log_theta_min = input_correlation_a.log_theta_min
        self.log_theta_max = input_correlation_a.log_theta_max
        self.bins_per_decade = bins_per_decade
        theta_deg = (np.power(10.0, self.log_theta_min)*rad_to_deg, np.power(10.0, self.log_theta_max)*rad_to_deg)
        self._theta_deg = theta_deg
        rows = engine.annular_bins(theta_deg[0], theta_deg[1], bins_per_decade)
        self.annular_bins = [AnnulusBin(r[0], r[1]) for r in rows]
        self.area = survey_area_deg2*deg2_to_strad
        self._survey_area_deg2 = survey_area_deg2
        self._n_a, self._n_b = n_a, n_b
        self.variance = variance
        self.nongaussian_cov = nongaussian_cov
        self.poisson_noise_only = poisson_noise_only
        self.ssc_cov = False
        self.halo_a, self.halo_b = input_correlation_a.halo, input_correlation_b.halo
        if input_halo_trispectrum is None:
            input_halo_trispectrum = halo_trispectrum.HaloTrispectrumOneHalo()
        self.halo_tri = input_halo_trispectrum
        if power_spec is None:
            power_spec = "linear_power"
        if power_spec not in _lib.POWER_SPEC or not hasattr(self.halo_a, power_spec):
            print("WARNING: Invalid input for power spectra variable,")      # covariance.py:184-191
            print("\t setting to linear_power")
            power_spec = "linear_power"
        self.power_spec = power_spec
        self._gpu = _facade.OnePoint()
        self._parts = None

    # ---- the reference's setters -----------------------------------------------------------------
    def set_cosmology(self, cosmo_dict):
        """covariance.py:241-262: both correlations and the trispectrum (moved to z_bar_NG) follow."""
        self.corr_a.set_cosmology(cosmo_dict)
        if not self.matching_corrs:
            self.corr_b.set_cosmology(cosmo_dict)
        self.halo_a, self.halo_b = self.corr_a.halo, self.corr_b.halo
        self._tri_follows_z_bar_ng = True
        self.halo_tri.cosmo_dict = cosmo_dict
        self._parts = None

    def get_cosmology(self):
        return self.corr_a.kernel.get_cosmology()

    # ---- evaluation ---------------------------------------------------------------------------------
    def _config_for(self, corr):
        """Limber set-up of the correlation's kernel + the halo-model switches of its halo object."""
        h = corr.halo
        h._ensure()
        cfg = corr.kernel._config()
        hc = h._gpu.eng.cfg
        for name in ("hod_kind", "exclusion", "extrapolate", "halo_precision", "use_halofit", "with_bao", "mass_function_kind"):
            setattr(cfg, name, getattr(hc, name))
        return cfg

    def _setup_for(self, cfg):
        survey = _Setup()
        survey.quadrature = dict(defaults.default_quadrature)
        survey.limits = dict(defaults.default_limits)
        survey.precision = dict(defaults.default_precision)
        survey.window = (cfg.window_kind[0], cfg.window_kind[1])
        setup = engine.CovarianceSetup(survey, self._theta_deg, self.bins_per_decade, self._survey_area_deg2,
                                       self._n_a, self._n_b, self.variance, self.nongaussian_cov, self.power_spec,
                                       self.poisson_noise_only)
        # the theta range is the correlation's own log10 values (covariance.py:93-97)
        setup.params.theta_min_rad = np.power(10.0, self.log_theta_min)
        setup.params.theta_max_rad = np.power(10.0, self.log_theta_max)
        return setup

    def _evaluate(self):
        if self._parts is not None:
            return
        if not self.matching_corrs:
            return self._evaluate_cross()
        corr, h = self.corr_a, self.corr_a.halo
        cfg = self._config_for(corr)
        cfg.tri_moment = _lib.TRISPECTRUM_MOMENT.get(self.halo_tri.power_spec, 0)
        tri = self.halo_tri
        if (self.nongaussian_cov and not self.poisson_noise_only and
                (tri.local_hod._kind != h.local_hod._kind or tri.local_hod._params() != h.local_hod._params() or
                 dict(tri.mass.halo_dict) != dict(h.mass.halo_dict))):
            # a trispectrum object with halo / HOD parameters of its own: the two-handle path carries them (the same
            # correlation on both handles gives the matching case's Gaussian term; the Poisson term is kept)
            return self._evaluate_cross(matching=True)
        self._gpu.configure(cfg)
        setup = self._setup_for(cfg)
        setup.params.halofit_z = float(getattr(h, "_fit_redshift", -1.0))
        self.equal_windows, self.density, self.cosmic_shear = setup.equal_windows, setup.density, setup.cosmic_shear
        tri_z = None if getattr(self, "_tri_follows_z_bar_ng", False) else [float(tri._redshift)]
        eng = self._gpu.eng
        out, parts = eng.covariance(_facade.cosmo_row(corr.kernel.cosmo.get_cosmology()),
                                    _facade.halo_row(h.mass.halo_dict, h._profile),
                                    _facade.hod_row(h.local_hod._kind, h.local_hod._params()), setup, tri_z=tri_z,
                                    parts=True)
        self.covar = out.cpu().numpy()[0]
        self._parts = parts.cpu().numpy()[0]
        self.D_z_NG = float(eng.table(_lib.T_D_NG, 1)[0, 0]) if setup.params.nongaussian and not setup.params.poisson_only else None
        self._centers = np.array([b.center for b in self.annular_bins])

    def _evaluate_cross(self, matching=False):
        """Two different correlations: one engine per correlation, a third for the trispectrum object."""
        ca, cb, tri = self.corr_a, self.corr_b, self.halo_tri
        ha, hb = ca.halo, cb.halo
        cfg_a, cfg_b = self._config_for(ca), self._config_for(cb)
        cfg_t = self._config_for(ca)
        cfg_t.hod_kind = tri.local_hod._kind
        cfg_t.halo_precision = getattr(tri.local_hod, "_halo_precision", cfg_t.halo_precision)
        cfg_t.exclusion, cfg_t.use_halofit = 0, 0
        cfg_t.tri_moment = _lib.TRISPECTRUM_MOMENT.get(tri.power_spec, 0)
        if getattr(self, "_gpu_b", None) is None:
            self._gpu_b, self._gpu_t = _facade.OnePoint(), _facade.OnePoint()
        self._gpu.configure(cfg_a)
        self._gpu_b.configure(cfg_b)
        self._gpu_t.configure(cfg_t)
        setup = self._setup_for(cfg_a)
        setup.params.halofit_z = float(getattr(ha, "_fit_redshift", getattr(hb, "_fit_redshift", -1.0)))
        if not matching:
            for i in range(6):                   # no window of one correlation is a window of the other
                setup.params.poisson[i] = 0.0
        conv = [w == _lib.WINDOW_CONVERGENCE for w in (cfg_a.window_kind[0], cfg_a.window_kind[1], cfg_b.window_kind[0],
                                                         cfg_b.window_kind[1])]
        self.equal_windows = setup.equal_windows if matching else [False]*6
        self.density = setup.density
        self.cosmic_shear = [bool(conv[0]*conv[1] or conv[2]*conv[3]), bool(conv[0]*conv[3] or conv[1]*conv[2])]
        tri_z = None if getattr(self, "_tri_follows_z_bar_ng", False) else [float(tri._redshift)]
        out, parts = self._gpu.eng.covariance_cross(
            self._gpu_b.eng, _facade.cosmo_row(ca.kernel.cosmo.get_cosmology()),
            _facade.halo_row(ha.mass.halo_dict, ha._profile), _facade.hod_row(ha.local_hod._kind, ha.local_hod._params()),
            _facade.halo_row(hb.mass.halo_dict, hb._profile), _facade.hod_row(hb.local_hod._kind, hb.local_hod._params()),
            setup, tri_engine=self._gpu_t.eng, halo_t=_facade.halo_row(tri.mass.halo_dict, tri._profile),
            hod_t=_facade.hod_row(tri.local_hod._kind, tri.local_hod._params()), tri_z=tri_z, parts=True)
        self.covar = out.cpu().numpy()[0]
        self._parts = parts.cpu().numpy()[0]
        self.D_z_NG = (float(self._gpu.eng.table(_lib.T_D_NG, 1)[0, 0])
                       if setup.params.nongaussian and not setup.params.poisson_only else None)
        self._centers = np.array([b.center for b in self.annular_bins])

    def get_covariance(self):
        """covariance.py:276-295."""
        self._evaluate()
        return self.covar

    def _index(self, theta):
        self._evaluate()
        i = int(np.argmin(np.abs(np.log(self._centers) - np.log(theta))))
        if abs(self._centers[i] - theta) > 1e-9*theta:
            raise NotImplementedError("covariance terms are tabulated at this object's annular bins only")
        return i

    def covariance(self, annular_bin_a, annular_bin_b):
        """covariance.py:297-321."""
        return float(self.covar[self._index(annular_bin_a.center), self._index(annular_bin_b.center)])

    def covariance_P(self, delta, theta, window_1=0, window_2=1):
        i = self._index(theta)
        return float(self._parts[0][i, i])

    def covariance_G(self, theta_a, theta_b, delta_a=None, delta_b=None):
        return float(self._parts[1][self._index(theta_a), self._index(theta_b)])

    def covariance_NG(self, theta_a_rad, theta_b_rad):
        return float(self._parts[2][self._index(theta_a_rad), self._index(theta_b_rad)])

    def write(self, output_file_name):
        cov = np.asarray(self.get_covariance(), dtype=float)
        with open(output_file_name, "w") as f:
            for row in cov:
                f.write(" ".join("%1.10g" % v for v in row) + "\n")


class CovarianceMulti(Covariance):
    """covariance.py:794-871: the auto-covariances of a list of correlations and the cross-covariances of every pair,
    assembled into one (n_corr n_bins) x (n_corr n_bins) matrix.  As in the reference every block is a Covariance
    built with the default ``power_spec`` ('power_mm') and the shared survey / trispectrum arguments."""

    def __init__(self, correlation_object_list, bins_per_decade=5, survey_area_deg2=4*np.pi*strad_to_deg2, n_a=1e6, n_b=1e6,
                 variance=1.0, nongaussian_cov=True, input_halo_trispectrum=None, poisson_noise_only=False, **kws):
        self.covariance_list = []
        n = len(correlation_object_list)
        for idx1 in range(n):
            row = []
            for idx2 in range(idx1, n):
                row.append(Covariance(input_correlation_a=correlation_object_list[idx1],
                                      input_correlation_b=correlation_object_list[idx2], bins_per_decade=bins_per_decade,
                                      survey_area_deg2=survey_area_deg2, n_a=n_a, n_b=n_b, variance=variance,
                                      nongaussian_cov=nongaussian_cov, input_halo_trispectrum=input_halo_trispectrum,
                                      poisson_noise_only=poisson_noise_only))
            self.covariance_list.append(row)
        self.annular_bins = self.covariance_list[0][0].annular_bins
        self.theta_bins = len(self.annular_bins)
        self.wcovar = np.empty((self.theta_bins*n, self.theta_bins*n))

    def get_covariance(self):
        """covariance.py:850-871 (block (idx1, idx1 + idx2) and its mirror image)."""
        nb = self.theta_bins
        for idx1, row in enumerate(self.covariance_list):
            for idx2, cov in enumerate(row):
                cov.get_covariance()
                r1, c1 = idx1*nb, (idx1 + idx2)*nb
                self.wcovar[r1:r1 + nb, c1:c1 + nb] = cov.covar
                self.wcovar[c1:c1 + nb, r1:r1 + nb] = cov.covar
        return self.wcovar
