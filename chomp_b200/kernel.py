"""Drop-in for the hot-path part of the reference's kernel.py: dNdzGaussian /
dNdzMagLim (kernel.py:26-179), WindowFunctionGalaxy / WindowFunctionConvergence
(:211-484), Kernel / GalaxyGalaxyLensingKernel (:559-839), numerics on the GPU.

Not provided (outside BASELINE.json's configs):
dNdChiGaussian, the delta-function and flat windows, KernelGalaxyDelta,
KernelCovariance.  ``force_quad`` is accepted and ignored:
every integral is a converged fixed-order rule.  Kernel.__init__ does not write
the reference's debug files ``test_window_before/after`` (kernel.py:606, 608).
"""
import copy

import numpy as np

from . import _facade, _lib, cosmology, defaults


class dNdz(object):
    """Base class of the redshift distributions (kernel.py:26-86)."""
    _kind = None

    def __init__(self, z_min, z_max):
        self.z_min = z_min
        self.z_max = z_max
        self.norm = 1.0
        self.normalize()

    def _device(self):
        gpu = getattr(self, "_gpu", None)
        if gpu is None:
            gpu = self._gpu = _facade.OnePoint()
        cfg = _facade.base_config(zk_min=0.0, zk_max=max(float(self.z_max), 1e-3))
        for i in range(2):
            cfg.window_kind[i], cfg.dndz_kind[i] = _lib.WINDOW_GALAXY, self._kind
            cfg.dndz_zmin[i], cfg.dndz_zmax[i] = float(self.z_min), float(self.z_max)
            for j, v in enumerate(self._params()):
                cfg.dndz_p[i][j] = float(v)
        self._upload(gpu.eng, 2)
        gpu.configure(cfg)
        gpu.eng.limber_tables(_facade.cosmo_row(defaults.default_cosmo_dict))
        return gpu

    def _upload(self, eng, which):
        """Hook for distributions that carry a table (dNdzInterpolation)."""

    def normalize(self):
        self.norm = float(self._device().table(_lib.T_DNDZ_NORM)[0])

    def raw_dndz(self, redshift):
        return _facade.like_input(redshift, self._device().ev(_lib.EVAL_DNDZ_A, redshift, aux=1.0))

    def dndz(self, redshift):
        return _facade.like_input(redshift, self._device().ev(_lib.EVAL_DNDZ_A, redshift))

    def set_limits(self, z_min=None, z_max=None, calc_norm=False):
        if z_min is not None:
            self.z_min = z_min
        if z_max is not None:
            self.z_max = z_max
        if calc_norm:
            self.normalize()


class dNdzGaussian(dNdz):
    """exp(-(z - z0)^2 / (2 sigma_z^2)), clipped to z0 +- 8 sigma_z (kernel.py:89-112)."""
    _kind = _lib.DNDZ_GAUSSIAN

    def __init__(self, z_min, z_max, z0, sigma_z):
        if z_min < z0 - 8.0*sigma_z:
            z_min = z0 - 8.0*sigma_z
        if z_max > z0 + 8.0*sigma_z:
            z_max = z0 + 8.0*sigma_z
        self.z0 = z0
        self.sigma_z = sigma_z
        dNdz.__init__(self, z_min, z_max)

    def _params(self):
        return (self.z0, self.sigma_z, 0.0)


class dNdzMagLim(dNdz):
    """z^a exp(-(z/z0)^b) (kernel.py:148-179)."""
    _kind = _lib.DNDZ_MAGLIM

    def __init__(self, z_min, z_max, a, z0, b):
        self.a = a
        self.z0 = z0
        self.b = b
        # Python-2 semantics of ``1/b`` (kernel.py:164-168): integer division for an int b
        inv_b = (1//b) if isinstance(b, (int, np.integer)) else 1.0/b
        tmp_zmax = np.power(-1*np.log(defaults.default_precision["dNdz_precision"]), inv_b)*z0
        if tmp_zmax < z_max:
            print("WARNING:: z_max requested could result in failed normalization...")
            print("\tReseting z_max from %.2f to %.2f..." % (z_max, tmp_zmax))
            z_max = tmp_zmax
        dNdz.__init__(self, z_min, z_max)

    def _params(self):
        return (self.a, self.z0, self.b)


class dNdzInterpolation(dNdz):
    """p(z) from an array of redshifts and probabilities (kernel.py:181-208): the reference's FITPACK
    spline (order 2 by default), evaluated on the device from its piecewise-polynomial form."""
    _kind = _lib.DNDZ_TABLE

    def __init__(self, z_array, p_array, weights=None, interpolation_order=2, smoothing=None):
        from . import engine
        self._breaks, self._coef = engine.fitpack_piecewise(z_array, p_array, weights, interpolation_order,
                                                            smoothing)
        dNdz.__init__(self, z_array[0], z_array[-1])

    def _params(self):
        return (0.0, 0.0, 0.0)

    def _upload(self, eng, which):
        eng.set_dndz_table(which, self._breaks, self._coef)


class WindowFunction(object):
    """Window function tabulated in comoving distance (kernel.py:211-357)."""
    _kind = None

    def __init__(self, z_min, z_max, cosmo_multi_epoch=None, **kws):
        self.initialized_spline = False
        eps = defaults.default_precision["window_precision"]
        if z_min < eps:
            z_min = eps
        self.z_min = z_min
        self.z_max = z_max
        if cosmo_multi_epoch is None:
            cosmo_multi_epoch = cosmology.MultiEpoch(z_min, z_max)
        self._gpu = _facade.OnePoint()
        self.set_cosmology_object(cosmo_multi_epoch)

    def __copy__(self):
        other = self.__class__.__new__(self.__class__)
        other.__dict__.update(self.__dict__)
        other._gpu = _facade.OnePoint()
        other.initialized_spline = False
        return other

    def get_cosmology(self):
        return self.cosmo.get_cosmology()

    def set_cosmology_object(self, cosmo_multi_epoch):
        """kernel.py:289-306: own copy of the cosmology, re-gridded on the window's z range."""
        self.cosmo = copy.copy(cosmo_multi_epoch)
        self.cosmo.set_redshift(self.z_min, self.z_max)
        eps = defaults.default_precision["window_precision"]
        self.chi_min = max(float(self.cosmo.comoving_distance(self.z_min)), eps)
        self.chi_max = float(self.cosmo.comoving_distance(self.z_max))
        self._chi_array = np.linspace(self.chi_min, self.chi_max, defaults.default_precision["window_npoints"])
        self.initialized_spline = False

    def _initialize_spline(self):
        cfg = _facade.base_config(zk_min=float(self.cosmo.z_min), zk_max=float(self.cosmo.z_max))
        _facade.set_window(cfg, 0, self)
        _facade.set_window(cfg, 1, self)
        self._gpu.configure(cfg)
        self._gpu.eng.limber_tables(_facade.cosmo_row(self.cosmo.get_cosmology()))
        self._wf_array = self._gpu.table(_lib.T_WINDOW_NODES)[:cfg.n_window].copy()
        self.initialized_spline = True

    def window_function(self, chi):
        if not self.initialized_spline:
            self._initialize_spline()
        return _facade.like_input(chi, self._gpu.ev(_lib.EVAL_WINDOW_A, chi))

    def write(self, output_file_name):
        if not self.initialized_spline:
            self._initialize_spline()
        with open(output_file_name, "w") as f:
            f.write("#ttype1 = chi [Mpc/h]/n#ttype2 = window function value\n")
            for chi, wf in zip(self._chi_array, self._wf_array):
                f.write("%1.10f %1.10f\n" % (chi, wf))


class WindowFunctionGalaxy(WindowFunction):
    """W(chi) = dN/dz dz/dchi (kernel.py:360-387)."""
    _kind = _lib.WINDOW_GALAXY

    def __init__(self, redshift_dist, cosmo_multi_epoch=None, **kws):
        self._redshift_dist = redshift_dist
        self._redshift_dist.normalize()
        WindowFunction.__init__(self, redshift_dist.z_min, redshift_dist.z_max, cosmo_multi_epoch)


class WindowFunctionConvergence(WindowFunction):
    """W(chi) = 3/2 Omega_m H0^2 chi g(chi) / a (kernel.py:409-484)."""
    _kind = _lib.WINDOW_CONVERGENCE

    def __init__(self, redshift_dist, cosmo_multi_epoch=None, **kws):
        self._redshift_dist = redshift_dist
        self._redshift_dist.normalize()
        WindowFunction.__init__(self, 0.0, redshift_dist.z_max, cosmo_multi_epoch, **kws)


class Kernel(object):
    """K(ln k theta) = int dchi W_a W_b D^2 J_0(k theta chi) (kernel.py:559-781)."""
    _bessel_order = 0

    def __init__(self, ktheta_min, ktheta_max, window_function_a, window_function_b,
                 cosmo_multi_epoch=None, force_quad=False, **kws):
        self.initialized_spline = False
        self.ln_ktheta_min = np.log(ktheta_min)
        self.ln_ktheta_max = np.log(ktheta_max)
        self._ktheta = (float(ktheta_min), float(ktheta_max))
        self.window_function_a = copy.copy(window_function_a)
        self.window_function_b = copy.copy(window_function_b)
        self.z_min = np.max([self.window_function_a.z_min, self.window_function_b.z_min])
        self.z_max = np.min([self.window_function_a.z_max, self.window_function_b.z_max])
        if cosmo_multi_epoch is None:
            cosmo_multi_epoch = cosmology.MultiEpoch(self.z_min, self.z_max)
        self.cosmo = cosmo_multi_epoch
        self._force_quad = force_quad
        self._ln_ktheta_array = np.linspace(self.ln_ktheta_min, self.ln_ktheta_max,
                                            defaults.default_precision["kernel_npoints"])
        self._gpu = _facade.OnePoint()
        self._rebuild()

    def _config(self):
        from . import engine
        cfg = _facade.base_config(zk_min=float(self.cosmo.z_min), zk_max=float(self.cosmo.z_max),
                                  bessel_order=self._bessel_order)
        _facade.set_window(cfg, 0, self.window_function_a)
        _facade.set_window(cfg, 1, self.window_function_b)
        cfg.ktheta_min, cfg.ktheta_max = self._ktheta
        cfg.bessel_limit = engine.bessel_limit(self._bessel_order,
                                               defaults.default_precision["kernel_bessel_limit"])
        return cfg

    def _rebuild(self):
        """kernel.py:598-633 + 641-647: windows on this cosmology, chi range, z_bar and the
        K table -- one launch of the Limber stage."""
        self.window_function_a.set_cosmology_object(self.cosmo)
        self.window_function_b.set_cosmology_object(self.cosmo)
        self._gpu.configure(self._config())
        self._gpu.eng.limber_tables(_facade.cosmo_row(self.cosmo.get_cosmology()))
        self.chi_min, self.chi_max = (float(v) for v in self._gpu.table(_lib.T_KERNEL_CHI))
        self.z_bar = float(self._gpu.table(_lib.T_ZBAR)[0])
        self._kernel_array = self._gpu.table(_lib.T_KERNEL_NODES).copy()
        self.initialized_spline = True

    def get_cosmology(self):
        return self.cosmo.get_cosmology()

    def set_cosmology(self, cosmo_dict):
        self.cosmo.set_cosmology(cosmo_dict)
        self._rebuild()

    def kernel(self, ln_ktheta):
        return _facade.like_input(ln_ktheta, self._gpu.ev(_lib.EVAL_KERNEL, ln_ktheta))

    def write(self, output_file_name):
        with open(output_file_name, "w") as f:
            f.write("#ttype1 = k*theta [h/Mpc*Radians]\n#ttype2 = kernel [(h/Mpc)^2]\n")
            for x, kv in zip(self._ln_ktheta_array, self._kernel_array):
                f.write("%1.10g %1.10g\n" % (np.exp(x), kv))


class GalaxyGalaxyLensingKernel(Kernel):
    """Same with J_2 (kernel.py:784-839)."""
    _bessel_order = 2
