"""Batched evaluation of the halo-model -> Limber -> Hankel path on one GPU.

`Engine` owns one C-ABI handle (device scratch + configuration) and exposes the
stages of include/chomp_b200.h on torch CUDA tensors.  `Survey` collects the
batch-invariant set-up that the reference spreads over constructor arguments:
the two redshift distributions / windows (kernel.py:89, 148, 372, 430), the
MultiEpoch range and k*theta range given to Kernel (kernel.py:584), the
angular bins of Correlation (correlation.py:65-90) and the module precision
dictionaries (defaults.py).

PyTorch is used for device memory and streams only.
"""
import ctypes
import math

import numpy as np

from . import _lib, defaults

# zeros of J0 and J2 (scipy.special.jn_zeros(order, 24)); the Limber integral stops
# at zero number defaults.default_precision["kernel_bessel_limit"] (kernel.py:627-629, 806-807)
_J0_ZEROS = (2.4048255576957724, 5.520078110286311, 8.653727912911013, 11.791534439014281,
             14.930917708487787, 18.071063967910924, 21.21163662987926, 24.352471530749302,
             27.493479132040253, 30.634606468431976, 33.77582021357357, 36.917098353664045,
             40.05842576462824, 43.19979171317673, 46.341188371661815, 49.482609897397815,
             52.624051841115, 55.76551075501998, 58.90698392608094, 62.048469190227166,
             65.18996480020687, 68.3314693298568, 71.47298160359374, 74.61450064370183)
_J2_ZEROS = (5.135622301840683, 8.417244140399866, 11.61984117214906, 14.795951782351262,
             17.959819494987826, 21.116997053021844, 24.2701123135731, 27.420573549984557,
             30.569204495516395, 33.7165195092227, 36.86285651128381, 40.008446733478195,
             43.153453778371464, 46.29799667723692, 49.442164110416876, 52.58602350681596,
             55.72962705320114, 58.87301577261216, 62.01622235921766, 65.1592731907578,
             68.30218978418345, 71.44498986635786, 74.5876881736024, 77.7302970569789)

DEG_TO_RAD = math.pi/180.0


def bessel_limit(order, n_zeros):
    zeros = _J0_ZEROS if order == 0 else _J2_ZEROS
    if not 1 <= n_zeros <= len(zeros):
        raise ValueError("kernel_bessel_limit must be in 1..%d" % len(zeros))
    return zeros[n_zeros - 1]


def theta_bins(theta_min_deg, theta_max_deg, bins_per_decade=5.0):
    """Geometric bin centres in radians, exactly as Correlation.__init__ builds
    them (correlation.py:69-90): edges at 10**(u/bpd) for integer u counted up
    from floor(log10 theta_min)*bpd, centre = geometric mean of the two edges,
    bins kept when theta_min <= lower edge < theta_max."""
    lo = math.log10(theta_min_deg*DEG_TO_RAD)
    hi = math.log10(theta_max_deg*DEG_TO_RAD)
    if theta_min_deg == theta_max_deg:
        return np.array([theta_min_deg*DEG_TO_RAD])
    bpd = 1.0*bins_per_decade
    t_lo, t_hi = np.power(10.0, lo), np.power(10.0, hi)
    centres = []
    u = np.floor(lo)*bins_per_decade
    edge = np.power(10.0, u/bpd)
    while edge < t_hi:
        if t_lo <= edge < t_hi:
            centres.append(10**(0.5*(np.log10(edge) + (u + 1.0)/bpd)))
        u += 1.0
        edge = np.power(10.0, u/bpd)
    return np.array(centres)


def _bspline_basis(t, k, x):
    """All B-splines of degree k on the knot vector t at the points x (Cox-de Boor): [len(x), len(t) - k - 1]."""
    x = np.asarray(x, dtype=np.float64)
    n = len(t) - k - 1
    # degree 0: indicator of [t_j, t_j+1), the last non-empty interval closed on the right
    B = np.zeros((len(x), len(t) - 1))
    last = np.max(np.nonzero(np.diff(t) > 0)[0])
    for j in range(len(t) - 1):
        if t[j + 1] > t[j]:
            hi = (x <= t[j + 1]) if j == last else (x < t[j + 1])
            B[:, j] = (x >= t[j]) & hi
    for d in range(1, k + 1):
        Bn = np.zeros((len(x), len(t) - 1 - d))
        for j in range(len(t) - 1 - d):
            left = t[j + d] - t[j]
            right = t[j + d + 1] - t[j + 1]
            if left > 0:
                Bn[:, j] += (x - t[j])/left*B[:, j]
            if right > 0:
                Bn[:, j] += (t[j + d + 1] - x)/right*B[:, j + 1]
        B = Bn
    return B[:, :n]


def interpolating_spline_piecewise(z_array, p_array, k=2):
    """The interpolating spline FITPACK builds for s = 0 (what InterpolatedUnivariateSpline is;
    reference kernel.py:197-200), restated: knots as in FITPACK's fpcurf -- for odd k the data
    abscissae without the (k + 1) / 2 outermost interior ones on each side ("not-a-knot"), for even
    k the mid-points of the interior data intervals -- and the B-spline coefficients from the
    collocation system.  Returned in piecewise-polynomial form: (breaks[n + 1], coef[n, 4])."""
    x = np.ascontiguousarray(z_array, dtype=np.float64)
    y = np.ascontiguousarray(p_array, dtype=np.float64)
    m = x.size
    if not 1 <= k <= 3:
        raise ValueError("interpolation_order must be 1, 2 or 3 (the device evaluates cubics)")
    if m <= k or np.any(np.diff(x) <= 0):
        raise ValueError("need more than k strictly increasing abscissae")
    if k % 2:
        interior = x[(k + 1)//2:m - (k + 1)//2]
    else:
        interior = 0.5*(x[k//2:m - k//2 - 1] + x[k//2 + 1:m - k//2])
    t = np.concatenate([[x[0]]*(k + 1), interior, [x[-1]]*(k + 1)])
    c = np.linalg.solve(_bspline_basis(t, k, x), y)
    breaks = np.unique(t)
    coef = np.zeros((breaks.size - 1, 4))
    # one polynomial of degree k per knot interval: exact fit through k + 1 of its own values
    u = 0.5*(1.0 - np.cos(np.pi*(np.arange(k + 1) + 0.5)/(k + 1)))           # Chebyshev points in (0, 1)
    V = np.vander(u, k + 1, increasing=True)
    for i in range(breaks.size - 1):
        h = breaks[i + 1] - breaks[i]
        vals = _bspline_basis(t, k, breaks[i] + h*u) @ c
        coef[i, :k + 1] = np.linalg.solve(V, vals)/h**np.arange(k + 1)
    return np.ascontiguousarray(breaks), np.ascontiguousarray(coef)


def fitpack_piecewise(z_array, p_array, weights=None, interpolation_order=2, smoothing=None):
    """dNdzInterpolation's spline (kernel.py:191-204) in piecewise-polynomial form: returns
    (breaks[n + 1], coef[n, 4]) with p(z) = sum_j coef[i, j] (z - breaks[i])**j.

    ``smoothing is None`` (the reference's default): the interpolating spline, built here
    (`interpolating_spline_piecewise`; positive weights do not change an interpolating spline).
    A smoothing spline (kernel.py:201-204) needs FITPACK's adaptive knot placement and is taken
    from scipy, as the reference does -- input preparation on a handful of numbers, once per survey."""
    if smoothing is None:
        if weights is not None and np.any(np.asarray(weights) <= 0):
            raise ValueError("weights must be positive")
        return interpolating_spline_piecewise(z_array, p_array, int(interpolation_order))
    try:
        from scipy.interpolate import PPoly, UnivariateSpline
    except ImportError as exc:   # pragma: no cover
        raise _lib.ChompError("a smoothing dNdzInterpolation needs scipy (FITPACK), as the reference does") from exc
    z_array = np.ascontiguousarray(z_array, dtype=np.float64)
    p_array = np.ascontiguousarray(p_array, dtype=np.float64)
    if not 1 <= int(interpolation_order) <= 3:
        raise ValueError("interpolation_order must be 1, 2 or 3 (the device evaluates cubics)")
    spl = UnivariateSpline(z_array, p_array, w=weights, k=interpolation_order, s=smoothing)
    pp = PPoly.from_spline(spl._eval_args)
    keep = np.diff(pp.x) > 0                      # drop the repeated end knots
    c = pp.c[::-1][:, keep]                       # ascending powers, [order + 1, n]
    coef = np.zeros((c.shape[1], 4))
    coef[:, :c.shape[0]] = c.T
    breaks = np.concatenate([pp.x[:-1][keep], pp.x[-1:]])
    return np.ascontiguousarray(breaks), np.ascontiguousarray(coef)


class RedshiftDistribution(object):
    """Parameters of a dNdzGaussian / dNdzMagLim after the constructors' range
    clipping (kernel.py:100-106, 160-175), or the table of a dNdzInterpolation."""

    def __init__(self, kind, z_min, z_max, p, breaks=None, coef=None):
        self.kind, self.z_min, self.z_max, self.p = kind, float(z_min), float(z_max), tuple(p)
        self.breaks, self.coef = breaks, coef

    @classmethod
    def table(cls, z_array, p_array, weights=None, interpolation_order=2, smoothing=None):
        """dNdzInterpolation(z_array, p_array, weights, interpolation_order, smoothing), kernel.py:191-205."""
        breaks, coef = fitpack_piecewise(z_array, p_array, weights, interpolation_order, smoothing)
        return cls(_lib.DNDZ_TABLE, z_array[0], z_array[-1], (0.0, 0.0, 0.0), breaks, coef)

    @classmethod
    def gaussian(cls, z_min, z_max, z0, sigma_z):
        z_min = max(z_min, z0 - 8.0*sigma_z)
        z_max = min(z_max, z0 + 8.0*sigma_z)
        return cls(_lib.DNDZ_GAUSSIAN, z_min, z_max, (z0, sigma_z, 0.0))

    @classmethod
    def maglim(cls, z_min, z_max, a, z0, b, dndz_precision=None):
        if dndz_precision is None:
            dndz_precision = defaults.default_precision["dNdz_precision"]
        # the reference computes 1/b under Python 2: integer division for int b (kernel.py:164-168)
        inv_b = (1//b) if isinstance(b, (int, np.integer)) else 1.0/b
        cap = (-1.0*math.log(dndz_precision))**inv_b*z0
        if cap < z_max:
            z_max = cap
        return cls(_lib.DNDZ_MAGLIM, z_min, z_max, (float(a), float(z0), float(b)))


class Survey(object):
    """Everything that is shared by all points of a batch."""

    def __init__(self, dist_a, dist_b=None, window_a="galaxy", window_b=None,
                 z_range=(0.0, 5.0), ktheta_range_deg=(1e-6, 100.0),
                 theta_deg=(0.001, 1.0), bins_per_decade=5.0, bessel_order=0,
                 power_spec="power_mm", hod="zheng", exclusion=False,
                 extrapolate=False, precision=None, limits=None, quadrature=None):
        self.dist = (dist_a, dist_b or dist_a)
        kinds = {"galaxy": _lib.WINDOW_GALAXY, "convergence": _lib.WINDOW_CONVERGENCE}
        self.window = (kinds[window_a], kinds[window_b or window_a])
        self.z_range = tuple(z_range)
        self.ktheta = (ktheta_range_deg[0]*DEG_TO_RAD, ktheta_range_deg[1]*DEG_TO_RAD)
        self.theta = theta_bins(theta_deg[0], theta_deg[1], bins_per_decade)
        self.bessel_order = int(bessel_order)
        self.power_spec = power_spec
        self.hod_kind = {"zheng": _lib.HOD_ZHENG, "mandelbaum": _lib.HOD_MANDELBAUM}[hod]
        self.exclusion, self.extrapolate = bool(exclusion), bool(extrapolate)
        self.precision = dict(precision or defaults.default_precision)
        self.limits = dict(limits or defaults.default_limits)
        self.quadrature = dict(quadrature or defaults.default_quadrature)

    def config(self):
        p, lim, q = self.precision, self.limits, self.quadrature
        c = _lib.Config()
        c.n_cosmo, c.n_mass, c.n_halo = p["cosmo_npoints"], p["mass_npoints"], p["halo_npoints"]
        c.n_window, c.n_kernel = p["window_npoints"], p["kernel_npoints"]
        c.nq_nu, c.nq_hankel, c.nq_limber, c.nq_lens = q["nu"], q["hankel"], q["limber"], q["lens"]
        c.hod_kind, c.bessel_order = self.hod_kind, self.bessel_order
        c.exclusion, c.extrapolate = int(self.exclusion), int(self.extrapolate)
        for i in range(2):
            c.window_kind[i] = self.window[i]
            c.dndz_kind[i] = self.dist[i].kind
            c.dndz_zmin[i], c.dndz_zmax[i] = self.dist[i].z_min, self.dist[i].z_max
            for j in range(3):
                c.dndz_p[i][j] = self.dist[i].p[j]
        c.halo_precision, c.cosmo_precision = p["halo_precision"], p["cosmo_precision"]
        c.window_precision = p["window_precision"]
        c.k_min, c.k_max = lim["k_min"], lim["k_max"]
        c.mass_min, c.mass_max = lim["mass_min"], lim["mass_max"]
        c.zk_min, c.zk_max = self.z_range
        c.ktheta_min, c.ktheta_max = self.ktheta
        c.bessel_limit = bessel_limit(self.bessel_order, p["kernel_bessel_limit"])
        c.corr_k_min = c.corr_k_max = -1.0
        c.tri_moment = -1          # >= 0 also builds the 1-halo trispectrum's node list
        return c


def annular_bins(theta_min_deg, theta_max_deg, bins_per_decade=5.0):
    """(inner, outer, center, delta) of the AnnulusBin list Covariance.__init__ builds
    (covariance.py:53-74, 1085-1103), radians, shape [n_bins, 4]."""
    lo = np.log10(theta_min_deg*DEG_TO_RAD)
    hi = np.log10(theta_max_deg*DEG_TO_RAD)
    u = np.floor(lo)*bins_per_decade
    t = np.power(10.0, u/(1.0*bins_per_decade))
    rows = []
    while t < np.power(10.0, hi):
        if t >= np.power(10.0, lo) and t < np.power(10.0, hi):
            outer = np.power(10.0, (u + 1.0)/(1.0*bins_per_decade))
            rows.append((t, outer, np.power(10.0, 0.5*(np.log10(t) + np.log10(outer))), outer - t))
        u += 1.0
        t = np.power(10.0, u/(1.0*bins_per_decade))
    return np.array(rows, dtype=np.float64).reshape(-1, 4)


class CovarianceSetup(object):
    """Batch-invariant arguments of covariance.Covariance (covariance.py:47-191) for
    ``input_correlation_a is input_correlation_b``: the annular bins, the survey, the shot-noise
    bookkeeping of ``equal_windows`` / ``density`` / ``cosmic_shear`` (covariance.py:100-123, 194-205)."""

    def __init__(self, survey, theta_deg=(0.001, 1.0), bins_per_decade=5.0, survey_area_deg2=20.0, n_a=1.0e4,
                 n_b=1.0e4, variance=1.0, nongaussian_cov=True, power_spec="power_mm", poisson_noise_only=False):
        self.bins = annular_bins(theta_deg[0], theta_deg[1], bins_per_decade)
        p = _lib.CovParams()
        p.n_bins = self.bins.shape[0]
        p.which = _lib.POWER_SPEC.get("linear_power" if power_spec is None else power_spec, _lib.P_LINEAR)
        p.nongaussian, p.poisson_only = int(bool(nongaussian_cov)), int(bool(poisson_noise_only))
        q = survey.quadrature
        p.nq_osc, p.osc_phase = q.get("cov_osc", 4), q.get("cov_phase", 3.0)
        p.nq_ng = q.get("cov_ng", 0)
        lim = survey.limits
        ln_k = np.linspace(np.log(lim["k_min"]), np.log(lim["k_max"]), survey.precision["kernel_npoints"])
        p.zero_last_ka = int(np.exp(ln_k[-1]) > lim["k_max"])       # halo_trispectrum.py:100-107 in numpy arithmetic
        p.theta_min_rad = np.power(10.0, np.log10(theta_deg[0]*DEG_TO_RAD))     # covariance.py:93-97
        p.theta_max_rad = np.power(10.0, np.log10(theta_deg[1]*DEG_TO_RAD))
        self.area = survey_area_deg2*DEG_TO_RAD*DEG_TO_RAD
        p.area_sr = self.area

        def pair(n):
            try:
                return float(n[0]), float(n[1])
            except (TypeError, IndexError):
                return float(n), float(n)
        n_a1, n_a2 = pair(n_a)
        n_b1, n_b2 = pair(n_b)
        # a1 == b1 and a2 == b2 are the same objects; the two windows of one Kernel never compare
        # equal (kernel.py:248-259, 592-593: their cosmology copies differ)
        self.equal_windows = [False, False, False, False, True, True]
        self.density = [n_a1/self.area, n_a2/self.area, n_b1/self.area, n_b2/self.area, n_a1/self.area, n_a2/self.area]
        self.variance = variance
        for i in range(6):
            p.poisson[i] = variance*variance/self.density[i] if self.equal_windows[i] else 0.0
        conv = [w == _lib.WINDOW_CONVERGENCE for w in survey.window]
        shear = [conv[0], conv[1], conv[0], conv[1]]
        self.cosmic_shear = [bool(shear[0]*shear[1] or shear[2]*shear[3]), bool(shear[0]*shear[3] or shear[1]*shear[2])]
        p.shot_wt[0], p.shot_wt[1] = 1.0 + self.cosmic_shear[0], 1.0 + self.cosmic_shear[1]
        p.bessel_limit = bessel_limit(0, survey.precision["kernel_bessel_limit"])
        p.halofit_z = -1.0
        # log-spaced bin centres (always, for the bins built here): lets the non-Gaussian term share kernel values
        lc = np.log(self.bins[:, 2])
        p.bin_log0, p.bin_dlog = float(lc[0]), 0.0
        if lc.size >= 2:
            d = (lc[-1] - lc[0])/(lc.size - 1)
            if d > 0 and np.max(np.abs(lc - (lc[0] + d*np.arange(lc.size)))) < 1e-11:
                p.bin_dlog = float(d)
        self.params = p


def pack_params(dicts, keys):
    """[B, len(keys)] float64 array from a list of parameter dictionaries
    (KeyError on a missing key, like the reference)."""
    return np.array([[float(d[k]) for k in keys] for d in dicts], dtype=np.float64)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.ChompError("chomp_b200 needs a CUDA device: there is no CPU fallback")
    return torch


class Engine(object):
    """One C-ABI handle on one device."""

    def __init__(self, config=None, device=None):
        self.torch = _torch()
        self.lib = _lib.load()
        if device is None:
            device = self.torch.cuda.current_device()
        self.device = int(device)
        self._h = ctypes.c_void_p()
        _lib.check(self.lib.chomp_b200_create(ctypes.byref(self._h), self.device))
        self.cfg = None
        self.n_theta = 0
        if config is not None:
            self.configure(config)

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self.lib.chomp_b200_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- plumbing -----------------------------------------------------------
    def set_dndz_table(self, which, breaks, coef):
        """Upload a tabulated dN/dz (which = 0 / 1: window a / b, 2: both); configure() afterwards."""
        breaks = np.ascontiguousarray(breaks, dtype=np.float64)
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        if coef.shape != (breaks.size - 1, 4):
            raise ValueError("coef must be [len(breaks) - 1, 4]")
        _lib.check(self.lib.chomp_b200_set_dndz_table(self._h, int(which), int(breaks.size - 1),
                                                     breaks.ctypes.data_as(ctypes.c_void_p),
                                                     coef.ctypes.data_as(ctypes.c_void_p)))

    def configure(self, config):
        if isinstance(config, Survey):
            da, db = config.dist
            if da.kind == _lib.DNDZ_TABLE and db is da:
                self.set_dndz_table(2, da.breaks, da.coef)
            else:
                for i, d in enumerate((da, db)):
                    if d.kind == _lib.DNDZ_TABLE:
                        self.set_dndz_table(i, d.breaks, d.coef)
            config = config.config()
        _lib.check(self.lib.chomp_b200_configure(self._h, ctypes.byref(config)))
        self.cfg = config

    def reserve(self, n_points):
        _lib.check(self.lib.chomp_b200_reserve(self._h, int(n_points)))

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, a, cols=None):
        t = self.torch
        if not isinstance(a, t.Tensor):
            a = t.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
        a = a.to(device="cuda:%d" % self.device, dtype=t.float64).contiguous()
        if cols is not None and (a.dim() != 2 or a.shape[1] != cols):
            raise ValueError("expected a [B, %d] array, got %s" % (cols, tuple(a.shape)))
        return a

    def _new(self, *shape, dtype=None):
        t = self.torch
        return t.empty(shape, dtype=dtype or t.float64, device="cuda:%d" % self.device)

    @staticmethod
    def _p(tensor):
        return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None else None

    # -- stages ---------------------------------------------------------------
    def limber_tables(self, cosmo, status=None):
        cosmo = self._dev(cosmo, _lib.N_COSMO)
        _lib.check(self.lib.chomp_b200_limber_tables(self._h, cosmo.shape[0], self._p(cosmo),
                                                     self._p(status), self._stream()))
        return cosmo.shape[0]

    def mass_tables(self, cosmo, halo, z=None, status=None):
        cosmo, halo = self._dev(cosmo, _lib.N_COSMO), self._dev(halo, _lib.N_HALO)
        zt = None if z is None else self._dev(z).reshape(-1)
        _lib.check(self.lib.chomp_b200_mass_tables(self._h, cosmo.shape[0], self._p(cosmo), self._p(halo),
                                                   self._p(zt), self._p(status), self._stream()))
        return cosmo.shape[0]

    def halo_tables(self, halo, hod, status=None):
        halo, hod = self._dev(halo, _lib.N_HALO), self._dev(hod, _lib.N_HOD)
        _lib.check(self.lib.chomp_b200_halo_tables(self._h, halo.shape[0], self._p(halo), self._p(hod),
                                                   self._p(status), self._stream()))
        return halo.shape[0]

    def power(self, B, which, k):
        k = self._dev(k).reshape(-1)
        out = self._new(B, k.numel())
        _lib.check(self.lib.chomp_b200_power(self._h, B, int(which), k.numel(), self._p(k), self._p(out),
                                             self._stream()))
        return out

    def wtheta_stage(self, B, which, theta, status=None):
        theta = self._dev(theta).reshape(-1)
        out = self._new(B, theta.numel())
        _lib.check(self.lib.chomp_b200_wtheta(self._h, B, int(which), theta.numel(), self._p(theta),
                                              self._p(out), self._p(status), self._stream()))
        return out

    def wtheta(self, cosmo, halo, hod, theta, which, out=None, status=None):
        """All four stages on device-resident inputs -> w [B, n_theta]."""
        cosmo, halo = self._dev(cosmo, _lib.N_COSMO), self._dev(halo, _lib.N_HALO)
        hod = self._dev(hod, _lib.N_HOD)
        theta = self._dev(theta).reshape(-1)
        B = cosmo.shape[0]
        if out is None:
            out = self._new(B, theta.numel())
        _lib.check(self.lib.chomp_b200_wtheta_batch(
            self._h, B, self._p(cosmo), self._p(halo), self._p(hod), int(which), theta.numel(),
            self._p(theta), self._p(out), self._p(status), self._stream()))
        return out

    def wtheta_grouped(self, cosmo, halo, group_index, hod, theta, which, out=None, status=None):
        """The MCMC fast / slow split: `cosmo` [G, 10] and `halo` [G, 6] rows shared by the B points of
        `hod` [B, 5] through `group_index` [B] (row of each point).  The cosmology-level stages (Limber
        tables, mass tables) run G times, the HOD-level ones B times; w [B, n_theta] is bit-identical to
        `wtheta` on the expanded arrays."""
        t = self.torch
        cosmo, halo = self._dev(cosmo, _lib.N_COSMO), self._dev(halo, _lib.N_HALO)
        hod = self._dev(hod, _lib.N_HOD)
        theta = self._dev(theta).reshape(-1)
        if not isinstance(group_index, t.Tensor):
            group_index = t.as_tensor(np.ascontiguousarray(group_index, dtype=np.int32))
        group_index = group_index.to(device="cuda:%d" % self.device, dtype=t.int32).contiguous().reshape(-1)
        B, G = hod.shape[0], cosmo.shape[0]
        if halo.shape[0] != G or group_index.numel() != B:
            raise ValueError("cosmo / halo need one row per group, group_index one entry per point")
        if out is None:
            out = self._new(B, theta.numel())
        _lib.check(self.lib.chomp_b200_wtheta_batch_grouped(
            self._h, G, self._p(cosmo), self._p(halo), B, self._p(group_index), self._p(hod), int(which),
            theta.numel(), self._p(theta), self._p(out), self._p(status), self._stream()))
        return out

    def wtheta_host(self, cosmo, halo, hod, theta, which):
        """Host numpy in, host numpy out (copies inside): the end-to-end call."""
        cosmo = np.ascontiguousarray(cosmo, dtype=np.float64)
        halo = np.ascontiguousarray(halo, dtype=np.float64)
        hod = np.ascontiguousarray(hod, dtype=np.float64)
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1)
        B = cosmo.shape[0]
        if cosmo.shape != (B, _lib.N_COSMO) or halo.shape != (B, _lib.N_HALO) or hod.shape != (B, _lib.N_HOD):
            raise ValueError("parameter arrays must be [B,10], [B,6], [B,5]")
        w = np.empty((B, theta.size), dtype=np.float64)
        status = np.zeros(B, dtype=np.int32)
        _lib.check(self.lib.chomp_b200_wtheta_batch_host(
            self._h, B, cosmo.ctypes.data_as(ctypes.c_void_p), halo.ctypes.data_as(ctypes.c_void_p),
            hod.ctypes.data_as(ctypes.c_void_p), int(which), theta.size,
            theta.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p),
            status.ctypes.data_as(ctypes.c_void_p)))
        return w, status

    def mass_second_order(self, B, status=None):
        """bias_2_norm [B] of MassFunctionSecondOrder for the epochs of the last mass_tables call."""
        out = self._new(B)
        _lib.check(self.lib.chomp_b200_mass_second_order(self._h, int(B), self._p(out), self._p(status), self._stream()))
        return out

    def halofit(self, B, fit_z=-1.0, status=None):
        """HALOFIT parameters [B, 16] for the epochs of the last mass_tables call."""
        out = self._new(B, len(_lib.HALOFIT_FIELDS))
        _lib.check(self.lib.chomp_b200_halofit(self._h, int(B), float(fit_z), self._p(out), self._p(status),
                                               self._stream()))
        return out

    def cl(self, B, which, ell):
        ell = self._dev(ell).reshape(-1)
        out = self._new(B, ell.numel())
        _lib.check(self.lib.chomp_b200_cl(self._h, int(B), int(which), ell.numel(), self._p(ell), self._p(out),
                                          self._stream()))
        return out

    def trispectrum_1h(self, B):
        """[B, n_halo, n_halo] table of the 1-halo trispectrum (after mass_tables + halo_tables)."""
        out = self._new(B, self.cfg.n_halo, self.cfg.n_halo)
        _lib.check(self.lib.chomp_b200_trispectrum_1h(self._h, int(B), self._p(out), self._stream()))
        return out

    def trispectrum_eval(self, k1, k2, point=0):
        k1 = self._dev(np.atleast_1d(np.asarray(k1, dtype=np.float64))).reshape(-1)
        k2 = self._dev(np.atleast_1d(np.asarray(k2, dtype=np.float64))).reshape(-1)
        out = self._new(k1.numel())
        _lib.check(self.lib.chomp_b200_trispectrum_eval(self._h, int(point), k1.numel(), self._p(k1), self._p(k2),
                                                        self._p(out), self._stream()))
        return out

    def cov_kernel_ng(self, B, setup, status=None):
        """K_NG table [B, n_kernel, n_kernel], z_bar_NG [B], D(z_bar_NG) [B] after limber_tables."""
        _lib.check(self.lib.chomp_b200_cov_kernel_ng(self._h, int(B), ctypes.byref(setup.params), self._p(status),
                                                     self._stream()))
        nk = self.cfg.n_kernel
        return (self.table(_lib.T_KNG, B).reshape(B, nk, nk), self.table(_lib.T_ZBAR_NG, B).reshape(B),
                self.table(_lib.T_D_NG, B).reshape(B))

    def covariance(self, cosmo, halo, hod, setup, tri_z=None, status=None, parts=False):
        """Covariance.get_covariance for every point: [B, n_bins, n_bins] (and [B, 3, n, n] = P, G, NG)."""
        cosmo, halo = self._dev(cosmo, _lib.N_COSMO), self._dev(halo, _lib.N_HALO)
        hod = self._dev(hod, _lib.N_HOD)
        B, nb = cosmo.shape[0], setup.bins.shape[0]
        centre = self._dev(np.ascontiguousarray(setup.bins[:, 2]))
        delta = self._dev(np.ascontiguousarray(setup.bins[:, 3]))
        zt = None if tri_z is None else self._dev(np.broadcast_to(np.asarray(tri_z, dtype=np.float64), (B,)).copy())
        out = self._new(B, nb, nb)
        pt = self._new(B, 3, nb, nb) if parts else None
        _lib.check(self.lib.chomp_b200_covariance(
            self._h, B, ctypes.byref(setup.params), self._p(centre), self._p(delta), self._p(zt), self._p(cosmo),
            self._p(halo), self._p(hod), self._p(out), self._p(pt), self._p(status), self._stream()))
        return (out, pt) if parts else out

    def covariance_cross(self, other, cosmo, halo_a, hod_a, halo_b, hod_b, setup, tri_engine=None, halo_t=None, hod_t=None,
                         tri_z=None, status=None, parts=False):
        """Covariance between the correlation configured on this engine (a) and the one on `other` (b) for every
        point (covariance.py:60-63, matching_corrs False): [B, n_bins, n_bins] (and [B, 3, n, n] = P, G, NG).
        `tri_engine`: the engine configured for the trispectrum object (default: this one), `halo_t` / `hod_t` its
        own parameters (default: those of correlation a)."""
        cosmo = self._dev(cosmo, _lib.N_COSMO)
        halo_a, hod_a = self._dev(halo_a, _lib.N_HALO), self._dev(hod_a, _lib.N_HOD)
        halo_b, hod_b = self._dev(halo_b, _lib.N_HALO), self._dev(hod_b, _lib.N_HOD)
        halo_t = None if halo_t is None else self._dev(halo_t, _lib.N_HALO)
        hod_t = None if hod_t is None else self._dev(hod_t, _lib.N_HOD)
        B, nb = cosmo.shape[0], setup.bins.shape[0]
        centre = self._dev(np.ascontiguousarray(setup.bins[:, 2]))
        delta = self._dev(np.ascontiguousarray(setup.bins[:, 3]))
        zt = None if tri_z is None else self._dev(np.broadcast_to(np.asarray(tri_z, dtype=np.float64), (B,)).copy())
        out = self._new(B, nb, nb)
        pt = self._new(B, 3, nb, nb) if parts else None
        _lib.check(self.lib.chomp_b200_covariance_cross(
            self._h, other._h, None if tri_engine is None else tri_engine._h, B, ctypes.byref(setup.params), self._p(centre),
            self._p(delta), self._p(zt), self._p(cosmo), self._p(halo_a), self._p(hod_a), self._p(halo_b), self._p(hod_b),
            self._p(halo_t), self._p(hod_t), self._p(out), self._p(pt), self._p(status), self._stream()))
        return (out, pt) if parts else out

    def halo_ssc(self, B, k, what=1, status=None):
        """HaloSuperSampleCovariance for the last halo_tables batch: what = 0 I^1_2(k), 1 dln P / d delta_b; [B, n_k]."""
        k = self._dev(k).reshape(-1)
        out = self._new(B, k.numel())
        _lib.check(self.lib.chomp_b200_halo_ssc(self._h, int(B), int(what), k.numel(), self._p(k), self._p(out),
                                                self._p(status), self._stream()))
        return out

    def xi3d(self, B, which, r, status=None):
        """Correlation3d.raw_correlation for the last batch: xi(r) [B, n_r]."""
        r = self._dev(r).reshape(-1)
        out = self._new(B, r.numel())
        _lib.check(self.lib.chomp_b200_xi3d(self._h, int(B), int(which), r.numel(), self._p(r), self._p(out),
                                            self._p(status), self._stream()))
        return out

    def set_params(self, cosmo=None, halo=None, hod=None):
        arrs = [None if a is None else self._dev(a, n) for a, n in
                ((cosmo, _lib.N_COSMO), (halo, _lib.N_HALO), (hod, _lib.N_HOD))]
        B = next(a.shape[0] for a in arrs if a is not None)
        _lib.check(self.lib.chomp_b200_set_params(self._h, B, self._p(arrs[0]), self._p(arrs[1]),
                                                  self._p(arrs[2]), self._stream()))

    def set_zbar(self, z):
        z = self._dev(np.atleast_1d(np.asarray(z, dtype=np.float64))).reshape(-1)
        _lib.check(self.lib.chomp_b200_set_zbar(self._h, z.numel(), self._p(z), self._stream()))

    # -- inspection -------------------------------------------------------------
    def table(self, table_id, B):
        n = ctypes.c_int(0)
        _lib.check(self.lib.chomp_b200_copy_table(self._h, B, int(table_id), None, ctypes.byref(n), None))
        out = self._new(B, n.value)
        _lib.check(self.lib.chomp_b200_copy_table(self._h, B, int(table_id), self._p(out), ctypes.byref(n),
                                                  self._stream()))
        return out

    def evaluate(self, what, x, point=0, aux=0.0):
        x = self._dev(np.atleast_1d(np.asarray(x, dtype=np.float64))).reshape(-1)
        out = self._new(x.numel())
        _lib.check(self.lib.chomp_b200_eval(self._h, int(point), int(what), x.numel(), self._p(x),
                                            float(aux), self._p(out), self._stream()))
        return out

    def dfma_peak_tflops(self, iters=20000):
        v = ctypes.c_double(0.0)
        _lib.check(self.lib.chomp_b200_dfma_peak(self._h, int(iters), ctypes.byref(v)))
        return v.value

    def set_timing(self, on=True):
        _lib.check(self.lib.chomp_b200_set_timing(self._h, int(bool(on))))

    def kernel_times_ms(self):
        """Device time of each kernel of the last wtheta() call, by kernel name."""
        ms = (ctypes.c_double*len(_lib.KERNEL_NAMES))()
        _lib.check(self.lib.chomp_b200_get_timing(self._h, ms))
        return dict(zip(_lib.KERNEL_NAMES, [float(v) for v in ms]))

    def cov_kernel_times_ms(self):
        """Device time of each kernel class of the last covariance() call (timing on), by kernel name."""
        names = _lib.COV_KERNEL_NAMES + _lib.KERNEL_NAMES
        ms = (ctypes.c_double*len(names))()
        _lib.check(self.lib.chomp_b200_get_cov_timing(self._h, ms))
        return dict(zip(names, [float(v) for v in ms]))

    def launch_count(self):
        return int(self.lib.chomp_b200_launch_count(self._h))
