"""CPU oracle: a numpy/scipy restatement of CHOMP's halo-model -> Limber ->
Hankel hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module; chomp_b200/ must never do so.

Every function cites the reference file:line it restates.  The restatement is
pinned two ways (tests/test_oracle_*.py):

* against the reference's own known-answer vectors that still hold at HEAD
  (unit_test.py:131-144, 183-188, 267-303, 319-335, 346-407, 469-511), and
* against the reference itself, run in this container from the generated copy
  oracle/_ref (oracle/make_ref.py), with golden outputs committed under
  tests/golden/.

Third-party arithmetic (not under /root/reference): scipy.special.sici / j0 /
jn / erf / erfinv and FITPACK ``InterpolatedUnivariateSpline`` (k=3, i.e. the
not-a-knot interpolating cubic).  The installed scipy 1.18.1 provides them, as
it does for the generated reference copy.

Python-2 semantics the reference depends on are restated explicitly:
``(Omb2)**(3/4)`` has exponent 0 (cosmology.py:464) and ``1/b`` is an integer
division when ``b`` is an int (kernel.py:164-168).

An ``integ`` strategy (oracle.quadrature.Romberg or .Tight) is threaded through
every integral: ``Romberg`` reproduces the reference's numbers, ``Tight`` gives
the converged value of the same integrands on the same node grids / splines and
is the 1e-5 comparison target for the CUDA path.
"""
import copy as _copy

import numpy as np
from scipy import special
from scipy.interpolate import InterpolatedUnivariateSpline as _IUS
from scipy.optimize import brentq

from .quadrature import Romberg, Tight  # noqa: F401

# ----------------------------------------------------------------------------
# defaults.py:6-92
# ----------------------------------------------------------------------------
DEFAULT_COSMO = {
    "omega_m0": 0.278 - 4.15e-5/0.7**2, "omega_b0": 0.046, "omega_l0": 0.722,
    "omega_r0": 4.15e-5/0.7**2, "cmb_temp": 2.726, "h": 0.7, "sigma_8": 0.811,
    "n_scalar": 0.960, "w0": -1.0, "wa": 0.0}
DEFAULT_HALO = {"stq": 0.3, "st_little_a": 0.707, "c0": 9.0, "beta": -0.13,
                "alpha": -1, "delta_v": -1.0}
DEFAULT_HOD = {"log_M_min": 12.14, "sigma": 0.15, "log_M_0": 12.14,
               "log_M_1p": 13.43, "alpha": 1.0}
DEFAULT_LIMITS = {"k_min": 0.001, "k_max": 100.0, "mass_min": -1,
                  "mass_max": -1}
DEFAULT_PRECISION = {
    "corr_npoints": 50, "corr_precision": 1.48e-6, "cosmo_npoints": 50,
    "cosmo_precision": 1.48e-8, "dNdz_precision": 1.48e-8, "halo_npoints": 50,
    "halo_precision": 1.48e-5, "halo_limit": 100, "kernel_npoints": 50,
    "kernel_precision": 1.48e-6, "kernel_limit": 100,
    "kernel_bessel_limit": 8, "mass_npoints": 50, "mass_precision": 1.48e-8,
    "window_npoints": 100, "window_precision": 1.48e-6,
    "global_precision": 1.48e-32, "divmax": 20}


def precision(**overrides):
    p = dict(DEFAULT_PRECISION)
    p.update(overrides)
    return p


def _spline(x, y):
    return _IUS(np.asarray(x, dtype=float), np.asarray(y, dtype=float), k=3)


# ----------------------------------------------------------------------------
# cosmology.SingleEpoch  (cosmology.py:25-729)
# ----------------------------------------------------------------------------
class SingleEpoch(object):
    def __init__(self, z, cosmo=None, prec=None, integ=None, limits=None, with_bao=False):
        self.with_bao = bool(with_bao)                    # cosmology.py:86, 556-571
        self.prec = prec or DEFAULT_PRECISION
        self.integ = integ or Romberg(self.prec["divmax"])
        self.limits = limits or DEFAULT_LIMITS
        self.c = dict(cosmo or DEFAULT_COSMO)
        self.z = max(float(z), 0.0)                       # cosmology.py:40-42
        c = self.c
        if c["w0"] != -1.0 or c["wa"] != 0.0:
            raise NotImplementedError("dynamical dark energy is out of scope "
                                      "(SURVEY.md 8(f) rank 4)")
        self.om, self.ob, self.ol, self.orad = (
            c["omega_m0"], c["omega_b0"], c["omega_l0"], c["omega_r0"])
        self.h, self.ns, self.s8, self.tcmb = (
            c["h"], c["n_scalar"], c["sigma_8"], c["cmb_temp"])
        self.H0 = 100.0/(2.998*10**5)                     # cosmology.py:59
        eps = self.prec["cosmo_precision"]
        tot = self.om + self.ol + self.orad               # cosmology.py:65-79
        self.flat = (1.0 - eps) <= tot <= (1.0 + eps)
        self.open = tot <= 1.0 - eps
        self.closed = tot > 1.0 + eps
        self.k_min, self.k_max = self.limits["k_min"], self.limits["k_max"]
        self.delta_H = (1.94e-5*self.om**(-0.785 - 0.05*np.log(self.om)) *
                        np.exp(-0.95*(self.ns - 1) - 0.169*(self.ns - 1)**2))
        # cosmology.py:93-119
        self.chi = self.integ(self.inv_hubble, 0.0, self.z,
                              self.prec["cosmo_precision"])
        self.growth_norm = self.growth_approx(1.0)
        self.growth = self.growth_approx(1.0/(1.0 + self.z))/self.growth_norm
        self.sigma_norm = 1.0
        self.sigma_norm = self.s8*self.growth/self.sigma_r(8.0)

    # cosmology.py:164-182 (no curvature term, Q4)
    def E0(self, z):
        a = 1.0/(1.0 + z)
        return self.ol + self.om/(a*a*a) + self.orad/(a*a*a*a)

    def inv_hubble(self, z):
        return 1.0/(self.H0*np.sqrt(self.E0(z)))

    # cosmology.py:215-231 -- what growth_factor_eval returns at HEAD (:326)
    def growth_approx(self, a):
        om = self.om/a**3
        den = self.ol + om
        Om, Ol = om/den, self.ol/den
        return (5.0*Om/(2.0/a))/(Om*(4.0/7.0) - Ol +
                                 (1.0 + 0.5*Om)*(1.0 + Ol/70.0))

    def omega_m(self):                                    # cosmology.py:375
        return self.om*(1.0 + self.z)**3/self.E0(self.z)

    def omega_l(self):                                    # cosmology.py:384
        return self.ol/self.E0(self.z)

    def delta_c(self):                                    # cosmology.py:393
        d = 0.15*(12.0*np.pi)**(2.0/3.0)
        if self.open:
            d *= self.omega_m()**0.0185
        if self.flat and self.om < 1.0001:
            d *= self.omega_m()**0.0055
        return d

    def delta_v(self):                                    # cosmology.py:409
        d = 178.0
        if self.open:
            d /= self.omega_m()**0.7
        if self.flat and self.om < 1.0001:
            d /= self.omega_m()**0.55
        return d/self.growth

    def rho_crit(self):                                   # cosmology.py:425
        return 1.879/1.989*3.086**3*1e10*self.E0(self.z)

    def rho_bar(self):                                    # cosmology.py:440
        return self.rho_crit()*self.omega_m()

    # cosmology.py:449-472 with the Python-2 exponent (3/4 == 0), Q1 and Q5
    def transfer(self, k):                                # cosmology.py:556-571
        if self.with_bao:
            return self.transfer_bao(k)
        theta = self.tcmb/2.7
        omh2 = self.om*self.h**2
        ombh2 = self.ob*self.h**2
        fb = self.ob/self.om
        s = 44.5*np.log(9.83/omh2)/np.sqrt(1 + 10.0*ombh2**0)
        alpha = (1 - 0.328*np.log(431.0*omh2)*fb +
                 0.38*np.log(22.3*omh2)*fb**2)
        gamma = self.om*self.h*(alpha + (1 - alpha)/(1 + 0.43*k*s)**4)
        q = k*theta/gamma
        L0 = np.log(2*np.e + 1.8*q)
        C0 = 14.2 + 731.0/(1 + 62.5*q)
        return L0/(L0 + C0*q*q)

    def transfer_bao(self, k):                            # cosmology.py:474-538 (Eisenstein & Hu 1998 with wiggles)
        k = np.asarray(k, dtype=float)
        theta = self.tcmb/2.7
        Ob, Om = self.ob, self.om
        Oc, h = Om - Ob, self.h
        Oh2, Obh2, ObO = Om*h**2, Ob*h**2, Ob/Om
        zeq = 2.5e4*Oh2*theta**(-4)
        keq = 7.46e-2*Oh2*theta**(-2)
        b1 = 0.313*Oh2**(-0.419)*(1. + 0.607*Oh2**0.674)
        b2 = 0.238*Oh2**0.223
        zd = 1291.*(Oh2**0.251/(1. + 0.659*Oh2**0.828))*(1. + b1*Obh2**b2)

        def R(z):
            return 31.5*Obh2*theta**(-4)*(1000./z)
        Req, Rd = R(zeq), R(zd)
        s = (2./(3.*keq))*np.sqrt(6./Req)*np.log((np.sqrt(1. + Rd) + np.sqrt(Rd + Req))/(1. + np.sqrt(Req)))
        ks = k*h*s
        kSilk = 1.6*Obh2**0.52*Oh2**0.73*(1. + (10.4*Oh2)**(-0.95))
        q = k*h/(13.41*keq)

        def G(y):
            return y*(-6.*np.sqrt(1. + y) + (2 + 3*y)*np.log((np.sqrt(1. + y) + 1.)/(np.sqrt(1. + y) - 1.)))
        alpha_b = 2.07*keq*s*(1. + Rd)**(-3./4.)*G((1. + zeq)/(1. + zd))
        beta_b = 0.5 + ObO + (3. - 2.*ObO)*np.sqrt((17.2*Oh2)**2 + 1.)

        def T0t(a, b):                                    # :513-516 (both logarithms at the same k)
            C = (14.2/a) + 386./(1. + 69.9*q**1.08)
            L = np.log(np.e + 1.8*b*q)
            return L/(L + C*q**2)
        a1 = (46.9*Oh2)**0.670*(1. + (32.1*Oh2)**(-0.532))
        a2 = (12.*Oh2)**0.424*(1. + (45.*Oh2)**(-0.582))
        alpha_c = a1**(-ObO)*a2**(-ObO**3)
        b1 = 0.944*(1. + (458.*Oh2)**(-0.708))**(-1)
        b2 = (0.395*Oh2)**(-0.0266)
        beta_c = 1./(1. + b1*((Oc/Om)**b2 - 1))
        f = 1./(1. + (ks/5.4)**4)
        Tc = f*T0t(1, beta_c) + (1. - f)*T0t(alpha_c, beta_c)
        beta_node = 8.41*(Oh2**0.435)
        stilde = s/(1. + (beta_node/ks)**3)**(1./3.)
        Tb1 = T0t(1., 1.)/(1. + (ks/5.2)**2)
        Tb2 = (alpha_b/(1. + (beta_b/ks)**3))*np.exp(-(k*h/kSilk)**1.4)
        Tb = np.sinc(k*stilde/np.pi)*(Tb1 + Tb2)          # sin(k s~) / (k s~): k without the h that ks carries (:536)
        return ObO*Tb + (Oc/Om)*Tc

    def delta_k(self, k):                                 # cosmology.py:574
        d = self.delta_H**2*(k/self.H0)**(3 + self.ns)*self.transfer(k)**2/self.h
        return d*(self.growth*self.growth*self.sigma_norm*self.sigma_norm)

    def linear_power(self, k):                            # cosmology.py:589
        k = np.asarray(k, dtype=float)
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(k > 1e-16,
                            2.0*np.pi*np.pi*self.delta_k(k)/(k*k*k), 1e-16)

    def sigma_limits(self, R):                            # cosmology.py:611-629
        k_lo, k_hi = self.k_min, self.k_max
        need_lo, need_hi = 1.0/R/10.0, 1.0/R*14.0662
        if need_lo <= k_lo:
            k_lo = need_lo if need_lo > self.k_min/100.0 else self.k_min/100.0
        if need_hi >= k_hi:
            k_hi = need_hi if need_hi < self.k_max*100.0 else self.k_max*100.0
        return k_lo, k_hi

    def _sigma_integrand(self, ln_k, R):                  # cosmology.py:640
        k = np.exp(ln_k)
        x = R*k
        W = 3.0*(np.sin(x)/x**3 - np.cos(x)/x**2)
        return k*self.linear_power(k)*W*W*k*k

    def sigma_r(self, R):                                 # cosmology.py:602
        k_lo, k_hi = self.sigma_limits(R)
        breaks = ()
        if self.integ.name == "tight":
            # resolve the top-hat window's oscillation: a panel per ~6 rad of
            # phase (2kR) plus decade marks below kR = 1
            x_hi = k_hi*R
            marks = [np.arange(1.0, x_hi, 3.0)/R, np.geomspace(k_lo, 1.0/R, 8)]
            breaks = np.log(np.concatenate(marks))
        s2 = self.integ(self._sigma_integrand, np.log(k_lo), np.log(k_hi),
                        self.prec["cosmo_precision"], breaks=breaks, args=(R,))
        return np.sqrt(s2/(2.0*np.pi*np.pi))

    def mass_to_scale(self, M):                           # cosmology.py:662
        return (3.0*M/(4.0*np.pi*self.rho_bar()))**(1.0/3.0)

    def sigma_m(self, M):
        return self.sigma_r(self.mass_to_scale(M))

    def nu_m(self, M):                                    # cosmology.py:687
        s = self.delta_c()/self.sigma_m(M)
        return s*s


# ----------------------------------------------------------------------------
# cosmology.MultiEpoch  (cosmology.py:731-1164)
# ----------------------------------------------------------------------------
class MultiEpoch(object):
    def __init__(self, z_min, z_max, cosmo=None, prec=None, integ=None,
                 limits=None, epoch0=None):
        self.prec = prec or DEFAULT_PRECISION
        self.integ = integ or Romberg(self.prec["divmax"])
        self.epoch0 = epoch0 or SingleEpoch(0.0, cosmo, self.prec, self.integ,
                                            limits)
        self.H0 = self.epoch0.H0
        self.om = self.epoch0.om
        self.set_redshift(z_min, z_max)

    def regrid(self, z_min, z_max):
        """What WindowFunction.set_cosmology_object does (kernel.py:289-306):
        a shallow copy sharing epoch0, re-tabulated on a new z range."""
        other = _copy.copy(self)
        other.set_redshift(z_min, z_max)
        return other

    def set_redshift(self, z_min, z_max):                 # cosmology.py:819-843
        self.z_min = max(z_min, 0.0)
        self.z_max = z_max
        n = self.prec["cosmo_npoints"]
        self.z_nodes = np.linspace(self.z_min, self.z_max, n)
        # cosmology.py:787-817
        self.chi_nodes = np.array([
            self.integ(self.epoch0.inv_hubble, 0.0, z,
                       self.prec["cosmo_precision"]) for z in self.z_nodes])
        self.growth_nodes = (self.epoch0.growth_approx(1.0/(1.0 + self.z_nodes))
                             / self.epoch0.growth_norm)
        self._chi_of_z = _spline(self.z_nodes, self.chi_nodes)
        self._z_of_chi = _spline(self.chi_nodes, self.z_nodes)
        self._growth_of_z = _spline(self.z_nodes, self.growth_nodes)

    def comoving_distance(self, z):                       # cosmology.py:873
        z = np.asarray(z, dtype=float)
        return np.where((z <= self.z_max) & (z >= self.z_min),
                        self._chi_of_z(z), 0.0)

    def redshift(self, chi):                              # cosmology.py:922
        return self._z_of_chi(chi)

    def growth_factor(self, z):                           # cosmology.py:934
        z = np.asarray(z, dtype=float)
        return np.where((z <= self.z_max) & (z >= self.z_min),
                        self._growth_of_z(z), 1.0)

    def inv_hubble(self, z):                              # cosmology.py:862
        return self.epoch0.inv_hubble(z)


# ----------------------------------------------------------------------------
# mass_function.MassFunction  (mass_function.py:25-363)
# ----------------------------------------------------------------------------
class MassFunction(object):
    def __init__(self, epoch, halo=None, prec=None, integ=None, limits=None):
        self.prec = prec or epoch.prec
        self.integ = integ or epoch.integ
        self.limits = limits or epoch.limits
        self.epoch = epoch
        self.set_halo_params(halo or DEFAULT_HALO)
        self.delta_c = epoch.delta_c()
        self._find_mass_limits()
        self._tabulate()
        self.normalize()

    def set_halo_params(self, halo):                      # mass_function.py:141
        self.halo = dict(halo)
        self.stq, self.sta = halo["stq"], halo["st_little_a"]

    def _find_mass_limits(self):                          # mass_function.py:160
        n = self.prec["mass_npoints"]
        lo, hi = 1.0e9, 1.0e16
        if self.limits["mass_min"] > 0 and self.limits["mass_max"] > 0:
            lo, hi = self.limits["mass_min"], self.limits["mass_max"]
        else:
            nu = self.epoch.nu_m
            self.walk_steps = 0
            while True:
                self.walk_steps += 1
                v = nu(lo)
                if 0.1*1.05 < v:
                    lo = lo/1.05
                    continue
                if 0.1*0.95 > v:
                    lo = lo*1.05
                    continue
                v = nu(hi)
                if 50.0*0.95 > v:
                    hi = hi*1.05
                    continue
                if 50.0*1.05 < v:
                    hi = hi/1.05
                    continue
                break
        self.ln_mass_min, self.ln_mass_max = np.log(lo), np.log(hi)
        self.ln_mass_nodes = np.linspace(self.ln_mass_min, self.ln_mass_max, n)

    def _tabulate(self):                                  # mass_function.py:205
        self.nu_nodes = np.array([self.epoch.nu_m(np.exp(lm))
                                  for lm in self.ln_mass_nodes])
        self.nu_min = 1.001*self.nu_nodes[0]
        self.nu_max = 0.999*self.nu_nodes[-1]
        self._nu_of_lnm = _spline(self.ln_mass_nodes, self.nu_nodes)
        self._lnm_of_nu = _spline(self.nu_nodes, self.ln_mass_nodes)
        self.m_star = self.mass(1.0)

    def normalize(self):                                  # mass_function.py:225
        self.f_norm = 1.0
        self.bias_norm = 1.0
        breaks = np.geomspace(self.nu_min, self.nu_max, 12)
        rt = self.prec["mass_precision"]
        self.f_norm = 1.0/self.integ(self.f_nu, self.nu_min, self.nu_max, rt,
                                     breaks=breaks)
        self.bias_norm = 1.0/self.integ(
            lambda v: self.f_nu(v)*self.bias_nu(v), self.nu_min, self.nu_max,
            rt, breaks=breaks)

    def f_nu(self, nu):                                   # mass_function.py:243
        nup = nu*self.sta
        return (self.f_norm*(1.0 + nup**(-1.0*self.stq))*np.sqrt(nup) *
                np.exp(-0.5*nup)/nu)

    def bias_nu(self, nu):                                # mass_function.py:290
        nup = nu*self.sta
        return self.bias_norm*(1.0 + (nup - 1.0)/self.delta_c +
                               2.0*self.stq/(self.delta_c*(1.0 + nup**self.stq)))

    def nu(self, M):                                      # mass_function.py:315
        return self._nu_of_lnm(np.log(M))

    def ln_mass(self, nu):                                # mass_function.py:326
        return self._lnm_of_nu(nu)

    def mass(self, nu):                                   # mass_function.py:337
        return np.exp(self._lnm_of_nu(nu))


class TinkerMassFunction(MassFunction):                   # mass_function.py:436-564 (Tinker et al. 2010)
    """f(nu) and b(nu) of Tinker et al. (2010); nu = (delta_c / sigma)^2 is the square of their variable.  The
    five shape parameters are cubic splines in ln(Delta_v) through the nine tabulated over-densities, scaled with
    redshift; only the bias is normalised (mass_function.py:528-542)."""
    DELTA = (200, 300, 400, 600, 800, 1200, 1600, 2400, 3200)
    TABLE = {"alpha": (0.368, 0.363, 0.385, 0.389, 0.393, 0.365, 0.379, 0.355, 0.327),
             "beta": (0.589, 0.585, 0.544, 0.543, 0.564, 0.632, 0.637, 0.673, 0.702),
             "gamma": (0.864, 0.922, 0.987, 1.09, 1.20, 1.34, 1.50, 1.68, 1.81),
             "phi": (-0.729, -0.789, -0.910, -1.05, -1.20, -1.26, -1.45, -1.50, -1.49),
             "eta": (-0.243, -0.261, -0.261, -0.273, -0.278, -0.301, -0.301, -0.319, -0.336)}
    Z_POWER = {"alpha": 0.0, "beta": 0.20, "phi": -0.08, "eta": 0.27, "gamma": -0.01}     # :544-564

    def __init__(self, epoch, halo=None, prec=None, integ=None, limits=None):
        self.prec = prec or epoch.prec
        self.integ = integ or epoch.integ
        self.limits = limits or epoch.limits
        self.epoch = epoch
        self.set_halo_params(halo or DEFAULT_HALO)
        self.delta_c = epoch.delta_c()
        self.delta_v = self.halo["delta_v"]
        if self.delta_v == -1:
            self.delta_v = epoch.delta_v()
        ld = np.log(self.DELTA)
        self.par = {k: float(_spline(ld, np.array(v))(np.log(self.delta_v)))*(1.0 + epoch.z)**self.Z_POWER[k]
                    for k, v in self.TABLE.items()}
        self._find_mass_limits()
        self._tabulate()
        self.normalize()

    def f_nu(self, nu):                                   # mass_function.py:493-508
        p = self.par
        s = np.sqrt(nu)
        return p["alpha"]*(1 + np.power(p["beta"]*s, -2*p["phi"]))*np.power(nu, p["eta"])*np.exp(-p["gamma"]*nu/2.0)/s

    def bias_nu(self, nu):                                # mass_function.py:510-526
        s = np.sqrt(nu)
        y = np.log10(self.delta_v)
        A = 1 + 0.24*y*np.exp(-(4.0/y)**4)
        a = 0.44*y - 0.88
        B, b = 0.183, 1.5
        C = 0.019 + 0.107*y + 0.19*np.exp(-(4.0/y)**4)
        c = 2.4
        return self.bias_norm*(1 - A*s**a/(s**a + self.delta_c**a) + B*s**b + C*s**c)

    def normalize(self):                                  # mass_function.py:528-542: the bias only
        self.f_norm = 1.0
        self.bias_norm = 1.0
        breaks = np.geomspace(self.nu_min, self.nu_max, 12)
        self.bias_norm = 1.0/self.integ(lambda v: self.f_nu(v)*self.bias_nu(v), self.nu_min, self.nu_max,
                                        self.prec["mass_precision"], breaks=breaks)


class MassFunctionSecondOrder(MassFunction):               # mass_function.py:365-433
    def _tabulate(self):                                  # mass_function.py:371-393
        MassFunction._tabulate(self)
        # sigma_m(M) at the nodes; nu = (delta_c / sigma)^2 there
        self.sigma_nodes = self.delta_c/np.sqrt(self.nu_nodes)
        self._sigma_of_nu = _spline(self.nu_nodes, self.sigma_nodes)

    def normalize(self):                                  # mass_function.py:395-421
        MassFunction.normalize(self)
        self.bias_2_norm = 0.0
        breaks = np.geomspace(self.nu_min, self.nu_max, 12)
        self.bias_2_norm = -self.integ(lambda v: self.f_nu(v)*self.bias_2_nu(v), self.nu_min, self.nu_max,
                                       self.prec["mass_precision"], breaks=breaks)

    def bias_2_nu(self, nu):                              # mass_function.py:423-430
        sigma = self._sigma_of_nu(nu)
        nup = nu*self.sta
        return self.bias_2_norm + (
            8.0/21.0*(self.bias_nu(nu) - 1.0) + (nu - 3.0)/(sigma*sigma) +
            2.0*self.stq/(self.delta_c**2*(1.0 + nup**self.stq))*(2.0*self.stq + 2*nup - 1.0))


# ----------------------------------------------------------------------------
# hod.HODZheng / hod.HODMandelbaum  (hod.py:141-299)
# ----------------------------------------------------------------------------
class HODZheng(object):
    kind = 0

    def __init__(self, hod=None, halo_precision=DEFAULT_PRECISION["halo_precision"]):
        p = dict(hod or DEFAULT_HOD)
        self.p = p
        self.log_M_min, self.sigma = p["log_M_min"], p["sigma"]
        self.log_M_0, self.log_M_1p, self.alpha = (
            p["log_M_0"], p["log_M_1p"], p["alpha"])
        # hod.py:176-185 (Q9: the 'secon_moment_zero' typo means no clamp; Q10)
        self.first_moment_zero = 10.0**(
            self.log_M_min + self.sigma*special.erfinv(2.0*halo_precision - 1.0))
        self.second_moment_zero = 10.0**self.log_M_0
        self.safe_norm = 10.0**(self.log_M_min + 1.0*self.sigma)

    def central(self, M):                                 # hod.py:196
        lm = np.log10(M)
        if self.sigma <= 0.0:
            return np.where(lm > self.log_M_min, 1.0, 0.0)
        return 0.5*(1 + special.erf((lm - self.log_M_min)/self.sigma))

    def satellite(self, M):                               # hod.py:214
        d = np.asarray(M - 10.0**self.log_M_0, dtype=float)
        with np.errstate(invalid="ignore"):
            return np.where(d > 0.0,
                            self.central(M)*np.power(
                                np.where(d > 0.0, d, 1.0)/10**self.log_M_1p,
                                self.alpha), 0.0)

    def first_moment(self, M):                            # hod.py:188
        return self.central(M) + self.satellite(M)

    def second_moment(self, M):                           # hod.py:192 (Q8)
        ns = self.satellite(M)
        return (2 + ns)*ns

    def nth_moment(self, M, n=3):                         # hod.py:68-92
        if n == 1:
            return self.first_moment(M)
        if n == 2:
            return self.second_moment(M)
        m1 = self.first_moment(M)
        with np.errstate(divide="ignore", invalid="ignore"):
            a2 = np.where(m1 != 0.0, self.second_moment(M)/m1**2, 0.0)
        out = m1**n
        for j in range(n):
            out = out*(j*a2 - j + 1)
        return out

    def kink_masses(self):
        return [10.0**self.log_M_0]


class HODMandelbaum(HODZheng):
    kind = 1

    def __init__(self, hod, halo_precision=None):
        # hod.py:250-261: with hod_dict None the base-class attributes are never
        # created and Halo cannot use the object; a dict is therefore required.
        p = dict(hod)
        self.p = p
        self.log_M_0 = p["log_M_0"]
        self.log_M_min = np.log10(3.0) + p["log_M_0"]
        self.w = p["w"]
        self.first_moment_zero = -1
        self.second_moment_zero = -1
        self.safe_norm = -1

    def central(self, M):                                 # hod.py:270
        return np.where(np.log10(M) >= self.log_M_0, 1.0, 0.0)

    def satellite(self, M):                               # hod.py:286
        r = M/10**self.log_M_min
        return np.where(np.log10(M) < self.log_M_min, r**2*self.w, r*self.w)

    def kink_masses(self):
        return [10.0**self.log_M_0, 10.0**self.log_M_min]


# ----------------------------------------------------------------------------
# halo.Halo  (halo.py:23-1086)
# ----------------------------------------------------------------------------
class Halo(object):
    """State-free construction: everything is built at ``epoch.z``."""

    exclusion = False

    def __init__(self, epoch, mass, hod, halo=None, prec=None, integ=None,
                 extrapolate=False, profile_halo=None):
        self.prec = prec or epoch.prec
        self.integ = integ or epoch.integ
        self.epoch, self.mass, self.hod = epoch, mass, hod
        halo = dict(halo or DEFAULT_HALO)
        # profile_halo lets tests reproduce Halo.set_halo (halo.py:196-212),
        # which refreshes the mass function but keeps stale c(M)/r_v(M) splines
        prof = dict(profile_halo or halo)
        self.k_min, self.k_max = epoch.limits["k_min"], epoch.limits["k_max"]
        self.ln_k_nodes = np.linspace(np.log(self.k_min), np.log(self.k_max),
                                      self.prec["halo_npoints"])
        self.c0 = prof["c0"]/(1.0 + epoch.z)              # halo.py:65
        self.beta = prof["beta"]
        if prof["alpha"] != -1:
            raise NotImplementedError("non-NFW profiles are out of scope")
        self.delta_v = prof["delta_v"]
        if self.delta_v == -1:
            self.delta_v = epoch.delta_v()                # halo.py:73-75
        self.rho_bar = epoch.rho_bar()
        self.extrapolate = extrapolate
        self._profile_splines()
        self._n_bar()
        self.tab = {}

    # halo.py:839-902: ln c and ln r_v are splined on the mass nodes (Q7)
    def _profile_splines(self):
        M = np.exp(self.mass.ln_mass_nodes)
        self._ln_c = _spline(self.mass.ln_mass_nodes,
                             np.log(self.c0*(M/self.mass.m_star)**self.beta))
        r3 = 3.0*M/(4.0*np.pi*self.delta_v*self.rho_bar)
        self._ln_rv = _spline(self.mass.ln_mass_nodes, np.log(r3**(1.0/3.0)))

    def concentration(self, M):                           # halo.py:451
        return np.exp(self._ln_c(np.log(M)))

    def virial_radius(self, M):                           # halo.py:441
        return np.exp(self._ln_rv(np.log(M)))

    def y(self, ln_k, M):                                 # halo.py:561-585
        k = np.exp(ln_k)
        c = self.concentration(M)
        cp = 1.0 + c
        z = k*self.virial_radius(M)/c
        si_z, ci_z = special.sici(z)
        si_cz, ci_cz = special.sici(cp*z)
        rho = (np.cos(z)*(ci_cz - ci_z) + np.sin(z)*(si_cz - si_z) -
               np.sin(c*z)/(cp*z))
        return rho/(np.log(cp) - c/cp)

    def mass_window(self, M, ln_k):                       # halo.py:1223-1233
        kR = np.exp(ln_k)*2*self.virial_radius(M)
        return ((kR*np.cos(kR) + kR*kR*kR*special.sici(kR)[1] +
                 (2 - kR*kR)*np.sin(kR))/(3.0*kR))

    # -- where the integrands stop being smooth (Tight strategy only) --------
    def _mass_to_lnnu(self, M):
        """ln nu at which mass(nu) == M on the *inverse* spline (the integrands
        are functions of nu through MassFunction.mass, mass_function.py:337)."""
        lo, hi = self.mass.nu_min, self.mass.nu_max
        g = lambda v: float(self.mass.ln_mass(v)) - np.log(M)
        if g(lo) >= 0.0 or g(hi) <= 0.0:
            return None
        return np.log(brentq(g, lo, hi, xtol=1e-15, rtol=1e-15))

    def _crossings(self, moment, lo, hi):
        """ln nu where moment(mass(nu)) crosses 1 (halo.py:1038-1041,1084-1086)."""
        grid = np.linspace(lo, hi, 4001)
        g = lambda x: float(moment(self.mass.mass(np.exp(x)))) - 1.0
        vals = moment(self.mass.mass(np.exp(grid))) - 1.0
        out = []
        for i in np.nonzero(np.sign(vals[1:]) != np.sign(vals[:-1]))[0]:
            if vals[i] == 0.0:
                out.append(grid[i])
                continue
            # a jump (step-function HOD) converges to the jump location
            out.append(brentq(g, grid[i], grid[i + 1], xtol=1e-15, rtol=1e-15))
        return out

    def _panel_hints(self, lo, hi, moment=None):
        if self.integ.name != "tight":
            return (), ()
        breaks = list(np.log(self.mass.nu_nodes))
        singular = []
        for M in self.hod.kink_masses():
            x = self._mass_to_lnnu(M)
            if x is not None:
                singular.append(x)
        if moment is not None:
            breaks += self._crossings(moment, lo, hi)
        return breaks, singular

    def _lower_limit(self, zero_mass):                    # halo.py:675-679 etc.
        if zero_mass > -1 and zero_mass > np.exp(self.mass.ln_mass_min):
            return float(self.mass.nu(zero_mass))
        return self.mass.nu_min

    def _safe_ln_nu(self):
        return np.log(self.mass.nu(self.hod.safe_norm))

    def _n_bar(self):                                     # halo.py:674-707
        m = self.mass
        lo = np.log(self._lower_limit(self.hod.first_moment_zero))
        hi = np.log(m.nu_max)
        f = lambda x, norm=1.0: self._nbar_integrand(x)*norm
        norm = 1.0
        sn = self.hod.safe_norm
        if sn != -1 and np.exp(m.ln_mass_min) < sn < np.exp(m.ln_mass_max):
            inv = f(self._safe_ln_nu())
            norm = 1.0/inv if inv > 1e-16 else 1.0
        br, sg = self._panel_hints(lo, hi)
        self.n_bar_over_rho_bar = self.integ(
            f, lo, hi, self.prec["halo_precision"], breaks=br, singular=sg,
            args=(norm,))/norm
        self.n_bar = self.n_bar_over_rho_bar*self.rho_bar

    def _nbar_integrand(self, ln_nu):                     # halo.py:704
        nu = np.exp(ln_nu)
        M = self.mass.mass(nu)
        return nu*self.hod.first_moment(M)*self.mass.f_nu(nu)/M

    # the five k-dependent mass integrals, halo.py:904-1086
    def _integrand(self, name, ln_nu, ln_k):
        nu = np.exp(ln_nu)
        m = self.mass
        M = m.mass(nu)
        y = self.y(ln_k, M)
        f = nu*m.f_nu(nu)
        if name == "h_m":                                 # halo.py:923
            v = f*m.bias_nu(nu)*y
            return v*self.mass_window(M, ln_k) if self.exclusion else v
        if name == "pp_mm":                               # halo.py:990
            return f*M*y*y
        if name == "h_g":                                 # halo.py:964
            v = f*m.bias_nu(nu)*y*self.hod.first_moment(M)/M
            return v*self.mass_window(M, ln_k) if self.exclusion else v
        if name == "pp_gg":                               # halo.py:1032
            n2 = self.hod.second_moment(M)
            return np.where(n2 < 1, f*n2*y/M, f*n2*y*y/M)
        if name == "pp_gm":                               # halo.py:1078
            n1 = self.hod.first_moment(M)
            return np.where(n1 < 1, f*n1*y, f*n1*y*y)
        raise KeyError(name)

    def table(self, name):
        if name in self.tab:
            return self.tab[name]
        m = self.mass
        hi = np.log(m.nu_max)
        moment = None
        if name in ("h_m", "pp_mm"):
            lo = np.log(m.nu_min)
        elif name in ("h_g", "pp_gm"):
            lo = np.log(self._lower_limit(self.hod.first_moment_zero))
            moment = self.hod.first_moment if name == "pp_gm" else None
        else:
            lo = np.log(self._lower_limit(self.hod.second_moment_zero))
            moment = self.hod.second_moment
        br, sg = self._panel_hints(lo, hi, moment)
        if name in ("h_m", "pp_mm"):
            sg = ()
        vals = np.empty_like(self.ln_k_nodes)
        rt = self.prec["halo_precision"]
        for i, ln_k in enumerate(self.ln_k_nodes):
            f = lambda x, lk=ln_k, norm=1.0: self._integrand(name, x, lk)*norm
            # the reference conditions Romberg's relative test by scaling the
            # integrand to ~1 at a reference point (halo.py:908, 943-949, ...)
            if name in ("h_m", "pp_mm"):
                norm = 1.0/float(f(0.0))
            else:
                norm = 1.0
                if self.hod.safe_norm != -1:
                    inv = float(f(self._safe_ln_nu()))
                    norm = 1.0/inv if inv > 1e-16 else 1.0
            vals[i] = self.integ(f, lo, hi, rt, breaks=br, singular=sg,
                                 args=(ln_k, norm))/norm
        if name == "pp_mm":
            vals /= self.rho_bar
        elif name == "h_g":
            vals /= self.n_bar_over_rho_bar
        elif name == "pp_gg":
            vals *= self.rho_bar/(self.n_bar*self.n_bar)
        elif name == "pp_gm":
            vals /= self.n_bar
        self.tab[name] = (vals, _spline(self.ln_k_nodes, vals))
        return self.tab[name]

    def _tab(self, name, k):                              # halo.py:649-672
        k = np.asarray(k, dtype=float)
        return np.where((k >= self.k_min) & (k <= self.k_max),
                        self.table(name)[1](np.log(k)), 0.0)

    def linear_power(self, k):
        return self.epoch.linear_power(k)

    def two_halo_power(self, k):
        return self.epoch.linear_power(k)

    # halo.py:277-439
    def _assemble(self, k, a, b, pp, power_law):
        k = np.asarray(k, dtype=float)
        P2 = self.two_halo_power
        kmin, kmax = self.k_min, self.k_max
        inside = P2(k)*self._tab(a, k)*self._tab(b, k) + self._tab(pp, k)
        low = P2(k)*(self._tab(a, kmin)*self._tab(b, kmin) +
                     self._tab(pp, kmin)/P2(kmin))
        if not self.extrapolate:
            return np.where(k < kmin, low, np.where(k <= kmax, inside, 0.0))
        at_max = (P2(kmax)*self._tab(a, kmax)*self._tab(b, kmax) +
                  self._tab(pp, kmax))
        if power_law:                                     # halo.py:343-351
            kk = np.exp(self.ln_k_nodes[-7:-1])
            lv = np.log(P2(kk)*self._tab(a, kk)*self._tab(b, kk) +
                        self._tab(pp, kk))
            slope = np.mean((lv[1:] - lv[:-1]) /
                            (self.ln_k_nodes[-6:-1] - self.ln_k_nodes[-7:-2]))
            high = np.power(k/kmax, slope)*at_max
        else:
            high = P2(k)*at_max/P2(kmax)
        return np.where(k < kmin, low, np.where(k < kmax, inside, high))

    def power_mm(self, k):
        return self._assemble(k, "h_m", "h_m", "pp_mm", False)

    def power_gm(self, k):
        return self._assemble(k, "h_g", "h_m", "pp_gm", True)

    def power_gg(self, k):
        return self._assemble(k, "h_g", "h_g", "pp_gg", True)

    def power(self, which, k):
        return getattr(self, which)(k)


class HaloSuperSampleCovariance(Halo):                    # halo.py:1089-1199 (Takada & Hu 2013)
    """Adds I^1_2(k) = rho_bar^-1 int dln nu nu f(nu) b(nu) y(k, M)^2 M and the response of the matter power
    spectrum to a super-survey over-density delta_b."""

    def __init__(self, epoch, mass, hod, halo=None, delta_b=0.0, **kw):
        Halo.__init__(self, epoch, mass, hod, halo, **kw)
        self.delta_b = delta_b
        self._i12 = None

    def _i_1_2_integrand(self, ln_nu, ln_k, norm=1.0):    # halo.py:1193-1199
        nu = np.exp(ln_nu)
        M = self.mass.mass(nu)
        y = self.y(ln_k, M)
        return nu*self.mass.f_nu(nu)*self.mass.bias_nu(nu)*y*y*M*norm

    def i_1_2_table(self):                                # halo.py:1174-1191
        if self._i12 is None:
            m = self.mass
            lo, hi = np.log(m.nu_min), np.log(m.nu_max)
            br, _ = self._panel_hints(lo, hi)
            vals = np.empty_like(self.ln_k_nodes)
            for i, ln_k in enumerate(self.ln_k_nodes):
                norm = 1.0/float(self._i_1_2_integrand(0.0, ln_k, 1.0))
                vals[i] = self.integ(self._i_1_2_integrand, lo, hi, self.prec["halo_precision"], breaks=br,
                                     args=(ln_k, norm))/(norm*self.rho_bar)
            self._i12 = (vals, _spline(self.ln_k_nodes, vals))
        return self._i12

    def i_1_2(self, k):                                   # halo.py:1169-1172
        k = np.asarray(k, dtype=float)
        return np.where((k >= self.k_min) & (k <= self.k_max), self.i_1_2_table()[1](np.log(k)), 0.0)

    def dln_power_ddelta_b(self, k):                      # halo.py:1138-1157
        k = np.asarray(k, dtype=float)
        inside = (k >= self.k_min) & (k <= self.k_max)
        hm = self._tab("h_m", k)
        with np.errstate(divide="ignore", invalid="ignore"):
            val = (68.0/21.0*hm*hm*self.linear_power(k) + self.i_1_2(k))/self.power("power_mm", k)
        return np.where(inside, val, 0.0)

    def power_mm_ssc(self, k):                            # halo.py:1159-1167
        return self.power("power_mm", k)*(1.0 + self.dln_power_ddelta_b(k)*self.delta_b)


class HaloExclusion(Halo):                                # halo.py:1201-1233
    exclusion = True


# ----------------------------------------------------------------------------
# kernel.dNdz*  (kernel.py:26-208)
# ----------------------------------------------------------------------------
class dNdz(object):
    def _init(self, z_min, z_max, prec, integ):
        self.prec = prec or DEFAULT_PRECISION
        self.integ = integ or Romberg(self.prec["divmax"])
        self.z_min, self.z_max = z_min, z_max
        self.normalize()

    def normalize(self):                                  # kernel.py:43-54
        self.norm = 1.0/self.integ(self.raw, self.z_min, self.z_max,
                                   self.prec["dNdz_precision"],
                                   breaks=self.smooth_marks())

    def smooth_marks(self):
        return ()

    def dndz(self, z):                                    # kernel.py:67-86
        z = np.asarray(z, dtype=float)
        return np.where((z <= self.z_max) & (z >= self.z_min),
                        self.norm*self.raw(z), 0.0)


class dNdzGaussian(dNdz):                                 # kernel.py:89-112
    kind = 0

    def __init__(self, z_min, z_max, z0, sigma_z, prec=None, integ=None):
        z_min = max(z_min, z0 - 8.0*sigma_z)
        z_max = min(z_max, z0 + 8.0*sigma_z)
        self.z0, self.sigma_z = z0, sigma_z
        self._init(z_min, z_max, prec, integ)

    def raw(self, z):
        return np.exp(-1.0*(z - self.z0)*(z - self.z0) /
                      (2.0*self.sigma_z*self.sigma_z))

    def smooth_marks(self):
        return self.z0 + self.sigma_z*np.arange(-8.0, 8.5, 1.0)

    def params(self):
        return [self.z0, self.sigma_z, 0.0]


class dNdzMagLim(dNdz):                                   # kernel.py:148-179
    kind = 1

    def __init__(self, z_min, z_max, a, z0, b, prec=None, integ=None):
        prec = prec or DEFAULT_PRECISION
        self.a, self.z0, self.b = a, z0, b
        if isinstance(b, (int, np.integer)):
            inv_b = 1//b                                  # Python 2 (Q2)
        else:
            inv_b = 1/b
        cap = np.power(-1*np.log(prec["dNdz_precision"]), inv_b)*z0
        if cap < z_max:
            z_max = cap
        self._init(z_min, z_max, prec, integ)

    def raw(self, z):
        return np.power(z, self.a)*np.exp(-1.0*np.power(z/self.z0, self.b))

    def smooth_marks(self):
        return np.linspace(self.z_min, self.z_max, 17)

    def params(self):
        return [float(self.a), self.z0, float(self.b)]


class dNdzInterpolation(dNdz):                            # kernel.py:181-208
    """p(z) tabulated at z_array, interpolated by FITPACK (scipy, the reference's own
    provider: third-party arithmetic)."""
    kind = 2

    def __init__(self, z_array, p_array, weights=None, interpolation_order=2,
                 smoothing=None, prec=None, integ=None):
        from scipy.interpolate import UnivariateSpline
        z_array = np.asarray(z_array, dtype=float)
        p_array = np.asarray(p_array, dtype=float)
        if smoothing is None:                             # kernel.py:197-200
            self._p_of_z = _IUS(z_array, p_array, w=weights,
                                                        k=interpolation_order)
        else:                                             # kernel.py:201-204
            self._p_of_z = UnivariateSpline(z_array, p_array, w=weights,
                                            k=interpolation_order, s=smoothing)
        self._init(z_array[0], z_array[-1], prec, integ)

    def raw(self, z):                                     # kernel.py:207-208
        return self._p_of_z(z)

    def smooth_marks(self):
        return self._p_of_z.get_knots()

    def piecewise(self):
        """(breaks[n+1], coef[n,4]): the spline as one cubic per knot interval,
        p(z) = c0 + c1 t + c2 t^2 + c3 t^3, t = z - breaks[i]."""
        from scipy.interpolate import PPoly
        pp = PPoly.from_spline(self._p_of_z._eval_args)
        keep = np.diff(pp.x) > 0
        c = pp.c[::-1, keep]                              # ascending powers
        coef = np.zeros((c.shape[1], 4))
        coef[:, :c.shape[0]] = c.T
        return np.concatenate([pp.x[:-1][keep], pp.x[-1:]]), coef


# ----------------------------------------------------------------------------
# kernel.WindowFunction*  (kernel.py:211-484)
# ----------------------------------------------------------------------------
class WindowFunction(object):
    def _init(self, z_min, z_max, cosmo, prec, integ):
        self.prec = prec or cosmo.prec
        self.integ = integ or cosmo.integ
        eps = self.prec["window_precision"]
        self.z_min = eps if z_min < eps else z_min        # kernel.py:236-238
        self.z_max = z_max
        self.set_cosmology_object(cosmo)

    def set_cosmology_object(self, cosmo):                # kernel.py:289-306
        eps = self.prec["window_precision"]
        self.cosmo = cosmo.regrid(self.z_min, self.z_max)
        self.chi_min = max(float(self.cosmo.comoving_distance(self.z_min)), eps)
        self.chi_max = float(self.cosmo.comoving_distance(self.z_max))
        self.chi_nodes = np.linspace(self.chi_min, self.chi_max,
                                     self.prec["window_npoints"])
        self._wf = None

    def rebuilt_on(self, cosmo):
        """Kernel.__init__ copies its windows and re-grids them on its own
        cosmology (kernel.py:592-608)."""
        other = _copy.copy(self)
        other.set_cosmology_object(cosmo)
        return other

    def _build(self):                                     # kernel.py:308-313
        self.wf_nodes = np.array([float(self.raw(c)) for c in self.chi_nodes])
        self._wf = _spline(self.chi_nodes, self.wf_nodes)

    def window_function(self, chi):                       # kernel.py:326-340
        if self._wf is None:
            self._build()
        chi = np.asarray(chi, dtype=float)
        return np.where((chi >= self.chi_min) & (chi <= self.chi_max),
                        self._wf(chi), 0.0)


class WindowFunctionGalaxy(WindowFunction):               # kernel.py:360-387
    kind = 0

    def __init__(self, dist, cosmo, prec=None, integ=None):
        self.dist = dist
        self._init(dist.z_min, dist.z_max, cosmo, prec, integ)

    def raw(self, chi):
        z = self.cosmo.redshift(chi)
        return self.dist.dndz(z)/self.cosmo.inv_hubble(z)


class WindowFunctionConvergence(WindowFunction):          # kernel.py:409-484
    kind = 1

    def __init__(self, dist, cosmo, prec=None, integ=None):
        self.dist = dist
        self._init(0.0, dist.z_max, cosmo, prec, integ)

    def set_cosmology_object(self, cosmo):
        WindowFunction.set_cosmology_object(self, cosmo)
        # kernel.py:437-441 computes g_chi_min once in __init__; it is kept
        # across set_cosmology_object there.  The oracle rebuilds it, which is
        # identical whenever the object is freshly built per cosmology.
        eps = self.prec["window_precision"]
        self.g_chi_min = max(float(self.cosmo.comoving_distance(self.dist.z_min)),
                             eps)

    def _lensing_integrand(self, chi, chi0):              # kernel.py:479-482
        z = self.cosmo.redshift(chi)
        return self.dist.dndz(z)/self.cosmo.inv_hubble(z)*(chi - chi0)/chi

    def raw(self, chi):                                   # kernel.py:443-477
        eps = self.prec["window_precision"]
        a = 1.0/(1.0 + self.cosmo.redshift(chi))
        lo = max(chi, self.g_chi_min)
        if lo <= eps:
            g = 0.0
        else:
            g = self.integ(self._lensing_integrand, lo, self.chi_max,
                           self.prec["window_precision"],
                           breaks=self.cosmo.chi_nodes, args=(chi,))
        g *= self.cosmo.H0*self.cosmo.H0*chi
        return 3.0/2.0*self.cosmo.om*g/a


# ----------------------------------------------------------------------------
# kernel.Kernel / GalaxyGalaxyLensingKernel  (kernel.py:559-839)
# ----------------------------------------------------------------------------
class Kernel(object):
    bessel_order = 0

    def __init__(self, ktheta_min, ktheta_max, window_a, window_b, cosmo,
                 prec=None, integ=None):
        self.prec = prec or cosmo.prec
        self.integ = integ or cosmo.integ
        self.ln_ktheta_min = np.log(ktheta_min)
        self.ln_ktheta_max = np.log(ktheta_max)
        self.cosmo = cosmo
        self.z_min = max(window_a.z_min, window_b.z_min)  # kernel.py:594-597
        self.z_max = min(window_a.z_max, window_b.z_max)
        self.wa = window_a.rebuilt_on(cosmo)
        self.wb = window_b.rebuilt_on(cosmo)
        eps = self.prec["window_precision"]
        self.chi_min = max(eps, float(cosmo.comoving_distance(self.z_min)))
        self.chi_max = float(cosmo.comoving_distance(self.z_max))
        n = self.prec["kernel_npoints"]
        self.ln_ktheta_nodes = np.linspace(self.ln_ktheta_min,
                                           self.ln_ktheta_max, n)
        self.bessel_limit = special.jn_zeros(
            self.bessel_order, self.prec["kernel_bessel_limit"])[-1]
        # kernel.py:635-639
        zg = np.linspace(self.z_min, self.z_max, n)
        # _find_z_bar always goes through Kernel._kernel_integrand, i.e. J0 (= 1 at
        # k theta = 0), also for the J2 kernel, which does not override it (kernel.py:635-639)
        self.z_bar = zg[np.argmax(self._weight(cosmo.comoving_distance(zg)))]
        self._k = None

    def bessel(self, x):
        return special.j0(x)

    def _weight(self, chi):
        D = self.cosmo.growth_factor(self.cosmo.redshift(chi))
        return self.wa.window_function(chi)*self.wb.window_function(chi)*D*D

    def _integrand(self, chi, ktheta):                    # kernel.py:707-712, 834-839
        return self._weight(chi)*self.bessel(ktheta*chi)

    def smooth_breaks(self, ktheta):
        if self.integ.name != "tight":
            return ()
        marks = [self.wa.chi_nodes, self.wb.chi_nodes, self.cosmo.chi_nodes,
                 [self.wa.chi_min, self.wa.chi_max, self.wb.chi_min,
                  self.wb.chi_max]]
        if ktheta > 0:
            marks.append(np.arange(1.0, self.bessel_limit + 2.0, 1.5)/ktheta)
        return np.concatenate([np.asarray(m, dtype=float) for m in marks])

    def raw_kernel(self, ln_ktheta):                      # kernel.py:678-705
        kt = np.exp(ln_ktheta)
        top = self.bessel_limit/kt
        if top >= self.chi_max:
            top = self.chi_max
        return self.integ(self._integrand, self.chi_min, top,
                          self.prec["kernel_precision"],
                          breaks=self.smooth_breaks(kt), args=(kt,))

    def _build(self):                                     # kernel.py:641-647
        self.kernel_nodes = np.array([self.raw_kernel(x)
                                      for x in self.ln_ktheta_nodes])
        self._k = _spline(self.ln_ktheta_nodes, self.kernel_nodes)

    def kernel(self, ln_ktheta):                          # kernel.py:714-729
        if self._k is None:
            self._build()
        x = np.asarray(ln_ktheta, dtype=float)
        return np.where(x < self.ln_ktheta_min, self._k(self.ln_ktheta_min),
                        np.where(x <= self.ln_ktheta_max, self._k(x), 0.0))


class GalaxyGalaxyLensingKernel(Kernel):                  # kernel.py:784-839
    bessel_order = 2

    def bessel(self, x):
        return special.jn(2, x)


# ----------------------------------------------------------------------------
# correlation.Correlation  (correlation.py:33-289)
# ----------------------------------------------------------------------------
def theta_bins(theta_min_deg, theta_max_deg, bins_per_decade=5.0):
    """correlation.py:69-90 (bin centres in radians)."""
    d2r = np.pi/180.0
    lmin = np.log10(theta_min_deg*d2r)
    lmax = np.log10(theta_max_deg*d2r)
    if theta_min_deg == theta_max_deg:
        return np.array([theta_min_deg*d2r])
    out = []
    u = np.floor(lmin)*bins_per_decade
    t = np.power(10.0, u/(1.0*bins_per_decade))
    while t < np.power(10.0, lmax):
        if t >= np.power(10.0, lmin) and t < np.power(10.0, lmax):
            out.append(10**(0.5*(np.log10(t) + (u + 1.0)/(1.0*bins_per_decade))))
        u += 1.0
        t = np.power(10.0, u/(1.0*bins_per_decade))
    return np.array(out)


class Correlation(object):
    def __init__(self, theta_min_deg, theta_max_deg, kernel, halo_factory,
                 power_spec="linear_power", bins_per_decade=5.0, prec=None,
                 integ=None, k_min=None, k_max=None):
        """``halo_factory(z)`` returns a Halo built at redshift z -- the
        state-free equivalent of ``halo.set_redshift(kernel.z_bar)``
        (correlation.py:103)."""
        self.prec = prec or kernel.prec
        self.integ = integ or kernel.integ
        self.theta = theta_bins(theta_min_deg, theta_max_deg, bins_per_decade)
        self.kernel = kernel
        self.D_z = float(kernel.cosmo.growth_factor(kernel.z_bar))
        self.halo = halo_factory(kernel.z_bar)
        if ((k_min is not None or k_max is not None) and
                not self.halo.extrapolate and
                ((k_min is None or k_min < self.halo.k_min) or     # Python 2: None < x is True,
                 (k_max is not None and k_max > self.halo.k_max))):  # None > x is False
            self.halo.extrapolate = True                  # correlation.py:104-107
        self.ln_k_min = np.log(self.halo.k_min if k_min is None else k_min)
        self.ln_k_max = np.log(self.halo.k_max if k_max is None else k_max)
        self.power_spec = power_spec

    def _integrand(self, ln_k, theta):                    # correlation.py:270
        k = np.exp(ln_k)
        return (k*k/(2.0*np.pi)*self.halo.power(self.power_spec, k) /
                (self.D_z*self.D_z)*self.kernel.kernel(np.log(k*theta)))

    def correlation(self, theta):                         # correlation.py:242
        theta = np.atleast_1d(np.asarray(theta, dtype=float))
        out = np.empty(theta.size)
        for i, t in enumerate(theta):
            breaks = ()
            if self.integ.name == "tight":
                breaks = np.concatenate([
                    self.halo.ln_k_nodes,
                    self.kernel.ln_ktheta_nodes - np.log(t)])
            out[i] = self.integ(self._integrand, self.ln_k_min, self.ln_k_max,
                                self.prec["corr_precision"], breaks=breaks,
                                args=(t,))
        return out

    def compute_correlation(self):                        # correlation.py:234
        self.wtheta = self.correlation(self.theta)
        return self.wtheta


# ----------------------------------------------------------------------------
# halo_trispectrum.HaloTrispectrumOneHalo  (halo_trispectrum.py:13-151)
# ----------------------------------------------------------------------------
TRISPECTRUM_MOMENT = {"power_mmmm": 0, "power_gmmm": 1, "power_ggmm": 2, "power_gggm": 3, "power_gggg": 4}


class HaloTrispectrumOneHalo(Halo):
    """I^0_4(k1, k1, k2, k2) = rho_bar^-3 int dln nu  nu f(nu) M^3 y(k1)^2 y(k2)^2 <moment>(M)
    on the n_k x n_k grid of ln k nodes.  The reference always integrates over the full
    [nu_min, nu_max] (the HOD-limited nu_min it computes is not used, Q18)."""

    def __init__(self, epoch, mass, hod, halo=None, power_spec="power_mmmm", **kw):
        Halo.__init__(self, epoch, mass, hod, halo, **kw)
        self.power_spec = power_spec

    def _moment(self, M):                                 # halo_trispectrum.py:141-151
        n = TRISPECTRUM_MOMENT.get(self.power_spec, 0)
        if n == 0:
            return np.ones(np.shape(M))
        return self.hod.nth_moment(M, n)

    def _i04_integrand(self, ln_nu, ln_k1, ln_k2):        # halo_trispectrum.py:131-140
        nu = np.exp(ln_nu)
        M = self.mass.mass(nu)
        y1, y2 = self.y(ln_k1, M), self.y(ln_k2, M)
        return nu*self.mass.f_nu(nu)*y1*y1*y2*y2*M*M*M*self._moment(M)

    def i_0_4(self, ln_k1, ln_k2):                        # halo_trispectrum.py:58-98
        lo, hi = np.log(self.mass.nu_min), np.log(self.mass.nu_max)
        br, sg = self._panel_hints(lo, hi)
        if TRISPECTRUM_MOMENT.get(self.power_spec, 0) == 0:
            sg = ()
        return self.integ(self._i04_integrand, lo, hi, self.prec["halo_precision"], breaks=br,
                          singular=sg, args=(ln_k1, ln_k2))/self.rho_bar**3

    def table_i04(self):                                  # halo_trispectrum.py:108-129
        n = self.ln_k_nodes.size
        T = np.empty((n, n))
        for i in range(n):
            for j in range(i, n):
                T[i, j] = T[j, i] = self.i_0_4(self.ln_k_nodes[i], self.ln_k_nodes[j])
        self._i04 = T
        return T

    def trispectrum_parallelogram(self, k1, k2):          # halo_trispectrum.py:53-57, 100-107
        from scipy.interpolate import RectBivariateSpline
        if not hasattr(self, "_i04"):
            self.table_i04()
        sp = RectBivariateSpline(self.ln_k_nodes, self.ln_k_nodes, self._i04, kx=3, ky=3, s=0)
        k1 = np.atleast_1d(np.asarray(k1, dtype=float))
        k2 = np.atleast_1d(np.asarray(k2, dtype=float))
        a = np.where(k1 < self.k_min, self.k_min, k1)
        b = np.where(k2 < self.k_min, self.k_min, k2)
        val = sp(np.log(a), np.log(b), grid=False)
        return np.where((a <= self.k_max) & (b <= self.k_max), val, 0.0)


# ----------------------------------------------------------------------------
# halo.HaloFit  (halo.py:1236-1412)
# ----------------------------------------------------------------------------
class HaloFit(Halo):
    """Takahashi et al. (2012) HALOFIT for power_mm; gm / gg use the halo-model
    tables with the HALOFIT spectrum as the 2-halo spectrum.  ``extrapolate`` is
    dropped by the reference's constructor (halo.py:1254-1259, Q17)."""

    def __init__(self, epoch, mass, hod, halo=None, fit_epoch=None, **kw):
        """``fit_epoch``: the SingleEpoch whose Omega_m(z), Omega_L(z) enter f_1..f_3.  The
        reference evaluates them once, in the constructor (halo.py:1261-1266), and never
        again -- not when Correlation*.__init__ moves the halo to z_bar -- so in the
        reference's own flow they belong to the construction redshift (z = 0 for
        ``halo.HaloFit()``, examples/shear_shear_spectrum.py:81)."""
        kw.pop("extrapolate", None)
        Halo.__init__(self, epoch, mass, hod, halo, **kw)
        epoch = fit_epoch or epoch
        om = epoch.omega_m()                                   # halo.py:1261-1266
        self.f1, self.f2, self.f3 = om**-0.0307, om**-0.0585, om**0.0743
        self.omega_l = epoch.omega_l()
        self.w = -1.0
        self._fit = None

    def _sigma2_integrand(self, ln_k, R):                      # halo.py:1321-1323
        k = np.exp(ln_k)
        return self.epoch.delta_k(k)*np.exp(-k*k*R*R)

    def _fit_params(self):                                     # halo.py:1268-1319
        if self._fit is not None:
            return self._fit
        n = self.prec["halo_npoints"]
        ln_R = np.linspace(np.log(0.1), np.log(10.0), n)
        lk0, lk1 = np.log(self.k_min), np.log(self.k_max)
        breaks = np.linspace(lk0, lk1, 33)
        ln_s2 = np.array([np.log(self.integ(self._sigma2_integrand, lk0, lk1, self.prec["halo_precision"],
                                            breaks=breaks, args=(np.exp(x),))) for x in ln_R])
        self.ln_R_nodes, self.ln_sigma2_nodes = ln_R, ln_s2
        k_s = 1.0/np.exp(_IUS(ln_s2[::-1], ln_R[::-1], k=3)(0.0))
        quintic = _IUS(ln_R, ln_s2, k=5)
        d1, d2 = quintic.derivatives(np.log(1.0/k_s))[1:3]
        ne, C = -d1 - 3.0, -d2
        ow = self.omega_l*(1 + self.w)
        self._fit = dict(
            k_s=k_s, n_eff=ne, C=C,
            a_n=10**(1.5222 + 2.8553*ne + 2.3706*ne*ne + 0.9903*ne**3 + 0.2250*ne**4 - 0.6038*C + 0.1749*ow),
            b_n=10**(-0.5642 + 0.5864*ne + 0.5716*ne*ne - 1.5474*C + 0.2279*ow),
            c_n=10**(0.3698 + 2.0404*ne + 0.8161*ne*ne + 0.5869*C),
            gamma_n=0.1971 - 0.0843*ne + 0.8460*C,
            alpha_n=np.fabs(6.0835 + 1.3373*ne - 0.1959*ne*ne - 5.5274*C),
            beta_n=2.0379 - 0.7354*ne + 0.3157*ne*ne + 1.2490*ne**3 + 0.3980*ne**4 - 0.1682*C,
            mu_n=0.0, nu_n=10**(5.2105 + 3.6902*ne))
        return self._fit

    def power_mm(self, k):                                     # halo.py:1325-1365
        p = self._fit_params()
        k = np.asarray(k, dtype=float)
        dk = self.epoch.delta_k(k)
        y = k/p["k_s"]
        dq = dk*(np.power(1 + dk, p["beta_n"])/(1 + p["alpha_n"]*dk)*np.exp(-(y/4.0 + y*y/8.0)))
        dh = (p["a_n"]*np.power(y, 3.0*self.f1)/(1.0 + p["b_n"]*np.power(y, self.f2) +
                                                 np.power(p["c_n"]*self.f3*y, 3.0 - p["gamma_n"])))
        dh = dh/(1.0 + p["mu_n"]/y + p["nu_n"]/(y*y))
        return 2.0*np.pi*np.pi/np.power(k, 3)*(dq + dh)

    def power_gm(self, k):                                     # halo.py:1367-1387
        return self.power_mm(k)*self._tab("h_g", k)*self._tab("h_m", k) + self._tab("pp_gm", k)

    def power_gg(self, k):                                     # halo.py:1400-1412
        return self.power_mm(k)*self._tab("h_g", k)*self._tab("h_g", k) + self._tab("pp_gg", k)


# ----------------------------------------------------------------------------
# correlation.CorrelationFourier  (correlation.py:297-405)
# ----------------------------------------------------------------------------
class Correlation3d(object):                              # correlation.py:408-510
    """xi(r) = int dln k k^2/(2 pi) P(k) J0(k r) over [ln k_min, ln k_max] -- the reference's integrand as written."""

    def __init__(self, r_min, r_max, halo, power_spec="linear_power", prec=None, integ=None):
        self.prec = prec or halo.prec
        self.integ = integ or halo.integ
        self.halo = halo
        self.r = np.logspace(np.log10(r_min), np.log10(r_max), self.prec["corr_npoints"])
        self.power_spec = power_spec
        self.ln_k_min, self.ln_k_max = np.log(halo.k_min), np.log(halo.k_max)

    def _integrand(self, ln_k, r):                        # correlation.py:493-500
        k = np.exp(ln_k)
        return k*k/(2.0*np.pi)*self.halo.power(self.power_spec, k)*special.j0(k*r)

    def raw_correlation(self, r):                         # correlation.py:467-491
        r = np.atleast_1d(np.asarray(r, dtype=float))
        out = np.empty(r.size)
        for i, v in enumerate(r):
            breaks = ()
            if self.integ.name == "tight":
                marks = [self.halo.ln_k_nodes]
                top = np.exp(self.ln_k_max)*v
                if top > 1.5:
                    marks.append(np.log(np.arange(1.0, top + 1.5, 1.5)/v))
                breaks = np.concatenate(marks)
            out[i] = self.integ(self._integrand, self.ln_k_min, self.ln_k_max, self.prec["corr_precision"],
                                breaks=breaks, args=(v,))
        return out


class CorrelationFourier(object):
    def __init__(self, kernel, halo_factory, power_spec="linear_power", prec=None, integ=None):
        self.prec = prec or kernel.prec
        self.integ = integ or kernel.integ
        self.kernel = kernel
        self.D_z = float(kernel.cosmo.growth_factor(kernel.z_bar))
        self.halo = halo_factory(kernel.z_bar)
        self.power_spec = power_spec

    def _integrand(self, chi, ell):                            # correlation.py:387-392
        k = self.kernel
        D = k.cosmo.growth_factor(k.cosmo.redshift(chi))
        return (self.halo.power(self.power_spec, ell/chi)/(self.D_z*self.D_z) *
                k.wa.window_function(chi)*k.wb.window_function(chi)*D*D/(chi*chi))

    def correlation(self, ell):                                # correlation.py:360-385
        ell = np.atleast_1d(np.asarray(ell, dtype=float))
        k = self.kernel
        out = np.empty(ell.size)
        for i, l in enumerate(ell):
            breaks = ()
            if self.integ.name == "tight":
                marks = [k.wa.chi_nodes, k.wb.chi_nodes, k.cosmo.chi_nodes]
                if self.power_spec != "linear_power" and not isinstance(self.halo, HaloFit):
                    marks.append(l/np.exp(self.halo.ln_k_nodes))
                elif self.power_spec in ("power_gm", "power_gg"):
                    marks.append(l/np.exp(self.halo.ln_k_nodes))
                breaks = np.concatenate([np.asarray(m, dtype=float) for m in marks])
            out[i] = self.integ(self._integrand, k.chi_min, k.chi_max, self.prec["corr_precision"],
                                breaks=breaks, args=(l,))
        return out
