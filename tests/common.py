"""Shared fixtures for the parity tests: the reference's unit-test dictionaries
(unit_test.py:17-119) and helpers that run the same inputs through the CPU
oracle."""
import numpy as np

from oracle import chomp_oracle as O
from oracle.quadrature import Tight

D2R = np.pi/180.0

C_DICT = {"omega_m0": 0.3 - 4.15e-5/0.7**2, "omega_b0": 0.046, "omega_l0": 0.7,
          "omega_r0": 4.15e-5/0.7**2, "cmb_temp": 2.726, "h": 0.7, "sigma_8": 0.8,
          "n_scalar": 0.960, "w0": -1.0, "wa": 0.0}
C_DICT_2 = dict(C_DICT, omega_m0=1.0 - 4.15e-5/0.7**2, omega_l0=0.0)
H_DICT = {"stq": 0.3, "st_little_a": 0.707, "c0": 9., "beta": -0.13, "alpha": -1.,
          "delta_v": -1.}
H_DICT_2 = {"stq": 0.5, "st_little_a": 0.5, "c0": 5., "beta": -0.2, "alpha": -1,
            "delta_v": 200.0}
HOD_DICT = {"log_M_min": 12.14, "sigma": 0.15, "log_M_0": 12.14, "log_M_1p": 13.43,
            "alpha": 1.0}
HOD_DICT_2 = {"log_M_min": 14.06, "sigma": 0.71, "log_M_0": 14.06, "log_M_1p": 14.80,
              "alpha": 1.0}


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)/np.maximum(np.abs(b), 1e-300)))


def w_err(w, ref, floor=1e-3):
    """Relative error of a correlation function that may cross zero: each bin
    relative to max(|ref_i|, floor * max|ref|)  (parity
    report)."""
    w, ref = np.asarray(w, dtype=float), np.asarray(ref, dtype=float)
    scale = np.maximum(np.abs(ref), floor*np.max(np.abs(ref)))
    return float(np.max(np.abs(w - ref)/scale))


def oracle_wtheta(cosmo, halo, hod, dist_a, dist_b=None, window_a="galaxy", window_b=None,
                  power_spec="power_gg", bins_per_decade=10.0, prec=None, integ=None,
                  z_range=(0.0, 5.0), theta_deg=(0.001, 1.0), bessel_order=0, hod_kind="zheng", k_min=None, k_max=None, with_bao=False):
    """One parameter point through the oracle; returns a dict of every table."""
    prec = prec or O.precision()
    integ = integ or Tight(40)

    def make_dist(spec):
        kind, args = spec
        cls = {"gaussian": O.dNdzGaussian, "table": O.dNdzInterpolation}.get(kind, O.dNdzMagLim)
        return cls(*args, prec=prec, integ=integ)

    def make_window(kind, dist, cm):
        cls = O.WindowFunctionGalaxy if kind == "galaxy" else O.WindowFunctionConvergence
        return cls(dist, cm)

    cm = O.MultiEpoch(z_range[0], z_range[1], cosmo, prec, integ)
    da = make_dist(dist_a)
    db = make_dist(dist_b) if dist_b is not None else da
    wa = make_window(window_a, da, cm)
    wb = make_window(window_b or window_a, db, cm)
    kcls = O.Kernel if bessel_order == 0 else O.GalaxyGalaxyLensingKernel
    kern = kcls(1e-6*D2R, 100*D2R, wa, wb, cm)
    hod_cls = O.HODZheng if hod_kind == "zheng" else O.HODMandelbaum

    def factory(z):
        se = O.SingleEpoch(z, cosmo, prec, integ, with_bao=with_bao)
        mf = O.MassFunction(se, halo)
        return O.Halo(se, mf, hod_cls(hod, prec["halo_precision"]), halo)

    corr = O.Correlation(theta_deg[0], theta_deg[1], kern, factory, power_spec,
                         bins_per_decade=bins_per_decade, k_min=k_min, k_max=k_max)
    w = corr.compute_correlation()
    h = corr.halo
    out = dict(w=w, theta=corr.theta, z_bar=kern.z_bar, D_z=corr.D_z,
               kernel_nodes=kern.kernel_nodes, chi_nodes=cm.chi_nodes,
               wa_nodes=kern.wa.wf_nodes, wb_nodes=kern.wb.wf_nodes,
               nu_nodes=h.mass.nu_nodes, ln_mass_nodes=h.mass.ln_mass_nodes,
               f_norm=h.mass.f_norm, bias_norm=h.mass.bias_norm,
               ln_m_star=np.log(h.mass.m_star), n_bar_over_rho_bar=h.n_bar_over_rho_bar,
               sigma_norm=h.epoch.sigma_norm, growth=h.epoch.growth, halo=h)
    for name in ("h_m", "pp_mm", "h_g", "pp_gm", "pp_gg"):
        out[name] = h.table(name)[0]
    return out


def oracle_covariance(cosmo, halo, hod, dist=("gaussian", (0.0, 2.0, 0.5, 0.1)), theta_deg=(0.01, 1.0),
                      bins_per_decade=5.0, tri_spec="power_gggg", cov_spec="power_gg", tri_z=None,
                      area_deg2=25.0, n_a=(1e10, 1e10), n_b=(1e10, 1e10), variance=1.0, prec=None, integ=None,
                      corr_bins_per_decade=5.0):
    """Config 5 through the oracle: galaxy x galaxy Limber kernel, Correlation, HaloTrispectrumOneHalo at
    ``tri_z`` (None: z_bar_NG, what Covariance.set_cosmology does) and Covariance of the correlation with itself."""
    from oracle import covariance_oracle as CO
    prec = prec or O.precision()
    integ = integ or Tight(16)
    kind, args = dist
    dcls = O.dNdzGaussian if kind == "gaussian" else O.dNdzMagLim
    cm = O.MultiEpoch(0.0, 5.0, cosmo, prec, integ)
    d = dcls(*args, prec=prec, integ=integ)
    wa, wb = O.WindowFunctionGalaxy(d, cm), O.WindowFunctionGalaxy(d, cm)
    kern = O.Kernel(1e-6*D2R, 100*D2R, wa, wb, cm)

    def factory(z, cls=O.Halo, **kw):
        se = O.SingleEpoch(z, cosmo, prec, integ)
        return cls(se, O.MassFunction(se, halo), O.HODZheng(hod, prec["halo_precision"]), halo, **kw)

    corr = O.Correlation(theta_deg[0], theta_deg[1], kern, factory, cov_spec, bins_per_decade=corr_bins_per_decade)
    cov = CO.Covariance(corr, theta_deg, bins_per_decade, area_deg2, n_a, n_b, variance, True, None, cov_spec)
    cov.tri = factory(cov.kernel.z_bar_NG if tri_z is None else tri_z, O.HaloTrispectrumOneHalo, power_spec=tri_spec)
    return cov


def cov_err(got, ref):
    """Relative error of a covariance block: each element relative to the geometric mean of the
    reference's diagonal (parity report)."""
    got, ref = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
    d = np.sqrt(np.abs(np.outer(np.diag(ref), np.diag(ref))))
    return float(np.max(np.abs(got - ref)/np.where(d > 0, d, 1.0)))


def interp_table():
    """A tabulated p(z) for the dNdzInterpolation cases: 25 samples of z^2 exp(-(z/0.45)^1.5)
    on [0.02, 1.7] (unnormalised, as a photometric-redshift histogram would be)."""
    z = np.linspace(0.02, 1.7, 25)
    return z, z*z*np.exp(-(z/0.45)**1.5)


def oracle_covariance_cross(cosmo, halo, hod_a, hod_b, dist_a, dist_b, hod_t=None, theta_deg=(0.01, 1.0), bins_per_decade=5.0,
                            tri_spec="power_gggg", cov_spec="power_gg", tri_z=None, area_deg2=25.0, n_a=(1e10, 1e10),
                            n_b=(1e10, 1e10), variance=1.0, prec=None, integ=None):
    """Covariance between two DIFFERENT galaxy-clustering correlations (Gaussian dN/dz `dist_a` / `dist_b`, HODs
    `hod_a` / `hod_b`) through the oracle; the trispectrum object carries `hod_t` (default `hod_a`)."""
    from oracle import covariance_oracle as CO
    prec = prec or O.precision()
    integ = integ or Tight(16)
    cm = O.MultiEpoch(0.0, 5.0, cosmo, prec, integ)

    def make_corr(dist, hod):
        d = O.dNdzGaussian(*dist, prec=prec, integ=integ)
        kern = O.Kernel(1e-6*D2R, 100*D2R, O.WindowFunctionGalaxy(d, cm), O.WindowFunctionGalaxy(d, cm), cm)

        def factory(z):
            se = O.SingleEpoch(z, cosmo, prec, integ)
            return O.Halo(se, O.MassFunction(se, halo), O.HODZheng(hod, prec["halo_precision"]), halo)
        return O.Correlation(theta_deg[0], theta_deg[1], kern, factory, cov_spec, bins_per_decade=bins_per_decade)

    corr_a, corr_b = make_corr(dist_a, hod_a), make_corr(dist_b, hod_b)
    cov = CO.Covariance(corr_a, theta_deg, bins_per_decade, area_deg2, n_a, n_b, variance, True, None, cov_spec, corr_b=corr_b)
    zt = cov.kernel.z_bar_NG if tri_z is None else tri_z
    se = O.SingleEpoch(zt, cosmo, prec, integ)
    cov.tri = O.HaloTrispectrumOneHalo(se, O.MassFunction(se, halo), O.HODZheng(hod_t or hod_a, prec["halo_precision"]), halo,
                                       power_spec=tri_spec)
    return cov
