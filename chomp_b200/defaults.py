"""Default cosmology / halo / HOD dictionaries, integration limits and precision
settings -- the same module-level, mutable, read-at-call-time dictionaries and
keys as the reference's defaults.py:6-92 (callers reconfigure by assignment,
e.g. unit_test.py:17).

``*_npoints`` size the device tables.  The Romberg tolerances
(``*_precision``) are accepted for compatibility; on the device every integral
is a fixed-order Gauss-Legendre rule that is converged well below them
(orders in ``default_quadrature``).  Two of them still change results exactly as
in the reference: ``halo_precision`` enters HODZheng.first_moment_zero
(hod.py:176-179) and ``window_precision`` is the z / chi floor of the windows
(kernel.py:236, 301).
"""
default_cosmo_dict = {
    "omega_m0": 0.278 - 4.15e-5/0.7**2,
    "omega_b0": 0.046,
    "omega_l0": 0.722,
    "omega_r0": 4.15e-5/0.7**2,
    "cmb_temp": 2.726,
    "h": 0.7,
    "sigma_8": 0.811,
    "n_scalar": 0.960,
    "w0": -1.0,
    "wa": 0.0,
}

default_halo_dict = {
    "stq": 0.3,
    "st_little_a": 0.707,
    "c0": 9.0,
    "beta": -0.13,
    "alpha": -1,
    "delta_v": -1.0,
}

default_hod_dict = {
    "log_M_min": 12.14,
    "sigma": 0.15,
    "log_M_0": 12.14,
    "log_M_1p": 13.43,
    "alpha": 1.0,
}

default_limits = {
    "k_min": 0.001,
    "k_max": 100.0,
    "mass_min": -1,
    "mass_max": -1,
}

default_precision = {
    "corr_npoints": 50,
    "corr_precision": 1.48e-6,
    "cosmo_npoints": 50,
    "cosmo_precision": 1.48e-8,
    "dNdz_precision": 1.48e-8,
    "halo_npoints": 50,
    "halo_precision": 1.48e-5,
    "halo_limit": 100,
    "kernel_npoints": 50,
    "kernel_precision": 1.48e-6,
    "kernel_limit": 100,
    "kernel_bessel_limit": 8,
    "mass_npoints": 50,
    "mass_precision": 1.48e-8,
    "window_npoints": 100,
    "window_precision": 1.48e-6,
    "global_precision": 1.48e-32,
    "divmax": 20,
}

# Gauss-Legendre orders per panel used by the CUDA kernels (no reference
# counterpart: they replace the adaptive Romberg refinement).
default_quadrature = {
    "nu": 8,        # mass integrals, per knot interval of ln M(nu)
    "hankel": 4,    # w(theta) k-integral, per piece (<= 0.0625 wide in ln k) of a halo-table interval
    "limber": 4,    # K(ln k theta) chi-integral, per knot interval
    "lens": 6,      # lensing-efficiency integral, per chi(z) knot interval
    "cov_osc": 4,   # J0 J0 integrals of the covariance (K_NG table, Gaussian term), per piece
    "cov_phase": 3.0,  # largest phase advance of the faster Bessel factor over one such piece
    "cov_ng": 4,    # inner k_b integral of the non-Gaussian covariance term, per piece (<= 0.0625 wide in ln k)
}
