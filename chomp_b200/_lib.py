"""ctypes binding of libchomp_b200.so (the C ABI in include/chomp_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
usable, every compute entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHOMP_B200_LIB", os.path.join(_HERE, "libchomp_b200.so"))

# enums of include/chomp_b200.h
N_COSMO, N_HALO, N_HOD = 10, 6, 5
COSMO_KEYS = ("omega_m0", "omega_b0", "omega_l0", "omega_r0", "cmb_temp", "h",
              "sigma_8", "n_scalar", "w0", "wa")
HALO_KEYS = ("stq", "st_little_a", "c0", "beta", "alpha", "delta_v")
HOD_ZHENG_KEYS = ("log_M_min", "sigma", "log_M_0", "log_M_1p", "alpha")
HOD_MANDELBAUM_KEYS = ("log_M_0", "w")
HOD_ZHENG, HOD_MANDELBAUM = 0, 1
P_LINEAR, P_MM, P_GM, P_GG = 0, 1, 2, 3
MF_SHETH_TORMEN, MF_TINKER = 0, 1
TRISPECTRUM_MOMENT = {"power_mmmm": 0, "power_gmmm": 1, "power_ggmm": 2, "power_gggm": 3, "power_gggg": 4}
POWER_SPEC = {"linear_power": P_LINEAR, "power_mm": P_MM, "power_gm": P_GM,
              "power_mg": P_GM, "power_gg": P_GG}
DNDZ_GAUSSIAN, DNDZ_MAGLIM, DNDZ_TABLE = 0, 1, 2
WINDOW_GALAXY, WINDOW_CONVERGENCE = 0, 1
ST_NONFINITE, ST_MASS_WALK, ST_NODE_OVERFLOW, ST_DOMAIN = 1, 2, 4, 8
(EVAL_LINEAR_POWER, EVAL_SIGMA_R, EVAL_NU_OF_MASS, EVAL_MASS_OF_NU, EVAL_F_NU,
 EVAL_BIAS_NU, EVAL_KERNEL, EVAL_WINDOW_A, EVAL_WINDOW_B, EVAL_Y_NFW,
 EVAL_FIRST_MOMENT, EVAL_SECOND_MOMENT, EVAL_NTH_MOMENT, EVAL_HOD_ZEROS, EVAL_CONCENTRATION,
 EVAL_VIRIAL_RADIUS, EVAL_CHI_OF_Z, EVAL_Z_OF_CHI, EVAL_GROWTH_OF_Z, EVAL_INV_HUBBLE, EVAL_E0,
 EVAL_GROWTH_APPROX, EVAL_DNDZ_A, EVAL_DNDZ_B, EVAL_SIGMA_OF_NU, EVAL_BIAS_2_NU) = range(26)
(T_ZBAR, T_DBAR, T_KERNEL_NODES, T_CHI_NODES, T_WINDOW_NODES, T_WINDOW_CHI,
 T_EPOCH, T_LNM_NODES, T_NU_NODES, T_HALO_NODES, T_NBAR, T_NU_QUAD_COUNT, T_KERNEL_CHI,
 T_DNDZ_NORM, T_KNG, T_ZBAR_NG, T_D_NG, T_KNG_MIN, T_PROJECTED) = range(19)
HALOFIT_FIELDS = ("k_s", "n_eff", "C", "a_n", "b_n", "c_n", "gamma_n", "alpha_n", "beta_n", "mu_n", "nu_n",
                  "f_1", "f_2", "f_3", "omega_l", "fit_z")
KERNEL_NAMES = ("limber_tables_kernel", "mass_tables_kernel", "nu_nodes_kernel",
                "halo_sums_kernel", "halo_splines_kernel", "wtheta_kernel")
COV_KERNEL_NAMES = ("cov_kng_kernel", "tri_profile_kernel", "tri_gram_kernel", "cov_projected_kernel", "cov_g_kernel",
                    "cov_tri_nodes_kernel", "cov_ng_kernel", "cov_finish_kernel")
EPOCH_FIELDS = ("z", "growth", "sigma_norm", "delta_c", "delta_v", "rho_bar",
                "ln_mass_min", "ln_mass_max", "nu_min", "nu_max", "f_norm",
                "bias_norm", "ln_m_star", "pk_amp", "chi", "walk_steps", "omega_m", "omega_l", "E0",
                "delta_v_cosmo", "rho_crit", "flat", "open", "sigma_8_z")


class Config(ctypes.Structure):
    """Mirror of ``chomp_b200_config``."""
    _fields_ = [
        ("n_cosmo", ctypes.c_int32), ("n_mass", ctypes.c_int32),
        ("n_halo", ctypes.c_int32), ("n_window", ctypes.c_int32),
        ("n_kernel", ctypes.c_int32),
        ("nq_nu", ctypes.c_int32), ("nq_hankel", ctypes.c_int32),
        ("nq_limber", ctypes.c_int32), ("nq_lens", ctypes.c_int32),
        ("hod_kind", ctypes.c_int32), ("bessel_order", ctypes.c_int32),
        ("exclusion", ctypes.c_int32), ("extrapolate", ctypes.c_int32),
        ("window_kind", ctypes.c_int32*2), ("dndz_kind", ctypes.c_int32*2),
        ("tri_moment", ctypes.c_int32), ("use_halofit", ctypes.c_int32), ("with_bao", ctypes.c_int32),
        ("halo_precision", ctypes.c_double), ("cosmo_precision", ctypes.c_double),
        ("window_precision", ctypes.c_double),
        ("k_min", ctypes.c_double), ("k_max", ctypes.c_double),
        ("mass_min", ctypes.c_double), ("mass_max", ctypes.c_double),
        ("zk_min", ctypes.c_double), ("zk_max", ctypes.c_double),
        ("dndz_zmin", ctypes.c_double*2), ("dndz_zmax", ctypes.c_double*2),
        ("dndz_p", (ctypes.c_double*3)*2),
        ("ktheta_min", ctypes.c_double), ("ktheta_max", ctypes.c_double),
        ("bessel_limit", ctypes.c_double),
        ("corr_k_min", ctypes.c_double), ("corr_k_max", ctypes.c_double),
        ("dndz_table", ctypes.c_void_p*2), ("dndz_table_n", ctypes.c_int32*2),
        ("reserved_d", ctypes.c_double*1),
        ("mass_function_kind", ctypes.c_int32), ("reserved_tail", ctypes.c_int32*1),
    ]


class CovParams(ctypes.Structure):
    """Mirror of ``chomp_b200_cov_params``."""
    _fields_ = [
        ("n_bins", ctypes.c_int32), ("which", ctypes.c_int32), ("nongaussian", ctypes.c_int32),
        ("poisson_only", ctypes.c_int32), ("nq_osc", ctypes.c_int32), ("zero_last_ka", ctypes.c_int32),
        ("nq_ng", ctypes.c_int32), ("reserved_i", ctypes.c_int32*1),
        ("theta_min_rad", ctypes.c_double), ("theta_max_rad", ctypes.c_double), ("area_sr", ctypes.c_double),
        ("poisson", ctypes.c_double*6), ("shot_wt", ctypes.c_double*2), ("bessel_limit", ctypes.c_double),
        ("osc_phase", ctypes.c_double), ("halofit_z", ctypes.c_double), ("bin_log0", ctypes.c_double), ("bin_dlog", ctypes.c_double),
    ]


class ChompError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    "chomp_b200_version": (ctypes.c_int, []),
    "chomp_b200_last_error": (ctypes.c_char_p, []),
    "chomp_b200_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "chomp_b200_destroy": (None, [ctypes.c_void_p]),
    "chomp_b200_configure": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Config)]),
    "chomp_b200_set_dndz_table": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_void_p]),
    "chomp_b200_reserve": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "chomp_b200_limber_tables": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_mass_tables": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p]),
    "chomp_b200_halo_tables": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_power": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_wtheta": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "chomp_b200_wtheta_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_wtheta_batch_grouped": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                       ctypes.c_void_p]),
    "chomp_b200_wtheta_batch_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_void_p]),
    "chomp_b200_eval": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "chomp_b200_mass_second_order": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_void_p]),
    "chomp_b200_halofit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_cl": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_trispectrum_1h": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_trispectrum_eval": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_set_params": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_set_zbar": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_copy_table": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.POINTER(ctypes.c_int),
                                             ctypes.c_void_p]),
    "chomp_b200_dfma_peak": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_double)]),
    "chomp_b200_set_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "chomp_b200_get_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]),
    "chomp_b200_get_cov_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]),
    "chomp_b200_launch_count": (ctypes.c_longlong, [ctypes.c_void_p]),
    "chomp_b200_cov_kernel_ng": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(CovParams),
                                                ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_covariance": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(CovParams)] +
                              [ctypes.c_void_p]*10),
    "chomp_b200_halo_ssc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_xi3d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "chomp_b200_covariance_cross": (ctypes.c_int, [ctypes.c_void_p]*3 + [ctypes.c_int, ctypes.POINTER(CovParams)] +
                                    [ctypes.c_void_p]*14),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))


def load():
    """Load the shared library (once) and declare the argument types."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ChompError(
                "%s is missing: build it with `make -C chomp_b200/csrc` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`; there is "
                "no CPU fallback" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise ChompError("chomp_b200: %s" % load().chomp_b200_last_error().decode())
