/* chomp_b200 -- C ABI of the B200-native halo-model / Limber / Hankel hot path.
 *
 * The reference (morriscb/chomp) is pure Python and has no FFI; its boundary is
 * the Python class API.  This library is what the drop-in Python classes in
 * chomp_b200/ bind through ctypes (see INTEGRATION.md for the stubs).  Each
 * entry point names the reference code it replaces (paths relative to
 * /root/reference).
 *
 * Conventions
 *   - every array argument is FP64, row-major; "dev" pointers are device
 *     pointers owned by the caller, "host" pointers are host memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - return 0 on success, nonzero on a call-level failure (bad argument, CUDA
 *     error; text from chomp_b200_last_error); per-point numerical trouble is
 *     reported in `status[B]` bit flags and never aborts the batch;
 *   - the library never frees caller memory and only synchronises the device
 *     in create / reserve / destroy and in the *_host convenience calls.
 */
#ifndef CHOMP_B200_H
#define CHOMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHOMP_B200_VERSION 100

/* cosmology parameter columns: keys of defaults.default_cosmo_dict (defaults.py:6-19) */
enum {
    CHOMP_C_OMEGA_M0 = 0, CHOMP_C_OMEGA_B0, CHOMP_C_OMEGA_L0, CHOMP_C_OMEGA_R0,
    CHOMP_C_CMB_TEMP, CHOMP_C_H, CHOMP_C_SIGMA_8, CHOMP_C_N_SCALAR, CHOMP_C_W0,
    CHOMP_C_WA, CHOMP_N_COSMO
};
/* halo parameter columns: defaults.default_halo_dict (defaults.py:21-30) */
enum { CHOMP_H_STQ = 0, CHOMP_H_ST_LITTLE_A, CHOMP_H_C0, CHOMP_H_BETA, CHOMP_H_ALPHA,
       CHOMP_H_DELTA_V, CHOMP_N_HALO };
/* HOD parameter columns (hod.py:141-186 Zheng; hod.py:232-261 Mandelbaum: log_M_0, w) */
enum { CHOMP_HOD_ZHENG = 0, CHOMP_HOD_MANDELBAUM = 1 };
#define CHOMP_N_HOD 5
/* power spectra (halo.py:266, 277, 322, 391) */
enum { CHOMP_P_LINEAR = 0, CHOMP_P_MM = 1, CHOMP_P_GM = 2, CHOMP_P_GG = 3 };
/* redshift distributions (kernel.py:89-112, 148-179) and windows (kernel.py:360-387, 409-484) */
enum { CHOMP_DNDZ_GAUSSIAN = 0, CHOMP_DNDZ_MAGLIM = 1,
       CHOMP_DNDZ_TABLE = 2 /* dNdzInterpolation (kernel.py:181-208): piecewise cubic set with chomp_b200_set_dndz_table */ };
enum { CHOMP_WINDOW_GALAXY = 0, CHOMP_WINDOW_CONVERGENCE = 1 };
enum { CHOMP_MF_SHETH_TORMEN = 0, CHOMP_MF_TINKER = 1 };
/* per-point status bits */
enum {
    CHOMP_ST_NONFINITE = 1,      /* a result is NaN/Inf                                  */
    CHOMP_ST_MASS_WALK = 2,      /* mass-limit search left its bracket                   */
    CHOMP_ST_NODE_OVERFLOW = 4,  /* nu-quadrature node list exceeded its capacity        */
    CHOMP_ST_DOMAIN = 8          /* parameter outside the supported domain (w0/wa, alpha) */
};

/* Batch-invariant configuration: the reference's defaults.default_precision /
 * default_limits (defaults.py:42-92) plus the survey set-up the reference takes
 * as constructor arguments (kernel.py:89, 148, 372, 430, 584; correlation.py:65). */
typedef struct chomp_b200_config {
    /* table sizes: cosmo_npoints, mass_npoints, halo_npoints, window_npoints, kernel_npoints */
    int32_t n_cosmo, n_mass, n_halo, n_window, n_kernel;
    /* Gauss-Legendre orders per panel: nu mass integrals, Hankel, Limber, lensing efficiency */
    int32_t nq_nu, nq_hankel, nq_limber, nq_lens;
    int32_t hod_kind;         /* CHOMP_HOD_*                                             */
    int32_t bessel_order;     /* 0: Kernel (kernel.py:559), 2: GalaxyGalaxyLensingKernel */
    int32_t exclusion;        /* 1: HaloExclusion mass window (halo.py:1201-1233)        */
    int32_t extrapolate;      /* Halo(extrapolate=...) (halo.py:42)                      */
    int32_t window_kind[2];   /* CHOMP_WINDOW_* for window a, b                          */
    int32_t dndz_kind[2];     /* CHOMP_DNDZ_*                                            */
    int32_t tri_moment;       /* HaloTrispectrumOneHalo power_spec: 0 mmmm, 1 gmmm, 2 ggmm, 3 gggm, 4 gggg;
                                 -1: no trispectrum (its node list is then not built)              */
    int32_t use_halofit;      /* 1: HaloFit two-halo spectrum (halo.py:1236-1412); run chomp_b200_halofit first */
    int32_t with_bao;         /* 1: Eisenstein & Hu (1998) transfer function with baryon wiggles, SingleEpoch(with_bao=True)
                                 (cosmology.py:474-538, 556-571); 0: the zero-baryon form (:449-472) */
    double halo_precision;    /* enters HODZheng.first_moment_zero (hod.py:176-179)      */
    double cosmo_precision;   /* flat/open/closed test (cosmology.py:65-79)              */
    double window_precision;  /* z / chi floor of the windows (kernel.py:236, 301, 612)  */
    double k_min, k_max;      /* defaults.default_limits                                 */
    double mass_min, mass_max;/* > 0: fixed mass limits (mass_function.py:163-171)       */
    double zk_min, zk_max;    /* MultiEpoch(z_min, z_max) handed to Kernel               */
    double dndz_zmin[2], dndz_zmax[2]; /* after the constructors' clipping (kernel.py:101-104, 163-174) */
    double dndz_p[2][3];      /* Gaussian: z0, sigma_z, -; MagLim: a, z0, b              */
    double ktheta_min, ktheta_max;     /* Kernel(ktheta_min, ktheta_max, ...)            */
    double bessel_limit;      /* special.jn_zeros(order, kernel_bessel_limit)[-1]        */
    double corr_k_min, corr_k_max;     /* Correlation(k_min=, k_max=); <= 0: halo limits */
    /* CHOMP_DNDZ_TABLE: filled in by chomp_b200_configure from the tables given to
     * chomp_b200_set_dndz_table (whatever the caller puts here is ignored) */
    const double* dndz_table[2];
    int32_t dndz_table_n[2];
    double reserved_d[1];
    int32_t mass_function_kind; /* CHOMP_MF_*: Sheth-Tormen (MassFunction, mass_function.py:25-363) or Tinker et al. 2010
                                   (TinkerMassFunction, :436-564: no f(nu) normalisation, its own bias)            */
    int32_t reserved_tail[1];
} chomp_b200_config;

int chomp_b200_version(void);
const char* chomp_b200_last_error(void);

/* handle = per-device scratch + constant tables; one handle per thread/stream */
int chomp_b200_create(void** handle, int device);
void chomp_b200_destroy(void* handle);
int chomp_b200_configure(void* handle, const chomp_b200_config* cfg);
/* Tabulated redshift distribution -- replaces dNdzInterpolation.__init__ / raw_dndz
 * (kernel.py:191-208): the reference's FITPACK spline handed over in piecewise-polynomial form,
 * p(z) = c[i][0] + c[i][1] t + c[i][2] t^2 + c[i][3] t^3 with t = z - breaks[i] on
 * [breaks[i], breaks[i+1]] (end pieces extrapolate).  HOST pointers: breaks[n_intervals + 1],
 * coef[n_intervals][4]; copied.  which = 0 / 1: window a / b, 2: both.  Call before chomp_b200_configure
 * with dndz_kind[which] = CHOMP_DNDZ_TABLE. */
int chomp_b200_set_dndz_table(void* handle, int which, int n_intervals, const double* breaks_host,
                              const double* coef_host);
/* (re)allocate device scratch for batches of up to max_points parameter points */
int chomp_b200_reserve(void* handle, int max_points);

/* Stage 1 -- replaces MultiEpoch._initialize_splines (cosmology.py:787-817),
 * dNdz.normalize (kernel.py:43-54), WindowFunction._initialize_spline (kernel.py:308-313),
 * WindowFunctionConvergence.raw_window_function (kernel.py:443-477), Kernel.__init__ /
 * _find_z_bar / raw_kernel / _initialize_spline (kernel.py:584-705, 812-839).
 * cosmo_dev: [B, CHOMP_N_COSMO].  Leaves z_bar, D(z_bar) and the K(ln k theta) spline
 * in the handle. */
int chomp_b200_limber_tables(void* handle, int B, const double* cosmo_dev, int32_t* status_dev, void* stream);

/* Stage 2 -- replaces SingleEpoch._initialize_defaults / sigma_r / nu_m
 * (cosmology.py:93-119, 602-699) and MassFunction._set_mass_limits / _initialize_splines /
 * _normalize (mass_function.py:160-241).  z_dev: [B] redshifts, or NULL to use stage 1's
 * z_bar (what Correlation.__init__ does, correlation.py:103).  halo_dev: [B, CHOMP_N_HALO]. */
int chomp_b200_mass_tables(void* handle, int B, const double* cosmo_dev, const double* halo_dev,
                           const double* z_dev, int32_t* status_dev, void* stream);

/* Stage 3 -- replaces Halo._calculate_n_bar, _initialize_h_m / _h_g / _pp_mm / _pp_gm / _pp_gg
 * and y_nfw (halo.py:561-585, 674-707, 904-1086) with the HOD moments of hod.py:188-230,
 * 262-299.  hod_dev: [B, CHOMP_N_HOD] (Zheng: log_M_min, sigma, log_M_0, log_M_1p, alpha;
 * Mandelbaum: log_M_0, w, -, -, -). */
int chomp_b200_halo_tables(void* handle, int B, const double* halo_dev, const double* hod_dev,
                           int32_t* status_dev, void* stream);

/* Halo.linear_power / power_mm / power_gm / power_gg (halo.py:266-439) at k_dev[n_k]
 * for every point: P_out_dev [B, n_k]. */
int chomp_b200_power(void* handle, int B, int which, int n_k, const double* k_dev, double* P_out_dev, void* stream);

/* Correlation.correlation (correlation.py:242-275): w_out_dev [B, n_theta], theta in radians. */
int chomp_b200_wtheta(void* handle, int B, int which, int n_theta, const double* theta_dev, double* w_out_dev,
                      int32_t* status_dev, void* stream);

/* All four stages back to back on `stream` (one MCMC step per point,
 * examples/example_script.py:141-143): inputs and outputs already on the device. */
int chomp_b200_wtheta_batch(void* handle, int B, const double* cosmo_dev, const double* halo_dev,
                            const double* hod_dev, int which, int n_theta, const double* theta_dev,
                            double* w_out_dev, int32_t* status_dev, void* stream);

/* The MCMC fast / slow split: n_groups rows of (cosmology, halo) parameters are shared
 * by B points; group_index_dev[B] (int32, device) maps a point to its row.  Stages 1 and 2 -- everything that
 * depends on cosmology and mass function only: chi / growth tables, windows, K(ln k theta), sigma(M), nu(M),
 * normalisations -- run once per ROW, stages 3 and 4 per point.  An HOD-only batch at fixed cosmology (n_groups
 * = 1) or a grid of cosmologies x HODs costs the slow stages n_groups times instead of B times.  Results are
 * bit-identical to chomp_b200_wtheta_batch on the expanded [B, .] arrays.  The reference rebuilds everything
 * at every corr.set_cosmology / set_hod (halo.py:134-192).  An index outside [0, n_groups) flags the point
 * with CHOMP_ST_DOMAIN. */
int chomp_b200_wtheta_batch_grouped(void* handle, int n_groups, const double* cosmo_dev, const double* halo_dev, int B,
                                    const int32_t* group_index_dev, const double* hod_dev, int which, int n_theta,
                                    const double* theta_dev, double* w_out_dev, int32_t* status_dev, void* stream);

/* Same with HOST buffers: pinned staging, H2D copies, the four stages, D2H copy, one
 * stream synchronise.  This is the call the end-to-end benchmark times. */
int chomp_b200_wtheta_batch_host(void* handle, int B, const double* cosmo_host, const double* halo_host,
                                 const double* hod_host, int which, int n_theta, const double* theta_host,
                                 double* w_out_host, int32_t* status_host);

/* Element-wise evaluators behind the drop-in classes' scalar methods (point index p of the
 * last batch): SingleEpoch.linear_power / sigma_r (cosmology.py:589, 602), MassFunction.nu /
 * mass / f_nu / bias_nu (mass_function.py:243-346), Kernel.kernel (kernel.py:714). */
enum { CHOMP_EVAL_LINEAR_POWER = 0, CHOMP_EVAL_SIGMA_R, CHOMP_EVAL_NU_OF_MASS, CHOMP_EVAL_MASS_OF_NU,
       CHOMP_EVAL_F_NU, CHOMP_EVAL_BIAS_NU, CHOMP_EVAL_KERNEL, CHOMP_EVAL_WINDOW_A, CHOMP_EVAL_WINDOW_B,
       CHOMP_EVAL_Y_NFW /* x = mass, aux = ln k */ , CHOMP_EVAL_FIRST_MOMENT, CHOMP_EVAL_SECOND_MOMENT,
       CHOMP_EVAL_NTH_MOMENT /* aux = n */, CHOMP_EVAL_HOD_ZEROS /* x = 0,1,2: first_moment_zero, second_moment_zero, _safe_norm */,
       CHOMP_EVAL_CONCENTRATION, CHOMP_EVAL_VIRIAL_RADIUS /* halo.py:441-463 */,
       CHOMP_EVAL_CHI_OF_Z, CHOMP_EVAL_Z_OF_CHI, CHOMP_EVAL_GROWTH_OF_Z /* MultiEpoch accessors, cosmology.py:873-953 */,
       CHOMP_EVAL_INV_HUBBLE /* E(z), cosmology.py:153 */, CHOMP_EVAL_E0, CHOMP_EVAL_GROWTH_APPROX,
       CHOMP_EVAL_DNDZ_A, CHOMP_EVAL_DNDZ_B /* dNdz.dndz (aux != 0: raw_dndz), kernel.py:56-86 */,
       CHOMP_EVAL_SIGMA_OF_NU, CHOMP_EVAL_BIAS_2_NU /* MassFunctionSecondOrder, mass_function.py:389, 423-430;
                                                       after chomp_b200_mass_second_order */ };
int chomp_b200_eval(void* handle, int point, int what, int n, const double* x_dev, double aux, double* out_dev,
                    void* stream);

/* MassFunctionSecondOrder._initialize_splines / _normalize (mass_function.py:371-421) on top of the last
 * chomp_b200_mass_tables: the sigma(nu) spline and bias_2_norm, b2_norm_out_dev [B] (may be NULL). */
int chomp_b200_mass_second_order(void* handle, int B, double* b2_norm_out_dev, int32_t* status_dev, void* stream);

/* HaloFit._initialize_halo_fit / _initialize_sigma_spline (halo.py:1261-1319) for the epochs of
 * the last chomp_b200_mass_tables: k_s, n_eff, C and the Takahashi et al. (2012) coefficients,
 * params_out_dev [B, 16] (may be NULL).  fit_z: redshift at which Omega_m(z), Omega_L(z) enter
 * f_1..f_3 -- the reference evaluates them once at construction (z = 0 for halo.HaloFit());
 * < 0: the epoch's own redshift.  With cfg.use_halofit = 1, chomp_b200_power / _wtheta / _cl use
 * the HALOFIT spectrum as power_mm and as the two-halo spectrum of power_gm / power_gg. */
int chomp_b200_halofit(void* handle, int B, double fit_z, double* params_out_dev, int32_t* status_dev, void* stream);

/* CorrelationFourier.correlation (correlation.py:360-392): C(l) = int dchi P(l/chi)/D(z_bar)^2
 * W_a W_b D^2 / chi^2, cl_out_dev [B, n_ell], for any CHOMP_P_* spectrum (HaloFit when cfg.use_halofit). */
int chomp_b200_cl(void* handle, int B, int which, int n_ell, const double* ell_dev, double* cl_out_dev, void* stream);

/* HaloTrispectrumOneHalo._initialize_i_0_4 (halo_trispectrum.py:58-140): the 1-halo trispectrum
 * I^0_4(k_i, k_i, k_j, k_j) on the n_halo x n_halo grid of ln k nodes, T_out_dev [B, n_halo, n_halo].
 * Needs stages 2 and 3 (chomp_b200_mass_tables, chomp_b200_halo_tables) of the same batch.
 * The (k x nu)(nu x k) contraction runs on the FP64 tensor-core path (DMMA m8n8k4). */
int chomp_b200_trispectrum_1h(void* handle, int B, double* T_out_dev, void* stream);
/* HaloTrispectrumOneHalo.trispectrum_parallelogram (halo_trispectrum.py:53-57, 100-107) for
 * point `point` of the last chomp_b200_trispectrum_1h: bicubic not-a-knot interpolation of the
 * table (RectBivariateSpline kx = ky = 3, s = 0), k clamped below k_min, 0 above k_max. */
int chomp_b200_trispectrum_eval(void* handle, int point, int n, const double* k1_dev, const double* k2_dev,
                                double* out_dev, void* stream);

/* Store parameter rows in the handle without running a stage (NULL = leave as is): lets the
 * HOD / cosmology closed forms be evaluated through chomp_b200_eval on their own. */
int chomp_b200_set_params(void* handle, int B, const double* cosmo_dev, const double* halo_dev,
                          const double* hod_dev, void* stream);

/* Correlation.set_redshift (correlation.py:144-153): force z_bar, D(z_bar) follows from the
 * kernel cosmology's growth spline of the last chomp_b200_limber_tables. */
int chomp_b200_set_zbar(void* handle, int B, const double* z_dev, void* stream);

/* Copy one of the per-point tables of the last batch into out_dev [B, len]; *len_out
 * receives the row length.  Table ids below. */
enum {
    CHOMP_T_ZBAR = 0, CHOMP_T_DBAR, CHOMP_T_KERNEL_NODES, CHOMP_T_CHI_NODES /* [3, n_cosmo]: kernel, a, b */,
    CHOMP_T_WINDOW_NODES /* [2, n_window] */, CHOMP_T_WINDOW_CHI /* [2,2] chi_min, chi_max */,
    CHOMP_T_EPOCH /* 16 scalars, see chomp_b200.cu */, CHOMP_T_LNM_NODES, CHOMP_T_NU_NODES,
    CHOMP_T_HALO_NODES /* [5, n_halo]: h_m, pp_mm, h_g, pp_gm, pp_gg */, CHOMP_T_NBAR /* n_bar/rho_bar */,
    CHOMP_T_NU_QUAD_COUNT /* number of nu quadrature nodes per k class (as double) */,
    CHOMP_T_KERNEL_CHI /* Kernel.chi_min, chi_max */, CHOMP_T_DNDZ_NORM /* dNdz.norm of both distributions */,
    /* covariance tables (after chomp_b200_cov_kernel_ng / chomp_b200_covariance) */
    CHOMP_T_KNG /* [n_kernel, n_kernel] K_NG */, CHOMP_T_ZBAR_NG, CHOMP_T_D_NG, CHOMP_T_KNG_MIN,
    CHOMP_T_PROJECTED /* [n_kernel] projected spectrum nodes (covariance.py:455-543) */
};
int chomp_b200_copy_table(void* handle, int B, int table, double* out_dev, int* len_out, void* stream);

/* Measured FP64 FMA peak of the device (dependent-free DFMA chains on all SMs): the
 * roofline denominator of bench.py, MEASURED_PEAKS.json carries no FP64 figure. */
int chomp_b200_dfma_peak(void* handle, int iters, double* tflops_out);

/* Per-kernel device times of the last chomp_b200_wtheta_batch on this handle, measured with
 * CUDA events recorded on the launching stream around every kernel (order below).  With
 * timing on, the events are recorded at every call; get_timing synchronises on the last. */
enum { CHOMP_K_LIMBER = 0, CHOMP_K_MASS, CHOMP_K_NODES, CHOMP_K_SUMS, CHOMP_K_SPLINES, CHOMP_K_WTHETA,
       CHOMP_N_KERNELS };
int chomp_b200_set_timing(void* handle, int on);
int chomp_b200_get_timing(void* handle, double* ms_out /* [CHOMP_N_KERNELS] */);
/* Same for the kernels of the last chomp_b200_covariance (classes below; a class launched several times -- the
 * chunked trispectrum / non-Gaussian kernels, the mass tables at two redshifts -- reports the sum). */
enum { CHOMP_KC_KNG = 0 /* cov_kng_kernel + its spline */, CHOMP_KC_TRI_PROFILE, CHOMP_KC_TRI_GRAM, CHOMP_KC_PROJECTED,
       CHOMP_KC_GAUSS, CHOMP_KC_TRI_NODES, CHOMP_KC_NG, CHOMP_KC_FINISH, CHOMP_N_COV_KERNELS };
int chomp_b200_get_cov_timing(void* handle, double* ms_out /* [CHOMP_N_COV_KERNELS + CHOMP_N_KERNELS]: the classes above, then
                                                               the w(theta)-path kernels launched inside the call */);

/* ---- covariance of w(theta): covariance.Covariance (covariance.py:23-683) over kernel.KernelCovariance
 * (kernel.py:864-1111) and HaloTrispectrumOneHalo, for input_correlation_a is input_correlation_b
 * (``matching_corrs``, covariance.py:60-63: the case of BASELINE config 5). ------------------------ */
typedef struct chomp_b200_cov_params {
    int32_t n_bins;          /* annular bins (covariance.py:53-74), <= 128                          */
    int32_t which;           /* CHOMP_P_*: Covariance(power_spec=), the projected spectrum          */
    int32_t nongaussian;     /* nongaussian_cov                                                      */
    int32_t poisson_only;    /* poisson_noise_only                                                   */
    int32_t nq_osc;          /* Gauss-Legendre order of the pieces of the J0 J0 integrals            */
    int32_t zero_last_ka;    /* outcome of exp(ln k_max) > k_max in the caller's arithmetic: the
                                reference's last ln k_a node then sees T = 0 (halo_trispectrum.py:100-107) */
    int32_t nq_ng;           /* Gauss-Legendre order per piece (<= 0.0625 in ln k_b) of the inner k_b integral of
                                the non-Gaussian term; <= 0: cfg.nq_hankel.  The integrand is a bicubic in ln k_b
                                times a smooth kernel: order 2 is converged to 1e-8                       */
    int32_t reserved_i[1];
    double theta_min_rad, theta_max_rad; /* 10**log_theta_min/max of the correlation (covariance.py:93-97) */
    double area_sr;          /* survey_area_deg2 * deg2_to_strad                                     */
    double poisson[6];       /* proj_power_poisson(window_pair = 0..5), covariance.py:352-357        */
    double shot_wt[2];       /* 1 + cosmic_shear[0], 1 + cosmic_shear[1], covariance.py:336-347      */
    double bessel_limit;     /* special.jn_zeros(0, kernel_bessel_limit)[-1]                         */
    double osc_phase;        /* largest phase advance of the fast Bessel factor over one piece      */
    double halofit_z;        /* cfg.use_halofit: fit_z of chomp_b200_halofit (the HaloFit object's
                                construction redshift, halo.py:1261-1266; < 0: each epoch's own)        */
    double bin_log0, bin_dlog; /* ln(center of bin 0) and the spacing of ln(center) when the bins are log-spaced to
                                rounding (what Covariance.__init__ builds, covariance.py:53-74): the non-Gaussian term then
                                shares its kernel values between the bins (cov_ng_shift_kernel).  bin_dlog <= 0: any bins */
} chomp_b200_cov_params;

/* KernelCovariance._find_z_bar / _initialize_NG_spline (kernel.py:961-972, 1016-1068) for the batch of
 * the last chomp_b200_limber_tables: z_bar_NG, D(z_bar_NG), the K_NG table and its log-space bicubic. */
int chomp_b200_cov_kernel_ng(void* handle, int B, const chomp_b200_cov_params* p, int32_t* status_dev, void* stream);

/* Covariance.get_covariance (covariance.py:276-321) for every point: Limber tables, K_NG, the 1-halo
 * trispectrum at tri_z_dev[B] (NULL: z_bar_NG, what Covariance.set_cosmology does, covariance.py:258;
 * needs cfg.tri_moment >= 0), the halo tables at z_bar, the projected spectra, then the Poisson,
 * Gaussian and non-Gaussian terms.  bin_center_dev / bin_delta_dev: [n_bins] AnnulusBin.center / .delta
 * in radians, ascending.  cov_out_dev [B, n_bins, n_bins]; parts_out_dev [B, 3, n_bins, n_bins] (P, G,
 * NG; may be NULL).  On return the handle holds the same state as after stages 1-3 of the w(theta) path
 * at z_bar, so chomp_b200_wtheta / chomp_b200_power may follow. */
int chomp_b200_covariance(void* handle, int B, const chomp_b200_cov_params* p, const double* bin_center_dev,
                          const double* bin_delta_dev, const double* tri_z_dev, const double* cosmo_dev,
                          const double* halo_dev, const double* hod_dev, double* cov_out_dev, double* parts_out_dev,
                          int32_t* status_dev, void* stream);

/* Covariance between two DIFFERENT correlations (covariance.py:60-63 ``matching_corrs`` False; the off-diagonal
 * blocks of covariance.CovarianceMulti, covariance.py:794-871).  handle_a / handle_b are configured with the windows
 * and the halo model (HOD kind, exclusion, HaloFit) of correlation a / b and share table sizes, k limits and the
 * MultiEpoch range; handle_t (NULL: handle_a) is configured with the trispectrum object's HOD kind and tri_moment.
 * K_NG uses all four windows (kernel.py:1102-1111), the Gaussian term the four projected spectra P_a, P_b, P_ab,
 * P_ba (covariance.py:421-453, 455-591); the Poisson term vanishes (covariance.py:313-315).  halo_t_dev / hod_t_dev:
 * the trispectrum object's own parameters (NULL: those of correlation a).  Outputs as chomp_b200_covariance. */
int chomp_b200_covariance_cross(void* handle_a, void* handle_b, void* handle_t, int B, const chomp_b200_cov_params* p,
                                const double* bin_center_dev, const double* bin_delta_dev, const double* tri_z_dev,
                                const double* cosmo_dev, const double* halo_a_dev, const double* hod_a_dev,
                                const double* halo_b_dev, const double* hod_b_dev, const double* halo_t_dev,
                                const double* hod_t_dev, double* cov_out_dev, double* parts_out_dev, int32_t* status_dev,
                                void* stream);

/* HaloSuperSampleCovariance (halo.py:1089-1199) for the batch of the last chomp_b200_halo_tables: what = 0 the table
 * I^1_2(k) = rho_bar^-1 int dln nu nu f(nu) b(nu) y(k, M)^2 M (halo.py:1159-1199, splined in ln k), what = 1
 * dln P / d delta_b = (68/21 h_m^2 P_lin + I^1_2) / P_mm (halo.py:1138-1157); both 0 outside [k_min, k_max].
 * k_dev [n_k], out_dev [B, n_k]. */
int chomp_b200_halo_ssc(void* handle, int B, int what, int n_k, const double* k_dev, double* out_dev, int32_t* status_dev,
                        void* stream);

/* Correlation3d.raw_correlation (correlation.py:467-500): xi(r) = int dln k k^2 / (2 pi) P(k) J0(k r) over the halo
 * model's k range for any CHOMP_P_* spectrum of the last batch; r_dev [n_r] in Mpc/h, xi_out_dev [B, n_r]. */
int chomp_b200_xi3d(void* handle, int B, int which, int n_r, const double* r_dev, double* xi_out_dev, int32_t* status_dev,
                    void* stream);

/* number of kernel launches issued by this handle since creation (bench.py's gpu_launches) */
long long chomp_b200_launch_count(void* handle);

#ifdef __cplusplus
}
#endif
#endif /* CHOMP_B200_H */
