// Shared device helpers: Gauss-Legendre tables, reductions, cosmology closed forms.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/chomp_b200.h"

namespace chomp {

#define CHOMP_MAX_GL 16
// nodes / weights on [-1, 1] for orders 1..16 (filled once per device by chomp_b200_create)
__constant__ double c_glx[CHOMP_MAX_GL + 1][CHOMP_MAX_GL];
__constant__ double c_glw[CHOMP_MAX_GL + 1][CHOMP_MAX_GL];

typedef chomp_b200_config Cfg;

// exp(x) for |x| < 700 to a few 1e-16 relative, with the constants as constant-bank operands:
// x = n ln 2 + r, |r| <= ln 2 / 2, degree-12 Taylor polynomial, 2^n added into the exponent field.
// No special cases (overflow, NaN, denormal results): the callers' arguments are table values.
#define EXPF_DEG 12
static const double h_k_expf[EXPF_DEG + 4] = {
    1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
    1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600};
__constant__ double k_expf[EXPF_DEG + 4];
static inline cudaError_t chomp_upload_expf_table() { return cudaMemcpyToSymbol(k_expf, h_k_expf, sizeof h_k_expf); }
__device__ __forceinline__ double exp_fast(double x) {
    const double n = rint(x * k_expf[0]);
    double r = fma(n, k_expf[1], x);
    r = fma(n, k_expf[2], r);
    double p = k_expf[3 + EXPF_DEG];
#pragma unroll
    for (int i = EXPF_DEG - 1; i >= 0; --i) p = fma(p, r, k_expf[3 + i]);
    return __hiloint2double(__double2hiint(p) + ((int)n * 1048576), __double2loint(p));
}


#define CHOMP_EPOCH_LEN 24
enum { EP_Z = 0, EP_GROWTH, EP_SIGMA_NORM, EP_DELTA_C, EP_DELTA_V, EP_RHO_BAR, EP_LNM_MIN, EP_LNM_MAX, EP_NU_MIN,
       EP_NU_MAX, EP_F_NORM, EP_B_NORM, EP_LNM_STAR, EP_PK_AMP, EP_CHI, EP_WALK,
       // scalars behind SingleEpoch's accessors (cosmology.py:366-447)
       EP_OMEGA_M, EP_OMEGA_L, EP_E0, EP_DELTA_V_COSMO, EP_RHO_CRIT, EP_FLAT, EP_OPEN, EP_SIGMA_8_Z };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums 16 per-lane values over the warp with 16 shuffles instead of 16 x 5: at each of the first
// four butterfly steps a lane keeps one half of its values and hands the other half to its
// partner.  Lanes 2j and 2j + 1 return the warp total of input index j.  Fixed order: reproducible.
template <int HALF>
__device__ __forceinline__ void warp_fold_step(double* v, int off) {
    const bool hi = (threadIdx.x & off) != 0;
#pragma unroll
    for (int j = 0; j < HALF; ++j) {
        const double keep = hi ? v[j + HALF] : v[j];
        const double send = hi ? v[j] : v[j + HALF];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
}
// eight values: lanes 4j .. 4j + 3 return the warp total of input index j
__device__ __forceinline__ double warp_fold8(double (&v)[8]) {
    warp_fold_step<4>(v, 16);
    warp_fold_step<2>(v, 8);
    warp_fold_step<1>(v, 4);
    double t = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 2);
    return t + __shfl_xor_sync(0xffffffffu, t, 1);
}
// eight values over each aligned group of eight lanes: lane j of a group returns the group total of
// input index j (a complete transpose-reduce in 7 shuffles)
__device__ __forceinline__ double group8_fold8(double (&v)[8]) {
    warp_fold_step<4>(v, 4);
    warp_fold_step<2>(v, 2);
    warp_fold_step<1>(v, 1);
    return v[0];
}
__device__ __forceinline__ double warp_fold16(double (&v)[16]) {
    warp_fold_step<8>(v, 16);
    warp_fold_step<4>(v, 8);
    warp_fold_step<2>(v, 4);
    warp_fold_step<1>(v, 2);
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// deterministic block-wide sum; `red` is shared scratch of >= 32 doubles; all threads get the result
__device__ inline double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

// ---------------------------------------------------------------------------
// cosmology closed forms (reference cosmology.py)
// ---------------------------------------------------------------------------
struct Cosmo {
    double om, ob, ol, orad, tcmb, h, s8, ns, H0;
    int flat, open;
    int bad;
};

__device__ inline Cosmo load_cosmo(const double* __restrict__ p, double cosmo_precision) {
    Cosmo c;
    c.om = p[CHOMP_C_OMEGA_M0]; c.ob = p[CHOMP_C_OMEGA_B0]; c.ol = p[CHOMP_C_OMEGA_L0];
    c.orad = p[CHOMP_C_OMEGA_R0]; c.tcmb = p[CHOMP_C_CMB_TEMP]; c.h = p[CHOMP_C_H];
    c.s8 = p[CHOMP_C_SIGMA_8]; c.ns = p[CHOMP_C_N_SCALAR];
    c.H0 = 100.0 / (2.998 * 100000.0);                           // cosmology.py:59
    const double tot = c.om + c.ol + c.orad;                       // cosmology.py:65-79
    c.flat = (tot <= 1.0 + cosmo_precision) && (tot >= 1.0 - cosmo_precision);
    c.open = (tot <= 1.0 - cosmo_precision);
    c.bad = (p[CHOMP_C_W0] != -1.0 || p[CHOMP_C_WA] != 0.0);       // dynamical dark energy: out of scope
    return c;
}

// (H(z)/H0)^2, no curvature term (cosmology.py:175-178): Omega_L + Omega_m / a^3 + Omega_r / a^4 with
// 1 / a = 1 + z multiplied out (no division)
__device__ __forceinline__ double E0(const Cosmo& c, double z) {
    const double y = 1.0 + z, y2 = y * y;
    return fma(c.orad * y2, y2, fma(c.om * y, y2, c.ol));
}
// 1/H(z) in Mpc/h (cosmology.py:153-162): reciprocal square root, 1 / H0 = 2998 Mpc/h
__device__ __forceinline__ double inv_hubble(const Cosmo& c, double z) { return rsqrt(E0(c, z)) * (2.998 * 100000.0 / 100.0); }

// Carroll et al. closed form, which is what growth_factor_eval returns (cosmology.py:215-231, 326)
__device__ __forceinline__ double growth_approx(const Cosmo& c, double a) {
    const double om = c.om / (a * a * a);
    const double den = c.ol + om;
    const double Om = om / den, Ol = c.ol / den;
    return (5.0 * Om / (2.0 / a)) / (Om * (4.0 / 7.0) - Ol + (1.0 + 0.5 * Om) * (1.0 + Ol / 70.0));
}

__device__ __forceinline__ double omega_m_z(const Cosmo& c, double z) {
    return c.om * (1.0 + z) * (1.0 + z) * (1.0 + z) / E0(c, z);
}
// cosmology.py:393-407
__device__ inline double delta_c_z(const Cosmo& c, double z) {
    double d = 0.15 * pow(12.0 * M_PI, 2.0 / 3.0);
    if (c.open) d *= pow(omega_m_z(c, z), 0.0185);
    if (c.flat && c.om < 1.0001) d *= pow(omega_m_z(c, z), 0.0055);
    return d;
}
// cosmology.py:409-423 (divided by the growth factor)
__device__ inline double delta_v_z(const Cosmo& c, double z, double growth) {
    double d = 178.0;
    if (c.open) d /= pow(omega_m_z(c, z), 0.7);
    if (c.flat && c.om < 1.0001) d /= pow(omega_m_z(c, z), 0.55);
    return d / growth;
}
// cosmology.py:425-447
__device__ __forceinline__ double rho_bar_z(const Cosmo& c, double z) {
    return 1.879 / 1.989 * (3.086 * 3.086 * 3.086) * 1e10 * E0(c, z) * omega_m_z(c, z);
}

// Eisenstein-Hu zero-baryon transfer function with the reference's Python-2
// arithmetic (cosmology.py:449-472: (Omb2)**(3/4) has exponent 0) and the
// dimensionless spectrum Delta^2(k) (cosmology.py:574-587).
// Eisenstein & Hu (1998) transfer function WITH baryon wiggles (SingleEpoch(with_bao=True), cosmology.py:474-538):
// the cosmology-level constants, built only when cfg.with_bao is set (CHOMP_ATTACH_BAO below)
struct BaoParams {
    double h, keq, s, ksilk, alpha_b, beta_b, alpha_c, beta_c, beta_node, fb, fc;
};

struct PkParams {
    double amp;      // delta_H^2 / h * growth^2 * sigma_norm^2
    double expo;     // 3 + n_scalar
    double ln_H0;
    double s, alpha, omh, theta;
    const BaoParams* bao;   // non-null: the wiggle transfer function replaces the zero-baryon one
};

__device__ inline PkParams make_pk(const Cosmo& c, double growth, double sigma_norm) {
    PkParams p;
    p.bao = nullptr;
    const double delta_H = 1.94e-5 * pow(c.om, -0.785 - 0.05 * log(c.om)) *
                           exp(-0.95 * (c.ns - 1.0) - 0.169 * (c.ns - 1.0) * (c.ns - 1.0));  // cosmology.py:83-85
    p.amp = delta_H * delta_H / c.h * growth * growth * sigma_norm * sigma_norm;
    p.expo = 3.0 + c.ns;
    p.ln_H0 = log(c.H0);
    const double omh2 = c.om * c.h * c.h;
    const double fb = c.ob / c.om;
    p.s = 44.5 * log(9.83 / omh2) / sqrt(1.0 + 10.0 * 1.0);
    p.alpha = 1.0 - 0.328 * log(431.0 * omh2) * fb + 0.38 * log(22.3 * omh2) * fb * fb;
    p.omh = c.om * c.h;
    p.theta = c.tcmb / 2.7;
    return p;
}

__device__ __forceinline__ double transfer_eh(const PkParams& p, double k) {
    // cosmology.py:466-472 with the two nested quotients cleared (two divisions instead of four):
    //   q = k theta / (Omega_m h (alpha + (1 - alpha) / t^4)),  T = L0 / (L0 + C0 q^2),
    //   C0 = 14.2 + 731 / (1 + 62.5 q)
    const double t = 1.0 + 0.43 * k * p.s;
    const double t2 = t * t, t4 = t2 * t2;
    const double q = k * p.theta * t4 / (p.omh * fma(p.alpha, t4, 1.0 - p.alpha));
    const double L0 = log(2.0 * M_E + 1.8 * q);
    const double u = fma(62.5, q, 1.0);
    const double L0u = L0 * u;
    return L0u / fma(fma(14.2, u, 731.0), q * q, L0u);
}

// Out of line, both of them: the wiggle form is a few hundred instructions of pow / log / sin that the default
// path must not carry in its instruction stream.
__device__ __noinline__ void make_bao(const Cosmo& c, BaoParams* b) {
    const double theta = c.tcmb / 2.7, t2 = theta * theta, t4 = t2 * t2;
    const double Oc = c.om - c.ob, h = c.h;
    const double Oh2 = c.om * h * h, Obh2 = c.ob * h * h, ObO = c.ob / c.om;
    const double zeq = 2.5e4 * Oh2 / t4;
    const double keq = 7.46e-2 * Oh2 / t2;
    double b1 = 0.313 * pow(Oh2, -0.419) * (1.0 + 0.607 * pow(Oh2, 0.674));
    double b2 = 0.238 * pow(Oh2, 0.223);
    const double zd = 1291.0 * (pow(Oh2, 0.251) / (1.0 + 0.659 * pow(Oh2, 0.828))) * (1.0 + b1 * pow(Obh2, b2));
    const double Req = 31.5 * Obh2 / t4 * (1000.0 / zeq), Rd = 31.5 * Obh2 / t4 * (1000.0 / zd);
    const double s = (2.0 / (3.0 * keq)) * sqrt(6.0 / Req) * log((sqrt(1.0 + Rd) + sqrt(Rd + Req)) / (1.0 + sqrt(Req)));
    const double y = (1.0 + zeq) / (1.0 + zd), sy = sqrt(1.0 + y);
    const double G = y * (-6.0 * sy + (2.0 + 3.0 * y) * log((sy + 1.0) / (sy - 1.0)));
    b->h = h; b->keq = keq; b->s = s;
    b->ksilk = 1.6 * pow(Obh2, 0.52) * pow(Oh2, 0.73) * (1.0 + pow(10.4 * Oh2, -0.95));
    b->alpha_b = 2.07 * keq * s * pow(1.0 + Rd, -0.75) * G;
    b->beta_b = 0.5 + ObO + (3.0 - 2.0 * ObO) * sqrt((17.2 * Oh2) * (17.2 * Oh2) + 1.0);
    const double a1 = pow(46.9 * Oh2, 0.670) * (1.0 + pow(32.1 * Oh2, -0.532));
    const double a2 = pow(12.0 * Oh2, 0.424) * (1.0 + pow(45.0 * Oh2, -0.582));
    b->alpha_c = pow(a1, -ObO) * pow(a2, -ObO * ObO * ObO);
    b1 = 0.944 / (1.0 + pow(458.0 * Oh2, -0.708));
    b2 = pow(0.395 * Oh2, -0.0266);
    b->beta_c = 1.0 / (1.0 + b1 * (pow(Oc / c.om, b2) - 1.0));
    b->beta_node = 8.41 * pow(Oh2, 0.435);
    b->fb = ObO; b->fc = Oc / c.om;
}
__device__ __noinline__ double transfer_bao(const BaoParams* b, double k) {
    const double kh = k * b->h, ks = kh * b->s, q = kh / (13.41 * b->keq), q2 = q * q;
    const double cq = 386.0 / (1.0 + 69.9 * pow(q, 1.08));
    // T0~(k, a, beta) = L / (L + C q^2),  L = ln(e + 1.8 beta q),  C = 14.2 / a + 386 / (1 + 69.9 q^1.08)   (:513-516)
    const double Lc = log(M_E + 1.8 * b->beta_c * q), L1 = log(M_E + 1.8 * q);
    const double t_c1 = Lc / (Lc + (14.2 + cq) * q2), t_ca = Lc / (Lc + (14.2 / b->alpha_c + cq) * q2);
    const double t_11 = L1 / (L1 + (14.2 + cq) * q2);
    const double r = ks / 5.4, r2 = r * r, f = 1.0 / (1.0 + r2 * r2);
    const double Tc = f * t_c1 + (1.0 - f) * t_ca;
    const double bn = b->beta_node / ks;
    const double stilde = b->s / cbrt(1.0 + bn * bn * bn);
    const double Tb1 = t_11 / (1.0 + (ks / 5.2) * (ks / 5.2));
    const double bb = b->beta_b / ks;
    const double Tb2 = (b->alpha_b / (1.0 + bb * bb * bb)) * exp(-pow(kh / b->ksilk, 1.4));
    const double x = k * stilde;                                   // k without the h that ks carries (cosmology.py:536)
    const double sinc = (x == 0.0) ? 1.0 : sin(x) / x;
    return b->fb * sinc * (Tb1 + Tb2) + b->fc * Tc;
}
__device__ __forceinline__ double transfer_any(const PkParams& p, double k) {
    return p.bao ? transfer_bao(p.bao, k) : transfer_eh(p, k);
}
// after `PkParams pk = make_pk(...)`: hang the wiggle constants on it when the configuration asks for them
#define CHOMP_ATTACH_BAO(cfg, cosmo_struct, pk_var)                         \
    BaoParams pk_var##_bao_store;                                            \
    if ((cfg).with_bao) { make_bao((cosmo_struct), &pk_var##_bao_store); (pk_var).bao = &pk_var##_bao_store; }

// Delta^2(k) = k^3 P(k) / (2 pi^2)
__device__ __forceinline__ double delta2(const PkParams& p, double k, double lnk) {
    const double T = transfer_any(p, k);
    return p.amp * exp(p.expo * (lnk - p.ln_H0)) * T * T;
}
// the same with the zero-baryon transfer function compiled in (default instantiations of the hot kernels)
__device__ __forceinline__ double delta2_eh(const PkParams& p, double k, double lnk) {
    const double T = transfer_eh(p, k);
    return p.amp * exp(p.expo * (lnk - p.ln_H0)) * T * T;
}
// P(k) (cosmology.py:589-600)
__device__ __forceinline__ double linear_power(const PkParams& p, double k) {
    if (!(k > 1e-16)) return 1e-16;
    return 2.0 * M_PI * M_PI * delta2(p, k, log(k)) / (k * k * k);
}

// Sheth-Tormen multiplicity and bias without their normalisations
// (mass_function.py:243-256, 290-303)
__device__ __forceinline__ void st_raw(double nu, double sta, double stq, double delta_c, double& nu_f, double& bias) {
    const double nup = nu * sta;
    const double pq = pow(nup, -stq);
    nu_f = (1.0 + pq) * sqrt(nup) * exp(-0.5 * nup);  // nu * f(nu) / f_norm
    bias = 1.0 + (nup - 1.0) / delta_c + 2.0 * stq / (delta_c * (1.0 + 1.0 / pq));
}

// The same from ln(nu): every power is the exponential of a known logarithm (a handful of
// constant-bank exp_fast calls instead of pow + sqrt + exp: the node loops stay small enough
// for the instruction cache).  ln_sta = ln(st_little_a).
__device__ __forceinline__ void st_raw_ln(double lnnu, double ln_sta, double stq, double delta_c, double& nu_f,
                                          double& bias) {
    const double lnup = lnnu + ln_sta;
    const double nup = exp_fast(lnup);
    const double pq = exp_fast(-stq * lnup);
    nu_f = (1.0 + pq) * exp_fast(0.5 * lnup - 0.5 * nup);
    bias = 1.0 + (nup - 1.0) / delta_c + 2.0 * stq / (delta_c * (1.0 + 1.0 / pq));
}

// ---------------------------------------------------------------------------------------------
// Tinker et al. (2010) mass function and bias (TinkerMassFunction, mass_function.py:436-564): cfg.mass_function_kind
// = CHOMP_MF_TINKER.  The five shape parameters are the reference's cubic splines in ln(Delta_v) through nine
// tabulated over-densities (piecewise-cubic form built on the host at create(), chomp_upload_tinker_tables),
// scaled with (1 + z); nu = (delta_c / sigma)^2 is the square of their variable.
// ---------------------------------------------------------------------------------------------
#define TINKER_N 9
__constant__ double k_tinker_breaks[TINKER_N];
__constant__ double k_tinker_coef[5][TINKER_N - 1][4];     // alpha, beta, gamma, phi, eta
static inline cudaError_t chomp_upload_tinker_tables() {
    const double delta[TINKER_N] = {200, 300, 400, 600, 800, 1200, 1600, 2400, 3200};
    const double tab[5][TINKER_N] = {{0.368, 0.363, 0.385, 0.389, 0.393, 0.365, 0.379, 0.355, 0.327},
                                     {0.589, 0.585, 0.544, 0.543, 0.564, 0.632, 0.637, 0.673, 0.702},
                                     {0.864, 0.922, 0.987, 1.09, 1.20, 1.34, 1.50, 1.68, 1.81},
                                     {-0.729, -0.789, -0.910, -1.05, -1.20, -1.26, -1.45, -1.50, -1.49},
                                     {-0.243, -0.261, -0.261, -0.273, -0.278, -0.301, -0.301, -0.319, -0.336}};
    const int n = TINKER_N;
    double x[TINKER_N], h[TINKER_N - 1], coef[5][TINKER_N - 1][4];
    for (int i = 0; i < n; ++i) x[i] = log(delta[i]);
    for (int i = 0; i < n - 1; ++i) h[i] = x[i + 1] - x[i];
    for (int t = 0; t < 5; ++t) {
        // not-a-knot cubic spline (what InterpolatedUnivariateSpline(k = 3) is): second derivatives M from a dense solve
        double A[TINKER_N][TINKER_N + 1] = {};
        const double* y = tab[t];
        A[0][0] = h[1]; A[0][1] = -(h[0] + h[1]); A[0][2] = h[0];
        for (int i = 1; i < n - 1; ++i) {
            A[i][i - 1] = h[i - 1]; A[i][i] = 2.0 * (h[i - 1] + h[i]); A[i][i + 1] = h[i];
            A[i][n] = 6.0 * ((y[i + 1] - y[i]) / h[i] - (y[i] - y[i - 1]) / h[i - 1]);
        }
        A[n - 1][n - 3] = h[n - 2]; A[n - 1][n - 2] = -(h[n - 3] + h[n - 2]); A[n - 1][n - 1] = h[n - 3];
        for (int c = 0; c < n; ++c) {                       // Gaussian elimination with partial pivoting
            int p = c;
            for (int r = c + 1; r < n; ++r) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
            for (int k = 0; k <= n; ++k) { const double tmp = A[c][k]; A[c][k] = A[p][k]; A[p][k] = tmp; }
            for (int r = c + 1; r < n; ++r) {
                const double f = A[r][c] / A[c][c];
                for (int k = c; k <= n; ++k) A[r][k] -= f * A[c][k];
            }
        }
        double M[TINKER_N];
        for (int r = n - 1; r >= 0; --r) {
            double v = A[r][n];
            for (int k = r + 1; k < n; ++k) v -= A[r][k] * M[k];
            M[r] = v / A[r][r];
        }
        for (int i = 0; i < n - 1; ++i) {
            coef[t][i][0] = y[i];
            coef[t][i][1] = (y[i + 1] - y[i]) / h[i] - h[i] * (2.0 * M[i] + M[i + 1]) / 6.0;
            coef[t][i][2] = 0.5 * M[i];
            coef[t][i][3] = (M[i + 1] - M[i]) / (6.0 * h[i]);
        }
    }
    cudaError_t e = cudaMemcpyToSymbol(k_tinker_breaks, x, sizeof x);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(k_tinker_coef, coef, sizeof coef);
}

// multiplicity / bias parameters of a point's mass function, whichever form is configured
struct MfParams {
    int kind;
    double ln_sta, stq, delta_c;                        // Sheth-Tormen
    double alpha, ln_beta, phi, eta, gamma;             // Tinker f(nu)
    double bA, ba, bC, dca;                             // Tinker b(nu): A, a, C, delta_c^a  (B = 0.183, b = 1.5, c = 2.4)
};
__device__ __noinline__ void tinker_params(double delta_v, double z, double delta_c, MfParams* p) {
    const double x = log(delta_v);
    int i = 0;
    while (i < TINKER_N - 2 && x >= k_tinker_breaks[i + 1]) ++i;          // end pieces extrapolate, as FITPACK does
    const double t = x - k_tinker_breaks[i];
    double v[5];
    for (int q = 0; q < 5; ++q) v[q] = k_tinker_coef[q][i][0] + t * (k_tinker_coef[q][i][1] + t * (k_tinker_coef[q][i][2] + t * k_tinker_coef[q][i][3]));
    const double lz = log(1.0 + z);                                       // mass_function.py:544-564
    p->alpha = v[0];
    p->ln_beta = log(v[1]) + 0.20 * lz;
    p->gamma = v[2] * exp(-0.01 * lz);
    p->phi = v[3] * exp(-0.08 * lz);
    p->eta = v[4] * exp(0.27 * lz);
    const double y = log10(delta_v), e4 = exp(-pow(4.0 / y, 4.0));       // mass_function.py:517-526
    p->bA = 1.0 + 0.24 * y * e4;
    p->ba = 0.44 * y - 0.88;
    p->bC = 0.019 + 0.107 * y + 0.19 * e4;
    p->dca = pow(delta_c, p->ba);
    p->delta_c = delta_c;
}
// nu f(nu) (unnormalised) and b(nu) / bias_norm from ln(nu)
__device__ __noinline__ void tinker_raw_ln(const MfParams& p, double lnnu, double& nu_f, double& bias) {
    nu_f = p.alpha * (1.0 + exp(-2.0 * p.phi * (p.ln_beta + 0.5 * lnnu))) * exp((p.eta + 0.5) * lnnu - 0.5 * p.gamma * exp(lnnu));
    const double sa = exp(0.5 * p.ba * lnnu);
    bias = 1.0 - p.bA * sa / (sa + p.dca) + 0.183 * exp(0.75 * lnnu) + p.bC * exp(1.2 * lnnu);
}
__device__ __forceinline__ MfParams mf_params(const Cfg& cfg, double sta, double stq, double delta_c, double delta_v, double z) {
    MfParams p;
    p.kind = cfg.mass_function_kind;
    p.stq = stq; p.delta_c = delta_c;
    p.ln_sta = log(sta);
    if (p.kind == CHOMP_MF_TINKER) tinker_params(delta_v, z, delta_c, &p);
    return p;
}
__device__ __forceinline__ void mf_raw_ln(const MfParams& p, double lnnu, double& nu_f, double& bias) {
    if (p.kind == CHOMP_MF_TINKER) tinker_raw_ln(p, lnnu, nu_f, bias);
    else st_raw_ln(lnnu, p.ln_sta, p.stq, p.delta_c, nu_f, bias);
}

}  // namespace chomp
