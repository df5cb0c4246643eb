// Device special functions for the halo-model hot path (FP64).
//
//   sici          sine / cosine integrals  -- replaces scipy.special.sici in the halo-exclusion
//                 window (halo.py:1232) and in the element-wise Halo.y accessor
//   nfw_rho_tab   NFW Fourier-profile numerator (reference halo.py:574-583) for the hot loops:
//                 branch-free, range tables of the Si/Ci auxiliary functions in shared memory
//   sincos_reduced, sici_series_c          building blocks with constant-bank coefficients
//   bessel_j      J0 / J2                  -- replaces scipy.special.j0 / jn(2, .)
//                 (reference kernel.py:712, 839): piecewise polynomials below the 8th zero
//
// Coefficients come from tools/gen_special.py, gen_nfw_tables.py, gen_bessel_tables.py
// (Chebyshev interpolants built with mpmath; errors printed by the generators).
#pragma once
#include <math.h>
#include "special_coeffs.cuh"
#include "nfw_coeffs.cuh"
#include "bessel_coeffs.cuh"

namespace chomp {

#define CHOMP_EULER 0.57721566490153286061
#define CHOMP_PI_2 1.57079632679489661923

// coefficient rows padded to an odd number of doubles so that lanes selecting
// different ranges hit different shared-memory banks
#define CHOMP_SICI_ROW (CHOMP_SICI_DEG_L + 2)

struct SiciTables {
    double F[3][CHOMP_SICI_ROW];
    double G[3][CHOMP_SICI_ROW];
    double urange[3][2];
    double si_small[CHOMP_SICI_DEG_S + 1];
    double ci_small[CHOMP_SICI_DEG_S + 1];
};

// cooperative copy of the coefficient tables from constant to shared memory
__device__ inline void sici_tables_load(SiciTables* t) {
    for (int i = threadIdx.x; i < 3 * (CHOMP_SICI_DEG_L + 1); i += blockDim.x) {
        int r = i / (CHOMP_SICI_DEG_L + 1), j = i % (CHOMP_SICI_DEG_L + 1);
        t->F[r][j] = k_sici_F[r][j];
        t->G[r][j] = k_sici_G[r][j];
    }
    for (int i = threadIdx.x; i < 6; i += blockDim.x) t->urange[i / 2][i % 2] = k_sici_urange[i / 2][i % 2];
    for (int i = threadIdx.x; i <= CHOMP_SICI_DEG_S; i += blockDim.x) {
        t->si_small[i] = k_si_small[i];
        t->ci_small[i] = k_ci_small[i];
    }
}

// x <= CHOMP_SICI_SMALL_X:  Si(x) = x*P(x^2),  Ci(x) - ln(x) = gamma + x^2*Q(x^2)
__device__ __forceinline__ void sici_series(const SiciTables* t, double x, double& si, double& ci_nolog) {
    const double xx = x * x;
    const double s = (xx - CHOMP_SI_SMALL_MID) * CHOMP_SI_SMALL_IHALF;
    double p = t->si_small[CHOMP_SICI_DEG_S];
    double q = t->ci_small[CHOMP_SICI_DEG_S];
#pragma unroll
    for (int i = CHOMP_SICI_DEG_S - 1; i >= 0; --i) {
        p = fma(p, s, t->si_small[i]);
        q = fma(q, s, t->ci_small[i]);
    }
    si = x * p;
    ci_nolog = fma(xx, q, CHOMP_EULER);
}

// x > CHOMP_SICI_SMALL_X, with sin(x), cos(x) supplied by the caller
__device__ __forceinline__ void sici_aux(const SiciTables* t, double x, double sx, double cx, double& si, double& ci) {
    const double ix = 1.0 / x;
    const double u = ix * ix;
    const int r = (x >= CHOMP_SICI_X2) ? 2 : ((x >= CHOMP_SICI_X1) ? 1 : 0);
    const double s = (u - t->urange[r][0]) * t->urange[r][1];
    const double* __restrict__ cf = t->F[r];
    const double* __restrict__ cg = t->G[r];
    double f = cf[CHOMP_SICI_DEG_L];
    double g = cg[CHOMP_SICI_DEG_L];
#pragma unroll
    for (int i = CHOMP_SICI_DEG_L - 1; i >= 0; --i) {
        f = fma(f, s, cf[i]);
        g = fma(g, s, cg[i]);
    }
    f *= ix;
    g *= u;
    si = CHOMP_PI_2 - f * cx - g * sx;
    ci = f * sx - g * cx;
}

// General-purpose entry (x > 0), used outside the hot loop.
__device__ inline void sici(const SiciTables* t, double x, double& si, double& ci) {
    if (x <= CHOMP_SICI_SMALL_X) {
        double c0;
        sici_series(t, x, si, c0);
        ci = c0 + log(x);
    } else {
        double sx, cx;
        sincos(x, &sx, &cx);
        sici_aux(t, x, sx, cx, si, ci);
    }
}

// NFW Fourier profile numerator (reference halo.py:574-583):
//   cos z [Ci((1+c)z) - Ci(z)] + sin z [Si((1+c)z) - Si(z)] - sin(cz)/((1+c)z)
// lncp = ln(1+c); the caller divides by  ln(1+c) - c/(1+c).
__device__ __forceinline__ double nfw_rho_k(const SiciTables* t, double z, double cp, double lncp) {
    const double z2 = cp * z;
    double s1, c1, s2, c2;
    sincos(z, &s1, &c1);
    sincos(z2, &s2, &c2);
    const double sin_cz = s2 * c1 - c2 * s1;  // sin((1+c)z - z)
    double dsi, dci;
    if (z2 <= CHOMP_SICI_SMALL_X) {
        double si1, ci1, si2, ci2;
        sici_series(t, z, si1, ci1);
        sici_series(t, z2, si2, ci2);
        dsi = si2 - si1;
        dci = lncp + (ci2 - ci1);
    } else if (z <= CHOMP_SICI_SMALL_X) {
        double si1, ci1, si2, ci2;
        sici_series(t, z, si1, ci1);
        sici_aux(t, z2, s2, c2, si2, ci2);
        dsi = si2 - si1;
        dci = ci2 - (ci1 + log(z));
    } else {
        double si1, ci1, si2, ci2;
        sici_aux(t, z, s1, c1, si1, ci1);
        sici_aux(t, z2, s2, c2, si2, ci2);
        dsi = si2 - si1;
        dci = ci2 - ci1;
    }
    return c1 * dci + s1 * dsi - sin_cz / z2;
}

// sin and cos for 0 <= x < 1e5: three-term Cody-Waite reduction by pi/2 with FMAs, then the
// classic degree-13 / degree-14 minimax kernels on |r| <= pi/4 (max abs error 1.2e-16, checked
// against mpmath in tools/gen_special.py's companion test).  The arguments k r_s (1 + c) of the
// halo tables stay below ~1e3, so the large-argument machinery of the library routine (and its
// register / instruction footprint in the hot loop) is not needed.  The constants live in a
// __constant__ array filled at create(): literals would be materialised with two UMOVs per DFMA,
// constant-bank operands cost no issue slot.
enum { SC_2_PI = 0, SC_PIO2_1, SC_PIO2_2, SC_PIO2_3, SC_S0, SC_C0 = SC_S0 + 6, SC_N = SC_C0 + 6 };
static const double h_k_sincos[SC_N] = {
    0.63661977236758134308, 1.5707963267948966, 6.123233995736766e-17, -1.4973849048591698e-33,
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};
__constant__ double k_sincos[SC_N];
static inline cudaError_t chomp_upload_sincos_table() {
    return cudaMemcpyToSymbol(k_sincos, h_k_sincos, sizeof h_k_sincos);
}

__device__ __forceinline__ void sincos_reduced(double x, double& s, double& c) {
    const double n = rint(x * k_sincos[SC_2_PI]);
    double r = fma(-n, k_sincos[SC_PIO2_1], x);
    r = fma(-n, k_sincos[SC_PIO2_2], r);
    r = fma(-n, k_sincos[SC_PIO2_3], r);
    const double z = r * r;
    double ps = k_sincos[SC_S0];
#pragma unroll
    for (int i = 1; i < 6; ++i) ps = fma(ps, z, k_sincos[SC_S0 + i]);
    const double sr = fma(r * z, ps, r);
    double pc = k_sincos[SC_C0];
#pragma unroll
    for (int i = 1; i < 6; ++i) pc = fma(pc, z, k_sincos[SC_C0 + i]);
    const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int q = (int)n;
    const double a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;
    s = (q & 2) ? -a : a;
    c = ((q + 1) & 2) ? -b : b;
}

// Same series with the coefficients as constant-bank operands (used by the both-arguments-small
// branch of the table-driven profile below).
__device__ __forceinline__ void sici_series_c(double x, double& si, double& ci_nolog) {
    const double xx = x * x;
    const double s = (xx - CHOMP_SI_SMALL_MID) * CHOMP_SI_SMALL_IHALF;
    double p = k_si_small[CHOMP_SICI_DEG_S];
    double q = k_ci_small[CHOMP_SICI_DEG_S];
#pragma unroll
    for (int i = CHOMP_SICI_DEG_S - 1; i >= 0; --i) {
        p = fma(p, s, k_si_small[i]);
        q = fma(q, s, k_ci_small[i]);
    }
    si = x * p;
    ci_nolog = fma(xx, q, CHOMP_EULER);
}

// ---------------------------------------------------------------------------------------
// Branch-free NFW profile numerator for the hot loops (tables from tools/gen_nfw_tables.py).
//   N(z, c) = g(z) - g(z2) cos(c z) + [f(z2) - 1/z2] sin(c z),   z2 = (1 + c) z
// (f, g: auxiliary functions of Si/Ci; derivation in the generator's header).  One sine/cosine
// pair, f and g at z2 and g at z, each a degree-NFW_DEG polynomial whose coefficients a lane
// fetches from shared memory by range index (nothing diverges).  The kernel is bound by the pipe
// that delivers these 16-byte fetches, hence the low degree on many narrow ranges (generator header).
// ---------------------------------------------------------------------------------------
struct NfwTables {
    double2 A[NFW_DEG + 2][NFW_NSLOT];
    double2 B[NFW_DEG + 2][NFW_NSLOT];
};
__device__ inline void nfw_tables_load(NfwTables* t) {
    const double2* a = (const double2*)&k_nfw_A[0][0][0];
    const double2* b = (const double2*)&k_nfw_B[0][0][0];
    for (int i = threadIdx.x; i < (NFW_DEG + 2) * NFW_NSLOT; i += blockDim.x) {
        (&t->A[0][0])[i] = a[i];
        (&t->B[0][0])[i] = b[i];
    }
}
// range of x >= 2 from the exponent and the top two mantissa bits of x^2: every octave of x^2 from
// [4, 8) on in NFW_SUB = 4 equal pieces, NFW_OCTAVES octaves, then one range up to infinity
__device__ __forceinline__ int nfw_range_large(double x2) {
    const int r = (__double2hiint(x2) >> 18) - (1025 << 2);
    return r < NFW_NLARGE - 1 ? r : NFW_NLARGE - 1;
}
// lnz = ln(z).  Any z > 0; z2 < 2 (both arguments small) takes the power series of Si and Ci,
// which the callers' own small-argument series make rare on the hot path.
__device__ __forceinline__ double nfw_rho_tab(const NfwTables* t, double z, double cp, double lnz) {
    const double z2 = cp * z;
    if (z2 < 2.0) {
        double s1, c1, s2, c2, si1, ci1, si2, ci2;
        sincos_reduced(z, s1, c1);
        sincos_reduced(z2, s2, c2);
        sici_series_c(z, si1, ci1);
        sici_series_c(z2, si2, ci2);
        return c1 * (log(cp) + (ci2 - ci1)) + s1 * (si2 - si1) - (s2 * c1 - c2 * s1) / z2;
    }
    const double iz2 = 1.0 / z2, iz = iz2 * cp, u2 = iz2 * iz2;
    double sc, cc;
    sincos_reduced(z2 - z, sc, cc);
    // f and g at z2
    const int ra = nfw_range_large(z2 * z2);
    const double2 ma = t->A[NFW_DEG + 1][ra];
    const double sa = (u2 - ma.x) * ma.y;
    double2 cf = t->A[NFW_DEG][ra];
    double ft = cf.x, g2 = cf.y;
#pragma unroll
    for (int j = NFW_DEG - 1; j >= 0; --j) {
        cf = t->A[j][ra];
        ft = fma(ft, sa, cf.x);
        g2 = fma(g2, sa, cf.y);
    }
    // g at z
    const bool small = z < 2.0;
    const double u1 = iz * iz;
    const int rb = small ? (int)(z * (0.5 * NFW_NSMALL)) : NFW_NSMALL + nfw_range_large(z * z);
    const double2 mb = t->B[NFW_DEG + 1][rb];
    const double sb = ((small ? z : u1) - mb.x) * mb.y;
    cf = t->B[NFW_DEG][rb];
    double p1 = cf.x, p2 = cf.y;
#pragma unroll
    for (int j = NFW_DEG - 1; j >= 0; --j) {
        cf = t->B[j][rb];
        p1 = fma(p1, sb, cf.x);
        p2 = fma(p2, sb, cf.y);
    }
    const double g1 = small ? fma(-lnz, p2, p1) : u1 * p1;
    return g1 - u2 * (g2 * cc - ft * iz2 * sc);
}

// J0 / J2 for x >= 0: piecewise degree-12 polynomials below x = 28 (tools/gen_bessel_tables.py;
// the Limber integrals stop at the 8th zero, j_{0,8} = 24.35 / j_{2,8} = 27.42), the library
// routines beyond.
__device__ __noinline__ double bessel_j_library(int order, double x) { return order == 0 ? j0(x) : jn(2, x); }
__device__ __forceinline__ double bessel_j(int order, double x) {
    if (x < BESSEL_NR * BESSEL_WIDTH) {
        const int r = (int)(x * (1.0 / BESSEL_WIDTH));
        const double s = fma(x, 2.0 / BESSEL_WIDTH, -(2.0 * r + 1.0));
        const double* __restrict__ c = &g_bessel_tab[order ? 1 : 0][0][r];
        double p = __ldg(c + BESSEL_DEG * BESSEL_NR);
#pragma unroll
        for (int j = BESSEL_DEG - 1; j >= 0; --j) p = fma(p, s, __ldg(c + j * BESSEL_NR));
        return p;
    }
    return bessel_j_library(order, x);     // out of line: one copy, away from the hot loops
}

}  // namespace chomp
