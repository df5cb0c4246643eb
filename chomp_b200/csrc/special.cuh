// Device special functions for the halo-model hot path (FP64).
//
//   sici        sine / cosine integrals  -- replaces scipy.special.sici in the
//               NFW Fourier profile (reference halo.py:578-579) and in the
//               halo-exclusion window (halo.py:1232)
//   bessel_j    J0 / J2                  -- replaces scipy.special.j0 / jn(2, .)
//               (reference kernel.py:712, 839)
//
// Coefficients come from tools/gen_special.py (Chebyshev interpolants built with
// mpmath, max relative error <= 4e-14 on f, g and <= 3e-16 on the series part).
#pragma once
#include <math.h>
#include "special_coeffs.cuh"

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
// register / instruction footprint in the hot loop) is not needed.
__device__ __forceinline__ void sincos_reduced(double x, double& s, double& c) {
    const double n = rint(x * 0.63661977236758134308);
    double r = fma(-n, 1.5707963267948966, x);
    r = fma(-n, 6.123233995736766e-17, r);
    r = fma(-n, -1.4973849048591698e-33, r);
    const double z = r * r;
    double ps = 1.58969099521155010221e-10;
    ps = fma(ps, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double sr = fma(r * z, ps, r);
    double pc = -1.13596475577881948265e-11;
    pc = fma(pc, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int q = (int)n;
    const double a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;
    s = (q & 2) ? -a : a;
    c = ((q + 1) & 2) ? -b : b;
}

// x <= CHOMP_SICI_TINY_X: same series, degree CHOMP_SICI_DEG_T
__device__ __forceinline__ void sici_series_tiny_c(double x, double& si, double& ci_nolog) {
    const double xx = x * x;
    const double s = (xx - CHOMP_SI_TINY_MID) * CHOMP_SI_TINY_IHALF;
    double p = k_si_tiny[CHOMP_SICI_DEG_T];
    double q = k_ci_tiny[CHOMP_SICI_DEG_T];
#pragma unroll
    for (int i = CHOMP_SICI_DEG_T - 1; i >= 0; --i) {
        p = fma(p, s, k_si_tiny[i]);
        q = fma(q, s, k_ci_tiny[i]);
    }
    si = x * p;
    ci_nolog = fma(xx, q, CHOMP_EULER);
}

// ---------------------------------------------------------------------------------------
// Warp-uniform fast path.  The nu nodes are ordered by mass, so the 32 lanes of a warp
// almost always fall into the same Si/Ci ranges; the polynomials are then evaluated with
// the coefficients as constant-bank operands of the DFMA instructions (no loads at all).
// ---------------------------------------------------------------------------------------
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

template <int R>
__device__ __forceinline__ void sici_aux_c(double x, double sx, double cx, double& si, double& ci) {
    const double ix = 1.0 / x;
    const double u = ix * ix;
    const double s = (u - k_sici_urange[R][0]) * k_sici_urange[R][1];
    double f = k_sici_F[R][CHOMP_SICI_DEG_L];
    double g = k_sici_G[R][CHOMP_SICI_DEG_L];
#pragma unroll
    for (int i = CHOMP_SICI_DEG_L - 1; i >= 0; --i) {
        f = fma(f, s, k_sici_F[R][i]);
        g = fma(g, s, k_sici_G[R][i]);
    }
    f *= ix;
    g *= u;
    si = CHOMP_PI_2 - f * cx - g * sx;
    ci = f * sx - g * cx;
}

__device__ __forceinline__ int sici_range(double x) {
    return (x >= CHOMP_SICI_X2) ? 2 : ((x >= CHOMP_SICI_X1) ? 1 : 0);
}

template <int R1, int R2>
__device__ __forceinline__ void nfw_large_large(double z, double z2, double s1, double c1, double s2, double c2,
                                                double& dsi, double& dci) {
    double si1, ci1, si2, ci2;
    sici_aux_c<R1>(z, s1, c1, si1, ci1);
    sici_aux_c<R2>(z2, s2, c2, si2, ci2);
    dsi = si2 - si1;
    dci = ci2 - ci1;
}
template <int R2>
__device__ __forceinline__ void nfw_small_large(double z, double z2, double s2, double c2, double& dsi, double& dci) {
    double si1, ci1, si2, ci2;
    sici_series_c(z, si1, ci1);
    sici_aux_c<R2>(z2, s2, c2, si2, ci2);
    dsi = si2 - si1;
    dci = ci2 - (ci1 + log(z));
}

// Same value as nfw_rho_k; all 32 lanes of the warp must call it together.
__device__ __forceinline__ double nfw_rho_k_warp(const SiciTables* t, double z, double cp, double lncp) {
    const double z2 = cp * z;
    double s1, c1, s2, c2;
    sincos_reduced(z, s1, c1);
    sincos_reduced(z2, s2, c2);
    const double sin_cz = s2 * c1 - c2 * s1;
    const bool small1 = z <= CHOMP_SICI_SMALL_X, small2 = z2 <= CHOMP_SICI_SMALL_X;
    const int key = (z2 <= CHOMP_SICI_TINY_X) ? 13
                    : (small2 ? 0 : (small1 ? 1 + sici_range(z2) : 4 + 3 * sici_range(z) + sici_range(z2)));
    const int key0 = __shfl_sync(0xffffffffu, key, 0);
    double dsi, dci;
    if (__all_sync(0xffffffffu, key == key0)) {
        switch (key0) {
            case 0: {
                double si1, ci1, si2, ci2;
                sici_series_c(z, si1, ci1);
                sici_series_c(z2, si2, ci2);
                dsi = si2 - si1;
                dci = lncp + (ci2 - ci1);
            } break;
            case 13: {
                double si1, ci1, si2, ci2;
                sici_series_tiny_c(z, si1, ci1);
                sici_series_tiny_c(z2, si2, ci2);
                dsi = si2 - si1;
                dci = lncp + (ci2 - ci1);
            } break;
            case 1: nfw_small_large<0>(z, z2, s2, c2, dsi, dci); break;
            case 2: nfw_small_large<1>(z, z2, s2, c2, dsi, dci); break;
            case 3: nfw_small_large<2>(z, z2, s2, c2, dsi, dci); break;
            case 4: nfw_large_large<0, 0>(z, z2, s1, c1, s2, c2, dsi, dci); break;
            case 5: nfw_large_large<0, 1>(z, z2, s1, c1, s2, c2, dsi, dci); break;
            case 6: nfw_large_large<0, 2>(z, z2, s1, c1, s2, c2, dsi, dci); break;
            case 8: nfw_large_large<1, 1>(z, z2, s1, c1, s2, c2, dsi, dci); break;
            case 9: nfw_large_large<1, 2>(z, z2, s1, c1, s2, c2, dsi, dci); break;
            default: nfw_large_large<2, 2>(z, z2, s1, c1, s2, c2, dsi, dci); break;   // 12
        }
    } else {
        double si1, ci1, si2, ci2;
        if (small2) {
            sici_series(t, z, si1, ci1);
            sici_series(t, z2, si2, ci2);
            dsi = si2 - si1;
            dci = lncp + (ci2 - ci1);
        } else if (small1) {
            sici_series(t, z, si1, ci1);
            sici_aux(t, z2, s2, c2, si2, ci2);
            dsi = si2 - si1;
            dci = ci2 - (ci1 + log(z));
        } else {
            sici_aux(t, z, s1, c1, si1, ci1);
            sici_aux(t, z2, s2, c2, si2, ci2);
            dsi = si2 - si1;
            dci = ci2 - ci1;
        }
    }
    return c1 * dci + s1 * dsi - sin_cz / z2;
}

__device__ __forceinline__ double bessel_j(int order, double x) {
    return order == 0 ? j0(x) : jn(2, x);
}

}  // namespace chomp
