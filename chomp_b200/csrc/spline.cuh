// Not-a-knot interpolating cubic splines on the device.
//
// Every table in the reference is a scipy InterpolatedUnivariateSpline(k=3)
// (e.g. cosmology.py:795-798, mass_function.py:215-218, halo.py:919, 960,
// kernel.py:310, 645).  For k=3 FITPACK's interpolating spline is the
// not-a-knot cubic, so the device builds exactly that: a tridiagonal solve for
// the second derivatives with the third derivative continuous across x[1] and
// x[n-2], stored as per-interval polynomial coefficients
//     y(x) = c0 + t (c1 + t (c2 + t c3)),   t = x - x[i].
#pragma once
#include <cuda_runtime.h>

namespace chomp {

// Build by ONE thread.  x[n], y[n] inputs; coef[4*(n-1)] output; work[2*n]
// scratch.  n >= 4.
__device__ __noinline__ void spline_build(int n, const double* __restrict__ x, const double* __restrict__ y,
                                    double* __restrict__ coef, double* __restrict__ work) {
    double* m = work;      // second derivatives
    double* cp = work + n; // Thomas scratch (modified upper diagonal)
    // unknowns m[1..n-2]; m[0], m[n-1] eliminated through the not-a-knot rows.  The chord slopes
    // are kept in coef[4 i + 1] so that every interval costs one division in the sweep and one
    // in the coefficient pass.
    const int last = n - 2;
    const double h0 = x[1] - x[0], h1 = x[2] - x[1];
    double slope_prev = (y[1] - y[0]) / h0;
    coef[1] = slope_prev;
    double hl = h0;
    for (int i = 1; i <= last; ++i) {
        const double hr = x[i + 1] - x[i];
        const double slope = (y[i + 1] - y[i]) / hr;
        coef[4 * i + 1] = slope;
        double lo = hl, dg = 2.0 * (hl + hr), u = hr;
        if (i == 1) {
            // m[0] = (1 + h0/h1) m[1] - (h0/h1) m[2]
            dg = hl * (1.0 + hl / hr) + 2.0 * (hl + hr);
            u = hr - hl * hl / hr;
            lo = 0.0;
        }
        if (i == last) {
            // m[n-1] = (1 + hr/hl) m[n-2] - (hr/hl) m[n-3]
            dg = hr * (1.0 + hr / hl) + 2.0 * (hl + hr);
            lo = hl - hr * hr / hl;
            u = 0.0;
        }
        const double r = 6.0 * (slope - slope_prev);
        const double cprev = (i == 1) ? 0.0 : cp[i - 1], mprev = (i == 1) ? 0.0 : m[i - 1];
        const double inv = 1.0 / (dg - lo * cprev);
        cp[i] = u * inv;
        m[i] = (r - lo * mprev) * inv;
        slope_prev = slope;
        hl = hr;
    }
    for (int i = last - 1; i >= 1; --i) m[i] -= cp[i] * m[i + 1];
    m[0] = (1.0 + h0 / h1) * m[1] - (h0 / h1) * m[2];
    {
        const double hl2 = x[n - 2] - x[n - 3], hr2 = x[n - 1] - x[n - 2];
        m[n - 1] = (1.0 + hr2 / hl2) * m[n - 2] - (hr2 / hl2) * m[n - 3];
    }
    for (int i = 0; i < n - 1; ++i) {
        const double h = x[i + 1] - x[i];
        coef[4 * i + 0] = y[i];
        coef[4 * i + 1] = coef[4 * i + 1] - h * (2.0 * m[i] + m[i + 1]) * (1.0 / 6.0);
        coef[4 * i + 2] = 0.5 * m[i];
        coef[4 * i + 3] = (m[i + 1] - m[i]) / (6.0 * h);
    }
}

#define SPLINE_PCR_ROWS 2      // rows per lane in the parallel cyclic reduction (n - 2 <= 64)
// Same spline, built by one WARP (all 32 lanes must call it; work[5*n] scratch in shared memory).
// Spacings, chord slopes, the tridiagonal rows and the coefficient pass are lane-parallel; only
// the two sweeps of the Thomas algorithm -- a serial chain of one reciprocal per row -- run on
// lane 0, with everything they need already in shared memory.
__device__ __noinline__ void spline_build_warp(int n, const double* __restrict__ x, const double* __restrict__ y,
                                               double* __restrict__ coef, double* __restrict__ work) {
    const int lane = threadIdx.x & 31;
    double* m = work;          // second derivatives (right-hand side first)
    double* cp = work + n;     // modified upper diagonal
    double* lo = work + 2 * n; // sub-diagonal
    double* dg = work + 3 * n; // diagonal
    double* up = work + 4 * n; // super-diagonal
    const int last = n - 2;
    for (int i = lane; i < n - 1; i += 32) coef[4 * i + 1] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    __syncwarp();
    for (int i = 1 + lane; i <= last; i += 32) {
        const double hl = x[i] - x[i - 1], hr = x[i + 1] - x[i];
        double l = hl, d = 2.0 * (hl + hr), u = hr;
        if (i == 1) {               // m[0] = (1 + h0/h1) m[1] - (h0/h1) m[2]
            d = hl * (1.0 + hl / hr) + 2.0 * (hl + hr);
            u = hr - hl * hl / hr;
            l = 0.0;
        }
        if (i == last) {            // m[n-1] = (1 + hr/hl) m[n-2] - (hr/hl) m[n-3]
            d = hr * (1.0 + hr / hl) + 2.0 * (hl + hr);
            l = hl - hr * hr / hl;
            u = 0.0;
        }
        lo[i] = l; dg[i] = d; up[i] = u;
        m[i] = 6.0 * (coef[4 * i + 1] - coef[4 * (i - 1) + 1]);
    }
    __syncwarp();
    if (last <= 32 * SPLINE_PCR_ROWS) {
        // parallel cyclic reduction over the rows 1 .. last (diagonally dominant: stable): log2(n)
        // steps, every lane eliminating the neighbours at distance s of its own rows -- instead of
        // a serial chain of one division per row
        for (int s = 1; s < last; s <<= 1) {
            double na[SPLINE_PCR_ROWS], nb[SPLINE_PCR_ROWS], nc[SPLINE_PCR_ROWS], nd[SPLINE_PCR_ROWS];
#pragma unroll
            for (int t = 0; t < SPLINE_PCR_ROWS; ++t) {
                const int i = 1 + lane + 32 * t;
                na[t] = nb[t] = nc[t] = nd[t] = 0.0;
                if (i <= last) {
                    double bb = dg[i], d = m[i], aa = 0.0, cc = 0.0;
                    const int im = i - s, ip = i + s;
                    if (im >= 1) { const double al = -lo[i] / dg[im]; aa = al * lo[im]; bb += al * up[im]; d += al * m[im]; }
                    if (ip <= last) { const double ga = -up[i] / dg[ip]; cc = ga * up[ip]; bb += ga * lo[ip]; d += ga * m[ip]; }
                    na[t] = aa; nb[t] = bb; nc[t] = cc; nd[t] = d;
                }
            }
            __syncwarp();
#pragma unroll
            for (int t = 0; t < SPLINE_PCR_ROWS; ++t) {
                const int i = 1 + lane + 32 * t;
                if (i <= last) { lo[i] = na[t]; dg[i] = nb[t]; up[i] = nc[t]; m[i] = nd[t]; }
            }
            __syncwarp();
        }
        for (int i = 1 + lane; i <= last; i += 32) m[i] = m[i] / dg[i];
        __syncwarp();
        if (lane == 0) {
            const double h0 = x[1] - x[0], h1 = x[2] - x[1];
            m[0] = (1.0 + h0 / h1) * m[1] - (h0 / h1) * m[2];
            const double hl2 = x[n - 2] - x[n - 3], hr2 = x[n - 1] - x[n - 2];
            m[n - 1] = (1.0 + hr2 / hl2) * m[n - 2] - (hr2 / hl2) * m[n - 3];
        }
    } else if (lane == 0) {
        double cprev = 0.0, mprev = 0.0;
        for (int i = 1; i <= last; ++i) {
            const double l = lo[i];
            const double inv = 1.0 / (dg[i] - l * cprev);
            cprev = up[i] * inv;
            mprev = (m[i] - l * mprev) * inv;
            cp[i] = cprev;
            m[i] = mprev;
        }
        for (int i = last - 1; i >= 1; --i) { mprev = m[i] - cp[i] * mprev; m[i] = mprev; }
        const double h0 = x[1] - x[0], h1 = x[2] - x[1];
        m[0] = (1.0 + h0 / h1) * m[1] - (h0 / h1) * m[2];
        const double hl2 = x[n - 2] - x[n - 3], hr2 = x[n - 1] - x[n - 2];
        m[n - 1] = (1.0 + hr2 / hl2) * m[n - 2] - (hr2 / hl2) * m[n - 3];
    }
    __syncwarp();
    for (int i = lane; i < n - 1; i += 32) {
        const double h = x[i + 1] - x[i];
        coef[4 * i + 0] = y[i];
        coef[4 * i + 1] = coef[4 * i + 1] - h * (2.0 * m[i] + m[i + 1]) * (1.0 / 6.0);
        coef[4 * i + 2] = 0.5 * m[i];
        coef[4 * i + 3] = (m[i + 1] - m[i]) / (6.0 * h);
    }
    __syncwarp();
}

// Uniform abscissae x_i = x0 + i h: the not-a-knot system reduces to  m_1 = r_1 / 6,
// m_{n-2} = r_{n-2} / 6,  m_{i-1} + 4 m_i + m_{i+1} = r_i  in between (r_i = 6 (y_{i+1} - 2 y_i +
// y_{i-1}) / h^2), whose Thomas factors c_i = 1 / (4 - c_{i-1}) depend on nothing: they come from a
// constant table (c_i is 2 - sqrt(3) to rounding from i = 24 on), so the serial sweeps are two
// dependent operations per row and free of divisions.  Warp-collective; work[n] scratch.
#define NAK_CP_TABLE 24
__constant__ double k_nak_cp[NAK_CP_TABLE];
static inline cudaError_t chomp_upload_spline_tables() {
    double cp[NAK_CP_TABLE];
    cp[0] = 0.0; cp[1] = 0.0;
    for (int i = 2; i < NAK_CP_TABLE; ++i) cp[i] = 1.0 / (4.0 - cp[i - 1]);
    return cudaMemcpyToSymbol(k_nak_cp, cp, sizeof cp);
}
__device__ __noinline__ void spline_build_uniform_warp(int n, double h, const double* __restrict__ y,
                                                       double* __restrict__ coef, double* __restrict__ work) {
    const int lane = threadIdx.x & 31;
    double* m = work;
    const double s = 6.0 / (h * h);
    for (int i = 1 + lane; i <= n - 2; i += 32) m[i] = s * ((y[i + 1] - y[i]) - (y[i] - y[i - 1]));
    __syncwarp();
    if (lane == 0) {
        double mprev = m[1] * (1.0 / 6.0);
        m[1] = mprev;
        for (int i = 2; i <= n - 3; ++i) {
            const double c = i < NAK_CP_TABLE ? k_nak_cp[i] : 0.26794919243112270647;
            mprev = (m[i] - mprev) * c;
            m[i] = mprev;
        }
        double mnext = m[n - 2] * (1.0 / 6.0);
        m[n - 2] = mnext;
        for (int i = n - 3; i >= 2; --i) {
            const double c = i < NAK_CP_TABLE ? k_nak_cp[i] : 0.26794919243112270647;
            mnext = m[i] - c * mnext;
            m[i] = mnext;
        }
        m[0] = 2.0 * m[1] - m[2];
        m[n - 1] = 2.0 * m[n - 2] - m[n - 3];
    }
    __syncwarp();
    const double ih = 1.0 / h;
    for (int i = lane; i < n - 1; i += 32) {
        coef[4 * i + 0] = y[i];
        coef[4 * i + 1] = (y[i + 1] - y[i]) * ih - h * (2.0 * m[i] + m[i + 1]) * (1.0 / 6.0);
        coef[4 * i + 2] = 0.5 * m[i];
        coef[4 * i + 3] = (m[i + 1] - m[i]) * (ih * (1.0 / 6.0));
    }
    __syncwarp();
}

__device__ __forceinline__ double spline_poly(const double* __restrict__ coef, int i, double t) {
    const double* c = coef + 4 * i;
    return fma(t, fma(t, fma(t, c[3], c[2]), c[1]), c[0]);
}

// index of the interval containing v on a uniform grid x0 + i*h (clamped)
__device__ __forceinline__ int uniform_index(double v, double x0, double inv_h, int n) {
    int i = (int)floor((v - x0) * inv_h);
    return i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
}

// index of the interval containing v on an increasing grid (clamped)
__device__ __forceinline__ int search_index(double v, const double* __restrict__ x, int n) {
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ double spline_eval_uniform(const double* __restrict__ coef, double v, double x0,
                                                      double h, int n) {
    int i = uniform_index(v, x0, 1.0 / h, n);
    // guard against floor() landing one interval off at a knot
    const double xi = x0 + i * h;
    return spline_poly(coef, i, v - xi);
}

__device__ __forceinline__ double spline_eval_search(const double* __restrict__ coef, double v,
                                                     const double* __restrict__ x, int n) {
    const int i = search_index(v, x, n);
    return spline_poly(coef, i, v - x[i]);
}

}  // namespace chomp
