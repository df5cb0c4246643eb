// 1-halo trispectrum  T(k_i, k_j) = I^0_4(k_i, k_i, k_j, k_j)
//     = rho_bar^-3 int dln nu  nu f(nu) M^3 <moment>(M)  y(k_i, M)^2  y(k_j, M)^2
// (reference HaloTrispectrumOneHalo._initialize_i_0_4 / i_0_4 / _i_0_4_integrand,
// halo_trispectrum.py:58-140: 1275 Romberg integrals per table there).  Here it is a dense
// (k x nu)(nu x k) weighted Gram product on the finest nu node list:
//   tri_profile_kernel   A[k, node] = y(k, node)^2           (same warp-per-k scheme as halo_sums)
//   tri_gram_kernel      T = A diag(w) A^T  with FP64 tensor-core MMAs (mma.sync m8n8k4.f64)
//   tri_eval_kernel      bicubic not-a-knot interpolation (RectBivariateSpline kx=ky=3, s=0)
#pragma once
#include "common.cuh"
#include "halo_tables.cuh"
#include "special.cuh"
#include "spline.cuh"

namespace chomp {

#ifndef TRI_CHUNK
#define TRI_CHUNK 128      // points whose y^2 tables are staged at a time (360 MB at n_halo = 200)
#endif

// grid (ceil(n_k / 8), B), 256 threads: warp w handles ln k node blockIdx.x * 8 + w
__global__ void __launch_bounds__(256, 3)
tri_profile_kernel(const Cfg cfg, int b0, int B, NodesOut nd, double* __restrict__ A /* [chunk, n_k, cap_last] */) {
    __shared__ NfwTables ntab;
    nfw_tables_load(&ntab);
    __syncthreads();
    const int b = b0 + blockIdx.y;       // the A buffer holds one chunk of points starting at b0
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ik = blockIdx.x * 8 + w;
    const int nk = cfg.n_halo;
    if (b >= B || ik >= nk) return;
    const int cls = TRI_LIST;
    const int cap = nd.cap[cls];
    const int nn = nd.n_nodes[(size_t)b * N_NODE_LISTS + cls];
    const double* __restrict__ g = nd.nodes + ((size_t)b * nd.cap_total + nd.off[cls]) * NODE_FIELDS;
    const double l0 = log(cfg.k_min), l1 = log(cfg.k_max);
    const double lnk = (ik == nk - 1) ? l1 : l0 + (l1 - l0) / (nk - 1) * ik;
    const double k = exp(lnk);
    double* __restrict__ row = A + ((size_t)(b - b0) * nk + ik) * cap;
    const int nn_pad = (nn + 31) & ~31;
    for (int i = lane; i < cap; i += 32) {
        double v = 0.0;
        if (i < nn_pad) {        // whole warps enter the collective profile routine
            const int ii = i < nn ? i : nn - 1;
            const double cp = g[(size_t)NF_CP * cap + ii], rs = g[(size_t)NF_RS * cap + ii];
            const double lncp = log(cp);
            const double rho = nfw_rho_tab(&ntab, k * rs, cp, lnk + g[(size_t)NF_LNRS * cap + ii]);
            const double y = rho / (lncp - (cp - 1.0) / cp);                   // halo.py:584-585
            v = i < nn ? y * y : 0.0;
        }
        row[i] = v;
    }
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// grid (tiles, B), 256 threads.  Tile = 64 x 64 block (ti <= tj) of T; warp w owns rows
// 16 (w / 2) ... + 15 and columns 32 (w % 2) ... + 31 of it: 2 x 4 m8n8k4 accumulator fragments.
__global__ void __launch_bounds__(256)
tri_gram_kernel(const Cfg cfg, int b0, int B, int cap, const double* __restrict__ A, const double* __restrict__ wts,
                double* __restrict__ T /* [B, n_k, n_k] */) {
    const int b = b0 + blockIdx.y;
    if (b >= B) return;
    const int nk = cfg.n_halo;
    const int nt = (nk + 63) / 64;
    // unpack the upper-triangular tile index
    int t = blockIdx.x, ti = 0;
    while (t >= nt - ti) { t -= nt - ti; ++ti; }
    const int tj = ti + t;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = ti * 64 + 16 * (w >> 1), j0 = tj * 64 + 32 * (w & 1);
    const int r = lane >> 2, c = lane & 3;
    const double* __restrict__ Ab = A + (size_t)(b - b0) * nk * cap;
    const double* __restrict__ wb = wts + (size_t)b * cap;
    double acc[2][4][2];
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
    const double* arow[2];
    const double* brow[4];
    bool aok[2], bok[4];
#pragma unroll
    for (int x = 0; x < 2; ++x) { const int i = i0 + 8 * x + r; aok[x] = i < nk; arow[x] = Ab + (size_t)(aok[x] ? i : 0) * cap; }
#pragma unroll
    for (int y = 0; y < 4; ++y) { const int j = j0 + 8 * y + r; bok[y] = j < nk; brow[y] = Ab + (size_t)(bok[y] ? j : 0) * cap; }
    for (int n = 0; n < cap; n += 4) {
        const double wn = wb[n + c];
        double a[2], bb[4];
#pragma unroll
        for (int x = 0; x < 2; ++x) a[x] = aok[x] ? arow[x][n + c] : 0.0;
#pragma unroll
        for (int y = 0; y < 4; ++y) bb[y] = bok[y] ? wn * brow[y][n + c] : 0.0;
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) dmma_m8n8k4(acc[x][y][0], acc[x][y][1], a[x], bb[y]);
    }
    double* __restrict__ Tb = T + (size_t)b * nk * nk;
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = i0 + 8 * x + r, j = j0 + 8 * y + 2 * c + e;
                // upper triangle only, mirrored: the table is exactly symmetric, as the
                // reference's fill is (halo_trispectrum.py:111-121)
                if (i < nk && j < nk && i <= j) {
                    Tb[(size_t)i * nk + j] = acc[x][y][e];
                    Tb[(size_t)j * nk + i] = acc[x][y][e];
                }
            }
}

// One thread per query: tensor-product not-a-knot cubic = a 1-D spline (in ln k1) through the
// values at ln k2 of the n_k row splines.
__global__ void tri_eval_kernel(const Cfg cfg, int n, const double* __restrict__ k1, const double* __restrict__ k2,
                                const double* __restrict__ T /* [n_k, n_k] of one point */,
                                double* __restrict__ scratch /* [n, 9 n_k] */, double* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int nk = cfg.n_halo;
    double a = k1[q], bq = k2[q];
    if (a < cfg.k_min) a = cfg.k_min;                     // halo_trispectrum.py:101-102
    if (bq < cfg.k_min) bq = cfg.k_min;
    if (!(a <= cfg.k_max && bq <= cfg.k_max)) { out[q] = 0.0; return; }
    const double l0 = log(cfg.k_min), l1 = log(cfg.k_max), h = (l1 - l0) / (nk - 1);
    double* x = scratch + (size_t)q * 9 * nk;   // nk
    double* v = x + nk;                         // nk
    double* coef = v + nk;                      // 4 nk
    double* work = coef + 4 * nk;               // 2 nk  (+ nk spare)
    for (int i = 0; i < nk; ++i) x[i] = (i == nk - 1) ? l1 : l0 + h * i;
    const double x1 = log(a), x2 = log(bq);
    int j2 = uniform_index(x2, l0, 1.0 / h, nk);
    for (int i = 0; i < nk; ++i) {
        spline_build(nk, x, T + (size_t)i * nk, coef, work);
        v[i] = spline_poly(coef, j2, x2 - x[j2]);
    }
    spline_build(nk, x, v, coef, work);
    const int j1 = uniform_index(x1, l0, 1.0 / h, nk);
    out[q] = spline_poly(coef, j1, x1 - x[j1]);
}

}  // namespace chomp
