// Two callers next to the hot path:
//
//   halo_i12_kernel / halo_ssc_eval_kernel   HaloSuperSampleCovariance (halo.py:1089-1199): the table
//       I^1_2(k) = rho_bar^-1 int dln nu  nu f(nu) b(nu) y(k, M)^2 M        (halo.py:1174-1199)
//       and  dln P / d delta_b = (68/21 h_m^2 P_lin + I^1_2) / P_mm         (halo.py:1138-1157, Takada & Hu 2013)
//   xi3d_kernel                              Correlation3d.raw_correlation (correlation.py:467-500):
//       xi(r) = int dln k  k^2 / (2 pi)  P(k)  J0(k r)   -- the reference's integrand as written (cylindrical J0)
#pragma once
#include "common.cuh"
#include "halo_tables.cuh"
#include "hankel.cuh"
#include "special.cuh"
#include "spline.cuh"

namespace chomp {

// grid (B), 256 threads: one warp per ln k node on the finest node list of the halo-tables stage (its per-panel orders
// follow k_max r_vir, so it serves every k); then one warp splines the table (not-a-knot, uniform ln k grid).
// dynamic shared memory: 2 n_halo doubles
__global__ void __launch_bounds__(256)
halo_i12_kernel(const Cfg cfg, int B, NodesOut nd, const double* __restrict__ epoch, double* __restrict__ tab /* [B, n_halo] */,
                double* __restrict__ coef /* [B, 4 n_halo] */, int32_t* __restrict__ status) {
    extern __shared__ double sm[];
    __shared__ NfwTables ntab;
    const int b = blockIdx.x;
    if (b >= B) return;
    nfw_tables_load(&ntab);
    __syncthreads();
    const int nk = cfg.n_halo, w = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int cls = N_KCLASS - 1;
    const int cap = nd.cap[cls];
    const int nn = nd.n_nodes[(size_t)b * N_NODE_LISTS + cls];
    const double* __restrict__ g = nd.nodes + ((size_t)b * nd.cap_total + nd.off[cls]) * NODE_FIELDS;
    const double* e = epoch + (size_t)b * CHOMP_EPOCH_LEN;
    const double m_coef = 4.0 * M_PI * e[EP_DELTA_V] / 3.0;           // M / rho_bar = (4 pi Delta_v / 3) r_vir^3  (halo.py:890-893)
    const double l0 = log(cfg.k_min), l1 = log(cfg.k_max), hk = (l1 - l0) / (nk - 1);
    double* y = sm;                     // n_halo
    double* work = sm + nk;             // n_halo
    for (int ik = w; ik < nk; ik += nwarp) {
        const double lnk = (ik == nk - 1) ? l1 : l0 + hk * ik;
        const double k = exp(lnk);
        double acc = 0.0;
        for (int i = lane; i < ((nn + 31) & ~31); i += 32) {
            const int ii = i < nn ? i : nn - 1;
            const double cp = g[(size_t)NF_CP * cap + ii], rs = g[(size_t)NF_RS * cap + ii];
            const double rho = nfw_rho_tab(&ntab, k * rs, cp, lnk + g[(size_t)NF_LNRS * cap + ii]);
            const double c = cp - 1.0, rv = rs * c;
            const double imk = 1.0 / (log(cp) - c / cp);                  // halo.py:584
            // W_HM = dln nu  nu f(nu) b(nu) / m(c): times rho^2 / m(c) = y^2 m(c) ... and M / rho_bar
            if (i < nn) acc = fma(g[(size_t)NF_W_HM * cap + ii] * imk * (m_coef * rv * rv * rv), rho * rho, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) y[ik] = acc;
    }
    __syncthreads();
    bool bad = false;
    for (int i = threadIdx.x; i < nk; i += blockDim.x) { tab[(size_t)b * nk + i] = y[i]; if (!isfinite(y[i])) bad = true; }
    if (bad && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
    if (w == 0) spline_build_uniform_warp(nk, hk, y, coef + (size_t)b * 4 * nk, work);
}

// what = 0: I^1_2(k);  1: dln P / d delta_b.  Both vanish outside [k_min, k_max] (halo.py:1153-1157, 1169-1172).
__global__ void __launch_bounds__(256)
halo_ssc_eval_kernel(const Cfg cfg, int B, int what, int n_k, const double* __restrict__ k_in, const double* __restrict__ cosmo,
                     const double* __restrict__ epoch, const double* __restrict__ htab, const double* __restrict__ hcoef,
                     const double* __restrict__ hfit, const double* __restrict__ i12_coef, double* __restrict__ out) {
    const int per = (n_k + blockDim.x - 1) / blockDim.x;
    const int b = blockIdx.x / per;
    const int i = (blockIdx.x - b * per) * blockDim.x + threadIdx.x;
    if (b >= B || i >= n_k) return;
    const double k = k_in[i];
    double r = 0.0;
    if (k >= cfg.k_min && k <= cfg.k_max) {
        const int nk = cfg.n_halo;
        const double l0 = log(cfg.k_min), h = (log(cfg.k_max) - l0) / (nk - 1), x = log(k);
        const int j = uniform_index(x, l0, 1.0 / h, nk);
        const double dx = x - (l0 + h * j);
        const double i12 = spline_poly(i12_coef + (size_t)b * 4 * nk, j, dx);
        if (what == 0) r = i12;
        else {
            const Cosmo c = load_cosmo(cosmo + (size_t)b * CHOMP_N_COSMO, cfg.cosmo_precision);
            const double* e = epoch + (size_t)b * CHOMP_EPOCH_LEN;
            PkParams pk = make_pk(c, e[EP_GROWTH], e[EP_SIGMA_NORM]);
            CHOMP_ATTACH_BAO(cfg, c, pk)
            HaloTabs T;
            T.nk = nk; T.l0 = l0; T.l1 = log(cfg.k_max); T.h = h; T.k_min = cfg.k_min; T.k_max = cfg.k_max;
            T.extrapolate = cfg.extrapolate; T.tab = htab + (size_t)b * 5 * nk; T.coef = hcoef + (size_t)b * 20 * nk;
            T.hf = hfit ? hfit + (size_t)b * HF_LEN : nullptr;
            const double hm = tab_at(T, 0, j, dx);
            r = (68.0 / 21.0 * hm * hm * linear_power(pk, k) + i12) / halo_power(T, pk, CHOMP_P_MM, k);
        }
    }
    out[(size_t)b * n_k + i] = r;
}

// xi(r): grid (n_r, B), 256 threads.  Every interval of the halo tables' ln k grid is cut into pieces no wider than
// HANKEL_MAX_PIECE over which k r advances by at most XI_PHASE; Gauss-Legendre nq_hankel per piece.
#define XI_PHASE 2.0
#define XI_MAX_INTERVALS 1024
__global__ void __launch_bounds__(256)
xi3d_kernel(const Cfg cfg, int B, int which, int n_r, const double* __restrict__ r_in, const double* __restrict__ cosmo,
            const double* __restrict__ epoch, const double* __restrict__ htab, const double* __restrict__ hcoef,
            const double* __restrict__ hfit, double* __restrict__ out, int32_t* __restrict__ status) {
    __shared__ int pfx[XI_MAX_INTERVALS + 1];
    __shared__ double red[32];
    const int b = blockIdx.y, ir = blockIdx.x;
    if (b >= B || ir >= n_r) return;
    const int tid = threadIdx.x;
    const int nk = cfg.n_halo, nq = cfg.nq_hankel;
    const double r = r_in[ir];
    const double l0 = log(cfg.k_min), l1 = log(cfg.k_max), hP = (l1 - l0) / (nk - 1);
    for (int i = tid; i < nk - 1; i += blockDim.x) {
        const double ka = exp(l0 + hP * i), kb = exp((i == nk - 2) ? l1 : l0 + hP * (i + 1));
        int cnt = (int)ceil(fmax(hP / HANKEL_MAX_PIECE - 1e-9, fmin(r * (kb - ka) / XI_PHASE, 1.0e6)));
        pfx[i + 1] = cnt < 1 ? 1 : cnt;
    }
    __syncthreads();
    if (tid == 0) {
        pfx[0] = 0;
        for (int i = 0; i < nk - 1; ++i) pfx[i + 1] += pfx[i];
    }
    __syncthreads();
    const Cosmo c = load_cosmo(cosmo + (size_t)b * CHOMP_N_COSMO, cfg.cosmo_precision);
    const double* e = epoch + (size_t)b * CHOMP_EPOCH_LEN;
    PkParams pk = make_pk(c, e[EP_GROWTH], e[EP_SIGMA_NORM]);
    CHOMP_ATTACH_BAO(cfg, c, pk)
    HaloTabs T;
    T.nk = nk; T.l0 = l0; T.l1 = l1; T.h = hP; T.k_min = cfg.k_min; T.k_max = cfg.k_max; T.extrapolate = cfg.extrapolate;
    T.tab = htab + (size_t)b * 5 * nk; T.coef = hcoef + (size_t)b * 20 * nk;
    T.hf = hfit ? hfit + (size_t)b * HF_LEN : nullptr;
    const long long total = (long long)pfx[nk - 1] * nq;
    double acc = 0.0;
    for (long long idx = tid; idx < total; idx += blockDim.x) {
        const int piece = (int)(idx / nq), q = (int)(idx - (long long)piece * nq);
        int lo = 0, hi = nk - 1;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pfx[mid] <= piece) lo = mid; else hi = mid; }
        const int n_sub = pfx[lo + 1] - pfx[lo], s = piece - pfx[lo];
        const double a = l0 + hP * lo, bb = (lo == nk - 2) ? l1 : l0 + hP * (lo + 1);
        const double pa = a + (bb - a) * s / n_sub, pb = (s == n_sub - 1) ? bb : a + (bb - a) * (s + 1) / n_sub;
        const double half = 0.5 * (pb - pa);
        const double x = 0.5 * (pa + pb) + half * c_glx[nq][q];
        const double k = exp(x);
        acc += half * c_glw[nq][q] * k * k * halo_power(T, pk, which, k) * j0(k * r);       // correlation.py:493-500
    }
    acc = block_sum(acc, red);
    if (tid == 0) {
        const double v = acc / (2.0 * M_PI);
        out[(size_t)b * n_r + ir] = v;
        if (!isfinite(v) && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
    }
}

}  // namespace chomp
