// Covariance between two DIFFERENT correlations (reference covariance.Covariance with
// input_correlation_a != input_correlation_b, covariance.py:60-63: ``matching_corrs`` False; the blocks
// CovarianceMulti assembles, covariance.py:794-871).  Correlation a lives on one handle (windows a1, a2, halo a at
// z_bar_a), correlation b on another (b1, b2, halo b at z_bar_b); both were built for the same cosmology rows.
//
//   cov_kng_cross_kernel        K_NG with all four windows: int dchi W_a1 W_a2 W_b1 W_b2 D^4 / chi^2 J0 J0
//                               (kernel.py:1102-1111), z_bar_NG on the common z range (kernel.py:909-915, 961-972)
//   cov_projected_cross_kernel  the four projected spectra P_a, P_b, P_ab, P_ba (covariance.py:455-591): P_ab / P_ba
//                               project sqrt(P_a P_b) with the windows (a1, b2) / (a2, b1)
//   cov_g_cross_kernel          Gaussian term with both two-point products (covariance.py:390-453)
// The non-Gaussian term and the assembly are the matching case's kernels (covariance.cuh); the Poisson term
// vanishes (covariance.py:313-315 adds it for matching correlations only).
#pragma once
#include "covariance.cuh"

namespace chomp {

struct Limber4 {
    EpochGrid g;
    Window w[4];           // a1, a2, b1, b2
    __device__ __forceinline__ double growth(double chi) const { return grid_growth(g, grid_z(g, chi)); }
    // W_i W_j D^2 (kernel.py:1070-1100)
    __device__ __forceinline__ double pair(int i, int j, double chi) const {
        const double D = growth(chi);
        return window_eval(w[i], chi) * window_eval(w[j], chi) * D * D;
    }
};
__host__ __device__ inline size_t limber_stage4_doubles(const Cfg& cfg) { return 13 * (size_t)cfg.n_cosmo + 16 * (size_t)cfg.n_window; }

// grid tables of handle a (the two handles share the cosmology) and the window splines of both, in shared memory
__device__ inline Limber4 limber_stage4(const Cfg& cfg, const LimberIn& ia, const LimberIn& ib, int b, double* sm) {
    const int nz = cfg.n_cosmo, nw = cfg.n_window;
    for (int i = threadIdx.x; i < 13 * nz; i += blockDim.x) sm[i] = ia.grid0[(size_t)b * 13 * nz + i];
    double* wc = sm + 13 * nz;
    for (int i = threadIdx.x; i < 8 * nw; i += blockDim.x) {
        wc[i] = ia.win_coef[(size_t)b * 8 * nw + i];
        wc[8 * nw + i] = ib.win_coef[(size_t)b * 8 * nw + i];
    }
    Limber4 F;
    F.g.n = nz; F.g.z_min = cfg.zk_min < 0.0 ? 0.0 : cfg.zk_min; F.g.z_max = cfg.zk_max;
    F.g.chi = sm; F.g.c_chi_z = sm + nz; F.g.c_z_chi = sm + 5 * nz; F.g.c_g_z = sm + 9 * nz;
    F.g.z = nullptr; F.g.growth = nullptr;
    const double* ca = ia.win_chi + (size_t)b * 4;
    const double* cb = ib.win_chi + (size_t)b * 4;
    F.w[0] = Window{nw, ca[0], ca[1], nullptr, wc};
    F.w[1] = Window{nw, ca[2], ca[3], nullptr, wc + 4 * nw};
    F.w[2] = Window{nw, cb[0], cb[1], nullptr, wc + 8 * nw};
    F.w[3] = Window{nw, cb[2], cb[3], nullptr, wc + 12 * nw};
    return F;
}

// Panel edges of both handles merged (each list is sorted), clipped to [lo, hi], duplicates dropped.  One thread.
__device__ inline int merge_edges(const double* ea, int na, const double* eb, int nb, double lo, double hi, double* out, int cap) {
    int i = 0, j = 0, n = 0;
    out[n++] = lo;
    while ((i < na || j < nb) && n < cap - 1) {
        double v;
        if (j >= nb || (i < na && ea[i] <= eb[j])) v = ea[i++]; else v = eb[j++];
        if (v <= lo || v >= hi) continue;
        if (v - out[n - 1] <= 1e-12 * fabs(v)) continue;
        out[n++] = v;
    }
    if (hi - out[n - 1] <= 1e-12 * fabs(hi)) out[n - 1] = hi; else out[n++] = hi;
    return n;
}

// common z range of the four windows and the chi range that goes with it (kernel.py:909-930)
__device__ inline void cross_chi_range(const Cfg& ca, const Cfg& cb, const EpochGrid& g, double& zlo, double& zhi, double& chi_lo,
                                       double& chi_hi) {
    double a0, a1, b0, b1;
    window_z_range(ca, a0, a1);
    window_z_range(cb, b0, b1);
    zlo = fmax(a0, b0); zhi = fmin(a1, b1);
    chi_lo = fmax(ca.window_precision, grid_chi(g, zlo));
    chi_hi = grid_chi(g, zhi);
}

struct KngS4 {
    Limber4 F;
    __device__ __forceinline__ double operator()(double chi) const {
        const double D = F.growth(chi), D2 = D * D;
        return window_eval(F.w[0], chi) * window_eval(F.w[1], chi) * window_eval(F.w[2], chi) * window_eval(F.w[3], chi) * D2 * D2 /
               (chi * chi);
    }
};

// grid (n_kernel, B): as cov_kng_kernel
__global__ void __launch_bounds__(COV_THREADS)
cov_kng_cross_kernel(const Cfg cfg, const Cfg cfg_b, const CovP cp, int B, LimberIn ia, LimberIn ib, CovOut out) {
    extern __shared__ double dyn[];
    __shared__ OscShared s;
    __shared__ int n_edge_s;
    const int b = blockIdx.y;
    if (b >= B) return;
    const int nk = cfg.n_kernel;
    const int i = nk - 1 - blockIdx.x;
    const int tid = threadIdx.x;
    KngS4 S{limber_stage4(cfg, ia, ib, b, dyn)};
    __syncthreads();
    double zlo, zhi, chi_min, chi_max;
    cross_chi_range(cfg, cfg_b, S.F.g, zlo, zhi, chi_min, chi_max);
    if (tid == 0)
        n_edge_s = (chi_max > chi_min)
                       ? merge_edges(ia.edges + (size_t)b * ia.edge_stride, ia.n_edges[b], ib.edges + (size_t)b * ib.edge_stride,
                                     ib.n_edges[b], chi_min, chi_max, s.edge, COV_MAX_EDGES)
                       : 0;
    const double x0 = log(cp.theta_min_rad * cfg.k_min), x1 = log(cp.theta_max_rad * cfg.k_max);
    const double hx = (x1 - x0) / (nk - 1);
    for (int j = tid; j <= i; j += blockDim.x) {
        const double kt = exp((j == nk - 1) ? x1 : x0 + hx * j);
        s.fj[j] = kt;
        s.top[j] = fmin(cp.bessel_limit / kt, chi_max);       // kernel.py:1047-1052
    }
    __syncthreads();
    const int n_edge = n_edge_s;
    if (blockIdx.x == 0 && tid < 32) {
        // z_bar_NG: arg-max of the weight on the n_kernel-point z grid (kernel.py:961-972)
        double best = -1e300; int besti = 0;
        for (int j = tid; j < nk; j += 32) {
            const double zj = (j == nk - 1) ? zhi : zlo + (zhi - zlo) / (nk - 1) * j;
            double chi = grid_chi(S.F.g, zj);
            if (!(chi > cfg.window_precision)) chi = cfg.window_precision;
            const double v = S(chi);
            if (v > best) { best = v; besti = j; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        if (tid == 0) {
            const double zb = (besti == nk - 1) ? zhi : zlo + (zhi - zlo) / (nk - 1) * besti;
            out.zbar_ng[b] = zb;
            out.d_ng[b] = grid_growth(S.F.g, zb);             // covariance.py:142
        }
    }
    double* K = out.kng + (size_t)b * nk * nk;
    if (n_edge < 2) {                                         // the four windows do not overlap: K_NG = 0 (kernel.py:1051-1052)
        for (int j = tid; j <= i; j += blockDim.x) { K[(size_t)i * nk + j] = 0.0; K[(size_t)j * nk + i] = 0.0; }
        return;
    }
    const double kt_i = exp((i == nk - 1) ? x1 : x0 + hx * i);
    osc_row(S, IdentityU(), s, n_edge, kt_i, i + 1, cp.nq_osc, cp.osc_phase);
    for (int j = tid; j <= i; j += blockDim.x) {
        const double v = (s.flag[j] & 1) ? 0.0 : s.acc[j];
        K[(size_t)i * nk + j] = v;
        K[(size_t)j * nk + i] = v;
    }
}

// halo tables + linear spectrum of one handle for point b
struct HaloSide {
    const double *cosmo, *epoch, *htab, *hcoef, *hfit;
    int extrapolate;
};
__device__ inline HaloTabs halo_side_tabs(const Cfg& cfg, const HaloSide& hs, int b, PkParams& pk, double cosmo_precision) {
    const Cosmo c = load_cosmo(hs.cosmo + (size_t)b * CHOMP_N_COSMO, cosmo_precision);
    const double* e = hs.epoch + (size_t)b * CHOMP_EPOCH_LEN;
    pk = make_pk(c, e[EP_GROWTH], e[EP_SIGMA_NORM]);
    HaloTabs T;
    T.nk = cfg.n_halo; T.l0 = log(cfg.k_min); T.l1 = log(cfg.k_max); T.h = (T.l1 - T.l0) / (T.nk - 1);
    T.k_min = cfg.k_min; T.k_max = cfg.k_max; T.extrapolate = hs.extrapolate;
    T.tab = hs.htab + (size_t)b * 5 * T.nk; T.coef = hs.hcoef + (size_t)b * 20 * T.nk;
    T.hf = hs.hfit ? hs.hfit + (size_t)b * HF_LEN : nullptr;
    return T;
}

// chi ranges of the two correlations (covariance.py:125-140) and the common ln K grid (:160-172)
struct CrossRanges {
    double chi_lo_a, chi_hi_a, chi_lo_b, chi_hi_b, lK0, lK1;
};
__device__ inline CrossRanges cross_ranges(const Cfg& cfg, const LimberIn& ia, const LimberIn& ib, int b) {
    CrossRanges r;
    r.chi_lo_a = ia.kchi[2 * b]; r.chi_hi_a = ia.kchi[2 * b + 1];
    r.chi_lo_b = ib.kchi[2 * b]; r.chi_hi_b = ib.kchi[2 * b + 1];
    r.lK0 = log(fmin(cfg.k_min * r.chi_lo_a, cfg.k_min * r.chi_lo_b));
    r.lK1 = log(fmax(cfg.k_max * r.chi_hi_a, cfg.k_max * r.chi_hi_b));
    return r;
}

// grid (4, B), 128 threads, one warp per ln K node: table t = blockIdx.x of a, b, ab, ba into proj4 [B, 4, 2, n_kernel]
__global__ void __launch_bounds__(128)
cov_projected_cross_kernel(const Cfg cfg, const CovP cp, int B, LimberIn ia, LimberIn ib, HaloSide ha, HaloSide hb,
                           double* __restrict__ proj4, int32_t* __restrict__ status) {
    extern __shared__ double dyn[];
    __shared__ double pe[PROJ_MAX_PIECES + 1];
    __shared__ double me[COV_MAX_EDGES];
    __shared__ int n_piece_s;
    __shared__ double vals[COV_MAX_COLS], fac[COV_MAX_COLS];
    const int b = blockIdx.y, t = blockIdx.x;
    if (b >= B) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    const int nk = cfg.n_kernel, nq = cfg.nq_limber;
    const Limber4 F = limber_stage4(cfg, ia, ib, b, dyn);
    PkParams pka, pkb;
    const HaloTabs Ta = halo_side_tabs(cfg, ha, b, pka, cfg.cosmo_precision);
    const HaloTabs Tb = halo_side_tabs(cfg, hb, b, pkb, cfg.cosmo_precision);
    BaoParams bao_store;                 // one cosmology: both sides share the wiggle constants
    if (cfg.with_bao) {
        make_bao(load_cosmo(ha.cosmo + (size_t)b * CHOMP_N_COSMO, cfg.cosmo_precision), &bao_store);
        pka.bao = &bao_store; pkb.bao = &bao_store;
    }
    const CrossRanges R = cross_ranges(cfg, ia, ib, b);
    // windows and chi range of this table (covariance.py:479-536)
    const int w1 = (t == 0 || t == 2) ? 0 : (t == 1 ? 2 : 1);
    const int w2 = (t == 0) ? 1 : (t == 1 ? 3 : (t == 2 ? 3 : 2));
    double c_lo, c_hi;
    if (t == 0) { c_lo = R.chi_lo_a; c_hi = R.chi_hi_a; }
    else if (t == 1) { c_lo = R.chi_lo_b; c_hi = R.chi_hi_b; }
    else { c_lo = fmax(R.chi_lo_a, R.chi_lo_b); c_hi = fmin(R.chi_hi_a, R.chi_hi_b); }
    __syncthreads();
    if (tid == 0) {
        int cnt = 0;
        bool over = false;
        int ne = 0;
        if (c_hi > c_lo)
            ne = merge_edges(ia.edges + (size_t)b * ia.edge_stride, ia.n_edges[b], ib.edges + (size_t)b * ib.edge_stride,
                             ib.n_edges[b], c_lo, c_hi, me, COV_MAX_EDGES);
        if (ne >= 2) {
            pe[cnt++] = me[0];
            for (int p = 0; p < ne - 1; ++p) {
                const double a = me[p], bb = me[p + 1];
                int ns = (int)ceil(log(bb / a) / COV_PIECE - 1e-9);
                if (ns < 1) ns = 1;
                if (cnt + ns > PROJ_MAX_PIECES) { ns = 1; over = true; }
                const double r = log(bb / a) / ns;
                for (int k2 = 1; k2 < ns; ++k2) pe[cnt++] = a * exp(r * k2);
                pe[cnt++] = bb;
            }
        }
        n_piece_s = cnt > 0 ? cnt - 1 : 0;
        if (over && status) atomicOr(status + b, CHOMP_ST_NODE_OVERFLOW);
        nak_uniform_factors(nk, fac);
    }
    __syncthreads();
    const int n_piece = n_piece_s;
    const double hK = (R.lK1 - R.lK0) / (nk - 1);
    for (int j = wid; j < nk; j += nwarp) {
        const double lK = (j == nk - 1) ? R.lK1 : R.lK0 + hK * j;
        const double K = exp(lK);
        const double lo = fmax(K / cfg.k_max, c_lo), hi = fmin(K / cfg.k_min, c_hi);
        double acc = 0.0;
        if (hi > lo) {
            for (int idx = lane; idx < n_piece * nq; idx += 32) {
                const int p = idx / nq, q = idx - p * nq;
                const double a = fmax(pe[p], lo), bb = fmin(pe[p + 1], hi);
                if (bb > a) {
                    const double half = 0.5 * (bb - a);
                    const double chi = 0.5 * (a + bb) + half * c_glx[nq][q];
                    double P;
                    if (t == 0) P = halo_power(Ta, pka, cp.which, K / chi);
                    else if (t == 1) P = halo_power(Tb, pkb, cp.which, K / chi);
                    else P = sqrt(halo_power(Ta, pka, cp.which, K / chi) * halo_power(Tb, pkb, cp.which, K / chi));   // covariance.py:569-591
                    acc += half * c_glw[nq][q] * P * F.pair(w1, w2, chi) / (chi * chi);
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) vals[j] = acc;
    }
    __syncthreads();
    double* P = proj4 + ((size_t)b * 4 + t) * 2 * nk;
    if (tid == 0) nak_uniform_solve(nk, hK, vals, 1, P + nk, 1, fac);
    bool bad = false;
    for (int j = tid; j < nk; j += blockDim.x) { P[j] = vals[j]; if (!isfinite(vals[j])) bad = true; }
    if (bad && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
}

// K^2 [ (P_a P_b + P_a N_2 + P_b N_0) + (P_ab P_ba + P_ab N_1 + P_ba N_3) ]   (covariance.py:421-453)
struct GaussS4 {
    const double* proj;   // [4][2 nk] in shared memory
    int nk;
    double x0, h, inv_da2, inv_db2, inv_dab, n0, n1, n2, n3;
    __device__ __forceinline__ double tab(int t, int j, double u) const {
        const double* p = proj + (size_t)t * 2 * nk;
        return nak_eval(p[j], p[j + 1], p[nk + j], p[nk + j + 1], h, u);
    }
    __device__ __forceinline__ double operator()(double x) const {
        int j = (int)floor((x - x0) / h);
        j = j < 0 ? 0 : (j > nk - 2 ? nk - 2 : j);
        const double u = (x - (x0 + h * j)) / h;
        const double Pa = tab(0, j, u) * inv_da2, Pb = tab(1, j, u) * inv_db2;
        const double Pab = tab(2, j, u) * inv_dab, Pba = tab(3, j, u) * inv_dab;
        const double K = exp(x);
        return K * K * ((Pa * Pb + Pa * n2 + Pb * n0) + (Pab * Pba + Pab * n1 + Pba * n3));
    }
};

__global__ void __launch_bounds__(COV_THREADS)
cov_g_cross_kernel(const Cfg cfg, const CovP cp, int B, LimberIn ia, LimberIn ib, const double* __restrict__ bin_center,
                   const double* __restrict__ proj4, CovOut out) {
    __shared__ OscShared s;
    __shared__ double pr[8 * COV_MAX_COLS];
    const int b = blockIdx.y;
    if (b >= B) return;
    const int nb = cp.n_bins, nk = cfg.n_kernel, tid = threadIdx.x;
    const int row = nb - 1 - blockIdx.x;
    for (int i = tid; i < 8 * nk; i += blockDim.x) pr[i] = proj4[(size_t)b * 8 * nk + i];
    const CrossRanges R = cross_ranges(cfg, ia, ib, b);
    const double lK0 = R.lK0, lK1 = R.lK1, hK = (lK1 - lK0) / (nk - 1);
    int sub = (int)ceil(hK / COV_PIECE - 1e-9);
    if (sub < 1) sub = 1;
    while ((nk - 1) * sub + 1 > COV_MAX_EDGES) --sub;
    const int n_edge = (nk - 1) * sub + 1;
    for (int eidx = tid; eidx < n_edge; eidx += blockDim.x) {
        const int i = eidx / sub, r = eidx - i * sub;
        const double a = lK0 + hK * i, bb = (i >= nk - 2) ? lK1 : lK0 + hK * (i + 1);
        s.edge[eidx] = (eidx == n_edge - 1) ? lK1 : a + (bb - a) * r / sub;
    }
    for (int a = tid; a <= row; a += blockDim.x) {
        const double th = bin_center[a];
        s.fj[a] = th;
        s.top[a] = fmin(log(cp.bessel_limit / th), lK1);         // covariance.py:375-380 (theta_a <= theta_b)
    }
    __syncthreads();
    const double Da = ia.dbar[b], Db = ib.dbar[b];
    GaussS4 S{pr, nk, lK0, hK, 1.0 / (Da * Da), 1.0 / (Db * Db), 1.0 / (Da * Db), cp.poisson[0], cp.poisson[1], cp.poisson[2],
              cp.poisson[3]};
    osc_row(S, ExpU(), s, n_edge, bin_center[row], row + 1, cp.nq_osc, cp.osc_phase);
    double* G = out.parts + ((size_t)b * 3 + 1) * nb * nb;
    const double norm = 1.0 / (2.0 * M_PI * cp.area_sr);
    for (int a = tid; a <= row; a += blockDim.x) {
        const double v = (s.flag[a] & 1) ? 0.0 : s.acc[a] * norm;
        G[(size_t)a * nb + row] = v;
        G[(size_t)row * nb + a] = v;
    }
}

}  // namespace chomp
