// Covariance of w(theta): Poisson + Gaussian + 1-halo non-Gaussian terms for matching
// correlations (reference covariance.Covariance, covariance.py:23-683, and
// kernel.KernelCovariance, kernel.py:864-1111).
//
//   cov_kng_kernel       K_NG(x_a, x_b) = int dchi (W_a W_b)^2 D^4 / chi^2 J0(e^{x_a} chi) J0(e^{x_b} chi)
//                        on the n_kernel x n_kernel grid of ln(k theta) nodes, z_bar_NG, D(z_bar_NG)
//                        (kernel.py:961-972, 1016-1068: 1275 Romberg integrals per cosmology there)
//   cov_kng_spline_kernel  ln(K - 10 K_min) and its not-a-knot second derivatives along the first
//                        axis (RectBivariateSpline of kernel.py:1025-1029)
//   cov_projected_kernel P_a(K) = int dchi P(K / chi) W_a W_b D^2 / chi^2 on the ln K nodes
//                        (covariance.py:455-543)
//   cov_tri_nodes_kernel bicubic T(k_a node, k_b quadrature node) from the 1-halo trispectrum table
//                        (halo_trispectrum.py:100-129) times the k_b quadrature weights
//   cov_g_kernel         Gaussian term, covariance.py:359-453
//   cov_ng_kernel        non-Gaussian term, covariance.py:593-683 (any bins)
//   cov_ng_shift_kernel  the same for log-spaced bins: kernel values shared by all theta_b (a Toeplitz product)
//   cov_finish_kernel    Poisson term (covariance.py:323-357) and assembly (:276-321)
//
// Both oscillatory families (K_NG rows, Gaussian-term rows) are integrals
//     I_ij = int_{x_0}^{top_j} S(x) J0(f_i u(x)) J0(f_j u(x)) dx,    f_j <= f_i,
// with the truncation top_j set by the slower Bessel factor.  One CTA owns a row i: it lays a
// Gauss-Legendre grid fine enough for f_i over the smooth panels of S once, tabulates
// S J0(f_i u) there, and every column j re-uses those nodes (one J0 per node and column).
#pragma once
#include "common.cuh"
#include "hankel.cuh"
#include "limber_tables.cuh"
#include "special.cuh"
#include "spline.cuh"

namespace chomp {

typedef chomp_b200_cov_params CovP;

#define COV_THREADS 256
#define COV_CHUNK 1024            // fine nodes staged per pass
#define COV_MAX_EDGES 1280        // smooth-panel edges of one row integral
#define COV_MAX_COLS 128          // columns per row (n_kernel or n_bins)
#define COV_PIECE 0.0625          // widest piece in ln k / ln K / ln chi (as HANKEL_MAX_PIECE)

// ---------------------------------------------------------------------------------------
// uniform-grid not-a-knot cubic splines through their second derivatives: the Thomas factors
// depend on n only, so a table family shares them and a solve needs no scratch
// ---------------------------------------------------------------------------------------
__device__ inline void nak_uniform_factors(int n, double* cp) {
    cp[0] = 0.0; cp[1] = 0.0;
    for (int i = 2; i <= n - 3; ++i) cp[i] = 1.0 / (4.0 - cp[i - 1]);
    cp[n - 2] = 0.0; cp[n - 1] = 0.0;
}
// M[i * sm] = second derivative at node i of the not-a-knot spline through y[i * sy], spacing h (n >= 4)
__device__ inline void nak_uniform_solve(int n, double h, const double* y, int sy, double* M, int sm, const double* cp) {
    const double s = 6.0 / (h * h);
#define NAK_R(i) (s * (y[((i) + 1) * sy] - 2.0 * y[(i) * sy] + y[((i) - 1) * sy]))
    M[sm] = NAK_R(1) / 6.0;
    for (int i = 2; i <= n - 3; ++i) M[i * sm] = (NAK_R(i) - M[(i - 1) * sm]) * cp[i];
    M[(n - 2) * sm] = NAK_R(n - 2) / 6.0;
#undef NAK_R
    for (int i = n - 3; i >= 1; --i) M[i * sm] -= cp[i] * M[(i + 1) * sm];
    M[0] = 2.0 * M[sm] - M[2 * sm];
    M[(n - 1) * sm] = 2.0 * M[(n - 2) * sm] - M[(n - 3) * sm];
}
// value on interval j at fraction t in [0, 1]
__device__ __forceinline__ double nak_eval(double y0, double y1, double m0, double m1, double h, double t) {
    const double a = 1.0 - t;
    return a * y0 + t * y1 + ((a * a * a - a) * m0 + (t * t * t - t) * m1) * (h * h * (1.0 / 6.0));
}

// ---------------------------------------------------------------------------------------
// per-point Limber context rebuilt from the stage-1 outputs
// ---------------------------------------------------------------------------------------
struct LimberIn {
    const double *grid0, *win_chi, *win_coef, *kchi, *edges, *zbar, *dbar;
    const int32_t* n_edges;
    int edge_stride;
};

// stage the chi(z) / z(chi) / D(z) tables and both window splines of point b in shared memory
// (13 n_cosmo + 8 n_window doubles) and return the W_a W_b D^2 functor on them
__device__ inline LimberF limber_stage(const Cfg& cfg, const LimberIn& in, int b, double* sm) {
    const int nz = cfg.n_cosmo, nw = cfg.n_window;
    for (int i = threadIdx.x; i < 13 * nz; i += blockDim.x) sm[i] = in.grid0[(size_t)b * 13 * nz + i];
    double* wc = sm + 13 * nz;
    for (int i = threadIdx.x; i < 8 * nw; i += blockDim.x) wc[i] = in.win_coef[(size_t)b * 8 * nw + i];
    LimberF F;
    F.g.n = nz; F.g.z_min = cfg.zk_min < 0.0 ? 0.0 : cfg.zk_min; F.g.z_max = cfg.zk_max;
    F.g.chi = sm; F.g.c_chi_z = sm + nz; F.g.c_z_chi = sm + 5 * nz; F.g.c_g_z = sm + 9 * nz;
    F.g.z = nullptr; F.g.growth = nullptr;
    const double* c4 = in.win_chi + (size_t)b * 4;
    F.a = Window{nw, c4[0], c4[1], nullptr, wc};
    F.b = Window{nw, c4[2], c4[3], nullptr, wc + 4 * nw};
    return F;
}
__host__ __device__ inline size_t limber_stage_doubles(const Cfg& cfg) { return 13 * (size_t)cfg.n_cosmo + 8 * (size_t)cfg.n_window; }

// z range shared by the windows (kernel.py:594-597, 909-915)
__device__ inline void window_z_range(const Cfg& cfg, double& zlo, double& zhi) {
    zlo = -1e300; zhi = 1e300;
    for (int i = 0; i < 2; ++i) {
        double a = (cfg.window_kind[i] == CHOMP_WINDOW_GALAXY) ? cfg.dndz_zmin[i] : 0.0;
        if (a < cfg.window_precision) a = cfg.window_precision;
        zlo = fmax(zlo, a);
        zhi = fmin(zhi, cfg.dndz_zmax[i]);
    }
}

// ---------------------------------------------------------------------------------------
// row-shared oscillatory integrals (see the header comment).  All threads of the CTA call it.
// ---------------------------------------------------------------------------------------
struct OscShared {
    double edge[COV_MAX_EDGES];
    int pfx[COV_MAX_EDGES + 1];
    double fj[COV_MAX_COLS], top[COV_MAX_COLS], ulim[COV_MAX_COLS], lim[COV_MAX_COLS], acc[COV_MAX_COLS];
    int flag[COV_MAX_COLS];        // bit 0: zero (top <= x_min), bit 1: partial piece present
    double u[COV_CHUNK], g[COV_CHUNK];
};

template <class SF, class UF>
__device__ void osc_row(const SF& S, const UF& U, OscShared& s, int n_edge, double f_i, int nj, int nq, double phase) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    // nothing is needed beyond the largest truncation point: clip the panel list there
    __shared__ int n_edge_s;
    if (tid == 0) {
        double tmax = -1e300;
        for (int j = 0; j < nj; ++j) tmax = fmax(tmax, s.top[j]);
        int ne = n_edge;
        if (tmax < s.edge[n_edge - 1] && tmax > s.edge[0]) {
            const int p = search_index(tmax, s.edge, n_edge);
            if (tmax > s.edge[p]) { s.edge[p + 1] = tmax; ne = p + 2; } else ne = p + 1;
        }
        n_edge_s = ne;
    }
    __syncthreads();
    n_edge = n_edge_s;
    const int n_pan = n_edge - 1;
    // pieces per panel: phase advance of the fast factor <= `phase` per piece
    for (int p = tid; p < n_pan; p += blockDim.x) {
        const double adv = f_i * (U(s.edge[p + 1]) - U(s.edge[p]));
        int cnt = (int)ceil(fmin(adv / phase, 65536.0));
        s.pfx[p + 1] = (cnt < 1 ? 1 : cnt) * nq;
    }
    __syncthreads();
    if (wid == 0) {                                   // exclusive prefix sums by one warp
        const int per = (n_pan + 31) / 32;
        const int p0 = lane * per, p1 = min(n_pan, p0 + per);
        int mine = 0;
        for (int p = p0; p < p1; ++p) mine += s.pfx[p + 1];
        int incl = mine;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - mine;
        for (int p = p0; p < p1; ++p) { const int c = s.pfx[p + 1]; s.pfx[p + 1] = run + c; run += c; }
        if (lane == 0) s.pfx[0] = 0;
    }
    __syncthreads();
    const double x_min = s.edge[0], x_max = s.edge[n_pan];
    for (int j = tid; j < nj; j += blockDim.x) {
        s.acc[j] = 0.0;
        const double top = s.top[j];
        int fl = 0;
        double lim = x_max, ulim = 1e300;
        if (!(top > x_min)) fl = 1;
        else if (top < x_max) {
            const int p = search_index(top, s.edge, n_edge);
            const int nsub = (s.pfx[p + 1] - s.pfx[p]) / nq;
            const double a = s.edge[p], d = (s.edge[p + 1] - a) / nsub;
            int k = (int)floor((top - a) / d);
            k = k < 0 ? 0 : (k > nsub - 1 ? nsub - 1 : k);
            lim = a + d * k;
            ulim = U(lim);
            if (top > lim) fl |= 2;
        }
        s.flag[j] = fl; s.lim[j] = lim; s.ulim[j] = ulim;
    }
    __syncthreads();
    const int total = s.pfx[n_pan];
    for (int c0 = 0; c0 < total; c0 += COV_CHUNK) {
        const int m = min(COV_CHUNK, total - c0);
        for (int idx = tid; idx < m; idx += blockDim.x) {
            const int gidx = c0 + idx;
            int lo = 0, hi = n_pan;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s.pfx[mid] <= gidx) lo = mid; else hi = mid; }
            const int local = gidx - s.pfx[lo];
            const int nsub = (s.pfx[lo + 1] - s.pfx[lo]) / nq;
            const int sidx = local / nq, q = local - sidx * nq;
            const double a = s.edge[lo], d = (s.edge[lo + 1] - a) / nsub, half = 0.5 * d;
            const double x = a + d * (sidx + 0.5) + half * c_glx[nq][q];
            const double u = U(x);
            s.u[idx] = u;
            s.g[idx] = half * c_glw[nq][q] * S(x) * j0(f_i * u);
        }
        __syncthreads();
        for (int j = wid; j < nj; j += nwarp) {
            if (s.flag[j] & 1) continue;
            const double f = s.fj[j], ul = s.ulim[j];
            double part = 0.0;
            for (int idx = lane; idx < m; idx += 32) {
                const double u = s.u[idx];
                if (u < ul) part = fma(s.g[idx], j0(f * u), part);
            }
            part = warp_sum(part);
            if (lane == 0) s.acc[j] += part;
        }
        __syncthreads();
    }
    // the piece cut by top_j
    for (int j = wid; j < nj; j += nwarp) {
        if ((s.flag[j] & 3) != 2) continue;
        double part = 0.0;
        if (lane < nq) {
            const double a = s.lim[j], half = 0.5 * (s.top[j] - a);
            const double x = a + half + half * c_glx[nq][lane];
            const double u = U(x);
            part = half * c_glw[nq][lane] * S(x) * j0(f_i * u) * j0(s.fj[j] * u);
        }
        part = warp_sum(part);
        if (lane == 0) s.acc[j] += part;
    }
    __syncthreads();
}

// Gauss-Legendre order of the k_b pieces of the non-Gaussian term and the node count that goes with it
__host__ __device__ inline int cov_ng_order(const Cfg& cfg, const CovP& cp) { return cp.nq_ng > 0 ? cp.nq_ng : cfg.nq_hankel; }
__host__ __device__ inline int cov_ng_nodes(const Cfg& cfg, const CovP& cp) {
    return (cfg.n_halo - 1) * hankel_subdiv(cfg) * cov_ng_order(cfg, cp);
}

struct CovOut {
    double *kng;        // [B, n_kernel, n_kernel]   K_NG table
    double *lkng;       // [B, n_kernel, n_kernel]   ln(K - 10 K_min)
    double *mkng;       // [B, n_kernel, n_kernel]   its second derivatives along axis 0
    double *kng_min;    // [B]
    double *zbar_ng;    // [B]
    double *d_ng;       // [B]
    double *proj;       // [B, 2, n_kernel]          projected spectrum nodes, second derivatives
    double *parts;      // [B, 3, n_bins, n_bins]    P, G, NG
};

struct IdentityU { __device__ __forceinline__ double operator()(double x) const { return x; } };
struct ExpU { __device__ __forceinline__ double operator()(double x) const { return exp(x); } };
struct KngS {      // (W_a W_b)^2 D^4 / chi^2, kernel.py:1102-1111 with a1 = b1, a2 = b2
    LimberF F;
    __device__ __forceinline__ double operator()(double chi) const { const double f = F(chi); return f * f / (chi * chi); }
};

// grid (n_kernel, B): row i = n_kernel - 1 - blockIdx.x (the long rows first)
__global__ void __launch_bounds__(COV_THREADS)
cov_kng_kernel(const Cfg cfg, const CovP cp, int B, LimberIn in, CovOut out) {
    extern __shared__ double dyn[];
    __shared__ OscShared s;
    const int b = blockIdx.y;
    if (b >= B) return;
    const int nk = cfg.n_kernel;
    const int i = nk - 1 - blockIdx.x;
    const int tid = threadIdx.x;
    KngS S{limber_stage(cfg, in, b, dyn)};
    const int n_edge = in.n_edges[b];
    for (int e = tid; e < n_edge; e += blockDim.x) s.edge[e] = in.edges[(size_t)b * in.edge_stride + e];
    const double x0 = log(cp.theta_min_rad * cfg.k_min), x1 = log(cp.theta_max_rad * cfg.k_max);
    const double hx = (x1 - x0) / (nk - 1);
    const double chi_min = in.kchi[2 * b], chi_max = in.kchi[2 * b + 1];
    for (int j = tid; j <= i; j += blockDim.x) {
        const double kt = exp((j == nk - 1) ? x1 : x0 + hx * j);
        s.fj[j] = kt;
        s.top[j] = fmin(cp.bessel_limit / kt, chi_max);       // kernel.py:1047-1052
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid < 32) {
        // z_bar_NG: arg-max of the weight on the n_kernel-point z grid (kernel.py:961-972)
        double zlo, zhi;
        window_z_range(cfg, zlo, zhi);
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
    const double kt_i = exp((i == nk - 1) ? x1 : x0 + hx * i);
    osc_row(S, IdentityU(), s, n_edge, kt_i, i + 1, cp.nq_osc, cp.osc_phase);
    double* K = out.kng + (size_t)b * nk * nk;
    for (int j = tid; j <= i; j += blockDim.x) {
        const double v = (s.flag[j] & 1) ? 0.0 : s.acc[j];
        K[(size_t)i * nk + j] = v;
        K[(size_t)j * nk + i] = v;
    }
    (void)chi_min;
}

// grid (B): K_min, ln(K - 10 K_min), second derivatives of every column along axis 0
__global__ void __launch_bounds__(COV_THREADS)
cov_kng_spline_kernel(const Cfg cfg, const CovP cp, int B, CovOut out, int32_t* __restrict__ status) {
    __shared__ double red[32];
    __shared__ double fac[COV_MAX_COLS];
    const int b = blockIdx.x;
    if (b >= B) return;
    const int nk = cfg.n_kernel, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double* K = out.kng + (size_t)b * nk * nk;
    double m = 1e300;
    for (int idx = tid; idx < nk * nk; idx += blockDim.x) m = fmin(m, K[idx]);
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[wid] = m;
    if (tid == 0) nak_uniform_factors(nk, fac);
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmin(m, red[w]);
    double* L = out.lkng + (size_t)b * nk * nk;
    bool bad = false;
    for (int idx = tid; idx < nk * nk; idx += blockDim.x) {
        const double v = log(K[idx] - 10.0 * m);               // kernel.py:1027-1029
        L[idx] = v;
        if (!isfinite(v)) bad = true;
    }
    if (tid == 0) out.kng_min[b] = m;
    if (bad && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
    __syncthreads();
    const double x0 = log(cp.theta_min_rad * cfg.k_min), x1 = log(cp.theta_max_rad * cfg.k_max);
    const double hx = (x1 - x0) / (nk - 1);
    double* M = out.mkng + (size_t)b * nk * nk;
    for (int col = tid; col < nk; col += blockDim.x) nak_uniform_solve(nk, hx, L + col, nk, M + col, nk, fac);
}

// K_NG evaluation pieces shared by the non-Gaussian kernel: values of the n_kernel column splines at x
// ---------------------------------------------------------------------------------------
// projected spectrum P_a(K) on the ln K nodes.  grid (B), 128 threads, one warp per node.
// Pieces: the base panels of the Limber integrals, cut geometrically so that no piece spans
// more than COV_PIECE in ln chi (P(K / chi) is a spline in ln k).
// ---------------------------------------------------------------------------------------
#define PROJ_MAX_PIECES 1536
__global__ void __launch_bounds__(128)
cov_projected_kernel(const Cfg cfg, const CovP cp, int B, LimberIn in, const double* __restrict__ cosmo,
                     const double* __restrict__ epoch, const double* __restrict__ htab, const double* __restrict__ hcoef,
                     const double* __restrict__ hfit, CovOut out, int32_t* __restrict__ status) {
    extern __shared__ double dyn[];
    __shared__ double pe[PROJ_MAX_PIECES + 1];
    __shared__ int n_piece_s;
    __shared__ double vals[COV_MAX_COLS], fac[COV_MAX_COLS];
    const int b = blockIdx.x;
    if (b >= B) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    const int nk = cfg.n_kernel, nq = cfg.nq_limber;
    LimberF F = limber_stage(cfg, in, b, dyn);
    const Cosmo c = load_cosmo(cosmo + (size_t)b * CHOMP_N_COSMO, cfg.cosmo_precision);
    const double* e = epoch + (size_t)b * CHOMP_EPOCH_LEN;
    PkParams pk = make_pk(c, e[EP_GROWTH], e[EP_SIGMA_NORM]);
    CHOMP_ATTACH_BAO(cfg, c, pk)
    HaloTabs T;
    T.nk = cfg.n_halo; T.l0 = log(cfg.k_min); T.l1 = log(cfg.k_max); T.h = (T.l1 - T.l0) / (T.nk - 1);
    T.k_min = cfg.k_min; T.k_max = cfg.k_max; T.extrapolate = cfg.extrapolate;
    T.tab = htab + (size_t)b * 5 * T.nk; T.coef = hcoef + (size_t)b * 20 * T.nk;
    T.hf = hfit ? hfit + (size_t)b * HF_LEN : nullptr;
    const double* ed = in.edges + (size_t)b * in.edge_stride;
    const int n_pan = in.n_edges[b] - 1;
    if (tid == 0) {
        int cnt = 0;
        bool over = false;
        pe[cnt++] = ed[0];
        for (int p = 0; p < n_pan; ++p) {
            const double a = ed[p], bb = ed[p + 1];
            int ns = (int)ceil(log(bb / a) / COV_PIECE - 1e-9);
            if (ns < 1) ns = 1;
            if (cnt + ns > PROJ_MAX_PIECES) { ns = 1; over = true; }
            const double r = log(bb / a) / ns;
            for (int k2 = 1; k2 < ns; ++k2) pe[cnt++] = a * exp(r * k2);
            pe[cnt++] = bb;
        }
        n_piece_s = cnt - 1;
        if (over && status) atomicOr(status + b, CHOMP_ST_NODE_OVERFLOW);
        nak_uniform_factors(nk, fac);
    }
    __syncthreads();
    const int n_piece = n_piece_s;
    // covariance.py:133-140, 160-172: chi range of the pair of windows, ln K nodes
    const double chi_lo = in.kchi[2 * b], chi_hi = in.kchi[2 * b + 1];
    const double lK0 = log(cfg.k_min * chi_lo), lK1 = log(cfg.k_max * chi_hi), hK = (lK1 - lK0) / (nk - 1);
    for (int j = wid; j < nk; j += nwarp) {
        const double lK = (j == nk - 1) ? lK1 : lK0 + hK * j;
        const double K = exp(lK);
        const double lo = fmax(K / cfg.k_max, chi_lo), hi = fmin(K / cfg.k_min, chi_hi);    // covariance.py:479-484
        double acc = 0.0;
        if (hi > lo) {
            for (int idx = lane; idx < n_piece * nq; idx += 32) {
                const int p = idx / nq, q = idx - p * nq;
                const double a = fmax(pe[p], lo), bb = fmin(pe[p + 1], hi);
                if (bb > a) {
                    const double half = 0.5 * (bb - a);
                    const double chi = 0.5 * (a + bb) + half * c_glx[nq][q];
                    acc += half * c_glw[nq][q] * halo_power(T, pk, cp.which, K / chi) * F(chi) / (chi * chi);
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) vals[j] = acc;
    }
    __syncthreads();
    double* P = out.proj + (size_t)b * 2 * nk;
    if (tid == 0) nak_uniform_solve(nk, hK, vals, 1, P + nk, 1, fac);
    bool bad = false;
    for (int j = tid; j < nk; j += blockDim.x) { P[j] = vals[j]; if (!isfinite(vals[j])) bad = true; }
    if (bad && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
}

// C(l) = int dchi P(l / chi) / D(z_bar)^2  W_a W_b D^2 / chi^2 over the kernel's chi range for any of the
// spectra (CorrelationFourier.correlation, correlation.py:360-392): the same integrand as the projected
// spectrum above without the clipping; every piece is split at chi = l / k_max and l / k_min, where the
// halo-model spectra change branch (halo.py:277-439).  grid (B), 128 threads, one warp per l.
__global__ void __launch_bounds__(128)
cl_table_kernel(const Cfg cfg, int which, int B, int n_ell, const double* __restrict__ ell, LimberIn in,
                const double* __restrict__ cosmo, const double* __restrict__ epoch, const double* __restrict__ htab,
                const double* __restrict__ hcoef, const double* __restrict__ hfit, double* __restrict__ out,
                int32_t* __restrict__ status) {
    extern __shared__ double dyn[];
    __shared__ double pe[PROJ_MAX_PIECES + 1];
    __shared__ int n_piece_s;
    const int b = blockIdx.x;
    if (b >= B) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    const int nq = cfg.nq_limber;
    LimberF F = limber_stage(cfg, in, b, dyn);
    const Cosmo c = load_cosmo(cosmo + (size_t)b * CHOMP_N_COSMO, cfg.cosmo_precision);
    const double* e = epoch + (size_t)b * CHOMP_EPOCH_LEN;
    PkParams pk = make_pk(c, e[EP_GROWTH], e[EP_SIGMA_NORM]);
    CHOMP_ATTACH_BAO(cfg, c, pk)
    HaloTabs T;
    T.nk = cfg.n_halo; T.l0 = log(cfg.k_min); T.l1 = log(cfg.k_max); T.h = (T.l1 - T.l0) / (T.nk - 1);
    T.k_min = cfg.k_min; T.k_max = cfg.k_max; T.extrapolate = cfg.extrapolate;
    T.tab = htab + (size_t)b * 5 * T.nk; T.coef = hcoef + (size_t)b * 20 * T.nk;
    T.hf = hfit ? hfit + (size_t)b * HF_LEN : nullptr;
    const double* ed = in.edges + (size_t)b * in.edge_stride;
    const int n_pan = in.n_edges[b] - 1;
    if (tid == 0) {
        int cnt = 0;
        bool over = false;
        pe[cnt++] = ed[0];
        for (int p = 0; p < n_pan; ++p) {
            const double a = ed[p], bb = ed[p + 1];
            int ns = (int)ceil(log(bb / a) / COV_PIECE - 1e-9);
            if (ns < 1) ns = 1;
            if (cnt + ns > PROJ_MAX_PIECES) { ns = 1; over = true; }
            const double r = log(bb / a) / ns;
            for (int k2 = 1; k2 < ns; ++k2) pe[cnt++] = a * exp(r * k2);
            pe[cnt++] = bb;
        }
        n_piece_s = cnt - 1;
        if (over && status) atomicOr(status + b, CHOMP_ST_NODE_OVERFLOW);
    }
    __syncthreads();
    const int n_piece = n_piece_s;
    const double D = in.dbar[b], inv_d2 = 1.0 / (D * D);
    for (int il = wid; il < n_ell; il += nwarp) {
        const double l = ell[il];
        const double c_lo = l / cfg.k_max, c_hi = l / cfg.k_min;
        double acc = 0.0;
        for (int idx = lane; idx < n_piece * 3 * nq; idx += 32) {
            const int p = idx / (3 * nq), r = idx - p * 3 * nq;
            const int reg = r / nq, q = r - reg * nq;
            const double r_lo = reg == 0 ? -1e300 : (reg == 1 ? c_lo : c_hi);
            const double r_hi = reg == 0 ? c_lo : (reg == 1 ? c_hi : 1e300);
            const double a = fmax(pe[p], r_lo), bb = fmin(pe[p + 1], r_hi);
            if (bb > a) {
                const double half = 0.5 * (bb - a);
                const double chi = 0.5 * (a + bb) + half * c_glx[nq][q];
                acc += half * c_glw[nq][q] * halo_power(T, pk, which, l / chi) * F(chi) / (chi * chi);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            out[(size_t)b * n_ell + il] = acc * inv_d2;
            if (!isfinite(acc) && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Gaussian term.  grid (n_bins, B): row = the wider bin theta_b (index n_bins - 1 - blockIdx.x),
// columns a <= b.  covariance.py:359-453 with matching correlations: both two-point terms
// are equal, P_a = P_b.
// ---------------------------------------------------------------------------------------
struct GaussS {
    const double* proj;   // [2 nk] nodes, second derivatives (shared memory)
    int nk;
    double x0, h, inv_d2, poiss;
    __device__ __forceinline__ double operator()(double x) const {
        int j = (int)floor((x - x0) / h);
        j = j < 0 ? 0 : (j > nk - 2 ? nk - 2 : j);
        const double t = (x - (x0 + h * j)) / h;
        const double P = nak_eval(proj[j], proj[j + 1], proj[nk + j], proj[nk + j + 1], h, t) * inv_d2;
        const double K = exp(x);
        return K * K * 2.0 * (P * P + P * poiss);
    }
};

__global__ void __launch_bounds__(COV_THREADS)
cov_g_kernel(const Cfg cfg, const CovP cp, int B, LimberIn in, const double* __restrict__ bin_center, CovOut out) {
    __shared__ OscShared s;
    __shared__ double pr[2 * COV_MAX_COLS];
    const int b = blockIdx.y;
    if (b >= B) return;
    const int nb = cp.n_bins, nk = cfg.n_kernel, tid = threadIdx.x;
    const int row = nb - 1 - blockIdx.x;
    for (int i = tid; i < 2 * nk; i += blockDim.x) pr[i] = out.proj[(size_t)b * 2 * nk + i];
    const double chi_lo = in.kchi[2 * b], chi_hi = in.kchi[2 * b + 1];
    const double lK0 = log(cfg.k_min * chi_lo), lK1 = log(cfg.k_max * chi_hi), hK = (lK1 - lK0) / (nk - 1);
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
    const double D = in.dbar[b];
    GaussS S{pr, nk, lK0, hK, 1.0 / (D * D), cp.poisson[2] + cp.poisson[0]};
    osc_row(S, ExpU(), s, n_edge, bin_center[row], row + 1, cp.nq_osc, cp.osc_phase);
    double* G = out.parts + ((size_t)b * 3 + 1) * nb * nb;
    const double norm = 1.0 / (2.0 * M_PI * cp.area_sr);
    for (int a = tid; a <= row; a += blockDim.x) {
        const double v = (s.flag[a] & 1) ? 0.0 : s.acc[a] * norm;
        G[(size_t)a * nb + row] = v;
        G[(size_t)row * nb + a] = v;
    }
}

// ---------------------------------------------------------------------------------------
// 1-halo trispectrum at (k_a node i, k_b quadrature node q), bicubic in (ln k_a, ln k_b)
// (RectBivariateSpline of halo_trispectrum.py:123-125), times w_q k_b^2 / D_NG^4.
// grid (chunk), 256 threads.  k_b nodes: the Hankel rule's (every interval of the table's
// ln k grid cut into pieces <= 0.0625, nq_hankel points each).
// ---------------------------------------------------------------------------------------
struct TriScratch {
    double *mcol;   // [chunk, n_halo, n_halo]  second derivatives of the columns along k_a
    double *r;      // [chunk, n_kernel, n_halo] table interpolated at the k_a nodes
    double *m2;     // [chunk, n_kernel, n_halo] second derivatives of those rows along k_b
    double *tw;     // [chunk, n_kernel, n_q]
};

__global__ void __launch_bounds__(COV_THREADS)
cov_tri_nodes_kernel(const Cfg cfg, const CovP cp, int b0, int nb_chunk, const double* __restrict__ T,
                     const double* __restrict__ d_ng, TriScratch ts) {
    __shared__ double fac[1024];
    const int cidx = blockIdx.x;
    if (cidx >= nb_chunk) return;
    const int b = b0 + cidx;
    const int nh = cfg.n_halo, nk = cfg.n_kernel, nq = cov_ng_order(cfg, cp), tid = threadIdx.x;
    const int sub = hankel_subdiv(cfg), ntot = (nh - 1) * sub * nq;
    const double l0 = log(cfg.k_min), l1 = log(cfg.k_max), hT = (l1 - l0) / (nh - 1), hA = (l1 - l0) / (nk - 1);
    const double* Tb = T + (size_t)b * nh * nh;
    double* mcol = ts.mcol + (size_t)cidx * nh * nh;
    double* R = ts.r + (size_t)cidx * nk * nh;
    double* M2 = ts.m2 + (size_t)cidx * nk * nh;
    double* TW = ts.tw + (size_t)cidx * nk * ntot;
    if (tid == 0) nak_uniform_factors(nh, fac);
    __syncthreads();
    for (int m = tid; m < nh; m += blockDim.x) nak_uniform_solve(nh, hT, Tb + m, nh, mcol + m, nh, fac);
    __syncthreads();
    for (int idx = tid; idx < nk * nh; idx += blockDim.x) {
        const int i = idx / nh, m = idx - i * nh;
        const double x = (i == nk - 1) ? l1 : l0 + hA * i;
        int j = (int)floor((x - l0) / hT);
        j = j < 0 ? 0 : (j > nh - 2 ? nh - 2 : j);
        const double t = (x - (l0 + hT * j)) / hT;
        R[idx] = nak_eval(Tb[(size_t)j * nh + m], Tb[(size_t)(j + 1) * nh + m], mcol[(size_t)j * nh + m],
                          mcol[(size_t)(j + 1) * nh + m], hT, t);
    }
    __syncthreads();
    for (int i = tid; i < nk; i += blockDim.x) nak_uniform_solve(nh, hT, R + (size_t)i * nh, 1, M2 + (size_t)i * nh, 1, fac);
    __syncthreads();
    const double D = d_ng[b];
    const double inv_d4 = 1.0 / (D * D * D * D);                  // covariance.py:651-652
    for (int idx = tid; idx < nk * ntot; idx += blockDim.x) {
        const int i = idx / ntot, qq = idx - i * ntot;
        const int j = qq / (sub * nq), r = qq - j * (sub * nq);
        const int sidx = r / nq, q = r - sidx * nq;
        const double a = l0 + hT * j, bb = (j == nh - 2) ? l1 : l0 + hT * (j + 1);
        const double pa = a + (bb - a) * sidx / sub, pb = (sidx == sub - 1) ? bb : a + (bb - a) * (sidx + 1) / sub;
        const double half = 0.5 * (pb - pa);
        const double x = 0.5 * (pa + pb) + half * c_glx[nq][q];
        const double t = (x - a) / hT;
        const double* Ri = R + (size_t)i * nh;
        const double* Mi = M2 + (size_t)i * nh;
        double v = nak_eval(Ri[j], Ri[j + 1], Mi[j], Mi[j + 1], hT, t);
        // exp(ln k_max) rounds above k_max in the reference's arithmetic: its last k_a node sees
        // T = 0 (halo_trispectrum.py:100-107); the host passes that comparison's outcome
        if (cp.zero_last_ka && i == nk - 1) v = 0.0;
        TW[idx] = v * half * c_glw[nq][q] * exp(2.0 * x) * inv_d4;
    }
}

// ---------------------------------------------------------------------------------------
// non-Gaussian term.  grid (n_bins, chunk): CTA (a, point) handles the bin pairs (a, b >= a).
// covariance.py:593-683: I_i = int dln k_b k_b^2 T(k_a,i, k_b) K_NG(ln k_a,i theta_a, ln k_b theta_b) / D_NG^4
// at the n_kernel ln k_a nodes, then int dln k_a k_a^2 spline(I) / (4 pi^2 A).
// dynamic shared memory: 2 n_kernel^2 + n_q + n_bins n_kernel * 2 + n_kernel doubles
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(COV_THREADS)
cov_ng_kernel(const Cfg cfg, const CovP cp, int b0, int nb_chunk, const double* __restrict__ bin_center,
              const double* __restrict__ tw, CovOut out) {
    extern __shared__ double dyn[];
    const int cidx = blockIdx.y;
    if (cidx >= nb_chunk) return;
    const int b = b0 + cidx;
    const int a_bin = blockIdx.x;
    const int nb = cp.n_bins, nk = cfg.n_kernel, nh = cfg.n_halo, nq = cov_ng_order(cfg, cp);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    const int sub = hankel_subdiv(cfg), ntot = (nh - 1) * sub * nq;
    double* U = dyn;                      // [nk (i), nk (m)]
    double* Mu = U + nk * nk;             // [nk, nk]
    double* xq = Mu + nk * nk;            // [ntot]
    double* I = xq + ntot;                // [nb, nk]
    double* MI = I + nb * nk;             // [nb, nk]
    double* fac = MI + nb * nk;           // [nk]
    __shared__ int rowzero[COV_MAX_COLS];
    const double x0 = log(cp.theta_min_rad * cfg.k_min), x1 = log(cp.theta_max_rad * cfg.k_max), hx = (x1 - x0) / (nk - 1);
    const double l0 = log(cfg.k_min), l1 = log(cfg.k_max), hT = (l1 - l0) / (nh - 1), hA = (l1 - l0) / (nk - 1);
    const double* L = out.lkng + (size_t)b * nk * nk;
    const double* M = out.mkng + (size_t)b * nk * nk;
    const double kmin10 = 10.0 * out.kng_min[b];
    const double lta = log(bin_center[a_bin]);
    if (tid == 0) nak_uniform_factors(nk, fac);
    for (int qq = tid; qq < ntot; qq += blockDim.x) {
        const int j = qq / (sub * nq), r = qq - j * (sub * nq);
        const int sidx = r / nq, q = r - sidx * nq;
        const double a = l0 + hT * j, bb = (j == nh - 2) ? l1 : l0 + hT * (j + 1);
        const double pa = a + (bb - a) * sidx / sub, pb = (sidx == sub - 1) ? bb : a + (bb - a) * (sidx + 1) / sub;
        xq[qq] = 0.5 * (pa + pb) + 0.5 * (pb - pa) * c_glx[nq][q];
    }
    // the column splines of ln(K - 10 K_min) at x = ln(k_a,i theta_a): a table over the second axis
    for (int idx = tid; idx < nk * nk; idx += blockDim.x) {
        const int i = idx / nk, m = idx - i * nk;
        double x = ((i == nk - 1) ? l1 : l0 + hA * i) + lta;
        if (m == 0) rowzero[i] = !(x <= x1);
        if (x < x0) x = x0;                                       // kernel.py:997-999
        if (x > x1) x = x1;
        int j = (int)floor((x - x0) / hx);
        j = j < 0 ? 0 : (j > nk - 2 ? nk - 2 : j);
        const double t = (x - (x0 + hx * j)) / hx;
        U[idx] = nak_eval(L[(size_t)j * nk + m], L[(size_t)(j + 1) * nk + m], M[(size_t)j * nk + m], M[(size_t)(j + 1) * nk + m], hx, t);
    }
    __syncthreads();
    for (int i = tid; i < nk; i += blockDim.x) nak_uniform_solve(nk, hx, U + i * nk, 1, Mu + i * nk, 1, fac);
    __syncthreads();
    const double* TW = tw + (size_t)cidx * nk * ntot;
    const int n_task = (nb - a_bin) * nk;
    for (int task = wid; task < n_task; task += nwarp) {
        const int bb = a_bin + task / nk, i = task - (task / nk) * nk;
        double acc = 0.0;
        if (!rowzero[i]) {
            const double ltb = log(bin_center[bb]);
            const double* Ui = U + i * nk;
            const double* Mi = Mu + i * nk;
            const double* Ti = TW + (size_t)i * ntot;
            const double ihx = 1.0 / hx;
            for (int qq = lane; qq < ntot; qq += 32) {
                double y = xq[qq] + ltb;
                if (y <= x1) {                                    // kernel.py:1004-1008
                    if (y < x0) y = x0;
                    int j = (int)((y - x0) * ihx);
                    j = j < 0 ? 0 : (j > nk - 2 ? nk - 2 : j);
                    const double t = (y - (x0 + hx * j)) * ihx;
                    // ln(K - 10 K_min) is a table value of moderate size: the constant-bank exp applies
                    const double v = exp_fast(nak_eval(Ui[j], Ui[j + 1], Mi[j], Mi[j + 1], hx, t)) + kmin10;
                    acc = fma(Ti[qq], v, acc);
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) I[bb * nk + i] = acc;
    }
    __syncthreads();
    // outer integral: not-a-knot spline through I over ln k_a, GL-8 on every interval (covariance.py:607-613)
    double* NG = out.parts + ((size_t)b * 3 + 2) * nb * nb;
    // one warp per theta_b: lane 0 solves for the spline, all lanes share the (interval, node) pairs -- as a loop of one
    // THREAD per theta_b this tail ran 400 exponentials in sequence on 30 threads while the rest of the CTA idled
    for (int bb = a_bin + wid; bb < nb; bb += nwarp) {
        const double* Ib = I + bb * nk;
        double* Mb = MI + bb * nk;
        if (lane == 0) nak_uniform_solve(nk, hA, Ib, 1, Mb, 1, fac);
        __syncwarp();
        double tot = 0.0;
        for (int idx = lane; idx < (nk - 1) * 8; idx += 32) {
            const int j = idx >> 3, q = idx & 7;
            const double xa = l0 + hA * j, xb = (j == nk - 2) ? l1 : l0 + hA * (j + 1), half = 0.5 * (xb - xa);
            const double x = 0.5 * (xa + xb) + half * c_glx[8][q];
            tot += half * c_glw[8][q] * exp(2.0 * x) * nak_eval(Ib[j], Ib[j + 1], Mb[j], Mb[j + 1], hA, (x - xa) / hA);
        }
        tot = warp_sum(tot);
        if (lane == 0) {
            const double v = tot / (4.0 * M_PI * M_PI * cp.area_sr);
            NG[(size_t)a_bin * nb + bb] = v;
            NG[(size_t)bb * nb + a_bin] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Non-Gaussian term on a theta-shift-aligned k_b grid (log-spaced bins only).
//
// The inner integral  I_i(theta_b) = int dln k_b  k_b^2 T(k_a,i, k_b) K_NG(ln k_a,i theta_a, ln k_b theta_b) / D_NG^4
// sees theta_b only through the shift ln theta_b of the kernel's second argument.  For the bins Covariance builds
// (covariance.py:53-74) ln theta_b = ln theta_0 + b Delta exactly, so on k_b pieces of width delta = Delta / m the
// kernel values of bin b + 1 are those of bin b four pieces further on: V_i[s] = K_NG(x_i, l0 + ln theta_0 + delta (s +
// node)) is tabulated ONCE per (theta_a, k_a node) -- (P + m (n_bins - 1)) nq exponentials instead of n_bins P nq --
// and every I_i(theta_b) is a dot product of the T k_b^2 weights with a window of that table (a Toeplitz product).
// The last, partial piece [l0 + delta P, l1] is not on the grid and is evaluated directly.
// ---------------------------------------------------------------------------------------
struct NgGrid {
    int m;          // pieces per bin spacing
    int P;          // full pieces of width delta in [l0, l1]
    int has_rem;    // a partial piece [l0 + delta P, l1] follows
    double delta, Delta, lt0;
};
// bins log-spaced to rounding?  (host and device: the host decides which kernel runs)
__host__ __device__ inline bool ng_grid_make(const Cfg& cfg, int n_bins, double lt0, double lt_last, NgGrid& g) {
    if (n_bins < 2) return false;
    g.Delta = (lt_last - lt0) / (n_bins - 1);
    if (!(g.Delta > 1e-6)) return false;
    g.m = (int)ceil(g.Delta / COV_PIECE - 1e-9);
    if (g.m < 1) g.m = 1;
    g.delta = g.Delta / g.m;
    const double len = log(cfg.k_max) - log(cfg.k_min);
    g.P = (int)floor(len / g.delta + 1e-9);
    g.has_rem = (len - g.delta * g.P) > 1e-9 * len;
    g.lt0 = lt0;
    return true;
}
__host__ __device__ inline int ng_grid_nodes(const NgGrid& g, int nq) { return (g.P + (g.has_rem ? 1 : 0)) * nq; }
__host__ __device__ inline int ng_grid_vlen(const NgGrid& g, int n_bins, int nq) { return (g.P + g.m * (n_bins - 1)) * nq; }

// T(k_a node i, k_b node) k_b^2 w / D_NG^4 on the shift-aligned grid: as cov_tri_nodes_kernel, other nodes
__global__ void __launch_bounds__(COV_THREADS)
cov_tri_nodes_shift_kernel(const Cfg cfg, const CovP cp, NgGrid G, int b0, int nb_chunk, const double* __restrict__ T,
                           const double* __restrict__ d_ng, TriScratch ts) {
    __shared__ double fac[1024];
    const int cidx = blockIdx.x;
    if (cidx >= nb_chunk) return;
    const int b = b0 + cidx;
    const int nh = cfg.n_halo, nk = cfg.n_kernel, nq = cov_ng_order(cfg, cp), tid = threadIdx.x;
    const int ntot = ng_grid_nodes(G, nq);
    const double l0 = log(cfg.k_min), l1 = log(cfg.k_max), hT = (l1 - l0) / (nh - 1), hA = (l1 - l0) / (nk - 1);
    const double* Tb = T + (size_t)b * nh * nh;
    double* mcol = ts.mcol + (size_t)cidx * nh * nh;
    double* R = ts.r + (size_t)cidx * nk * nh;
    double* M2 = ts.m2 + (size_t)cidx * nk * nh;
    double* TW = ts.tw + (size_t)cidx * nk * ntot;
    if (tid == 0) nak_uniform_factors(nh, fac);
    __syncthreads();
    for (int m = tid; m < nh; m += blockDim.x) nak_uniform_solve(nh, hT, Tb + m, nh, mcol + m, nh, fac);
    __syncthreads();
    for (int idx = tid; idx < nk * nh; idx += blockDim.x) {
        const int i = idx / nh, m = idx - i * nh;
        const double x = (i == nk - 1) ? l1 : l0 + hA * i;
        int j = (int)floor((x - l0) / hT);
        j = j < 0 ? 0 : (j > nh - 2 ? nh - 2 : j);
        const double t = (x - (l0 + hT * j)) / hT;
        R[idx] = nak_eval(Tb[(size_t)j * nh + m], Tb[(size_t)(j + 1) * nh + m], mcol[(size_t)j * nh + m],
                          mcol[(size_t)(j + 1) * nh + m], hT, t);
    }
    __syncthreads();
    for (int i = tid; i < nk; i += blockDim.x) nak_uniform_solve(nh, hT, R + (size_t)i * nh, 1, M2 + (size_t)i * nh, 1, fac);
    __syncthreads();
    const double D = d_ng[b];
    const double inv_d4 = 1.0 / (D * D * D * D);                  // covariance.py:651-652
    for (int idx = tid; idx < nk * ntot; idx += blockDim.x) {
        const int i = idx / ntot, qq = idx - i * ntot;
        const int p = qq / nq, q = qq - p * nq;
        const double pa = l0 + G.delta * p, pb = (p < G.P) ? l0 + G.delta * (p + 1) : l1;
        const double half = 0.5 * (pb - pa);
        const double x = 0.5 * (pa + pb) + half * c_glx[nq][q];
        int j = (int)floor((x - l0) / hT);
        j = j < 0 ? 0 : (j > nh - 2 ? nh - 2 : j);
        const double t = (x - (l0 + hT * j)) / hT;
        const double* Ri = R + (size_t)i * nh;
        const double* Mi = M2 + (size_t)i * nh;
        double v = nak_eval(Ri[j], Ri[j + 1], Mi[j], Mi[j + 1], hT, t);
        if (cp.zero_last_ka && i == nk - 1) v = 0.0;              // halo_trispectrum.py:100-107, see cov_tri_nodes_kernel
        TW[idx] = v * half * c_glw[nq][q] * exp(2.0 * x) * inv_d4;
    }
}

// grid (n_bins, chunk), 256 threads: CTA (a, point) handles the bin pairs (a, b >= a).
// dynamic shared memory: 2 n_kernel^2 + vlen + nodes + 2 n_bins n_kernel + n_kernel doubles
__global__ void __launch_bounds__(COV_THREADS)
cov_ng_shift_kernel(const Cfg cfg, const CovP cp, NgGrid G, int b0, int nb_chunk, const double* __restrict__ bin_center,
                    const double* __restrict__ tw, CovOut out) {
    extern __shared__ double dyn[];
    const int cidx = blockIdx.y;
    if (cidx >= nb_chunk) return;
    const int b = b0 + cidx;
    const int a_bin = blockIdx.x;
    const int nb = cp.n_bins, nk = cfg.n_kernel, nq = cov_ng_order(cfg, cp);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    const int ntot = ng_grid_nodes(G, nq), nfull = G.P * nq, vlen = ng_grid_vlen(G, nb, nq);
    double* U = dyn;                      // [nk (i), nk (m)]
    double* Mu = U + nk * nk;             // [nk, nk]
    double* Vall = Mu + nk * nk;          // [vlen + ntot]: kernel values, then the T k_b^2 weights of the current k_a node
    double* I = Vall + (size_t)(vlen + ntot);   // [nb, nk]
    double* MI = I + nb * nk;             // [nb, nk]
    double* fac = MI + nb * nk;           // [nk]
    __shared__ int rowzero[COV_MAX_COLS];
    const double x0 = log(cp.theta_min_rad * cfg.k_min), x1 = log(cp.theta_max_rad * cfg.k_max), hx = (x1 - x0) / (nk - 1);
    const double l0 = log(cfg.k_min), l1 = log(cfg.k_max), hA = (l1 - l0) / (nk - 1);
    const double* L = out.lkng + (size_t)b * nk * nk;
    const double* M = out.mkng + (size_t)b * nk * nk;
    const double kmin10 = 10.0 * out.kng_min[b];
    const double lta = log(bin_center[a_bin]);
    if (tid == 0) nak_uniform_factors(nk, fac);
    // the column splines of ln(K - 10 K_min) at x = ln(k_a,i theta_a): a table over the second axis
    for (int idx = tid; idx < nk * nk; idx += blockDim.x) {
        const int i = idx / nk, m = idx - i * nk;
        double x = ((i == nk - 1) ? l1 : l0 + hA * i) + lta;
        if (m == 0) rowzero[i] = !(x <= x1);
        if (x < x0) x = x0;                                       // kernel.py:997-999
        if (x > x1) x = x1;
        int j = (int)floor((x - x0) / hx);
        j = j < 0 ? 0 : (j > nk - 2 ? nk - 2 : j);
        const double t = (x - (x0 + hx * j)) / hx;
        U[idx] = nak_eval(L[(size_t)j * nk + m], L[(size_t)(j + 1) * nk + m], M[(size_t)j * nk + m], M[(size_t)(j + 1) * nk + m], hx, t);
    }
    __syncthreads();
    for (int i = tid; i < nk; i += blockDim.x) nak_uniform_solve(nk, hx, U + i * nk, 1, Mu + i * nk, 1, fac);
    __syncthreads();
    const double* TW = tw + (size_t)cidx * nk * ntot;
    const double ihx = 1.0 / hx;
    double* V = Vall;                       // kernel values of the current k_a node on the shared grid
    double* Tsm = V + vlen;                 // its T k_b^2 weights
    const double ybase = l0 + G.lt0;
    // One k_a node at a time for the whole CTA: all threads tabulate the kernel values and stage the weights, then the
    // warps share the theta_b dot products.  (One node per WARP needed 8 x 16 KB of tables: one CTA per SM, 22 % issue
    // utilisation -- ncu r3c.)
    for (int i = 0; i < nk; ++i) {
        if (rowzero[i]) {                                         // the same for every thread of the CTA
            for (int bb = a_bin + tid; bb < nb; bb += blockDim.x) I[bb * nk + i] = 0.0;
            continue;
        }
        const double* Ui = U + i * nk;
        const double* Mi = Mu + i * nk;
        // kernel values on the shared grid (kernel.py:993-1014: clamp below, zero above)
        for (int idx = tid; idx < vlen; idx += blockDim.x) {
            const int s = idx / nq, q = idx - s * nq;
            double y = ybase + G.delta * (s + 0.5 + 0.5 * c_glx[nq][q]);
            double v = 0.0;
            if (y <= x1) {
                if (y < x0) y = x0;
                int j = (int)((y - x0) * ihx);
                j = j < 0 ? 0 : (j > nk - 2 ? nk - 2 : j);
                const double t = (y - (x0 + hx * j)) * ihx;
                v = exp_fast(nak_eval(Ui[j], Ui[j + 1], Mi[j], Mi[j + 1], hx, t)) + kmin10;
            }
            V[idx] = v;
        }
        for (int idx = tid; idx < ntot; idx += blockDim.x) Tsm[idx] = TW[(size_t)i * ntot + idx];
        __syncthreads();
        for (int bb = a_bin + wid; bb < nb; bb += nwarp) {
            const double* Vb = V + (size_t)G.m * bb * nq;
            double acc = 0.0;
            for (int idx = lane; idx < nfull; idx += 32) acc = fma(Tsm[idx], Vb[idx], acc);
            if (G.has_rem && lane < nq) {                          // the partial last piece, off the grid
                const double pa = l0 + G.delta * G.P, half = 0.5 * (l1 - pa);
                double y = pa + half + half * c_glx[nq][lane] + (G.lt0 + G.Delta * bb);
                if (y <= x1) {
                    if (y < x0) y = x0;
                    int j = (int)((y - x0) * ihx);
                    j = j < 0 ? 0 : (j > nk - 2 ? nk - 2 : j);
                    const double t = (y - (x0 + hx * j)) * ihx;
                    acc = fma(Tsm[nfull + lane], exp_fast(nak_eval(Ui[j], Ui[j + 1], Mi[j], Mi[j + 1], hx, t)) + kmin10, acc);
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) I[bb * nk + i] = acc;
        }
        __syncthreads();
    }
    __syncthreads();
    // outer integral: not-a-knot spline through I over ln k_a, GL-8 on every interval (covariance.py:607-613)
    double* NG = out.parts + ((size_t)b * 3 + 2) * nb * nb;
    // one warp per theta_b: lane 0 solves for the spline, all lanes share the (interval, node) pairs -- as a loop of one
    // THREAD per theta_b this tail ran 400 exponentials in sequence on 30 threads while the rest of the CTA idled
    for (int bb = a_bin + wid; bb < nb; bb += nwarp) {
        const double* Ib = I + bb * nk;
        double* Mb = MI + bb * nk;
        if (lane == 0) nak_uniform_solve(nk, hA, Ib, 1, Mb, 1, fac);
        __syncwarp();
        double tot = 0.0;
        for (int idx = lane; idx < (nk - 1) * 8; idx += 32) {
            const int j = idx >> 3, q = idx & 7;
            const double xa = l0 + hA * j, xb = (j == nk - 2) ? l1 : l0 + hA * (j + 1), half = 0.5 * (xb - xa);
            const double x = 0.5 * (xa + xb) + half * c_glx[8][q];
            tot += half * c_glw[8][q] * exp(2.0 * x) * nak_eval(Ib[j], Ib[j + 1], Mb[j], Mb[j + 1], hA, (x - xa) / hA);
        }
        tot = warp_sum(tot);
        if (lane == 0) {
            const double v = tot / (4.0 * M_PI * M_PI * cp.area_sr);
            NG[(size_t)a_bin * nb + bb] = v;
            NG[(size_t)bb * nb + a_bin] = v;
        }
    }
}

// Poisson term (covariance.py:323-357) and assembly (covariance.py:276-321)
__global__ void cov_finish_kernel(const CovP cp, int B, const double* __restrict__ bin_center,
                                  const double* __restrict__ bin_delta, CovOut out, double* __restrict__ cov,
                                  int32_t* __restrict__ status) {
    const int nb = cp.n_bins;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * nb * nb) return;
    const int b = (int)(idx / ((size_t)nb * nb));
    const int r = (int)(idx - (size_t)b * nb * nb), i = r / nb, j = r - i * nb;
    double* P = out.parts + ((size_t)b * 3) * nb * nb;
    const double* G = P + (size_t)nb * nb;
    const double* NG = G + (size_t)nb * nb;
    double p = 0.0;
    if (i == j) {
        const double t1 = cp.poisson[0] * cp.poisson[2] * cp.shot_wt[0];
        const double t2 = cp.poisson[3] * cp.poisson[1] * cp.shot_wt[1];
        const double t3 = cp.poisson[4] * cp.poisson[5] * cp.shot_wt[1];
        p = (t1 + t2 + t3) / (2.0 * M_PI * cp.area_sr * bin_center[i] * bin_delta[i]);
    }
    P[r] = p;
    double v = p;
    if (!cp.poisson_only) {
        v += G[r];
        if (cp.nongaussian) v += NG[r];
    }
    cov[idx] = v;
    if (!isfinite(v) && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
}

}  // namespace chomp
