// Stage 1: comoving-distance / growth tables, redshift windows, z_bar and the Limber
// kernel table K(ln k theta).  One CTA (128 threads) per parameter point.
//
// Replaces (reference): MultiEpoch._initialize_splines cosmology.py:787-817 and accessors
// :873-953; dNdz.normalize kernel.py:43-54; dNdzGaussian :89-112; dNdzMagLim :148-179;
// WindowFunction.set_cosmology_object / _initialize_spline / window_function :289-340;
// WindowFunctionGalaxy.raw_window_function :382-387; WindowFunctionConvergence
// .raw_window_function / _lensing_integrand :443-484; Kernel.__init__ / _find_z_bar /
// raw_kernel / _kernel_integrand / _initialize_spline :584-712; GalaxyGalaxyLensingKernel
// :784-839.
//
// Where the reference runs an adaptive Romberg per table node, the device integrates
// panel-wise between the knots of the splines that make up each integrand (those are the
// only places the integrands are not smooth) with fixed Gauss-Legendre orders.
#pragma once
#include "common.cuh"
#include "special.cuh"
#include "spline.cuh"

namespace chomp {

#ifndef LIMBER_THREADS
#define LIMBER_THREADS 128
#endif
#define DNDZ_PANELS 16

struct EpochGrid {     // one MultiEpoch tabulation (cosmology.py:747-817)
    int n;
    double z_min, z_max;
    double *z, *chi, *growth;           // nodes
    double *c_chi_z, *c_z_chi, *c_g_z;  // spline coefficients
};

// MultiEpoch.comoving_distance (cosmology.py:873-893)
__device__ __forceinline__ double grid_chi(const EpochGrid& g, double z) {
    if (!(z <= g.z_max && z >= g.z_min)) return 0.0;
    return spline_eval_uniform(g.c_chi_z, z, g.z_min, (g.z_max - g.z_min) / (g.n - 1), g.n);
}
// MultiEpoch.redshift (cosmology.py:922-932)
__device__ __forceinline__ double grid_z(const EpochGrid& g, double chi) {
    return spline_eval_search(g.c_z_chi, chi, g.chi, g.n);
}
// MultiEpoch.growth_factor (cosmology.py:934-953)
__device__ __forceinline__ double grid_growth(const EpochGrid& g, double z) {
    if (!(z <= g.z_max && z >= g.z_min)) return 1.0;
    return spline_eval_uniform(g.c_g_z, z, g.z_min, (g.z_max - g.z_min) / (g.n - 1), g.n);
}

struct Dndz {
    int kind;
    double z_min, z_max, p0, p1, p2, norm;
    const double* tab;      // CHOMP_DNDZ_TABLE: breaks[n + 1] then coef[n][4] (global memory)
    int n;
};
__device__ __forceinline__ Dndz make_dndz(const Cfg& cfg, int i, double norm) {
    return Dndz{cfg.dndz_kind[i], cfg.dndz_zmin[i], cfg.dndz_zmax[i], cfg.dndz_p[i][0], cfg.dndz_p[i][1],
                cfg.dndz_p[i][2], norm, cfg.dndz_table[i], cfg.dndz_table_n[i]};
}
// dNdzInterpolation.raw_dndz (kernel.py:207-208): the piece that holds z, end pieces extrapolating
__device__ __noinline__ double dndz_table_eval(const double* __restrict__ tab, int n, double z) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(tab + mid) <= z) lo = mid; else hi = mid;
    }
    const double t = z - __ldg(tab + lo);
    const double* c = tab + (n + 1) + 4 * lo;
    return fma(t, fma(t, fma(t, __ldg(c + 3), __ldg(c + 2)), __ldg(c + 1)), __ldg(c));
}
// integral of piece i over its intersection with [a, b]
__device__ __forceinline__ double dndz_table_piece(const double* __restrict__ tab, int n, int i, double a, double b) {
    const double x0 = tab[i], lo = fmax(a, x0) - x0, hi = fmin(b, tab[i + 1]) - x0;
    if (!(hi > lo)) return 0.0;
    const double* c = tab + (n + 1) + 4 * i;
    auto F = [&](double t) { return t * (c[0] + t * (c[1] * 0.5 + t * (c[2] * (1.0 / 3.0) + t * (c[3] * 0.25)))); };
    return F(hi) - F(lo);
}
// out of line: the pow / exp of the analytic forms are inlined once, not at every call site
__device__ __noinline__ double dndz_raw(const Dndz& d, double z) {
    if (d.kind == CHOMP_DNDZ_TABLE) return dndz_table_eval(d.tab, d.n, z);
    if (d.kind == CHOMP_DNDZ_GAUSSIAN)                          // kernel.py:110-112
        return exp(-1.0 * (z - d.p0) * (z - d.p0) / (2.0 * d.p1 * d.p1));
    return pow(z, d.p0) * exp(-1.0 * pow(z / d.p1, d.p2));      // kernel.py:177-179
}
__device__ __forceinline__ double dndz_eval(const Dndz& d, double z) {   // kernel.py:67-86
    return (z <= d.z_max && z >= d.z_min) ? d.norm * dndz_raw(d, z) : 0.0;
}

struct Window {
    int n;
    double chi_min, chi_max;
    double *wf, *coef;
};
// WindowFunction.window_function (kernel.py:326-340)
__device__ __forceinline__ double window_eval(const Window& w, double chi) {
    if (!(chi >= w.chi_min && chi <= w.chi_max)) return 0.0;
    return spline_eval_uniform(w.coef, chi, w.chi_min, (w.chi_max - w.chi_min) / (w.n - 1), w.n);
}

struct LimberF {   // W_a W_b D^2 of Kernel._kernel_integrand (kernel.py:707-712) without the Bessel factor
    Window a, b;
    EpochGrid g;
    __device__ __forceinline__ double operator()(double chi) const {
        const double D = grid_growth(g, grid_z(g, chi));
        return window_eval(a, chi) * window_eval(b, chi) * D * D;
    }
};

struct LimberOut {
    double *zbar, *dbar;       // [B]
    double *knodes;            // [B, n_kernel]
    double *kcoef;             // [B, 4 n_kernel]
    double *chi_nodes;         // [B, 3, n_cosmo]
    double *win_nodes;         // [B, 2, n_window]
    double *win_chi;           // [B, 4]  chi_min_a, chi_max_a, chi_min_b, chi_max_b
    double *win_coef;          // [B, 2, 4 n_window]
    double *kchi;              // [B, 2]  kernel chi_min, chi_max
    double *grid0;             // [B, 13 n_cosmo]  chi nodes + chi(z), z(chi), D(z) coefficients of the kernel's MultiEpoch
    double *dndz_norm;         // [B, 2]
    double *edges;             // [B, 2 n_window + n_cosmo + 4] base panel edges of the chi integrals
    int32_t *n_edges;          // [B]
};

__host__ __device__ inline size_t limber_work_doubles(const Cfg& cfg) {
    // warp-built splines (5 n scratch each): one table spline per warp at a time, then 2 window
    // splines (+ n abscissae each), then the K spline (+ n)
    size_t a = 5 * (size_t)cfg.n_cosmo * (LIMBER_THREADS / 32), b = 12 * (size_t)cfg.n_window, c = 6 * (size_t)cfg.n_kernel;
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
}
__host__ __device__ inline size_t limber_edge_cap(const Cfg& cfg, int same_window) {
    return (same_window ? 1 : 2) * (size_t)cfg.n_window + cfg.n_cosmo + 4;     // max base panels + 1
}
__host__ __device__ inline size_t limber_smem_doubles(const Cfg& cfg, int same_window) {
    const size_t nz = cfg.n_cosmo, nw = cfg.n_window, nk = cfg.n_kernel;
    const size_t nb = limber_edge_cap(cfg, same_window);
    const size_t n_grid = same_window ? 2 : 3, n_win = same_window ? 1 : 2;
    return n_grid * (3 * nz + 12 * nz) + n_win * (nw + 4 * nw) + 2 * nz /*lens sums*/ + nb /*edges*/ +
           2 * nb * cfg.nq_limber /*chi_q, Fw_q*/ + nk + 4 * nk + limber_work_doubles(cfg) +
           128 /*red + misc*/ + (LIMBER_THREADS / 32) * (nb / 2 + 2) /*per-warp int prefix sums*/;
}

#define LIMBER_SER_TERMS 12
// (-1)^m / (m! (m + n)!) for n = 0, 2: coefficients of (k theta chi / 2)^(2m + n) in J_n
__constant__ double k_limber_ser[2][LIMBER_SER_TERMS];
static inline cudaError_t chomp_upload_limber_tables() {
    double t[2][LIMBER_SER_TERMS];
    for (int o = 0; o < 2; ++o) {
        const int n = 2 * o;
        double fm = 1.0, fmn = (n == 0) ? 1.0 : 2.0;      // m!, (m + n)!
        for (int m = 0; m < LIMBER_SER_TERMS; ++m) {
            if (m > 0) { fm *= m; fmn *= (m + n); }
            t[o][m] = ((m & 1) ? -1.0 : 1.0) / (fm * fmn);
        }
    }
    return cudaMemcpyToSymbol(k_limber_ser, t, sizeof t);
}

// number of elements of the increasing sequence s[0..n) that are < x (strict) or <= x
__device__ __forceinline__ int count_below(const double* __restrict__ s, int n, double x, bool or_equal) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const bool below = or_equal ? (s[mid] <= x) : (s[mid] < x);
        if (below) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// the same for the uniform sequence a0 + h i (evaluated exactly as the callers evaluate it)
__device__ __forceinline__ int count_below_uniform(double a0, double h, int n, double x, bool or_equal) {
    int c = (int)ceil((x - a0) / h);
    c = c < 0 ? 0 : (c > n ? n : c);
    while (c > 0 && !(or_equal ? (a0 + h * (c - 1) <= x) : (a0 + h * (c - 1) < x))) --c;
    while (c < n && (or_equal ? (a0 + h * c <= x) : (a0 + h * c < x))) ++c;
    return c;
}
#ifndef LIMBER_MIN_BLOCKS
#define LIMBER_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(LIMBER_THREADS, LIMBER_MIN_BLOCKS)
limber_tables_kernel(const Cfg cfg, int B, int same_window, const double* __restrict__ cosmo, LimberOut out,
                     int32_t* __restrict__ status) {
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    if (b >= B) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    const int nz = cfg.n_cosmo, nw = cfg.n_window, nk = cfg.n_kernel, nq = cfg.nq_limber, nql = cfg.nq_lens;
    const double eps = cfg.window_precision;
    // ---- carve shared memory ------------------------------------------------------------------
    double* p = sm;
    const int n_grid = same_window ? 2 : 3;
    EpochGrid g[3];
    for (int i = 0; i < n_grid; ++i) {
        g[i].n = nz;
        g[i].z = p; p += nz; g[i].chi = p; p += nz; g[i].growth = p; p += nz;
        g[i].c_chi_z = p; p += 4 * nz; g[i].c_z_chi = p; p += 4 * nz; g[i].c_g_z = p; p += 4 * nz;
    }
    if (same_window) g[2] = g[1];          // one window: its grid and table serve as both
    Window win[2];
    for (int i = 0; i < (same_window ? 1 : 2); ++i) { win[i].n = nw; win[i].wf = p; p += nw; win[i].coef = p; p += 4 * nw; }
    if (same_window) win[1] = win[0];
    double* lens0 = p; p += nz;       // suffix sums of  w f          over the window-cosmology panels
    double* lens1 = p; p += nz;       //                 w f / chi'
    const int nb_max = (int)limber_edge_cap(cfg, same_window);
    const int edge_stride = 2 * nw + nz + 4;      // row length of the global copy
    double* edge = p; p += nb_max;
    double* chi_q = p; p += (size_t)nb_max * nq;
    double* fw_q = p; p += (size_t)nb_max * nq;
    double* kn = p; p += nk;
    double* kc = p; p += 4 * nk;
    double* work = p; p += limber_work_doubles(cfg);
    double* red = p; p += 128;
    int* pfx_all = (int*)p;           // (LIMBER_THREADS / 32) x (nb_max + 1) ints
    __shared__ int n_edge_s;
    __shared__ double s_misc[8];

    const Cosmo c = load_cosmo(cosmo + (size_t)b * CHOMP_N_COSMO, cfg.cosmo_precision);
    // ---- redshift ranges of the three tabulations ------------------------------------------
    // [0] the MultiEpoch handed to Kernel; [1], [2] the windows' own copies, re-gridded on
    // their z range (kernel.py:236-240, 296-297, 372-375, 434-436)
    g[0].z_min = cfg.zk_min < 0.0 ? 0.0 : cfg.zk_min; g[0].z_max = cfg.zk_max;
    Dndz dist[2];
    for (int i = 0; i < 2; ++i) {
        dist[i] = make_dndz(cfg, i, 1.0);
        double zlo = (cfg.window_kind[i] == CHOMP_WINDOW_GALAXY) ? dist[i].z_min : 0.0;
        if (zlo < eps) zlo = eps;
        g[1 + i].z_min = zlo; g[1 + i].z_max = dist[i].z_max;
    }
    // ---- chi(z), D(z) nodes: chi_i = chi_{i-1} + GL-8 over [z_{i-1}, z_i] ---------------------
    // One (grid, node, Gauss-Legendre index) item per thread and round, the eight lanes of a node
    // reduced by shuffles; the first node integrates [0, z_min] on four panels.
    const double g1 = growth_approx(c, 1.0);
    for (int base = 0; base < n_grid * nz * 8; base += blockDim.x) {
        const int idx = base + tid;
        const bool live = idx < n_grid * nz * 8;
        const int node = live ? idx >> 3 : 0, q = idx & 7;
        const int gi = node / nz, i = node - gi * nz;
        // (no dynamic indexing of g[]: that would put the array in local memory)
        const double gz_min = gi == 0 ? g[0].z_min : (gi == 1 ? g[1].z_min : g[2].z_min);
        const double gz_max = gi == 0 ? g[0].z_max : (gi == 1 ? g[1].z_max : g[2].z_max);
        double* gbase = sm + (size_t)gi * 15 * nz;          // z | chi | growth | coefficients
        const double hz = (gz_max - gz_min) / (nz - 1);
        const double zi = (i == nz - 1) ? gz_max : gz_min + hz * i;
        double acc = 0.0;
        if (i == 0) {
            if (zi > 0.0)
                for (int pnl = 0; pnl < 4; ++pnl) {
                    const double a = zi * pnl / 4.0, bb = zi * (pnl + 1) / 4.0, half = 0.5 * (bb - a);
                    acc += half * c_glw[8][q] * inv_hubble(c, 0.5 * (a + bb) + half * c_glx[8][q]);
                }
        } else {
            const double a = gz_min + hz * (i - 1), half = 0.5 * (zi - a);
            acc = half * c_glw[8][q] * inv_hubble(c, 0.5 * (a + zi) + half * c_glx[8][q]);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (live && q == 0) {
            gbase[i] = zi;
            gbase[2 * nz + i] = growth_approx(c, 1.0 / (1.0 + zi)) / g1;
            gbase[nz + i] = acc;
        }
    }
    __syncthreads();
    if (wid < n_grid) {      // running sums chi_i = sum_{j <= i} (panel integrals): one warp scan per grid
        EpochGrid& G = g[wid];
        double carry = 0.0;
        for (int base = 0; base < nz; base += 32) {
            const int i = base + lane;
            double v = i < nz ? G.chi[i] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += up;
            }
            v += carry;
            if (i < nz) G.chi[i] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    for (int sp = wid; sp < 3 * n_grid; sp += nwarp) {      // one spline per warp at a time
        EpochGrid& G = g[sp / 3];
        double* wk = work + (size_t)wid * 5 * nz;
        switch (sp % 3) {
            // the z grid is uniform: chi(z) and D(z) take the division-free uniform builder
            case 0: spline_build_uniform_warp(nz, (G.z_max - G.z_min) / (nz - 1), G.chi, G.c_chi_z, wk); break;   // cosmology.py:795-796
            case 1: spline_build_warp(nz, G.chi, G.z, G.c_z_chi, wk); break;                                      // :797-798
            default: spline_build_uniform_warp(nz, (G.z_max - G.z_min) / (nz - 1), G.growth, G.c_g_z, wk); break; // :814-815
        }
    }
    __syncthreads();
    // ---- dN/dz normalisations (kernel.py:43-54): 16 panels x GL-8 --------------------------------
    for (int i = 0; i < (same_window ? 1 : 2); ++i) {
        double v = 0.0;
        if (dist[i].kind == CHOMP_DNDZ_TABLE) {
            // piecewise polynomial: every piece integrated in closed form
            for (int j = tid; j < dist[i].n; j += blockDim.x)
                v += dndz_table_piece(dist[i].tab, dist[i].n, j, dist[i].z_min, dist[i].z_max);
        } else if (tid < DNDZ_PANELS * 8) {
            const int pnl = tid >> 3, q = tid & 7;
            const double a = dist[i].z_min + (dist[i].z_max - dist[i].z_min) * pnl / DNDZ_PANELS;
            const double bb = dist[i].z_min + (dist[i].z_max - dist[i].z_min) * (pnl + 1) / DNDZ_PANELS;
            const double half = 0.5 * (bb - a);
            v = half * c_glw[8][q] * dndz_raw(dist[i], 0.5 * (a + bb) + half * c_glx[8][q]);
        }
        dist[i].norm = 1.0 / block_sum(v, red);
    }
    if (same_window) dist[1].norm = dist[0].norm;
    // ---- window tables ---------------------------------------------------------------------------
    for (int i = 0; i < (same_window ? 1 : 2); ++i) {
        const EpochGrid& G = g[1 + i];
        Window& W = win[i];
        // kernel.py:298-305 (comoving_distance at the grid ends is the node value)
        W.chi_min = G.chi[0] < eps ? eps : G.chi[0];
        W.chi_max = G.chi[nz - 1];
        const double hw = (W.chi_max - W.chi_min) / (nw - 1);
        if (cfg.window_kind[i] == CHOMP_WINDOW_GALAXY) {
            // W = dN/dz dz/dchi (kernel.py:382-387)
            for (int j = tid; j < nw; j += blockDim.x) {
                const double chi = (j == nw - 1) ? W.chi_max : W.chi_min + hw * j;
                const double z = grid_z(G, chi);
                W.wf[j] = dndz_eval(dist[i], z) / inv_hubble(c, z);
            }
        } else {
            // lensing efficiency g(chi) = int_{max(chi, g_chi_min)}^{chi_max} dchi' f(chi') (chi' - chi)/chi'
            // with f = dN/dz dz/dchi'  (kernel.py:443-484).  Full panels between the knots of
            // this window's chi(z) table are summed once (suffix sums), the partial panel
            // above each node is integrated on the spot.
            for (int pnl = tid; pnl < nz - 1; pnl += blockDim.x) {
                const double a = G.chi[pnl], bb = G.chi[pnl + 1], half = 0.5 * (bb - a);
                double s0 = 0.0, s1 = 0.0;
                for (int q = 0; q < nql; ++q) {
                    const double x = 0.5 * (a + bb) + half * c_glx[nql][q];
                    const double z = grid_z(G, x);
                    const double f = half * c_glw[nql][q] * dndz_eval(dist[i], z) / inv_hubble(c, z);
                    s0 += f; s1 += f / x;
                }
                lens0[pnl] = s0; lens1[pnl] = s1;
            }
            __syncthreads();
            if (wid < 2) {       // suffix sums (entry nz - 1 = 0): warp scans over the reversed arrays
                double* ls = wid == 0 ? lens0 : lens1;
                double carry = 0.0;
                for (int base = 0; base < nz; base += 32) {
                    const int r = base + lane;               // reversed index: element nz - 1 - r
                    double v = (r > 0 && r < nz) ? ls[nz - 1 - r] : 0.0;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double up = __shfl_up_sync(0xffffffffu, v, o);
                        if (lane >= o) v += up;
                    }
                    v += carry;
                    if (r < nz) ls[nz - 1 - r] = v;
                    carry = __shfl_sync(0xffffffffu, v, 31);
                }
            }
            __syncthreads();
            double g_chi_min = grid_chi(G, dist[i].z_min);            // kernel.py:437-441
            if (g_chi_min < eps) g_chi_min = eps;
            for (int j = tid; j < nw; j += blockDim.x) {
                const double chi = (j == nw - 1) ? W.chi_max : W.chi_min + hw * j;
                const double a_scale = 1.0 / (1.0 + grid_z(G, chi));
                double lo = chi < g_chi_min ? g_chi_min : chi;
                double gl = 0.0;
                if (lo > eps && lo < W.chi_max) {
                    int k = search_index(lo, G.chi, nz);
                    const double bb = G.chi[k + 1], half = 0.5 * (bb - lo);
                    for (int q = 0; q < nql; ++q) {
                        const double x = 0.5 * (lo + bb) + half * c_glx[nql][q];
                        const double z = grid_z(G, x);
                        gl += half * c_glw[nql][q] * dndz_eval(dist[i], z) / inv_hubble(c, z) * (x - chi) / x;
                    }
                    gl += lens0[k + 1] - chi * lens1[k + 1];
                }
                W.wf[j] = 1.5 * c.om * (gl * c.H0 * c.H0 * chi) / a_scale;
            }
        }
        __syncthreads();
    }
    // ---- window splines (warps 0 [, 1]) side by side with the merge of the base-panel edges -----------
    // Base panels of the chi integrals: union of the knots of both windows and of the kernel's chi(z)
    // table.  The three sorted sequences are merged by rank: every candidate counts the members of the
    // other two sequences below it (ties: window a, window b, table -- the order of a sequential merge).
    const int n_spl = same_window ? 1 : 2;
    const double zmin_k = fmax(g[1].z_min, g[2].z_min), zmax_k = fmin(g[1].z_max, g[2].z_max);
    const double chi_min_k = fmax(eps, grid_chi(g[0], zmin_k)), chi_max_k = grid_chi(g[0], zmax_k);
    const int nA = nw, nB = same_window ? 0 : nw, nC = nz, n_cand = nA + nB + nC;
    double* cand = chi_q;                       // sorted candidates (chi_q / fw_q are filled later)
    if (wid < n_spl) {
        Window& W = win[wid];
        double* wk = work + (size_t)wid * 6 * nw;
        // uniform chi nodes
        spline_build_uniform_warp(nw, (W.chi_max - W.chi_min) / (nw - 1), W.wf, W.coef, wk);
    } else {
        const double a0 = win[0].chi_min, ha = (win[0].chi_max - win[0].chi_min) / (nw - 1);
        const double b0 = win[1].chi_min, hb = (win[1].chi_max - win[1].chi_min) / (nw - 1);
        const double* cs = g[0].chi;
        for (int id = tid - 32 * n_spl; id < n_cand; id += blockDim.x - 32 * n_spl) {
            double x;
            int rank;
            if (id < nA) {
                x = a0 + ha * id;
                rank = id + (nB ? count_below_uniform(b0, hb, nB, x, false) : 0) + count_below(cs, nC, x, false);
            } else if (id < nA + nB) {
                x = b0 + hb * (id - nA);
                rank = (id - nA) + count_below_uniform(a0, ha, nA, x, true) + count_below(cs, nC, x, false);
            } else {
                x = cs[id - nA - nB];
                rank = (id - nA - nB) + count_below_uniform(a0, ha, nA, x, true) +
                       (nB ? count_below_uniform(b0, hb, nB, x, true) : 0);
            }
            cand[rank] = x;
        }
    }
    __syncthreads();
    if (same_window) win[1] = win[0];
    {
        // keep a candidate if it lies inside (chi_min, chi_max) and clear of its predecessor; positions
        // by a block-wide exclusive scan of the flags (ballots inside the warps, warp totals in shared memory)
        const double tol = 1e-9 * chi_max_k;
        int* wtot = (int*)red;                  // nwarp + 1 ints
        int run = 1;                            // edge[0] = chi_min_k
        if (tid == 0) edge[0] = chi_min_k;
        for (int base = 0; base < n_cand; base += blockDim.x) {
            const int r = base + tid;
            double x = 0.0;
            bool keep = false;
            if (r < n_cand) {
                x = cand[r];
                keep = x > chi_min_k + tol && x < chi_max_k - tol && (r == 0 || x > cand[r - 1] + tol);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const int before = __popc(bal & ((1u << lane) - 1u));
            __syncthreads();
            if (lane == 0) wtot[wid] = __popc(bal);
            __syncthreads();
            int off = run;
            for (int ww = 0; ww < wid; ++ww) off += wtot[ww];
            if (keep) edge[off + before] = x;
            for (int ww = 0; ww < nwarp; ++ww) run += wtot[ww];
        }
        if (tid == 0) { edge[run] = chi_max_k; n_edge_s = run + 1; }
    }
    __syncthreads();
    // ---- kernel range, z_bar (kernel.py:594-639) -----------------------------------------------------
    LimberF F{win[0], win[1], g[0]};
    {
        double best = -1e300; int besti = 0;
        for (int j = tid; j < nk; j += blockDim.x) {
            const double zj = (j == nk - 1) ? zmax_k : zmin_k + (zmax_k - zmin_k) / (nk - 1) * j;
            const double v = F(grid_chi(g[0], zj));
            if (v > best) { best = v; besti = j; }
        }
        // arg-max with first-index tie break (numpy.argmax)
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        __syncthreads();
        if (lane == 0) { red[wid] = best; red[32 + wid] = (double)besti; }
        __syncthreads();
        if (tid == 0) {
            double bv = red[0]; int bi = (int)red[32];
            for (int i = 1; i < nwarp; ++i) {
                const int oi = (int)red[32 + i];
                if (red[i] > bv || (red[i] == bv && oi < bi)) { bv = red[i]; bi = oi; }
            }
            const double zb = (bi == nk - 1) ? zmax_k : zmin_k + (zmax_k - zmin_k) / (nk - 1) * bi;
            s_misc[0] = zb;
            s_misc[1] = grid_growth(g[0], zb);           // Correlation.D_z, correlation.py:94
        }
    }
    __syncthreads();
    const int n_pan = n_edge_s - 1;
    for (int idx = tid; idx < n_pan * nq; idx += blockDim.x) {
        const int pnl = (nq == 4) ? idx >> 2 : idx / nq, q = idx - pnl * nq;
        const double a = edge[pnl], bb = edge[pnl + 1], half = 0.5 * (bb - a);
        const double x = 0.5 * (a + bb) + half * c_glx[nq][q];
        chi_q[idx] = x;
        fw_q[idx] = half * c_glw[nq][q] * F(x);
    }
    __syncthreads();
    // ---- K(ln k theta) nodes: one warp per node (kernel.py:678-705) ----------------------------------
    const double x0 = log(cfg.ktheta_min), x1 = log(cfg.ktheta_max);
    const int order = cfg.bessel_order;
    // widest base panel: below kt * width <= 2 no panel needs sub-division
    double wmax = 0.0;
    for (int pnl = tid; pnl < n_pan; pnl += blockDim.x) wmax = fmax(wmax, edge[pnl + 1] - edge[pnl]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    __syncthreads();
    if (lane == 0) red[96 + wid] = wmax;
    __syncthreads();
    wmax = 0.0;
    for (int ww = 0; ww < nwarp; ++ww) wmax = fmax(wmax, red[96 + ww]);
    // ---- small k theta: J_n(k theta chi) as a power series in (k theta chi / 2)^2 ------------------
    // For k theta chi_max <= 2 every panel is whole, and
    //   K = sum_m (-1)^m t^m / (m! (m + n)!) * sum_nodes fw (chi / chi_max)^(2 m + n),  t = (k theta chi_max / 2)^2
    // (n = 0 or 2; 12 terms: truncation < 1e-17): the nodes are summed ONCE into 12 moments instead
    // of once per K node.  That covers the lower ~60 % of the ln(k theta) table.
    {
        double mom[LIMBER_SER_TERMS];
#pragma unroll
        for (int m = 0; m < LIMBER_SER_TERMS; ++m) mom[m] = 0.0;
        const double icm = 1.0 / chi_max_k;
        for (int idx = tid; idx < n_pan * nq; idx += blockDim.x) {
            const double r = chi_q[idx] * icm, r2 = r * r;
            double pw = fw_q[idx] * (order == 0 ? 1.0 : r2);
#pragma unroll
            for (int m = 0; m < LIMBER_SER_TERMS; ++m) { mom[m] += pw; pw *= r2; }
        }
        // fixed-order reduction: lanes, then warps
#pragma unroll
        for (int m = 0; m < LIMBER_SER_TERMS; ++m) mom[m] = warp_sum(mom[m]);
        __syncthreads();
        if (lane == 0)
            for (int m = 0; m < LIMBER_SER_TERMS; ++m) red[wid * LIMBER_SER_TERMS + m] = mom[m];
        __syncthreads();
    }
    int* pfx = pfx_all + wid * (nb_max + 1);     // per-warp prefix sums of node counts
    for (int j = wid; j < nk; j += nwarp) {
        const double lkt = (j == nk - 1) ? x1 : x0 + (x1 - x0) / (nk - 1) * j;
        const double kt = exp(lkt);
        if (kt * chi_max_k <= 2.0 && cfg.bessel_limit >= 2.0) {
            // series in t with the moments above (lane m holds term m)
            const double t = 0.25 * kt * kt * chi_max_k * chi_max_k;
            double term = 0.0;
            if (lane < LIMBER_SER_TERMS) {
                double mm = 0.0;
                for (int ww = 0; ww < nwarp; ++ww) mm += red[ww * LIMBER_SER_TERMS + lane];
                // (-1)^m t^m / (m! (m + n)!)  (n = 2: one more power of t); the factorials are tabulated
                double cfac = (order == 0) ? 1.0 : t;
                for (int i = 1; i <= lane; ++i) cfac *= t;
                term = cfac * k_limber_ser[order ? 1 : 0][lane] * mm;
            }
            // sum the terms from the smallest up (ascending magnitude is descending m)
            double accs = 0.0;
            for (int m = LIMBER_SER_TERMS - 1; m >= 0; --m) accs += __shfl_sync(0xffffffffu, term, m);
            if (lane == 0) kn[j] = accs;
            continue;
        }
        double top = cfg.bessel_limit / kt;
        if (top >= chi_max_k) top = chi_max_k;
        double acc = 0.0;
        if (top >= chi_max_k && kt * wmax <= 2.0) {
            // every panel whole and un-divided: W_a W_b D^2 was tabulated at these nodes
            for (int idx = lane; idx < n_pan * nq; idx += 32) acc += fw_q[idx] * bessel_j(order, kt * chi_q[idx]);
        } else {
            // clipped at the Bessel-zero limit and / or sub-divided: spread all (panel, piece,
            // node) triples evenly over the lanes
            const int per = (n_pan + 31) / 32;
            const int p_begin = lane * per, p_end = min(n_pan, (lane + 1) * per);
            auto count = [&](int pnl) -> int {
                const double a = edge[pnl];
                if (!(a < top)) return 0;
                const double bb = fmin(edge[pnl + 1], top);
                return (int)fmax(1.0, ceil(kt * (bb - a) / 2.0)) * nq;
            };
            int mine = 0;
            for (int pnl = p_begin; pnl < p_end; ++pnl) mine += count(pnl);
            int incl = mine;
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int run = incl - mine;
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            for (int pnl = p_begin; pnl < p_end; ++pnl) { pfx[pnl] = run; run += count(pnl); }
            if (lane == 31) pfx[n_pan] = total;
            __syncwarp();
            for (int idx = lane; idx < total; idx += 32) {
                int lo = 0, hi = n_pan;            // panel with pfx[lo] <= idx < pfx[lo + 1]
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pfx[mid] <= idx) lo = mid; else hi = mid; }
                const int pnl = lo, local = idx - pfx[pnl];
                int nsub, sidx, q;
                if (nq == 4) { nsub = (pfx[pnl + 1] - pfx[pnl]) >> 2; sidx = local >> 2; q = local & 3; }
                else { nsub = (pfx[pnl + 1] - pfx[pnl]) / nq; sidx = local / nq; q = local - sidx * nq; }
                const double a = edge[pnl], bfull = edge[pnl + 1];
                if (nsub == 1 && bfull <= top) {
                    acc += fw_q[pnl * nq + q] * bessel_j(order, kt * chi_q[pnl * nq + q]);
                } else {
                    const double bb = fmin(bfull, top);
                    const double d = (bb - a) / nsub, half = 0.5 * d;
                    const double x = a + d * (sidx + 0.5) + half * c_glx[nq][q];
                    acc += half * c_glw[nq][q] * F(x) * bessel_j(order, kt * x);
                }
            }
            __syncwarp();
        }
        acc = warp_sum(acc);
        if (lane == 0) kn[j] = acc;
    }
    __syncthreads();
    if (wid == 0) spline_build_uniform_warp(nk, (x1 - x0) / (nk - 1), kn, kc, work);   // kernel.py:645-646
    __syncthreads();
    // ---- write out ------------------------------------------------------------------------------------
    if (tid == 0) {
        out.zbar[b] = s_misc[0];
        out.dbar[b] = s_misc[1];
        out.kchi[2 * b] = chi_min_k; out.kchi[2 * b + 1] = chi_max_k;
        out.win_chi[4 * b + 0] = win[0].chi_min; out.win_chi[4 * b + 1] = win[0].chi_max;
        out.win_chi[4 * b + 2] = win[1].chi_min; out.win_chi[4 * b + 3] = win[1].chi_max;
        int st = c.bad ? CHOMP_ST_DOMAIN : 0;
        if (!isfinite(s_misc[1])) st |= CHOMP_ST_NONFINITE;
        if (status && st) atomicOr(status + b, st);
    }
    bool bad = false;
    for (int j = tid; j < nk; j += blockDim.x) { out.knodes[(size_t)b * nk + j] = kn[j]; if (!isfinite(kn[j])) bad = true; }
    if (bad && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
    for (int j = tid; j < 4 * (nk - 1); j += blockDim.x) out.kcoef[(size_t)b * 4 * nk + j] = kc[j];
    for (int idx = tid; idx < 3 * nz; idx += blockDim.x) out.chi_nodes[(size_t)b * 3 * nz + idx] = g[idx / nz].chi[idx % nz];
    for (int idx = tid; idx < 13 * nz; idx += blockDim.x) {
        const double* src = idx < nz ? g[0].chi : (idx < 5 * nz ? g[0].c_chi_z : (idx < 9 * nz ? g[0].c_z_chi : g[0].c_g_z));
        const int off = idx < nz ? idx : (idx < 5 * nz ? idx - nz : (idx < 9 * nz ? idx - 5 * nz : idx - 9 * nz));
        out.grid0[(size_t)b * 13 * nz + idx] = src[off];
    }
    if (tid < 2) out.dndz_norm[2 * b + tid] = dist[tid].norm;
    for (int idx = tid; idx < n_edge_s; idx += blockDim.x) out.edges[(size_t)b * edge_stride + idx] = edge[idx];
    if (tid == 0) out.n_edges[b] = n_edge_s;
    for (int idx = tid; idx < 2 * nw; idx += blockDim.x) out.win_nodes[(size_t)b * 2 * nw + idx] = win[idx / nw].wf[idx % nw];
    for (int idx = tid; idx < 2 * 4 * nw; idx += blockDim.x) {
        const int i = idx / (4 * nw), j = idx % (4 * nw);
        if (j < 4 * (nw - 1)) out.win_coef[(size_t)b * 8 * nw + idx] = win[i].coef[j];
    }
}

}  // namespace chomp
