// Stage 4: halo-model power spectra from the splined tables and the Hankel-type k integral
//     w(theta) = int dln k  k^2/(2 pi) P(k)/D(z_bar)^2  K(ln k theta)
// Replaces (reference): Halo.linear_power / power_mm / power_gm / power_gg halo.py:266-439,
// _h_m ... _pp_gg wrappers :649-672; Correlation.correlation / _correlation_integrand
// correlation.py:242-275.
//
// The integrand is a product of splines: smooth between the ln k knots of the halo tables
// and the (theta-shifted) ln k theta knots of the kernel table.  Each halo-table interval is
// cut into pieces no wider than HANKEL_MAX_PIECE with an nq_hankel-point Gauss-Legendre rule
// each; at that resolution the third-derivative jumps of the (much coarser) kernel spline
// contribute < 1e-7, so the nodes are shared by all theta.
#pragma once
#include "common.cuh"
#include "halofit.cuh"
#include "spline.cuh"

namespace chomp {

struct HaloTabs {          // per-point views
    int nk;
    double l0, l1, h;       // ln k_min, ln k_max, knot spacing
    double k_min, k_max;
    const double* tab;      // [5, nk] node values
    const double* coef;     // [5, 4 nk]
    int extrapolate;
    const double* hf;       // HALOFIT parameters of the point, or nullptr (halo.py:1236)
};

__device__ __forceinline__ void which_tables(int which, int& a, int& b, int& pp) {
    // h_m = 0, pp_mm = 1, h_g = 2, pp_gm = 3, pp_gg = 4
    if (which == CHOMP_P_MM) { a = 0; b = 0; pp = 1; }
    else if (which == CHOMP_P_GM) { a = 2; b = 0; pp = 3; }
    else { a = 2; b = 2; pp = 4; }
}

// value of table t at ln k inside [l0, l1], interval index i given
__device__ __forceinline__ double tab_at(const HaloTabs& T, int t, int i, double dx) {
    return spline_poly(T.coef + (size_t)t * 4 * T.nk, i, dx);
}

// Halo.power_xx(k) for any k (halo.py:277-439)
__device__ inline double halo_power(const HaloTabs& T, const PkParams& pk, int which, double k) {
    if (T.hf) {
        // HaloFit: power_mm is the closed form; gm / gg are P_halofit h h + pp with the table
        // wrappers returning 0 outside [k_min, k_max] (halo.py:1325-1412, 649-672)
        if (which == CHOMP_P_LINEAR) return linear_power(pk, k);
        const double pm = halofit_power(T.hf, pk, k);
        if (which == CHOMP_P_MM) return pm;
        if (!(k >= T.k_min && k <= T.k_max)) return 0.0;
        int a, b, pp;
        which_tables(which, a, b, pp);
        const double x = log(k);
        const int i = uniform_index(x, T.l0, 1.0 / T.h, T.nk);
        const double dx = x - (T.l0 + T.h * i);
        return pm * tab_at(T, a, i, dx) * tab_at(T, b, i, dx) + tab_at(T, pp, i, dx);
    }
    const double pl = linear_power(pk, k);
    if (which == CHOMP_P_LINEAR) return pl;
    int a, b, pp;
    which_tables(which, a, b, pp);
    const int nk = T.nk;
    if (k < T.k_min) {
        // P_lin(k) * [h_a h_b + pp / P_lin] at k_min
        const double va = T.tab[a * nk], vb = T.tab[b * nk], vp = T.tab[pp * nk];
        return pl * (va * vb + vp / linear_power(pk, T.k_min));
    }
    const bool inside = T.extrapolate ? (k < T.k_max) : (k <= T.k_max);
    if (inside) {
        const double x = log(k);
        const int i = uniform_index(x, T.l0, 1.0 / T.h, nk);
        const double dx = x - (T.l0 + T.h * i);
        return pl * tab_at(T, a, i, dx) * tab_at(T, b, i, dx) + tab_at(T, pp, i, dx);
    }
    if (!T.extrapolate) return 0.0;
    const double pmax = linear_power(pk, T.k_max);
    const double at_max = pmax * T.tab[a * nk + nk - 1] * T.tab[b * nk + nk - 1] + T.tab[pp * nk + nk - 1];
    if (which == CHOMP_P_MM) return pl * at_max / pmax;                       // halo.py:308-312
    // power law with the mean log-slope over nodes [-7:-1] (halo.py:343-351, 407-415)
    double lv[6];
    for (int j = 0; j < 6; ++j) {
        const int i = nk - 7 + j;
        const double kk = exp(T.l0 + T.h * i);
        lv[j] = log(linear_power(pk, kk) * T.tab[a * nk + i] * T.tab[b * nk + i] + T.tab[pp * nk + i]);
    }
    double slope = 0.0;
    for (int j = 0; j < 5; ++j) slope += (lv[j + 1] - lv[j]) / T.h;
    slope /= 5.0;
    return pow(k / T.k_max, slope) * at_max;
}

struct KernelTab {
    int n;
    double x0, x1, h;       // ln ktheta_min, ln ktheta_max, spacing
    const double* nodes;    // [n]
    const double* coef;     // [4 n]
};
// Kernel.kernel (kernel.py:714-729)
__device__ __forceinline__ double kernel_eval(const KernelTab& K, double u) {
    if (u < K.x0) return K.nodes[0];
    if (!(u <= K.x1)) return 0.0;
    const int i = uniform_index(u, K.x0, 1.0 / K.h, K.n);
    return spline_poly(K.coef, i, u - (K.x0 + K.h * i));
}

__global__ void __launch_bounds__(256)
power_kernel(const Cfg cfg, int B, int which, int n_k, const double* __restrict__ k_in,
             const double* __restrict__ cosmo, const double* __restrict__ epoch,
             const double* __restrict__ htab, const double* __restrict__ hcoef, const double* __restrict__ hfit,
             double* __restrict__ P_out) {
    const int per = (n_k + blockDim.x - 1) / blockDim.x;        // grid.x = per x B (one-dimensional: any B)
    const int b = blockIdx.x / per;
    const int i = (blockIdx.x - b * per) * blockDim.x + threadIdx.x;
    if (b >= B || i >= n_k) return;
    const Cosmo c = load_cosmo(cosmo + (size_t)b * CHOMP_N_COSMO, cfg.cosmo_precision);
    const double* e = epoch + (size_t)b * CHOMP_EPOCH_LEN;
    PkParams pk = make_pk(c, e[EP_GROWTH], e[EP_SIGMA_NORM]);
    CHOMP_ATTACH_BAO(cfg, c, pk)
    HaloTabs T;
    T.nk = cfg.n_halo; T.l0 = log(cfg.k_min); T.l1 = log(cfg.k_max); T.h = (T.l1 - T.l0) / (T.nk - 1);
    T.k_min = cfg.k_min; T.k_max = cfg.k_max; T.extrapolate = cfg.extrapolate;
    T.tab = htab + (size_t)b * 5 * T.nk; T.coef = hcoef + (size_t)b * 20 * T.nk;
    T.hf = hfit ? hfit + (size_t)b * HF_LEN : nullptr;
    P_out[(size_t)b * n_k + i] = halo_power(T, pk, which, k_in[i]);
}

// Widest piece of a halo-table interval used by the Hankel rule: pieces this narrow make the
// kernel-table knots (spacing ~0.38 in ln k theta) harmless, so the nodes do not depend on
// theta and k^2 P(k) is evaluated once per node instead of once per (node, theta).
#define HANKEL_MAX_PIECE 0.0625

__host__ __device__ inline int hankel_subdiv(const Cfg& cfg) {
    const double hP = (log(cfg.k_max) - log(cfg.k_min)) / (cfg.n_halo - 1);
    int m = (int)ceil(hP / HANKEL_MAX_PIECE - 1e-9);
    return m < 1 ? 1 : m;
}
__host__ __device__ inline int hankel_nodes(const Cfg& cfg) { return (cfg.n_halo - 1) * hankel_subdiv(cfg) * cfg.nq_hankel; }

// Correlation(k_min=, k_max=) (correlation.py:104-112): the k integral runs over [lc0, lc1] instead of the
// halo-table range [l0, l1].  Panels: n_lo equal pieces of [lc0, l0] (spectrum from the k < k_min branch,
// halo.py:283-288), the table intervals i_first .. i_first + n_mid - 1 with the outermost two clipped at
// lc0 / lc1, n_hi equal pieces of [l1, lc1] (extrapolated spectrum, or 0).  With the default limits this
// is exactly the table-interval layout.
struct HankelLayout {
    double l0, l1, hP, lc0, lc1;
    int n_lo, i_first, n_mid, n_hi, sub, per, total;
};
__host__ __device__ inline HankelLayout hankel_layout(const Cfg& cfg) {
    HankelLayout L;
    L.l0 = log(cfg.k_min); L.l1 = log(cfg.k_max); L.hP = (L.l1 - L.l0) / (cfg.n_halo - 1);
    L.lc0 = cfg.corr_k_min > 0.0 ? log(cfg.corr_k_min) : L.l0;
    L.lc1 = cfg.corr_k_max > 0.0 ? log(cfg.corr_k_max) : L.l1;
    if (cfg.corr_k_min > 0.0 && cfg.corr_k_min == cfg.k_min) L.lc0 = L.l0;
    if (cfg.corr_k_max > 0.0 && cfg.corr_k_max == cfg.k_max) L.lc1 = L.l1;
    L.sub = hankel_subdiv(cfg);
    L.per = L.sub * cfg.nq_hankel;
    L.n_lo = L.lc0 < L.l0 ? (int)ceil((fmin(L.l0, L.lc1) - L.lc0) / L.hP - 1e-9) : 0;
    L.n_hi = L.lc1 > L.l1 ? (int)ceil((L.lc1 - fmax(L.l1, L.lc0)) / L.hP - 1e-9) : 0;
    if (L.n_lo < 0) L.n_lo = 0;
    if (L.n_hi < 0) L.n_hi = 0;
    L.i_first = 0;
    L.n_mid = 0;
    if (L.lc0 < L.l1 && L.lc1 > L.l0) {
        int i0 = L.lc0 > L.l0 ? (int)floor((L.lc0 - L.l0) / L.hP) : 0;
        if (i0 > cfg.n_halo - 2) i0 = cfg.n_halo - 2;
        int i1 = L.lc1 < L.l1 ? (int)ceil((L.lc1 - L.l0) / L.hP - 1e-12) : cfg.n_halo - 1;     // one past the last interval
        if (i1 > cfg.n_halo - 1) i1 = cfg.n_halo - 1;
        if (i1 <= i0) i1 = i0 + 1;
        L.i_first = i0;
        L.n_mid = i1 - i0;
    }
    L.total = (L.n_lo + L.n_mid + L.n_hi) * L.per;
    return L;
}

// spectrum on the panels outside the halo table (Correlation(k_min=, k_max=) only): out of line, with its own
// table view, so that the common path of wtheta_kernel carries none of it
__device__ __noinline__ double halo_power_outside(int n_halo, double k_min, double k_max, int extrapolate, const double* tab,
                                                  const double* coef, const double* hf, PkParams pk, int which, double k) {
    HaloTabs T;
    T.nk = n_halo; T.l0 = log(k_min); T.l1 = log(k_max); T.h = (T.l1 - T.l0) / (T.nk - 1);
    T.k_min = k_min; T.k_max = k_max; T.extrapolate = extrapolate; T.tab = tab; T.coef = coef; T.hf = hf;
    return halo_power(T, pk, which, k);
}

// grid (B), 256 threads.  Phase 1: G_q = w_q k^2 P(k_q) / (2 pi D^2) at the theta-independent
// Gauss-Legendre nodes (each halo-table interval cut into `sub` equal pieces).  Phase 2: one
// warp per theta sums G_q K(x_q + ln theta).
#ifndef WTHETA_MIN_BLOCKS
#define WTHETA_MIN_BLOCKS 4
#endif
// LIMITS = false: the table-interval layout (no Correlation k limits), nothing of the general path is compiled in
template <bool LIMITS>
__global__ void __launch_bounds__(256, WTHETA_MIN_BLOCKS)
wtheta_kernel(const Cfg cfg, int B, int which, int n_theta, const double* __restrict__ theta,
              const double* __restrict__ cosmo, const double* __restrict__ epoch, const double* __restrict__ dbar,
              const double* __restrict__ htab, const double* __restrict__ hcoef,
              const double* __restrict__ knodes, const double* __restrict__ kcoef, const double* __restrict__ hfit,
              double* __restrict__ w_out, int32_t* __restrict__ status, const int32_t* __restrict__ group) {
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    if (b >= B) return;
    const int gb = group ? group[b] : b;     // row of the cosmology-level tables (fast / slow split)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
    const int nk = cfg.n_halo, nkt = cfg.n_kernel, nq = cfg.nq_hankel;
    const HankelLayout L = hankel_layout(cfg);
    const int sub = L.sub;
    const double* hf = hfit ? hfit + (size_t)gb * HF_LEN : nullptr;
    const int total = L.total;
    double* s_x = sm;                 // total
    double* s_g = s_x + total;        // total
    double* s_kc = s_g + total;       // 4 nkt
    int ta = 0, tb = 0, tpp = 1;
    if (which != CHOMP_P_LINEAR) which_tables(which, ta, tb, tpp);
    const double* hc = hcoef + (size_t)b * 20 * nk;
    const double* ca = hc + (size_t)ta * 4 * nk;
    const double* cb = hc + (size_t)tb * 4 * nk;
    const double* cpp = hc + (size_t)tpp * 4 * nk;
    for (int i = tid; i < 4 * (nkt - 1); i += blockDim.x) s_kc[i] = kcoef[(size_t)gb * 4 * nkt + i];
    const Cosmo c = load_cosmo(cosmo + (size_t)gb * CHOMP_N_COSMO, cfg.cosmo_precision);
    const double* e = epoch + (size_t)gb * CHOMP_EPOCH_LEN;
    PkParams pk = make_pk(c, e[EP_GROWTH], e[EP_SIGMA_NORM]);
    BaoParams bao_store;
    if (LIMITS && cfg.with_bao) { make_bao(c, &bao_store); pk.bao = &bao_store; }    // the general instantiation serves with_bao too
    const double D = dbar[gb];
    const double inv_norm = 1.0 / (2.0 * M_PI * D * D);                 // correlation.py:270-275
    const double l0 = L.l0, l1 = L.l1, hP = L.hP;
    const bool limits = LIMITS;                                          // Correlation(k_min=, k_max=)
    for (int idx = tid; idx < total; idx += blockDim.x) {
        const int j = idx / (sub * nq), r = idx - j * (sub * nq);
        const int s = r / nq, q = r - s * nq;
        const int i = L.i_first + (j - L.n_lo);                        // table interval (middle panels)
        const bool mid = !LIMITS || (j >= L.n_lo && j < L.n_lo + L.n_mid);
        double a = l0 + hP * i;
        double bb = (i == nk - 2) ? l1 : l0 + hP * (i + 1);
        double ea = a, eb = bb;                                       // ends of the panel
        if (limits) {
            if (mid) { ea = fmax(a, L.lc0); eb = fmin(bb, L.lc1); }
            else if (j < L.n_lo) {
                const double top = fmin(l0, L.lc1);
                ea = L.lc0 + (top - L.lc0) * j / L.n_lo;
                eb = (j == L.n_lo - 1) ? top : L.lc0 + (top - L.lc0) * (j + 1) / L.n_lo;
            } else {
                const int jj = j - L.n_lo - L.n_mid;
                const double bot = fmax(l1, L.lc0);
                ea = bot + (L.lc1 - bot) * jj / L.n_hi;
                eb = (jj == L.n_hi - 1) ? L.lc1 : bot + (L.lc1 - bot) * (jj + 1) / L.n_hi;
            }
        }
        const double pa = ea + (eb - ea) * s / sub, pb = (s == sub - 1) ? eb : ea + (eb - ea) * (s + 1) / sub;
        const double half = 0.5 * (pb - pa);
        const double x = 0.5 * (pa + pb) + half * c_glx[nq][q];
        const double k = exp_fast(x);
        const double dx = x - a;
        double P;
        if (mid) {
            P = 2.0 * M_PI * M_PI * (LIMITS ? delta2(pk, k, x) : delta2_eh(pk, k, x)) / (k * k * k);
            if (which != CHOMP_P_LINEAR) {
                if (LIMITS && hf) P = halofit_power(hf, pk, k);          // HaloFit runs on the general instantiation
                if (!(LIMITS && hf && which == CHOMP_P_MM))
                    P = P * spline_poly(ca, i, dx) * spline_poly(cb, i, dx) + spline_poly(cpp, i, dx);
            }
        } else {
            P = halo_power_outside(nk, cfg.k_min, cfg.k_max, cfg.extrapolate, htab + (size_t)b * 5 * nk, hc, hf, pk, which, k);
        }
        s_x[idx] = x;
        s_g[idx] = half * c_glw[nq][q] * k * k * P * inv_norm;
    }
    __syncthreads();
    const double x0 = log(cfg.ktheta_min), x1 = log(cfg.ktheta_max), hK = (x1 - x0) / (nkt - 1);
    const double ihK = 1.0 / hK;
    const double k_first = knodes[(size_t)gb * nkt];
    for (int it = wid; it < n_theta; it += nwarp) {
        const double lt = log(theta[it]);
        double acc = 0.0;
        for (int idx = lane; idx < total; idx += 32) {
            const double u = s_x[idx] + lt;
            double Kv;                                                   // Kernel.kernel, kernel.py:714-729
            if (u < x0) Kv = k_first;
            else if (u > x1) Kv = 0.0;
            else {
                int j = (int)((u - x0) * ihK);
                j = j > nkt - 2 ? nkt - 2 : j;
                Kv = spline_poly(s_kc, j, u - (x0 + hK * j));
            }
            acc = fma(s_g[idx], Kv, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            w_out[(size_t)b * n_theta + it] = acc;
            if (!isfinite(acc) && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
        }
    }
}

}  // namespace chomp
