// Stage 3: the nu-space halo-model mass integrals.
//
//   nu_nodes_kernel    per point: kink-aware Gauss-Legendre node list in ln(nu) with every
//                      k-independent factor folded into per-node weights, plus n_bar
//   halo_sums_kernel   one CTA per (point, chunk of ln k nodes): small-argument series of the NFW
//                      Fourier profile summed into moments once per chunk, table-driven profile
//                      at the remaining nodes, one warp per ln k node; the five sums h_m, pp_mm,
//                      h_g, pp_gm, pp_gg                                       (the hot kernel)
//   halo_splines_kernel per point: normalisations and the five not-a-knot splines in ln k
//
// Replaces (reference): Halo._calculate_n_bar halo.py:674-707; _initialize_h_m :904-927,
// _h_g :929-969, _pp_mm :971-994, _pp_gg :996-1041, _pp_gm :1043-1086; y_nfw :561-585;
// _concentration / _virial_radius :869-902; HaloExclusion._mass_window :1223-1233;
// HODZheng hod.py:141-230; HODMandelbaum hod.py:232-299.
//
// The reference integrates each of these with Romberg over ln(nu) and ignores the kinks of
// the integrands; here the range is cut at every point where an integrand is not smooth
// (knots of the ln M(nu) spline, HOD lower limits, the satellite turn-on M_0, the
// <N> = 1 and <N(N-1)> = 1 exponent switches) and each panel gets a Gauss-Legendre rule whose
// order follows the local phase k r_vir of the profile (three node lists for three k classes).
#pragma once
#include "common.cuh"
#include "special.cuh"
#include "spline.cuh"

namespace chomp {

// Per-node record (structure of arrays).  1/m(c) = 1/(ln(1+c) - c/(1+c)) is folded into the
// weights, so the sums run over rho_k = m(c) * y directly.  The sign of W_GM / W_GG carries the
// reference's exponent switch: negative = "moment < 1, first power of y" (halo.py:1038-1041,
// 1084-1086), positive = second power.
#define NODE_FIELDS 8
enum { NF_CP = 0, NF_RS, NF_LNRS /* ln r_s: ln z = ln k + ln r_s without a logarithm per node */, NF_W_HM, NF_W_PMM,
       NF_W_HG, NF_W_GM, NF_W_GG };
#define MAX_EXTRA_BREAKS 8
// k classes: the panel order grows with phi = k * r_vir(M_max), the phase of the profile's
// oscillation in mass across the table.  {smooth panels, panels inside the central-galaxy
// erf edge log_M_min +- 3.5 sigma}
#define N_KCLASS 3
#define KCLASS_PHI_1 45.0
#define KCLASS_PHI_2 180.0
// per-panel order inside a class: local phase k_top(class) * r_vir(panel top)
#ifndef KPANEL_PHI_1
#define KPANEL_PHI_1 12.0
#define KPANEL_PHI_2 60.0
#endif
// A fourth node list serves the 1-halo trispectrum only (built when cfg.tri_moment >= 0): its
// M^3 weighting moves the integrand to high masses, where y(k, M) oscillates fastest, so every
// panel is cut in two halves of order 16 ("32").
#define N_NODE_LISTS 4
#define TRI_LIST 3
__device__ __constant__ int k_class_base[N_NODE_LISTS] = {4, 8, 16, 32};
__device__ __constant__ int k_class_sharp[N_NODE_LISTS] = {10, 10, 16, 32};
#define KCLASS_MAX_ORDER 16
#define SING_MIN_ORDER 8
#ifndef SUMS_K_PER_CTA
#define SUMS_K_PER_CTA 128
#endif

struct HodP {
    int kind;
    double log_M_min, sigma, log_M_0, log_M_1p, alpha, w;
    double M0, M1p, Mmin;          // 10**log_M_0, 10**log_M_1p, 10**log_M_min
    double first_zero, second_zero; // hod.py:176-185; -1 for Mandelbaum
};

__device__ __noinline__ HodP load_hod(int kind, const double* __restrict__ p, double halo_precision) {
    const double ln10 = 2.302585092994046;
    HodP h;
    h.kind = kind;
    if (kind == CHOMP_HOD_ZHENG) {
        h.log_M_min = p[0]; h.sigma = p[1]; h.log_M_0 = p[2]; h.log_M_1p = p[3]; h.alpha = p[4]; h.w = 0.0;
        h.first_zero = exp_fast(ln10 * (h.log_M_min + h.sigma * erfinv(2.0 * halo_precision - 1.0)));
        h.second_zero = exp_fast(ln10 * h.log_M_0);   // the reference's clamp is a typo'd no-op (hod.py:183-184)
    } else {
        h.log_M_0 = p[0]; h.w = p[1]; h.log_M_min = log10(3.0) + p[0];
        h.sigma = 0.0; h.log_M_1p = 0.0; h.alpha = 0.0;
        h.first_zero = -1.0; h.second_zero = -1.0;
    }
    h.M0 = exp_fast(ln10 * h.log_M_0);
    h.M1p = exp_fast(ln10 * h.log_M_1p);
    h.Mmin = exp_fast(ln10 * h.log_M_min);
    return h;
}

// <N>, <N(N-1)>  (hod.py:188-230 Zheng, 262-299 Mandelbaum); lm = ln(M) is supplied by the
// callers, who all have it
__device__ __noinline__ void hod_moments(const HodP& h, double M, double lm, double& n1, double& n2) {
    const double lg = lm * 0.43429448190325182765;
    double nc, ns;
    if (h.kind == CHOMP_HOD_ZHENG) {
        if (h.sigma <= 0.0) nc = (lg > h.log_M_min) ? 1.0 : 0.0;
        else nc = 0.5 * (1.0 + erf((lg - h.log_M_min) / h.sigma));
        const double d = M - h.M0;
        ns = (d > 0.0) ? nc * exp_fast(h.alpha * log(d / h.M1p)) : 0.0;
    } else {
        nc = (lg >= h.log_M_0) ? 1.0 : 0.0;
        const double r = M / h.Mmin;
        ns = (lg < h.log_M_min) ? r * r * h.w : r * h.w;
    }
    n1 = nc + ns;
    n2 = (2.0 + ns) * ns;
}

struct NuTab {   // shared-memory view of stage 2's tables for one point
    int n;
    const double *lnm, *nu, *c_lnm_nu, *c_nu_lnm;
};

__device__ __forceinline__ double mass_of_nu_ln(const NuTab& t, double nu) {
    return spline_eval_search(t.c_lnm_nu, nu, t.nu, t.n);    // MassFunction.ln_mass, mass_function.py:326
}
__device__ __forceinline__ double nu_of_lnm(const NuTab& t, double lnm) {
    const double h = (t.lnm[t.n - 1] - t.lnm[0]) / (t.n - 1);
    return spline_eval_uniform(t.c_nu_lnm, lnm, t.lnm[0], h, t.n);  // MassFunction.nu, :315
}

// ln(nu) at which exp(ln_mass(nu)) equals the target mass; NaN if outside (nu_min, nu_max).
__device__ __noinline__ double lnnu_of_lnm_inverse(const NuTab& t, double lnm_t, double nu_min, double nu_max) {
    if (!(mass_of_nu_ln(t, nu_min) < lnm_t) || !(mass_of_nu_ln(t, nu_max) > lnm_t)) return nan("");
    int i = 0;
    while (i < t.n - 2 && t.lnm[i + 1] <= lnm_t) ++i;       // node values bracket the root
    // regula falsi (Illinois variant) on the monotone cubic of this interval
    double lo = t.nu[i], hi = t.nu[i + 1];
    double flo = t.lnm[i] - lnm_t, fhi = t.lnm[i + 1] - lnm_t;
    if (flo == 0.0) return log(lo);
    for (int it = 0; it < 60 && hi - lo > 4e-16 * hi; ++it) {
        double mid = (lo * fhi - hi * flo) / (fhi - flo);
        if (!(mid > lo && mid < hi)) mid = 0.5 * (lo + hi);
        const double fm = spline_poly(t.c_lnm_nu, i, mid - t.nu[i]) - lnm_t;
        if (fm == 0.0) { lo = hi = mid; break; }
        if (fm < 0.0) { lo = mid; flo = fm; fhi *= 0.5; } else { hi = mid; fhi = fm; flo *= 0.5; }
    }
    return log(0.5 * (lo + hi));
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ln(nu) in (a, b) where moment(mass(nu)) crosses 1 (which = 1: <N>, 2: <N(N-1)>); NaN if none.
// Both moments are non-decreasing in mass.  Warp-collective (all 32 lanes, result in every lane):
// the knots of the ln M(nu) table are tested in parallel to find the knot interval, which is then
// cut into 33 pieces per round, every lane testing one cut, until the bracket has closed to
// rounding.  A step in the moment (sigma <= 0, Mandelbaum centrals) converges onto the step.
__device__ __noinline__ double lnnu_moment_crossing(const NuTab& t, const HodP& h, int which, double a, double b) {
    const int lane = threadIdx.x & 31;
    double n1, n2;
    // the two ends, on lanes 0 and 1
    {
        const double lm = mass_of_nu_ln(t, exp_fast(lane == 0 ? a : b));
        hod_moments(h, exp_fast(lm), lm, n1, n2);
    }
    const bool lt_end = (which == 1 ? n1 : n2) < 1.0;
    const unsigned ends = __ballot_sync(0xffffffffu, lt_end);
    if (!(ends & 1u) || (ends & 2u)) return nan("");
    double lo = a, hi = b;
    for (int base = 1; base < t.n - 1; base += 32) {              // interior knots
        const int i = base + lane;
        const bool have = i < t.n - 1;
        const int ii = have ? i : 1;
        const double x = log(t.nu[ii]);
        hod_moments(h, exp_fast(t.lnm[ii]), t.lnm[ii], n1, n2);
        const bool inside = have && x > a && x < b;
        const bool lt = (which == 1 ? n1 : n2) < 1.0;
        lo = fmax(lo, warp_max(inside && lt ? x : a));
        hi = fmin(hi, warp_min(inside && !lt ? x : b));
    }
    const int kn = search_index(exp_fast(0.5 * (lo + hi)), t.nu, t.n);
    for (int round = 0; round < 16 && hi - lo > 1e-15 * fabs(hi); ++round) {
        const double x = lo + (hi - lo) * ((lane + 1) * (1.0 / 33.0));
        const double v = exp_fast(x);
        const double lm = spline_poly(t.c_lnm_nu, kn, v - t.nu[kn]);
        hod_moments(h, exp_fast(lm), lm, n1, n2);
        const bool lt = (which == 1 ? n1 : n2) < 1.0;
        const double nlo = warp_max(lt ? x : lo), nhi = warp_min(lt ? hi : x);
        lo = nlo; hi = nhi;
    }
    return hi;   // first abscissa on the ">= 1" side
}

struct NodesOut {
    double* nodes;     // [B, cap_total * NODE_FIELDS]; class c occupies [off_c*NF, (off_c+cap_c)*NF), fields SoA
    int32_t* n_nodes;  // [B, N_NODE_LISTS]
    double* nbar;      // [B] n_bar / rho_bar
    double* rv_max;    // [B, 3] r_vir at the upper end of the mass table, and the first ln k index of k classes 1 and 2
    double* tri_w;     // [B, cap of the last class] weights of the 1-halo trispectrum integral (0-padded)
    int cap[N_NODE_LISTS];
    int off[N_NODE_LISTS];
    int cap_total;
};

__host__ __device__ inline int kclass_cap(int c, int n_mass, bool with_trispectrum) {
    const int base[N_NODE_LISTS] = {4, 8, 16, 32}, sharp[N_NODE_LISTS] = {10, 10, 16, 32};
    if (c == TRI_LIST && !with_trispectrum) return 32;
    int m = base[c] > sharp[c] ? base[c] : sharp[c];
    if (m < SING_MIN_ORDER) m = SING_MIN_ORDER;
    // at most six of the MAX_EXTRA_BREAKS slots are ever filled (two lower limits, M_0, the step /
    // slope change, two moment crossings)
    return (((n_mass - 1 + 6) * m + 31) / 32) * 32;
}

// first ln k node index whose phi = k * rv_max reaches `phi`
__device__ __forceinline__ int kclass_first_index(double phi, double rv_max, double l0, double hk, int nk) {
    const double x = (log(phi / rv_max) - l0) / hk;
    if (!(x > 0.0)) return 0;
    if (x >= (double)nk) return nk;
    return (int)ceil(x);
}

// weight of a node of the 1-halo trispectrum integral: d ln(nu) nu f(nu) M^3 <moment> / rho_bar^3
// (halo_trispectrum.py:131-151; full nu range).  Out of line: the trispectrum list is built only on
// request and its pow() should not sit in the node loop's instruction stream.
__device__ __noinline__ double tri_node_weight(int tri_moment, double wt, double mr, double n1, double n2) {
    double mom = 1.0;
    if (tri_moment == 1) mom = n1;
    else if (tri_moment == 2) mom = n2;
    else if (tri_moment >= 3) {
        const double a2 = (n1 != 0.0) ? n2 / (n1 * n1) : 0.0;       // HOD.nth_moment, hod.py:68-92
        mom = pow(n1, (double)tri_moment);
        for (int jm = 0; jm < tri_moment; ++jm) mom *= (jm * a2 - jm + 1);
    }
    return wt * mr * mr * mr * mom;
}

#ifndef NODES_MIN_BLOCKS
#define NODES_MIN_BLOCKS 8
#endif
// VAR = false: Sheth-Tormen only (the default instantiation carries nothing of the Tinker form)
template <bool VAR>
__global__ void __launch_bounds__(128, NODES_MIN_BLOCKS)
nu_nodes_kernel(const Cfg cfg, int B, const double* __restrict__ halo, const double* __restrict__ hod,
                const double* __restrict__ epoch, const double* __restrict__ g_lnm, const double* __restrict__ g_nu,
                const double* __restrict__ g_c1, const double* __restrict__ g_c2, NodesOut out,
                int32_t* __restrict__ status, const int32_t* __restrict__ group /* [B] or null */,
                const int32_t* __restrict__ group_status) {
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    if (b >= B) return;
    // fast / slow split: the cosmology-level tables (halo parameters, epoch scalars, nu(M) splines) of
    // point b live in row gb of their arrays; everything HOD-level is per point
    const int gb = group ? group[b] : b;
    const int n = cfg.n_mass, tid = threadIdx.x;
    const int max_edge = n + MAX_EXTRA_BREAKS;
    double* lnm = sm;
    double* nu = lnm + n;
    double* c1 = nu + n;
    double* c2 = c1 + 4 * n;
    double* edge = c2 + 4 * n;                 // max_edge
    double* extra = edge + max_edge;           // MAX_EXTRA_BREAKS
    double* red = extra + MAX_EXTRA_BREAKS;    // 64
    int* pstart = (int*)(red + 64);            // [N_NODE_LISTS][max_edge] first node of each panel
    int* pknot = pstart + N_NODE_LISTS * max_edge; // [max_edge] knot interval of the ln M(nu) spline holding the panel
    __shared__ int n_edge;
    __shared__ double x_singular;
    for (int i = tid; i < n; i += blockDim.x) { lnm[i] = g_lnm[(size_t)gb * n + i]; nu[i] = g_nu[(size_t)gb * n + i]; }
    for (int i = tid; i < 4 * (n - 1); i += blockDim.x) {
        c1[i] = g_c1[(size_t)gb * 4 * n + i];
        c2[i] = g_c2[(size_t)gb * 4 * n + i];
    }
    // the HOD constants (an erfinv and a few exponentials) once per CTA
    __shared__ HodP s_hod;
    if (tid == blockDim.x - 1) s_hod = load_hod(cfg.hod_kind, hod + (size_t)b * CHOMP_N_HOD, cfg.halo_precision);
    __syncthreads();
    NuTab t{n, lnm, nu, c1, c2};
    const double* e = epoch + (size_t)gb * CHOMP_EPOCH_LEN;
    const double nu_min = e[EP_NU_MIN], nu_max = e[EP_NU_MAX];
    const double l_min = log(nu_min), l_max = log(nu_max);
    const HodP h = s_hod;
    // lower limits of the galaxy integrals (halo.py:675-679, 935-939, 1002-1006, 1049-1053)
    double x_lo1 = l_min, x_lo2 = l_min;
    if (h.first_zero > -1.0 && h.first_zero > exp(e[EP_LNM_MIN])) x_lo1 = log(nu_of_lnm(t, log(h.first_zero)));
    if (h.second_zero > -1.0 && h.second_zero > exp(e[EP_LNM_MIN])) x_lo2 = log(nu_of_lnm(t, log(h.second_zero)));
    // ---- break points: the two expensive root finders on warps of their own (lanes of one warp
    //      would run the different searches one after the other) ------------------------------
    {
        const int wid = tid >> 5, lane = tid & 31;
        if (wid == 0 && lane < 4) {
            double x = nan("");
            if (lane == 0) x = x_lo1;
            else if (lane == 1) x = x_lo2;
            else if (lane == 2) x = lnnu_of_lnm_inverse(t, log(h.M0), nu_min, nu_max);   // satellite turn-on / central step
            else if (h.kind == CHOMP_HOD_MANDELBAUM || h.sigma <= 0.0)
                x = lnnu_of_lnm_inverse(t, log(h.Mmin), nu_min, nu_max);   // Mandelbaum slope change at 3 M_0; Zheng step
            extra[lane] = x;
        }
        if (wid == 1) {                                                                // halo.py:1084-1086
            const double x = lnnu_moment_crossing(t, h, 1, fmax(x_lo1, l_min), l_max);
            if (lane == 0) extra[4] = x;
        }
        if (wid == 2) {                                                                // halo.py:1038-1041
            const double x = lnnu_moment_crossing(t, h, 2, fmax(x_lo2, l_min), l_max);
            if (lane == 0) extra[5] = x;
        }
        if (wid == 3 && lane < 2) extra[6 + lane] = nan("");
    }
    // ln(nu) of the knots, in parallel (staged in the tail of the edge array)
    double* lognu = edge + MAX_EXTRA_BREAKS;     // edge has n + MAX_EXTRA_BREAKS slots; the merge below
    for (int i = tid; i < n; i += blockDim.x) lognu[i] = log(nu[i]);   // writes edge[k] only for k <= i + extras so far
    __syncthreads();
    if (tid == 0) {
        int cnt = 0;
        edge[cnt++] = l_min;
        for (int i = 1; i < n - 1; ++i) {
            const double x = lognu[i];
            if (x > l_min && x < l_max) edge[cnt++] = x;      // cnt <= i < i + MAX_EXTRA_BREAKS: never overtakes lognu
        }
        edge[cnt++] = l_max;
        double x_sing = nan("");
        if (h.kind == CHOMP_HOD_ZHENG) x_sing = extra[2];
        for (int j = 0; j < MAX_EXTRA_BREAKS; ++j) {
            const double x = extra[j];
            if (!(x > l_min && x < l_max)) continue;
            int pos = 0;
            while (pos < cnt && edge[pos] < x) ++pos;
            const double tol = 1e-12;
            if ((pos < cnt && fabs(edge[pos] - x) < tol) || (pos > 0 && fabs(edge[pos - 1] - x) < tol)) continue;
            for (int k = cnt; k > pos; --k) edge[k] = edge[k - 1];
            edge[pos] = x;
            ++cnt;
        }
        n_edge = cnt;
        x_singular = x_sing;
    }
    __syncthreads();
    // per panel (in parallel): spline interval and the Gauss-Legendre order in each k class.  The
    // class order (set by the phase phi = k r_vir at the TOP of the mass table) is only needed where
    // the profile really oscillates that fast: a panel whose own phase k_top(class) r_vir(panel top)
    // stays below KPANEL_PHI_1 / _2 gets order 4 / 8 (tools/adaptive_orders.py: same accuracy,
    // less than half the nodes in the two fine lists).  Panels inside the erf edge of the central
    // occupation get at least the "sharp" order.
    const double rv_coef = 3.0 / (4.0 * M_PI * e[EP_DELTA_V] * e[EP_RHO_BAR]);
    const double rvm = cbrt(rv_coef * exp(lnm[n - 1]));
    const int nkh = cfg.n_halo;
    const double lk0 = log(cfg.k_min), hkh = (log(cfg.k_max) - lk0) / (nkh - 1);
    const int ic1 = kclass_first_index(KCLASS_PHI_1, rvm, lk0, hkh, nkh);
    const int ic2 = max(ic1, kclass_first_index(KCLASS_PHI_2, rvm, lk0, hkh, nkh));
    {
        const double ln10 = 2.302585092994046;
        const double sharp_lo = (h.log_M_min - 3.5 * h.sigma) * ln10, sharp_hi = (h.log_M_min + 3.5 * h.sigma) * ln10;
        // largest k of each class's ln k nodes (the trispectrum list serves every k)
        const double k_top[N_NODE_LISTS] = {exp(lk0 + hkh * max(ic1 - 1, 0)), exp(lk0 + hkh * max(ic2 - 1, 0)), cfg.k_max, cfg.k_max};
        for (int p = tid; p < n_edge - 1; p += blockDim.x) {
            const double xm = 0.5 * (edge[p] + edge[p + 1]);
            const int kn = search_index(exp(xm), nu, n);
            pknot[p] = kn;
            const double lm_lo = spline_poly(c1, kn, exp(edge[p]) - nu[kn]);
            const double lm_hi = spline_poly(c1, kn, exp(edge[p + 1]) - nu[kn]);
            const bool sharp = (h.kind == CHOMP_HOD_ZHENG) && h.sigma > 0.0 && lm_hi > sharp_lo && lm_lo < sharp_hi;
            const bool sing = edge[p] >= x_singular - 1e-12 && edge[p] <= x_singular + 0.02;
            const double rv_p = cbrt(rv_coef * exp(lm_hi));
            for (int c = 0; c < N_NODE_LISTS; ++c) {
                int o = k_class_base[c];
                if (c < N_KCLASS) {
                    const double phi_p = k_top[c] * rv_p;
                    const int o_local = phi_p < KPANEL_PHI_1 ? 4 : (phi_p < KPANEL_PHI_2 ? 8 : 16);
                    if (o_local < o) o = o_local;
                }
                if (sharp && o < k_class_sharp[c]) o = (c < N_KCLASS) ? max(o, 10) : k_class_sharp[c];
                if (sing && o < SING_MIN_ORDER) o = SING_MIN_ORDER;
                pstart[c * max_edge + p] = o;
            }
        }
    }
    __syncthreads();
    if (tid < N_NODE_LISTS) {    // exclusive prefix sums, one list per thread
        int acc = 0;
        int* ps = pstart + tid * max_edge;
        for (int p = 0; p < n_edge - 1; ++p) { const int o = ps[p]; ps[p] = acc; acc += o; }
        ps[n_edge - 1] = acc;
    }
    __syncthreads();
    // ---- nodes ------------------------------------------------------------------------------
    const int n_pan = n_edge - 1;
    const double* hp = halo + (size_t)gb * CHOMP_N_HALO;
    const double stq = hp[CHOMP_H_STQ], sta = hp[CHOMP_H_ST_LITTLE_A], beta = hp[CHOMP_H_BETA];
    const double c0 = hp[CHOMP_H_C0] / (1.0 + e[EP_Z]);                   // halo.py:65
    const double f_norm = e[EP_F_NORM], b_norm = e[EP_B_NORM], delta_c = e[EP_DELTA_C];
    const double rho_bar = e[EP_RHO_BAR], lnm_star = e[EP_LNM_STAR];
    const double ln_rv_coef = log(rv_coef), ln_c0 = log(c0), ln_sta = log(sta);
    MfParams mf;
    mf.kind = CHOMP_MF_SHETH_TORMEN;
    if (VAR) mf = mf_params(cfg, sta, stq, delta_c, e[EP_DELTA_V], e[EP_Z]);
    double nbar = 0.0;
    int st = 0;
    const int n_lists = cfg.tri_moment >= 0 ? N_NODE_LISTS : N_KCLASS;
#pragma unroll 1
    for (int c = 0; c < n_lists; ++c) {
        const int* ps = pstart + c * max_edge;
        const int total = ps[n_pan];
        const int cap = out.cap[c];
        if (total > cap) st |= CHOMP_ST_NODE_OVERFLOW;
        double* rec = out.nodes + ((size_t)b * out.cap_total + out.off[c]) * NODE_FIELDS;
        // one thread per (panel, node): walk the panels with a running index (not unrolled: four
        // copies of this body overflow the instruction cache)
#pragma unroll 1
        for (int idx = tid; idx < total && idx < cap; idx += blockDim.x) {
            int p = 0, hi = n_pan;          // panel with ps[p] <= idx < ps[p+1]
            while (hi - p > 1) { const int mid = (p + hi) >> 1; if (ps[mid] <= idx) p = mid; else hi = mid; }
            int nq = ps[p + 1] - ps[p], q = idx - ps[p];
            double a = edge[p], bb = edge[p + 1];
            const double xmid = 0.5 * (a + bb);
            const bool sing_panel = a >= x_singular - 1e-12 && a <= x_singular + 0.02;
            if (nq == 32) {          // two halves of order 16 (trispectrum list)
                const double mid2 = 0.5 * (a + bb);
                if (q < 16) bb = mid2; else { a = mid2; q -= 16; }
                nq = 16;
            }
            double x, wq;
            if (sing_panel && a == edge[p]) {
                // x = a + (b - a) t^4 removes the (M - M0)^alpha end-point behaviour (hod.py:226-230);
                // also applied when the panel starts just above M0 (lower limit from the forward spline)
                const double tt = 0.5 * (c_glx[nq][q] + 1.0);
                const double t2 = tt * tt;
                x = a + (bb - a) * t2 * t2;
                wq = (bb - a) * 4.0 * t2 * tt * 0.5 * c_glw[nq][q];
            } else {
                const double half = 0.5 * (bb - a);
                x = 0.5 * (a + bb) + half * c_glx[nq][q];
                wq = half * c_glw[nq][q];
            }
            const double v = exp_fast(x);
            const int kn = pknot[p];
            const double lm = spline_poly(c1, kn, v - nu[kn]);      // MassFunction.ln_mass, mass_function.py:326
            const double M = exp_fast(lm);
            double nf, bias;
            if (VAR) mf_raw_ln(mf, x, nf, bias);
            else st_raw_ln(x, ln_sta, stq, delta_c, nf, bias);
            const double wt = wq * nf * f_norm;          // d ln(nu) * nu f(nu)
            bias *= b_norm;
            const double con = c0 * exp_fast(beta * (lm - lnm_star));                      // halo.py:869-873
            const double r_v = exp_fast((ln_rv_coef + lm) * (1.0 / 3.0));                  // halo.py:890-893
            const double cp = 1.0 + con;
            const double lncp = log(cp);
            const double imk = 1.0 / (lncp - con / cp);                                    // halo.py:584
            double n1, n2;
            hod_moments(h, M, lm, n1, n2);
            const double in1 = (xmid > x_lo1) ? 1.0 : 0.0, in2 = (xmid > x_lo2) ? 1.0 : 0.0;
            rec[NF_CP * cap + idx] = cp;
            rec[NF_RS * cap + idx] = r_v / con;
            rec[NF_LNRS * cap + idx] = (ln_rv_coef + lm) * (1.0 / 3.0) - (ln_c0 + beta * (lm - lnm_star));
            rec[NF_W_HM * cap + idx] = wt * bias * imk;                                     // halo.py:923-927
            rec[NF_W_PMM * cap + idx] = wt * M / rho_bar * imk * imk;                       // halo.py:990-994, :988
            rec[NF_W_HG * cap + idx] = in1 * wt * bias * n1 / M * imk;                      // halo.py:964-969
            const double wgm = in1 * wt * n1;                                               // halo.py:1078-1086
            rec[NF_W_GM * cap + idx] = (n1 < 1.0) ? -wgm * imk : wgm * imk * imk;
            const double wgg = in2 * wt * n2 / M;                                           // halo.py:1032-1041
            rec[NF_W_GG * cap + idx] = (n2 < 1.0) ? -wgg * imk : wgg * imk * imk;
            if (c == N_KCLASS - 1) nbar += in1 * wt * n1 / M;                               // halo.py:704-707
            if (c == TRI_LIST) out.tri_w[(size_t)b * cap + idx] = tri_node_weight(cfg.tri_moment, wt, M / rho_bar, n1, n2);
        }
        if (c == TRI_LIST)
            for (int idx = total + tid; idx < cap; idx += blockDim.x) out.tri_w[(size_t)b * cap + idx] = 0.0;
        if (tid == 0) out.n_nodes[(size_t)b * N_NODE_LISTS + c] = total < cap ? total : cap;
    }
    nbar = block_sum(nbar, red);
    if (tid == 0) {
        out.nbar[b] = nbar;
        out.rv_max[3 * b] = rvm;
        out.rv_max[3 * b + 1] = (double)ic1;
        out.rv_max[3 * b + 2] = (double)ic2;
        if (!isfinite(nbar)) st |= CHOMP_ST_NONFINITE;
        if (group_status) st |= group_status[gb];
        if (status && st) atomicOr(status + b, st);
    }
}

// HaloExclusion._mass_window (halo.py:1223-1233), kR = 2 k r_v
__device__ __forceinline__ double exclusion_window(const SiciTables* t, double kR) {
    double si, ci, s, c;
    sici(t, kR, si, ci);
    sincos(kR, &s, &c);
    return (kR * c + kR * kR * kR * ci + (2.0 - kR * kR) * s) / (3.0 * kR);
}

// ---- small-argument series of the NFW profile numerator ---------------------------------------
// rho(k, M) = int_0^c dx x/(1+x)^2 j0(k r_s x)  (the integral halo.py:574-583 is the closed form of)
//           = sum_n a_n(c) t^n,   t = (k r_s c)^2 = (k r_vir)^2,
//   a_n(c) = (-1)^n J_{2n+1}(c) / ((2n+1)! c^{2n}),   J_m(c) = int_0^c x^m / (1+x)^2 dx.
// J_m / c^m follows the forward recurrence (stable for c >= 1, accurate enough down to c = 0.3)  j_m = (l_{m-1} - j_{m-1}) / c,
// l_m = 1/m - l_{m-1} / c  with  l_0 = ln(1+c), j_0 = c / (1+c).  With t <= SER_X^2 and degree
// SER_DEG the truncation error is below 4e-15 (tools/series_check.py).
// Because every k-independent factor of the five integrands is already folded into the node
// weights, the part of each sum that comes from nodes with k r_vir <= SER_X for EVERY k of a CTA's
// chunk collapses into 5 (SER_DEG + 1) moments  S[s][n] = sum_i w_s(i) coef_n(i) (k_hi r_vir,i)^2n
// (coef = a for the sums linear in rho, a*a for the quadratic ones), evaluated per k as a
// polynomial in (k / k_hi)^2: those (k, node) pairs are never visited.
#ifndef SER_DEG
#define SER_DEG 11
#endif
#ifndef SER_X
#define SER_X 3.0
#endif
#define SER_NC (SER_DEG + 1)
#define SER_MIN_C 0.3          // the forward recurrence loses c^-(2n+1) in a_n: still 7e-15 overall at c = 0.3
                               // (tools/series_check.py); nodes below take the general path

__device__ __forceinline__ void nfw_series_coeffs(double c, double cp, double lncp, double (&a)[SER_NC]) {
    const double ic = 1.0 / c;
    double l = lncp, j = c / cp, fact = 1.0;
#pragma unroll
    for (int m = 1; m <= 2 * SER_DEG + 1; ++m) {
        const double jn = (l - j) * ic;
        l = 1.0 / (double)m - l * ic;
        j = jn;
        fact *= (double)m;
        if (m & 1) {
            const int n = (m - 1) >> 1;
            a[n] = ((n & 1) ? -j : j) * c / fact;
        }
    }
}

// grid (chunks x B): a CTA stages the node list of one k class in shared memory and its 8
// warps take the ln k nodes of one SUMS_K_PER_CTA chunk of that class; lanes stride the nodes.
#ifndef SUMS_MIN_BLOCKS
#define SUMS_MIN_BLOCKS 3
#endif
#define SER_ROUNDS ((SER_NC + 2) / 3)     // moment orders are reduced three at a time
#define SER_MOM (16 * SER_ROUNDS)          // moment of order n, sum s at [16 (n / 3) + 5 (n % 3) + s]
#define SUMS_EXTRA_DOUBLES (SER_MOM + 8 * SER_MOM)   // moments + per-warp partial moments
__global__ void __launch_bounds__(256, SUMS_MIN_BLOCKS)
halo_sums_kernel(const Cfg cfg, int B, NodesOut nd, int smem_doubles, double* __restrict__ raw /* [B, 5, n_halo] */) {
    extern __shared__ double srec[];      // NODE_FIELDS * nn_pad records | series coefficients | moments
    __shared__ NfwTables ntab;
    __shared__ int s_first_bad;
    // grid.x = chunks per point x B (a one-dimensional grid: B is not bounded by the 65 535 of gridDim.y)
    const int n_chunk_slots = (cfg.n_halo + SUMS_K_PER_CTA - 1) / SUMS_K_PER_CTA + N_KCLASS - 1;
    const int b = blockIdx.x / n_chunk_slots;
    if (b >= B) return;
    const int nk = cfg.n_halo;
    const double l0 = log(cfg.k_min), l1 = log(cfg.k_max), hk = (l1 - l0) / (nk - 1);
    // which class / k range does this CTA own?
    const int i1 = (int)nd.rv_max[3 * b + 1], i2 = (int)nd.rv_max[3 * b + 2];     // set by nu_nodes_kernel
    const int first[N_KCLASS + 1] = {0, i1, i2, nk};
    int chunk = blockIdx.x - b * n_chunk_slots, cls = -1, k_begin = 0, k_end = 0;
    for (int c = 0; c < N_KCLASS; ++c) {
        const int cnt = first[c + 1] - first[c];
        const int nch = (cnt + SUMS_K_PER_CTA - 1) / SUMS_K_PER_CTA;
        if (chunk < nch) {
            cls = c;
            k_begin = first[c] + chunk * SUMS_K_PER_CTA;
            k_end = min(first[c + 1], k_begin + SUMS_K_PER_CTA);
            break;
        }
        chunk -= nch;
    }
    if (cls < 0) return;
    const int cap = nd.cap[cls];
    const int nn = nd.n_nodes[(size_t)b * N_NODE_LISTS + cls];
    const double* __restrict__ g = nd.nodes + ((size_t)b * nd.cap_total + nd.off[cls]) * NODE_FIELDS;
    // stage: field f of node i at srec[f * nn_pad + i].  Asynchronous copies (no register round
    // trip): the coefficient tables are loaded while the records are in flight.
    const int nn_pad = (nn + 31) & ~31;
#pragma unroll
    for (int f = 0; f < NODE_FIELDS; ++f)
        for (int i = threadIdx.x; i < nn; i += blockDim.x) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(srec + f * nn_pad + i);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(g + (size_t)f * cap + i) : "memory");
        }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // halo-exclusion window only: its Si/Ci tables sit at the end of the dynamic shared memory
    SiciTables* tabs = (SiciTables*)(srec + smem_doubles);
    if (cfg.exclusion) sici_tables_load(tabs);
    nfw_tables_load(&ntab);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int i = nn + threadIdx.x; i < nn_pad; i += blockDim.x) {     // padding: valid shape, zero weight
#pragma unroll
        for (int f = 0; f < NODE_FIELDS; ++f) srec[f * nn_pad + i] = (f >= NF_W_HM) ? 0.0 : srec[f * nn_pad + nn - 1];
    }
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const double* __restrict__ s_cp = srec + NF_CP * nn_pad;
    const double* __restrict__ s_rs = srec + NF_RS * nn_pad;
    const double* __restrict__ s_lr = srec + NF_LNRS * nn_pad;
    const double* __restrict__ s_hm = srec + NF_W_HM * nn_pad;
    const double* __restrict__ s_pm = srec + NF_W_PMM * nn_pad;
    const double* __restrict__ s_hg = srec + NF_W_HG * nn_pad;
    const double* __restrict__ s_gm = srec + NF_W_GM * nn_pad;
    const double* __restrict__ s_gg = srec + NF_W_GG * nn_pad;
    // ---- series region ---------------------------------------------------------------------
    // ser_n leading nodes get series coefficients (as many as the shared memory left over holds);
    // i_lo of them satisfy k r_vir <= SER_X for the largest k of the chunk and go into the moments.
    double* s_a = srec + NODE_FIELDS * nn_pad;                        // [SER_NC][ser_n]
    int ser_n = (smem_doubles - NODE_FIELDS * nn_pad - SUMS_EXTRA_DOUBLES) / SER_NC;
    ser_n = cfg.exclusion ? 0 : min(nn_pad, ser_n & ~31);
    if (ser_n < 0) ser_n = 0;
    double* s_mom = s_a + SER_NC * ser_n;                              // [SER_MOM]
    double* s_part = s_mom + SER_MOM;                                  // [8 warps][SER_MOM]
    const double k_hi = exp((k_end - 1 == nk - 1) ? l1 : l0 + hk * (k_end - 1));
    const double k_lo = exp(l0 + hk * k_begin);
    if (threadIdx.x == 0) s_first_bad = ser_n;
    __syncthreads();
    int i_lo = 0;
    if (ser_n > 0) {
        // pass 1: coefficients for every node some k of the chunk can use, and the end of the prefix
        for (int i = threadIdx.x; i < ser_n; i += blockDim.x) {
            const double cp = s_cp[i], c = cp - 1.0, q = s_rs[i] * c;
            const bool ok = c >= SER_MIN_C;
            if (!(ok && q * k_hi <= SER_X)) atomicMin(&s_first_bad, i);
            if (ok && q * k_lo <= SER_X) {
                double a[SER_NC];
                nfw_series_coeffs(c, cp, log(cp), a);
#pragma unroll
                for (int n = 0; n < SER_NC; ++n) s_a[n * ser_n + i] = a[n];
            } else {
                s_a[i] = nan("");          // a_0 = NaN marks "no series for this node"
            }
        }
        __syncthreads();
        i_lo = s_first_bad;
        // pass 2: moments of the prefix [0, i_lo).  Three orders x five sums at a time go through
        // one 16-value warp fold (node order inside a warp, then warp order: bit-reproducible).
        for (int idx = lane; idx < SER_MOM; idx += 32) s_part[w * SER_MOM + idx] = 0.0;
        __syncwarp();
        for (int base = 0; base < i_lo; base += blockDim.x) {      // uniform trip count
            const int i = base + threadIdx.x;
            const bool in = i < i_lo;
            const int ii = in ? i : 0;
            if (base + 32 * w < i_lo) {                             // warp-uniform: this warp has nodes
                double a[SER_NC];
#pragma unroll
                for (int n = 0; n < SER_NC; ++n) a[n] = s_a[n * ser_n + ii];
                const double q = s_rs[ii] * (s_cp[ii] - 1.0) * k_hi, u = q * q;
                const double whm = in ? s_hm[ii] : 0.0, wpm = in ? s_pm[ii] : 0.0, whg = in ? s_hg[ii] : 0.0;
                const double wgm = in ? s_gm[ii] : 0.0, wgg = in ? s_gg[ii] : 0.0;
                const double agm = fabs(wgm), agg = fabs(wgg);
                double pw = 1.0;
#pragma unroll
                for (int r = 0; r < SER_ROUNDS; ++r) {
                    double v[16];
                    v[15] = 0.0;
#pragma unroll
                    for (int m = 0; m < 3; ++m) {
                        const int n = 3 * r + m;
                        double an = 0.0, bn = 0.0;
                        if (n < SER_NC) {
#pragma unroll
                            for (int j = 0; j <= n; ++j) bn = fma(a[j], a[n - j], bn);
                            an = a[n] * pw;
                            bn *= pw;
                            pw *= u;
                        }
                        v[5 * m + 0] = whm * an;
                        v[5 * m + 1] = wpm * bn;
                        v[5 * m + 2] = whg * an;
                        v[5 * m + 3] = agm * ((wgm < 0.0) ? an : bn);
                        v[5 * m + 4] = agg * ((wgg < 0.0) ? an : bn);
                    }
                    const double tot = warp_fold16(v);
                    if ((lane & 1) == 0) s_part[w * SER_MOM + 16 * r + (lane >> 1)] += tot;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < SER_MOM) {
            double v = 0.0;
            for (int ww = 0; ww < nwarp; ++ww) v += s_part[ww * SER_MOM + threadIdx.x];
            s_mom[threadIdx.x] = v;
        }
    }
    __syncthreads();
    const double ik_hi = 1.0 / k_hi;
    // Large chunks: four ln k nodes per warp pass, eight lanes each striding the nodes -- everything
    // below is per-lane (no warp-uniform dispatch), the per-k epilogue is shared by four k and few
    // lanes idle when only a handful of nodes lie beyond the moment prefix.  Small chunks (the two
    // fine k classes): one k per warp, so that every warp of the CTA has work.
    const int kl = (k_end - k_begin >= 64) ? 8 : 32;             // lanes per ln k node
    const int kpw = 32 / kl, grp = lane / kl, gl = lane & (kl - 1);
    for (int ik0 = k_begin + kpw * w; ik0 < k_end; ik0 += kpw * nwarp) {
        const int ik = ik0 + grp;
        const bool live = ik < k_end;
        const int ikc = live ? ik : k_end - 1;
        const double lnk = (ikc == nk - 1) ? l1 : l0 + hk * ikc;                    // halo.py:49-51
        const double k = exp_fast(lnk);
        double a_hm = 0.0, a_pmm = 0.0, a_hg = 0.0, a_gm = 0.0, a_gg = 0.0;
        for (int base = i_lo; base < nn_pad; base += kl) {
            const bool valid = base + gl < nn_pad;
            const int i = valid ? base + gl : nn_pad - 1;
            const double cp = s_cp[i], rs = s_rs[i];
            const double z = k * rs;
            const double zc = z * (cp - 1.0);
            double rho = 0.0;
            bool ser = !valid;                                   // lanes past the end: nothing to do, rho = 0
            if (valid && i < ser_n && zc <= SER_X) {
                double p = s_a[SER_DEG * ser_n + i];
                ser = s_a[i] == s_a[i];
                const double t = zc * zc;
#pragma unroll
                for (int n = SER_DEG - 1; n >= 0; --n) p = fma(p, t, s_a[n * ser_n + i]);
                if (ser) rho = p;
            }
            if (!ser) rho = nfw_rho_tab(&ntab, z, cp, lnk + s_lr[i]);               // halo.py:574-583
            const double rho2 = rho * rho;
            double rho_h = rho;
            if (cfg.exclusion) rho_h = rho * exclusion_window(tabs, 2.0 * z * (cp - 1.0));
            a_hm = fma(s_hm[i], rho_h, a_hm);
            a_pmm = fma(s_pm[i], rho2, a_pmm);
            a_hg = fma(s_hg[i], rho_h, a_hg);
            const double wgm = s_gm[i], wgg = s_gg[i];
            a_gm = fma(fabs(wgm), (wgm < 0.0) ? rho : rho2, a_gm);
            a_gg = fma(fabs(wgg), (wgg < 0.0) ? rho : rho2, a_gg);
        }
        double v8[8] = {a_hm, a_pmm, a_hg, a_gm, a_gg, 0.0, 0.0, 0.0};
        double v;
        int s5;                                                   // which of the five sums this lane holds
        if (kl == 8) { v = group8_fold8(v8); s5 = gl; }           // lane s of the group holds sum s
        else { v = warp_fold8(v8); s5 = (lane & 3) ? 8 : (lane >> 2); }   // lane 4 s holds sum s
        if (s5 < 5 && live) {
            if (i_lo > 0) {
                const double r = k * ik_hi, sc = r * r;
                double p = s_mom[16 * (SER_DEG / 3) + 5 * (SER_DEG % 3) + s5];
#pragma unroll
                for (int n = SER_DEG - 1; n >= 0; --n) p = fma(p, sc, s_mom[16 * (n / 3) + 5 * (n % 3) + s5]);
                v += p;
            }
            raw[((size_t)b * 5 + s5) * nk + ik] = v;
        }
    }
}

// Normalise (halo.py:919, 958, 988, 1028, 1074) and spline the five tables in ln k
// (uniform knots).  One CTA of five warps per point, one warp per table: loads, second
// differences, coefficients and stores are lane-parallel and coalesced; only the two sweeps of
// the tridiagonal solve (a serial chain) run on lane 0, in shared memory.
__global__ void __launch_bounds__(160)
halo_splines_kernel(const Cfg cfg, int B, const double* __restrict__ raw, const double* __restrict__ nbar,
                    const double* __restrict__ epoch, double* __restrict__ tab /* [B,5,n_halo] */,
                    double* __restrict__ coef /* [B,5,4 n_halo] */, int32_t* __restrict__ status,
                    const int32_t* __restrict__ group) {
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    if (b >= B) return;
    const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = cfg.n_halo;
    double* sy = sm + (size_t)t * 3 * n;   // node values
    double* sm2 = sy + n;                  // second derivatives
    const double h = (log(cfg.k_max) - log(cfg.k_min)) / (n - 1);
    const double nb = nbar[b], rho_bar = epoch[(size_t)(group ? group[b] : b) * CHOMP_EPOCH_LEN + EP_RHO_BAR];
    const double scale = t < 2 ? 1.0 : (t == 2 ? 1.0 / nb : (t == 3 ? 1.0 / (nb * rho_bar) : 1.0 / (nb * nb * rho_bar)));
    const double* __restrict__ y_in = raw + ((size_t)b * 5 + t) * n;
    double* __restrict__ y_out = tab + ((size_t)b * 5 + t) * n;
    double* __restrict__ c = coef + ((size_t)b * 5 + t) * 4 * n;
    bool bad = false;
    for (int i = lane; i < n; i += 32) {
        const double v = y_in[i] * scale;
        sy[i] = v;
        y_out[i] = v;
        if (!isfinite(v)) bad = true;
    }
    if (bad && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
    __syncwarp();
    // not-a-knot spline on the uniform ln k grid: division-free sweeps (spline.cuh)
    spline_build_uniform_warp(n, h, sy, c, sm2);
}

}  // namespace chomp
