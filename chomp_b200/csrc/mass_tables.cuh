// Stage 2: SingleEpoch scalars, sigma(R), mass limits, nu(M) tables, Sheth-Tormen
// normalisations.  One CTA per parameter point; one warp per sigma(R) evaluation.
//
// Replaces (reference): SingleEpoch._initialize_defaults cosmology.py:93-119,
// sigma_r / _sigma_integrand cosmology.py:602-660, nu_m :687-699,
// MassFunction._set_mass_limits / _initialize_splines / _normalize
// mass_function.py:160-241.
#pragma once
#include "common.cuh"
#include "special.cuh"
#include "spline.cuh"

namespace chomp {

#define SIG_NQ 16          // Gauss-Legendre order per linear sigma(R) panel (W^2 oscillates: 2.5 periods per panel)
#define SIG_NQ_S 8         // order on the x < 1 and tail panels (smooth integrands; error < 1e-10 of sigma^2)
#define SIG_LOW_DL 2.0     // width in ln x of the fixed x < 1 panels [-2 (j + 1), -2 j]
#define SIG_LOW_MAX 6
#define SIG_XSPLIT 48.0    // beyond x = kR = 48 (96 for R >= 8 Mpc/h) only the non-oscillatory part of W^2 is integrated
#define SIG_DX 8.0         // panel width in x between 1 and SIG_XSPLIT
#define SIG_NLOW 4         // geometric panels below x_one when the fixed lattice does not apply
#define SIG_NTAIL 2        // geometric panels beyond SIG_XSPLIT

// squared top-hat window W^2(x), W = 3 (sin x / x^3 - cos x / x^2)   (cosmology.py:651-652)
__device__ __forceinline__ double tophat2(double x) {
    double W;
    if (x < 0.1) {
        const double x2 = x * x;
        W = 1.0 + x2 * (-0.1 + x2 * (1.0 / 280.0 + x2 * (-1.0 / 15120.0 + x2 / 1330560.0)));
    } else {
        double s, c;
        sincos_reduced(x, s, c);      // x = k R stays below ~1e4 here
        W = 3.0 * (s - x * c) / (x * x * x);
    }
    return W * W;
}

// The Gauss-Legendre nodes of the full-width panels [1 + 8 j, 9 + 8 j] in x = k R do not depend on
// R (nor on the parameter point): ln x and  W^2(x) dx/x  at those nodes are tabulated once, at
// create().  Only Delta^2(x / R) is left to evaluate per node.
#define SIG_LIN_MAX 12     // panels up to x = 97 (2 SIG_XSPLIT)
// two more rows: the truncated last panels [41, 48] and [89, 96] of the two split points
__device__ double g_sig_lnx[(SIG_LIN_MAX + 2) * SIG_NQ];
__device__ double g_sig_w2w[(SIG_LIN_MAX + 2) * SIG_NQ];
// same for the fixed panels below x = 1 (Gauss-Legendre in ln x): ln x and W^2(x) dln x
__device__ double g_sig_low_lnx[SIG_LOW_MAX * SIG_NQ_S];
__device__ double g_sig_low_w2w[SIG_LOW_MAX * SIG_NQ_S];
static inline cudaError_t chomp_upload_sigma_tables(const double* glx16, const double* glw16, const double* glx8,
                                                    const double* glw8) {
    static double lnx[(SIG_LIN_MAX + 2) * SIG_NQ], w2w[(SIG_LIN_MAX + 2) * SIG_NQ];
    static double llnx[SIG_LOW_MAX * SIG_NQ_S], lw2w[SIG_LOW_MAX * SIG_NQ_S];
    for (int j = 0; j < SIG_LOW_MAX; ++j)
        for (int q = 0; q < SIG_NQ_S; ++q) {
            const double half = 0.5 * SIG_LOW_DL;
            const double lx = -SIG_LOW_DL * j - half + half * glx8[q];
            const double x = exp(lx), x2 = x * x;
            const double W = x < 0.1 ? 1.0 + x2 * (-0.1 + x2 * (1.0 / 280.0 + x2 * (-1.0 / 15120.0 + x2 / 1330560.0)))
                                     : 3.0 * (sin(x) - x * cos(x)) / (x * x * x);
            llnx[j * SIG_NQ_S + q] = lx;
            lw2w[j * SIG_NQ_S + q] = W * W * half * glw8[q];
        }
    cudaError_t e0 = cudaMemcpyToSymbol(g_sig_low_lnx, llnx, sizeof llnx);
    if (e0 != cudaSuccess) return e0;
    e0 = cudaMemcpyToSymbol(g_sig_low_w2w, lw2w, sizeof lw2w);
    if (e0 != cudaSuccess) return e0;
    for (int j = 0; j < SIG_LIN_MAX + 2; ++j)
        for (int q = 0; q < SIG_NQ; ++q) {
            double a = 1.0 + SIG_DX * j, b = a + SIG_DX;
            if (j == SIG_LIN_MAX) { b = SIG_XSPLIT; a = b - fmod(b - 1.0, SIG_DX); }                 // [41, 48]
            if (j == SIG_LIN_MAX + 1) { b = 2.0 * SIG_XSPLIT; a = b - fmod(b - 1.0, SIG_DX); }       // [89, 96]
            const double half = 0.5 * (b - a);
            const double x = a + half + half * glx16[q];
            const double W = 3.0 * (sin(x) - x * cos(x)) / (x * x * x);
            lnx[j * SIG_NQ + q] = log(x);
            w2w[j * SIG_NQ + q] = W * W * half * glw16[q] / x;
        }
    cudaError_t e = cudaMemcpyToSymbol(g_sig_lnx, lnx, sizeof lnx);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(g_sig_w2w, w2w, sizeof w2w);
}

// Delta^2(k) providers for the sigma(R) integrals.  The mass-tables kernel evaluates ~14 000
// quadrature nodes per parameter point; it tabulates ln Delta^2 once on a fine uniform ln k grid
// in shared memory and interpolates (4-point Lagrange, error 0.0234 h^4 |d4f| < 3e-10 at
// h = 0.0101) instead of re-evaluating the transfer function at every node.
#define D2_TABLE_N 2048
struct D2Direct {
    PkParams pk;
    __device__ __forceinline__ double operator()(double k, double lnk) const { return delta2(pk, k, lnk); }
    __device__ __forceinline__ double at_lnk(double lnk) const { return delta2(pk, exp(lnk), lnk); }
};
struct D2Table {
    const double* tab;       // ln Delta^2 at l0 + j h
    double l0, inv_h;
    __device__ __forceinline__ double operator()(double k, double lnk) const {
        const double pos = (lnk - l0) * inv_h;
        int j = (int)pos;
        j = j < 1 ? 1 : (j > D2_TABLE_N - 3 ? D2_TABLE_N - 3 : j);
        const double u = pos - (double)j;          // in [0, 1) except at the clamped ends
        const double um = u + 1.0, u1 = u - 1.0, u2 = u - 2.0;
        const double f = tab[j - 1] * (-(1.0 / 6.0) * u * u1 * u2) + tab[j] * (0.5 * um * u1 * u2) +
                         tab[j + 1] * (-0.5 * um * u * u2) + tab[j + 2] * ((1.0 / 6.0) * um * u * u1);
        return exp_fast(f);
    }
    __device__ __forceinline__ double at_lnk(double lnk) const { return (*this)(0.0, lnk); }
};

// sigma^2(R) = int dlnk Delta^2(k) W^2(kR) over the reference's k range
// (cosmology.py:611-638); executed by one full warp, result in every lane.
// Partial sum (this thread's share) of
//   sigma^2(R) = int dlnk Delta^2(k) W^2(kR) over the reference's k range (cosmology.py:611-638).
// `rank` / `size`: position of the thread in the group (a warp or a slice of the CTA) that
// evaluates this sigma together; the caller reduces over the group.
// k range of the sigma(R) integrals with its logarithms (taken once per kernel)
struct SigLim {
    double k_min, k_max, ln_k_min, ln_k_max;
};
__device__ __forceinline__ SigLim make_siglim(double k_min, double k_max) { return SigLim{k_min, k_max, log(k_min), log(k_max)}; }

// lnR = ln(R) comes from the caller (who has ln M); every other logarithm the panels need follows
// from it and from constants.
template <class D2>
__device__ __noinline__ double sigma2_partial(const D2& pk, double R, double lnR, const SigLim& lim, int rank, int size) {
    // integration range rules, cosmology.py:611-629
    const double k_min = lim.k_min, k_max = lim.k_max;
    double k_lo = k_min, k_hi = k_max, ln_k_lo = lim.ln_k_min, ln_k_hi = lim.ln_k_max;
    const double iR = 1.0 / R;
    const double need_lo = iR / 10.0, need_hi = iR * 14.0662;
    if (need_lo <= k_lo) {
        if (need_lo > k_min / 100.0) { k_lo = need_lo; ln_k_lo = -2.3025850929940455 - lnR; }
        else { k_lo = k_min / 100.0; ln_k_lo = lim.ln_k_min - 4.605170185988092; }
    }
    if (need_hi >= k_hi) {
        if (need_hi < k_max * 100.0) { k_hi = need_hi; ln_k_hi = 2.643774756468092 - lnR; }     // ln 14.0662
        else { k_hi = k_max * 100.0; ln_k_hi = lim.ln_k_max + 4.605170185988092; }
    }
    const double x_lo = k_lo * R, x_hi = k_hi * R;
    // large spheres: nu ~ 40 there and ln f(nu) amplifies an error in sigma 14-fold, so the
    // oscillatory part is carried twice as far
    const double xs = fmin(R >= 8.0 ? 2.0 * SIG_XSPLIT : SIG_XSPLIT, x_hi);
    const double x_one = fmin(fmax(1.0, x_lo), xs);   // low (geometric) panels cover [x_lo, x_one]
    int n_lin = 0;
    if (xs > x_one) n_lin = (int)ceil((xs - x_one) / SIG_DX - 1e-9);
    if (n_lin < 0) n_lin = 0;
    const int n_tail = (x_hi > xs) ? SIG_NTAIL : 0;
    // logarithms of the panel ends
    const double l_lo = ln_k_lo + lnR, l_hi = ln_k_hi + lnR;
    const double l_s = (xs < x_hi) ? (R >= 8.0 ? 4.564348191467836 : 3.871201010907891) : l_hi;   // ln 96, ln 48
    const double l_one = (x_one == 1.0) ? 0.0 : (x_one == xs ? l_s : l_lo);
    const bool lattice = x_one == 1.0;     // the linear panels sit on the tabulated lattice
    // x < 1: fixed panels of the tabulated lattice down to the one that holds x_lo, which is cut
    // at x_lo and integrated on the spot
    int j_lo = 0;
    bool lat_low = lattice && x_lo < 1.0;
    if (lat_low) {
        j_lo = (int)floor(-l_lo * (1.0 / SIG_LOW_DL));
        if (j_lo >= SIG_LOW_MAX) lat_low = false;
    }
    const int n_low = lat_low ? j_lo + 1 : SIG_NLOW;
    // nodes are dealt out in slots of 8: one per low / tail panel, two per linear panel
    const int n_slots = n_low + 2 * n_lin + n_tail;
    const int n_nodes = n_slots * SIG_NQ_S;
    // the truncated last linear panel is on the lattice too when it ends at one of the two split points
    const int last_tab = !lattice ? -1
                         : ((xs == SIG_XSPLIT && n_lin == 6) ? SIG_LIN_MAX
                            : ((xs == 2.0 * SIG_XSPLIT && n_lin == 12) ? SIG_LIN_MAX + 1 : -1));
    double acc = 0.0;
    // One loop, ONE inlined copy of the Delta^2 interpolation (several copies of it thrash the instruction cache:
    // measured, +0.15 ms).  Flat node index t: first the nodes on the tabulated lattices -- rows 0 .. j_lo - 1 of
    // the x < 1 tables, the full-width rows of the linear tables, the truncated last linear panel when it ends
    // at a split point -- for which only Delta^2 is left to evaluate; then the nodes evaluated on the spot: the
    // cut (or non-lattice) low panels, linear panels off the lattice, the tail panels and the two end-point
    // terms of the oscillatory tail (two more tail nodes).
    const int n_low_tab = lat_low ? j_lo * SIG_NQ_S : 0;
    int lin_rows = 0;                                                            // full rows of the linear tables
    if (lattice && n_lin > 0) lin_rows = (n_lin - 1 < SIG_LIN_MAX) ? n_lin - 1 : SIG_LIN_MAX;
    const int n_lin_tab = lin_rows * SIG_NQ;
    const bool last_on_tab = last_tab >= 0;
    const int n_tab = n_low_tab + n_lin_tab + (last_on_tab ? SIG_NQ : 0);
    // ranges of the slot-of-8 node index `idx` the tables do not cover
    const int r0_lo = n_low_tab, r0_hi = n_low * SIG_NQ_S;                                    // low panels
    const int r1_lo = (n_low + 2 * lin_rows) * SIG_NQ_S;                                      // linear panels off the lattice
    const int r1_hi = (n_low + 2 * (last_on_tab ? n_lin - 1 : n_lin)) * SIG_NQ_S;
    const int r2_lo = (n_low + 2 * n_lin) * SIG_NQ_S, r2_hi = n_nodes + (n_tail ? 2 : 0);         // tail + end points
    const int c0 = r0_hi - r0_lo, c1 = (r1_hi > r1_lo ? r1_hi - r1_lo : 0), c2 = r2_hi - r2_lo;
    for (int t = rank; t < n_tab + c0 + c1 + c2; t += size) {
        double lnk, wv;
        if (t < n_tab) {
            const double* __restrict__ tl = g_sig_low_lnx;
            const double* __restrict__ tw = g_sig_low_w2w;
            int i = t;
            if (t >= n_low_tab) {
                tl = g_sig_lnx; tw = g_sig_w2w;
                i = t - n_low_tab;
                if (i >= n_lin_tab) i += last_tab * SIG_NQ - n_lin_tab;
            }
            lnk = tl[i] - lnR;
            wv = tw[i];
        } else {
            const int u = t - n_tab;
            const int idx = u < c0 ? r0_lo + u : (u < c0 + c1 ? r1_lo + (u - c0) : r2_lo + (u - c0 - c1));
            const int slot = idx >> 3, q8 = idx & 7;
            const bool edge = idx >= n_nodes;
            const bool lin = !edge && slot >= n_low && slot < n_low + 2 * n_lin;
            double x, wgt, w2;
            if (lin) {
                // Gauss-Legendre in x on [x_one + 8 j, x_one + 8 (j + 1)]: dlnk = dx / x
                const int jp = (slot - n_low) >> 1, q = ((slot - n_low) & 1) * SIG_NQ_S + q8;
                const double a = x_one + SIG_DX * jp;
                const double b = (jp == n_lin - 1) ? xs : a + SIG_DX;
                const double half = 0.5 * (b - a);
                x = 0.5 * (a + b) + half * c_glx[SIG_NQ][q];
                lnk = log(x) - lnR;
                wgt = half * c_glw[SIG_NQ][q] / x;
                w2 = tophat2(x);
            } else if (edge) {
                // first-order end-point term of the oscillatory tail:  +- F(x)/x [B sin 2x - C cos 2x] / 2,
                // B = 9 (x^2 - 1) / (2 x^6),  C = -9 / x^5, at x_hi (+) and at the split point (-)
                const bool top = idx == n_nodes;
                x = top ? x_hi : xs;
                lnk = (top ? l_hi : l_s) - lnR;
                const double x2 = x * x;
                double s2, c2;
                sincos_reduced(2.0 * x, s2, c2);
                w2 = 9.0 * (x2 - 1.0) / (2.0 * x2 * x2 * x2) * s2 + 9.0 / (x2 * x2 * x) * c2;
                wgt = (top ? 0.5 : -0.5) / x;
            } else {
                double a, b;
                const bool tail = slot >= n_low;
                if (!tail) {
                    if (lat_low) {
                        a = l_lo;
                        b = -SIG_LOW_DL * j_lo;
                    } else {
                        a = l_lo + (l_one - l_lo) * slot / SIG_NLOW;
                        b = l_lo + (l_one - l_lo) * (slot + 1) / SIG_NLOW;
                    }
                } else {
                    const int jp = slot - n_low - 2 * n_lin;
                    a = l_s + (l_hi - l_s) * jp / SIG_NTAIL;
                    b = l_s + (l_hi - l_s) * (jp + 1) / SIG_NTAIL;
                }
                const double half = 0.5 * (b - a);
                const double lx = 0.5 * (a + b) + half * c_glx[SIG_NQ_S][q8];
                x = exp_fast(lx);
                lnk = lx - lnR;
                wgt = half * c_glw[SIG_NQ_S][q8];
                if (tail) {
                    const double x2 = x * x;
                    w2 = 9.0 * (1.0 + x2) / (2.0 * x2 * x2 * x2);
                } else {
                    w2 = tophat2(x);
                }
            }
            wv = wgt * w2;
        }
        acc += wv * pk.at_lnk(lnk);
    }
    return acc;
}

// sigma^2(R) for a spectrum with baryon wiggles (SingleEpoch(with_bao=True), cosmology.py:474-538).  The wiggles
// (period 2 pi / s ~ 0.06 h/Mpc in k) sit exactly where W^2(kR) has its weight and the tabulated x lattices above
// cannot follow them: here the part below the split point is a plain composite rule in ln k -- pieces no wider than
// SIG_FINE_DL, Gauss-Legendre 4, Delta^2 evaluated directly -- and the tail beyond it is the smooth-spectrum one.
// About 2 000 nodes per call instead of 250: a variant, not the hot path.
#define SIG_FINE_DL 0.02
__device__ __noinline__ double sigma2_fine(const PkParams& pkp, double R, double lnR, const SigLim& lim, int rank, int size) {
    const double k_min = lim.k_min, k_max = lim.k_max;
    double k_lo = k_min, k_hi = k_max, ln_k_lo = lim.ln_k_min, ln_k_hi = lim.ln_k_max;
    const double iR = 1.0 / R;
    const double need_lo = iR / 10.0, need_hi = iR * 14.0662;           // cosmology.py:611-629
    if (need_lo <= k_lo) {
        if (need_lo > k_min / 100.0) { k_lo = need_lo; ln_k_lo = -2.3025850929940455 - lnR; }
        else { k_lo = k_min / 100.0; ln_k_lo = lim.ln_k_min - 4.605170185988092; }
    }
    if (need_hi >= k_hi) {
        if (need_hi < k_max * 100.0) { k_hi = need_hi; ln_k_hi = 2.643774756468092 - lnR; }
        else { k_hi = k_max * 100.0; ln_k_hi = lim.ln_k_max + 4.605170185988092; }
    }
    const double x_hi = k_hi * R;
    const double xs = fmin(R >= 8.0 ? 2.0 * SIG_XSPLIT : SIG_XSPLIT, x_hi);
    const double l_lo = ln_k_lo + lnR, l_hi = ln_k_hi + lnR, l_s = (xs < x_hi) ? log(xs) : l_hi;   // in ln x
    const int n_pan = (int)ceil((l_s - l_lo) / SIG_FINE_DL);
    const double d = (l_s - l_lo) / n_pan;
    double acc = 0.0;
    for (int idx = rank; idx < 4 * n_pan; idx += size) {
        const int p = idx >> 2, q = idx & 3;
        const double lx = l_lo + d * (p + 0.5) + 0.5 * d * c_glx[4][q];
        const double x = exp(lx);
        acc += 0.5 * d * c_glw[4][q] * delta2(pkp, x * iR, lx - lnR) * tophat2(x);
    }
    if (x_hi > xs) {
        // beyond the split point: the non-oscillatory part of W^2 on SIG_NTAIL geometric panels and the first-order
        // end-point terms of the oscillatory part, as in sigma2_partial
        for (int idx = rank; idx < SIG_NTAIL * SIG_NQ_S + 2; idx += size) {
            double x, lnk, wgt, w2;
            if (idx >= SIG_NTAIL * SIG_NQ_S) {
                const bool top = idx == SIG_NTAIL * SIG_NQ_S;
                x = top ? x_hi : xs;
                lnk = (top ? l_hi : l_s) - lnR;
                const double x2 = x * x;
                double s2, c2;
                sincos_reduced(2.0 * x, s2, c2);
                w2 = 9.0 * (x2 - 1.0) / (2.0 * x2 * x2 * x2) * s2 + 9.0 / (x2 * x2 * x) * c2;
                wgt = (top ? 0.5 : -0.5) / x;
            } else {
                const int jp = idx >> 3, q8 = idx & 7;
                const double a = l_s + (l_hi - l_s) * jp / SIG_NTAIL, b = l_s + (l_hi - l_s) * (jp + 1) / SIG_NTAIL;
                const double half = 0.5 * (b - a);
                const double lx = 0.5 * (a + b) + half * c_glx[SIG_NQ_S][q8];
                x = exp(lx);
                lnk = lx - lnR;
                wgt = half * c_glw[SIG_NQ_S][q8];
                const double x2 = x * x;
                w2 = 9.0 * (1.0 + x2) / (2.0 * x2 * x2 * x2);
            }
            acc += wgt * delta2(pkp, x * iR, lnk) * w2;
        }
    }
    return acc;
}

// one full warp; result in every lane
template <class D2>
__device__ inline double warp_sigma2(const D2& pk, double R, double lnR, const SigLim& lim) {
    return warp_sum(sigma2_partial(pk, R, lnR, lim, threadIdx.x & 31, 32));
}
__device__ inline double warp_sigma2(const PkParams& pk, double R, double k_min, double k_max) {
    if (pk.bao) return warp_sum(sigma2_fine(pk, R, log(R), make_siglim(k_min, k_max), threadIdx.x & 31, 32));
    return warp_sigma2(D2Direct{pk}, R, log(R), make_siglim(k_min, k_max));
}

// A "team" is one half of the 256-thread CTA (4 warps) with its own named barrier, so that
// the two mass-limit walks run side by side.
struct Team {
    int id, rank;        // 0 / 1, thread rank inside the team
    double* red;         // 8 doubles of shared scratch per team
};
#define TEAM_SIZE 128
__device__ __forceinline__ void team_sync(const Team& t) {
    asm volatile("bar.sync %0, %1;" ::"r"(t.id + 1), "r"(TEAM_SIZE) : "memory");
}
__device__ inline double team_sum(const Team& t, double v) {
    v = warp_sum(v);
    team_sync(t);
    if ((t.rank & 31) == 0) t.red[t.rank >> 5] = v;
    team_sync(t);
    return (t.red[0] + t.red[1]) + (t.red[2] + t.red[3]);
}
template <class D2>
__device__ inline double team_sigma2(const Team& t, const D2& pk, double R, double lnR, const SigLim& lim) {
    return team_sum(t, sigma2_partial(pk, R, lnR, lim, t.rank, TEAM_SIZE));
}


struct MassCtx {
    const PkParams* fine;   // non-null: spectrum with baryon wiggles, every sigma(R) through sigma2_fine
    D2Table pk;
    SigLim lim;
    double delta_c, rho_bar;
    double ln_r_coef;      // ln(3 / (4 pi rho_bar)): ln R = (ln M + ln_r_coef) / 3  (cosmology.py:662-672)
};
// the share of thread `rank` of `size` in sigma^2(R), whichever rule the epoch asks for.  VAR = false: the default
// configuration (zero-baryon transfer function, Sheth-Tormen) -- nothing of the variants is compiled into that
// instantiation (their out-of-line calls alone grew the stack frame of the kernel and put 0.4 GB of local-memory traffic
// on a launch: ncu r2z)
template <bool VAR>
__device__ __forceinline__ double sigma2_share(const MassCtx& m, double R, double lnR, int rank, int size);

// nu(M) = (delta_c / sigma(M))^2 from ln M, warp-collective (cosmology.py:662-699)
template <bool VAR>
__device__ inline double warp_nu_lm(const MassCtx& m, double lm) {
    const double lnR = (lm + m.ln_r_coef) * (1.0 / 3.0);
    const double s2 = warp_sum(sigma2_share<VAR>(m, exp_fast(lnR), lnR, threadIdx.x & 31, 32));
    return m.delta_c * m.delta_c / s2;
}
// same, team-collective
template <bool VAR>
__device__ inline double team_nu_lm(const Team& t, const MassCtx& m, double lm) {
    const double lnR = (lm + m.ln_r_coef) * (1.0 / 3.0);
    const double s2 = team_sum(t, sigma2_share<VAR>(m, exp_fast(lnR), lnR, t.rank, TEAM_SIZE));
    return m.delta_c * m.delta_c / s2;
}

// The reference walks a mass limit in factors of 1.05 until nu(M) enters a window
// (mass_function.py:172-194).  nu(M) is monotone, so the walk stops at the first step j whose
// mass passes the window edge `thr`; that step is found here by solving ln nu(M0 e^t) = ln thr
// with the secant method (ln nu is almost linear in ln M) and then checking the two
// neighbouring steps with the reference's own comparison.  Team-collective (every sigma(R) is
// spread over 128 threads: the walk is a serial chain, so latency is what counts); returns the
// number of steps (0 = already inside), or -1 if no step within +-J satisfies it.
template <bool VAR>
__device__ inline int team_walk(const Team& tm, const MassCtx& m, double nu_scale, double M0, double nu0,
                                double lo_edge, double hi_edge, int J, double* mass_out) {
    *mass_out = M0;
    int dir;          // +1: multiply by 1.05, -1: divide
    double thr;
    bool want_le;
    if (hi_edge < nu0) { dir = -1; thr = hi_edge; want_le = true; }        // "too high": M /= 1.05 until nu <= hi_edge
    else if (lo_edge > nu0) { dir = +1; thr = lo_edge; want_le = false; }  // "too low":  M *= 1.05 until nu >= lo_edge
    else return 0;
    const double ln_step = log(1.05), target = log(thr), ln_M0 = log(M0);
    // secant on g(t) = ln nu(M0 e^t) - target
    double t0 = 0.0, g0 = log(nu0) - target;
    double t1 = -g0 / 0.35;                       // d ln nu / d ln M is 0.15 ... 0.7
    double g1 = 0.0;
    const double t_max = J * ln_step;
    // sigma(R) carries quadrature noise of ~1e-9, and only the step index is needed: the search
    // stops as soon as the root t* is pinned between two steps of the walk.  ln nu is nearly linear
    // in ln M, so the secant-corrected estimate t_c = t1 - g1 / slope is off by far less than the
    // correction itself; a quarter of it (plus the noise floor) is taken as the bound.
    double slope = 0.35;
    int j = 0;
    bool certain = false;
    for (int it = 0; it < 10; ++it) {
        t1 = fmax(-t_max, fmin(t_max, t1));
        g1 = log(team_nu_lm<VAR>(tm, m, ln_M0 + t1) * nu_scale) - target;
        if (it > 0 && t1 != t0) slope = (g1 - g0) / (t1 - t0);
        const double corr = (fabs(slope) > 0.05) ? g1 / slope : g1 / 0.05;
        const double tc = t1 - corr;
        const double s = fabs(tc) / ln_step, fl = floor(s);
        const double d = fmin(s - fl, fl + 1.0 - s) * ln_step;
        if (it > 0 && tc * dir > 0.0 && fabs(tc) < t_max && d > 4.0 * (0.25 * fabs(corr) + 2e-7)) {
            j = (int)fl + 1;                 // first step beyond the root
            certain = true;
            break;
        }
        if (fabs(g1) < 2e-8 || fabs(t1 - t0) < 1e-7) break;
        const double t2 = t1 - g1 * (t1 - t0) / (g1 - g0);
        t0 = t1; g0 = g1; t1 = t2;
    }
    if (!certain) {
        j = (int)ceil(fabs(t1) / ln_step - 1e-7);
        if (j < 1) j = 1;
        // settle on the first step that passes, exactly as the sequential walk would
        for (int guard = 0; guard < 8; ++guard) {
            if (j > J) return -1;
            const double nu_j = team_nu_lm<VAR>(tm, m, ln_M0 + dir * j * ln_step) * nu_scale;
            const bool ok_j = want_le ? (nu_j <= thr) : (nu_j >= thr);
            if (!ok_j) { ++j; continue; }
            if (j == 1) break;
            const double nu_p = team_nu_lm<VAR>(tm, m, ln_M0 + dir * (j - 1) * ln_step) * nu_scale;
            const bool ok_p = want_le ? (nu_p <= thr) : (nu_p >= thr);
            if (ok_p) { --j; continue; }
            break;
        }
    }
    *mass_out = M0 * pow(1.05, (double)(dir * j));
    return j;
}

template <bool VAR>
__device__ __forceinline__ double sigma2_share(const MassCtx& m, double R, double lnR, int rank, int size) {
    if (VAR && m.fine) return sigma2_fine(*m.fine, R, lnR, m.lim, rank, size);
    return sigma2_partial(m.pk, R, lnR, m.lim, rank, size);
}

struct MassOut {
    double* epoch;     // [B, CHOMP_EPOCH_LEN]
    double* lnm_nodes; // [B, n_mass]
    double* nu_nodes;  // [B, n_mass]
    double* c_lnm_nu;  // [B, 4 n_mass]  ln M as a function of nu
    double* c_nu_lnm;  // [B, 4 n_mass]  nu as a function of ln M
};

#ifndef MASS_MIN_BLOCKS
#define MASS_MIN_BLOCKS 4
#endif
template <bool VAR>
__global__ void __launch_bounds__(256, MASS_MIN_BLOCKS)
mass_tables_kernel(const Cfg cfg, int B, const double* __restrict__ cosmo, const double* __restrict__ halo,
                   const double* __restrict__ z_in, const double* __restrict__ zbar, MassOut out,
                   int32_t* __restrict__ status) {
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    if (b >= B) return;
    const int n = cfg.n_mass;
    double* lnm = sm;            // n
    double* nu = lnm + n;        // n
    double* c1 = nu + n;         // 4n
    double* c2 = c1 + 4 * n;     // 4n
    double* work = c2 + 4 * n;   // 10n (two warp-built splines)
    double* red = work + 10 * n; // 64
    double* d2tab = red + 64;    // D2_TABLE_N
    const int tid = threadIdx.x, w = tid >> 5, nw = blockDim.x >> 5, lane = tid & 31;

    const Cosmo c = load_cosmo(cosmo + (size_t)b * CHOMP_N_COSMO, cfg.cosmo_precision);
    double z = z_in ? z_in[b] : zbar[b];
    if (z < 0.0) z = 0.0;                                          // cosmology.py:40-41
    const double* hp = halo + (size_t)b * CHOMP_N_HALO;
    const double growth = growth_approx(c, 1.0 / (1.0 + z)) / growth_approx(c, 1.0);
    MassCtx m;
    m.delta_c = delta_c_z(c, z);
    m.rho_bar = rho_bar_z(c, z);
    m.lim = make_siglim(cfg.k_min, cfg.k_max);
    m.ln_r_coef = log(3.0 / (4.0 * M_PI * m.rho_bar));
    // Every sigma(R) is evaluated with sigma_norm = 1; nu scales as 1 / sigma_norm^2
    // (cosmology.py:118-119, 574-587).  Round 0: sigma_8, nu(1e9) and nu(1e16) on three warps.
    PkParams pk1 = make_pk(c, growth, 1.0);
    BaoParams bao_store;
    m.fine = nullptr;
    if (VAR && cfg.with_bao) { make_bao(c, &bao_store); pk1.bao = &bao_store; m.fine = &pk1; }
    {
        // ln Delta^2 table over every k the sigma(R) range rules can reach: [k_min / 100, 100 k_max]
        const double t0 = log(cfg.k_min / 100.0) - 0.05, t1 = log(cfg.k_max * 100.0) + 0.05;
        const double th = (t1 - t0) / (D2_TABLE_N - 1);
        const double ln_amp = log(pk1.amp);
        for (int j = tid; j < D2_TABLE_N; j += blockDim.x) {
            const double lk = t0 + th * j;
            // ln Delta^2 = ln amp + (3 + n)(ln k - ln H0) + 2 ln T(k): no exp / log round trip
            d2tab[j] = ln_amp + pk1.expo * (lk - pk1.ln_H0) + 2.0 * log(VAR ? transfer_any(pk1, exp_fast(lk)) : transfer_eh(pk1, exp_fast(lk)));
        }
        m.pk.tab = d2tab; m.pk.l0 = t0; m.pk.inv_h = 1.0 / th;
    }
    __syncthreads();
    double m_lo = 1.0e9, m_hi = 1.0e16;
    const bool fixed_limits = cfg.mass_min > 0.0 && cfg.mass_max > 0.0;
    if (fixed_limits) { m_lo = cfg.mass_min; m_hi = cfg.mass_max; }
    // sigma_8, nu(1e9), nu(1e16) side by side on groups of 2 / 3 / 3 warps (warp partials in
    // red[8 ..], summed in warp order)
    {
        const int grp = (w < 2 || fixed_limits) ? 0 : (w < 5 ? 1 : 2);
        const int g_first = grp == 0 ? 0 : (grp == 1 ? 2 : 5);
        const int g_warps = fixed_limits ? nw : (grp == 0 ? 2 : 3);
        const double lnR = grp == 0 ? 2.0794415416798357 : (log(grp == 1 ? m_lo : m_hi) + m.ln_r_coef) * (1.0 / 3.0);
        const double R = grp == 0 ? 8.0 : exp_fast(lnR);
        const double part = warp_sum(sigma2_share<VAR>(m, R, lnR, tid - 32 * g_first, 32 * g_warps));
        if (lane == 0) red[8 + w] = part;
    }
    __syncthreads();
    if (tid == 0) {
        if (fixed_limits) {
            double v = 0.0;
            for (int i = 0; i < nw; ++i) v += red[8 + i];
            red[0] = v;
        } else {
            red[0] = red[8] + red[9];
            red[1] = m.delta_c * m.delta_c / ((red[10] + red[11]) + red[12]);
            red[2] = m.delta_c * m.delta_c / ((red[13] + red[14]) + red[15]);
        }
    }
    __syncthreads();
    const double s8_raw = sqrt(red[0]);
    const double sigma_norm = c.s8 * growth / s8_raw;
    const double nu_scale = 1.0 / (sigma_norm * sigma_norm);

    int st = 0;
    if (c.bad || hp[CHOMP_H_ALPHA] != -1.0) st |= CHOMP_ST_DOMAIN;
    int walk_steps = 0;
    // ---- mass limits (mass_function.py:160-203): the two walks are independent -------------
    if (!fixed_limits) {
        const double nu_lo0 = red[1] * nu_scale, nu_hi0 = red[2] * nu_scale;
        __syncthreads();
        Team tm{tid / TEAM_SIZE, tid % TEAM_SIZE, red + 16 + 8 * (tid / TEAM_SIZE)};
        double mm;
        int j;
        // one call site (two inlined copies of the walk are 2 x 600 instructions in a kernel that stalls on fetch)
        const bool low = tm.id == 0;
        j = team_walk<VAR>(tm, m, nu_scale, low ? m_lo : m_hi, low ? nu_lo0 : nu_hi0, (low ? 0.1 : 50.0) * (1.0 - 0.05),
                      (low ? 0.1 : 50.0) * (1.0 + 0.05), 512, &mm);
        if (tm.rank == 0) { red[4 + 2 * tm.id] = mm; red[5 + 2 * tm.id] = (double)j; }
        __syncthreads();
        m_lo = red[4]; m_hi = red[6];
        if (red[5] < 0.0 || red[7] < 0.0) st |= CHOMP_ST_MASS_WALK;
        walk_steps = (int)(fmax(red[5], 0.0) + fmax(red[7], 0.0));
        __syncthreads();
    }
    const double lnm_min = log(m_lo), lnm_max = log(m_hi);
    const double hM = (lnm_max - lnm_min) / (n - 1);
    // ---- nu at the mass nodes (mass_function.py:205-209) -----------------------------
    // the warps draw the nodes from a shared counter, heaviest (largest R: most panels) first
    __shared__ int next_node;
    if (tid == 0) next_node = 0;
    __syncthreads();
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&next_node, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n) break;
        const int i = n - 1 - t;
        const double lm = (i == n - 1) ? lnm_max : lnm_min + hM * i;
        const double v = warp_nu_lm<VAR>(m, lm) * nu_scale;
        if (lane == 0) { lnm[i] = lm; nu[i] = v; }
    }
    __syncthreads();
    // Two warps build the two splines; the other six meanwhile integrate the Sheth-Tormen
    // normalisations and the comoving distance, which need the nu nodes only.
    const double nu_min = 1.001 * nu[0], nu_max = 0.999 * nu[n - 1];  // mass_function.py:212-213
    const double stq = hp[CHOMP_H_STQ], sta = hp[CHOMP_H_ST_LITTLE_A];
    double sf = 0.0, sfb = 0.0, chi = 0.0;
    if (w == 0) spline_build_uniform_warp(n, hM, nu, c2, work);   // nu(ln M): uniform ln M grid
    else if (w == 1) spline_build_warp(n, nu, lnm, c1, work + 5 * n);  // ln M(nu)
    else {
        // ---- f_norm, bias_norm (mass_function.py:225-241): int f dnu = int nu f dln nu over
        //      [nu_min, nu_max], 8-point Gauss-Legendre on every knot interval of ln nu
        const int t2 = tid - 64, nt2 = blockDim.x - 64;
        // ln(nu) at the knots once (the Delta^2 table is no longer needed), ends moved to nu_min / nu_max
        double* lognu = d2tab;
        for (int i = t2; i < n; i += nt2) lognu[i] = log(i == 0 ? nu_min : (i == n - 1 ? nu_max : nu[i]));
        MfParams mf;
        mf.kind = CHOMP_MF_SHETH_TORMEN; mf.ln_sta = log(sta); mf.stq = stq; mf.delta_c = m.delta_c;
        if (VAR) {
            double dv_mf = hp[CHOMP_H_DELTA_V];
            if (dv_mf == -1.0) dv_mf = delta_v_z(c, z, growth);
            mf = mf_params(cfg, sta, stq, m.delta_c, dv_mf, z);
        }
        asm volatile("bar.sync 3, %0;" ::"r"(nt2) : "memory");
        for (int idx = t2; idx < (n - 1) * 8; idx += nt2) {
            const int i = idx >> 3, q = idx & 7;
            const double a = lognu[i], bb = lognu[i + 1];
            const double half = 0.5 * (bb - a);
            const double x = 0.5 * (a + bb) + half * c_glx[8][q];
            double nf, bi;
            if (VAR) mf_raw_ln(mf, x, nf, bi);
            else st_raw_ln(x, mf.ln_sta, stq, m.delta_c, nf, bi);
            const double wgt = half * c_glw[8][q];
            sf += wgt * nf;
            sfb += wgt * nf * bi;
        }
        // comoving distance to z (SingleEpoch._chi, cosmology.py:106-110): 8 panels x GL-8
        if (t2 < 64) {
            const int p = t2 >> 3, q = t2 & 7;
            const double a = z * p / 8.0, bb = z * (p + 1) / 8.0;
            const double half = 0.5 * (bb - a);
            chi = half * c_glw[8][q] * inv_hubble(c, 0.5 * (a + bb) + half * c_glx[8][q]);
        }
    }
    sf = block_sum(sf, red);
    sfb = block_sum(sfb, red + 32);
    chi = block_sum(chi, red);
    // Tinker's multiplicity function carries its fitted amplitude: only the bias is normalised (mass_function.py:528-542)
    const double f_norm = (VAR && cfg.mass_function_kind == CHOMP_MF_TINKER) ? 1.0 : 1.0 / sf;
    const double b_norm = 1.0 / (f_norm * sfb);
    const double lnm_star = spline_eval_search(c1, 1.0, nu, n);        // m_star = mass(1.0), :223

    // ---- write out ----------------------------------------------------------------------
    for (int i = tid; i < n; i += blockDim.x) {
        out.lnm_nodes[(size_t)b * n + i] = lnm[i];
        out.nu_nodes[(size_t)b * n + i] = nu[i];
    }
    for (int i = tid; i < 4 * (n - 1); i += blockDim.x) {
        out.c_lnm_nu[(size_t)b * 4 * n + i] = c1[i];
        out.c_nu_lnm[(size_t)b * 4 * n + i] = c2[i];
    }
    if (tid == 0) {
        double* e = out.epoch + (size_t)b * CHOMP_EPOCH_LEN;
        double dv = hp[CHOMP_H_DELTA_V];
        if (dv == -1.0) dv = delta_v_z(c, z, growth);           // mass_function.py:51-53, halo.py:73-75
        e[EP_Z] = z; e[EP_GROWTH] = growth; e[EP_SIGMA_NORM] = sigma_norm; e[EP_DELTA_C] = m.delta_c;
        e[EP_DELTA_V] = dv; e[EP_RHO_BAR] = m.rho_bar; e[EP_LNM_MIN] = lnm_min; e[EP_LNM_MAX] = lnm_max;
        e[EP_NU_MIN] = nu_min; e[EP_NU_MAX] = nu_max; e[EP_F_NORM] = f_norm; e[EP_B_NORM] = b_norm;
        e[EP_LNM_STAR] = lnm_star; e[EP_PK_AMP] = pk1.amp * sigma_norm * sigma_norm; e[EP_CHI] = chi; e[EP_WALK] = (double)walk_steps;
        e[EP_OMEGA_M] = omega_m_z(c, z); e[EP_OMEGA_L] = c.ol / E0(c, z); e[EP_E0] = E0(c, z);
        e[EP_DELTA_V_COSMO] = delta_v_z(c, z, growth);
        e[EP_RHO_CRIT] = 1.879 / 1.989 * (3.086 * 3.086 * 3.086) * 1e10 * E0(c, z);
        e[EP_FLAT] = c.flat; e[EP_OPEN] = c.open; e[EP_SIGMA_8_Z] = s8_raw * sigma_norm;
        if (!(isfinite(f_norm) && isfinite(b_norm) && isfinite(lnm_star))) st |= CHOMP_ST_NONFINITE;
        if (status && st) atomicOr(status + b, st);
    }
}

}  // namespace chomp
