// HALOFIT (Smith et al. 2003 with the Takahashi et al. 2012 coefficients) and the Limber
// C(l) integral.  Replaces (reference): HaloFit._initialize_halo_fit / _initialize_sigma_spline /
// power_mm / _delta2_Q / _delta2_H (halo.py:1261-1365), HaloFit.power_gm / power_gg
// (:1367-1412), CorrelationFourier.correlation / _correlation_integrand (correlation.py:360-392).
//
// n_eff and the curvature C are the first and second derivative of a k = 5 FITPACK
// interpolating spline of ln sigma^2(ln R) (halo.py:1289-1294); the same quintic spline
// (knots x_0 x6, x_3 ... x_{n-4}, x_{n-1} x6; collocation solve; B-spline derivative
// formulas) is built here.
#pragma once
#include "common.cuh"
#include "limber_tables.cuh"
#include "spline.cuh"

namespace chomp {

#define HF_LEN 16
enum { HF_KS = 0, HF_NEFF, HF_C, HF_AN, HF_BN, HF_CN, HF_GAMMA, HF_ALPHA, HF_BETA, HF_MU, HF_NU, HF_F1, HF_F2, HF_F3 };
#define HF_PANELS 32
#define HF_NQ 8

// values of the deg+1 B-splines of degree `deg` that are non-zero on [t[l], t[l+1]) at x:
// b[r] = B_{l-deg+r, deg}(x)   (de Boor's BSPLVB)
__device__ inline void bspline_values(const double* __restrict__ t, int deg, double x, int l, double* b) {
    double dr[6], dl[6];
    b[0] = 1.0;
    for (int j = 0; j < deg; ++j) {
        dr[j] = t[l + j + 1] - x;
        dl[j] = x - t[l - j];
        double saved = 0.0;
        for (int r = 0; r <= j; ++r) {
            const double term = b[r] / (dr[r] + dl[j - r]);
            b[r] = saved + dr[r] * term;
            saved = dl[j - r] * term;
        }
        b[j + 1] = saved;
    }
}

// First and second derivative at xq of the quintic interpolating spline through (x0 + i h, y[i]),
// i < n (n >= 8).  ONE thread.  work: (n + 6) + 9 n + n doubles.
__device__ inline void quintic_derivatives(int n, double x0, double h, const double* __restrict__ y, double xq,
                                           double* __restrict__ work, double& d1, double& d2) {
    double* t = work;                 // n + 6 knots
    double* band = t + n + 6;         // [n][9]: entry (i, j) at band[i * 9 + (j - i + 4)]
    double* c = band + 9 * n;         // coefficients
    const double xe = x0 + h * (n - 1);
    for (int i = 0; i < 6; ++i) { t[i] = x0; t[n + i] = xe; }
    for (int i = 6; i < n; ++i) t[i] = x0 + h * (i - 3);
    for (int i = 0; i < 9 * n; ++i) band[i] = 0.0;
    for (int i = 0; i < n; ++i) {
        const double x = (i == n - 1) ? xe : x0 + h * i;
        int l = i + 3;                 // x_i = t[i + 3] for 3 <= i <= n - 4
        if (l < 5) l = 5;
        if (l > n - 1) l = n - 1;      // last interval [t[n-1], t[n]] also serves x = xe
        double b[6];
        bspline_values(t, 5, x, l, b);
        for (int r = 0; r < 6; ++r) {
            const int j = l - 5 + r;
            const int off = j - i + 4;
            if (off >= 0 && off < 9) band[i * 9 + off] = b[r];
        }
        c[i] = y[i];
    }
    // banded elimination without pivoting (the collocation matrix is totally positive)
    for (int p = 0; p < n; ++p) {
        const double piv = band[p * 9 + 4];
        for (int i = p + 1; i <= p + 4 && i < n; ++i) {
            const double f = band[i * 9 + (p - i + 4)] / piv;
            if (f == 0.0) continue;
            for (int j = p; j <= p + 4 && j < n; ++j) {
                const int oi = j - i + 4, op = j - p + 4;
                if (oi >= 0 && oi < 9) band[i * 9 + oi] -= f * band[p * 9 + op];
            }
            c[i] -= f * c[p];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = c[i];
        for (int j = i + 1; j <= i + 4 && j < n; ++j) s -= band[i * 9 + (j - i + 4)] * c[j];
        c[i] = s / band[i * 9 + 4];
    }
    // knot interval of xq
    int l = 5;
    while (l < n - 1 && t[l + 1] <= xq) ++l;
    // s'(x) = sum c1_j B_{j,4}(x),  c1_j = 5 (c_j - c_{j-1}) / (t_{j+5} - t_j)
    // s''(x) = sum c2_j B_{j,3}(x), c2_j = 4 (c1_j - c1_{j-1}) / (t_{j+4} - t_j)
    double c1[6];                      // c1_{l-4 .. l}  (+ one extra below for c2)
    for (int r = 0; r < 5; ++r) {
        const int j = l - 4 + r;
        c1[r] = 5.0 * (c[j] - c[j - 1]) / (t[j + 5] - t[j]);
    }
    double b4[5], b3[4];
    bspline_values(t, 4, xq, l, b4);
    bspline_values(t, 3, xq, l, b3);
    d1 = 0.0;
    for (int r = 0; r < 5; ++r) d1 += c1[r] * b4[r];
    d2 = 0.0;
    for (int r = 0; r < 4; ++r) {
        const int j = l - 3 + r;       // c2_j uses c1_j (index r + 1) and c1_{j-1} (index r)
        const double c2 = 4.0 * (c1[r + 1] - c1[r]) / (t[j + 4] - t[j]);
        d2 += c2 * b3[r];
    }
}

// HaloFit.power_mm (halo.py:1325-1365); no k-range guards, as in the reference
__device__ __forceinline__ double halofit_power(const double* __restrict__ hf, const PkParams& pk, double k) {
    const double dk = delta2(pk, k, log(k));
    const double y = k / hf[HF_KS];
    const double dq = dk * (pow(1.0 + dk, hf[HF_BETA]) / (1.0 + hf[HF_ALPHA] * dk) * exp(-(y / 4.0 + y * y / 8.0)));
    double dh = hf[HF_AN] * pow(y, 3.0 * hf[HF_F1]) /
                (1.0 + hf[HF_BN] * pow(y, hf[HF_F2]) + pow(hf[HF_CN] * hf[HF_F3] * y, 3.0 - hf[HF_GAMMA]));
    dh = dh / (1.0 + hf[HF_MU] / y + hf[HF_NU] / (y * y));
    return 2.0 * M_PI * M_PI / (k * k * k) * (dq + dh);
}

// grid B, 256 threads.  fit_z < 0: f_1..f_3 at the epoch's own redshift.
__global__ void __launch_bounds__(256)
halofit_kernel(const Cfg cfg, int B, double fit_z, const double* __restrict__ cosmo, const double* __restrict__ epoch,
               double* __restrict__ hfit /* [B, HF_LEN] */, double* __restrict__ ln_sigma2_out /* [B, n_halo] */,
               int32_t* __restrict__ status) {
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    if (b >= B) return;
    const int n = cfg.n_halo, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    const int NQ = HF_PANELS * HF_NQ;
    double* s_k2 = sm;                 // k^2 at the quadrature nodes
    double* s_w = s_k2 + NQ;           // weight * Delta^2(k)
    double* ls2 = s_w + NQ;            // n   ln sigma^2
    double* lnR = ls2 + n;             // n
    double* rev_x = lnR + n;           // n   ln sigma^2 reversed (increasing)
    double* rev_y = rev_x + n;         // n
    double* coef = rev_y + n;          // 4 n
    double* work = coef + 4 * n;       // max(2 n, 11 n + 6)
    const Cosmo c = load_cosmo(cosmo + (size_t)b * CHOMP_N_COSMO, cfg.cosmo_precision);
    const double* e = epoch + (size_t)b * CHOMP_EPOCH_LEN;
    PkParams pk = make_pk(c, e[EP_GROWTH], e[EP_SIGMA_NORM]);
    CHOMP_ATTACH_BAO(cfg, c, pk)
    const double l0 = log(cfg.k_min), l1 = log(cfg.k_max);
    for (int idx = tid; idx < NQ; idx += blockDim.x) {
        const int p = idx / HF_NQ, q = idx - p * HF_NQ;
        const double a = l0 + (l1 - l0) * p / HF_PANELS, bb = l0 + (l1 - l0) * (p + 1) / HF_PANELS;
        const double half = 0.5 * (bb - a);
        const double x = 0.5 * (a + bb) + half * c_glx[HF_NQ][q];
        const double k = exp(x);
        s_k2[idx] = k * k;
        s_w[idx] = half * c_glw[HF_NQ][q] * delta2(pk, k, x);
    }
    __syncthreads();
    const double r0 = log(0.1), r1 = log(10.0), hR = (r1 - r0) / (n - 1);
    for (int i = w; i < n; i += nw) {                              // halo.py:1269-1283
        const double lr = (i == n - 1) ? r1 : r0 + hR * i;
        const double R2 = exp(2.0 * lr);
        double acc = 0.0;
        for (int idx = lane; idx < NQ; idx += 32) acc += s_w[idx] * exp(-s_k2[idx] * R2);
        acc = warp_sum(acc);
        if (lane == 0) { ls2[i] = log(acc); lnR[i] = lr; }
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        rev_x[i] = ls2[n - 1 - i];
        rev_y[i] = lnR[n - 1 - i];
        ln_sigma2_out[(size_t)b * n + i] = ls2[i];
    }
    __syncthreads();
    if (tid == 0) {
        double* h = hfit + (size_t)b * HF_LEN;
        spline_build(n, rev_x, rev_y, coef, work);                 // ln R as a function of ln sigma^2
        const double ln_r_star = spline_eval_search(coef, 0.0, rev_x, n);
        const double k_s = 1.0 / exp(ln_r_star);                   // halo.py:1284-1286
        double d1, d2;
        quintic_derivatives(n, r0, hR, ls2, log(1.0 / k_s), work, d1, d2);
        const double ne = -d1 - 3.0, C = -d2;                      // halo.py:1291-1294
        const double zf = fit_z < 0.0 ? e[EP_Z] : fit_z;
        const double om = omega_m_z(c, zf), ol = c.ol / E0(c, zf);
        const double ow = ol * (1.0 + (-1.0));                     // w = -1 on the supported domain
        h[HF_KS] = k_s; h[HF_NEFF] = ne; h[HF_C] = C;
        h[HF_AN] = pow(10.0, 1.5222 + 2.8553 * ne + 2.3706 * ne * ne + 0.9903 * ne * ne * ne +
                                 0.2250 * ne * ne * ne * ne - 0.6038 * C + 0.1749 * ow);
        h[HF_BN] = pow(10.0, -0.5642 + 0.5864 * ne + 0.5716 * ne * ne - 1.5474 * C + 0.2279 * ow);
        h[HF_CN] = pow(10.0, 0.3698 + 2.0404 * ne + 0.8161 * ne * ne + 0.5869 * C);
        h[HF_GAMMA] = 0.1971 - 0.0843 * ne + 0.8460 * C;
        h[HF_ALPHA] = fabs(6.0835 + 1.3373 * ne - 0.1959 * ne * ne - 5.5274 * C);
        h[HF_BETA] = 2.0379 - 0.7354 * ne + 0.3157 * ne * ne + 1.2490 * ne * ne * ne + 0.3980 * ne * ne * ne * ne - 0.1682 * C;
        h[HF_MU] = 0.0;
        h[HF_NU] = pow(10.0, 5.2105 + 3.6902 * ne);
        h[HF_F1] = pow(om, -0.0307); h[HF_F2] = pow(om, -0.0585); h[HF_F3] = pow(om, 0.0743);   // halo.py:1261-1264
        h[14] = ol; h[15] = zf;
        if (!(isfinite(k_s) && isfinite(ne) && isfinite(C)) && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
    }
}

// C(l) = int dchi P(l / chi) / D(z_bar)^2  W_a W_b D^2 / chi^2   (correlation.py:360-392) for the
// smooth spectra: linear_power and HaloFit power_mm.  grid B, 128 threads, one warp per l.
__global__ void __launch_bounds__(128)
cl_kernel(const Cfg cfg, int B, int use_halofit, int n_ell, const double* __restrict__ ell,
          const double* __restrict__ cosmo, const double* __restrict__ epoch, const double* __restrict__ dbar,
          const double* __restrict__ hfit, const double* __restrict__ grid0, const double* __restrict__ win_chi,
          const double* __restrict__ win_coef, const double* __restrict__ edges, const int32_t* __restrict__ n_edges,
          int edge_stride, double* __restrict__ out) {
    const int b = blockIdx.x;
    if (b >= B) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nz = cfg.n_cosmo, nwin = cfg.n_window, nq = cfg.nq_limber;
    const Cosmo c = load_cosmo(cosmo + (size_t)b * CHOMP_N_COSMO, cfg.cosmo_precision);
    const double* e = epoch + (size_t)b * CHOMP_EPOCH_LEN;
    PkParams pk = make_pk(c, e[EP_GROWTH], e[EP_SIGMA_NORM]);
    CHOMP_ATTACH_BAO(cfg, c, pk)
    const double* hf = hfit + (size_t)b * HF_LEN;
    LimberF F;
    F.g.n = nz; F.g.z_min = cfg.zk_min < 0.0 ? 0.0 : cfg.zk_min; F.g.z_max = cfg.zk_max;
    F.g.chi = const_cast<double*>(grid0) + (size_t)b * 13 * nz; F.g.c_chi_z = F.g.chi + nz;
    F.g.c_z_chi = F.g.chi + 5 * nz; F.g.c_g_z = F.g.chi + 9 * nz; F.g.z = nullptr; F.g.growth = nullptr;
    const double* wc = win_chi + (size_t)b * 4;
    F.a = Window{nwin, wc[0], wc[1], nullptr, const_cast<double*>(win_coef) + (size_t)b * 8 * nwin};
    F.b = Window{nwin, wc[2], wc[3], nullptr, const_cast<double*>(win_coef) + (size_t)b * 8 * nwin + 4 * nwin};
    const double* ed = edges + (size_t)b * edge_stride;
    const int n_pan = n_edges[b] - 1;
    const double inv_d2 = 1.0 / (dbar[b] * dbar[b]);
    // every panel cut in 4: P(l / chi) varies fast near the observer
    const int sub = 4;
    for (int il = w; il < n_ell; il += nw) {
        const double l = ell[il];
        double acc = 0.0;
        for (int idx = lane; idx < n_pan * sub * nq; idx += 32) {
            const int p = idx / (sub * nq), r = idx - p * sub * nq;
            const int s = r / nq, q = r - s * nq;
            const double a = ed[p], bb = ed[p + 1];
            const double pa = a + (bb - a) * s / sub, pb = a + (bb - a) * (s + 1) / sub;
            const double half = 0.5 * (pb - pa);
            const double chi = 0.5 * (pa + pb) + half * c_glx[nq][q];
            const double k = l / chi;
            const double P = use_halofit ? halofit_power(hf, pk, k) : 2.0 * M_PI * M_PI * delta2(pk, k, log(k)) / (k * k * k);
            acc += half * c_glw[nq][q] * P * F(chi) / (chi * chi);
        }
        acc = warp_sum(acc);
        if (lane == 0) out[(size_t)b * n_ell + il] = acc * inv_d2;
    }
}

}  // namespace chomp
