// MassFunctionSecondOrder (reference mass_function.py:365-433): the sigma(nu) spline and the
// normalisation of the second-order Sheth-Tormen bias b_2(nu), on top of the stage-2 tables.
// One CTA (64 threads) per parameter point; not part of the w(theta) path.
#pragma once
#include "common.cuh"
#include "spline.cuh"

namespace chomp {

// b_2(nu) - bias_2_norm (mass_function.py:423-430)
__device__ __forceinline__ double bias2_raw(double nu, double sigma, double sta, double stq, double delta_c, double b_norm) {
    double nf, bi;
    st_raw(nu, sta, stq, delta_c, nf, bi);
    const double nup = nu * sta;
    return 8.0 / 21.0 * (bi * b_norm - 1.0) + (nu - 3.0) / (sigma * sigma) +
           2.0 * stq / (delta_c * delta_c * (1.0 + pow(nup, stq))) * (2.0 * stq + 2.0 * nup - 1.0);
}

__global__ void __launch_bounds__(64)
mass_second_order_kernel(const Cfg cfg, int B, const double* __restrict__ halo, const double* __restrict__ epoch,
                         const double* __restrict__ nu_nodes, double* __restrict__ sig_coef /* [B, 4 n_mass] */,
                         double* __restrict__ b2_norm /* [B] */, int32_t* __restrict__ status) {
    extern __shared__ double sm[];
    __shared__ double red[64];
    const int b = blockIdx.x;
    if (b >= B) return;
    const int n = cfg.n_mass, tid = threadIdx.x;
    double* nu = sm;             // n
    double* sg = nu + n;         // n
    double* coef = sg + n;       // 4 n
    double* work = coef + 4 * n; // 2 n
    const double* e = epoch + (size_t)b * CHOMP_EPOCH_LEN;
    const double* hp = halo + (size_t)b * CHOMP_N_HALO;
    const double delta_c = e[EP_DELTA_C];
    for (int i = tid; i < n; i += blockDim.x) {
        const double v = nu_nodes[(size_t)b * n + i];
        nu[i] = v;
        sg[i] = delta_c / sqrt(v);            // nu = (delta_c / sigma)^2, mass_function.py:375-379
    }
    __syncthreads();
    if (tid == 0) spline_build(n, nu, sg, coef, work);      // _sigma_spline, mass_function.py:389-390
    __syncthreads();
    // bias_2_norm = - int f(nu) b2_raw(nu) dnu over [nu_min, nu_max] (mass_function.py:413-420):
    // nu f dln nu, GL-8 on every knot interval of ln nu as for f_norm / bias_norm
    const double stq = hp[CHOMP_H_STQ], sta = hp[CHOMP_H_ST_LITTLE_A];
    const double l_min = log(e[EP_NU_MIN]), l_max = log(e[EP_NU_MAX]);
    double acc = 0.0;
    for (int idx = tid; idx < (n - 1) * 8; idx += blockDim.x) {
        const int i = idx >> 3, q = idx & 7;
        const double a = (i == 0) ? l_min : log(nu[i]);
        const double bb = (i == n - 2) ? l_max : log(nu[i + 1]);
        const double half = 0.5 * (bb - a);
        const double v = exp(0.5 * (a + bb) + half * c_glx[8][q]);
        double nf, bi;
        st_raw(v, sta, stq, delta_c, nf, bi);
        const double sigma = spline_eval_search(coef, v, nu, n);
        acc += half * c_glw[8][q] * nf * e[EP_F_NORM] * bias2_raw(v, sigma, sta, stq, delta_c, e[EP_B_NORM]);
    }
    acc = block_sum(acc, red);
    if (tid == 0) {
        b2_norm[b] = -acc;
        if (!isfinite(acc) && status) atomicOr(status + b, CHOMP_ST_NONFINITE);
    }
    for (int i = tid; i < 4 * (n - 1); i += blockDim.x) sig_coef[(size_t)b * 4 * n + i] = coef[i];
}

}  // namespace chomp
