// K4: per-profile thermodynamic integrals (trapezoid rule over xi^3), one CTA per shell.
//   va_kinetic_energy_density     pttools/bubble/thermo.py:169-181
//   va_trace_anomaly_diff         thermo.py:213-221
//   va_thermal_energy_density_diff thermo.py:206-210
//   va_entropy_density_diff       thermo.py:156-166
//   wbar                          thermo.py:139-149
//   ubarf_squared / mean_enthalpy_change (old bag API)   pttools/bubble/quantities.py:409-445, 510-527
// The 4 pi / 3 prefactors and the algebra on the integrals stay on the host (a handful of scalars per point).
#include "ptt_host.cuh"

namespace ptt {

constexpr int TH_OUT = 8;

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// pp2[P,12] = {v_wall, mu_s, mu_b, V_s, V_b, a_s, a_b, T_ref, phase_rule, 0, 0, 0}
// out[P,8]  = {I_kin, I_theta, I_dw, I_ds, wbar, i_wall, n_broken, w_last}
__global__ void __launch_bounds__(256) profile_integrals_kernel(
    const double* __restrict__ rows_v, const double* __restrict__ rows_w, const double* __restrict__ rows_xi,
    const long long row_stride, const int* __restrict__ off, const int* __restrict__ npts,
    const double* __restrict__ pp2, double* __restrict__ out) {
    __shared__ double s_part[5][8];
    __shared__ int s_iwall, s_imax;
    const int p = blockIdx.x;
    const int tid = threadIdx.x;
    const int N = npts[p];
    double* o = out + (long long)p * TH_OUT;
    if (N < 3) {
        if (tid < TH_OUT) o[tid] = NAN;
        return;
    }
    const long long base = (long long)p * row_stride + off[p];
    const double* v = rows_v + base;
    const double* w = rows_w + base;
    const double* xi = rows_xi + base;
    const double v_wall = pp2[12 * p + 0], mu_s = pp2[12 * p + 1], mu_b = pp2[12 * p + 2];
    const double V_s = pp2[12 * p + 3], V_b = pp2[12 * p + 4], a_s = pp2[12 * p + 5], a_b = pp2[12 * p + 6];
    const double T_ref = pp2[12 * p + 7];
    const int phase_rule = (int)pp2[12 * p + 8];
    if (tid == 0) { s_iwall = N; s_imax = -1; }
    __syncthreads();
    const double w_last = w[N - 1];
    {
        int iw = N, im = -1;
        for (int k = tid; k < N; k += blockDim.x) {
            if (xi[k] >= v_wall && k < iw) iw = k;
            if (w[k] != w_last && k > im) im = k;
        }
        atomicMin(&s_iwall, iw);
        atomicMax(&s_imax, im);
    }
    __syncthreads();
    int i_wall = s_iwall;
    if (i_wall >= N) i_wall = 0;
    int n_broken = -1;
    if (phase_rule == 1) {                                   // props.find_phase, props.py:22-32
        if (i_wall == 0) n_broken = 0;
        else {
            n_broken = i_wall - 1;
            if (fabs(xi[i_wall] - v_wall) <= 1e-8 + 1e-5 * fabs(v_wall)) n_broken = i_wall;
        }
    }
    // wbar: i_max = last index with w != w[-1]  (thermo.py:141-145); if that is index 0 -> -1 (whole array)
    int i_max = s_imax;
    if (i_max <= 0) i_max = N - 1;

    const double theta_n = 0.25 * (1.0 - 4.0 / mu_s) * w_last + V_s;        // theta(w_n, symmetric)
    // entropy s(w, phase) = mu a (T/T0)^(mu-4) T^3,  T = T0 (w / (mu a T0^4))^(1/mu)   (const_cs.py:623-672; mu=4: bag.py)
    // (mu = 4 in both phases - the bag equation of state - needs no pow: T = T_ref (w / (4 a T_ref^4))^(1/4), s = 4 a T^3)
    const bool bag_like = (mu_s == 4.0) && (mu_b == 4.0);
    auto entropy = [&](const double ww, const bool broken) {
        const double mu = broken ? mu_b : mu_s, a = broken ? a_b : a_s;
        const double arg = ww / (mu * a * T_ref * T_ref * T_ref * T_ref);
        if (bag_like) {
            const double T = T_ref * sqrt(sqrt(arg));
            return 4.0 * a * T * T * T;
        }
        // (T/T0)^(mu-4) T^3 with T = T0 arg^(1/mu) is T0^3 arg^(1 - 1/mu): one pow per sample instead of two
        return mu * a * (T_ref * T_ref * T_ref) * pow(arg, 1.0 - 1.0 / mu);
    };
    const double s_n = entropy(w_last, false);

    // np.trapezoid(f, xi^3) as a weighted sum over samples, weight_k = (xi_{k+1}^3 - xi_{k-1}^3) / 2 (one-sided at the
    // ends): every integrand is evaluated once per sample
    double a_kin = 0.0, a_th = 0.0, a_dw = 0.0, a_ds = 0.0, a_wb = 0.0;
    for (int k = tid; k < N; k += blockDim.x) {
        const double xk = xi[k];
        const double xm = (k > 0) ? xi[k - 1] : xk, xp = (k < N - 1) ? xi[k + 1] : xk;
        const double x3m = xm * xm * xm, x3p = xp * xp * xp;
        const double wgt = 0.5 * (x3p - x3m);
        const double vk = v[k], wk = w[k];
        const bool bk = (phase_rule == 1) ? (k < n_broken) : (xk < v_wall);
        a_kin += wgt * (wk * vk * vk / (1.0 - vk * vk));
        const double th = (bk ? 0.25 * (1.0 - 4.0 / mu_b) * wk + V_b : 0.25 * (1.0 - 4.0 / mu_s) * wk + V_s) - theta_n;
        a_th += wgt * th;
        a_dw += wgt * (wk - w_last);
        a_ds += wgt * (entropy(wk, bk) - s_n);
        if (k <= i_max) {
            const double x3p_wb = (k < i_max) ? x3p : xk * xk * xk;      // the wbar integral stops at sample i_max
            a_wb += 0.5 * (x3p_wb - x3m) * wk;
        }
    }
    double vals[5] = {a_kin, a_th, a_dw, a_ds, a_wb};
    const int lane = tid & 31, wid = tid >> 5;
    for (int j = 0; j < 5; ++j) {
        const double r = warp_sum(vals[j]);
        if (lane == 0) s_part[j][wid] = r;
    }
    __syncthreads();
    if (tid == 0) {
        double tot[5];
        for (int j = 0; j < 5; ++j) {
            double s = 0.0;
            for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += s_part[j][q];
            tot[j] = s;
        }
        const double xm = xi[i_max];
        o[0] = tot[0]; o[1] = tot[1]; o[2] = tot[2]; o[3] = tot[3];
        o[4] = tot[4] / (xm * xm * xm);
        o[5] = (double)i_wall;
        o[6] = (double)n_broken;
        o[7] = w_last;
    }
}

}  // namespace ptt

extern "C" int ptt_profile_integrals_batch(int P, const double* rows_v, const double* rows_w, const double* rows_xi,
                                           int64_t row_stride, const int32_t* off, const int32_t* npts,
                                           const double* pp2, double* out, void* stream_) {
    if (P == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (P < 0 || !rows_v || !rows_w || !rows_xi || !off || !npts || !pp2 || !out)
        return ptt::set_error(PTT_E_ARG, "ptt_profile_integrals_batch: null pointer or bad size");
    if (P == 0) return PTT_OK;
    ptt::profile_integrals_kernel<<<P, 256, 0, (cudaStream_t)stream_>>>(rows_v, rows_w, rows_xi, row_stride, off, npts,
                                                                        pp2, out);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ptt::set_cuda_error("profile_integrals_kernel", e);
    return PTT_OK;
}
