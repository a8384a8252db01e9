// Post-processing epilogues of a sweep: the suppression factor of Omega_gw,0 and the LISA
// signal-to-noise ratio, evaluated on the device next to the spectra.
//   Suppression.suppression -> scipy.interpolate.griddata(method="linear")       pttools/omgw0/suppression/suppression.py:53-96
//   signal_to_noise_ratio(f, signal, noise) = sqrt(T_obs * trapz(signal^2 / noise^2, f))   pttools/omgw0/noise.py:7-52
//   Spectrum.signal_to_noise_ratio / _instrument                                 pttools/omgw0/omgw0_ssm.py:143-147
#include "ptt_host.cuh"

namespace ptt {

// Piecewise-linear interpolation on a triangulation (what griddata / LinearNDInterpolator evaluate): tri[t] = the three
// vertices (x0, y0, x1, y1, x2, y2) and their values; NaN outside the hull.  One thread per point, brute force over the
// (few dozen) triangles.  mode 0: no extension; mode 1: 1 below alpha_min (SuppressionMethod.EXT_CONSTANT).
__global__ void suppression_kernel(const int n, const double* __restrict__ vw, const double* __restrict__ an,
                                   const int n_tri, const double* __restrict__ tri, const double* __restrict__ val,
                                   const int mode, const double alpha_min, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = vw[i], y = an[i];
    double res = NAN, best = -INFINITY;
    for (int t = 0; t < n_tri; ++t) {
        const double* q = tri + 6 * t;
        const double x0 = q[0], y0 = q[1], x1 = q[2], y1 = q[3], x2 = q[4], y2 = q[5];
        const double det = (y1 - y2) * (x0 - x2) + (x2 - x1) * (y0 - y2);
        const double l0 = ((y1 - y2) * (x - x2) + (x2 - x1) * (y - y2)) / det;
        const double l1 = ((y2 - y0) * (x - x2) + (x0 - x2) * (y - y2)) / det;
        const double l2 = 1.0 - l0 - l1;
        const double lmin = fmin(l0, fmin(l1, l2));
        // inside (with the tolerance Qhull's point location uses for points on edges); the most interior triangle wins
        if (lmin >= -1e-12 && lmin > best) {
            best = lmin;
            res = l0 * val[3 * t] + l1 * val[3 * t + 1] + l2 * val[3 * t + 2];
        }
    }
    if (mode == 1 && y < alpha_min) res = 1.0;
    out[i] = res;
}

// snr[p, w] = sqrt( sum_k weight[w, k] * signal[p, k]^2 ),  weight = T_obs * trapezoid weight(f) / noise(f)^2.
// One warp per row, all weights in one pass over the row (HBM bound: 8 n_y bytes per point).
constexpr int SNR_MAXW = 4;
__global__ void __launch_bounds__(256) snr_kernel(const int P, const int n_y, const double* __restrict__ signal,
                                                  const long long stride, const int n_w,
                                                  const double* __restrict__ weight, double* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= P) return;
    const double* s = signal + (size_t)warp * stride;
    double acc[SNR_MAXW] = {0.0, 0.0, 0.0, 0.0};
    for (int k = lane; k < n_y; k += 32) {
        const double v = s[k];
        const double v2 = v * v;
#pragma unroll
        for (int w = 0; w < SNR_MAXW; ++w)
            if (w < n_w) acc[w] = fma(weight[(size_t)w * n_y + k], v2, acc[w]);
    }
#pragma unroll
    for (int w = 0; w < SNR_MAXW; ++w) {
        double a = acc[w];
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
        if (lane == 0 && w < n_w) out[(size_t)warp * n_w + w] = sqrt(a);
    }
}

}  // namespace ptt

using namespace ptt;

extern "C" int ptt_suppression_batch(int n, const double* v_wall, const double* alpha_n, int n_tri, const double* tri,
                                     const double* val, int mode, double alpha_min, double* out, void* stream_) {
    if (n == 0) return PTT_OK;
    if (n < 0 || n_tri < 0 || !v_wall || !alpha_n || !out || (n_tri > 0 && (!tri || !val)) || (mode != 0 && mode != 1))
        return set_error(PTT_E_ARG, "ptt_suppression_batch: null pointer or bad size");
    suppression_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(n, v_wall, alpha_n, n_tri, tri, val, mode,
                                                                            alpha_min, out);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("suppression_kernel", e);
    return PTT_OK;
}

extern "C" int ptt_snr_batch(int P, int n_y, const double* signal, int64_t stride, int n_w, const double* weight,
                             double* out, void* stream_) {
    if (P == 0) return PTT_OK;
    if (P < 0 || n_y < 1 || !signal || !weight || !out || n_w < 1 || n_w > SNR_MAXW || stride < n_y)
        return set_error(PTT_E_ARG, "ptt_snr_batch: null pointer or bad size (1 <= n_w <= 4)");
    const long long threads = (long long)P * 32;
    snr_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(P, n_y, signal, stride, n_w, weight, out);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("snr_kernel", e);
    return PTT_OK;
}
