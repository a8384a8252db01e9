// K7': spec_den_gw_scaled for large batches, "cell" formulation.        pttools/ssmtools/spectrum.py:327-368
//
// The reference integrates  G(zeta) P_v(y zeta) P_v(y zeta')  over n_int samples zeta_k per wavenumber y, with P_v
// linearly interpolated from z_lookup (np.interp).  On a maximal run of samples k over which BOTH lookups stay inside
// the same segments of z_lookup (a "cell"), P_v(y zeta_k) = a1 + b1 y zeta_k and P_v(y zeta'_k) = a2 + b2 y zeta'_k
// with constant (a, b), so the cell's contribution to the trapezoid sum is
//     a1 a2 S0 + b1 a2 S1 + a1 b2 S2 + b1 b2 S3,   S0 = sum G_k, S1 = sum G_k y zeta_k, S2 = sum G_k y zeta'_k,
//                                                   S3 = sum G_k y^2 zeta_k zeta'_k.
// The cell boundaries and S0..S3 depend on (y, cs, z_lookup, n_int) only - not on the profile - so they are built once
// per plan (gw_cells_build_kernel); the sum is the same trapezoid sum over the same samples, regrouped.  With the
// default sizes a wavenumber has ~2500 cells instead of 10^4 samples, and a cell costs 6 FMAs and one 16-byte load
// per profile instead of ~11 FP64 instructions and two shared-memory lookups per sample.
//
// Mapping: lanes <-> profiles (GC_PPL per lane), warp <-> one wavenumber, CTA = GC_WARPS adjacent wavenumbers of the
// same 32 * GC_PPL profiles.  The (a, b) tables are stored entry-major, T[e][p], so that a warp reads one entry of 32 profiles as one
// coalesced 512-byte request; the 8 warps of a CTA walk nearly the same entries, which keeps them in L1.
#include "ptt_host.cuh"

namespace ptt {

#ifndef PTT_GC_WARPS
#define PTT_GC_WARPS 8
#endif
#ifndef PTT_GC_PPL
#define PTT_GC_PPL 2
#endif
#ifndef PTT_GC_PF
#define PTT_GC_PF 0
#endif
constexpr int GC_WARPS = PTT_GC_WARPS;    // adjacent wavenumbers per CTA
constexpr int GC_PPL = PTT_GC_PPL;        // profiles per lane
constexpr int GC_PPW = 32 * GC_PPL;       // profiles per warp
constexpr int GC_PF = PTT_GC_PF;          // L1 prefetch distance in events; 0 = off (measured: 8..32 are 15-25 % slower,
                                          // the kernel is bound by L1/L2 throughput, not by latency)

// One thread per wavenumber: walk the samples in order and emit one event per change of either lookup segment.
// meta = (entry << 1) | which  (which = 0: lookup of y*zeta, 1: lookup of y*zeta'), coef = S0..S3 of the samples from
// this event up to the next one.  Entries are indices into the (a, b) table: 0 and n_zl are the np.interp end clamps.
__global__ void gw_cells_build_kernel(const int n_zl, const int n_y, const int n_int, const double* __restrict__ y,
                                      const double* __restrict__ yterm, const double* __restrict__ L1,
                                      const double* __restrict__ L2, const double* __restrict__ Z1,
                                      const double* __restrict__ Z2, const double* __restrict__ G, const int cap,
                                      int* __restrict__ ncell, int* __restrict__ meta, double* __restrict__ coef) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_y) return;
    const double yi = y[i], ub = yterm[i] + 1.0;
    int* m = meta + (long long)i * cap;
    double* cf = coef + (long long)i * cap * 4;
    int c = -1, e1p = -1, e2p = -1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    bool overflow = false;
    for (int k = 0; k < n_int; ++k) {
        int e1 = __double2int_rd(ub + L1[k]);
        int e2 = __double2int_rd(ub + L2[k]);
        e1 = max(0, min(n_zl, e1));
        e2 = max(0, min(n_zl, e2));
        for (int which = 0; which < 2; ++which) {
            const int e = which ? e2 : e1;
            if (e == (which ? e2p : e1p)) continue;
            if (c >= 0 && !overflow) {
                cf[4 * c + 0] = s0; cf[4 * c + 1] = s1; cf[4 * c + 2] = s2; cf[4 * c + 3] = s3;
            }
            s0 = s1 = s2 = s3 = 0.0;
            ++c;
            if (c >= cap) overflow = true;
            else m[c] = (e << 1) | which;
            if (which) e2p = e; else e1p = e;
        }
        const double g = G[k], a = yi * Z1[k], b = yi * Z2[k];
        s0 += g;
        s1 = fma(g, a, s1);
        s2 = fma(g, b, s2);
        s3 = fma(g * a, b, s3);
    }
    if (c >= 0 && !overflow) {
        cf[4 * c + 0] = s0; cf[4 * c + 1] = s1; cf[4 * c + 2] = s2; cf[4 * c + 3] = s3;
    }
    ncell[i] = overflow ? -1 : c + 1;
}

// (a, b) tables, entry-major: T[e][p] = {f_k - b x_k, b} on segment k = e-1 of z_lookup; e = 0 and e = n_zl clamp.
// 32 entries x 32 profiles per CTA, transposed through shared memory so that both sides are coalesced.
__global__ void __launch_bounds__(1024) gw_table_kernel(const int n_zl, const int P, const int Ppad,
                                                        const double* __restrict__ z_lookup,
                                                        const double* __restrict__ pv, const long long pv_stride,
                                                        double2* __restrict__ T) {
    __shared__ double s_f[32][33];
    const int e0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    // s_f[pp][j] = f[e0 - 1 + j], j = 0..32
    {
        const int p = p0 + ty;
        for (int j = tx; j < 33; j += 32) {
            const int k = e0 - 1 + j;
            s_f[ty][j] = (p < P && k >= 0 && k < n_zl) ? pv[(long long)p * pv_stride + k] : 0.0;
        }
    }
    __syncthreads();
    const int e = e0 + ty, p = p0 + tx;
    if (e > n_zl || p >= Ppad) return;
    double2 ab = make_double2(0.0, 0.0);
    if (p < P) {
        if (e == 0) ab.x = s_f[tx][1];
        else if (e == n_zl) ab.x = s_f[tx][ty];
        else {
            const double x0 = z_lookup[e - 1], x1 = z_lookup[e], f0 = s_f[tx][ty], f1 = s_f[tx][ty + 1];
            const double b = (f1 - f0) / (x1 - x0);
            ab = make_double2(f0 - b * x0, b);
        }
    }
    T[(long long)e * Ppad + p] = ab;
}

__global__ void __launch_bounds__(GC_WARPS * 32) spec_den_gw_cells_kernel(
    const int n_y, const int P, const int Ppad, const double* __restrict__ y, const int cap,
    const int* __restrict__ ncell, const int* __restrict__ meta, const double* __restrict__ coef,
    const double2* __restrict__ T, const double prefactor, const double* __restrict__ slf,
    const double* __restrict__ scale, double* __restrict__ sd_gw, double* __restrict__ pow_gw,
    double* __restrict__ omgw0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.y * GC_WARPS + warp;
    if (i >= n_y) return;
    const int p0 = blockIdx.x * GC_PPW + lane;
    const int nc = ncell[i];
    const int* __restrict__ m = meta + (long long)i * cap;
    const double2* __restrict__ cf = reinterpret_cast<const double2*>(coef + (long long)i * cap * 4);
    const double2* __restrict__ Tp = T + p0;
    double2 A[GC_PPL], B[GC_PPL];
    double acc[GC_PPL];
#pragma unroll
    for (int h = 0; h < GC_PPL; ++h) {
        A[h] = make_double2(0.0, 0.0);
        B[h] = A[h];
        acc[h] = 0.0;
    }
    // tuning knob: request the rows of the events GC_PF ahead
    if (GC_PF > 0) {
        for (int c = 0; c < min(GC_PF, nc); ++c) {
            const double2* row = Tp + (long long)(__ldg(m + c) >> 1) * Ppad;
#pragma unroll
            for (int h = 0; h < GC_PPL; ++h) asm volatile("prefetch.global.L1 [%0];" ::"l"(row + 32 * h));
        }
    }
#pragma unroll 4
    for (int c = 0; c < nc; ++c) {
        if (GC_PF > 0 && c + GC_PF < nc) {
            const double2* prow = Tp + (long long)(__ldg(m + c + GC_PF) >> 1) * Ppad;
#pragma unroll
            for (int h = 0; h < GC_PPL; ++h) asm volatile("prefetch.global.L1 [%0];" ::"l"(prow + 32 * h));
        }
        const int mm = __ldg(m + c);
        const double2 s01 = __ldg(cf + 2 * c), s23 = __ldg(cf + 2 * c + 1);
        const double2* row = Tp + (long long)(mm >> 1) * Ppad;
        double2 nw[GC_PPL];
#pragma unroll
        for (int h = 0; h < GC_PPL; ++h) nw[h] = __ldg(row + 32 * h);
        if (mm & 1) {                                     // warp-uniform
#pragma unroll
            for (int h = 0; h < GC_PPL; ++h) B[h] = nw[h];
        } else {
#pragma unroll
            for (int h = 0; h < GC_PPL; ++h) A[h] = nw[h];
        }
        // a1 (a2 S0 + b2 S2) + b1 (a2 S1 + b2 S3)
#pragma unroll
        for (int h = 0; h < GC_PPL; ++h) {
            const double t = fma(B[h].y, s23.x, B[h].x * s01.x), u = fma(B[h].y, s23.y, B[h].x * s01.y);
            acc[h] = fma(A[h].x, t, acc[h]);
            acc[h] = fma(A[h].y, u, acc[h]);
        }
    }
    const double yi = y[i];
#pragma unroll
    for (int h = 0; h < GC_PPL; ++h) {
        const int p = p0 + 32 * h;
        if (p >= P) continue;
        const double a = (nc < 0) ? NAN : acc[h];
        const double sd = prefactor * yi * yi * a * slf[p];
        const long long o = (long long)p * n_y + i;
        if (sd_gw) sd_gw[o] = sd;
        const double pw = yi * yi * yi / (2.0 * M_PI * M_PI) * sd;
        if (pow_gw) pow_gw[o] = pw;
        if (omgw0 && scale) omgw0[o] = scale[p] * pw;
    }
}

}  // namespace ptt

using namespace ptt;

extern "C" int ptt_gw_cells_capacity(int n_zl, int n_int, double lrange) {
    // every event moves one of the two lookups to a new entry; each lookup visits at most lrange + 2 entries
    // (and never more than n_int or n_zl + 1)
    long long per = (long long)ceil(fabs(lrange)) + 3;
    if (per > n_int) per = n_int;
    if (per > (long long)n_zl + 1) per = (long long)n_zl + 1;
    long long cap = 2 * per;
    cap = (cap + 3) / 4 * 4;
    return cap > 0x7fffffff ? -1 : (int)cap;
}

extern "C" int ptt_gw_cells_build(const ptt_gw_plan_t* plan, int cap, int32_t* ncell, int32_t* meta, double* coef,
                                  void* stream_) {
    if (!plan || cap <= 0 || !ncell || !meta || !coef || plan->n_y <= 0 || plan->n_int <= 0 || plan->n_zl < 2)
        return set_error(PTT_E_ARG, "ptt_gw_cells_build: null pointer or bad size");
    cudaStream_t stream = (cudaStream_t)stream_;
    gw_cells_build_kernel<<<(plan->n_y + 63) / 64, 64, 0, stream>>>(plan->n_zl, plan->n_y, plan->n_int, plan->y,
                                                                    plan->yterm, plan->L1, plan->L2, plan->Z1, plan->Z2,
                                                                    plan->G, cap, ncell, meta, coef);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("gw_cells_build_kernel", e);
    return PTT_OK;
}

extern "C" int64_t ptt_gw_cells_table_size(int n_zl, int P) {
    const int64_t Ppad = ((int64_t)P + GC_PPW - 1) / GC_PPW * GC_PPW;
    return ((int64_t)n_zl + 1) * Ppad * 2;
}

extern "C" int ptt_spec_den_gw_cells_batch(const ptt_gw_plan_t* plan, const ptt_gw_cells_t* cells, int P,
                                           const double* pv, int64_t pv_stride, const double* slf, const double* scale,
                                           double* table, double* sd_gw, double* pow_gw, double* omgw0, void* stream_) {
    if (!plan || !cells || P < 0 || !pv || !slf || !table || !pow_gw || !cells->ncell || !cells->meta || !cells->coef ||
        cells->cap <= 0 || pv_stride < plan->n_zl)
        return set_error(PTT_E_ARG, "ptt_spec_den_gw_cells_batch: null pointer or bad size");
    if (P == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int Ppad = (P + GC_PPW - 1) / GC_PPW * GC_PPW;
    {
        dim3 grid((plan->n_zl + 1 + 31) / 32, Ppad / 32), block(32, 32);
        gw_table_kernel<<<grid, block, 0, stream>>>(plan->n_zl, P, Ppad, plan->z_lookup, pv, pv_stride,
                                                    reinterpret_cast<double2*>(table));
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_cuda_error("gw_table_kernel", e);
    }
    dim3 grid(Ppad / GC_PPW, (plan->n_y + GC_WARPS - 1) / GC_WARPS);
    spec_den_gw_cells_kernel<<<grid, GC_WARPS * 32, 0, stream>>>(plan->n_y, P, Ppad, plan->y, cells->cap, cells->ncell,
                                                                 cells->meta, cells->coef,
                                                                 reinterpret_cast<const double2*>(table), plan->prefactor,
                                                                 slf, scale, sd_gw, pow_gw, omgw0);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("spec_den_gw_cells_kernel", e);
    return PTT_OK;
}
