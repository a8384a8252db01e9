// K7': spec_den_gw_scaled for large batches, "cell" formulation.        pttools/ssmtools/spectrum.py:327-368
//
// The reference integrates  G(zeta) P_v(y zeta) P_v(y zeta')  over n_int samples zeta_k per wavenumber y, with P_v
// linearly interpolated from z_lookup (np.interp).  On a maximal run of samples k over which BOTH lookups stay inside
// the same segments of z_lookup (a "cell"), with knot values (f0, f1) of the first and (g0, g1) of the second segment,
//     P_v(y zeta_k)  = f0 (1 - u_k)  + f1 u_k ,       u_k  = (y zeta_k  - x0)  / (x1 - x0),
//     P_v(y zeta'_k) = g0 (1 - u'_k) + g1 u'_k,
// so the cell's contribution to the trapezoid sum is
//     f0 (g0 C00 + g1 C01) + f1 (g0 C10 + g1 C11),   C00 = sum G_k (1-u_k)(1-u'_k), C01 = sum G_k (1-u_k) u'_k, ...
// The cell boundaries and C00..C11 depend on (y, cs, z_lookup, n_int) only - not on the profile - so they are built
// once per plan (gw_cells_build_kernel); the sum is the same trapezoid sum over the same samples, regrouped.  With the
// default sizes a wavenumber has ~2500 cells instead of 10^4 samples, and a cell costs 6 FMAs and ONE 8-byte load per
// profile (a step to the next segment brings in one new knot) instead of ~11 FP64 instructions and two
// shared-memory lookups per sample.
//
// Mapping: lanes <-> profiles (GC_PPL per lane), warp <-> one wavenumber, CTA = GC_WARPS adjacent wavenumbers of the
// same 32 * GC_PPL profiles.  P_v is stored knot-major, F[j][p] (j = 0 and j = n_zl + 1 duplicate the end knots:
// np.interp's clamps), so that a warp reads one knot of 32 profiles as one coalesced 256-byte request.
#include "ptt_host.cuh"

namespace ptt {

#ifndef PTT_GC_WARPS
#define PTT_GC_WARPS 4
#endif
#ifndef PTT_GC_PPL
#define PTT_GC_PPL 4
#endif
#ifndef PTT_GC_UNROLL
#define PTT_GC_UNROLL 4
#endif
constexpr int GC_UNROLL = PTT_GC_UNROLL;  // events in flight per warp
constexpr int GC_WARPS = PTT_GC_WARPS;    // adjacent wavenumbers per CTA
constexpr int GC_PPL = PTT_GC_PPL;        // profiles per lane
constexpr int GC_PPW = 32 * GC_PPL;       // profiles per warp

// One thread per wavenumber: walk the samples in order and emit one event per unit step of either lookup segment.
//   segment e of a lookup = knots (F[e], F[e+1]) of the extended table; e = 0 and e = n_zl are the clamps (F[0] = F[1],
//   F[n_zl] = F[n_zl+1]), e = 1 .. n_zl-1 the segments [z_lookup[e-1], z_lookup[e]].
//   event = (row << 1) | which:  which = 0: first lookup steps UP   (f0 <- f1, f1 <- F[row]),
//                                which = 1: second lookup steps DOWN (g1 <- g0, g0 <- F[row]).
//   coef[c] = C00, C01, C10, C11 of the samples from event c up to the next event.
// The stream starts with the two steps per lookup that load its initial pair of knots; a lookup that moves by more
// than one segment between two samples is walked in unit steps (zero coefficients in between).  The first lookup
// only ever moves up and the second only down (zeta ascends, zeta' = zeta_+ + zeta_- - zeta descends); anything else
// flags the wavenumber (ncell = -2).
__global__ void gw_cells_build_kernel(const int n_zl, const int n_y, const int n_int,
                                      const double* __restrict__ z_lookup, const double* __restrict__ y,
                                      const double* __restrict__ yterm, const double* __restrict__ L1,
                                      const double* __restrict__ L2, const double* __restrict__ Z1,
                                      const double* __restrict__ Z2, const double* __restrict__ G, const int cap,
                                      int* __restrict__ ncell, int* __restrict__ meta, double* __restrict__ coef) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_y) return;
    const double yi = y[i], ub = yterm[i] + 1.0;
    int* m = meta + (long long)i * cap;
    double* cf = coef + (long long)i * cap * 4;
    int c = -1, e1p = 0, e2p = 0;
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
    bool overflow = false, bad = false;
    auto emit = [&](const int row, const int which) {
        if (c >= 0 && !overflow) { cf[4 * c + 0] = c00; cf[4 * c + 1] = c01; cf[4 * c + 2] = c10; cf[4 * c + 3] = c11; }
        c00 = c01 = c10 = c11 = 0.0;
        ++c;
        if (c >= cap) overflow = true;
        else m[c] = (row << 1) | which;
    };
    double x0a = 0.0, ia = 0.0, x0b = 0.0, ib = 0.0;      // origin and 1/width of the current segments (0: clamp)
    auto seg = [&](const int e, double& x0, double& inv) {
        if (e <= 0 || e >= n_zl) { x0 = 0.0; inv = 0.0; }
        else { x0 = z_lookup[e - 1]; inv = 1.0 / (z_lookup[e] - x0); }
    };
    for (int k = 0; k < n_int; ++k) {
        int e1 = __double2int_rd(ub + L1[k]);
        int e2 = __double2int_rd(ub + L2[k]);
        e1 = max(0, min(n_zl, e1));
        e2 = max(0, min(n_zl, e2));
        if (k == 0) {
            emit(e1, 0);            // f1 <- F[e1]
            emit(e1 + 1, 0);        // f0 <- F[e1], f1 <- F[e1 + 1]
            emit(e2 + 1, 1);        // g0 <- F[e2 + 1]
            emit(e2, 1);            // g1 <- F[e2 + 1], g0 <- F[e2]
            e1p = e1; e2p = e2;
        } else {
            if (e1 < e1p || e2 > e2p) { bad = true; e1p = e1; e2p = e2; }
            while (e1p < e1) { ++e1p; emit(e1p + 1, 0); }
            while (e2p > e2) { --e2p; emit(e2p, 1); }
        }
        seg(e1, x0a, ia);
        seg(e2, x0b, ib);
        const double g = G[k];
        const double u = (yi * Z1[k] - x0a) * ia, up = (yi * Z2[k] - x0b) * ib;      // 0 on the clamps
        const double gu = g * u, g1u = g - gu;                                        // G u, G (1 - u)
        c11 = fma(gu, up, c11);
        c10 += gu - gu * up;
        c01 = fma(g1u, up, c01);
        c00 += g1u - g1u * up;
    }
    if (c >= 0 && !overflow) { cf[4 * c + 0] = c00; cf[4 * c + 1] = c01; cf[4 * c + 2] = c10; cf[4 * c + 3] = c11; }
    ncell[i] = overflow ? -1 : (bad ? -2 : c + 1);
}

// Knot-major copy of P_v with duplicated end knots: F[j][p] = pv[p][clamp(j - 1, 0, n_zl - 1)], j = 0 .. n_zl + 1.
// 32 knots x 32 profiles per CTA, transposed through shared memory so that both sides are coalesced.
__global__ void __launch_bounds__(1024) gw_table_kernel(const int n_zl, const int P, const int Ppad,
                                                        const double* __restrict__ pv, const long long pv_stride,
                                                        double* __restrict__ F) {
    __shared__ double s_f[32][33];
    const int j0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    {
        const int p = p0 + ty, j = j0 + tx;
        const int k = max(0, min(n_zl - 1, j - 1));
        s_f[ty][tx] = (p < P && j <= n_zl + 1) ? pv[(long long)p * pv_stride + k] : 0.0;
    }
    __syncthreads();
    const int j = j0 + ty, p = p0 + tx;
    if (j > n_zl + 1 || p >= Ppad) return;
    F[(long long)j * Ppad + p] = s_f[tx][ty];
}

__global__ void __launch_bounds__(GC_WARPS * 32) spec_den_gw_cells_kernel(
    const int n_y, const int P, const int Ppad, const double* __restrict__ y, const int cap,
    const int* __restrict__ ncell, const int* __restrict__ meta, const double* __restrict__ coef,
    const double* __restrict__ F, const double prefactor, const double* __restrict__ slf,
    const double* __restrict__ scale, double* __restrict__ sd_gw, double* __restrict__ pow_gw,
    double* __restrict__ omgw0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.y * GC_WARPS + warp;
    if (i >= n_y) return;
    const int p0 = blockIdx.x * GC_PPW + lane;
    const int nc = ncell[i];
    const int* __restrict__ m = meta + (long long)i * cap;
    const double2* __restrict__ cf = reinterpret_cast<const double2*>(coef + (long long)i * cap * 4);
    const double* __restrict__ Fp = F + p0;
    double f0[GC_PPL], f1[GC_PPL], g0[GC_PPL], g1[GC_PPL], acc[GC_PPL];
#pragma unroll
    for (int h = 0; h < GC_PPL; ++h) f0[h] = f1[h] = g0[h] = g1[h] = acc[h] = 0.0;
    // points without a shell carry NaN spectra (sweeps put them at the end of every v_wall row): a warp whose
    // profiles are all NaN at the first knot has nothing to integrate
    bool live = false;
#pragma unroll
    for (int h = 0; h < GC_PPL; ++h) live = live || (p0 + 32 * h < P && Fp[32 * h] == Fp[32 * h]);
    const int n_ev = __any_sync(0xffffffffu, live) ? nc : 0;
    if (n_ev == 0 && nc > 0) {
#pragma unroll
        for (int h = 0; h < GC_PPL; ++h) acc[h] = NAN;
    }
#pragma unroll GC_UNROLL
    for (int c = 0; c < n_ev; ++c) {
        const int mm = __ldg(m + c);
        const double2 c0 = __ldg(cf + 2 * c), c1 = __ldg(cf + 2 * c + 1);     // {C00, C01}, {C10, C11}
        const double* row = Fp + (long long)(mm >> 1) * Ppad;
        double nw[GC_PPL];
#pragma unroll
        for (int h = 0; h < GC_PPL; ++h) nw[h] = __ldg(row + 32 * h);
        if (mm & 1) {                                     // warp-uniform
#pragma unroll
            for (int h = 0; h < GC_PPL; ++h) { g1[h] = g0[h]; g0[h] = nw[h]; }
        } else {
#pragma unroll
            for (int h = 0; h < GC_PPL; ++h) { f0[h] = f1[h]; f1[h] = nw[h]; }
        }
#pragma unroll
        for (int h = 0; h < GC_PPL; ++h) {
            const double t = fma(g1[h], c0.y, g0[h] * c0.x), u = fma(g1[h], c1.y, g0[h] * c1.x);
            acc[h] = fma(f0[h], t, acc[h]);
            acc[h] = fma(f1[h], u, acc[h]);
        }
    }
    const double yi = y[i];
#pragma unroll
    for (int h = 0; h < GC_PPL; ++h) {
        const int p = p0 + 32 * h;
        if (p >= P) continue;
        const double a = (nc < 0) ? NAN : acc[h];
        // p_gw = pref/y * trapz(...) with z = y zeta  =>  pref * y^2 * sum_k G_k P(y zeta_k) P(y zeta'_k)
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
    // every event moves one of the two lookups by one segment; each lookup visits at most lrange + 2 segments
    // (never more than n_zl + 1), plus the two initial events per lookup
    long long per = (long long)ceil(fabs(lrange)) + 3;
    if (per > (long long)n_zl + 1) per = (long long)n_zl + 1;
    long long cap = 2 * per + 4;
    cap = (cap + 3) / 4 * 4;
    (void)n_int;
    return cap > 0x7fffffff ? -1 : (int)cap;
}

extern "C" int ptt_gw_cells_build(const ptt_gw_plan_t* plan, int cap, int32_t* ncell, int32_t* meta, double* coef,
                                  void* stream_) {
    if (!plan || cap <= 0 || !ncell || !meta || !coef || plan->n_y <= 0 || plan->n_int <= 0 || plan->n_zl < 2)
        return set_error(PTT_E_ARG, "ptt_gw_cells_build: null pointer or bad size");
    cudaStream_t stream = (cudaStream_t)stream_;
    gw_cells_build_kernel<<<(plan->n_y + 63) / 64, 64, 0, stream>>>(plan->n_zl, plan->n_y, plan->n_int, plan->z_lookup,
                                                                    plan->y, plan->yterm, plan->L1, plan->L2, plan->Z1,
                                                                    plan->Z2, plan->G, cap, ncell, meta, coef);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("gw_cells_build_kernel", e);
    return PTT_OK;
}

extern "C" int64_t ptt_gw_cells_table_size(int n_zl, int P) {
    const int64_t Ppad = ((int64_t)P + GC_PPW - 1) / GC_PPW * GC_PPW;
    return ((int64_t)n_zl + 2) * Ppad;
}

extern "C" int ptt_spec_den_gw_cells_batch(const ptt_gw_plan_t* plan, const ptt_gw_cells_t* cells, int P,
                                           const double* pv, int64_t pv_stride, const double* slf, const double* scale,
                                           double* table, double* sd_gw, double* pow_gw, double* omgw0, void* stream_) {
    if (P == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (!plan || !cells || P < 0 || !pv || !slf || !table || !pow_gw || !cells->ncell || !cells->meta || !cells->coef ||
        cells->cap <= 0 || pv_stride < plan->n_zl)
        return set_error(PTT_E_ARG, "ptt_spec_den_gw_cells_batch: null pointer or bad size");
    if (P == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int Ppad = (P + GC_PPW - 1) / GC_PPW * GC_PPW;
    {
        dim3 grid((plan->n_zl + 2 + 31) / 32, Ppad / 32), block(32, 32);
        gw_table_kernel<<<grid, block, 0, stream>>>(plan->n_zl, P, Ppad, pv, pv_stride, table);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_cuda_error("gw_table_kernel", e);
    }
    dim3 grid(Ppad / GC_PPW, (plan->n_y + GC_WARPS - 1) / GC_WARPS);
    spec_den_gw_cells_kernel<<<grid, GC_WARPS * 32, 0, stream>>>(plan->n_y, P, Ppad, plan->y, cells->cap, cells->ncell,
                                                                 cells->meta, cells->coef, table, plan->prefactor, slf,
                                                                 scale, sd_gw, pow_gw, omgw0);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("spec_den_gw_cells_kernel", e);
    return PTT_OK;
}
