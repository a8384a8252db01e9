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
// once per plan (gw_cells_build_kernel); the sum is the same trapezoid sum over the same samples, regrouped.  The cells
// are collected per segment of the first lookup, where (f0, f1) are fixed and the second lookup walks down a few
// consecutive knots g_j:   f0 * sum_j g_j D0_j + f1 * sum_j g_j D1_j.   With the default sizes a wavenumber has ~1250
// such segments and ~1250 steps of the second lookup instead of 10^4 samples; per profile a segment costs one 8-byte
// load and 2 FMAs, a step one 8-byte load and 2 FMAs, plus 4 FMAs per segment - against ~11 FP64 instructions and two
// shared-memory lookups per sample.
//
// Mapping: lanes <-> profiles (GC_PPL per lane), warp <-> one wavenumber, CTA = GC_WARPS adjacent wavenumbers of the
// same 32 * GC_PPL profiles.  P_v is stored knot-major, F[j][p] (j = 0 and j = n_zl + 1 duplicate the end knots:
// np.interp's clamps), so that a warp reads one knot of 32 profiles as one coalesced 256-byte request, and both lookups
// walk consecutive rows (the first upwards, the second downwards).
#include "ptt_host.cuh"

namespace ptt {

#ifndef PTT_GC_WARPS
#define PTT_GC_WARPS 4
#endif
#ifndef PTT_GC_PPL
#define PTT_GC_PPL 6
#endif
constexpr int GC_WARPS = PTT_GC_WARPS;    // adjacent wavenumbers per CTA
constexpr int GC_PPL = PTT_GC_PPL;        // profiles per lane
constexpr int GC_PPW = 32 * GC_PPL;       // profiles per warp

// One thread per wavenumber: walk the samples in order and group them by segment of the FIRST lookup ("A-segments":
// knots f0 = F[e1], f1 = F[e1+1]).  While the first lookup sits in one segment the second one visits m + 1 consecutive
// segments downwards, i.e. m + 2 knots g_0 > g_1 > ... (g_0 = F[e2_start + 1]); collecting the cell sums per knot,
//     contribution of the A-segment = f0 * sum_j g_j D0_j + f1 * sum_j g_j D1_j,
//     D0_j = C01(cell j) + C00(cell j-1),   D1_j = C11(cell j) + C10(cell j-1)        (cell j: second lookup at e2_start - j).
// Per wavenumber: meta[0] = e1 of the first A-segment, meta[1] = e2 at the start, meta[2 + a] = m of A-segment a;
// coef = the (D0_j, D1_j) pairs of all A-segments back to back; ncell = number of A-segments (-1: a capacity exceeded,
// -2: a lookup moved the wrong way).  Segment e of a lookup = knots (F[e], F[e+1]) of the extended table; e = 0 and
// e = n_zl are np.interp's clamps, e = 1 .. n_zl-1 the segments [z_lookup[e-1], z_lookup[e]].  A lookup that moves by
// more than one segment between two samples is walked in unit steps (empty cells / empty A-segments).
constexpr int GC_MAXK = 40;     // knots of the second lookup per A-segment (m + 2) the build kernel can hold

__global__ void gw_cells_build_kernel(const int n_zl, const int n_y, const int n_int,
                                      const double* __restrict__ z_lookup, const double* __restrict__ y,
                                      const double* __restrict__ yterm, const double* __restrict__ L1,
                                      const double* __restrict__ L2, const double* __restrict__ Z1,
                                      const double* __restrict__ Z2, const double* __restrict__ G, const int cap,
                                      const int cap_c, int* __restrict__ ncell, int* __restrict__ meta,
                                      double* __restrict__ coef) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_y) return;
    const double yi = y[i], ub = yterm[i] + 1.0;
    int* mt = meta + (long long)i * (cap + 2);
    double* cf = coef + (long long)i * cap_c * 2;
    double D0[GC_MAXK], D1[GC_MAXK];
    for (int j = 0; j < GC_MAXK; ++j) { D0[j] = 0.0; D1[j] = 0.0; }
    int a = 0, m = 0, co = 0;               // A-segment index, steps of the second lookup inside it, coef pairs written
    int e1p = 0, e2p = 0;
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
    bool overflow = false, bad = false;
    auto flush_cell = [&]() {              // current cell = second lookup m steps below its start: knots j = m, m + 1
        D0[m] += c01; D0[m + 1] += c00;
        D1[m] += c11; D1[m + 1] += c10;
        c00 = c01 = c10 = c11 = 0.0;
    };
    auto close_segment = [&]() {
        if (a >= cap || co + m + 2 > cap_c) overflow = true;
        if (!overflow) {
            mt[2 + a] = m;
            for (int j = 0; j < m + 2; ++j) { cf[2 * (co + j)] = D0[j]; cf[2 * (co + j) + 1] = D1[j]; }
        }
        co += m + 2;
        for (int j = 0; j < m + 2; ++j) { D0[j] = 0.0; D1[j] = 0.0; }
        ++a;
        m = 0;
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
            mt[0] = e1; mt[1] = e2;
            e1p = e1; e2p = e2;
        } else {
            if (e1 < e1p || e2 > e2p) { bad = true; e1p = e1; e2p = e2; }
            while (e1p < e1) { flush_cell(); close_segment(); ++e1p; }
            while (e2p > e2) {
                flush_cell();
                if (m + 3 > GC_MAXK) { overflow = true; m = 0; }
                ++m;
                --e2p;
            }
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
    flush_cell();
    close_segment();
    ncell[i] = overflow ? -1 : (bad ? -2 : a);
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

#ifndef PTT_GC_MINB
#define PTT_GC_MINB 4      // 128 registers: 4 CTAs per SM (measured: 3.7 ms per 2000 profiles against 4.4 - 4.8 for
                           // 4, 6 or 8 profiles per lane without the cap)
#endif
__global__ void __launch_bounds__(GC_WARPS * 32, PTT_GC_MINB) spec_den_gw_cells_kernel(
    const int n_y, const int P, const int Ppad, const double* __restrict__ y, const int cap, const int cap_c,
    const int* __restrict__ ncell, const int* __restrict__ meta, const double* __restrict__ coef,
    const double* __restrict__ F, const double prefactor, const double* __restrict__ slf,
    const double* __restrict__ scale, double* __restrict__ sd_gw, double* __restrict__ pow_gw,
    double* __restrict__ omgw0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // wavenumber blocks are the fast grid index: the CTAs resident together walk the rows of two or three profile
    // groups (tens of MB, L2-resident) instead of a window of every group (more than the L2 holds)
    const int i = blockIdx.x * GC_WARPS + warp;
    if (i >= n_y) return;
    const int p0 = blockIdx.y * GC_PPW + lane;
    const int nc = ncell[i];
    const int* __restrict__ mt = meta + (long long)i * (cap + 2);
    const double2* __restrict__ cp = reinterpret_cast<const double2*>(coef + (long long)i * cap_c * 2);
    const double* __restrict__ Fp = F + p0;
    double f0[GC_PPL], f1[GC_PPL], glo[GC_PPL], ghi[GC_PPL], acc[GC_PPL];
#pragma unroll
    for (int h = 0; h < GC_PPL; ++h) f0[h] = f1[h] = glo[h] = ghi[h] = acc[h] = 0.0;
    // points without a shell carry NaN spectra (sweeps put them at the end of every v_wall row): a warp whose
    // profiles are all NaN at the first knot has nothing to integrate
    bool live = false;
#pragma unroll
    for (int h = 0; h < GC_PPL; ++h) live = live || (p0 + 32 * h < P && Fp[32 * h] == Fp[32 * h]);
    const int n_seg = __any_sync(0xffffffffu, live) ? nc : 0;
    if (n_seg == 0 && nc > 0) {
#pragma unroll
        for (int h = 0; h < GC_PPL; ++h) acc[h] = NAN;
    }
    if (n_seg > 0) {
        const int e1 = __ldg(mt), e2 = __ldg(mt + 1);
        const double* rowA = Fp + (long long)e1 * Ppad;          // next knot of the first lookup: one row up per A-segment
        const double* rowB = Fp + (long long)e2 * Ppad;          // next knot of the second lookup: one row down per step
#pragma unroll
        for (int h = 0; h < GC_PPL; ++h) {
            f1[h] = __ldg(rowA + 32 * h);                        // becomes f0 of the first A-segment below
            glo[h] = __ldg(rowB + 32 * h);
            ghi[h] = __ldg(rowB + Ppad + 32 * h);
        }
        rowA += Ppad;
        rowB -= Ppad;
        for (int a = 0; a < n_seg; ++a) {
            const int m = __ldg(mt + 2 + a);
#pragma unroll
            for (int h = 0; h < GC_PPL; ++h) { f0[h] = f1[h]; f1[h] = __ldg(rowA + 32 * h); }
            rowA += Ppad;
            const double2 d0 = __ldg(cp), d1 = __ldg(cp + 1);
            double t0[GC_PPL], t1[GC_PPL];
#pragma unroll
            for (int h = 0; h < GC_PPL; ++h) {
                t0[h] = fma(glo[h], d1.x, ghi[h] * d0.x);
                t1[h] = fma(glo[h], d1.y, ghi[h] * d0.y);
            }
            for (int s = 0; s < m; ++s) {                        // warp-uniform trip count (0 .. 3 for the default grids)
                const double2 d = __ldg(cp + 2 + s);
#pragma unroll
                for (int h = 0; h < GC_PPL; ++h) {
                    const double gn = __ldg(rowB + 32 * h);
                    t0[h] = fma(gn, d.x, t0[h]);
                    t1[h] = fma(gn, d.y, t1[h]);
                    ghi[h] = glo[h];
                    glo[h] = gn;
                }
                rowB -= Ppad;
            }
            cp += m + 2;
#pragma unroll
            for (int h = 0; h < GC_PPL; ++h) {
                acc[h] = fma(f0[h], t0[h], acc[h]);
                acc[h] = fma(f1[h], t1[h], acc[h]);
            }
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
    long long cap = per + 2;                  // A-segments per wavenumber
    cap = (cap + 3) / 4 * 4;
    (void)n_int;
    return cap > 0x3fffffff ? -1 : (int)cap;
}

// coefficient pairs per wavenumber: one per knot of the second lookup per A-segment = steps + 2 per segment
extern "C" int ptt_gw_cells_coef_capacity(int cap) { return 3 * cap + 8; }

extern "C" int ptt_gw_cells_build(const ptt_gw_plan_t* plan, int cap, int32_t* ncell, int32_t* meta, double* coef,
                                  void* stream_) {
    if (!plan || cap <= 0 || !ncell || !meta || !coef || plan->n_y <= 0 || plan->n_int <= 0 || plan->n_zl < 2)
        return set_error(PTT_E_ARG, "ptt_gw_cells_build: null pointer or bad size");
    cudaStream_t stream = (cudaStream_t)stream_;
    gw_cells_build_kernel<<<(plan->n_y + 63) / 64, 64, 0, stream>>>(plan->n_zl, plan->n_y, plan->n_int, plan->z_lookup,
                                                                    plan->y, plan->yterm, plan->L1, plan->L2, plan->Z1,
                                                                    plan->Z2, plan->G, cap,
                                                                    ptt_gw_cells_coef_capacity(cap), ncell, meta, coef);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("gw_cells_build_kernel", e);
    return PTT_OK;
}

// Row stride of the knot-major table: the profiles rounded up to whole warps' worth, plus 32 when that would make
// the stride a multiple of 2 KB (consecutive knots of one profile group then never share L2 sets; measured neutral).
static inline int64_t gw_row_stride(const int64_t P) {
    int64_t s = (P + GC_PPW - 1) / GC_PPW * GC_PPW;
    if (s % 256 == 0) s += 32;
    return s;
}

extern "C" int64_t ptt_gw_cells_table_size(int n_zl, int P) { return ((int64_t)n_zl + 2) * gw_row_stride(P); }

extern "C" int ptt_spec_den_gw_cells_batch(const ptt_gw_plan_t* plan, const ptt_gw_cells_t* cells, int P,
                                           const double* pv, int64_t pv_stride, const double* slf, const double* scale,
                                           double* table, double* sd_gw, double* pow_gw, double* omgw0, void* stream_) {
    if (P == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (!plan || !cells || P < 0 || !pv || !slf || !table || !pow_gw || !cells->ncell || !cells->meta || !cells->coef ||
        cells->cap <= 0 || pv_stride < plan->n_zl)
        return set_error(PTT_E_ARG, "ptt_spec_den_gw_cells_batch: null pointer or bad size");
    if (P == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int Ppad = (int)gw_row_stride(P);
    const int n_groups = (P + GC_PPW - 1) / GC_PPW;
    {
        dim3 grid((plan->n_zl + 2 + 31) / 32, (Ppad + 31) / 32), block(32, 32);
        gw_table_kernel<<<grid, block, 0, stream>>>(plan->n_zl, P, Ppad, pv, pv_stride, table);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_cuda_error("gw_table_kernel", e);
    }
    dim3 grid((plan->n_y + GC_WARPS - 1) / GC_WARPS, n_groups);
    spec_den_gw_cells_kernel<<<grid, GC_WARPS * 32, 0, stream>>>(plan->n_y, P, Ppad, plan->y, cells->cap,
                                                                 ptt_gw_cells_coef_capacity(cells->cap), cells->ncell,
                                                                 cells->meta, cells->coef, table, plan->prefactor, slf,
                                                                 scale, sd_gw, pow_gw, omgw0);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("spec_den_gw_cells_kernel", e);
    return PTT_OK;
}
