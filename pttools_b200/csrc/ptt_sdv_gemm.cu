// K6': spec_den_v for a group of profiles that share v_wall, as a banded FP64 matrix product.
//
// _spec_den_v_core (pttools/ssmtools/spectrum.py:451-486) evaluates, for every wavenumber z_i,
//     P(z_i) = factor * sum_j W_j * interp(A2)(s_i T_j),   s_i = z_i / (b_R v_w),
// and np.interp is linear in the table values:  interp(A2)(x) = (1-t) A2[e] + t A2[e+1].  The weights depend on
// (z_i, T_j, v_w) only -- not on the profile -- so for the n_p profiles of a sweep that share v_w
//     P[i, p] = factor * sum_k Omega[i, k] * A2[p, k],      Omega[i, k] = sum_j W_j * (weight of knot k at s_i T_j).
// Omega is banded (each z_i sees the 3.3 decades of T-tilde): it is built once per group (sdv_weights_kernel, the
// same index arithmetic as the gather kernel) and multiplied with the A2 table by an FP64 tensor-core GEMM
// (mma.sync m8n8k4 f64 -- there is no tcgen05 kind for FP64).  Identical linear-interpolation weights, different
// summation order; the gather kernel (ptt_ssm.cu) remains for profiles that do not share v_w.
#include <cuda.h>              // CUtensorMap (types only: the encoder is fetched through the runtime, no -lcuda)

#include "ptt_host.cuh"

namespace ptt {

constexpr int GM_BM = 128;      // rows (wavenumbers) per CTA
constexpr int GM_BN = 128;      // columns (profiles) per CTA
constexpr int GM_BK = 16;
constexpr int GM_PITCH = GM_BK + 4;   // doubles; conflict-free fragment loads (half-warp covers all 32 banks)
constexpr int GM_THREADS = 512;      // 16 warps: 4 (m) x 4 (n), warp tile 32 x 32 -> 4 warps per scheduler, <= 128 registers

// Omega, tile-banded: tile t (GM_BM rows) stores [GM_BM][kw] weights for knots e_base[t] + kk, kk in [0, kw).
// One thread per (row, knot), pull form: the T-tilde samples whose lookup lands in the segment left or right of the
// knot are found arithmetically (LT is a linspace: ~1.4 samples per knot) and confirmed with the same floor() test
// as the gather kernel, so every sample is counted exactly once; no atomics, no zero-fill, coalesced stores.
// np.interp's end clamps put the whole weight of a sample on knot 0 (left) or knot n_q - 1 (right).
// 1 / (qT[k+1] - qT[k]) and 1 / qT[k] for the weights kernel (one division per knot instead of one per sample)
__global__ void sdv_recip_kernel(const int n_q, const double* __restrict__ qT, double* __restrict__ inv_dq,
                                 double* __restrict__ inv_q) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_q - 1) inv_dq[k] = 1.0 / (qT[k + 1] - qT[k]);
    if (k < n_q) inv_q[k] = 1.0 / qT[k];
}

constexpr int WT_KPT = 4;                      // knots per thread
constexpr int WT_SPAN = 256 * WT_KPT;         // knots per CTA (one row of Omega)

// One CTA = one row i of Omega and WT_SPAN consecutive knots; thread t handles knots e0 + t + 256 k, k < WT_KPT, so the
// per-row quantities (s_i, zterm_i + shift, the reciprocal spacing of LT) are formed once for four entries and the
// stores of a warp stay coalesced.
__global__ void __launch_bounds__(256) sdv_weights_kernel(
    const int n_q, const int n_z, const int nt, const double* __restrict__ qT, const double* __restrict__ z,
    const double* __restrict__ zterm, const double* __restrict__ T, const double* __restrict__ LT,
    const double* __restrict__ W, const double bv, const double shift, const int kw,
    const int* __restrict__ e_base, const double* __restrict__ inv_dq, const double* __restrict__ inv_q,
    const double2* __restrict__ omega_ab, double* __restrict__ omega) {
    const int i = blockIdx.y;
    const int kk0 = blockIdx.x * WT_SPAN;                            // first entry of this CTA within the band
    const double s = z[i] / bv;
    const double ub = zterm[i] + shift;
    const int e0 = e_base[i / GM_BM] + kk0;                          // knot of entry kk0
    const double lt0 = LT[0];
    const double inv_dlt = (nt > 1) ? (double)(nt - 1) / (LT[nt - 1] - lt0) : 0.0;     // LT is a linspace
    // number of kinks LT_j below x(knot) = knot - (zterm_i + shift): x moves by exactly one per knot, so the three counts
    // a knot needs (at x - 1, x, x + 1) are the counts of its neighbours: one per entry + a halo of two, through shared
    // memory
    __shared__ int s_cnt[WT_SPAN + 2];
    auto count_below = [&](const int d) {                            // knot e0 + d
        const double yy = ((double)(e0 + d) - ub - lt0) * inv_dlt;
        const int c = (yy > 0.0) ? __double2int_ru(yy) : 0;
        return min(c, nt);
    };
    if (omega_ab) {
#pragma unroll
        for (int k = 0; k < WT_KPT; ++k) s_cnt[threadIdx.x + 256 * k + 1] = count_below(threadIdx.x + 256 * k);
        if (threadIdx.x == 0) s_cnt[0] = count_below(-1);
        if (threadIdx.x == 32) s_cnt[WT_SPAN + 1] = count_below(WT_SPAN);
    }
    __syncthreads();
    double* __restrict__ row = omega + (size_t)i * kw;
#pragma unroll
    for (int k = 0; k < WT_KPT; ++k) {
        const int d = threadIdx.x + 256 * k;
        const int kk = kk0 + d, e = e0 + d;
        if (kk >= kw) break;
        double acc = 0.0;
        if (omega_ab && e > 0 && e < n_q - 1) {
            // Closed form.  As a function of x = e - (zterm_i + shift) the weight of an interior knot is
            //     omega(x) = alpha_p + beta_p * (s_i / qT[e])
            // on every piece p between two consecutive kinks LT_j - 1, LT_j, LT_j + 1 (the set of contributing samples
            // is fixed there and np.interp's weight (X - x0)/(x1 - x0) on a log grid is linear in X / qT[e]); the plan
            // carries (alpha, beta) per piece, and the piece index is the number of kinks below x.
            const int piece = s_cnt[d] + s_cnt[d + 1] + s_cnt[d + 2];
            const double2 ab = __ldg(omega_ab + piece);
            acc = fma(ab.y, s * __ldg(inv_q + e), ab.x);
        } else if (e < n_q) {
            // np.interp's end clamps put the whole weight of a sample on knot 0 (left) or knot n_q - 1 (right): these
            // two knots (and every knot of a plan without pieces) walk their samples, found arithmetically (LT is a
            // linspace) and confirmed with the same floor() test as the gather kernel
            int j = 0;
            if (e > 0) {
                const double jf = ((double)(e - 1) - ub - lt0) * inv_dlt;
                j = (jf > 1.0) ? ((jf < (double)nt) ? (int)jf - 1 : nt) : 0;
            }
            for (; j < nt; ++j) {
                const int er = __double2int_rd(ub + LT[j]);
                const double wj = W[j];
                int eff;
                double w0, w1;
                if (er < 0) { eff = 0; w0 = wj; w1 = 0.0; }                          // np.interp left clamp
                else if (er >= n_q - 1) { eff = n_q - 2; w0 = 0.0; w1 = wj; }          // right clamp: all on the last knot
                else {
                    // the arithmetic index may be off by one on a razor edge: the weights then extrapolate the
                    // neighbouring segment by ~1e-12 of its length, i.e. the same value (continuous interpolant)
                    eff = er;
                    const double t = (s * T[j] - qT[er]) * inv_dq[er];
                    w0 = wj * (1.0 - t);
                    w1 = wj * t;
                }
                if (eff > e) break;                 // eff is non-decreasing in j
                if (eff == e) acc += w0;
                else if (eff == e - 1) acc += w1;
            }
        }
        row[kk] = acc;
    }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// out[p, i] = factor * sum_kk omega[i, kk] * a2[p, e_base[tile(i)] + kk]
// Grid: one CTA per (row tile, column tile), column tiles fastest: the CTAs that are resident together cover a few row
// tiles and ALL their column tiles, so a row tile of Omega is fetched from HBM once and served to its other column
// tiles by the L2.  In the ragged last column tile a warp multiplies only the 8-profile fragments that hold profiles
// (the valid profiles of a v_wall row are trimmed to n_p, typically ~940 of 1000: 7 % of the DMMA work otherwise).
__global__ void __launch_bounds__(GM_THREADS) sdv_gemm_kernel(
    const int n_z, const int n_p, const int n_q, const int kw, const int n_tn, const int* __restrict__ e_base,
    const double* __restrict__ omega, const double* __restrict__ a2, const long long a2_stride, const double factor,
    double* __restrict__ out, const long long out_stride) {
    extern __shared__ double gm_smem[];
    double (*As)[GM_BM * GM_PITCH] = reinterpret_cast<double (*)[GM_BM * GM_PITCH]>(gm_smem);
    double (*Bs)[GM_BN * GM_PITCH] = reinterpret_cast<double (*)[GM_BN * GM_PITCH]>(gm_smem + 2 * GM_BM * GM_PITCH);
    const int tile_m = blockIdx.x / n_tn, tile_n = blockIdx.x % n_tn;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;          // 4 x 4 warps; warp tile 32 (m) x 32 (n)
    const int nb = min(4, max(0, (n_p - (tile_n * GM_BN + wn * 32) + 7) >> 3));      // 8-profile fragments with profiles
    const int eb = e_base[tile_m];
    const double* Ag = omega + (size_t)tile_m * GM_BM * kw;
    // global -> shared: thread loads 4 consecutive k of one row (A) and of one column/profile (B)
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    const int p_load = tile_n * GM_BN + lr;
    const bool p_ok = p_load < n_p;
    const double* Bg = a2 + (long long)(p_ok ? p_load : 0) * a2_stride;
    double ra[4], rb[4];
    auto load_regs = [&](const int k0) {
        const double* ap = Ag + (size_t)lr * kw + k0 + lk;
#pragma unroll
        for (int q = 0; q < 4; ++q) ra[q] = (k0 + lk + q < kw) ? ap[q] : 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = eb + k0 + lk + q;
            rb[q] = (p_ok && k >= 0 && k < n_q && k0 + lk + q < kw) ? Bg[k] : 0.0;
        }
    };
    auto store_smem = [&](const int buf) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            As[buf][lr * GM_PITCH + lk + q] = ra[q];
            Bs[buf][lr * GM_PITCH + lk + q] = rb[q];
        }
    };
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }

    const int n_k = (kw + GM_BK - 1) / GM_BK;
    load_regs(0);
    store_smem(0);
    __syncthreads();
    const int fr = lane >> 2, fk = lane & 3;
    for (int kt = 0; kt < n_k; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_k) load_regs((kt + 1) * GM_BK);
        const double* Asb = As[buf] + (wm * 32 + fr) * GM_PITCH + fk;
        const double* Bsb = Bs[buf] + (wn * 32 + fr) * GM_PITCH + fk;
#pragma unroll
        for (int k4 = 0; k4 < GM_BK; k4 += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = Asb[a * 8 * GM_PITCH + k4];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = Bsb[b * 8 * GM_PITCH + k4];
            if (nb == 4) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (b < nb) {
#pragma unroll
                        for (int a = 0; a < 4; ++a) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
                    }
                }
            }
        }
        if (kt + 1 < n_k) {
            store_smem(buf ^ 1);
        }
        __syncthreads();
    }
    // C fragment: row = lane / 4, cols = 2 * (lane % 4) + {0, 1}
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = tile_m * GM_BM + wm * 32 + a * 8 + fr;
        if (i >= n_z) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int p0 = tile_n * GM_BN + wn * 32 + b * 8 + 2 * fk;
            if (p0 < n_p) out[(long long)p0 * out_stride + i] = factor * acc[a][b][0];
            if (p0 + 1 < n_p) out[(long long)(p0 + 1) * out_stride + i] = factor * acc[a][b][1];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same product with the operands moved by the TMA.  Per 16-knot step one thread issues two tensor-map copies
// (cp.async.bulk.tensor.2d: the 128 x 16 tile of Omega and the 128 x 16 tile of A^2, 16 KB each, 128-byte swizzle)
// into a ring of GT_STAGES stages, GT_STAGES - 1 steps ahead; the 16 warps (4 x 4, warp tile 32 x 32, as above) wait on
// the stage's "full" mbarrier, load their fragments conflict-free from the swizzled tiles, issue the DMMAs and release
// the stage on its "empty" mbarrier (the issuing thread refills the stage the CTA consumed one step earlier, so it only
// ever waits for the skew between warps).  No register staging, no shared-memory stores by the SMs, no CTA-wide barrier
// in the loop.  (A seventeenth, producer-only warp puts five warps on one scheduler and caps the kernel at 96 registers.)
// Out-of-range parts of a box (knots beyond the band or the table, profiles beyond n_p) are zero-filled by the TMA,
// which is what the predicated loads of the kernel above do.  Needs 16-byte aligned rows of A^2 (even row stride).
// ---------------------------------------------------------------------------------------------------------------
#ifndef PTT_GT_STAGES
#define PTT_GT_STAGES 4
#endif
constexpr int GT_STAGES = PTT_GT_STAGES;
constexpr int GT_CONSUMERS = 16;                          // warps
constexpr int GT_THREADS = GT_CONSUMERS * 32;
constexpr int GT_TILE_BYTES = GM_BM * GM_BK * 8;          // 16 KB: one operand tile of one stage
constexpr int GT_SMEM = GT_STAGES * 2 * GT_TILE_BYTES + 1024 /* alignment */ + 2 * GT_STAGES * 8;

__device__ __forceinline__ uint32_t gt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gt_mbar_init(uint64_t* bar, const int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gt_mbar_expect_tx(uint64_t* bar, const uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gt_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gt_mbar_wait(uint64_t* bar, const uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(gt_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// box (c0 = fastest coordinate, c1) of a 2-D tensor map -> shared memory, completion on the mbarrier
__device__ __forceinline__ void gt_tma_load_2d(void* dst, const CUtensorMap* map, const int c0, const int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(gt_smem_u32(dst)), "l"(map), "r"(gt_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// map_om: Omega scratch as [rows = n_tiles * 128][kw] doubles; map_a2: the A^2 table as [n_p][row stride] doubles with
// its first k_off columns belonging to somebody else (column offset of this pass + alignment).
__global__ void __launch_bounds__(GT_THREADS, 1) sdv_gemm_tma_kernel(
    const __grid_constant__ CUtensorMap map_om, const __grid_constant__ CUtensorMap map_a2, const int n_z, const int n_p,
    const int kw, const int n_tn, const int k_off, const int* __restrict__ e_base, const double factor,
    double* __restrict__ out, const long long out_stride) {
    extern __shared__ unsigned char gt_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gt_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + GT_STAGES * 2 * GT_TILE_BYTES);
    uint64_t* empty = full + GT_STAGES;
    const int tile_m = blockIdx.x / n_tn, tile_n = blockIdx.x % n_tn;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_k = (kw + GM_BK - 1) / GM_BK;
    if (tid == 0) {
        for (int s = 0; s < GT_STAGES; ++s) { gt_mbar_init(full + s, 1); gt_mbar_init(empty + s, GT_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int eb = e_base[tile_m] + k_off;
    auto issue = [&](const int kt) {          // thread 0: the two tiles of step kt into stage kt % GT_STAGES
        const int st = kt % GT_STAGES;
        gt_mbar_expect_tx(full + st, 2 * GT_TILE_BYTES);
        gt_tma_load_2d(base + (size_t)st * 2 * GT_TILE_BYTES, &map_om, kt * GM_BK, tile_m * GM_BM, full + st);
        gt_tma_load_2d(base + (size_t)st * 2 * GT_TILE_BYTES + GT_TILE_BYTES, &map_a2, eb + kt * GM_BK, tile_n * GM_BN, full + st);
    };
    if (tid == 0)
        for (int kt = 0; kt < min(GT_STAGES - 1, n_k); ++kt) issue(kt);
    // ---- consumers
    const int wm = warp & 3, wn = warp >> 2;          // 4 x 4 warps; warp tile 32 (m) x 32 (n)
    const int nb = min(4, max(0, (n_p - (tile_n * GM_BN + wn * 32) + 7) >> 3));      // 8-profile fragments with profiles
    const int fr = lane >> 2, fk = lane & 3;
    // element (row r, knot c) of a tile sits at r * 128 + (((c >> 1) ^ (r & 7)) << 4) + (c & 1) * 8 bytes (128-byte
    // swizzle); r & 7 == fr for every fragment row of this lane.  The four knots of DMMA step q are the columns
    // {2q, 2q + 1, 2q + 8, 2q + 9} of the tile (the order of the 16 knots inside a tile is free): the 16-byte chunks of a
    // half-warp, (q ^ fr) and (q ^ fr) ^ 4 for fr = 0..3, are then all different and a fragment load touches every bank
    // once (with the columns {4q .. 4q + 3} rows fr and fr ^ 1 share their chunk positions: 2-way conflicts).
    int koff[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) koff[q] = (((q + 4 * (fk >> 1)) ^ fr) << 4) + ((fk & 1) << 3);
    const int row_a = (wm * 32 + fr) * 128, row_b = (wn * 32 + fr) * 128;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    int s = 0;
    uint32_t ph = 0;
    for (int kt = 0; kt < n_k; ++kt) {
        if (tid == 0) {
            const int nk = kt + GT_STAGES - 1;          // refill the stage consumed in step kt - 1
            if (nk < n_k) {
                if (nk >= GT_STAGES) gt_mbar_wait(empty + nk % GT_STAGES, (uint32_t)((nk / GT_STAGES - 1) & 1));
                issue(nk);
            }
        }
        __syncwarp();
        gt_mbar_wait(full + s, ph);
        const unsigned char* As = base + (size_t)s * 2 * GT_TILE_BYTES + row_a;
        const unsigned char* Bs = base + (size_t)s * 2 * GT_TILE_BYTES + GT_TILE_BYTES + row_b;
        if (nb > 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double af[4], bf[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) af[a] = *reinterpret_cast<const double*>(As + a * 8 * 128 + koff[q]);
#pragma unroll
                for (int b = 0; b < 4; ++b) bf[b] = *reinterpret_cast<const double*>(Bs + b * 8 * 128 + koff[q]);
                if (nb == 4) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
                } else {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (b < nb) {
#pragma unroll
                            for (int a = 0; a < 4; ++a) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) gt_mbar_arrive(empty + s);
        if (++s == GT_STAGES) { s = 0; ph ^= 1u; }
    }
    // C fragment: row = lane / 4, cols = 2 * (lane % 4) + {0, 1}
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int i = tile_m * GM_BM + wm * 32 + a * 8 + fr;
        if (i >= n_z) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int p0 = tile_n * GM_BN + wn * 32 + b * 8 + 2 * fk;
            if (p0 < n_p) out[(long long)p0 * out_stride + i] = factor * acc[a][b][0];
            if (p0 + 1 < n_p) out[(long long)(p0 + 1) * out_stride + i] = factor * acc[a][b][1];
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*gt_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static gt_encode_fn gt_encoder() {
    static gt_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<gt_encode_fn>(p);
        cudaGetLastError();
    }
    return fn;
}
// [rows][cols] doubles, row stride ld doubles, box 16 (cols) x 128 (rows), 128-byte swizzle, zero fill outside
static bool gt_make_map(CUtensorMap* m, const double* base, const uint64_t cols, const uint64_t rows, const uint64_t ld) {
    gt_encode_fn enc = gt_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) || (ld & 1) || cols == 0 || rows == 0) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld * sizeof(double)};
    const cuuint32_t box[2] = {GM_BK, GM_BM};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ptt

using namespace ptt;

// a2: [n_p, a2_stride] tables of the group's profiles (device); omega: scratch of n_tiles*128*kw doubles;
// e_base: [n_tiles] int32 (device) first knot of every row tile for this v_wall (host-computed from the plan).
extern "C" int ptt_spec_den_v_group(const ptt_sdv_plan_t* plan, int n_p, const double* a2, int64_t a2_stride,
                                    double v_wall, int kw, const int32_t* e_base, double* omega, double* out,
                                    int64_t out_stride, void* stream_) {
    return ptt_spec_den_v_group_ex(plan, n_p, a2, a2_stride, v_wall, kw, e_base, omega, out, out_stride, 0, stream_);
}

extern "C" int ptt_spec_den_v_group_ex(const ptt_sdv_plan_t* plan, int n_p, const double* a2, int64_t a2_stride,
                                       double v_wall, int kw, const int32_t* e_base, double* omega, double* out,
                                       int64_t out_stride, int flags, void* stream_) {
    if (n_p == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (!plan || n_p < 0 || !a2 || !e_base || !omega || !out || kw < 2)
        return set_error(PTT_E_ARG, "ptt_spec_den_v_group: null pointer or bad size");
    if (n_p == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n_tiles = (plan->n_z + GM_BM - 1) / GM_BM;
    const double b_R = cbrt(8.0 * M_PI);
    const double bv = b_R * v_wall;
    // rows of the last tile beyond n_z are never written: the GEMM multiplies them into outputs it does not store
    double* inv_dq = omega + (size_t)n_tiles * GM_BM * (size_t)kw;
    double* inv_q = inv_dq + plan->n_q;
    sdv_recip_kernel<<<(plan->n_q + 255) / 256, 256, 0, stream>>>(plan->n_q, plan->qT, inv_dq, inv_q);
    sdv_weights_kernel<<<dim3((kw + WT_SPAN - 1) / WT_SPAN, plan->n_z), 256, 0, stream>>>(
        plan->n_q, plan->n_z, plan->nt, plan->qT, plan->z, plan->zterm, plan->T, plan->LT, plan->W, bv,
        -log10(bv) * plan->inv_dlog, kw, e_base, inv_dq, inv_q,
        reinterpret_cast<const double2*>(plan->omega_ab), omega);
    const double factor = 2.0 / (bv * bv * bv * bv * bv * bv);
    const int n_tn = (n_p + GM_BN - 1) / GM_BN;
    dim3 grid(n_tiles * n_tn);
#ifndef PTT_GEMM_NO_TMA
    if (flags & PTT_SDV_ALIGNED_BAND) {
        // operands by tensor-map copies: the caller has aligned every tile's band start (a box of doubles must start on
        // a 16-byte boundary: an odd start coordinate is an illegal instruction) and padded the row stride; a2 itself
        // may start one double past an aligned address (the lookup pass follows the y pass in the table)
        const int k_off = (reinterpret_cast<uintptr_t>(a2) & 15) ? 1 : 0;
        CUtensorMap map_om, map_a2;
        if (gt_make_map(&map_om, omega, (uint64_t)kw, (uint64_t)n_tiles * GM_BM, (uint64_t)kw) &&
            gt_make_map(&map_a2, a2 - k_off, (uint64_t)plan->n_q + k_off, (uint64_t)n_p, (uint64_t)a2_stride)) {
            const cudaError_t ea = cudaFuncSetAttribute(sdv_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM);
            if (ea != cudaSuccess) return set_cuda_error("sdv_gemm_tma_kernel smem", ea);
            sdv_gemm_tma_kernel<<<grid, GT_THREADS, GT_SMEM, stream>>>(map_om, map_a2, plan->n_z, n_p, kw, n_tn, k_off, e_base,
                                                                     factor, out, out_stride);
            const cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return set_cuda_error("sdv_gemm_tma_kernel", e);
            return PTT_OK;
        }
    }
#endif
    const int smem = 2 * (GM_BM + GM_BN) * GM_PITCH * (int)sizeof(double);
    {
        const cudaError_t e = cudaFuncSetAttribute(sdv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return set_cuda_error("sdv_gemm_kernel smem", e);
    }
    sdv_gemm_kernel<<<grid, GM_THREADS, smem, stream>>>(plan->n_z, n_p, plan->n_q, kw, n_tn, e_base, omega, a2, a2_stride, factor,
                                                     out, out_stride);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("ptt_spec_den_v_group", e);
    return PTT_OK;
}
