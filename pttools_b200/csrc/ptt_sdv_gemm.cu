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
            rb[q] = (p_ok && k < n_q && k0 + lk + q < kw) ? Bg[k] : 0.0;
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

}  // namespace ptt

using namespace ptt;

// a2: [n_p, a2_stride] tables of the group's profiles (device); omega: scratch of n_tiles*128*kw doubles;
// e_base: [n_tiles] int32 (device) first knot of every row tile for this v_wall (host-computed from the plan).
extern "C" int ptt_spec_den_v_group(const ptt_sdv_plan_t* plan, int n_p, const double* a2, int64_t a2_stride,
                                    double v_wall, int kw, const int32_t* e_base, double* omega, double* out,
                                    int64_t out_stride, void* stream_) {
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
