// CUDA kernels for the Sound Shell Model spectrum chain (all FP64).
// K5 sin_transform / A^2    <- pttools/ssmtools/calculators.py:17-258, ssm.py:35-140
// K6 spec_den_v             <- pttools/ssmtools/spectrum.py:451-486
// K7 spec_den_gw_scaled     <- spectrum.py:327-368
// K8 pow_spec / omgw0       <- spectrum.py:230-240, pttools/omgw0/omgw0_ssm.py:123-131
//
// Grids that depend only on (y, cs, nt, ...) come from the host mirror; here is the per-profile arithmetic.
// Lookups into the log-uniform tables replace np.interp's binary search by an arithmetic index; the table is
// staged in shared memory as (a_k, b_k) with  interp(x) = a_k + b_k x  on [x_k, x_{k+1}).
#include "ptt_host.cuh"

namespace ptt {

constexpr int ENV_SIZE = 16;   // xi1, xi_w, xi2, f1, f_m, f_p, f2, valid, ...
// Block-moment representation of a sampled function for the exact part of the sine transform (see K5b').
constexpr int MOM_NB = 32;                     // blocks over xi in [0, 1]
constexpr int MOM_J = 16;                      // moments per block: sum_k g_k delta_k^j / j!,  j = 0..15
constexpr int MOM_HDR = 8;                     // {b_lo, b_hi, ...}
constexpr int MOM_LEVELS = 6;                  // level l: 2^l blocks of width 2^-l; level 5 is the finest (MOM_NB)
constexpr int MOM_BLOCKS = 2 * MOM_NB - 1;     // pyramid, level l starts at block 2^l - 1
constexpr int MOM_SIZE = MOM_HDR + MOM_BLOCKS * MOM_J;
constexpr double MOM_W = 1.0 / MOM_NB;
constexpr double MOM_ZW = 1.6;                 // a block of width W serves z W <= 1.6
constexpr double MOM_ZMAX = 51.2;              // (z W / 2)^16 / 16! < 1.4e-15 for z <= 51.2 = MOM_ZW * MOM_NB
static_assert(MOM_NB == 1 << (MOM_LEVELS - 1) && MOM_J % 2 == 0, "pyramid shape");
constexpr double FOUR_PI = 12.566370614359172953850573533118;

// --- block-wide helpers -------------------------------------------------------------------------------------

template <class T, class Op>
__device__ T block_reduce(T val, Op op, T* sm) {
    // sm: blockDim.x elements
    const int tid = threadIdx.x;
    sm[tid] = val;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (tid < s) sm[tid] = op(sm[tid], sm[tid + s]);
        __syncthreads();
    }
    const T r = sm[0];
    __syncthreads();
    return r;
}

struct MinD { __device__ double operator()(double a, double b) const { return fmin(a, b); } };
struct MaxD { __device__ double operator()(double a, double b) const { return fmax(a, b); } };
struct MinI { __device__ int operator()(int a, int b) const { return a < b ? a : b; } };
struct MaxFirst {
    // (value, index) pairs packed in double2: larger value wins, ties -> smaller index (np.argmax)
    __device__ double2 operator()(double2 a, double2 b) const {
        if (b.x > a.x || (b.x == a.x && b.y < a.y)) return b;
        return a;
    }
};

// envelope() of calculators.py:17-84 for one function sampled on n points. Result in env[0..7].
__device__ void block_envelope(const double* __restrict__ xi, const double* __restrict__ f, const int n,
                               double* __restrict__ env, double* smd, double2* smd2, int* smi) {
    const int tid = threadIdx.x;
    double xmin = INFINITY, xmax = -INFINITY;
    double2 best = make_double2(-INFINITY, 1e300);
    bool has_nan = false;
    for (int k = tid; k < n; k += blockDim.x) {
        const double fk = f[k];
        if (fk != fk) has_nan = true;
        if (fk != 0.0) {
            xmin = fmin(xmin, xi[k]);
            xmax = fmax(xmax, xi[k]);
        }
        if (fk > best.x) best = make_double2(fk, (double)k);   // increasing k: first maximum is kept
    }
    xmin = block_reduce(xmin, MinD(), smd);
    xmax = block_reduce(xmax, MaxD(), smd);
    best = block_reduce(best, MaxFirst(), smd2);
    const int nan_any = __syncthreads_or(has_nan ? 1 : 0);
    int ind1 = n, ind2 = n;
    for (int k = tid; k < n; k += blockDim.x) {
        const double x = xi[k];
        if (x == xmin && k < ind1) ind1 = k;
        if (x == xmax && k < ind2) ind2 = k;
    }
    ind1 = block_reduce(ind1, MinI(), smi);
    ind2 = block_reduce(ind2, MinI(), smi);
    if (tid == 0) {
        if (nan_any || !(xmin <= xmax) || ind1 >= n || ind2 >= n || n < 3) {
            for (int j = 0; j < 8; ++j) env[j] = NAN;
        } else {
            const int i_max = (int)best.y;
            const int i1m = (ind1 - 1 + n) % n;                 // python negative-index wrap
            const double f1 = f[i1m];
            const double xi1 = xi[i1m];
            const double f_max = f[i_max];
            const double xi_w = xi[i_max];
            double df;
            if (i_max + 1 == n) df = f[i_max] - f[(i_max - 2 + n) % n];
            else df = f[i_max + 1] - f[(i_max - 1 + n) % n];
            double f_m, f_p, f2;
            if (df > 0.0) {
                f_m = f[(i_max - 1 + n) % n];
                f_p = f_max;
                f2 = f[ind2];
            } else {
                f_m = f_max;
                f_p = 0.0;
                f2 = 0.0;
            }
            env[0] = xi1; env[1] = xi_w; env[2] = xmax; env[3] = f1; env[4] = f_m; env[5] = f_p; env[6] = f2;
            env[7] = 1.0;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double envelope_transform(const double* __restrict__ env, const double z) {
    // sin_transform_approx, calculators.py:231-258
    const double xi1 = env[0], xi_w = env[1], xi2 = env[2], f1 = env[3], f_m = env[4], f_p = env[5], f2 = env[6];
    const double cw = cos(z * xi_w);
    double r = -(f2 * cos(z * xi2) - f_p * cw) / z;
    r += -(f_m * cw - f1 * cos(z * xi1)) / z;
    return r;
}

// --- K5a: per-profile preparation ------------------------------------------------------------------------------
// work row layout (doubles): [gv: row_cap][lam: row_cap][fl: nxi][hl: nxi][env_v: 16][env_l: 16][F: n_z][L: n_z]
struct WorkLayout {
    int row_cap, nxi, n_z;
    __host__ __device__ long long gv() const { return 0; }
    __host__ __device__ long long lam() const { return row_cap; }
    __host__ __device__ long long fl() const { return 2LL * row_cap; }
    __host__ __device__ long long hl() const { return 2LL * row_cap + nxi; }
    __host__ __device__ long long env_v() const { return 2LL * row_cap + 2LL * nxi; }
    __host__ __device__ long long env_l() const { return env_v() + ENV_SIZE; }
    __host__ __device__ long long F() const { return env_l() + ENV_SIZE; }
    __host__ __device__ long long L() const { return F() + n_z; }
    __host__ __device__ long long mom_v() const { return L() + n_z; }
    __host__ __device__ long long mom_l() const { return mom_v() + MOM_SIZE; }
    __host__ __device__ long long total() const { return mom_l() + MOM_SIZE; }
};

__global__ void __launch_bounds__(256) a2_prep_kernel(
    const int nxi, const int phase_rule, const double* __restrict__ xi_re, const double* __restrict__ rows_v,
    const double* __restrict__ rows_w, const double* __restrict__ rows_xi, const long long row_stride,
    const int* __restrict__ off, const int* __restrict__ npts, const double* __restrict__ pp,
    double* __restrict__ work, const long long work_stride, const WorkLayout lay) {
    __shared__ double smd[256];
    __shared__ double2 smd2[256];
    __shared__ int smi[256];
    const int p = blockIdx.x;
    const int tid = threadIdx.x;
    const int N = npts[p];
    const long long base = (long long)p * row_stride + off[p];
    const double* v = rows_v + base;
    const double* w = rows_w + base;
    const double* xi = rows_xi + base;
    double* wk = work + (long long)p * work_stride;
    double* gv = wk + lay.gv();
    double* lam = wk + lay.lam();
    double* fl = wk + lay.fl();
    double* hl = wk + lay.hl();
    double* env_v = wk + lay.env_v();
    double* env_l = wk + lay.env_l();
    if (N < 3) {   // failed shell (the reference returns a length-1 NaN array): poison everything downstream
        if (tid < ENV_SIZE) { env_v[tid] = NAN; env_l[tid] = NAN; }
        for (int m = tid; m < nxi; m += blockDim.x) { fl[m] = NAN; hl[m] = NAN; }
        return;
    }
    const double v_wall = pp[8 * p + 0];
    const double ecs = pp[8 * p + 2], ecb = pp[8 * p + 3], Vs = pp[8 * p + 4], Vb = pp[8 * p + 5];

    // phase (0 symmetric, 1 broken)
    int i_wall = N;
    for (int k = tid; k < N; k += blockDim.x)
        if (xi[k] >= v_wall && k < i_wall) i_wall = k;
    i_wall = block_reduce(i_wall, MinI(), smi);
    if (i_wall >= N) i_wall = 0;            // np.argmax of an all-False mask
    int n_broken;                           // samples [0, n_broken) are in the broken phase (rule 1)
    if (phase_rule == 1) {
        // props.find_phase, props.py:22-32
        if (i_wall == 0) n_broken = 0;
        else {
            n_broken = i_wall - 1;
            const double a = xi[i_wall];
            if (fabs(a - v_wall) <= 1e-8 + 1e-5 * fabs(v_wall)) n_broken = i_wall;   // np.isclose
        }
    } else {
        n_broken = -1;                      // rule 0: per-sample test xi < v_wall (boundary.get_phase)
    }
    const double w_last = w[N - 1];
    double e_last;
    {
        const double xl = xi[N - 1];
        const bool broken = (phase_rule == 1) ? (N - 1 < n_broken) : (xl < v_wall);
        e_last = broken ? ecb * w_last + Vb : ecs * w_last + Vs;
    }
    for (int k = tid; k < N; k += blockDim.x) {
        const double xk = xi[k], vk = v[k], wk_ = w[k];
        const bool broken = (phase_rule == 1) ? (k < n_broken) : (xk < v_wall);
        const double e = broken ? ecb * wk_ + Vb : ecs * wk_ + Vs;
        // ssm.py:59-62: lam = (e - e[-1])/w[-1] + w v^2 / w[-1]
        double l = (e - e_last) / w_last;
        l += wk_ * vk * vk / w_last;
        lam[k] = l;
        const double xl = (k > 0) ? xi[k - 1] : xk;
        const double xr = (k < N - 1) ? xi[k + 1] : xk;
        gv[k] = 0.5 * (xr - xl) * vk;       // trapezoid weight times f
    }
    __syncthreads();
    // resample_uniform_xi (calculators.py:87-101): np.interp(xi_re, xi, lam)
    const double x_first = xi[0], x_last = xi[N - 1];
    for (int m = tid; m < nxi; m += blockDim.x) {
        const double x = xi_re[m];
        double val;
        if (x <= x_first) val = lam[0];                // left fill (x == xp[0] gives fp[0] as well)
        else if (x >= x_last) val = lam[N - 1];
        else {
            int lo = 0, hi = N;                         // upper bound: first index with xi > x
            while (lo < hi) {
                const int mid = lo + ((hi - lo) >> 1);
                if (x >= xi[mid]) lo = mid + 1; else hi = mid;
            }
            const int j = lo - 1;
            const double slope = (lam[j + 1] - lam[j]) / (xi[j + 1] - xi[j]);
            val = slope * (x - xi[j]) + lam[j];
        }
        const double f = x * val;
        fl[m] = f;
        const double xl = (m > 0) ? xi_re[m - 1] : x;
        const double xr = (m < nxi - 1) ? xi_re[m + 1] : x;
        hl[m] = 0.5 * (xr - xl) * f;
    }
    __syncthreads();
    block_envelope(xi, v, N, env_v, smd, smd2, smi);
    block_envelope(xi_re, fl, nxi, env_l, smd, smd2, smi);
}

// Generic variant for the array operator sin_transform(z, xi, f): weights and envelope only.
__global__ void __launch_bounds__(256) st_prep_kernel(const double* __restrict__ xi_all, const double* __restrict__ f_all,
                                                      const long long stride, const int* __restrict__ off,
                                                      const int* __restrict__ npts, double* __restrict__ g_all,
                                                      double* __restrict__ env_all) {
    __shared__ double smd[256];
    __shared__ double2 smd2[256];
    __shared__ int smi[256];
    const int p = blockIdx.x;
    const int N = npts[p];
    const long long base = (long long)p * stride + off[p];
    const double* xi = xi_all + base;
    const double* f = f_all + base;
    double* g = g_all + base;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const double xk = xi[k];
        const double xl = (k > 0) ? xi[k - 1] : xk;
        const double xr = (k < N - 1) ? xi[k + 1] : xk;
        g[k] = 0.5 * (xr - xl) * f[k];
    }
    __syncthreads();
    block_envelope(xi, f, N, env_all + (long long)p * ENV_SIZE, smd, smd2, smi);
}

// --- K5b: sine transforms -----------------------------------------------------------------------------------
// One thread per wavenumber; the profile streams through shared memory in tiles and every lane accumulates
// sum_k g_k sin(z xi_k).  Function A is the ragged profile (own xi grid), function B (optional) lives on the
// shared uniform grid xi_re.  Output = (4 pi / z) * [ (1-frac) * exact + frac * asymptotic ]  when scale4pi,
// else the bare transform.
constexpr int ST_THREADS = 128;
constexpr int ST_TILE = 1024;

__global__ void __launch_bounds__(ST_THREADS) sin_transform_kernel(
    const int n_z, const double* __restrict__ z_all, const double* __restrict__ frac_all,
    const double* __restrict__ xiA_all, const long long strideA, const double* __restrict__ gA_all,
    const long long g_strideA, const int g_has_off,
    const int* __restrict__ off, const int* __restrict__ npts, const double* __restrict__ envA_all,
    const long long env_strideA, double* __restrict__ outA, const long long out_strideA,
    const int nB, const double* __restrict__ xiB, const double* __restrict__ gB_all, const long long strideB,
    const double* __restrict__ envB_all, const long long env_strideB, double* __restrict__ outB,
    const long long out_strideB, const int scale4pi) {
    __shared__ double s_x[ST_TILE];
    __shared__ double s_g[ST_TILE];
    const int p = blockIdx.y;
    const int iz = blockIdx.x * ST_THREADS + threadIdx.x;
    const bool active = iz < n_z;
    const double z = active ? z_all[iz] : 1.0;
    const double frac = active ? frac_all[iz] : 1.0;
    const int need_exact = __syncthreads_or((active && frac < 1.0) ? 1 : 0);

    const int N = npts[p];
    const double* envA = envA_all + (long long)p * env_strideA;
    double accA = 0.0, accB = 0.0;
    if (need_exact) {
        const double* xi = xiA_all + (long long)p * strideA + off[p];
        const double* g = gA_all + (long long)p * g_strideA + (g_has_off ? off[p] : 0);
        for (int t0 = 0; t0 < N; t0 += ST_TILE) {
            const int m = min(ST_TILE, N - t0);
            for (int k = threadIdx.x; k < m; k += ST_THREADS) { s_x[k] = xi[t0 + k]; s_g[k] = g[t0 + k]; }
            __syncthreads();
            if (frac < 1.0) {
#pragma unroll 4
                for (int k = 0; k < m; ++k) {
                    const double gk = s_g[k];
                    if (gk != 0.0) accA = fma(gk, sin(z * s_x[k]), accA);
                }
            }
            __syncthreads();
        }
        if (outB) {
            const double* g2 = gB_all + (long long)p * strideB;
            for (int t0 = 0; t0 < nB; t0 += ST_TILE) {
                const int m = min(ST_TILE, nB - t0);
                for (int k = threadIdx.x; k < m; k += ST_THREADS) { s_x[k] = xiB[t0 + k]; s_g[k] = g2[t0 + k]; }
                __syncthreads();
                if (frac < 1.0) {
#pragma unroll 4
                    for (int k = 0; k < m; ++k) accB = fma(s_g[k], sin(z * s_x[k]), accB);
                }
                __syncthreads();
            }
        }
    }
    if (!active) return;
    const double pref = scale4pi ? FOUR_PI / z : 1.0;
    {
        double r;
        if (envA[7] != 1.0 || N < 3) r = NAN;
        else if (frac <= 0.0) r = accA;
        else if (frac >= 1.0) r = envelope_transform(envA, z);
        else r = envelope_transform(envA, z) * frac + accA * (1.0 - frac);
        outA[(long long)p * out_strideA + iz] = pref * r;
    }
    if (outB) {
        const double* envB = envB_all + (long long)p * env_strideB;
        double r;
        if (envB[7] != 1.0) r = NAN;
        else if (frac <= 0.0) r = accB;
        else if (frac >= 1.0) r = envelope_transform(envB, z);
        else r = envelope_transform(envB, z) * frac + accB * (1.0 - frac);
        outB[(long long)p * out_strideB + iz] = pref * r;
    }
}


// --- K5b': exact part of the sine transform by block moments ------------------------------------------------
// sum_k g_k sin(z xi_k) is the reference's trapezoid sum (calculators.py:208-228) with g_k = weight_k f_k.  With
// xi_k = c_b + delta_k inside block b (centre c_b, |delta_k| <= W/2):
//     sum_{k in b} g_k sin(z xi_k) = sin(z c_b) C_b(z) + cos(z c_b) S_b(z),
//     C_b = sum_m (-1)^m z^{2m}   M_{b,2m},     S_b = sum_m (-1)^m z^{2m+1} M_{b,2m+1},     M_{b,j} = sum_k g_k delta_k^j / j!
// The moments do not depend on z, 16 of them carry the series to 1.4e-15 for z W / 2 <= 0.8, and sin/cos(z c_b) advance by a
// constant rotation.  Same sum, ~N_xi / 6 times fewer flops than one sine per (z, xi) pair.
__device__ void block_moments(const double* __restrict__ xi, const double* __restrict__ g, const int n,
                              double* __restrict__ mom /* MOM_SIZE, global */, double* s_pyr /* MOM_BLOCKS*MOM_J shared */,
                              int* s_lohi /* 2 shared */) {
    // Warps walk the samples in chunks of MOM_CHUNK coalesced tiles of 32.  Consecutive samples nearly always fall in
    // the same block, so a warp keeps the moments of its current block in registers and adds them to shared memory
    // (a folding shuffle reduction + one atomic per moment) only when the block changes or the chunk ends; a tile that
    // straddles blocks is walked once per distinct block.
    constexpr int MOM_CHUNK = 4;
    __shared__ double s_pw[(MOM_LEVELS - 1) * MOM_J];   // s_pw[l][d] = (half child width at level l)^d / d!
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, n_warps = nt >> 5;
    double* s_m = s_pyr + (MOM_NB - 1) * MOM_J;       // finest level
    for (int i = tid; i < MOM_NB * MOM_J; i += nt) s_m[i] = 0.0;
    if (tid == 0) { s_lohi[0] = MOM_NB; s_lohi[1] = -1; }
    if (tid < (MOM_LEVELS - 1) * MOM_J) {
        const int l = tid / MOM_J, d = tid - l * MOM_J;
        const double sh = 0.25 / (double)(1 << l);
        double pw = 1.0;
        for (int i = 1; i <= d; ++i) pw *= sh / (double)i;
        s_pw[tid] = pw;
    }
    __syncthreads();
    double m[MOM_J];
#pragma unroll
    for (int j = 0; j < MOM_J; ++j) m[j] = 0.0;
    int cur = -1;                                      // warp-uniform: block whose moments sit in m
    static_assert(MOM_J == 16, "the butterfly below folds 16 -> 8 -> 4 -> 2 -> 1 values per lane");
    auto flush = [&]() {
        if (cur < 0) return;
        // Butterfly that halves the values per lane at every step (15 shuffles instead of 80): afterwards lane L holds
        // the warp total of moment j(L) = bit1 + 2 bit2 + 4 bit3 + 8 bit4 of the lane index (bit0: a duplicate).
        const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
        double a8[8], a4[4], a2[2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double send = h4 ? m[i] : m[i + 8], keep = h4 ? m[i + 8] : m[i];
            a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double send = h3 ? a8[i] : a8[i + 4], keep = h3 ? a8[i + 4] : a8[i];
            a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double send = h2 ? a4[i] : a4[i + 2], keep = h2 ? a4[i + 2] : a4[i];
            a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        const double send = h1 ? a2[0] : a2[1], keep = h1 ? a2[1] : a2[0];
        double tot = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
        if (!(lane & 1)) atomicAdd(&s_m[cur * MOM_J + (h1 ? 1 : 0) + (h2 ? 2 : 0) + (h3 ? 4 : 0) + (h4 ? 8 : 0)], tot);
#pragma unroll
        for (int j = 0; j < MOM_J; ++j) m[j] = 0.0;
        if (lane == 0) { atomicMin(&s_lohi[0], cur); atomicMax(&s_lohi[1], cur); }
        cur = -1;
    };
    const int n_tiles = (n + 31) >> 5;
    for (int t0 = warp * MOM_CHUNK; t0 < n_tiles; t0 += n_warps * MOM_CHUNK) {
        const int t1 = min(n_tiles, t0 + MOM_CHUNK);
        for (int t = t0; t < t1; ++t) {
            const int k = (t << 5) + lane;
            const double gk = (k < n) ? g[k] : 0.0;
            const double x = (k < n) ? xi[k] : 0.0;
            const int b = (gk != 0.0) ? max(0, min(MOM_NB - 1, (int)(x * MOM_NB))) : -1;
            unsigned act = __ballot_sync(0xffffffffu, b >= 0);
            while (act) {                              // one pass per distinct block in the tile (nearly always one)
                const int b0 = __shfl_sync(0xffffffffu, b, __ffs(act) - 1);
                if (b0 != cur) { flush(); cur = b0; }
                if (b == b0) {
                    const double d = x - ((double)b0 + 0.5) * MOM_W;
                    double tt = gk;
#pragma unroll
                    for (int j = 0; j < MOM_J; ++j) {
                        m[j] += tt;
                        tt *= d * (1.0 / (double)(j + 1));      // compile-time reciprocals: no division in the loop
                    }
                }
                act &= ~__ballot_sync(0xffffffffu, b == b0);
            }
        }
        flush();
    }
    __syncthreads();
    // Coarser levels by merging pairs of blocks: with delta' = delta + s (s = -+ half a child width),
    //   M'_j = sum_children sum_{i<=j} s^(j-i)/(j-i)! M_i   - exact algebra, no new truncation.
    for (int l = MOM_LEVELS - 2; l >= 0; --l) {
        const int nb = 1 << l;
        const double* pwt = s_pw + l * MOM_J;
        const double* child = s_pyr + (2 * nb - 1) * MOM_J;
        double* par = s_pyr + (nb - 1) * MOM_J;
        for (int i = tid; i < nb * MOM_J; i += nt) {
            const int b = i / MOM_J, j = i - b * MOM_J;
            const double* cl = child + (2 * b) * MOM_J;
            const double* cr = cl + MOM_J;
            double accl = 0.0, accr = 0.0;
            for (int q = j; q >= 0; --q) {
                // left child: centre = parent centre - sh  =>  delta' = delta - sh;  right child: delta' = delta + sh
                const double pw = pwt[j - q];
                accl = fma(((j - q) & 1) ? -pw : pw, cl[q], accl);
                accr = fma(pw, cr[q], accr);
            }
            par[i] = accl + accr;
        }
        __syncthreads();
    }
    for (int i = tid; i < MOM_BLOCKS * MOM_J; i += nt) mom[MOM_HDR + i] = s_pyr[i];
    if (tid == 0) {
        mom[0] = (double)s_lohi[0];
        mom[1] = (double)s_lohi[1];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) a2_moments_kernel(const int nxi, const double* __restrict__ xi_re,
                                                         const double* __restrict__ rows_xi, const long long row_stride,
                                                         const int* __restrict__ off, const int* __restrict__ npts,
                                                         double* __restrict__ work, const long long work_stride,
                                                         const WorkLayout lay) {
    __shared__ double s_m[MOM_BLOCKS * MOM_J];
    __shared__ int s_lohi[2];
    const int p = blockIdx.x;
    const int N = npts[p];
    double* wk = work + (long long)p * work_stride;
    if (N < 3) {
        if (threadIdx.x == 0) { wk[lay.mom_v()] = 1.0; wk[lay.mom_v() + 1] = 0.0; wk[lay.mom_l()] = 1.0; wk[lay.mom_l() + 1] = 0.0; }
        return;
    }
    const double* xi = rows_xi + (long long)p * row_stride + off[p];
    block_moments(xi, wk + lay.gv(), N, wk + lay.mom_v(), s_m, s_lohi);
    block_moments(xi_re, wk + lay.hl(), nxi, wk + lay.mom_l(), s_m, s_lohi);
}

constexpr int STM_THREADS = 256;

// exact transform of one function from its moment pyramid, for this thread's z: the coarsest level whose blocks
// satisfy z W <= MOM_ZW
__device__ __forceinline__ double moment_transform(const double* __restrict__ mom, const double z) {
    const int b_lo7 = (int)mom[0], b_hi7 = (int)mom[1];
    if (b_hi7 < b_lo7) return 0.0;
    int l = 0;
    while (l < MOM_LEVELS - 1 && z > MOM_ZW * (double)(1 << l)) ++l;
    const int nb = 1 << l;
    const double W = 1.0 / (double)nb;
    const int b_lo = b_lo7 >> (MOM_LEVELS - 1 - l), b_hi = b_hi7 >> (MOM_LEVELS - 1 - l);
    const double* lev = mom + MOM_HDR + (nb - 1) * MOM_J;
    const double z2 = -z * z;
    double sb, cb;
    sincos(z * ((double)b_lo + 0.5) * W, &sb, &cb);
    double sw, cw;
    sincos(z * W, &sw, &cw);
    double acc = 0.0;
    for (int b = b_lo; b <= b_hi; ++b) {
        const double* m = lev + b * MOM_J;
        // Horner in -z^2: even and odd moments
        double C = m[MOM_J - 2], S = m[MOM_J - 1];
#pragma unroll
        for (int j = MOM_J - 4; j >= 0; j -= 2) { C = fma(C, z2, m[j]); S = fma(S, z2, m[j + 1]); }
        acc = fma(sb, C, acc);
        acc = fma(cb * z, S, acc);
        const double sn = fma(sb, cw, cb * sw);
        cb = fma(cb, cw, -sb * sw);
        sb = sn;
    }
    return acc;
}

// Two wavenumbers per thread from one pass over the blocks (the level is chosen for the larger of the two): the
// moment loads - uniform across the warp, but each still a full write-back into the register file - are shared.
__device__ __forceinline__ void moment_transform2(const double* __restrict__ mom, const double za, const double zb,
                                                  double& ra, double& rb) {
    ra = 0.0; rb = 0.0;
    const int b_lo7 = (int)mom[0], b_hi7 = (int)mom[1];
    if (b_hi7 < b_lo7) return;
    const double zmax = fmax(za, zb);
    int l = 0;
    while (l < MOM_LEVELS - 1 && zmax > MOM_ZW * (double)(1 << l)) ++l;
    const int nb = 1 << l;
    const double W = 1.0 / (double)nb;
    const int b_lo = b_lo7 >> (MOM_LEVELS - 1 - l), b_hi = b_hi7 >> (MOM_LEVELS - 1 - l);
    const double* lev = mom + MOM_HDR + (nb - 1) * MOM_J;
    const double za2 = -za * za, zb2 = -zb * zb;
    const double c0 = ((double)b_lo + 0.5) * W;
    double sa, ca, sb, cb, swa, cwa, swb, cwb;
    sincos(za * c0, &sa, &ca);
    sincos(zb * c0, &sb, &cb);
    sincos(za * W, &swa, &cwa);
    sincos(zb * W, &swb, &cwb);
    for (int b = b_lo; b <= b_hi; ++b) {
        const double* m = lev + b * MOM_J;
        double mj[MOM_J];
#pragma unroll
        for (int j = 0; j < MOM_J; ++j) mj[j] = m[j];
        double C = mj[MOM_J - 2], S = mj[MOM_J - 1];
#pragma unroll
        for (int j = MOM_J - 4; j >= 0; j -= 2) { C = fma(C, za2, mj[j]); S = fma(S, za2, mj[j + 1]); }
        ra = fma(sa, C, ra);
        ra = fma(ca * za, S, ra);
        C = mj[MOM_J - 2]; S = mj[MOM_J - 1];
#pragma unroll
        for (int j = MOM_J - 4; j >= 0; j -= 2) { C = fma(C, zb2, mj[j]); S = fma(S, zb2, mj[j + 1]); }
        rb = fma(sb, C, rb);
        rb = fma(cb * zb, S, rb);
        double sn = fma(sa, cwa, ca * swa);
        ca = fma(ca, cwa, -sa * swa);
        sa = sn;
        sn = fma(sb, cwb, cb * swb);
        cb = fma(cb, cwb, -sb * swb);
        sb = sn;
    }
}

// Same interface role as sin_transform_kernel for the A^2 path (functions A = v on the profile grid, B = xi*lambda
// on the uniform grid), exact part from block moments.
__global__ void __launch_bounds__(STM_THREADS) sin_transform_moments_kernel(
    const int n_z, const double* __restrict__ z_all, const double* __restrict__ frac_all,
    const int* __restrict__ npts, double* __restrict__ work, const long long work_stride, const WorkLayout lay) {
    const int p = blockIdx.y;
    double* wk = work + (long long)p * work_stride;
    const int iz0 = 2 * (blockIdx.x * STM_THREADS + threadIdx.x);        // two adjacent wavenumbers per thread
    if (iz0 >= n_z) return;
    const int n_here = min(2, n_z - iz0);
    const int N = npts[p];
    const double* envA = wk + lay.env_v();
    const double* envB = wk + lay.env_l();
    double z[2], frac[2], accA[2] = {0.0, 0.0}, accB[2] = {0.0, 0.0};
    for (int q = 0; q < 2; ++q) {
        z[q] = (q < n_here) ? z_all[iz0 + q] : 1.0;
        frac[q] = (q < n_here) ? frac_all[iz0 + q] : 1.0;
    }
    if (N >= 3) {
        // the lanes of a warp walk the same blocks of the same level (z is ascending): uniform, L1-resident loads
        if (frac[0] < 1.0 && frac[1] < 1.0) {
            moment_transform2(wk + lay.mom_v(), z[0], z[1], accA[0], accA[1]);
            moment_transform2(wk + lay.mom_l(), z[0], z[1], accB[0], accB[1]);
        } else {
            for (int q = 0; q < 2; ++q)
                if (frac[q] < 1.0) {
                    accA[q] = moment_transform(wk + lay.mom_v(), z[q]);
                    accB[q] = moment_transform(wk + lay.mom_l(), z[q]);
                }
        }
    }
    for (int q = 0; q < n_here; ++q) {
        const double pref = FOUR_PI / z[q];
        double rA, rB;
        if (envA[7] != 1.0 || N < 3) rA = NAN;
        else if (frac[q] <= 0.0) rA = accA[q];
        else if (frac[q] >= 1.0) rA = envelope_transform(envA, z[q]);
        else rA = envelope_transform(envA, z[q]) * frac[q] + accA[q] * (1.0 - frac[q]);
        if (envB[7] != 1.0 || N < 3) rB = NAN;
        else if (frac[q] <= 0.0) rB = accB[q];
        else if (frac[q] >= 1.0) rB = envelope_transform(envB, z[q]);
        else rB = envelope_transform(envB, z[q]) * frac[q] + accB[q] * (1.0 - frac[q]);
        wk[lay.F() + iz0 + q] = pref * rA;
        wk[lay.L() + iz0 + q] = pref * rB;
    }
}

// Array-operator variant (sin_transform(z, xi, f) on profiles with 0 <= xi <= 1): moments of g = weight * f, then the
// same per-wavenumber evaluation.  A profile with a sample outside [0, 1] is flagged and comes back as NaN.
__global__ void __launch_bounds__(256) st_moments_kernel(const double* __restrict__ xi_all,
                                                         const double* __restrict__ g_all, const long long stride,
                                                         const int* __restrict__ off, const int* __restrict__ npts,
                                                         double* __restrict__ mom_all) {
    __shared__ double s_m[MOM_BLOCKS * MOM_J];
    __shared__ int s_lohi[2];
    const int p = blockIdx.x;
    const int N = npts[p];
    const long long base = (long long)p * stride + off[p];
    const double* xi = xi_all + base;
    double* mom = mom_all + (long long)p * MOM_SIZE;
    int bad = 0;
    for (int k = threadIdx.x; k < N; k += blockDim.x)
        if (!(xi[k] >= 0.0 && xi[k] <= 1.0)) bad = 1;
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) mom[2] = bad ? 1.0 : 0.0;
    if (bad) {
        if (threadIdx.x == 0) { mom[0] = 1.0; mom[1] = 0.0; }
        return;
    }
    block_moments(xi, g_all + base, N, mom, s_m, s_lohi);
}

__global__ void __launch_bounds__(STM_THREADS) st_transform_moments_kernel(
    const int n_z, const double* __restrict__ z_all, const double* __restrict__ frac_all, const int* __restrict__ npts,
    const double* __restrict__ mom_all, const double* __restrict__ env_all, double* __restrict__ out) {
    const int p = blockIdx.y;
    const int iz = blockIdx.x * STM_THREADS + threadIdx.x;
    if (iz >= n_z) return;
    const double z = z_all[iz], frac = frac_all[iz];
    const double* mom = mom_all + (long long)p * MOM_SIZE;
    const double* env = env_all + (long long)p * ENV_SIZE;
    const int N = npts[p];
    double r;
    if (env[7] != 1.0 || N < 3 || mom[2] != 0.0 || (frac < 1.0 && z > MOM_ZMAX)) r = NAN;
    else if (frac <= 0.0) r = moment_transform(mom, z);
    else if (frac >= 1.0) r = envelope_transform(env, z);
    else r = envelope_transform(env, z) * frac + moment_transform(mom, z) * (1.0 - frac);
    out[(long long)p * n_z + iz] = r;
}

// --- K5c: |A|^2 from f and l (ssm.py:56, 74-81; speedup/functions.py:10-20) ----------------------------------
constexpr int A2F_PPT = 8;      // profiles per thread: the stencil and 1/dz of a wavenumber are shared by all profiles

__global__ void __launch_bounds__(256) a2_finish_kernel(
    const int P, const int n_z, const int n_seg, const int* __restrict__ seg, const double* __restrict__ z,
    const double* __restrict__ work, const long long work_stride, const long long offF, const long long offL,
    const double* __restrict__ pp, double* __restrict__ a2, const long long a2_stride, double* __restrict__ fp2_2,
    double* __restrict__ lam2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_z) return;
    int s0 = 0, s1 = n_z;
    for (int s = 0; s < n_seg; ++s)
        if (i >= seg[s] && i < seg[s + 1]) { s0 = seg[s]; s1 = seg[s + 1]; }
    // np.gradient stencil of this wavenumber: (F[ia] - F[ib]) * wgt / gz, the factor wgt / gz formed once (<= 1 ulp from
    // dividing per profile)
    int ia, ib;
    double wgt, gz;
    if (s1 - s0 < 2) { ia = ib = i; wgt = NAN; gz = 1.0; }
    else if (i == s0) { ia = i + 1; ib = i; wgt = 1.0; gz = z[i + 1] - z[i]; }
    else if (i == s1 - 1) { ia = i; ib = i - 1; wgt = 1.0; gz = z[i] - z[i - 1]; }
    else { ia = i + 1; ib = i - 1; wgt = 0.5; gz = (z[i + 1] - z[i - 1]) / 2.0; }
    const double scale = wgt / gz;
    const int p0 = blockIdx.y * A2F_PPT, p1 = min(P, p0 + A2F_PPT);
    for (int p = p0; p < p1; ++p) {
        const double* F = work + (long long)p * work_stride + offF;
        const double* L = work + (long long)p * work_stride + offL;
        const double v_ft = (F[ia] - F[ib]) * scale;
        const double cs = pp[8 * p + 1];
        const double cl = cs * L[i];
        const long long o = (long long)p * n_z + i;
        a2[(long long)p * a2_stride + i] = 0.25 * (v_ft * v_ft + cl * cl);
        if (fp2_2) fp2_2[o] = v_ft * v_ft / 2.0;
        if (lam2) lam2[o] = cl * cl / 2.0;
    }
}

// --- K6: spec_den_v -------------------------------------------------------------------------------------------
// CTA = (z-tile, point).  The window of the A^2 lookup that the tile can touch is staged in shared memory as
// (a_k, b_k); entry 0 and entry n_q are the np.interp end clamps.  Thread <-> z_i, sequential loop over T_j.
constexpr int SDV_THREADS = 512;
constexpr int SDV_TT = 256;     // T samples staged per tile

__global__ void __launch_bounds__(SDV_THREADS) spec_den_v_kernel(
    const int n_q, const int n_z, const int nt, const int z_tile, const double* __restrict__ qT,
    const double* __restrict__ z, const double* __restrict__ zterm, const double* __restrict__ T,
    const double* __restrict__ LT, const double* __restrict__ W, const double* __restrict__ a2,
    const long long a2_stride, const double* __restrict__ v_wall, double* __restrict__ out,
    const long long out_stride, const double inv_dlog, const int win_cap) {
    extern __shared__ double2 s_tab[];                 // [win_cap] lookup window
    __shared__ double s_T[SDV_TT], s_LT[SDV_TT], s_W[SDV_TT];
    const int p = blockIdx.y;
    const int i0 = blockIdx.x * z_tile;
    const int i1 = min(n_z, i0 + z_tile);
    const double vw = v_wall[p];
    const double b_R = cbrt(8.0 * M_PI);               // spectrum.py:474
    const double bv = b_R * vw;
    const double shift = -log10(bv) * inv_dlog;
    const double* fp = a2 + (long long)p * a2_stride;

    // window of table entries (entry index = knot index + 1)
    double umin = INFINITY, umax = -INFINITY;
    for (int i = i0 + threadIdx.x; i < i1; i += SDV_THREADS) {
        const double zt = zterm[i];
        umin = fmin(umin, zt);
        umax = fmax(umax, zt);
    }
    __shared__ double s_red[SDV_THREADS];
    umin = block_reduce(umin, MinD(), s_red);
    umax = block_reduce(umax, MaxD(), s_red);
    int e_lo = (int)floor(umin + LT[0] + shift) - 1 + 1;          // -1 safety, +1 entry offset
    int e_hi = (int)floor(umax + LT[nt - 1] + shift) + 1 + 1;
    e_lo = max(0, min(n_q, e_lo));
    e_hi = max(0, min(n_q, e_hi));
    const int n_win = e_hi - e_lo + 1;
    if (n_win > win_cap) {                                        // host mis-sized the tile: fail loudly
        for (int i = i0 + threadIdx.x; i < i1; i += SDV_THREADS) out[(long long)p * out_stride + i] = NAN;
        return;
    }
    for (int e = e_lo + threadIdx.x; e <= e_hi; e += SDV_THREADS) {
        double2 ab;
        if (e == 0) ab = make_double2(fp[0], 0.0);
        else if (e == n_q) ab = make_double2(fp[n_q - 1], 0.0);
        else {
            const int k = e - 1;
            const double x0 = qT[k], x1 = qT[k + 1], f0 = fp[k], f1 = fp[k + 1];
            const double b = (f1 - f0) / (x1 - x0);
            ab = make_double2(f0 - b * x0, b);
        }
        s_tab[e - e_lo] = ab;
    }
    __syncthreads();

    const double factor = 2.0 / (bv * bv * bv * bv * bv * bv);    // spectrum.py:480-481
    for (int ib = i0; ib < i1; ib += SDV_THREADS) {
        const int i = ib + threadIdx.x;
        const bool act = i < i1;
        const double s = act ? z[i] / bv : 1.0;
        const double ub = act ? zterm[i] + shift + 1.0 - (double)e_lo : 0.0;   // entry-relative index base
        double acc = 0.0;
        for (int j0 = 0; j0 < nt; j0 += SDV_TT) {
            const int m = min(SDV_TT, nt - j0);
            __syncthreads();
            for (int j = threadIdx.x; j < m; j += SDV_THREADS) { s_T[j] = T[j0 + j]; s_LT[j] = LT[j0 + j]; s_W[j] = W[j0 + j]; }
            __syncthreads();
            if (act) {
#pragma unroll 4
                for (int j = 0; j < m; ++j) {
                    const double u = ub + s_LT[j];
                    int e = __double2int_rd(u);
                    e = max(0, min(n_win - 1, e));
                    const double2 ab = s_tab[e];
                    const double x = s * s_T[j];
                    acc = fma(fma(ab.y, x, ab.x), s_W[j], acc);
                }
            }
        }
        if (act) out[(long long)p * out_stride + i] = acc * factor;
    }
}

// --- K7: spec_den_gw_scaled -----------------------------------------------------------------------------------
// CTA = (tile of GW_YT wavenumbers y, point): the window of the P_v lookup the tile can touch sits in shared memory as
// (a_k, b_k).  A warp owns GW_R consecutive y_i; its lanes stride over the integration variable zeta_k, so the
// per-k tables (zeta_k, log zeta'_k, G_k) are read coalesced once per GW_R evaluations and neighbouring lanes hit the
// same or adjacent lookup entries (broadcast, no bank conflicts); a shuffle reduction closes the trapezoid sum.
constexpr int GW_THREADS = 256;
constexpr int GW_R = 4;
constexpr int GW_YT = (GW_THREADS / 32) * GW_R;     // 32 wavenumbers per CTA

__global__ void __launch_bounds__(GW_THREADS) spec_den_gw_kernel(
    const int n_zl, const int n_y, const int n_int, const int y_tile, const double* __restrict__ z_lookup,
    const double* __restrict__ y, const double* __restrict__ yterm, const double* __restrict__ L1,
    const double* __restrict__ L2, const double* __restrict__ Z1, const double* __restrict__ Z2,
    const double* __restrict__ G, const double prefactor, const double* __restrict__ pv, const long long pv_stride,
    const double* __restrict__ slf, const double* __restrict__ scale, double* __restrict__ sd_gw,
    double* __restrict__ pow_gw, double* __restrict__ omgw0, const int win_cap) {
    extern __shared__ double2 s_tab[];
    __shared__ double s_red[GW_THREADS];
    const int p = blockIdx.y;
    const int i0 = blockIdx.x * y_tile;
    const int i1 = min(n_y, i0 + y_tile);
    const double* fp = pv + (long long)p * pv_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    double umin = INFINITY, umax = -INFINITY;
    for (int i = i0 + threadIdx.x; i < i1; i += GW_THREADS) {
        const double t = yterm[i];
        umin = fmin(umin, t);
        umax = fmax(umax, t);
    }
    umin = block_reduce(umin, MinD(), s_red);
    umax = block_reduce(umax, MaxD(), s_red);
    // zeta ranges: L1 ascending from log zeta_- to log zeta_+, L2 descending between the same bounds
    const double lmin = fmin(L1[0], L2[n_int - 1]), lmax = fmax(L1[n_int - 1], L2[0]);
    int e_lo = (int)floor(umin + lmin) - 1 + 1;
    int e_hi = (int)floor(umax + lmax) + 1 + 1;
    e_lo = max(0, min(n_zl, e_lo));
    e_hi = max(0, min(n_zl, e_hi));
    const int n_win = e_hi - e_lo + 1;
    if (n_win > win_cap) {                       // host mis-sized the tile: fail loudly
        for (int i = i0 + threadIdx.x; i < i1; i += GW_THREADS) {
            const long long o = (long long)p * n_y + i;
            if (sd_gw) sd_gw[o] = NAN;
            if (pow_gw) pow_gw[o] = NAN;
            if (omgw0) omgw0[o] = NAN;
        }
        return;
    }
    for (int e = e_lo + threadIdx.x; e <= e_hi; e += GW_THREADS) {
        double2 ab;
        if (e == 0) ab = make_double2(fp[0], 0.0);
        else if (e == n_zl) ab = make_double2(fp[n_zl - 1], 0.0);
        else {
            const int k = e - 1;
            const double x0 = z_lookup[k], x1 = z_lookup[k + 1], f0 = fp[k], f1 = fp[k + 1];
            const double b = (f1 - f0) / (x1 - x0);
            ab = make_double2(f0 - b * x0, b);
        }
        s_tab[e - e_lo] = ab;
    }
    __syncthreads();

    for (int ib = i0 + warp * GW_R; ib < i1; ib += (GW_THREADS / 32) * GW_R) {
        double yi[GW_R], ub[GW_R], acc[GW_R];
#pragma unroll
        for (int r = 0; r < GW_R; ++r) {
            const int i = min(ib + r, n_y - 1);
            yi[r] = y[i];
            ub[r] = yterm[i] + 1.0 - (double)e_lo;
            acc[r] = 0.0;
        }
        // L1 is linspace(log zeta_-, log zeta_+)/dlog: regenerate it (same formula as numpy's linspace), and
        // zeta'_k = (zeta_+ + zeta_-) - zeta_k: two of the five per-k tables need no load
        const double l1_0 = L1[0];
        const double dl1 = (n_int > 1) ? (L1[n_int - 1] - l1_0) / (double)(n_int - 1) : 0.0;
        const double zsum = Z1[0] + Z2[0];
#pragma unroll 2
        for (int k = lane; k < n_int; k += 32) {
            const double l2 = L2[k], z1 = Z1[k], g = G[k];
            const double l1 = fma((double)k, dl1, l1_0);
            const double z2 = zsum - z1;
#pragma unroll
            for (int r = 0; r < GW_R; ++r) {
                int e1 = __double2int_rd(ub[r] + l1);
                int e2 = __double2int_rd(ub[r] + l2);
                e1 = max(0, min(n_win - 1, e1));
                e2 = max(0, min(n_win - 1, e2));
                const double2 t1 = s_tab[e1];
                const double2 t2 = s_tab[e2];
                const double p1 = fma(t1.y, yi[r] * z1, t1.x);
                const double p2 = fma(t2.y, yi[r] * z2, t2.x);
                acc[r] = fma(g * p1, p2, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < GW_R; ++r) {
            double a = acc[r];
            for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
            const int i = ib + r;
            if (lane == 0 && i < i1) {
                // p_gw = pref/y * trapz(...) with z = y zeta  =>  pref * y^2 * sum_k G_k P(y zeta_k) P(y zeta'_k)
                const double sd = prefactor * yi[r] * yi[r] * a * slf[p];
                const long long o = (long long)p * n_y + i;
                if (sd_gw) sd_gw[o] = sd;
                const double pw = yi[r] * yi[r] * yi[r] / (2.0 * M_PI * M_PI) * sd;
                if (pow_gw) pow_gw[o] = pw;
                if (omgw0 && scale) omgw0[o] = scale[p] * pw;
            }
        }
    }
}

__global__ void pow_spec_kernel(const int n_z, const double* __restrict__ z, const double* __restrict__ sd,
                                const long long stride, double* __restrict__ out) {
    const int p = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_z) return;
    const double zi = z[i];
    out[(long long)p * stride + i] = zi * zi * zi / (2.0 * M_PI * M_PI) * sd[(long long)p * stride + i];
}

// FP64 FMA probe: 8 independent chains per thread.
__global__ void fp64_probe_kernel(const int iters, double* sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

}  // namespace ptt

using namespace ptt;

static int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(what, e);
    return PTT_OK;
}

// The batch index of these operators is gridDim.y (limit 65535): larger batches are processed in slices.
constexpr int MAX_BATCH_SLICE = 65535;

// The array operators take their scratch from the device's default memory pool (stream-ordered).  By default the pool
// hands freed memory back to the driver at every synchronisation, so each call would pay a fresh allocation of hundreds
// of MB: keep it (once per device).
static void keep_pool_memory() {
    static int done_for = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    done_for = dev;
}

extern "C" int ptt_sin_transform_batch(int P, const double* xi, const double* f, int64_t stride, const int32_t* off,
                                       const int32_t* npts, int n_z, const double* z, const double* frac,
                                       double* out, void* stream_) {
    if (P == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (P < 0 || n_z < 0 || !xi || !f || !off || !npts || !z || !frac || !out)
        return set_error(PTT_E_ARG, "ptt_sin_transform_batch: null pointer or bad size");
    for (int p0 = 0; P > MAX_BATCH_SLICE && p0 < P; p0 += MAX_BATCH_SLICE) {
        const int rc = ptt_sin_transform_batch(min(MAX_BATCH_SLICE, P - p0), xi + (size_t)p0 * stride, f + (size_t)p0 * stride,
                                               stride, off + p0, npts + p0, n_z, z, frac, out + (size_t)p0 * n_z, stream_);
        if (rc != PTT_OK || p0 + MAX_BATCH_SLICE >= P) return rc;
    }
    if (P == 0 || n_z == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    double* g = nullptr;
    double* env = nullptr;
    keep_pool_memory();
    cudaError_t e = cudaMallocAsync((void**)&g, sizeof(double) * (size_t)P * (size_t)stride, stream);
    if (e != cudaSuccess) return set_cuda_error("ptt_sin_transform_batch alloc", e);
    e = cudaMallocAsync((void**)&env, sizeof(double) * (size_t)P * ENV_SIZE, stream);
    if (e != cudaSuccess) { cudaFreeAsync(g, stream); return set_cuda_error("ptt_sin_transform_batch alloc", e); }
    st_prep_kernel<<<P, 256, 0, stream>>>(xi, f, stride, off, npts, g, env);
    dim3 grid((n_z + ST_THREADS - 1) / ST_THREADS, P);
    sin_transform_kernel<<<grid, ST_THREADS, 0, stream>>>(n_z, z, frac, xi, stride, g, stride, 1, off, npts, env, ENV_SIZE,
                                                          out, n_z, 0, nullptr, nullptr, 0, nullptr, 0, nullptr, 0, 0);
    const int rc = check_launch("sin_transform_kernel");
    cudaFreeAsync(g, stream);
    cudaFreeAsync(env, stream);
    return rc;
}

extern "C" int ptt_sin_transform_unit_batch(int P, const double* xi, const double* f, int64_t stride, const int32_t* off,
                                            const int32_t* npts, int n_z, const double* z, const double* frac,
                                            double* out, void* stream_) {
    if (P == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (P < 0 || n_z < 0 || !xi || !f || !off || !npts || !z || !frac || !out)
        return set_error(PTT_E_ARG, "ptt_sin_transform_unit_batch: null pointer or bad size");
    for (int p0 = 0; P > MAX_BATCH_SLICE && p0 < P; p0 += MAX_BATCH_SLICE) {
        const int rc = ptt_sin_transform_unit_batch(min(MAX_BATCH_SLICE, P - p0), xi + (size_t)p0 * stride,
                                                    f + (size_t)p0 * stride, stride, off + p0, npts + p0, n_z, z, frac,
                                                    out + (size_t)p0 * n_z, stream_);
        if (rc != PTT_OK || p0 + MAX_BATCH_SLICE >= P) return rc;
    }
    if (P == 0 || n_z == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    double* g = nullptr;
    double* env = nullptr;
    double* mom = nullptr;
    keep_pool_memory();
    cudaError_t e = cudaMallocAsync((void**)&g, sizeof(double) * (size_t)P * (size_t)stride, stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&env, sizeof(double) * (size_t)P * ENV_SIZE, stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&mom, sizeof(double) * (size_t)P * MOM_SIZE, stream);
    if (e != cudaSuccess) {
        if (g) cudaFreeAsync(g, stream);
        if (env) cudaFreeAsync(env, stream);
        return set_cuda_error("ptt_sin_transform_unit_batch alloc", e);
    }
    st_prep_kernel<<<P, 256, 0, stream>>>(xi, f, stride, off, npts, g, env);
    st_moments_kernel<<<P, 256, 0, stream>>>(xi, g, stride, off, npts, mom);
    dim3 grid((n_z + STM_THREADS - 1) / STM_THREADS, P);
    st_transform_moments_kernel<<<grid, STM_THREADS, 0, stream>>>(n_z, z, frac, npts, mom, env, out);
    const int rc = check_launch("st_transform_moments_kernel");
    cudaFreeAsync(g, stream);
    cudaFreeAsync(env, stream);
    cudaFreeAsync(mom, stream);
    return rc;
}

extern "C" int ptt_a2_batch(const ptt_a2_plan_t* plan, int P, const double* rows_v, const double* rows_w,
                            const double* rows_xi, int64_t row_stride, const int32_t* off, const int32_t* npts,
                            const double* pp, double* work, int64_t work_stride, double* a2, double* fp2_2,
                            double* lam2, void* stream_) {
    return ptt::a2_batch_strided(plan, P, rows_v, rows_w, rows_xi, row_stride, off, npts, pp, work, work_stride, a2,
                                 plan ? plan->n_z : 0, fp2_2, lam2, stream_);
}

// ptt_a2_batch with a row stride of its own for the A^2 table (the sweep driver pads it to an even number of doubles:
// rows that start on 16-byte boundaries can be fetched by the tensor-map copies of K6').  fp2_2 / lam2: stride n_z.
int ptt::a2_batch_strided(const ptt_a2_plan_t* plan, int P, const double* rows_v, const double* rows_w,
                          const double* rows_xi, int64_t row_stride, const int32_t* off, const int32_t* npts,
                          const double* pp, double* work, int64_t work_stride, double* a2, int64_t a2_stride,
                          double* fp2_2, double* lam2, void* stream_) {
    if (P == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (!plan || P < 0 || !rows_v || !rows_w || !rows_xi || !off || !npts || !pp || !work || !a2 || a2_stride < plan->n_z)
        return set_error(PTT_E_ARG, "ptt_a2_batch: null pointer or bad size");
    for (int p0 = 0; P > MAX_BATCH_SLICE && p0 < P; p0 += MAX_BATCH_SLICE) {
        const size_t o = (size_t)p0;
        const int rc = a2_batch_strided(plan, min(MAX_BATCH_SLICE, P - p0), rows_v + o * row_stride, rows_w + o * row_stride,
                                        rows_xi + o * row_stride, row_stride, off + p0, npts + p0, pp + o * 8,
                                        work + o * work_stride, work_stride, a2 + o * a2_stride, a2_stride,
                                        fp2_2 ? fp2_2 + o * plan->n_z : nullptr, lam2 ? lam2 + o * plan->n_z : nullptr, stream_);
        if (rc != PTT_OK || p0 + MAX_BATCH_SLICE >= P) return rc;
    }
    if (P == 0) return PTT_OK;
    WorkLayout lay;
    lay.row_cap = (int)row_stride;
    lay.nxi = plan->nxi;
    lay.n_z = plan->n_z;
    if (work_stride < lay.total()) return set_error(PTT_E_ARG, "ptt_a2_batch: work_stride too small");
    cudaStream_t stream = (cudaStream_t)stream_;
    a2_prep_kernel<<<P, 256, 0, stream>>>(plan->nxi, plan->phase_rule, plan->xi_re, rows_v, rows_w, rows_xi, row_stride,
                                          off, npts, pp, work, work_stride, lay);
    if (plan->z_exact_max <= MOM_ZMAX && !plan->force_direct) {
        // profiles live on xi in [0, 1]: exact part by block moments
        a2_moments_kernel<<<P, 256, 0, stream>>>(plan->nxi, plan->xi_re, rows_xi, row_stride, off, npts, work, work_stride,
                                                 lay);
        dim3 grid((plan->n_z + 2 * STM_THREADS - 1) / (2 * STM_THREADS), P);
        sin_transform_moments_kernel<<<grid, STM_THREADS, 0, stream>>>(plan->n_z, plan->z, plan->frac, npts, work,
                                                                       work_stride, lay);
    } else {
        dim3 grid((plan->n_z + ST_THREADS - 1) / ST_THREADS, P);
        sin_transform_kernel<<<grid, ST_THREADS, 0, stream>>>(
            plan->n_z, plan->z, plan->frac, rows_xi, row_stride, work + lay.gv(), work_stride, 0, off, npts,
            work + lay.env_v(), work_stride,
            work + lay.F(), work_stride, plan->nxi, plan->xi_re, work + lay.hl(), work_stride, work + lay.env_l(),
            work_stride, work + lay.L(), work_stride, 1);
    }
    dim3 grid2((plan->n_z + 255) / 256, (P + A2F_PPT - 1) / A2F_PPT);
    a2_finish_kernel<<<grid2, 256, 0, stream>>>(P, plan->n_z, plan->n_seg, plan->seg, plan->z, work, work_stride, lay.F(),
                                                lay.L(), pp, a2, a2_stride, fp2_2, lam2);
    return check_launch("ptt_a2_batch");
}

extern "C" int ptt_a2_work_stride(int row_stride, int nxi, int n_z) {
    WorkLayout lay;
    lay.row_cap = row_stride;
    lay.nxi = nxi;
    lay.n_z = n_z;
    return (int)lay.total();
}

extern "C" int ptt_spec_den_v_batch(const ptt_sdv_plan_t* plan, int P, const double* a2, int64_t a2_stride,
                                    const double* v_wall, double* out, int64_t out_stride, void* stream_) {
    if (P == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (!plan || P < 0 || !a2 || !v_wall || !out || plan->z_tile < 1)
        return set_error(PTT_E_ARG, "ptt_spec_den_v_batch: null pointer or bad size");
    for (int p0 = 0; P > MAX_BATCH_SLICE && p0 < P; p0 += MAX_BATCH_SLICE) {
        const int rc = ptt_spec_den_v_batch(plan, min(MAX_BATCH_SLICE, P - p0), a2 + (size_t)p0 * a2_stride, a2_stride,
                                            v_wall + p0, out + (size_t)p0 * out_stride, out_stride, stream_);
        if (rc != PTT_OK || p0 + MAX_BATCH_SLICE >= P) return rc;
    }
    if (P == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int smem = plan->smem_bytes;
    {   // static + dynamic shared memory may exceed the 48 KB default: always opt in
        const cudaError_t e = cudaFuncSetAttribute(spec_den_v_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return set_cuda_error("spec_den_v_kernel smem", e);
    }
    dim3 grid((plan->n_z + plan->z_tile - 1) / plan->z_tile, P);
    spec_den_v_kernel<<<grid, SDV_THREADS, smem, stream>>>(plan->n_q, plan->n_z, plan->nt, plan->z_tile, plan->qT, plan->z,
                                                           plan->zterm, plan->T, plan->LT, plan->W, a2, a2_stride, v_wall,
                                                           out, out_stride, plan->inv_dlog, smem / (int)sizeof(double2));
    return check_launch("spec_den_v_kernel");
}

extern "C" int ptt_spec_den_gw_batch(const ptt_gw_plan_t* plan, int P, const double* pv, int64_t pv_stride,
                                     const double* slf, const double* scale, double* sd_gw, double* pow_gw,
                                     double* omgw0, void* stream_) {
    if (P == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (!plan || P < 0 || !pv || !slf || plan->y_tile < 1 || plan->y_tile % GW_R != 0)
        return set_error(PTT_E_ARG, "ptt_spec_den_gw_batch: null pointer or bad size (y_tile must be a multiple of 4)");
    for (int p0 = 0; P > MAX_BATCH_SLICE && p0 < P; p0 += MAX_BATCH_SLICE) {
        const size_t o = (size_t)p0 * plan->n_y;
        const int rc = ptt_spec_den_gw_batch(plan, min(MAX_BATCH_SLICE, P - p0), pv + (size_t)p0 * pv_stride, pv_stride,
                                             slf + p0, scale ? scale + p0 : nullptr, sd_gw ? sd_gw + o : nullptr,
                                             pow_gw ? pow_gw + o : nullptr, omgw0 ? omgw0 + o : nullptr, stream_);
        if (rc != PTT_OK || p0 + MAX_BATCH_SLICE >= P) return rc;
    }
    if (P == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int smem = plan->smem_bytes;
    {
        const cudaError_t e = cudaFuncSetAttribute(spec_den_gw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return set_cuda_error("spec_den_gw_kernel smem", e);
    }
    dim3 grid((plan->n_y + plan->y_tile - 1) / plan->y_tile, P);
    spec_den_gw_kernel<<<grid, GW_THREADS, smem, stream>>>(plan->n_zl, plan->n_y, plan->n_int, plan->y_tile, plan->z_lookup,
                                                           plan->y, plan->yterm, plan->L1, plan->L2, plan->Z1, plan->Z2,
                                                           plan->G, plan->prefactor, pv, pv_stride, slf, scale, sd_gw,
                                                           pow_gw, omgw0, smem / (int)sizeof(double2));
    return check_launch("spec_den_gw_kernel");
}

// spec_den_gw_scaled on an ARBITRARY ascending z_lookup (and any order of y): np.interp by binary search per sample
// (spectrum.py:327-368, 361-363).  The log-uniform grids the reference itself always passes take the arithmetic-index
// kernels above; this is the general form of the operator.  CTA = one (wavenumber, profile) pair.
__device__ __forceinline__ double interp_search(const double x, const int n, const double* __restrict__ xp,
                                                const double* __restrict__ fp) {
    if (!(x > xp[0])) return (x != x) ? x : fp[0];           // np.interp: left clamp (NaN stays NaN)
    if (x >= xp[n - 1]) return fp[n - 1];
    int lo = 0, hi = n - 1;                                   // xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid; else hi = mid;
    }
    const double x0 = xp[lo], f0 = fp[lo];
    return fma((fp[lo + 1] - f0) / (xp[lo + 1] - x0), x - x0, f0);
}

__global__ void __launch_bounds__(128) spec_den_gw_search_kernel(
    const int n_zl, const int n_y, const int n_int, const double* __restrict__ z_lookup, const double* __restrict__ y,
    const double* __restrict__ Z1, const double* __restrict__ Z2, const double* __restrict__ G, const double prefactor,
    const double* __restrict__ pv, const long long pv_stride, const double* __restrict__ slf,
    const double* __restrict__ scale, double* __restrict__ sd_gw, double* __restrict__ pow_gw, double* __restrict__ omgw0) {
    __shared__ double s_red[128];
    const int i = blockIdx.x;
    const size_t p = blockIdx.y;
    const double* fp = pv + p * pv_stride;
    const double yi = y[i];
    double acc = 0.0;
    for (int k = threadIdx.x; k < n_int; k += blockDim.x) {
        const double p1 = interp_search(yi * Z1[k], n_zl, z_lookup, fp);
        const double p2 = interp_search(yi * Z2[k], n_zl, z_lookup, fp);
        acc = fma(G[k] * p1, p2, acc);
    }
    s_red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double sd = prefactor * yi * yi * s_red[0] * slf[p];
        const size_t o = p * n_y + i;
        if (sd_gw) sd_gw[o] = sd;
        const double pw = yi * yi * yi / (2.0 * M_PI * M_PI) * sd;
        if (pow_gw) pow_gw[o] = pw;
        if (omgw0 && scale) omgw0[o] = scale[p] * pw;
    }
}

extern "C" int ptt_spec_den_gw_general_batch(int n_zl, const double* z_lookup, int n_y, const double* y, int n_int,
                                             const double* Z1, const double* Z2, const double* G, double prefactor, int P,
                                             const double* pv, int64_t pv_stride, const double* slf, const double* scale,
                                             double* sd_gw, double* pow_gw, double* omgw0, void* stream_) {
    if (P == 0 || n_y == 0) return PTT_OK;
    if (P < 0 || n_y < 0 || n_zl < 2 || n_int < 2 || !z_lookup || !y || !Z1 || !Z2 || !G || !pv || !slf || pv_stride < n_zl)
        return set_error(PTT_E_ARG, "ptt_spec_den_gw_general_batch: null pointer or bad size");
    for (int p0 = 0; p0 < P; p0 += MAX_BATCH_SLICE) {
        const int np_ = min(MAX_BATCH_SLICE, P - p0);
        const size_t o = (size_t)p0 * n_y;
        spec_den_gw_search_kernel<<<dim3(n_y, np_), 128, 0, (cudaStream_t)stream_>>>(
            n_zl, n_y, n_int, z_lookup, y, Z1, Z2, G, prefactor, pv + (size_t)p0 * pv_stride, pv_stride, slf + p0,
            scale ? scale + p0 : nullptr, sd_gw ? sd_gw + o : nullptr, pow_gw ? pow_gw + o : nullptr,
            omgw0 ? omgw0 + o : nullptr);
    }
    return check_launch("spec_den_gw_search_kernel");
}

extern "C" int ptt_pow_spec_batch(int P, int n_z, const double* z, const double* spec_den, int64_t stride, double* out,
                                  void* stream_) {
    if (P == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (P < 0 || n_z < 0 || !z || !spec_den || !out) return set_error(PTT_E_ARG, "ptt_pow_spec_batch: null pointer");
    for (int p0 = 0; P > MAX_BATCH_SLICE && p0 < P; p0 += MAX_BATCH_SLICE) {
        const int rc = ptt_pow_spec_batch(min(MAX_BATCH_SLICE, P - p0), n_z, z, spec_den + (size_t)p0 * stride, stride,
                                          out + (size_t)p0 * stride, stream_);
        if (rc != PTT_OK || p0 + MAX_BATCH_SLICE >= P) return rc;
    }
    if (P == 0 || n_z == 0) return PTT_OK;
    dim3 grid((n_z + 255) / 256, P);
    pow_spec_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(n_z, z, spec_den, stride, out);
    return check_launch("pow_spec_kernel");
}

// mode 0: DFMA only, 1: DMMA (m8n8k4) only, 2: both interleaved (are the FP64 FMA pipe and the FP64 tensor pipe the
// same units?).  Flops counted: 2 per FMA lane-op, 512 per warp-level DMMA.
__global__ void fp64_pipe_probe_kernel(const int mode, const int iters, double* sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double c[8][2];
    for (int q = 0; q < 8; ++q) { c[q][0] = 0.0; c[q][1] = 0.0; }
    const double b = 1.0000001, cc = 1e-9, fa = 1e-3 * (threadIdx.x & 7), fb = 1e-3;
    for (int i = 0; i < iters; ++i) {
        if (mode != 1) {
            a0 = fma(a0, b, cc); a1 = fma(a1, b, cc); a2 = fma(a2, b, cc); a3 = fma(a3, b, cc);
            a4 = fma(a4, b, cc); a5 = fma(a5, b, cc); a6 = fma(a6, b, cc); a7 = fma(a7, b, cc);
        }
        if (mode != 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(fa), "d"(fb));
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1];
    if (s == 123.456) sink[0] = s;
}

extern "C" int ptt_fp64_pipe_probe(int mode, int iters, double* out_flops_per_s, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (mode < 0 || mode > 2 || iters < 1 || !out_flops_per_s) return set_error(PTT_E_ARG, "ptt_fp64_pipe_probe: bad argument");
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* sink = nullptr;
    {
        const cudaError_t e = cudaMalloc((void**)&sink, sizeof(double));
        if (e != cudaSuccess) return set_cuda_error("fp64 probe: cudaMalloc", e);
    }
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int blocks = sms * 8, threads = 256;
    fp64_pipe_probe_kernel<<<blocks, threads, 0, stream>>>(mode, iters / 10 + 1, sink);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a, stream);
        fp64_pipe_probe_kernel<<<blocks, threads, 0, stream>>>(mode, iters, sink);
        cudaEventRecord(b, stream);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double per_thread = (mode != 1 ? 2.0 * 8.0 : 0.0) + (mode != 0 ? 8.0 * 512.0 / 32.0 : 0.0);
        const double flops = per_thread * (double)iters * (double)blocks * (double)threads / (ms * 1e-3);
        if (flops > best) best = flops;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(sink);
    *out_flops_per_s = best;
    return check_launch("fp64_pipe_probe_kernel");
}

extern "C" int ptt_fp64_peak_probe(int iters, double* out_flops_per_s, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* sink = nullptr;
    {
        const cudaError_t e = cudaMalloc((void**)&sink, sizeof(double));
        if (e != cudaSuccess) return set_cuda_error("fp64 probe: cudaMalloc", e);
    }
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int blocks = sms * 8, threads = 256;
    fp64_probe_kernel<<<blocks, threads, 0, stream>>>(iters / 10 + 1, sink);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a, stream);
        fp64_probe_kernel<<<blocks, threads, 0, stream>>>(iters, sink);
        cudaEventRecord(b, stream);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads / (ms * 1e-3);
        if (flops > best) best = flops;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(sink);
    *out_flops_per_s = best;
    return check_launch("fp64_probe_kernel");
}
