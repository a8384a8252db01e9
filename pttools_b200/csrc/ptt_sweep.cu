// Seam B, the performance seam: a whole (v_wall, alpha_n, model) sweep behind ONE C call.
//   create_spectra / create_bubbles -> run_parallel(one future per point)   pttools/analysis/parallel.py:89-187,
//                                                                            pttools/speedup/parallel.py:67-196
//   per point: Bubble(model, v_wall, alpha_n).solve()                        pttools/bubble/bubble.py:232-352
//              Spectrum(bubble, r_star=...)                                  pttools/ssmtools/spectrum.py:96-136,
//                                                                            pttools/omgw0/omgw0_ssm.py:123-131
// Host C++ only drives: ordering by v_wall, chunking against device memory, the K6' groups on side streams, the retry
// search folded in under the spectrum stages, results streamed out per chunk.  All arithmetic is in the kernels.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <numeric>
#include <vector>

#include "ptt_host.cuh"

namespace ptt {

constexpr int SW_PG = 16;        // columns of pg (ptt_sweep_shells_generic)
constexpr int SW_PM = 4;         // columns of pm
constexpr int SW_GS = 16;        // scalars of the shell solver
constexpr int SW_NS = PTT_SWEEP_NSCAL;
constexpr int SW_GROUP_MIN = 24;     // default smallest run of equal v_wall worth the banded matrix product
constexpr int SW_GW_CELLS_MIN = 128; // smallest batch for the cell formulation of the GW integral
constexpr int SW_GM_BM = 128;        // row tile of K6' (ptt_sdv_gemm.cu)

// ---------------------------------------------------------------------------------------------------------------
// small kernels: parameter rows, thermodynamic algebra, result movement
// ---------------------------------------------------------------------------------------------------------------

// pp[n,8]  = {v_wall, cs_b, e_coef_s, e_coef_b, V_s, V_b, 0, 0}                    (ptt_a2_batch)
// pp2[n,12] = {v_wall, mu_s, mu_b, V_s, V_b, a_s, a_b, T_ref, phase_rule = 1, 0, 0, 0}  (ptt_profile_integrals_batch)
__global__ void sweep_prep_kernel(const int n, const double* __restrict__ pg, const double* __restrict__ pm,
                                  double* __restrict__ pp, double* __restrict__ pp2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* q = pg + (size_t)i * SW_PG;
    const double* m = pm + (size_t)i * SW_PM;
    const double mu_s = 1.0 + 1.0 / q[9], mu_b = 1.0 + 1.0 / q[10];
    double* a = pp + (size_t)i * 8;
    a[0] = q[0]; a[1] = sqrt(q[10]); a[2] = 1.0 - 1.0 / mu_s; a[3] = 1.0 - 1.0 / mu_b; a[4] = q[11]; a[5] = q[12];
    a[6] = 0.0; a[7] = 0.0;
    double* b = pp2 + (size_t)i * 12;
    b[0] = q[0]; b[1] = mu_s; b[2] = mu_b; b[3] = q[11]; b[4] = q[12]; b[5] = m[0]; b[6] = m[1]; b[7] = m[2];
    b[8] = 1.0; b[9] = 0.0; b[10] = 0.0; b[11] = 0.0;
}

// The volume-averaged quantities of pttools/bubble/thermo.py from the K4 integrals, the source lifetime factor of
// SSMSpectrum (spectrum.py:123-136) and the row of the result table.  One thread per entry; entry f reads point
// row[f] of pg and writes row row[f] of the outputs (row == NULL: f); entries with mask[f] == 0 are skipped.
__global__ void thermo_finish_kernel(const int n, const double* __restrict__ pg, const int* __restrict__ row_of,
                                     const int* __restrict__ mask, const double* __restrict__ gscal,
                                     const double* __restrict__ I, const int* __restrict__ npts, const double r_star,
                                     const double lifetime_mult, double* __restrict__ table, double* __restrict__ slf) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    if (mask && !mask[f]) return;
    const int row = row_of ? row_of[f] : f;
    const double* q = pg + (size_t)row * SW_PG;
    const double* g = gscal + (size_t)f * SW_GS;
    const double* in = I + (size_t)f * 8;
    double* t = table + (size_t)row * SW_NS;
    for (int k = 0; k < 14; ++k) t[k] = g[k];
    const double v_wall = q[0], wn = q[3];
    const double mu_s = 1.0 + 1.0 / q[9], mu_b = 1.0 + 1.0 / q[10], V_s = q[11], V_b = q[12];
    const double four_pi_3 = 4.0 * M_PI / 3.0;
    const double vw3 = v_wall * v_wall * v_wall;
    const double va_kin = four_pi_3 * in[0];                       // thermo.py:169-181
    const double va_trace = four_pi_3 * in[1];                     // :213-221
    const double va_thd = four_pi_3 * 0.75 * in[2];                // :206-210
    const double va_s = four_pi_3 * in[3];                         // :156-166
    const double wbar = in[4];                                     // :139-149
    const double ebar = (1.0 - 1.0 / mu_s) * wn + V_s;             // :89-91  e(wn, symmetric)
    const double ked = 3.0 / (4.0 * M_PI * vw3) * va_kin;          // :42-46
    const double kef = ked / ebar;                                 // :49-53
    const double v_sh = g[4];
    const double va_th = M_PI * wn * v_sh * v_sh * v_sh - va_kin - va_trace;   // :191-195
    const double thd = 3.0 / (4.0 * M_PI * vw3) * va_th;           // :56-60
    const double va_enth = 4.0 / 3.0 * thd;                        // :153-155
    const double kappa = va_kin / fabs(va_trace);                  // :94-106
    const double omega = va_thd / fabs(va_trace);                  // :122-133
    const double ubarf2 = ked / in[7];                             // :136 (w[-1])
    // nu of Giombi et al. at the volume-averaged enthalpy, broken phase (models/model.py:688-698)
    const double om = (va_enth / mu_b - V_b) / ((1.0 - 1.0 / mu_b) * va_enth + V_b);
    const double nu = (1.0 - 3.0 * om) / (1.0 + 3.0 * om);
    double f_life = 1.0 / (1.0 + 2.0 * nu);                        // spectrum.py:123-136
    if (r_star > 0.0 && !isinf(lifetime_mult))
        f_life *= 1.0 - pow(1.0 + lifetime_mult * 2.0 * r_star / sqrt(kef), -1.0 - 2.0 * nu);
    t[14] = va_kin; t[15] = va_trace; t[16] = va_thd; t[17] = va_s; t[18] = wbar; t[19] = ebar; t[20] = ked;
    t[21] = kef; t[22] = va_th; t[23] = thd; t[24] = va_enth; t[25] = kappa; t[26] = omega; t[27] = ubarf2;
    t[28] = wbar / ebar; t[29] = nu; t[30] = f_life; t[31] = (double)npts[f];
    slf[row] = f_life;
}

// rows src[f, 0:n_col] (stride src_stride) -> dst[row[f], 0:n_col] (stride dst_stride); row == NULL: row f.
template <class T>
__global__ void scatter_rows_kernel(const int n_col, const T* __restrict__ src, const long long src_stride,
                                    const long long* __restrict__ row, T* __restrict__ dst, const long long dst_stride) {
    const size_t f = blockIdx.x;                 // rows on grid.x (no 65535 limit), column blocks on grid.y
    const T* s = src + f * src_stride;
    T* d = dst + (size_t)(row ? row[f] : (long long)f) * dst_stride;
    for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < n_col; c += gridDim.y * blockDim.x) d[c] = s[c];
}

__global__ void fill_rows_kernel(const int n_col, double* __restrict__ dst, const long long stride, const double value) {
    double* d = dst + (size_t)blockIdx.x * stride;
    for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < n_col; c += gridDim.y * blockDim.x) d[c] = value;
}

// dst[r, :] = src[idx[r], :]
template <class I>
__global__ void gather_rows_kernel(const long long n, const int n_col, const double* __restrict__ src,
                                   const I* __restrict__ idx, double* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n_col) return;
    const long long r = i / n_col;
    const int c = (int)(i - r * n_col);
    dst[i] = src[(long long)idx[r] * n_col + c];
}

// Writes the rescued points of the retry search (compact entry f, rescued[f] != 0) over their rows idx[f] of the shell
// chunk: profile rows, off, npts, solver scalars, status.
__global__ void rescue_writeback_kernel(const int* __restrict__ idx, const int* __restrict__ rescued, const long long cap,
                                        const double* __restrict__ c_v, const double* __restrict__ c_w,
                                        const double* __restrict__ c_x, const int* __restrict__ c_off,
                                        const int* __restrict__ c_np, const double* __restrict__ c_sc,
                                        const int* __restrict__ c_st, double* __restrict__ r_v, double* __restrict__ r_w,
                                        double* __restrict__ r_x, int* __restrict__ off, int* __restrict__ npts,
                                        double* __restrict__ sc, int* __restrict__ st) {
    const size_t f = blockIdx.x;
    if (!rescued[f]) return;
    const size_t d = (size_t)idx[f];
    const double* src = (blockIdx.y == 0 ? c_v : blockIdx.y == 1 ? c_w : c_x) + f * cap;
    double* dst = (blockIdx.y == 0 ? r_v : blockIdx.y == 1 ? r_w : r_x) + d * cap;
    for (long long c = threadIdx.x; c < cap; c += blockDim.x) dst[c] = src[c];
    if (blockIdx.y == 0) {
        if (threadIdx.x < SW_GS) sc[d * SW_GS + threadIdx.x] = c_sc[f * SW_GS + threadIdx.x];
        if (threadIdx.x == 32) { off[d] = c_off[f]; npts[d] = c_np[f]; st[d] = c_st[f]; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// workspace: grow-only device / pinned buffers and the streams, kept between calls
// ---------------------------------------------------------------------------------------------------------------

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <class T>
    T* get(const size_t n) {
        const size_t b = (n ? n : 1) * sizeof(T);
        if (b > cap) {
            if (p) cudaFree(p);
            p = nullptr; cap = 0;
            const size_t want = b + b / 32 + 4096;
            if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; cudaGetLastError(); return nullptr; }
            cap = want;
        }
        return (T*)p;
    }
    template <class T> T* as() const { return (T*)p; }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <class T>
    T* get(const size_t n) {
        const size_t b = (n ? n : 1) * sizeof(T);
        if (b > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr; cap = 0;
            const size_t want = b + b / 32 + 4096;
            if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { p = nullptr; cudaGetLastError(); return nullptr; }
            cap = want;
        }
        return (T*)p;
    }
    template <class T> T* as() const { return (T*)p; }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct ShellSlot {
    // the shell chunk: profile rows and per-point tables
    DevBuf rows[3], off, npts, gscal, gstatus, I, pp, pp2, slf, table;
    // compact results of the retry search
    DevBuf c_rows[3], c_off, c_npts, c_scal, c_status, c_rescued, c_idx, c_I, c_pp2;
    PinBuf h_status, h_npts, h_rescued, h_idx;
    cudaEvent_t solved = nullptr, consumed = nullptr, retried = nullptr;
    std::vector<DevBuf*> dev() {
        return {&rows[0], &rows[1], &rows[2], &off, &npts, &gscal, &gstatus, &I, &pp, &pp2, &slf, &table, &c_rows[0],
                &c_rows[1], &c_rows[2], &c_off, &c_npts, &c_scal, &c_status, &c_rescued, &c_idx, &c_I, &c_pp2};
    }
    std::vector<PinBuf*> pin() { return {&h_status, &h_npts, &h_rescued, &h_idx}; }
};

struct TimingRec { int stage; cudaEvent_t a, b; };

struct Workspace {
    int device = -1;
    std::mutex mu;
    cudaStream_t s_shell = nullptr, s_retry = nullptr, s_copy = nullptr, s_side[2] = {nullptr, nullptr};
    ShellSlot slot[2];
    DevBuf pg, pm, scale, perm, stage_in, a2, work, pv_y, pv_l, omega[2], ebase, gw_table, out[2][3], full[3], full_scal,
        full_status, vw_tmp, snr, snr_perm, trows;
    PinBuf h_pg, h_ebase, h_scal, h_status;
    cudaEvent_t out_free[2] = {nullptr, nullptr}, ev_fork = nullptr, ev_join[2] = {nullptr, nullptr}, ev_tmp = nullptr,
                ev_emit = nullptr;
    // Stage events come in two sets used by alternate sweeps: a device-table sweep returns with its tail in flight and
    // its events unread, and the next sweep - queued behind it without waiting - needs events of its own.
    std::vector<cudaEvent_t> timing[2];
    size_t timing_used = 0;
    int tset = 0;                                // the set of the sweep that runs (or ran last)
    struct Pending { std::vector<TimingRec> recs; ptt_sweep_stats_t stats; int64_t serial = 0; };
    Pending pending[2];                          // per set: events not read yet, with the stats they complete
    std::vector<ptt_sweep_stats_t> done;         // completed stats of earlier timed sweeps (serial inside), newest last
    int64_t serial = 0;
    PinBuf h_trows;

    std::vector<cudaStream_t*> streams() { return {&s_shell, &s_retry, &s_copy, &s_side[0], &s_side[1]}; }
    std::vector<cudaEvent_t*> events() {
        return {&out_free[0], &out_free[1], &ev_fork, &ev_join[0], &ev_join[1], &ev_tmp, &ev_emit, &slot[0].solved,
                &slot[0].consumed, &slot[0].retried, &slot[1].solved, &slot[1].consumed, &slot[1].retried};
    }
    std::vector<DevBuf*> dev() {
        return {&pg, &pm, &scale, &perm, &stage_in, &a2, &work, &pv_y, &pv_l, &omega[0], &omega[1], &ebase, &gw_table,
                &out[0][0], &out[0][1], &out[0][2], &out[1][0], &out[1][1], &out[1][2], &full[0], &full[1], &full[2],
                &full_scal, &full_status, &vw_tmp, &snr, &snr_perm, &trows};
    }
    int init() {
        int dev_now = 0;
        cudaError_t e = cudaGetDevice(&dev_now);
        if (e != cudaSuccess) return set_cuda_error("cudaGetDevice", e);
        if (device == dev_now) return PTT_OK;
        if (device >= 0) release();                 // the process moved to another device: start over there
        for (cudaStream_t* s : streams())
            if ((e = cudaStreamCreateWithFlags(s, cudaStreamNonBlocking)) != cudaSuccess)
                return set_cuda_error("cudaStreamCreate", e);
        for (cudaEvent_t* x : events())
            if ((e = cudaEventCreateWithFlags(x, cudaEventDisableTiming)) != cudaSuccess)
                return set_cuda_error("cudaEventCreate", e);
        device = dev_now;
        return PTT_OK;
    }
    cudaEvent_t timing_event() {
        std::vector<cudaEvent_t>& t = timing[tset];
        if (timing_used == t.size()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            t.push_back(e);
        }
        return t[timing_used++];
    }
    void release() {
        if (device < 0) return;
        cudaDeviceSynchronize();
        for (ShellSlot& s : slot) {
            for (DevBuf* b : s.dev()) b->release();
            for (PinBuf* b : s.pin()) b->release();
        }
        for (DevBuf* b : dev()) b->release();
        h_pg.release(); h_ebase.release(); h_scal.release(); h_status.release();
        for (cudaStream_t* s : streams()) { cudaStreamDestroy(*s); *s = nullptr; }
        for (cudaEvent_t* e : events()) { cudaEventDestroy(*e); *e = nullptr; }
        for (int k = 0; k < 2; ++k) {
            for (cudaEvent_t e : timing[k]) cudaEventDestroy(e);
            timing[k].clear();
            pending[k].recs.clear();
        }
        h_trows.release();
        timing_used = 0;
        device = -1;
    }
};

static Workspace g_ws;
static ptt_sweep_stats_t g_stats;

#define SW_CUDA(call)                                            \
    do {                                                         \
        const cudaError_t e_ = (call);                           \
        if (e_ != cudaSuccess) return set_cuda_error(#call, e_); \
    } while (0)
#define SW_RC(call)                    \
    do {                               \
        const int rc_ = (call);        \
        if (rc_ != PTT_OK) return rc_; \
    } while (0)
#define SW_ALLOC(ptr)                                                                                  \
    do {                                                                                               \
        if (!(ptr)) return set_error(PTT_E_NOMEM, "ptt_sweep_spectra: out of device or pinned memory"); \
    } while (0)
#define SW_LAUNCHED(name)                                          \
    do {                                                           \
        const cudaError_t e_ = cudaGetLastError();                 \
        if (e_ != cudaSuccess) return set_cuda_error(name, e_);    \
    } while (0)

// First knot of every 128-row tile and the common band width of K6' for one wall speed (host side of
// ptt_spec_den_v_group).
// Every tile's first knot is given the parity of the pass' first column in the A^2 table (one knot earlier where needed;
// -1 = a knot before the table, weight 0): with the even row stride of the table its entries then start on a 16-byte
// boundary, which the tensor-map copies of K6' need (PTT_SDV_ALIGNED_BAND).
static int group_band(const ptt_sweep_plan_t& pl, const int ip, const double v_wall, int32_t* e_base) {
    const ptt_sdv_plan_t& sp = pl.sdv[ip];
    const double* zterm = pl.h_zterm[ip];
    const double b_r = cbrt(8.0 * M_PI);
    const double shift = -log10(b_r * v_wall) / pl.dl[ip];
    const int n_z = sp.n_z, n_q = sp.n_q;
    const int n_tiles = (n_z + SW_GM_BM - 1) / SW_GM_BM;
    long long kw = 0;
    const long long hi = std::max(n_q - 2, 0);
    for (int t = 0; t < n_tiles; ++t) {
        const int i0 = t * SW_GM_BM, i1 = std::min(i0 + SW_GM_BM, n_z) - 1;
        const long long e_first = (long long)floor(zterm[i0] + pl.lt_first[ip] + shift) - 1;
        long long eb = std::min(std::max(e_first, 0LL), hi);
        eb -= (eb + pl.seg[ip]) & 1;
        long long e_last = (long long)floor(zterm[i1] + pl.lt_last[ip] + shift) + 1;
        e_last = std::min(std::max(e_last, 0LL), hi) + 1;
        e_base[t] = (int32_t)eb;
        kw = std::max(kw, e_last - eb);
    }
    kw += 3;
    kw = (kw + 15) / 16 * 16;
    return (int)kw;
}

struct Timer {
    Workspace& ws;
    bool on;
    typedef TimingRec Rec;
    std::vector<Rec> recs;
    Timer(Workspace& w, const bool enabled) : ws(w), on(enabled) {}
    cudaEvent_t begin(cudaStream_t s) {
        if (!on) return nullptr;
        cudaEvent_t a = ws.timing_event();
        if (a) cudaEventRecord(a, s);
        return a;
    }
    void end(const int stage, cudaEvent_t a, cudaStream_t s) {
        if (!on || !a) return;
        cudaEvent_t b = ws.timing_event();
        if (!b) return;
        cudaEventRecord(b, s);
        recs.push_back({stage, a, b});
    }
    void collect(ptt_sweep_stats_t& st) { collect_recs(recs, st); }
    static void collect_recs(const std::vector<Rec>& recs, ptt_sweep_stats_t& st) {
        const char* trace = getenv("PTT_SWEEP_TRACE");       // diagnostics: timeline of the stages on stderr
        for (const Rec& r : recs) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
                st.stage_ms[r.stage] += ms;
                st.stage_calls[r.stage] += 1;
            }
            if (trace && !recs.empty()) {
                float t0 = 0.f;
                cudaEventElapsedTime(&t0, recs[0].a, r.a);
                fprintf(stderr, "[ptt trace] stage %d start %9.3f ms dur %8.3f ms\n", r.stage, t0, ms);
            }
        }
        if (trace) fprintf(stderr, "[ptt trace] end\n");
    }
};

// One sweep call.  Rows are processed in v_wall order ("sorted positions"); perm maps a sorted position to the
// caller's row.
struct Sweep {
    Workspace& ws;
    const ptt_sweep_plan_t* plan;
    const ptt_sweep_opts_t& opts;
    const ptt_sweep_out_t& out;
    ptt_chunk_cb cb;
    void* user;
    const bool host;
    cudaStream_t s_main;
    Timer timer;
    ptt_sweep_stats_t stats;

    long long n = 0, shell_chunk = 0, n_shell_chunks = 0, C = 0, cap = 0;
    int n_xi = 5000, thin = 100, chunk = 8000, group_min = SW_GROUP_MIN, n_pass = 0, n_y = 0, n_zl = 0, n_a2 = 0, ws_stride = 0;
    long long a2_ld = 0;             // row stride of the A^2 table: n_a2 padded to an even number (16-byte aligned rows for K6')
    double t_end = 50.0, wn_rtol = 1e-4;
    bool sorted = true, want_spec = false, want_rows = false, use_cells = false, retry = true;
    double* h_pg = nullptr;
    std::vector<long long> perm;
    const double *d_pg = nullptr, *d_pm = nullptr, *d_scale = nullptr;
    long long* d_perm = nullptr;
    std::vector<long long> run_start;
    int n_tiles[2] = {0, 0};
    std::vector<int> run_kw[2];
    size_t eb_off[2] = {0, 0};
    int32_t* d_ebase = nullptr;
    double *d_a2 = nullptr, *d_work = nullptr, *d_pvy = nullptr, *d_pvl = nullptr, *d_gwtab = nullptr;
    double* o_tab[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    double* d_full[3] = {nullptr, nullptr, nullptr};
    double* d_full_scal = nullptr;
    int* d_full_status = nullptr;
    double* d_snr = nullptr;     // [n, n_snr] in processing order
    long long* d_trows = nullptr;   // destination row of every sorted position in the spectrum tables (out.table_rows)
    long long emitted = 0;      // spectrum chunks emitted so far (selects the output slot)
    bool out_used[2] = {false, false};

    Sweep(Workspace& w, const ptt_sweep_plan_t* p, const ptt_sweep_opts_t& o, const ptt_sweep_out_t& ot, ptt_chunk_cb c,
          void* u, const int mem, cudaStream_t s)
        : ws(w), plan(p), opts(o), out(ot), cb(c), user(u), host(mem == PTT_MEM_HOST), s_main(s),
          timer(w, (o.flags & PTT_SWEEP_TIMING) != 0) {
        memset(&stats, 0, sizeof(stats));
    }

    double hq(const long long i, const int c) const { return h_pg[(sorted ? i : perm[i]) * SW_PG + c]; }

    int setup(const long long n_, const double* pg, const double* pm, const double* scale);
    int solve(const long long c);
    int launch_retry(const long long c, const std::vector<int>& failed);
    int thermo(const long long c);
    int fold(const long long c, const std::vector<int>& failed, std::vector<char>& valid);
    int spectra_chunk(const long long c, const long long s, const long long e, const std::vector<char>& valid);
    int sdv_pass(const int ip, const long long c, const long long s, const long long e, const std::vector<char>& valid,
                 double* out_pv);
    int finish_shell_chunk(const long long c);
    int run();
};

int Sweep::setup(const long long n_, const double* pg, const double* pm, const double* scale) {
    n = n_;
    want_spec = out.pow_v || out.pow_gw || out.omgw0;
    want_rows = out.rows_v != nullptr;
    retry = !(opts.flags & PTT_SWEEP_NO_RETRY);
    n_xi = opts.n_xi > 0 ? opts.n_xi : 5000;
    cap = ptt_generic_row_capacity(n_xi);
    t_end = opts.t_end > 0 ? opts.t_end : 50.0;
    thin = opts.thin_shell_limit > 0 ? opts.thin_shell_limit : 100;
    wn_rtol = opts.wn_rtol > 0 ? opts.wn_rtol : 1e-4;
    chunk = opts.chunk > 0 ? opts.chunk : 8000;
    group_min = opts.group_min > 0 ? opts.group_min : SW_GROUP_MIN;
    const size_t nn = (size_t)n;

    // ---- inputs on the host: grouping needs v_wall, the retry filter the solution type
    h_pg = ws.h_pg.get<double>(nn * SW_PG); SW_ALLOC(h_pg);
    if (host) memcpy(h_pg, pg, nn * SW_PG * sizeof(double));
    else if (opts.host_v_wall && opts.host_sol_type) {
        // the two columns the host looks at (hq(i, 0), hq(i, 2)) come from the caller: nothing to wait for
        for (size_t i = 0; i < nn; ++i) {
            h_pg[i * SW_PG] = opts.host_v_wall[i];
            h_pg[i * SW_PG + 2] = (double)opts.host_sol_type[i];
        }
    } else {
        SW_CUDA(cudaMemcpyAsync(h_pg, pg, nn * SW_PG * sizeof(double), cudaMemcpyDeviceToHost, s_main));
        SW_CUDA(cudaStreamSynchronize(s_main));
    }
    sorted = true;
    for (size_t i = 1; i < nn && sorted; ++i) sorted = h_pg[i * SW_PG] >= h_pg[(i - 1) * SW_PG];
    if (!sorted) {
        perm.resize(nn);
        std::iota(perm.begin(), perm.end(), 0LL);
        std::stable_sort(perm.begin(), perm.end(), [&](long long a, long long b) { return h_pg[a * SW_PG] < h_pg[b * SW_PG]; });
        d_perm = ws.perm.get<long long>(nn); SW_ALLOC(d_perm);
        SW_CUDA(cudaMemcpyAsync(d_perm, perm.data(), nn * sizeof(long long), cudaMemcpyHostToDevice, s_main));
    }
    // ---- device copies in processing order
    {
        const double* src[3] = {pg, pm, scale};
        const int ncol[3] = {SW_PG, SW_PM, 1};
        DevBuf* buf[3] = {&ws.pg, &ws.pm, &ws.scale};
        const double** dst[3] = {&d_pg, &d_pm, &d_scale};
        for (int k = 0; k < 3; ++k) {
            if (!src[k]) continue;
            const size_t cnt = nn * ncol[k];
            if (!host && sorted) { *dst[k] = src[k]; continue; }
            double* d = buf[k]->get<double>(cnt); SW_ALLOC(d);
            if (sorted) SW_CUDA(cudaMemcpyAsync(d, src[k], cnt * sizeof(double), cudaMemcpyHostToDevice, s_main));
            else {
                const double* from = src[k];
                if (host) {
                    double* t = ws.stage_in.get<double>(cnt); SW_ALLOC(t);
                    SW_CUDA(cudaMemcpyAsync(t, src[k], cnt * sizeof(double), cudaMemcpyHostToDevice, s_main));
                    from = t;
                }
                gather_rows_kernel<long long><<<(unsigned)((cnt + 255) / 256), 256, 0, s_main>>>((long long)nn, ncol[k], from, d_perm, d);
                SW_LAUNCHED("gather_rows_kernel");
            }
            *dst[k] = d;
        }
    }
    // ---- runs of equal v_wall over the whole sweep and the bands of K6' (both passes), uploaded once
    n_pass = want_spec ? plan->n_pass : 0;
    run_start.clear();
    run_start.push_back(0);
    for (long long i = 1; i < n; ++i)
        if (hq(i, 0) != hq(i - 1, 0)) run_start.push_back(i);
    run_start.push_back(n);
    const size_t n_runs = run_start.size() - 1;
    if (n_pass) {
        size_t total = 0;
        for (int ip = 0; ip < n_pass; ++ip) {
            n_tiles[ip] = (plan->sdv[ip].n_z + SW_GM_BM - 1) / SW_GM_BM;
            eb_off[ip] = total;
            total += n_runs * (size_t)n_tiles[ip];
            run_kw[ip].assign(n_runs, 0);
        }
        int32_t* h_eb = ws.h_ebase.get<int32_t>(total); SW_ALLOC(h_eb);
        memset(h_eb, 0, total * sizeof(int32_t));
        for (int ip = 0; ip < n_pass; ++ip)
            for (size_t r = 0; r < n_runs; ++r)
                if (run_start[r + 1] - run_start[r] >= group_min)
                    run_kw[ip][r] = group_band(*plan, ip, hq(run_start[r], 0), h_eb + eb_off[ip] + r * n_tiles[ip]);
        d_ebase = ws.ebase.get<int32_t>(total); SW_ALLOC(d_ebase);
        SW_CUDA(cudaMemcpyAsync(d_ebase, h_eb, total * sizeof(int32_t), cudaMemcpyHostToDevice, s_main));
        // weight scratch of K6' (one per side stream), sized for the widest band of the sweep
        size_t need = 0;
        for (int ip = 0; ip < n_pass; ++ip)
            for (size_t r = 0; r < n_runs; ++r)
                if (run_kw[ip][r] > 0)
                    need = std::max(need, (size_t)n_tiles[ip] * SW_GM_BM * run_kw[ip][r] + 2 * (size_t)plan->sdv[ip].n_q);
        if (need) { SW_ALLOC(ws.omega[0].get<double>(need)); SW_ALLOC(ws.omega[1].get<double>(need)); }
    }
    // ---- sizes and scratch
    C = std::min<long long>(chunk, n);
    if (want_spec) {
        n_y = plan->gw.n_y; n_zl = plan->gw.n_zl; n_a2 = plan->a2.n_z;
        ws_stride = ptt_a2_work_stride((int)cap, plan->a2.nxi, n_a2);
        use_cells = plan->cells.cap > 0;
        a2_ld = (n_a2 + 1) / 2 * 2;
        d_a2 = ws.a2.get<double>((size_t)C * a2_ld); SW_ALLOC(d_a2);
        d_work = ws.work.get<double>((size_t)C * ws_stride); SW_ALLOC(d_work);
        d_pvl = ws.pv_l.get<double>((size_t)C * n_zl); SW_ALLOC(d_pvl);
        if (n_pass == 2 && plan->has_y_pass) { d_pvy = ws.pv_y.get<double>((size_t)C * plan->sdv[0].n_z); SW_ALLOC(d_pvy); }
        if (use_cells) { d_gwtab = ws.gw_table.get<double>((size_t)ptt_gw_cells_table_size(n_zl, (int)C)); SW_ALLOC(d_gwtab); }
        double* const user_tab[3] = {out.pow_v, out.pow_gw, out.omgw0};
        for (int k = 0; k < 3; ++k) {
            if (!user_tab[k] && k != 1) continue;      // pow_gw is always produced by the GW kernel
            for (int b = 0; b < 2; ++b) { o_tab[b][k] = ws.out[b][k].get<double>((size_t)C * n_y); SW_ALLOC(o_tab[b][k]); }
            if (user_tab[k] && host && !sorted) { d_full[k] = ws.full[k].get<double>(nn * n_y); SW_ALLOC(d_full[k]); }
        }
    }
    if (out.table_rows && want_spec) {
        // (pinned staging of the workspace: the copy is over when the first shell solve of this sweep is, and the call
        // does not return before that)
        long long* tr = ws.h_trows.get<long long>(nn); SW_ALLOC(tr);
        for (size_t i = 0; i < nn; ++i) tr[i] = (long long)out.table_rows[sorted ? i : (size_t)perm[i]];
        d_trows = ws.trows.get<long long>(nn); SW_ALLOC(d_trows);
        SW_CUDA(cudaMemcpyAsync(d_trows, tr, nn * sizeof(long long), cudaMemcpyHostToDevice, s_main));
    }
    if (out.snr && want_spec) { d_snr = ws.snr.get<double>(nn * out.n_snr); SW_ALLOC(d_snr); }
    if (host && !sorted) {
        if (out.scalars) { d_full_scal = ws.full_scal.get<double>(nn * SW_NS); SW_ALLOC(d_full_scal); }
        d_full_status = ws.full_status.get<int>(nn); SW_ALLOC(d_full_status);
    }
    size_t free_b = 0, total_b = 0;
    SW_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // memory the shell slots already hold is part of what they may use
    size_t held = 0;
    for (ShellSlot& s : ws.slot) for (DevBuf* b : s.dev()) held += b->cap;
    shell_chunk = opts.shell_chunk > 0 ? opts.shell_chunk : ptt_sweep_shell_chunk(n_xi, (int64_t)(free_b + held));
    shell_chunk = std::min<long long>(std::max<long long>(shell_chunk, 1), n);
    if (want_spec && shell_chunk > chunk && shell_chunk < n) shell_chunk = shell_chunk / chunk * chunk;
    n_shell_chunks = (n + shell_chunk - 1) / shell_chunk;
    for (int b = 0; b < (n_shell_chunks > 1 ? 2 : 1); ++b) {
        ShellSlot& sl = ws.slot[b];
        const size_t S = (size_t)shell_chunk;
        for (int k = 0; k < 3; ++k) SW_ALLOC(sl.rows[k].get<double>(S * cap));
        SW_ALLOC(sl.off.get<int>(S)); SW_ALLOC(sl.npts.get<int>(S)); SW_ALLOC(sl.gscal.get<double>(S * SW_GS));
        SW_ALLOC(sl.gstatus.get<int>(S)); SW_ALLOC(sl.I.get<double>(S * 8)); SW_ALLOC(sl.pp.get<double>(S * 8));
        SW_ALLOC(sl.pp2.get<double>(S * 12)); SW_ALLOC(sl.slf.get<double>(S)); SW_ALLOC(sl.table.get<double>(S * SW_NS));
        SW_ALLOC(sl.h_status.get<int>(S)); SW_ALLOC(sl.h_npts.get<int>(S));
    }
    if (host) {
        if (out.scalars && sorted) SW_ALLOC(ws.h_scal.get<double>(1));
    }
    // uploads queued on s_main must be visible to the other streams
    SW_CUDA(cudaEventRecord(ws.ev_tmp, s_main));
    for (cudaStream_t* s : ws.streams()) SW_CUDA(cudaStreamWaitEvent(*s, ws.ev_tmp, 0));
    return PTT_OK;
}

int Sweep::solve(const long long c) {
    ShellSlot& sl = ws.slot[c & 1];
    const long long s0 = c * shell_chunk, S = std::min(n, s0 + shell_chunk) - s0;
    if (c >= 2) SW_CUDA(cudaStreamWaitEvent(ws.s_shell, sl.consumed, 0));      // the slot's previous tenant is done
    cudaEvent_t t0 = timer.begin(ws.s_shell);
    SW_RC(ptt_sweep_shells_generic((int)S, d_pg + (size_t)s0 * SW_PG, n_xi, t_end, thin, wn_rtol, sl.rows[0].as<double>(),
                                   sl.rows[1].as<double>(), sl.rows[2].as<double>(), cap, sl.off.as<int32_t>(),
                                   sl.npts.as<int32_t>(), sl.gscal.as<double>(), sl.gstatus.as<int32_t>(), PTT_MEM_DEVICE,
                                   ws.s_shell));
    timer.end(PTT_STAGE_SHELLS, t0, ws.s_shell);
    stats.launches += 1;
    SW_CUDA(cudaMemcpyAsync(sl.h_status.p, sl.gstatus.p, S * sizeof(int), cudaMemcpyDeviceToHost, ws.s_shell));
    SW_CUDA(cudaMemcpyAsync(sl.h_npts.p, sl.npts.p, S * sizeof(int), cudaMemcpyDeviceToHost, ws.s_shell));
    SW_CUDA(cudaEventRecord(sl.solved, ws.s_shell));
    return PTT_OK;
}

// The reference's retries for the points whose first root search failed (fsolve_vary, speedup/solvers.py:12-55; backup
// scan of the hybrid solver, fluid.py:731-795), on a stream of their own under the spectrum stages.
int Sweep::launch_retry(const long long c, const std::vector<int>& failed) {
    ShellSlot& sl = ws.slot[c & 1];
    const long long s0 = c * shell_chunk;
    const int n_f = (int)failed.size();
    // capacity in entries: generous and doubling, so that the buffers are (re)allocated practically never -- cudaFree /
    // cudaMalloc / cudaHostAlloc in the middle of a sweep stall the host for tens of milliseconds each
    size_t cap_f = 256;
    while (cap_f < (size_t)n_f) cap_f *= 2;
    int* h_idx = sl.h_idx.get<int>(cap_f); SW_ALLOC(h_idx);
    SW_ALLOC(sl.h_rescued.get<int>(cap_f));
    memcpy(h_idx, failed.data(), n_f * sizeof(int));
    for (int k = 0; k < 3; ++k) SW_ALLOC(sl.c_rows[k].get<double>(cap_f * cap));
    SW_ALLOC(sl.c_off.get<int>(cap_f)); SW_ALLOC(sl.c_npts.get<int>(cap_f)); SW_ALLOC(sl.c_scal.get<double>(cap_f * SW_GS));
    SW_ALLOC(sl.c_status.get<int>(cap_f)); SW_ALLOC(sl.c_rescued.get<int>(cap_f)); SW_ALLOC(sl.c_idx.get<int>(cap_f));
    SW_ALLOC(sl.c_I.get<double>(cap_f * 8)); SW_ALLOC(sl.c_pp2.get<double>(cap_f * 12));
    cudaStream_t s = ws.s_retry;
    SW_CUDA(cudaStreamWaitEvent(s, sl.solved, 0));
    SW_CUDA(cudaMemcpyAsync(sl.c_idx.p, h_idx, n_f * sizeof(int), cudaMemcpyHostToDevice, s));
    SW_CUDA(cudaMemsetAsync(sl.c_off.p, 0, n_f * sizeof(int), s));
    SW_CUDA(cudaMemsetAsync(sl.c_npts.p, 0, n_f * sizeof(int), s));      // < 3 samples: "no profile" for K4
    cudaEvent_t t0 = timer.begin(s);
    SW_RC(ptt_generic_rescue(n_f, sl.c_idx.as<int32_t>(), d_pg + (size_t)s0 * SW_PG, n_xi, t_end, thin, wn_rtol,
                             sl.c_rows[0].as<double>(), sl.c_rows[1].as<double>(), sl.c_rows[2].as<double>(), cap,
                             sl.c_off.as<int32_t>(), sl.c_npts.as<int32_t>(), sl.c_scal.as<double>(),
                             sl.c_status.as<int32_t>(), sl.c_rescued.as<int32_t>(), s));
    timer.end(PTT_STAGE_RETRY, t0, s);
    stats.launches += 1;
    stats.retried += n_f;
    SW_CUDA(cudaMemcpyAsync(sl.h_rescued.p, sl.c_rescued.p, n_f * sizeof(int), cudaMemcpyDeviceToHost, s));
    SW_CUDA(cudaEventRecord(sl.retried, s));
    return PTT_OK;
}

int Sweep::thermo(const long long c) {
    ShellSlot& sl = ws.slot[c & 1];
    const long long s0 = c * shell_chunk, S = std::min(n, s0 + shell_chunk) - s0;
    SW_CUDA(cudaStreamWaitEvent(s_main, sl.solved, 0));
    cudaEvent_t t0 = timer.begin(s_main);
    sweep_prep_kernel<<<(unsigned)((S + 127) / 128), 128, 0, s_main>>>((int)S, d_pg + (size_t)s0 * SW_PG,
                                                                          d_pm + (size_t)s0 * SW_PM, sl.pp.as<double>(),
                                                                          sl.pp2.as<double>());
    SW_LAUNCHED("sweep_prep_kernel");
    SW_RC(ptt_profile_integrals_batch((int)S, sl.rows[0].as<double>(), sl.rows[1].as<double>(), sl.rows[2].as<double>(),
                                      cap, sl.off.as<int32_t>(), sl.npts.as<int32_t>(), sl.pp2.as<double>(),
                                      sl.I.as<double>(), s_main));
    thermo_finish_kernel<<<(unsigned)((S + 127) / 128), 128, 0, s_main>>>(
        (int)S, d_pg + (size_t)s0 * SW_PG, nullptr, nullptr, sl.gscal.as<double>(), sl.I.as<double>(), sl.npts.as<int>(),
        opts.r_star, opts.lifetime_multiplier, sl.table.as<double>(), sl.slf.as<double>());
    SW_LAUNCHED("thermo_finish_kernel");
    timer.end(PTT_STAGE_THERMO, t0, s_main);
    stats.launches += 3;
    return PTT_OK;
}

// Waits for the retry search of chunk c and writes the rescued points into the chunk: profiles, solver scalars,
// status, thermodynamic quantities, table rows.
int Sweep::fold(const long long c, const std::vector<int>& failed, std::vector<char>& valid) {
    ShellSlot& sl = ws.slot[c & 1];
    const long long s0 = c * shell_chunk;
    const int n_f = (int)failed.size();
    {
        const auto t0 = std::chrono::steady_clock::now();
        SW_CUDA(cudaEventSynchronize(sl.retried));
        stats.host_wait_retry_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    const int* h_resc = sl.h_rescued.as<int>();
    int n_r = 0;
    for (int f = 0; f < n_f; ++f)
        if (h_resc[f]) { valid[failed[f]] = 1; ++n_r; }
    if (!n_r) return PTT_OK;
    stats.rescued += n_r;
    SW_CUDA(cudaStreamWaitEvent(s_main, sl.retried, 0));
    cudaEvent_t t0 = timer.begin(s_main);
    gather_rows_kernel<int><<<(unsigned)(((size_t)n_f * 12 + 255) / 256), 256, 0, s_main>>>((long long)n_f, 12, sl.pp2.as<double>(),
                                                                                          sl.c_idx.as<int>(), sl.c_pp2.as<double>());
    SW_LAUNCHED("gather_rows_kernel");
    SW_RC(ptt_profile_integrals_batch(n_f, sl.c_rows[0].as<double>(), sl.c_rows[1].as<double>(), sl.c_rows[2].as<double>(),
                                      cap, sl.c_off.as<int32_t>(), sl.c_npts.as<int32_t>(), sl.c_pp2.as<double>(),
                                      sl.c_I.as<double>(), s_main));
    rescue_writeback_kernel<<<dim3(n_f, 3), 256, 0, s_main>>>(
        sl.c_idx.as<int>(), sl.c_rescued.as<int>(), cap, sl.c_rows[0].as<double>(), sl.c_rows[1].as<double>(),
        sl.c_rows[2].as<double>(), sl.c_off.as<int>(), sl.c_npts.as<int>(), sl.c_scal.as<double>(), sl.c_status.as<int>(),
        sl.rows[0].as<double>(), sl.rows[1].as<double>(), sl.rows[2].as<double>(), sl.off.as<int>(), sl.npts.as<int>(),
        sl.gscal.as<double>(), sl.gstatus.as<int>());
    SW_LAUNCHED("rescue_writeback_kernel");
    thermo_finish_kernel<<<(n_f + 127) / 128, 128, 0, s_main>>>(
        n_f, d_pg + (size_t)s0 * SW_PG, sl.c_idx.as<int>(), sl.c_rescued.as<int>(), sl.c_scal.as<double>(),
        sl.c_I.as<double>(), sl.c_npts.as<int>(), opts.r_star, opts.lifetime_multiplier, sl.table.as<double>(),
        sl.slf.as<double>());
    SW_LAUNCHED("thermo_finish_kernel");
    timer.end(PTT_STAGE_RETRY_FOLD, t0, s_main);
    stats.launches += 4;
    return PTT_OK;
}

// spec_den_v of one pass for the chunk rows [s, e): runs of >= SW_GROUP_MIN profiles with the same v_wall go through
// the banded matrix product (on two side streams), everything else through the gather kernel.
int Sweep::sdv_pass(const int ip, const long long c, const long long s, const long long e, const std::vector<char>& valid,
                    double* out_pv) {
    const long long s0 = c * shell_chunk;
    const ptt_sdv_plan_t& sp = plan->sdv[ip];
    const int n_z = sp.n_z;
    const long long P = e - s;
    const double* a2p = d_a2 + plan->seg[ip];
    ShellSlot& sl = ws.slot[c & 1];
    struct G { long long a, b; size_t run; };
    std::vector<G> big;
    std::vector<std::pair<long long, long long>> small_, nanfill;
    // runs intersecting [s, e)
    size_t r = std::upper_bound(run_start.begin(), run_start.end(), s) - run_start.begin() - 1;
    long long small_start = -1;
    for (; r + 1 < run_start.size() && run_start[r] < e; ++r) {
        const long long a = std::max(run_start[r], s), b = std::min(run_start[r + 1], e);
        const bool is_big = (b - a) >= group_min && run_kw[ip][r] > 0;
        if (is_big) {
            if (small_start >= 0) { small_.push_back({small_start, a}); small_start = -1; }
            long long b_eff = a;
            for (long long i = b; i > a; --i)
                if (valid[i - 1 - s0]) { b_eff = i; break; }
            if (b_eff < b) nanfill.push_back({b_eff, b});
            if (b_eff > a) big.push_back({a, b_eff, r});
        } else if (small_start < 0) small_start = a;
    }
    if (small_start >= 0) small_.push_back({small_start, e});
    const int stage = (ip == n_pass - 1) ? PTT_STAGE_SDV : PTT_STAGE_SDV_Y;
    cudaEvent_t t0 = timer.begin(s_main);
    // v_wall per profile for the gather kernel: a contiguous array
    for (const auto& sm : small_) {
        // pp rows have stride 8: copy the wall speeds of the range into the head of the (unused here) omega scratch
        const long long m = sm.second - sm.first;
        double* vw_c = ws.vw_tmp.get<double>((size_t)C);
        SW_ALLOC(vw_c);
        SW_CUDA(cudaMemcpy2DAsync(vw_c, sizeof(double), sl.pp.as<double>() + (size_t)(sm.first - s0) * 8, 8 * sizeof(double),
                                  sizeof(double), (size_t)m, cudaMemcpyDeviceToDevice, s_main));
        SW_RC(ptt_spec_den_v_batch(&sp, (int)m, a2p + (size_t)(sm.first - s) * a2_ld, a2_ld, vw_c,
                                   out_pv + (size_t)(sm.first - s) * n_z, n_z, s_main));
        stats.launches += 1;
    }
    for (const auto& nf : nanfill) {
        fill_rows_kernel<<<dim3((unsigned)(nf.second - nf.first), 4), 256, 0, s_main>>>(n_z, out_pv + (size_t)(nf.first - s) * n_z, n_z, NAN);
        SW_LAUNCHED("fill_rows_kernel");
        stats.launches += 1;
    }
    if (!big.empty()) {
        // independent groups on two side streams: the tail wave of one matrix product overlaps the next group's work
        const int n_side = big.size() > 1 ? 2 : 1;
        SW_CUDA(cudaEventRecord(ws.ev_fork, s_main));
        for (int k = 0; k < n_side; ++k) SW_CUDA(cudaStreamWaitEvent(ws.s_side[k], ws.ev_fork, 0));
        for (size_t k = 0; k < big.size(); ++k) {
            const G& g = big[k];
            const int side = (int)(k % n_side);
            const int kw = run_kw[ip][g.run];
            const size_t need = (size_t)n_tiles[ip] * SW_GM_BM * kw + 2 * (size_t)sp.n_q;
            double* om = ws.omega[side].get<double>(need);
            SW_ALLOC(om);
            SW_RC(ptt_spec_den_v_group_ex(&sp, (int)(g.b - g.a), a2p + (size_t)(g.a - s) * a2_ld, a2_ld, hq(g.a, 0), kw,
                                          d_ebase + eb_off[ip] + g.run * n_tiles[ip], om, out_pv + (size_t)(g.a - s) * n_z,
                                          n_z, PTT_SDV_ALIGNED_BAND, ws.s_side[side]));
            stats.launches += 2;
            if (ip == n_pass - 1) {
                stats.group_profiles += g.b - g.a;
                stats.group_count += 1;
                stats.group_kw_sum += kw;
                stats.group_macs += (long long)sp.n_z * kw * (g.b - g.a);
            }
        }
        for (int k = 0; k < n_side; ++k) {
            SW_CUDA(cudaEventRecord(ws.ev_join[k], ws.s_side[k]));
            SW_CUDA(cudaStreamWaitEvent(s_main, ws.ev_join[k], 0));
        }
    }
    timer.end(stage, t0, s_main);
    (void)P;
    return PTT_OK;
}

int Sweep::spectra_chunk(const long long c, const long long s, const long long e, const std::vector<char>& valid) {
    ShellSlot& sl = ws.slot[c & 1];
    const long long s0 = c * shell_chunk;
    const int P = (int)(e - s);
    const size_t lo = (size_t)(s - s0);
    const int b = (int)(emitted & 1);
    // the output slot's previous tenant has been delivered
    if (out_used[b]) SW_CUDA(cudaStreamWaitEvent(s_main, ws.out_free[b], 0));
    cudaEvent_t t0 = timer.begin(s_main);
    SW_RC(a2_batch_strided(&plan->a2, P, sl.rows[0].as<double>() + lo * cap, sl.rows[1].as<double>() + lo * cap,
                           sl.rows[2].as<double>() + lo * cap, cap, sl.off.as<int32_t>() + lo, sl.npts.as<int32_t>() + lo,
                           sl.pp.as<double>() + lo * 8, d_work, ws_stride, d_a2, a2_ld, nullptr, nullptr, s_main));
    timer.end(PTT_STAGE_A2, t0, s_main);
    stats.launches += 3;
    stats.a2_profiles += P;
    stats.gw_profiles += P;
    if (n_pass == 2 && plan->has_y_pass && out.pow_v) {
        SW_RC(sdv_pass(0, c, s, e, valid, d_pvy));
        SW_RC(ptt_pow_spec_batch(P, plan->sdv[0].n_z, plan->gw.y, d_pvy, plan->sdv[0].n_z, o_tab[b][0], s_main));
        stats.launches += 1;
    }
    SW_RC(sdv_pass(n_pass - 1, c, s, e, valid, d_pvl));
    t0 = timer.begin(s_main);
    const double* scale_c = (out.omgw0 && d_scale) ? d_scale + s : nullptr;
    if (use_cells && P >= SW_GW_CELLS_MIN) {
        SW_RC(ptt_spec_den_gw_cells_batch(&plan->gw, &plan->cells, P, d_pvl, n_zl, sl.slf.as<double>() + lo, scale_c, d_gwtab,
                                          nullptr, o_tab[b][1], scale_c ? o_tab[b][2] : nullptr, s_main));
        stats.launches += 2;
    } else {
        SW_RC(ptt_spec_den_gw_batch(&plan->gw, P, d_pvl, n_zl, sl.slf.as<double>() + lo, scale_c, nullptr, o_tab[b][1],
                                    scale_c ? o_tab[b][2] : nullptr, s_main));
        stats.launches += 1;
    }
    if (d_snr && scale_c) {
        SW_RC(ptt_snr_batch(P, n_y, o_tab[b][2], n_y, out.n_snr, out.snr_weight, d_snr + (size_t)s * out.n_snr, s_main));
        stats.launches += 1;
    }
    timer.end(PTT_STAGE_GW, t0, s_main);
    // ---- deliver: the copy stream takes over, the compute stream goes on with the next chunk
    SW_CUDA(cudaEventRecord(ws.ev_emit, s_main));
    SW_CUDA(cudaStreamWaitEvent(ws.s_copy, ws.ev_emit, 0));
    t0 = timer.begin(ws.s_copy);
    double* const user_tab[3] = {out.pow_v, out.pow_gw, out.omgw0};
    for (int k = 0; k < 3; ++k) {
        if (!user_tab[k]) continue;
        const size_t bytes = (size_t)P * n_y * sizeof(double);
        if (d_trows) {
            scatter_rows_kernel<double><<<dim3(P, 4), 256, 0, ws.s_copy>>>(n_y, o_tab[b][k], n_y, d_trows + s, user_tab[k], n_y);
            SW_LAUNCHED("scatter_rows_kernel");
            stats.launches += 1;
        } else if (sorted) {
            SW_CUDA(cudaMemcpyAsync(user_tab[k] + (size_t)s * n_y, o_tab[b][k], bytes,
                                    host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ws.s_copy));
        } else {
            double* dst = host ? d_full[k] : user_tab[k];
            scatter_rows_kernel<double><<<dim3(P, 4), 256, 0, ws.s_copy>>>(n_y, o_tab[b][k], n_y, d_perm + s, dst, n_y);
            SW_LAUNCHED("scatter_rows_kernel");
            stats.launches += 1;
        }
    }
    timer.end(PTT_STAGE_EMIT, t0, ws.s_copy);
    SW_CUDA(cudaEventRecord(ws.out_free[b], ws.s_copy));
    out_used[b] = true;
    ++emitted;
    if (cb && sorted) cb(user, s, e, ws.s_copy);
    return PTT_OK;
}

// status, the table of scalars and (optionally) the profile rows of a shell chunk -> caller
int Sweep::finish_shell_chunk(const long long c) {
    ShellSlot& sl = ws.slot[c & 1];
    const long long s0 = c * shell_chunk, S = std::min(n, s0 + shell_chunk) - s0;
    cudaStream_t s = s_main;
    if (sorted) {
        const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        SW_CUDA(cudaMemcpyAsync(out.status + s0, sl.gstatus.p, S * sizeof(int), kind, s));
        if (out.scalars) SW_CUDA(cudaMemcpyAsync(out.scalars + (size_t)s0 * SW_NS, sl.table.p, S * SW_NS * sizeof(double), kind, s));
    } else {
        int* st_dst = host ? d_full_status : out.status;
        scatter_rows_kernel<int><<<dim3((unsigned)S, 1), 32, 0, s>>>(1, sl.gstatus.as<int>(), 1, d_perm + s0, st_dst, 1);
        SW_LAUNCHED("scatter_rows_kernel");
        if (out.scalars) {
            double* sc_dst = host ? d_full_scal : out.scalars;
            scatter_rows_kernel<double><<<dim3((unsigned)S, 1), 32, 0, s>>>(SW_NS, sl.table.as<double>(), SW_NS, d_perm + s0, sc_dst, SW_NS);
            SW_LAUNCHED("scatter_rows_kernel");
        }
    }
    if (want_rows) {
        double* const dst[3] = {out.rows_v, out.rows_w, out.rows_xi};
        for (int k = 0; k < 3; ++k) {
            if (sorted) SW_CUDA(cudaMemcpy2DAsync(dst[k] + (size_t)s0 * out.row_stride, out.row_stride * sizeof(double),
                                                  sl.rows[k].p, cap * sizeof(double), cap * sizeof(double), (size_t)S,
                                                  cudaMemcpyDeviceToDevice, s));
            else {
                scatter_rows_kernel<double><<<dim3((unsigned)S, 8), 256, 0, s>>>((int)cap, sl.rows[k].as<double>(), cap, d_perm + s0, dst[k], out.row_stride);
                SW_LAUNCHED("scatter_rows_kernel");
            }
        }
        if (sorted) {
            SW_CUDA(cudaMemcpyAsync(out.off + s0, sl.off.p, S * sizeof(int), cudaMemcpyDeviceToDevice, s));
            SW_CUDA(cudaMemcpyAsync(out.npts + s0, sl.npts.p, S * sizeof(int), cudaMemcpyDeviceToDevice, s));
        } else {
            scatter_rows_kernel<int><<<dim3((unsigned)S, 1), 32, 0, s>>>(1, sl.off.as<int>(), 1, d_perm + s0, out.off, 1);
            scatter_rows_kernel<int><<<dim3((unsigned)S, 1), 32, 0, s>>>(1, sl.npts.as<int>(), 1, d_perm + s0, out.npts, 1);
            SW_LAUNCHED("scatter_rows_kernel");
        }
    }
    SW_CUDA(cudaEventRecord(sl.consumed, s_main));
    return PTT_OK;
}

int Sweep::run() {
    SW_RC(solve(0));
    for (long long c = 0; c < n_shell_chunks; ++c) {
        ShellSlot& sl = ws.slot[c & 1];
        const long long s0 = c * shell_chunk, e0 = std::min(n, s0 + shell_chunk), S = e0 - s0;
        {
            const auto t0 = std::chrono::steady_clock::now();
            SW_CUDA(cudaEventSynchronize(sl.solved));
            stats.host_wait_shells_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        const int* h_status = sl.h_status.as<int>();
        const int* h_npts = sl.h_npts.as<int>();
        std::vector<int> failed;
        if (retry)
            for (long long i = 0; i < S; ++i) {
                const int code = (int)hq(s0 + i, 2);
                if ((h_status[i] == PTT_ST_SOLVER_FAILED || h_status[i] == PTT_ST_ROOT_NOT_CONVERGED) && (code == 0 || code == 1))
                    failed.push_back((int)i);
            }
        if (!failed.empty()) SW_RC(launch_retry(c, failed));
        // the next shell chunk is solved under this chunk's spectrum stages
        if (c + 1 < n_shell_chunks) SW_RC(solve(c + 1));
        SW_RC(thermo(c));
        std::vector<char> valid((size_t)S);
        for (long long i = 0; i < S; ++i) valid[i] = h_npts[i] >= 3;      // has a profile (solver_failed points do)
        bool folded = failed.empty();
        if (want_spec) {
            // Spectrum chunks without retried points go first: the retries finish under them and are written into the
            // chunk before the chunks that hold such points start.
            std::vector<std::pair<long long, long long>> chunks, later;
            for (long long s = s0; s < e0; s += chunk) {
                const long long e = std::min(e0, s + chunk);
                bool has_failed = false;
                for (const int f : failed) has_failed |= (s0 + f >= s && s0 + f < e);
                (has_failed ? later : chunks).push_back({s, e});
            }
            for (const auto& ch : chunks) SW_RC(spectra_chunk(c, ch.first, ch.second, valid));
            if (!folded) { SW_RC(fold(c, failed, valid)); folded = true; }
            for (const auto& ch : later) SW_RC(spectra_chunk(c, ch.first, ch.second, valid));
            stats.chunks_before_fold += (long long)chunks.size();
            stats.chunks_after_fold += (long long)later.size();
        }
        if (!folded) SW_RC(fold(c, failed, valid));
        SW_RC(finish_shell_chunk(c));
    }
    // ---- wrap up: everything the side streams still hold is ordered before the end of the call on s_main
    SW_CUDA(cudaEventRecord(ws.ev_tmp, ws.s_copy));
    SW_CUDA(cudaStreamWaitEvent(s_main, ws.ev_tmp, 0));
    if (host && !sorted) {
        const size_t nn = (size_t)n;
        double* const user_tab[3] = {out.pow_v, out.pow_gw, out.omgw0};
        for (int k = 0; k < 3; ++k)
            if (user_tab[k]) SW_CUDA(cudaMemcpyAsync(user_tab[k], d_full[k], nn * n_y * sizeof(double), cudaMemcpyDeviceToHost, s_main));
        SW_CUDA(cudaMemcpyAsync(out.status, d_full_status, nn * sizeof(int), cudaMemcpyDeviceToHost, s_main));
        if (out.scalars) SW_CUDA(cudaMemcpyAsync(out.scalars, d_full_scal, nn * SW_NS * sizeof(double), cudaMemcpyDeviceToHost, s_main));
    }
    if (d_snr) {
        const size_t cnt = (size_t)n * out.n_snr;
        const double* src = d_snr;
        if (!sorted) {
            double* t = ws.snr_perm.get<double>(cnt); SW_ALLOC(t);
            scatter_rows_kernel<double><<<dim3((unsigned)n, 1), 32, 0, s_main>>>(out.n_snr, d_snr, out.n_snr, d_perm, t, out.n_snr);
            SW_LAUNCHED("scatter_rows_kernel");
            src = t;
        }
        SW_CUDA(cudaMemcpyAsync(out.snr, src, cnt * sizeof(double), host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s_main));
    }
    if (cb && !sorted) cb(user, 0, n, s_main);
    if (host) {
        SW_CUDA(cudaStreamSynchronize(s_main));
        SW_CUDA(cudaStreamSynchronize(ws.s_shell));
        SW_CUDA(cudaStreamSynchronize(ws.s_retry));
        timer.collect(stats);
    } else if (timer.on) {
        // device tables: the call returns with its last chunks still in flight (the caller's next host work overlaps
        // them); the stage times are read from the events when somebody asks (ptt_sweep_timing_sync, or the next sweep)
        ws.pending[ws.tset].recs.swap(timer.recs);
    }
    return PTT_OK;
}

}  // namespace ptt

using namespace ptt;

extern "C" int ptt_ipc_alloc(int64_t bytes, void** ptr, unsigned char* handle64) {
    if (bytes <= 0 || !ptr || !handle64) return set_error(PTT_E_ARG, "ptt_ipc_alloc: null pointer or bad size");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) return set_cuda_error("ptt_ipc_alloc: cudaMalloc", e);
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return set_cuda_error("ptt_ipc_alloc: cudaIpcGetMemHandle", e); }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return PTT_OK;
}

extern "C" int ptt_ipc_open(const unsigned char* handle64, void** ptr) {
    if (!handle64 || !ptr) return set_error(PTT_E_ARG, "ptt_ipc_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return set_cuda_error("ptt_ipc_open: cudaIpcOpenMemHandle", e);
    *ptr = p;
    return PTT_OK;
}

extern "C" int ptt_ipc_close(void* ptr) {
    if (!ptr) return PTT_OK;
    const cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) return set_cuda_error("ptt_ipc_close", e);
    return PTT_OK;
}

extern "C" int ptt_ipc_free(void* ptr) {
    if (!ptr) return PTT_OK;
    const cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) return set_cuda_error("ptt_ipc_free", e);
    return PTT_OK;
}

extern "C" int ptt_sweep_release(void) {
    std::lock_guard<std::mutex> lock(g_ws.mu);
    g_ws.release();
    return PTT_OK;
}

// Stage times of a device-table sweep are read from its events on demand: waits for that sweep to finish.  The completed
// stats go to the list of finished sweeps (and to g_stats if it was the last sweep).
static int timing_complete_locked(const int set) {
    Workspace::Pending& p = g_ws.pending[set];
    if (p.recs.empty()) return PTT_OK;
    for (const TimingRec& r : p.recs) {
        const cudaError_t e = cudaEventSynchronize(r.b);
        if (e != cudaSuccess) { p.recs.clear(); return set_cuda_error("ptt_sweep_timing_sync", e); }
    }
    Timer::collect_recs(p.recs, p.stats);
    p.recs.clear();
    if (g_stats.timing_serial == p.stats.timing_serial) g_stats = p.stats;
    g_ws.done.push_back(p.stats);
    if (g_ws.done.size() > 8) g_ws.done.erase(g_ws.done.begin());
    return PTT_OK;
}
static int timing_sync_locked() {
    const int older = 1 - g_ws.tset;
    SW_RC(timing_complete_locked(older));
    return timing_complete_locked(g_ws.tset);
}

extern "C" int ptt_sweep_timing_sync(void) {
    std::lock_guard<std::mutex> lock(g_ws.mu);
    return timing_sync_locked();
}

extern "C" int ptt_sweep_timing_collect(int64_t serial, ptt_sweep_stats_t* out) {
    if (!out) return set_error(PTT_E_ARG, "ptt_sweep_timing_collect: null pointer");
    std::lock_guard<std::mutex> lock(g_ws.mu);
    for (int k = 0; k < 2; ++k)
        if (!g_ws.pending[k].recs.empty() && g_ws.pending[k].stats.timing_serial == serial) SW_RC(timing_complete_locked(k));
    for (const ptt_sweep_stats_t& st : g_ws.done)
        if (st.timing_serial == serial) { *out = st; return PTT_OK; }
    if (g_stats.timing_serial == serial) { *out = g_stats; return PTT_OK; }      // a sweep that had nothing pending
    return set_error(PTT_E_ARG, "ptt_sweep_timing_collect: no sweep with this serial among the last eight timed ones");
}

extern "C" int ptt_sweep_last_stats(ptt_sweep_stats_t* out) {
    if (!out) return set_error(PTT_E_ARG, "ptt_sweep_last_stats: null pointer");
    *out = g_stats;
    return PTT_OK;
}

extern "C" int64_t ptt_sweep_shell_chunk(int n_xi, int64_t free_bytes) {
    const int64_t cap = ptt_generic_row_capacity(n_xi);
    const int64_t per_point = 2 * (3 * 8 * cap + 1024);      // two slots of profile rows + the small tables
    int64_t c = (int64_t)(0.45 * (double)free_bytes) / per_point;
    c = std::min<int64_t>(c, 65536);
    return std::max<int64_t>(c, 256);
}

extern "C" int ptt_sweep_spectra(const ptt_sweep_plan_t* plan, const ptt_sweep_opts_t* opts, int64_t n, const double* pg,
                                 const double* pm, const double* scale, const ptt_sweep_out_t* out, ptt_chunk_cb cb,
                                 void* user, int mem, void* stream_) {
    if (n == 0) return PTT_OK;      // an empty sweep is valid, whatever the pointers
    if (!opts || n < 0 || n > 0x7fffffffLL || !pg || !pm || !out || !out->status)
        return set_error(PTT_E_ARG, "ptt_sweep_spectra: null pointer or bad size");
    if (mem != PTT_MEM_HOST && mem != PTT_MEM_DEVICE) return set_error(PTT_E_ARG, "ptt_sweep_spectra: bad mem");
    const bool want_spec = out->pow_v || out->pow_gw || out->omgw0;
    if (want_spec && !plan) return set_error(PTT_E_ARG, "ptt_sweep_spectra: spectra requested without a plan");
    if (want_spec && (plan->n_pass < 1 || plan->n_pass > 2 || !plan->h_zterm[plan->n_pass - 1]))
        return set_error(PTT_E_ARG, "ptt_sweep_spectra: malformed plan");
    if (out->omgw0 && !scale) return set_error(PTT_E_ARG, "ptt_sweep_spectra: omgw0 needs scale[n]");
    if (out->snr && (!out->omgw0 || !out->snr_weight || out->n_snr < 1 || out->n_snr > 4))
        return set_error(PTT_E_ARG, "ptt_sweep_spectra: snr needs omgw0, snr_weight and 1 <= n_snr <= 4");
    if (out->pow_v && !(plan->n_pass == 2 && plan->has_y_pass))
        return set_error(PTT_E_ARG, "ptt_sweep_spectra: pow_v needs a plan with the y pass");
    const bool want_rows = out->rows_v || out->rows_w || out->rows_xi;
    const int n_xi = opts->n_xi > 0 ? opts->n_xi : 5000;
    if (n_xi < 32) return set_error(PTT_E_ARG, "ptt_sweep_spectra: n_xi < 32");
    if (want_rows && (!out->rows_v || !out->rows_w || !out->rows_xi || !out->off || !out->npts ||
                      out->row_stride < ptt_generic_row_capacity(n_xi)))
        return set_error(PTT_E_ARG, "ptt_sweep_spectra: profile output needs rows_v/w/xi, off, npts and row_stride >= "
                                    "ptt_generic_row_capacity(n_xi)");
    if (want_rows && mem != PTT_MEM_DEVICE)
        return set_error(PTT_E_ARG, "ptt_sweep_spectra: profile rows are returned in device memory only");
    if (out->table_rows && mem != PTT_MEM_DEVICE)
        return set_error(PTT_E_ARG, "ptt_sweep_spectra: table_rows needs device tables");
    std::lock_guard<std::mutex> lock(g_ws.mu);
    SW_RC(g_ws.init());
    // the other set of stage events: whatever is still unread there belongs to the sweep before the last one, which is over
    g_ws.tset = 1 - g_ws.tset;
    SW_RC(timing_complete_locked(g_ws.tset));
    g_ws.timing_used = 0;
    Sweep sw(g_ws, plan, *opts, *out, cb, user, mem, (cudaStream_t)stream_);
    int rc = sw.setup(n, pg, pm, scale);
    if (rc == PTT_OK) rc = sw.run();
    if (rc != PTT_OK) {
        cudaDeviceSynchronize();      // leave no work in flight that references the workspace
        g_ws.pending[0].recs.clear(); g_ws.pending[1].recs.clear();
        return rc;
    }
    sw.stats.timing_serial = ++g_ws.serial;
    g_stats = sw.stats;
    g_ws.pending[g_ws.tset].stats = sw.stats;
    return PTT_OK;
}
