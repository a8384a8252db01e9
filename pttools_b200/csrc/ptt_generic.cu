// K3: model-independent shell solve, one (model, v_wall, alpha_n) point per thread.
//   sound_shell_generic     pttools/bubble/fluid.py:848-1052
#include "ptt_generic.cuh"
#include "ptt_host.cuh"

namespace ptt {

constexpr int GEN_PARAMS = 16;
constexpr int GEN_SCALARS = 16;

// pg[n,16] = {v_wall, alpha_n, sol_type, wn, v_cj, guess_vp, guess_vp_tilde, guess_wp, guess_wm,
//             css2, csb2, V_s, V_b, bag_name, s_coef_s, s_coef_b}   (s_coef: entropy s(w) = s_coef w^(1-1/mu), used by
//             the retry search only)
// scalars[n,16] = {vp, vm, vp_tilde, vm_tilde, v_sh, vm_sh, vm_tilde_sh, wp, wm, wm_sh, v_cj, wn, wn_estimate,
//                  solver_failed, n_resid, 0}   (the reference's 17-tuple, fluid.py:1052; n_resid: residual
//                  evaluations of the root search, diagnostics)
// The integration gets a register allocation of its own: the machine's state (root finder, residual in flight, results)
// is live across every integration and, inlined, competes with the Runge-Kutta stages for the 255 registers.
__device__ __noinline__ int generic_run_job(const double cs2, const State y0, const double te, const int n_xi,
                                            JobVisitor* pv) {
    JobVisitor vis = *pv;
    const int rc = integrate_grid(cs2, y0, te, n_xi, vis);
    *pv = vis;
    return rc;
}

__global__ void __launch_bounds__(64) generic_solve_kernel(
    const int n, const double* __restrict__ pg, const int n_xi, const double t_end, const int thin,
    const double wn_rtol, double* __restrict__ rows_v, double* __restrict__ rows_w, double* __restrict__ rows_xi,
    const long long row_stride, int* __restrict__ off, int* __restrict__ npts, double* __restrict__ scalars,
    int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#ifdef PTT_K3_CLOCKS
    const long long clk0 = clock64();      // diagnostics (tools/k3_clocks.py): cycles this lane's warp spent until it was done
#endif
    const double* q = pg + (size_t)i * GEN_PARAMS;
    RowOut row;
    row.v = rows_v + (size_t)i * row_stride;
    row.w = rows_w + (size_t)i * row_stride;
    row.xi = rows_xi + (size_t)i * row_stride;
    double* sc = scalars + (size_t)i * GEN_SCALARS;
    const double v_wall = q[0], alpha_n = q[1];
    const int sol_type = (int)q[2];
    const double wn = q[3], v_cj = q[4];
    bool valid = (v_wall > 0.0 && v_wall < 1.0) && (alpha_n > 0.0) && (wn > 0.0) &&
                 (sol_type == SOL_SUB_DEF || sol_type == SOL_HYBRID || sol_type == SOL_DETON);
    GenericOut o;
    if (valid) {
        EosDev m;
        m.css2 = q[9]; m.csb2 = q[10];
        m.mu_s = 1.0 + 1.0 / m.css2; m.mu_b = 1.0 + 1.0 / m.csb2;
        m.V_s = q[11]; m.V_b = q[12];
        m.bag_name = (int)q[13];
        GenericGuess g;
        g.vp = q[5]; g.vp_tilde = q[6]; g.wp = q[7]; g.wm = q[8];
        GenericMachine mach;
        mach.start(m, v_wall, sol_type, wn, v_cj, g, n_xi, t_end, thin, wn_rtol, row);
        while (mach.phase != GenericMachine::PH_DONE) {
            const int rc = generic_run_job(mach.job_cs2, mach.job_y0, mach.job_te, n_xi, &mach.vis);
            mach.advance(rc);
        }
        o = mach.o;
    } else {
        o.status = ST_INVALID_INPUT;
    }
    if (o.status != ST_OK) {
        row.set(0, NAN, NAN, NAN);
        off[i] = 0; npts[i] = 1;
        for (int k = 0; k < GEN_SCALARS; ++k) sc[k] = NAN;
        sc[13] = 1.0;
        status[i] = o.status;
        return;
    }
    off[i] = o.off;
    npts[i] = o.npts;
    sc[0] = o.vp; sc[1] = o.vm; sc[2] = o.vp_tilde; sc[3] = o.vm_tilde; sc[4] = o.v_sh; sc[5] = o.vm_sh;
    sc[6] = o.vm_tilde_sh; sc[7] = o.wp; sc[8] = o.wm; sc[9] = o.wm_sh; sc[10] = o.v_cj; sc[11] = o.wn;
    sc[12] = o.wn_estimate; sc[13] = (double)o.solver_failed; sc[14] = (double)o.n_resid; sc[15] = 0.0;
#ifdef PTT_K3_CLOCKS
    sc[15] = (double)(clock64() - clk0);
#endif
    status[i] = o.solver_failed ? ST_SOLVER_FAILED : ST_OK;
}

// K3r: what the reference does when its first root search fails (fsolve_vary with two varied starts,
// speedup/solvers.py:12-55; for hybrids then the 20-point backup scan, fluid.py:731-795).  One warp per failed point,
// one candidate start per lane, all root searches at once; the first candidate IN THE REFERENCE'S ORDER whose fsolve
// converges gives wm, and only that lane evaluates the final profile.  Outputs are compact (row f of the failed
// list); rescued[f] = 1 when a candidate converged and its profile was written.
constexpr int RESCUE_WARPS = 8;      // failed points per CTA: the retries of a batch then sit on one or two SMs, not on
                                     // one SM per point (a resident retry warp keeps a 512-thread GEMM CTA off its SM)
__global__ void __launch_bounds__(32 * RESCUE_WARPS) generic_rescue_kernel(
    const int n_f, const int* __restrict__ idx, const double* __restrict__ pg, const int n_xi, const double t_end,
    const int thin, const double wn_rtol, double* __restrict__ rows_v, double* __restrict__ rows_w,
    double* __restrict__ rows_xi, const long long row_stride, int* __restrict__ off, int* __restrict__ npts,
    double* __restrict__ scalars, int* __restrict__ status, int* __restrict__ rescued) {
    const int f = blockIdx.x * RESCUE_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= n_f) return;
    const double* q = pg + (size_t)idx[f] * GEN_PARAMS;
    RowOut row;
    row.v = rows_v + (size_t)f * row_stride;
    row.w = rows_w + (size_t)f * row_stride;
    row.xi = rows_xi + (size_t)f * row_stride;
    const double v_wall = q[0];
    const int sol_type = (int)q[2];
    const double wn = q[3], v_cj = q[4];
    if (lane == 0) rescued[f] = 0;
    if (!(sol_type == SOL_SUB_DEF || sol_type == SOL_HYBRID)) return;
    EosDev m;
    m.css2 = q[9]; m.csb2 = q[10];
    m.mu_s = 1.0 + 1.0 / m.css2; m.mu_b = 1.0 + 1.0 / m.csb2;
    m.V_s = q[11]; m.V_b = q[12];
    m.bag_name = (int)q[13];
    GenericGuess g;
    g.vp = q[5]; g.vp_tilde = q[6]; g.wp = q[7]; g.wm = q[8];
    double x0 = NAN;
    const bool exists = generic_retry_start(m, v_wall, sol_type, sqrt(m.css2), g.wm, g.vp_tilde, q[14], q[15], lane, x0);
    GenericMachine mach;
    mach.phase = GenericMachine::PH_HOLD;
    mach.converged = false;
    if (exists) {
        g.wm = x0;
        mach.start(m, v_wall, sol_type, wn, v_cj, g, n_xi, t_end, thin, wn_rtol, row, true);
    }
    int winner = -1;
    for (;;) {
        const bool running = mach.phase != GenericMachine::PH_DONE && mach.phase != GenericMachine::PH_HOLD;
        const unsigned run_mask = __ballot_sync(0xffffffffu, running);
        const unsigned conv_mask = __ballot_sync(0xffffffffu, exists && mach.phase == GenericMachine::PH_HOLD &&
                                                                   mach.converged);
        if (conv_mask) {
            const int c = __ffs(conv_mask) - 1;
            if (!(run_mask & ((1u << c) - 1u))) { winner = c; break; }      // every earlier candidate has failed
        }
        if (!run_mask) break;
        if (running) {
            const int rc = generic_run_job(mach.job_cs2, mach.job_y0, mach.job_te, n_xi, &mach.vis);
            mach.advance(rc);
        }
    }
    if (lane != winner) return;
    mach.resume_final();
    while (mach.phase != GenericMachine::PH_DONE) {
        const int rc = generic_run_job(mach.job_cs2, mach.job_y0, mach.job_te, n_xi, &mach.vis);
        mach.advance(rc);
    }
    const GenericOut o = mach.o;
    if (o.status != ST_OK) return;
    double* sc = scalars + (size_t)f * GEN_SCALARS;
    off[f] = o.off;
    npts[f] = o.npts;
    sc[0] = o.vp; sc[1] = o.vm; sc[2] = o.vp_tilde; sc[3] = o.vm_tilde; sc[4] = o.v_sh; sc[5] = o.vm_sh;
    sc[6] = o.vm_tilde_sh; sc[7] = o.wp; sc[8] = o.wm; sc[9] = o.wm_sh; sc[10] = o.v_cj; sc[11] = o.wn;
    sc[12] = o.wn_estimate; sc[13] = (double)o.solver_failed; sc[14] = (double)o.n_resid; sc[15] = (double)(winner + 1);
    status[f] = o.solver_failed ? ST_SOLVER_FAILED : ST_OK;
    rescued[f] = 1;
}

// Wall quantities of a bag shell as the reference reads them off the arrays (props.v_and_w_from_solution,
// props.py:52-87) -- used to build the starting-guess table of fluid_reference.py:166-196.
__global__ void wall_values_kernel(const int n, const double* __restrict__ rows_v, const double* __restrict__ rows_w,
                                   const double* __restrict__ rows_xi, const long long row_stride,
                                   const int* __restrict__ off, const int* __restrict__ npts,
                                   const double* __restrict__ v_wall, const int* __restrict__ sol_type,
                                   double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* o = out + (size_t)i * 8;
    const int N = npts[i];
    if (N < 3) { for (int k = 0; k < 8; ++k) o[k] = NAN; return; }
    const double* v = rows_v + (size_t)i * row_stride + off[i];
    const double* w = rows_w + (size_t)i * row_stride + off[i];
    int i_wall = 0;
    double vmax = v[0];
    for (int k = 1; k < N; ++k) if (v[k] > vmax) { vmax = v[k]; i_wall = k; }
    if (sol_type[i] == SOL_DETON) i_wall += 1;
    if (i_wall < 1 || i_wall >= N) { for (int k = 0; k < 8; ++k) o[k] = NAN; return; }
    const double vw = v_wall[i];
    const double vp = v[i_wall], vm = v[i_wall - 1];
    o[0] = vp; o[1] = vm; o[2] = lorentz(vw, vp); o[3] = lorentz(vw, vm); o[4] = w[i_wall]; o[5] = w[i_wall - 1];
    o[6] = w[N - 1]; o[7] = 0.0;
}

// Nearest table entry of the same set for every query point (brute force, no temporaries): the batched form of the
// NearestNDInterpolator lookups of FluidReference.get (fluid_reference.py:60-62, 152-163).  Ties: the first entry.
__global__ void nearest_kernel(const int n, const double* __restrict__ qx, const double* __restrict__ qy,
                               const int* __restrict__ sel, const int n_sets, const int* __restrict__ set_off,
                               const double* __restrict__ cx, const double* __restrict__ cy, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = sel[i];
    int best = -1;
    if (s >= 0 && s < n_sets) {
        const double x = qx[i], y = qy[i];
        double d_best = INFINITY;
        for (int k = set_off[s]; k < set_off[s + 1]; ++k) {
            const double dx = x - cx[k], dy = y - cy[k];
            const double d = dx * dx + dy * dy;
            if (d < d_best) { d_best = d; best = k; }
        }
    }
    out[i] = best;
}

}  // namespace ptt

using namespace ptt;

extern "C" int ptt_nearest_batch(int n, const double* qx, const double* qy, const int32_t* sel, int n_sets,
                                 const int32_t* set_off, const double* cx, const double* cy, int32_t* out_idx,
                                 void* stream_) {
    if (n == 0) return PTT_OK;
    if (n < 0 || n_sets < 0 || !qx || !qy || !sel || !set_off || !out_idx || (n_sets > 0 && (!cx || !cy)))
        return set_error(PTT_E_ARG, "ptt_nearest_batch: null pointer or bad size");
    nearest_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(n, qx, qy, sel, n_sets, set_off, cx, cy, out_idx);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("nearest_kernel", e);
    return PTT_OK;
}

extern "C" int ptt_generic_row_capacity(int n_xi) { return gen_row_capacity(n_xi); }

extern "C" int ptt_sweep_shells_generic(int n, const double* pg, int n_xi, double t_end, int thin_shell_limit,
                                        double wn_rtol, double* rows_v, double* rows_w, double* rows_xi,
                                        int64_t row_stride, int32_t* off, int32_t* npts, double* scalars,
                                        int32_t* status, int mem, void* stream_) {
    if (n == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (n < 0 || n_xi < 32 || !pg || !rows_v || !rows_w || !rows_xi || !off || !npts || !scalars || !status ||
        row_stride < gen_row_capacity(n_xi))
        return set_error(PTT_E_ARG, "ptt_sweep_shells_generic: null pointer, bad size or row_stride too small");
    if (n == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    Stager sg(mem, stream);
    const size_t nn = (size_t)n, nr = (size_t)n * (size_t)row_stride;
    const double* d_pg = sg.in(pg, nn * GEN_PARAMS);
    double* d_rv = sg.out(rows_v, nr); double* d_rw = sg.out(rows_w, nr); double* d_rx = sg.out(rows_xi, nr);
    int* d_off = sg.out(off, nn); int* d_np = sg.out(npts, nn);
    double* d_sc = sg.out(scalars, nn * GEN_SCALARS);
    int* d_st = sg.out(status, nn);
    if (!sg.ok()) return sg.fail();
    generic_solve_kernel<<<(n + 63) / 64, 64, 0, stream>>>(n, d_pg, n_xi, t_end, thin_shell_limit, wn_rtol, d_rv, d_rw,
                                                              d_rx, (long long)row_stride, d_off, d_np, d_sc, d_st);
    return sg.finish("generic_solve_kernel");
}

extern "C" int ptt_generic_rescue(int n_f, const int32_t* idx, const double* pg, int n_xi, double t_end,
                                  int thin_shell_limit, double wn_rtol, double* rows_v, double* rows_w, double* rows_xi,
                                  int64_t row_stride, int32_t* off, int32_t* npts, double* scalars, int32_t* status,
                                  int32_t* rescued, void* stream_) {
    if (n_f == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (n_f < 0 || n_xi < 32 || !idx || !pg || !rows_v || !rows_w || !rows_xi || !off || !npts || !scalars || !status ||
        !rescued || row_stride < gen_row_capacity(n_xi))
        return set_error(PTT_E_ARG, "ptt_generic_rescue: null pointer, bad size or row_stride too small");
    generic_rescue_kernel<<<(n_f + RESCUE_WARPS - 1) / RESCUE_WARPS, 32 * RESCUE_WARPS, 0, (cudaStream_t)stream_>>>(n_f, idx, pg, n_xi, t_end, thin_shell_limit, wn_rtol,
                                                                  rows_v, rows_w, rows_xi, (long long)row_stride, off,
                                                                  npts, scalars, status, rescued);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("generic_rescue_kernel", e);
    return PTT_OK;
}

extern "C" int ptt_shell_wall_values(int n, const double* rows_v, const double* rows_w, const double* rows_xi,
                                     int64_t row_stride, const int32_t* off, const int32_t* npts, const double* v_wall,
                                     const int32_t* sol_type, double* out, void* stream_) {
    if (n == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (n < 0 || !rows_v || !rows_w || !rows_xi || !off || !npts || !v_wall || !sol_type || !out)
        return set_error(PTT_E_ARG, "ptt_shell_wall_values: null pointer");
    if (n == 0) return PTT_OK;
    wall_values_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(n, rows_v, rows_w, rows_xi, row_stride, off,
                                                                            npts, v_wall, sol_type, out);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("wall_values_kernel", e);
    return PTT_OK;
}
