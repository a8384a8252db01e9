// CUDA kernels for the fluid-shell solve: one (v_wall, alpha_n) point per thread, per-lane adaptive step control.
// K1 integrate_kernel         <- fluid_integrate_param          pttools/bubble/integrate.py:81-135
// K2 anmax / root / profile   <- alpha_n_max_deflagration_bag, find_alpha_plus_bag, sound_shell_alpha_plus
//                                pttools/bubble/alpha.py:44-50, 305-331; fluid_bag.py:83-222
#include "ptt_bag.cuh"
#include "ptt_host.cuh"

namespace ptt {

struct GridEmitter {
    double* v;
    double* w;
    double* xi;
    __device__ bool skip(const State&) const { return false; }
    __device__ bool visit(const int i, const State& y) {
        v[i] = y.v; w[i] = y.w; xi[i] = y.xi;
        return true;
    }
};

__global__ void __launch_bounds__(128) integrate_kernel(
    const int n, const double* __restrict__ v0, const double* __restrict__ w0, const double* __restrict__ xi0,
    const double* __restrict__ phase, const double* __restrict__ t_end, const int n_xi, const double css2,
    const double csb2, double* __restrict__ out_v, double* __restrict__ out_w, double* __restrict__ out_xi,
    int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ph = phase[i];
    if (ph < 0.0) ph = 0.0;    // integrate.py:106-108: an unset phase is taken as symmetric
    const double cs2 = ph * csb2 + (1.0 - ph) * css2;
    State y0;
    y0.v = v0[i]; y0.w = w0[i]; y0.xi = xi0[i];
    GridEmitter em;
    em.v = out_v + (size_t)i * n_xi;
    em.w = out_w + (size_t)i * n_xi;
    em.xi = out_xi + (size_t)i * n_xi;
    const int rc = integrate_grid(cs2, y0, t_end[i], n_xi, em);
    if (rc < 0) {
        // the reference raises RuntimeError("Integration failed"): NaN-fill, flag the lane
        for (int k = 0; k < n_xi; ++k) { em.v[k] = NAN; em.w[k] = NAN; em.xi[k] = NAN; }
        status[i] = ST_INTEGRATION_FAILED;
    } else {
        status[i] = ST_OK;
    }
}

__global__ void __launch_bounds__(64) bag_anmax_kernel(const int n, const double* __restrict__ v_wall,
                                                        double* __restrict__ out, int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double vw = v_wall[i];
    if (!(vw > 0.0 && vw < 1.0)) {
        out[i] = NAN;
        if (status) status[i] = ST_INVALID_INPUT;
        return;
    }
    int st;
    BagMachine mach;
    mach.start_anmax(vw);
    mach.run();
    out[i] = mach.anmax(st);
    if (status) status[i] = st;
}

__global__ void __launch_bounds__(64) bag_root_kernel(
    const int n, const double* __restrict__ v_wall, const double* __restrict__ alpha_n, const int n_xi,
    const double* __restrict__ an_max, double* __restrict__ alpha_plus, int* __restrict__ sol_type,
    int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double vw = v_wall[i], an = alpha_n[i];
    if (!(vw > 0.0 && vw < 1.0) || !(an > 0.0)) {
        alpha_plus[i] = NAN; sol_type[i] = SOL_ERROR; status[i] = ST_INVALID_INPUT;
        return;
    }
    double am;
    if (an_max) {
        am = an_max[i];
    } else {
        int st;
        am = alpha_n_max_deflagration(vw, st);
        if (st != ST_OK) { alpha_plus[i] = NAN; sol_type[i] = SOL_ERROR; status[i] = st; return; }
    }
    BagMachine mach;
    RowOut none;
    none.v = none.w = none.xi = nullptr;
    mach.start_root(vw, an, n_xi, am, false, 1.0, none);
    mach.run();
    const BagRoot& r = mach.r;
    alpha_plus[i] = (r.status == ST_OK) ? r.alpha_plus : NAN;
    sol_type[i] = r.sol_type;
    status[i] = r.status;
}

__global__ void __launch_bounds__(64) bag_profile_kernel(
    const int n, const double* __restrict__ v_wall, const double* __restrict__ alpha_plus,
    const int* __restrict__ sol_type, const int n_xi, const double* __restrict__ w_n, double* __restrict__ rows_v,
    double* __restrict__ rows_w, double* __restrict__ rows_xi, const long long row_stride, int* __restrict__ off,
    int* __restrict__ npts, double* __restrict__ scalars, int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RowOut row;
    row.v = rows_v + (size_t)i * row_stride;
    row.w = rows_w + (size_t)i * row_stride;
    row.xi = rows_xi + (size_t)i * row_stride;
    const int st_in = status[i];
    const int type = sol_type[i];
    const double ap = alpha_plus[i];
    if (st_in != ST_OK || type == SOL_ERROR || !(ap == ap)) {
        // reference: nan_arr of length 1 (fluid_bag.py:61-73)
        row.set(0, NAN, NAN, NAN);
        off[i] = 0; npts[i] = 1;
        if (scalars) { scalars[4 * i] = NAN; scalars[4 * i + 1] = NAN; scalars[4 * i + 2] = 0.0; scalars[4 * i + 3] = 0.0; }
        if (st_in == ST_OK) status[i] = ST_NO_SOLUTION;
        return;
    }
    BagMachine mach;
    mach.start_profile(v_wall[i], ap, type, n_xi, w_n ? w_n[i] : 1.0, row);
    mach.run();
    const BagProfile& p = mach.p;
    if (p.status != ST_OK) {
        row.set(0, NAN, NAN, NAN);
        off[i] = 0; npts[i] = 1;
        status[i] = p.status;
        if (scalars) { scalars[4 * i] = NAN; scalars[4 * i + 1] = NAN; scalars[4 * i + 2] = 0.0; scalars[4 * i + 3] = 0.0; }
        return;
    }
    off[i] = p.off;
    npts[i] = p.npts;
    if (scalars) {
        scalars[4 * i] = p.xi_sh;
        scalars[4 * i + 1] = p.wn_raw;
        scalars[4 * i + 2] = (double)p.n_wall;
        scalars[4 * i + 3] = 0.0;
    }
}

__global__ void set_status_kernel(const int n, int* status, const int value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) status[i] = value;
}

}  // namespace ptt

using namespace ptt;

extern "C" int ptt_profile_row_capacity(int n_xi) { return profile_row_capacity(n_xi); }

extern "C" int ptt_integrate_batch(int n, const double* v0, const double* w0, const double* xi0, const double* phase,
                                   const double* t_end, int n_xi, const ptt_eos_t* eos, double* out_v, double* out_w,
                                   double* out_xi, int32_t* status, int mem, void* stream_) {
    if (n == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (n < 0 || n_xi < 1 || !v0 || !w0 || !xi0 || !phase || !t_end || !eos || !out_v || !out_w || !out_xi || !status)
        return set_error(PTT_E_ARG, "ptt_integrate_batch: null pointer or bad size");
    if (n == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    Stager sg(mem, stream);
    const size_t nn = (size_t)n, no = (size_t)n * n_xi;
    const double* d_v0 = sg.in(v0, nn); const double* d_w0 = sg.in(w0, nn); const double* d_xi0 = sg.in(xi0, nn);
    const double* d_ph = sg.in(phase, nn); const double* d_te = sg.in(t_end, nn);
    double* d_ov = sg.out(out_v, no); double* d_ow = sg.out(out_w, no); double* d_ox = sg.out(out_xi, no);
    int* d_st = sg.out(status, nn);
    if (!sg.ok()) return sg.fail();
    integrate_kernel<<<(n + 127) / 128, 128, 0, stream>>>(n, d_v0, d_w0, d_xi0, d_ph, d_te, n_xi, eos->css2, eos->csb2,
                                                          d_ov, d_ow, d_ox, d_st);
    return sg.finish("integrate_kernel");
}

extern "C" int ptt_bag_alpha_n_max(int n, const double* v_wall, double* out, int32_t* status, int mem, void* stream_) {
    if (n == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (n < 0 || !v_wall || !out) return set_error(PTT_E_ARG, "ptt_bag_alpha_n_max: null pointer or bad size");
    if (n == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    Stager sg(mem, stream);
    const double* d_vw = sg.in(v_wall, (size_t)n);
    double* d_out = sg.out(out, (size_t)n);
    int* d_st = status ? sg.out(status, (size_t)n) : nullptr;
    if (!sg.ok()) return sg.fail();
    bag_anmax_kernel<<<(n + 63) / 64, 64, 0, stream>>>(n, d_vw, d_out, d_st);
    return sg.finish("bag_anmax_kernel");
}

extern "C" int ptt_bag_find_alpha_plus(int n, const double* v_wall, const double* alpha_n, int n_xi,
                                       const double* an_max, double* alpha_plus, int32_t* sol_type, int32_t* status,
                                       int mem, void* stream_) {
    if (n == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (n < 0 || n_xi < 3 || !v_wall || !alpha_n || !alpha_plus || !sol_type || !status)
        return set_error(PTT_E_ARG, "ptt_bag_find_alpha_plus: null pointer or bad size");
    if (n == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    Stager sg(mem, stream);
    const double* d_vw = sg.in(v_wall, (size_t)n);
    const double* d_an = sg.in(alpha_n, (size_t)n);
    const double* d_am = an_max ? sg.in(an_max, (size_t)n) : nullptr;
    double* d_ap = sg.out(alpha_plus, (size_t)n);
    int* d_ty = sg.out(sol_type, (size_t)n);
    int* d_st = sg.out(status, (size_t)n);
    if (!sg.ok()) return sg.fail();
    bag_root_kernel<<<(n + 63) / 64, 64, 0, stream>>>(n, d_vw, d_an, n_xi, d_am, d_ap, d_ty, d_st);
    return sg.finish("bag_root_kernel");
}

extern "C" int ptt_bag_profiles(int n, const double* v_wall, const double* alpha_plus, const int32_t* sol_type,
                                int n_xi, const double* w_n, double* rows_v, double* rows_w, double* rows_xi,
                                int64_t row_stride, int32_t* off, int32_t* npts, double* scalars, int32_t* status,
                                int mem, void* stream_) {
    if (n == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (n < 0 || n_xi < 3 || !v_wall || !alpha_plus || !sol_type || !rows_v || !rows_w || !rows_xi || !off || !npts ||
        !status || row_stride < profile_row_capacity(n_xi))
        return set_error(PTT_E_ARG, "ptt_bag_profiles: null pointer, bad size or row_stride too small");
    if (n == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    Stager sg(mem, stream);
    const size_t nn = (size_t)n, nr = (size_t)n * (size_t)row_stride;
    const double* d_vw = sg.in(v_wall, nn);
    const double* d_ap = sg.in(alpha_plus, nn);
    const int* d_ty = sg.in(sol_type, nn);
    const double* d_wn = w_n ? sg.in(w_n, nn) : nullptr;
    double* d_rv = sg.out(rows_v, nr); double* d_rw = sg.out(rows_w, nr); double* d_rx = sg.out(rows_xi, nr);
    int* d_off = sg.out(off, nn); int* d_np = sg.out(npts, nn);
    double* d_sc = scalars ? sg.out(scalars, 4 * nn) : nullptr;
    int* d_st = sg.inout(status, nn);
    if (!sg.ok()) return sg.fail();
    bag_profile_kernel<<<(n + 63) / 64, 64, 0, stream>>>(n, d_vw, d_ap, d_ty, n_xi, d_wn, d_rv, d_rw, d_rx,
                                                            (long long)row_stride, d_off, d_np, d_sc, d_st);
    return sg.finish("bag_profile_kernel");
}

extern "C" int ptt_sweep_shells_bag(int n, const double* v_wall, const double* alpha_n, int n_xi, const double* an_max,
                                    double* rows_v, double* rows_w, double* rows_xi, int64_t row_stride, int32_t* off,
                                    int32_t* npts, double* alpha_plus, int32_t* sol_type, double* scalars,
                                    int32_t* status, int mem, void* stream_) {
    if (n == 0) return PTT_OK;      // an empty batch is valid, whatever the pointers
    if (n < 0 || n_xi < 3 || !v_wall || !alpha_n || !rows_v || !rows_w || !rows_xi || !off || !npts || !alpha_plus ||
        !sol_type || !status || row_stride < profile_row_capacity(n_xi))
        return set_error(PTT_E_ARG, "ptt_sweep_shells_bag: null pointer, bad size or row_stride too small");
    if (n == 0) return PTT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    Stager sg(mem, stream);
    const size_t nn = (size_t)n, nr = (size_t)n * (size_t)row_stride;
    const double* d_vw = sg.in(v_wall, nn);
    const double* d_an = sg.in(alpha_n, nn);
    const double* d_am = an_max ? sg.in(an_max, nn) : nullptr;
    double* d_rv = sg.out(rows_v, nr); double* d_rw = sg.out(rows_w, nr); double* d_rx = sg.out(rows_xi, nr);
    int* d_off = sg.out(off, nn); int* d_np = sg.out(npts, nn);
    double* d_ap = sg.out(alpha_plus, nn);
    int* d_ty = sg.out(sol_type, nn);
    double* d_sc = scalars ? sg.out(scalars, 4 * nn) : nullptr;
    int* d_st = sg.out(status, nn);
    if (!sg.ok()) return sg.fail();
    const int grid = (n + 63) / 64;
    // two launches on purpose: in the root search the lanes visit one sample per step, in the materialisation every
    // tau-grid sample; fused into one kernel the lanes of a warp mix the two and the warp runs 11x slower (measured)
    bag_root_kernel<<<grid, 64, 0, stream>>>(n, d_vw, d_an, n_xi, d_am, d_ap, d_ty, d_st);
    bag_profile_kernel<<<grid, 64, 0, stream>>>(n, d_vw, d_ap, d_ty, n_xi, nullptr, d_rv, d_rw, d_rx,
                                                 (long long)row_stride, d_off, d_np, d_sc, d_st);
    return sg.finish("ptt_sweep_shells_bag");
}
