// TEST-ONLY host emulation of the device algorithm headers (pttools_b200/csrc/*.cuh compiled with g++).
// It exists so that the per-thread device logic can be unit-tested in a container without a GPU.
// The product package never loads this library; it only ever runs the nvcc-built CUDA kernels.
#include "../../pttools_b200/csrc/ptt_generic.cuh"

using namespace ptt;

namespace {
struct ArrayEmitter {
    double *v, *w, *xi;
    bool skip(const State&) const { return false; }
    bool visit(int i, const State& y) { v[i] = y.v; w[i] = y.w; xi[i] = y.xi; return true; }
};
}

extern "C" {

int emu_integrate(double v0, double w0, double xi0, double cs2, double t_end, int n,
                  double* v, double* w, double* xi) {
    State y0{v0, w0, xi0};
    ArrayEmitter em{v, w, xi};
    return integrate_grid(cs2, y0, t_end, n, em);
}

double emu_anmax(double v_wall, int* status) { return alpha_n_max_deflagration(v_wall, *status); }

double emu_anmax_machine(double v_wall, int* status) {
    BagMachine m;
    m.start_anmax(v_wall);
    m.run();
    return m.anmax(*status);
}

int emu_bag_forward(double v_wall, double ap, int sol_type, int n_xi, double* out) {
    BagForward f = bag_forward(v_wall, ap, sol_type, n_xi);
    out[0] = f.i1; out[1] = f.t_end2; out[2] = f.n_fwd; out[3] = f.last.v; out[4] = f.last.w; out[5] = f.last.xi;
    out[6] = f.w1; out[7] = f.wn_raw;
    return f.status;
}

int emu_bag_find_alpha_plus(double v_wall, double alpha_n, int n_xi, double anmax, double* ap, int* sol_type, int* iters) {
    BagRoot r = bag_find_alpha_plus(v_wall, alpha_n, n_xi, anmax);
    *ap = r.alpha_plus; *sol_type = r.sol_type; *iters = r.iters;
    return r.status;
}

// rows must hold profile_row_capacity(n_xi) doubles each
int emu_bag_profile(double v_wall, double ap, int sol_type, int n_xi, double w_n,
                    double* v, double* w, double* xi, int* off, int* npts, int* n_wall) {
    RowOut row{v, w, xi};
    BagProfile p = bag_emit_profile(v_wall, ap, sol_type, n_xi, w_n, row);
    *off = p.off; *npts = p.npts; *n_wall = p.n_wall;
    return p.status;
}

// BagMachine: root search + profile in one resumable machine (mode 0), or profile only (mode 1: ap_in, sol_type_in)
int emu_bag_machine(int mode, double v_wall, double alpha_n, int n_xi, double anmax, double ap_in, int sol_type_in,
                    double w_n, double* v, double* w, double* xi, int* off, int* npts, int* n_wall, double* ap,
                    int* sol_type, int* iters, double* xi_sh, double* wn_raw, int* root_status) {
    RowOut row{v, w, xi};
    BagMachine m;
    if (mode == 0) m.start_root(v_wall, alpha_n, n_xi, anmax, true, w_n, row);
    else m.start_profile(v_wall, ap_in, sol_type_in, n_xi, w_n, row);
    m.run();
    *ap = m.r.alpha_plus; *sol_type = m.r.sol_type; *iters = m.r.iters; *root_status = m.r.status;
    *off = m.p.off; *npts = m.p.npts; *n_wall = m.p.n_wall; *xi_sh = m.p.xi_sh; *wn_raw = m.p.wn_raw;
    return m.p.status;
}

int emu_row_capacity(int n_xi) { return profile_row_capacity(n_xi); }
int emu_gen_row_capacity(int n_xi) { return gen_row_capacity(n_xi); }

// eos[7] = {css2, csb2, mu_s, mu_b, V_s, V_b, bag_name}; guess[4] = {vp, vp_tilde, wp, wm}; scal[16] out
int emu_generic_solve(const double* eos, double v_wall, double alpha_n, int sol_type, double wn, double v_cj,
                      const double* guess, int n_xi, double t_end, int thin, double wn_rtol,
                      double* v, double* w, double* xi, int* off, int* npts, double* scal) {
    EosDev m{eos[0], eos[1], eos[2], eos[3], eos[4], eos[5], (int)eos[6]};
    GenericGuess g{guess[0], guess[1], guess[2], guess[3]};
    RowOut row{v, w, xi};
    GenericOut o = generic_solve_point(m, v_wall, alpha_n, sol_type, wn, v_cj, g, n_xi, t_end, thin, wn_rtol, row);
    *off = o.off; *npts = o.npts;
    double vals[16] = {o.vp, o.vm, o.vp_tilde, o.vm_tilde, o.v_sh, o.vm_sh, o.vm_tilde_sh, o.wp, o.wm, o.wm_sh, o.v_cj,
                       o.wn, o.wn_estimate, (double)o.solver_failed, 0, 0};
    for (int i = 0; i < 16; ++i) scal[i] = vals[i];
    return o.status;
}


// The retry search of generic_rescue_kernel, candidate by candidate in the reference's order (the kernel runs the
// candidates in the lanes of a warp and takes the first converged one whose predecessors have all failed: same choice).
// Returns the winning candidate (0-based) or -1; outputs as emu_generic_solve when a candidate won.
int emu_generic_rescue(const double* eos, double v_wall, double alpha_n, int sol_type, double wn, double v_cj,
                       const double* guess, int n_xi, double t_end, int thin, double wn_rtol,
                       double* v, double* w, double* xi, int* off, int* npts, double* scal, int* status) {
    (void)alpha_n;
    EosDev m{eos[0], eos[1], eos[2], eos[3], eos[4], eos[5], (int)eos[6]};
    RowOut row{v, w, xi};
    *status = -1;
    if (!(sol_type == SOL_SUB_DEF || sol_type == SOL_HYBRID)) return -1;
    for (int c = 0; c < GEN_RETRY_CANDIDATES; ++c) {
        GenericGuess g{guess[0], guess[1], guess[2], guess[3]};
        double x0;
        if (!generic_retry_start(m, v_wall, sol_type, sqrt(m.css2), g.wm, g.vp_tilde, eos[7], eos[8], c, x0)) continue;
        g.wm = x0;
        GenericMachine mach;
        mach.start(m, v_wall, sol_type, wn, v_cj, g, n_xi, t_end, thin, wn_rtol, row, true);
        while (mach.phase != GenericMachine::PH_DONE && mach.phase != GenericMachine::PH_HOLD) {
            const int rc = integrate_grid(mach.job_cs2, mach.job_y0, mach.job_te, n_xi, mach.vis);
            mach.advance(rc);
        }
        if (!(mach.phase == GenericMachine::PH_HOLD && mach.converged)) continue;
        mach.resume_final();
        while (mach.phase != GenericMachine::PH_DONE) {
            const int rc = integrate_grid(mach.job_cs2, mach.job_y0, mach.job_te, n_xi, mach.vis);
            mach.advance(rc);
        }
        const GenericOut o = mach.o;
        *status = o.status;
        *off = o.off; *npts = o.npts;
        double vals[16] = {o.vp, o.vm, o.vp_tilde, o.vm_tilde, o.v_sh, o.vm_sh, o.vm_tilde_sh, o.wp, o.wm, o.wm_sh, o.v_cj,
                           o.wn, o.wn_estimate, (double)o.solver_failed, (double)o.n_resid, (double)(c + 1)};
        for (int i = 0; i < 16; ++i) scal[i] = vals[i];
        return c;
    }
    return -1;
}

}
