// Model-independent fluid shell solver (BagModel and ConstCSModel equations of state), device-side restatement of
//   sound_shell_generic and its solvers      pttools/bubble/fluid.py:48-1052
//   solve_junction                           pttools/bubble/boundary.py:232-373
//   find_shock_index / v_shock / solve_shock pttools/bubble/shock.py:49-248, 297-408
// One (model, v_wall, alpha_n) point per thread.  Host/device portable (see ptt_ode.cuh).
//
// Both in-scope models have p(w, phase) = w/mu_phase - V_phase with mu = 1 + 1/cs2, which makes the junction
// conditions  w1 g1^2 v1 = w2 g2^2 v2,  w1 g1^2 v1^2 + p1 = w2 g2^2 v2^2 + p2  a quadratic in v2:
//     A (mu2 - 1) v2^2 - mu2 (B + V2) v2 + A = 0,   A = w1 g1^2 v1,  B = A v1 + p1,   w2 = A (1 - v2^2) / v2.
// The reference solves them numerically (2-D MINPACK hybrd from a starting guess, boundary.py:360-373); here the
// closed form is used and the root is selected the way the reference's guess selects it.
#pragma once

#include "ptt_bag.cuh"

namespace ptt {

struct EosDev {
    double css2, csb2;   // sound speeds squared (symmetric / broken)
    double mu_s, mu_b;   // 1 + 1/cs2
    double V_s, V_b;     // potentials
    int bag_name;        // model.name == "bag": the reference then uses the closed-form bag shock curve (shock.py:63-70)
};

PTT_HD double eos_p(const EosDev& m, const double w, const int broken) {
    return broken ? w / m.mu_b - m.V_b : w / m.mu_s - m.V_s;
}

struct Junction {
    double v2, w2;
    bool ok;
};

// want: -1 = root below the sound speed of phase 2 (sqrt(1/(mu2-1))), +1 = root above it, 0 = root nearest v2_guess
PTT_HD Junction solve_junction(const EosDev& m, const double v1, const double w1, const int broken1, const int broken2,
                               const double v2_guess, const int want) {
    Junction j;
    j.ok = false; j.v2 = NAN; j.w2 = NAN;
    if (!(v1 > 0.0 && v1 < 1.0) || !(w1 > 0.0)) return j;
    const double mu2 = broken2 ? m.mu_b : m.mu_s;
    const double V2 = broken2 ? m.V_b : m.V_s;
    const double A = w1 * gamma2(v1) * v1;
    const double B = A * v1 + eos_p(m, w1, broken1);
    const double a = A * (mu2 - 1.0);
    const double b = -mu2 * (B + V2);
    const double c = A;
    const double disc = b * b - 4.0 * a * c;
    if (!(disc >= 0.0)) return j;
    const double q = -0.5 * (b + ((b >= 0.0) ? 1.0 : -1.0) * sqrt(disc));
    double r1 = q / a, r2 = c / q;
    if (r1 > r2) { const double t = r1; r1 = r2; r2 = t; }      // r1 <= cs2 <= r2
    double v2;
    if (want < 0) v2 = r1;
    else if (want > 0) v2 = r2;
    else v2 = (fabs(r1 - v2_guess) <= fabs(r2 - v2_guess)) ? r1 : r2;
    const double w2 = A * (1.0 - v2 * v2) / v2;
    if (!(v2 >= 0.0 && v2 <= 1.0) || !(w2 >= 0.0)) return j;    // boundary.py:329-337
    j.v2 = v2; j.w2 = w2; j.ok = true;
    return j;
}

// v_shock (shock.py:388-408): fluid speed just behind a shock at xi moving into still fluid of the symmetric phase.
// The same-phase junction has the trivial root v2 = v1 and the shock root v2 = cs^2 / v1.
PTT_HD double v_shock_generic(const EosDev& m, const double xi, const double cs_n) {
    if (xi <= cs_n || fabs(xi - cs_n) <= 1e-8 + 1e-5 * fabs(cs_n)) return 0.0;     // np.isclose(xi, cs_n)
    if (fabs(xi - 1.0) <= 1e-8 + 1e-5) return 1.0;
    const double vt = m.css2 / xi;
    const double r = lorentz(xi, vt);
    return (r < 0.0) ? 0.0 : r;
}

// First tau-grid index behind the shock curve (find_shock_index, shock.py:49-248), located in one ordered pass.
struct GenericShockLocator {
    EosDev m;
    double v_wall, cs_n;
    bool full;
    int found;          // i_shock (> 0) or 0 = not found
    int last_i;
    int first_close;    // first sample close to (cs_n, 0)
    bool turned;
    State cur, prev, prev2;

    PTT_HD void reset(const EosDev& m_, const double v_wall_, const double cs_n_, const bool full_) {
        m = m_; v_wall = v_wall_; cs_n = cs_n_; full = full_;
        found = 0; last_i = -1; first_close = -1; turned = false;
    }
    PTT_HD bool behind(const State& y) const {
        if (m.bag_name) return (y.xi > v_wall) && (y.v <= v_shock_bag(y.xi));          // shock.py:63-66
        if (!(y.xi > cs_n)) return false;
        return y.v <= v_shock_generic(m, y.xi, cs_n);
    }
    PTT_HD bool near_cs(const State& y) const {
        const double c = m.bag_name ? cs0() : cs_n;
        return fabs(y.xi - c) <= 1e-8 + 1e-5 * fabs(c) && fabs(y.v) <= 1e-8;
    }
    PTT_HD bool skip(const State& y_end) const {
        return !full && !behind(y_end) && !near_cs(y_end) && (last_i < 0 || y_end.xi > prev.xi);
    }
    PTT_HD bool visit(const int i, const State& y) {
        if (i > 0) {
            if (behind(y)) { found = i; cur = y; return false; }
            if (!m.bag_name) {
                if (first_close < 0 && near_cs(y)) first_close = i;
                if (y.xi > cs_n && first_close > 0) { found = first_close; cur = y; return false; }   // shock.py:114-121
                if (y.xi < prev.xi) {
                    // the curve turned back before reaching the shock curve: i_right = argmax(xi) (shock.py:134-141)
                    turned = true;
                    found = last_i;
                    cur = prev;
                    prev = prev2;
                    return false;
                }
            }
        }
        prev2 = prev;
        prev = y;
        last_i = i;
        return true;
    }
};

struct DeflagResult {
    int status;         // ST_OK or a failure code (the reference returns DEFLAGRATION_NAN)
    double vp, vp_tilde, wp;
    double t_end_final; // end time of the accepted integration (after the thin-shell loop)
    int i_shock;        // samples [0, i_shock) are kept
    State sh;           // sample i_shock - 1: (vm_sh, wm_sh, xi_sh)
    double wn_estimate;
};

// sound_shell_deflagration_common (fluid.py:121-288) in locate mode: everything but the profile arrays.
PTT_HD DeflagResult deflagration_common(const EosDev& m, const double v_wall, const double vm_tilde, const double wn,
                                        const double wm, const double cs_n, const double v_cj,
                                        const double vp_tilde_guess, const int sol_type, const int n_xi,
                                        const double t_end, const int thin_shell_limit) {
    DeflagResult r;
    r.status = ST_SOLVER_FAILED;
    r.vp = r.vp_tilde = r.wp = r.wn_estimate = NAN;
    r.i_shock = 0; r.t_end_final = -t_end;
    r.sh.v = r.sh.w = r.sh.xi = NAN;
    if (!(v_wall >= 0.0 && v_wall <= 1.0) || !(vm_tilde >= 0.0 && vm_tilde <= 1.0) || !(wn >= 0.0) || !(wm >= 0.0) ||
        !(vp_tilde_guess >= 0.0 && vp_tilde_guess <= 1.0) || v_wall > v_cj)
        return r;
    // wall junction, broken -> symmetric: a deflagration accelerates the flow through the wall, v+ < v-  (root below cs_s)
    const Junction j = solve_junction(m, vm_tilde, wm, 1, 0, vp_tilde_guess, -1);
    if (!j.ok) return r;
    r.vp_tilde = j.v2;
    r.wp = j.w2;
    r.vp = lorentz(v_wall, j.v2);        // -lorentz(vp_tilde, v_wall)
    if (!(r.vp >= 0.0) || !(r.wp >= 0.0)) return r;
    State y0;
    y0.v = r.vp; y0.w = r.wp; y0.xi = v_wall;
    GenericShockLocator loc;
    double te = -t_end;
    int i_shock = 0;
    State sh;
    for (int attempt = 0; attempt < 6; ++attempt) {     // first pass + up to 5 thin-shell refinements (fluid.py:223-271)
        loc.reset(m, v_wall, cs_n, false);
        const int rc = integrate_grid(m.css2, y0, te, n_xi, loc);
        if (rc < 0) { r.status = ST_INTEGRATION_FAILED; return r; }
        const int i_new = loc.found;
        if (attempt == 0) {
            if (i_new == 0 || i_new + 1 >= n_xi) return r;
        } else {
            if (i_new == 0 || i_new + 20 >= n_xi) break;           // keep the previous solution
        }
        i_shock = i_new;
        sh = loc.prev;
        r.t_end_final = te;
        if (i_shock >= thin_shell_limit) break;
        if (attempt == 5) break;
        te = grid_time(te, n_xi, i_shock + 20);
    }
    if (i_shock <= 1) return r;
    r.i_shock = i_shock;
    r.sh = sh;
    const double vm_tilde_sh = lorentz(sh.xi, sh.v);
    r.wn_estimate = sh.w * gamma2(vm_tilde_sh) * vm_tilde_sh / (gamma2(sh.xi) * sh.xi);   // w2_junction, boundary.py:638
    r.status = ST_OK;
    return r;
}

struct GenericGuess {
    double vp, vp_tilde, wp, wm;    // starting guesses (fluid reference table scaled by wn, fluid.py:925-949)
};

struct GenericRoot {
    int status;
    int sol_type;
    double wm;          // enthalpy just behind the wall (w_center for subsonic deflagrations)
    double vm_tilde;
    double vp_tilde_guess;
    bool solver_failed;
    int iters;
    DeflagResult d;     // last evaluation at wm
};

PTT_HD double generic_residual(const EosDev& m, const int sol_type, const double v_wall, const double wn,
                               const double wm, const double cs_n, const double v_cj, const double vp_tilde_guess,
                               const int n_xi, const double t_end, const int thin, DeflagResult& d) {
    if (!(wm >= 0.0)) { d.status = ST_SOLVER_FAILED; return NAN; }
    const double vm_tilde = (sol_type == SOL_HYBRID) ? sqrt(m.csb2) : v_wall;
    d = deflagration_common(m, v_wall, vm_tilde, wn, wm, cs_n, v_cj, vp_tilde_guess, sol_type, n_xi, t_end, thin);
    return d.wn_estimate - wn;
}

// sound_shell_solver_deflagration / _hybrid (fluid.py:606-669, 697-843): root of wn_estimate(wm) - wn.
// Subsonic deflagrations: scipy's secant (root_scalar, x0 = 0.99 wm_guess, x1 = 1.01 wm_guess, atol 1.48e-8),
// mirrored step by step.  Hybrids: the reference calls hybrd, which in one dimension is a finite-difference Newton
// step followed by secant (Broyden) updates; the same is done here.
PTT_HD GenericRoot generic_find_wm(const EosDev& m, const int sol_type, const double v_wall, const double wn,
                                   const double cs_n, const double v_cj, const GenericGuess& g, const int n_xi,
                                   const double t_end, const int thin, const double wn_rtol) {
    GenericRoot r;
    r.status = ST_OK; r.sol_type = sol_type; r.solver_failed = true; r.iters = 0;
    r.wm = NAN;
    r.vm_tilde = (sol_type == SOL_HYBRID) ? sqrt(m.csb2) : v_wall;
    double vp_guess = g.vp;
    double vpt_guess = g.vp_tilde;
    if (sol_type == SOL_SUB_DEF) {
        if (vp_guess > v_wall) vp_guess = 0.95 * v_wall;           // fluid.py:613-617
        vpt_guess = lorentz(v_wall, vp_guess);                     // -lorentz(vp_guess, v_wall), fluid.py:79
    }
    r.vp_tilde_guess = vpt_guess;
    DeflagResult d0, d1, d2;
    double p0, p1, q0, q1;
    bool converged = false;
    double p = NAN;
    // pass 0: the reference's primary solver; pass 1: its fallback (fsolve from wm_guess, fluid.py:629-640)
    for (int pass = 0; pass < 2 && !converged; ++pass) {
    const bool newton_start = (sol_type != SOL_SUB_DEF) || (pass == 1);
    if (!newton_start) {
        p0 = 0.99 * g.wm;
        p1 = 1.01 * g.wm;
        q0 = generic_residual(m, sol_type, v_wall, wn, p0, cs_n, v_cj, vpt_guess, n_xi, t_end, thin, d0);
        q1 = generic_residual(m, sol_type, v_wall, wn, p1, cs_n, v_cj, vpt_guess, n_xi, t_end, thin, d1);
    } else {
        p0 = g.wm;
        q0 = generic_residual(m, sol_type, v_wall, wn, p0, cs_n, v_cj, vpt_guess, n_xi, t_end, thin, d0);
        const double h = 1.4901161193847656e-08 * fabs(p0);
        double qh = generic_residual(m, sol_type, v_wall, wn, p0 + h, cs_n, v_cj, vpt_guess, n_xi, t_end, thin, d1);
        const double J = (qh - q0) / h;
        p1 = p0 - q0 / J;
        if (!(p1 > 0.0) || !(p1 == p1)) p1 = 1.01 * p0;
        // hybrd's first trust region is factor * |x| = 100 |x|: no clipping in practice
        q1 = generic_residual(m, sol_type, v_wall, wn, p1, cs_n, v_cj, vpt_guess, n_xi, t_end, thin, d1);
    }
    p = p1;
    if (q0 == q0 && q1 == q1) {
        if (fabs(q1) < fabs(q0)) {
            double t = p0; p0 = p1; p1 = t;
            t = q0; q0 = q1; q1 = t;
            const DeflagResult td = d0; d0 = d1; d1 = td;
        }
        const double atol = 1.48e-8;
        const double rtol_h = 1.4901161193847656e-08;     // hybrd xtol (relative) for the hybrid branch
        for (int it = 0; it < 50; ++it) {
            r.iters = it + 1;
            if (q1 == q0) {
                p = 0.5 * (p1 + p0);
                converged = (p1 == p0);
                break;
            }
            if (fabs(q1) > fabs(q0)) p = (-q0 / q1 * p1 + p0) / (1.0 - q0 / q1);
            else p = (-q1 / q0 * p0 + p1) / (1.0 - q1 / q0);
            const double tol = newton_start ? rtol_h * fabs(p) : atol;
            if (fabs(p - p1) <= tol) { converged = true; break; }
            p0 = p1; q0 = q1; d0 = d1;
            p1 = p;
            q1 = generic_residual(m, sol_type, v_wall, wn, p1, cs_n, v_cj, vpt_guess, n_xi, t_end, thin, d1);
            if (!(q1 == q1)) break;
        }
    }
    if (sol_type != SOL_SUB_DEF) break;
    }
    if (!(p == p) || !(p > 0.0)) { r.status = ST_ROOT_NOT_CONVERGED; return r; }
    // final evaluation at the root (fluid.py:642-653, 797-811)
    const double qf = generic_residual(m, sol_type, v_wall, wn, p, cs_n, v_cj, vpt_guess, n_xi, t_end, thin, d2);
    r.wm = p;
    r.d = d2;
    if (d2.status != ST_OK) { r.status = d2.status; return r; }
    const bool close = fabs(d2.wn_estimate - wn) <= 1e-8 + wn_rtol * fabs(wn);     // np.isclose(wn_estimate, wn, rtol)
    r.solver_failed = !(converged && close);
    (void)qf;
    return r;
}

// --- profile materialisation ----------------------------------------------------------------------------------
// Row layout for the generic solver: the part behind the wall (2 still-fluid samples + up to n_xi samples) ends at
// column gen_back_cap(n_xi); the wall -> shock part starts there.
PTT_HD int gen_back_cap(const int n_xi) { return n_xi + 2; }
PTT_HD int gen_row_capacity(const int n_xi) { return 2 * n_xi + 4; }

struct OffsetEmitter {
    RowOut row;
    int base;       // column of sample 0
    int dir;        // +1 forward, -1 reversed
    int limit;      // stop after this many samples (forward part: i_shock)
    PTT_HD bool skip(const State&) const { return false; }
    PTT_HD bool visit(const int i, const State& y) {
        row.set(base + dir * i, y.v, y.w, y.xi);
        return i + 1 < limit;
    }
};

struct CsTrimEmitter {      // trim_fluid_wall_to_cs for detonations (trim.py:20-73, n_start = 0)
    RowOut row;
    int base;
    double cs2b;
    int hit;
    int last_i;
    PTT_HD bool skip(const State&) const { return false; }
    PTT_HD bool visit(const int i, const State& y) {
        last_i = i;
        row.set(base - i, y.v, y.w, y.xi);
        if ((y.v <= 0.0) || (y.xi * y.xi <= cs2b)) { hit = i; return false; }
        return true;
    }
};

struct GenericOut {
    int status;
    int off, npts;
    int solver_failed;
    // the reference's 17-tuple scalars (fluid.py:1052)
    double vp, vm, vp_tilde, vm_tilde, v_sh, vm_sh, vm_tilde_sh, wp, wm, wm_sh, v_cj, wn, wn_estimate;
};

// sound_shell_generic (fluid.py:848-1052) for one point, default arguments (no user guesses, no bag/Giese solver).
// vm_tilde_ref / wm_ref: reference-table guesses for detonations (unscaled, fluid.py:969-973).
PTT_HD GenericOut generic_solve_point(const EosDev& m, const double v_wall, const double alpha_n, const int sol_type,
                                      const double wn, const double v_cj, const GenericGuess& g, const int n_xi,
                                      const double t_end, const int thin, const double wn_rtol, const RowOut& row) {
    GenericOut o;
    o.status = ST_OK; o.off = 0; o.npts = 0; o.solver_failed = 1;
    o.vp = o.vm = o.vp_tilde = o.vm_tilde = o.v_sh = o.vm_sh = o.vm_tilde_sh = o.wp = o.wm = o.wm_sh = NAN;
    o.v_cj = v_cj; o.wn = wn; o.wn_estimate = NAN;
    const double cs_n = sqrt(m.css2);
    const double dxi = 1.0 / (double)n_xi;
    const int bc = gen_back_cap(n_xi);
    int nb = 0, nf = 0;             // samples behind / ahead of the wall (without the still-fluid points)
    double w_first;                 // w of the first (innermost) moving sample

    if (sol_type == SOL_DETON) {
        // fluid.py:351-478.  Weak detonation: exit speed above the sound speed of the broken phase (root >= cs_b).
        const Junction j = solve_junction(m, v_wall, wn, 0, 1, 0.0, +1);
        if (!j.ok) { o.status = ST_SOLVER_FAILED; return o; }
        const double vm_tilde = j.v2, wm = j.w2;
        const double vm = lorentz(v_wall, vm_tilde);
        State y0;
        y0.v = vm; y0.w = wm; y0.xi = v_wall;
        CsTrimEmitter em;
        em.row = row; em.base = bc - 1; em.cs2b = m.csb2; em.hit = -1; em.last_i = -1;
        const int rc = integrate_grid(m.csb2, y0, -t_end, n_xi, em);
        if (rc < 0) { o.status = ST_INTEGRATION_FAILED; return o; }
        nb = (em.hit < 0) ? n_xi - 2 : ((em.hit == 0) ? 1 : em.hit);
        if (nb < 1) { o.status = ST_INTEGRATION_FAILED; return o; }
        o.vp = 0.0; o.vm = vm; o.vp_tilde = v_wall; o.vm_tilde = vm_tilde; o.v_sh = v_wall; o.vm_sh = vm;
        o.vm_tilde_sh = vm_tilde; o.wp = wn; o.wm = wm; o.wm_sh = wm;
        o.solver_failed = 0;
        nf = 0;
        w_first = row.w[bc - nb];
        row.set(bc, 0.0, wn, v_wall + dxi);
        row.set(bc + 1, 0.0, wn, 1.0);
    } else {
        const GenericRoot r = generic_find_wm(m, sol_type, v_wall, wn, cs_n, v_cj, g, n_xi, t_end, thin, wn_rtol);
        if (r.status != ST_OK) { o.status = r.status; return o; }
        const DeflagResult& d = r.d;
        // wall -> shock samples [0, i_shock) of the accepted integration
        State y0;
        y0.v = d.vp; y0.w = d.wp; y0.xi = v_wall;
        OffsetEmitter fe;
        fe.row = row; fe.base = bc; fe.dir = 1; fe.limit = d.i_shock;
        int rc = integrate_grid(m.css2, y0, d.t_end_final, n_xi, fe);
        if (rc < 0) { o.status = ST_INTEGRATION_FAILED; return o; }
        nf = d.i_shock;
        const double xi_last = row.xi[bc + nf - 1];
        row.set(bc + nf, 0.0, wn, xi_last + dxi);
        row.set(bc + nf + 1, 0.0, wn, 1.0);
        o.vp = d.vp; o.vp_tilde = d.vp_tilde; o.vm_tilde = r.vm_tilde; o.vm = lorentz(r.vm_tilde, v_wall);
        o.v_sh = d.sh.xi; o.vm_sh = d.sh.v; o.vm_tilde_sh = lorentz(d.sh.xi, d.sh.v);
        o.wp = d.wp; o.wm = r.wm; o.wm_sh = d.sh.w; o.wn_estimate = d.wn_estimate;
        o.solver_failed = r.solver_failed ? 1 : 0;
        if (sol_type == SOL_HYBRID) {
            // rarefaction tail behind the wall, n_xi samples, NOT trimmed (fluid.py:831-841)
            o.vm = lorentz(v_wall, sqrt(m.csb2));
            State yb;
            yb.v = o.vm; yb.w = r.wm; yb.xi = v_wall;
            OffsetEmitter be;
            be.row = row; be.base = bc - 1; be.dir = -1; be.limit = n_xi;
            rc = integrate_grid(m.csb2, yb, -t_end, n_xi, be);
            if (rc < 0) { o.status = ST_INTEGRATION_FAILED; return o; }
            nb = n_xi;
        } else {
            nb = 0;
        }
        w_first = (nb > 0) ? row.w[bc - nb] : row.w[bc];
    }
    const double xi_first = (nb > 0) ? row.xi[bc - nb] : row.xi[bc];
    const double w_center = fmin(o.wm, w_first);                   // fluid.py:1039
    row.set(bc - nb - 1, 0.0, w_center, xi_first - dxi);
    row.set(bc - nb - 2, 0.0, w_center, 0.0);
    o.off = bc - nb - 2;
    o.npts = 2 + nb + nf + 2;
    return o;
}

}  // namespace ptt
