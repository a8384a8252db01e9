// Bag-model fluid shell: device-side restatement of the reference's old solver
//   sound_shell_bag            pttools/bubble/fluid_bag.py:33-79
//   sound_shell_alpha_plus     fluid_bag.py:83-222
//   find_alpha_plus_bag        pttools/bubble/alpha.py:305-331 (MINPACK hybrd there; safeguarded secant here)
//   alpha_n_max_deflagration   alpha.py:44-50
//   trims / zoom               trim.py:20-116, shock.py:266-294
//   wall and shock relations   boundary.py:89-158, 376-408, 504-518; shock.py:423-430
// One (v_wall, alpha_n) point per thread.  Host/device portable (see ptt_ode.cuh).
#pragma once

#include "ptt_ode.cuh"

namespace ptt {

enum SolType : int { SOL_SUB_DEF = 0, SOL_HYBRID = 1, SOL_DETON = 2, SOL_ERROR = -1 };

// Per-point status codes of the C ABI (include/pttools_b200.h)
enum Status : int {
    ST_OK = 0,
    ST_NO_SOLUTION = 1,        // alpha_n >= alpha_n_max_deflagration(v_wall): the reference returns nan_arr
    ST_INTEGRATION_FAILED = 2, // reference: RuntimeError("Integration failed")
    ST_ROOT_NOT_CONVERGED = 3,
    ST_INVALID_INPUT = 4,
    ST_SOLVER_FAILED = 5       // generic solver: solution outside wn_rtol (fluid.py:650)
};

constexpr int N_XI_DEFAULT = 5000;       // bubble/const.py:12
constexpr double T_END_DEFAULT = 50.0;   // bubble/const.py:18

PTT_HD double cs0() { return 1.0 / sqrt(3.0); }          // const.py:29 (1/np.sqrt(3))
PTT_HD double cs0_root() { return sqrt(1.0 / 3.0); }     // cs2_fun(w, 0)**0.5 in fluid_bag.py:147,202

PTT_HD double gamma2(const double v) { return 1.0 / (1.0 - v * v); }
PTT_HD double lorentz(const double xi, const double v) { return (xi - v) / (1.0 - v * xi); }
PTT_HD double enthalpy_ratio(const double v_m, const double v_p) {
    return gamma2(v_m) * v_m / (gamma2(v_p) * v_p);
}

PTT_HD double v_plus_bag(const double vm, const double ap, const int sol_type) {
    const double x = vm + 1.0 / (3.0 * vm);
    const double b = (sol_type == SOL_DETON) ? 1.0 : -1.0;
    if (b < 0.0 && ap > 1.0 / 3.0) return NAN;
    return (0.5 / (1.0 + ap)) * (x + b * sqrt(x * x + 4.0 * ap * ap + (8.0 / 3.0) * ap - (4.0 / 3.0)));
}

PTT_HD double v_minus_bag(const double vp, const double ap, const int sol_type) {
    if (vp < 0.0) return NAN;
    const double vp2 = vp * vp;
    const double y = vp2 + 1.0 / 3.0;
    const double z = y - ap * (1.0 - vp2);
    const double x = (4.0 / 3.0) * vp2;
    const double b = (sol_type == SOL_DETON) ? 1.0 : -1.0;
    return (0.5 / vp) * (z + b * sqrt(z * z - x));
}

PTT_HD double v_shock_bag(const double xi) {
    if (xi < cs0()) return NAN;
    return (3.0 * xi * xi - 1.0) / (2.0 * xi);
}

PTT_HD double alpha_plus_max_detonation(const double v_wall) {
    if (v_wall < cs0()) return 0.0;
    const double b = 1.0 - sqrt(3.0) * v_wall;
    return b * b / (3.0 * (1.0 - v_wall * v_wall));
}

PTT_HD double alpha_plus_min_hybrid(const double v_wall) {
    if (v_wall < cs0()) return 0.0;
    const double b = 1.0 - sqrt(3.0) * v_wall;
    return b * b / (9.0 * v_wall * v_wall - 1.0);
}

struct WallSpeeds {
    double vfp_w, vfm_w, vfp_p, vfm_p;
};

PTT_HD WallSpeeds fluid_speeds_at_wall(const double v_wall, const double alpha_p, const int sol_type) {
    WallSpeeds s;
    if (sol_type == SOL_SUB_DEF) {
        s.vfp_w = v_plus_bag(v_wall, alpha_p, sol_type);
        s.vfm_w = v_wall;
    } else if (sol_type == SOL_HYBRID) {
        s.vfp_w = v_plus_bag(cs0(), alpha_p, sol_type);
        s.vfm_w = cs0();
    } else {
        s.vfp_w = v_wall;
        s.vfm_w = v_minus_bag(v_wall, alpha_p, SOL_DETON);
    }
    s.vfp_p = lorentz(v_wall, s.vfp_w);
    s.vfm_p = lorentz(v_wall, s.vfm_w);
    return s;
}

// --- wall -> shock ---------------------------------------------------------------------------------------------

// Finds the first tau-grid index with v <= v_sh(xi) (trim_fluid_wall_to_shock, trim.py:94-99).
struct ShockLocator {
    bool full;          // visit every sample (needed to reproduce the "never found" slice semantics)
    bool zero_hit;      // sample 0 already satisfies the condition: the reference promotes index 0 to 1
    int found;          // n_shock (>= 1) or -1
    int last_i;
    State cur, prev, prev2, y1;

    PTT_HD void reset(const bool full_) {
        full = full_;
        zero_hit = false;
        found = -1;
        last_i = -1;
    }
    PTT_HD static bool crossed(const State& y) { return y.v <= v_shock_bag(y.xi); }   // NaN compares false
    PTT_HD bool skip(const State& y_end) const { return !full && !zero_hit && !crossed(y_end); }
    // The crossing inside one integrator step by bisection (see integrate_grid): on thin shells one step holds thousands
    // of tau-grid samples.  Same assumption as skip(): within a step the curve crosses the shock curve once.
    template <class Sampler>
    PTT_HD int search(const int i_lo, const int i_hi, const Sampler& sample) {
        if (full || zero_hit || last_i != i_lo - 1) return -1;
        const State y_hi = sample(i_hi);
        if (!crossed(y_hi)) return (i_lo > 1) ? -2 : -1;
        if (crossed(prev)) return -1;
        int lo = i_lo - 1, hi = i_hi;
        State y_found = y_hi;
        while (hi - lo > 1) {
            const int mid = lo + ((hi - lo) >> 1);
            const State y = sample(mid);
            if (crossed(y)) { hi = mid; y_found = y; } else { lo = mid; }
        }
        if (i_lo == 1) y1 = (hi == 1) ? y_found : sample(1);
        if (hi - 1 >= i_lo) {
            const State p1 = sample(hi - 1);
            prev2 = (hi - 2 >= i_lo) ? sample(hi - 2) : prev;
            prev = p1;
        }
        found = hi; cur = y_found; last_i = hi - 1;
        return hi;
    }
    PTT_HD bool visit(const int i, const State& y) {
        if (i == 1) y1 = y;
        if (i == 0) {
            zero_hit = crossed(y);
        } else if (zero_hit || crossed(y)) {
            found = zero_hit ? 1 : i;
            cur = y;
            return false;
        }
        prev2 = prev;
        prev = y;
        last_i = i;
        return true;
    }
};

struct BagForward {
    int status;
    int i1;             // trim index of the first (5000-point) pass
    double t_end2;      // = t[i1] of the first pass: end time of the refined pass (fluid_bag.py:166)
    int n_fwd;          // number of wall->shock samples kept after the second trim
    State last;         // last kept sample (after shock_zoom_last_element)
    double w1;          // w of the second kept sample (index n_wall+1 of the assembled shell)
    double wn_raw;      // enthalpy ahead of the shock in the wp=1 normalisation (fluid_bag.py:175-177)
};

// Runs one wall->shock pass in locate mode; out-params describe the trimmed array v[:n+1].
PTT_HD int forward_pass(const double cs2, const State& y0, const double t_end, const int n,
                        int& n_keep, State& last, State& prev, State& y1) {
    ShockLocator loc;
    loc.reset(false);
    int rc = integrate_grid(cs2, y0, t_end, n, loc);
    if (rc < 0) return ST_INTEGRATION_FAILED;
    if (loc.found < 0) {
        // Never crossed: the reference's index stays -2 and the slice [: -1] silently drops the last sample.
        loc.reset(true);
        rc = integrate_grid(cs2, y0, t_end, n, loc);
        if (rc < 0) return ST_INTEGRATION_FAILED;
    }
    y1 = loc.y1;
    if (loc.found >= 0) {
        n_keep = loc.found + 1;
        last = loc.cur;
        prev = loc.prev;
    } else {
        n_keep = n - 1;
        last = loc.prev2;      // sample n-2
        prev = loc.prev2;      // (sample n-3 is not needed: no zoom can apply, see below)
        if (n_keep < 1) return ST_INTEGRATION_FAILED;
    }
    return ST_OK;
}

PTT_HD void shock_zoom(const State& p, State& l) {
    // shock.py:281-293
    const double vs1 = v_shock_bag(l.xi), vs0 = v_shock_bag(p.xi);
    if (l.v < vs1 && p.v > vs0 && l.xi > p.xi) {
        const double dxi = l.xi - p.xi;
        const double dv = l.v - p.v;
        const double dv_sh = vs1 - vs0;
        const double dw_sh = l.w - p.w;
        const double dxi_sh = dxi * (p.v - vs0) / (dv_sh - dv);
        l.xi = p.xi + dxi_sh;
        l.v = p.v + (dv_sh / dxi) * dxi_sh;
        l.w = p.w + (dw_sh / dxi) * dxi_sh;
    }
}

// Both wall->shock passes of sound_shell_alpha_plus (fluid_bag.py:158-181) without materialising the profile.
PTT_HD BagForward bag_forward(const double v_wall, const double alpha_plus, const int sol_type, const int n_xi) {
    BagForward r;
    r.status = ST_OK;
    r.i1 = 0; r.t_end2 = 0.0; r.n_fwd = 0; r.w1 = NAN; r.wn_raw = NAN;
    r.last.v = r.last.w = r.last.xi = NAN;
    const WallSpeeds ws = fluid_speeds_at_wall(v_wall, alpha_plus, sol_type);
    if (!(ws.vfp_p == ws.vfp_p)) { r.status = ST_INVALID_INPUT; return r; }
    State y0;
    y0.v = ws.vfp_p; y0.w = 1.0; y0.xi = v_wall;
    const double cs2 = 1.0 / 3.0;
    int n1;
    State last, prev, y1;
    r.status = forward_pass(cs2, y0, -T_END_DEFAULT, N_XI_DEFAULT, n1, last, prev, y1);
    if (r.status != ST_OK) return r;
    r.i1 = n1 - 1;
    r.t_end2 = grid_time(-T_END_DEFAULT, N_XI_DEFAULT, r.i1);
    int n2;
    r.status = forward_pass(cs2, y0, r.t_end2, n_xi, n2, last, prev, y1);
    if (r.status != ST_OK) return r;
    if (n2 >= 2) shock_zoom(prev, last);
    r.n_fwd = n2;
    r.last = last;
    r.w1 = (n2 == 2) ? last.w : y1.w;
    const double vfp_s = last.xi;
    const double vfm_s = 1.0 / (3.0 * vfp_s);
    r.wn_raw = last.w * enthalpy_ratio(vfm_s, vfp_s);
    return r;
}

// alpha_n_max_deflagration_bag(v_wall) = w[n_wall+1]/3 of the alpha_+ = 1/3 - 1e-10 shell (alpha.py:44-50).
// Always evaluated on the default 5000-point grid, whatever n_xi the caller uses (alpha.py:316, transition.py:27).
PTT_HD double alpha_n_max_deflagration(const double v_wall, int& status) {
    const int sol_type = (v_wall > cs0()) ? SOL_HYBRID : SOL_SUB_DEF;
    const BagForward f = bag_forward(v_wall, 1.0 / 3.0 - 1.0e-10, sol_type, N_XI_DEFAULT);
    status = f.status;
    if (f.status != ST_OK) return NAN;
    return (f.w1 / f.wn_raw) * (1.0 / 3.0);
}

PTT_HD int identify_solution_type_bag(const double v_wall, const double alpha_n, const double an_max_defl) {
    // transition.py:19-44
    if (alpha_n < alpha_plus_max_detonation(v_wall)) return SOL_DETON;
    if (alpha_n < an_max_defl) return (v_wall <= cs0()) ? SOL_SUB_DEF : SOL_HYBRID;
    return SOL_ERROR;
}

// alpha_n(alpha_+) = alpha_+ * w[n_wall]/w[-1]  (alpha.py:229-255); w[n_wall] is the wp=1 wall sample.
PTT_HD double bag_alpha_n_of_alpha_plus(const double v_wall, const double ap, const int sol_type, const int n_xi,
                                       int& status) {
    const BagForward f = bag_forward(v_wall, ap, sol_type, n_xi);
    status = f.status;
    return ap / f.wn_raw;
}

struct BagRoot {
    int status;
    int sol_type;
    double alpha_plus;
    int iters;
};

// find_alpha_plus_bag (alpha.py:305-331).  The reference calls hybrd (xtol=1e-6, factor=0.1) from
// alpha_plus_initial_guess (alpha.py:168-188); measured on the reference, hybrd's final iterate is within 3e-10
// (relative) of the exact root, so the device solver simply converges: bracketed secant on (0, 1/3).
PTT_HD BagRoot bag_find_alpha_plus(const double v_wall, const double alpha_n, const int n_xi, const double an_max_defl) {
    BagRoot r;
    r.iters = 0;
    r.alpha_plus = NAN;
    r.sol_type = identify_solution_type_bag(v_wall, alpha_n, an_max_defl);
    if (r.sol_type == SOL_ERROR) { r.status = ST_NO_SOLUTION; return r; }
    if (r.sol_type == SOL_DETON) { r.status = ST_OK; r.alpha_plus = alpha_n; return r; }

    double guess;
    if (alpha_n < 0.05) {
        guess = alpha_n;
    } else {
        const double ap_min = alpha_plus_min_hybrid(v_wall);
        const double an_min = alpha_plus_max_detonation(v_wall);
        const double slope = (1.0 / 3.0 - ap_min) / (an_max_defl - an_min);
        guess = ap_min + slope * (alpha_n - an_min);
    }
    const double hi_lim = 1.0 / 3.0 - 1.0e-10;    // the reference's own upper probe (alpha.py:47)
    double lo = 0.0, hi = hi_lim;                  // residual < 0 at lo, > 0 at hi (alpha_n < an_max_defl)
    if (guess <= lo) guess = 1e-3 * alpha_n;
    if (guess >= hi) guess = 0.5 * (lo + hi);

    int st = ST_OK;
    double x0 = guess;
    double f0 = bag_alpha_n_of_alpha_plus(v_wall, x0, r.sol_type, n_xi, st) - alpha_n;
    if (st != ST_OK || !(f0 == f0)) { r.status = (st != ST_OK) ? st : ST_ROOT_NOT_CONVERGED; return r; }
    if (f0 < 0.0) lo = x0; else hi = x0;
    double x1 = x0 * ((f0 < 0.0) ? 1.001 : 0.999);
    if (x1 <= lo || x1 >= hi) x1 = 0.5 * (lo + hi);
    double f1 = bag_alpha_n_of_alpha_plus(v_wall, x1, r.sol_type, n_xi, st) - alpha_n;
    if (st != ST_OK || !(f1 == f1)) { r.status = (st != ST_OK) ? st : ST_ROOT_NOT_CONVERGED; return r; }
    if (f1 < 0.0) { if (x1 > lo) lo = x1; } else { if (x1 < hi) hi = x1; }

    r.status = ST_ROOT_NOT_CONVERGED;
    for (int it = 0; it < 60; ++it) {
        r.iters = it + 1;
        if (f1 == 0.0) { r.status = ST_OK; break; }
        double x2 = (f1 != f0) ? x1 - f1 * (x1 - x0) / (f1 - f0) : 0.5 * (lo + hi);
        if (!(x2 > lo && x2 < hi)) x2 = 0.5 * (lo + hi);
        const double f2 = bag_alpha_n_of_alpha_plus(v_wall, x2, r.sol_type, n_xi, st) - alpha_n;
        if (st != ST_OK || !(f2 == f2)) { r.status = (st != ST_OK) ? st : ST_ROOT_NOT_CONVERGED; return r; }
        if (f2 < 0.0) { if (x2 > lo) lo = x2; } else { if (x2 < hi) hi = x2; }
        const double dx = fabs(x2 - x1);
        x0 = x1; f0 = f1;
        x1 = x2; f1 = f2;
        if (dx <= 1e-13 * fabs(x2) || fabs(f2) <= 1e-14 * alpha_n || (hi - lo) <= 1e-14 * hi) { r.status = ST_OK; break; }
        if (dx <= 1e-10 * fabs(x2) && fabs(f2) <= 1e-11 * alpha_n) { r.status = ST_OK; break; }
    }
    r.alpha_plus = x1;
    return r;
}

// --- profile materialisation --------------------------------------------------------------------------------

// A profile row has room for  2 + N_XI_DEFAULT  samples behind the wall (two still-fluid points, then the
// reversed wall->cs integration) followed by  n_xi + 2  samples ahead of it.  The wall->shock part always starts
// at column ROW_BACK_CAP, so the assembled shell is the contiguous slice [off, off + npts) of the row.
constexpr int ROW_BACK_CAP = N_XI_DEFAULT + 2;
PTT_HD int profile_row_capacity(const int n_xi) { return ROW_BACK_CAP + n_xi + 2; }

struct RowOut {
    double* v;
    double* w;
    double* xi;
    PTT_HD void set(const int k, const double v_, const double w_, const double xi_) const {
        v[k] = v_; w[k] = w_; xi[k] = xi_;
    }
};

struct ForwardEmitter {
    RowOut row;
    bool zero_hit;
    int found;
    int last_i;
    PTT_HD bool skip(const State&) const { return false; }
    PTT_HD bool visit(const int i, const State& y) {
        row.set(ROW_BACK_CAP + i, y.v, y.w, y.xi);
        last_i = i;
        if (i == 0) {
            zero_hit = ShockLocator::crossed(y);
            return true;
        }
        if (zero_hit || ShockLocator::crossed(y)) {
            found = zero_hit ? 1 : i;
            return false;
        }
        return true;
    }
};

struct BackwardEmitter {
    RowOut row;
    double cs2b;
    int n_start;    // 1 for hybrids (the wall sample is skipped, trim.py:69-71), 0 for detonations
    int hit;        // first index with v <= 0 or xi^2 <= cs2 (trim.py:52-56), or -1
    int last_i;
    PTT_HD bool skip(const State&) const { return false; }
    PTT_HD bool visit(const int i, const State& y) {
        last_i = i;
        const bool stop = (y.v <= 0.0) || (y.xi * y.xi <= cs2b);
        if (i >= n_start) row.set(ROW_BACK_CAP - 1 - (i - n_start), y.v, y.w, y.xi);
        if (stop && hit < 0) {
            hit = i;
            // hybrids keep one more sample than the hit index when the hit is sample 0 (n_stop = 2)
            if (!(i == 0 && n_start == 1)) return false;
        }
        if (hit == 0 && i >= 1) return false;
        return true;
    }
};

struct BagProfile {
    int status;
    int off;        // first column of the assembled shell inside the row
    int npts;       // number of samples
    int n_wall;     // index (relative to off) of the first sample with xi >= v_wall (props.find_v_index)
    double xi_sh;   // shock position (v_wall for detonations)
    double wn_raw;  // w[-1] before the final rescaling
};

// sound_shell_alpha_plus (fluid_bag.py:83-222): writes v, w, xi of the shell into `row`.
PTT_HD BagProfile bag_emit_profile(const double v_wall, const double alpha_plus, const int sol_type, const int n_xi,
                                   const double w_n, const RowOut& row) {
    BagProfile p;
    p.status = ST_OK; p.off = 0; p.npts = 0; p.n_wall = 0; p.xi_sh = v_wall; p.wn_raw = NAN;
    const WallSpeeds ws = fluid_speeds_at_wall(v_wall, alpha_plus, sol_type);
    const double wp = 1.0;
    const double wm = wp / enthalpy_ratio(ws.vfm_w, ws.vfp_w);
    const double dxi = 1.0 / (double)n_xi;
    const double cs2 = 1.0 / 3.0;

    int nf = 0;                          // samples ahead of the wall, excluding the two still-fluid points
    double xif0 = v_wall + dxi, wf = wp;
    if (sol_type != SOL_DETON) {
        const BagForward f = bag_forward(v_wall, alpha_plus, sol_type, n_xi);   // locates i1 -> t_end2
        if (f.status != ST_OK) { p.status = f.status; return p; }
        State y0;
        y0.v = ws.vfp_p; y0.w = wp; y0.xi = v_wall;
        ForwardEmitter em;
        em.row = row; em.zero_hit = false; em.found = -1; em.last_i = -1;
        const int rc = integrate_grid(cs2, y0, f.t_end2, n_xi, em);
        if (rc < 0) { p.status = ST_INTEGRATION_FAILED; return p; }
        nf = (em.found >= 0) ? em.found + 1 : n_xi - 1;
        if (nf >= 2) {
            State l, q;
            const int kl = ROW_BACK_CAP + nf - 1;
            l.v = row.v[kl]; l.w = row.w[kl]; l.xi = row.xi[kl];
            q.v = row.v[kl - 1]; q.w = row.w[kl - 1]; q.xi = row.xi[kl - 1];
            shock_zoom(q, l);
            row.set(kl, l.v, l.w, l.xi);
        }
        const int kl = ROW_BACK_CAP + nf - 1;
        const double vfp_s = row.xi[kl];
        const double vfm_s = 1.0 / (3.0 * vfp_s);
        wf = row.w[kl] * enthalpy_ratio(vfm_s, vfp_s);
        xif0 = vfp_s;
        p.xi_sh = vfp_s;
    }
    row.set(ROW_BACK_CAP + nf, 0.0, wf, xif0);
    row.set(ROW_BACK_CAP + nf + 1, 0.0, wf, 1.0);

    int nb = 0;
    double xib0 = fmin(cs0_root(), v_wall) - dxi, wb = wm;
    if (sol_type != SOL_SUB_DEF) {
        State y0;
        y0.v = ws.vfm_p; y0.w = wm; y0.xi = v_wall;
        BackwardEmitter em;
        em.row = row; em.cs2b = cs2; em.hit = -1; em.last_i = -1;
        em.n_start = (sol_type == SOL_DETON) ? 0 : 1;
        const int rc = integrate_grid(cs2, y0, -T_END_DEFAULT, N_XI_DEFAULT, em);
        if (rc < 0) { p.status = ST_INTEGRATION_FAILED; return p; }
        int n_stop;
        if (em.hit < 0) n_stop = N_XI_DEFAULT - 2 + em.n_start;       // slice [: -2] (+1 for hybrids)
        else n_stop = ((em.hit == 0) ? 1 : em.hit) + em.n_start;
        nb = n_stop - em.n_start;
        if (nb < 1) { p.status = ST_INTEGRATION_FAILED; return p; }
        wb = row.w[ROW_BACK_CAP - nb];        // last kept sample of the wall->cs integration
        xib0 = cs0_root();
    }
    row.set(ROW_BACK_CAP - nb - 1, 0.0, wb, xib0);
    row.set(ROW_BACK_CAP - nb - 2, 0.0, wb, 0.0);

    p.off = ROW_BACK_CAP - nb - 2;
    p.npts = 2 + nb + nf + 2;
    p.wn_raw = wf;
    const double scale = w_n / wf;
    for (int k = p.off; k < p.off + p.npts; ++k) row.w[k] = row.w[k] * scale;
    int nw = 0;
    for (int k = 0; k < p.npts; ++k) {
        if (row.xi[p.off + k] >= v_wall) { nw = k; break; }
    }
    p.n_wall = nw;
    return p;
}

// ---------------------------------------------------------------------------------------------------------------
// The bag solver as a resumable machine (same idea as GenericMachine in ptt_generic.cuh): find_alpha_plus_bag and
// sound_shell_alpha_plus are nests of loops around the ODE integration; written as nested calls every lane of a warp
// sits in its own inlined copy of the integration loop and the warp serialises.  Here the only integration call site
// is the driver loop of bag_machine_run, and the logic around it consumes one integration and prepares the next.
// Arithmetic and order are those of bag_find_alpha_plus / bag_forward / bag_emit_profile above (bit-identical
// results, checked through the host emulation); the wall->shock passes of the root's last evaluation are reused for
// the profile instead of being run again.

enum BagJob : int { BJ_LOCATE = 0, BJ_EMIT_FWD = 1, BJ_EMIT_BACK = 2 };

struct BagVisitor {
    int kind;
    ShockLocator loc;
    ForwardEmitter fe;
    BackwardEmitter be;
    PTT_HD bool skip(const State& y_end) const { return kind == BJ_LOCATE && loc.skip(y_end); }
    template <class Sampler>
    PTT_HD int search(const int i_lo, const int i_hi, const Sampler& sample) {
        return (kind == BJ_LOCATE) ? loc.search(i_lo, i_hi, sample) : -1;
    }
    PTT_HD bool visit(const int i, const State& y) {
        if (kind == BJ_LOCATE) return loc.visit(i, y);
        if (kind == BJ_EMIT_FWD) return fe.visit(i, y);
        return be.visit(i, y);
    }
};

struct BagMachine {
    enum Phase : int { PH_FWD1 = 0, PH_FWD2 = 1, PH_EMIT_FWD = 2, PH_EMIT_BACK = 3, PH_DONE = 4 };
    enum RootState : int { R_NONE = 0, R_F0 = 1, R_F1 = 2, R_IT = 3 };

    // --- point
    double v_wall, alpha_n, w_n;
    int n_xi;
    bool want_profile;
    RowOut row;
    // --- pending integration
    int phase;
    State job_y0;
    double job_te;
    int job_n;
    BagVisitor vis;
    bool full_retry;            // the pending locate pass is the "never crossed" re-run in full mode
    bool forward_only;          // stop after the wall->shock evaluation (alpha_n_max_deflagration)
    // --- wall->shock evaluation in flight (bag_forward)
    double ap_eval;
    WallSpeeds ws;
    BagForward f;
    // --- root finder (bag_find_alpha_plus)
    int rstate, it;
    double lo, hi, x0, f0, x1, f1, x2;
    // --- results
    BagRoot r;
    BagProfile p;
    int nf, nb;
    double xif0, wf;

    PTT_HD void finish_failed(const int status) { p.status = status; phase = PH_DONE; }

    // ---- bag_forward, resumable ---------------------------------------------------------------------------------
    PTT_HD void locate_job(const double te, const int n, const bool full) {
        vis.kind = BJ_LOCATE;
        vis.loc.reset(full);
        job_te = te; job_n = n;
        full_retry = full;
    }
    // returns true if an integration is pending, false if the evaluation is already over (f.status says why)
    PTT_HD bool forward_begin(const double ap, const int sol_type) {
        ap_eval = ap;
        f.status = ST_OK;
        f.i1 = 0; f.t_end2 = 0.0; f.n_fwd = 0; f.w1 = NAN; f.wn_raw = NAN;
        f.last.v = f.last.w = f.last.xi = NAN;
        ws = fluid_speeds_at_wall(v_wall, ap, sol_type);
        if (!(ws.vfp_p == ws.vfp_p)) { f.status = ST_INVALID_INPUT; return false; }
        job_y0.v = ws.vfp_p; job_y0.w = 1.0; job_y0.xi = v_wall;
        locate_job(-T_END_DEFAULT, N_XI_DEFAULT, false);
        phase = PH_FWD1;
        return true;
    }
    // forward_pass's epilogue; returns false on failure
    PTT_HD bool pass_result(const int n, int& n_keep, State& last, State& prev, State& y1) {
        y1 = vis.loc.y1;
        if (vis.loc.found >= 0) {
            n_keep = vis.loc.found + 1;
            last = vis.loc.cur;
            prev = vis.loc.prev;
        } else {
            n_keep = n - 1;
            last = vis.loc.prev2;
            prev = vis.loc.prev2;
            if (n_keep < 1) return false;
        }
        return true;
    }
    // After a locate integration of the evaluation in flight.  Returns true if another integration is pending,
    // false when the evaluation is complete (f filled, f.status).
    PTT_HD bool forward_after(const int rc) {
        if (rc < 0) { f.status = ST_INTEGRATION_FAILED; return false; }
        if (vis.loc.found < 0 && !full_retry) {
            // never crossed: the reference's index stays -2 and the slice [: -1] silently drops the last sample
            locate_job(job_te, job_n, true);
            return true;
        }
        int nk;
        State last, prev, y1;
        if (!pass_result(job_n, nk, last, prev, y1)) { f.status = ST_INTEGRATION_FAILED; return false; }
        if (phase == PH_FWD1) {
            f.i1 = nk - 1;
            f.t_end2 = grid_time(-T_END_DEFAULT, N_XI_DEFAULT, f.i1);
            locate_job(f.t_end2, n_xi, false);
            phase = PH_FWD2;
            return true;
        }
        if (nk >= 2) shock_zoom(prev, last);
        f.n_fwd = nk;
        f.last = last;
        f.w1 = (nk == 2) ? last.w : y1.w;
        const double vfp_s = last.xi;
        const double vfm_s = 1.0 / (3.0 * vfp_s);
        f.wn_raw = last.w * enthalpy_ratio(vfm_s, vfp_s);
        return false;
    }

    // ---- bag_find_alpha_plus, resumable ---------------------------------------------------------------------------
    PTT_HD void root_fail(const int st) {
        r.status = (st != ST_OK) ? st : ST_ROOT_NOT_CONVERGED;
        rstate = R_NONE;
    }
    // Consumes a completed evaluation.  Returns true if an integration is pending, false if the root search is over.
    PTT_HD bool root_on_eval() {
        const double fv = ap_eval / f.wn_raw - alpha_n;          // bag_alpha_n_of_alpha_plus(...) - alpha_n
        const int st = f.status;
        for (;;) {
            if (rstate == R_F0) {
                f0 = fv;
                if (st != ST_OK || !(f0 == f0)) { root_fail(st); return false; }
                if (f0 < 0.0) lo = x0; else hi = x0;
                x1 = x0 * ((f0 < 0.0) ? 1.001 : 0.999);
                if (x1 <= lo || x1 >= hi) x1 = 0.5 * (lo + hi);
                rstate = R_F1;
                if (forward_begin(x1, r.sol_type)) return true;
                root_fail(f.status);          // (an evaluation that cannot start has a NaN residual)
                return false;
            }
            if (rstate == R_F1) {
                f1 = fv;
                if (st != ST_OK || !(f1 == f1)) { root_fail(st); return false; }
                if (f1 < 0.0) { if (x1 > lo) lo = x1; } else { if (x1 < hi) hi = x1; }
                r.status = ST_ROOT_NOT_CONVERGED;
                it = 0;
            } else {        // R_IT: fv belongs to x2
                const double f2 = fv;
                if (st != ST_OK || !(f2 == f2)) { root_fail(st); return false; }
                if (f2 < 0.0) { if (x2 > lo) lo = x2; } else { if (x2 < hi) hi = x2; }
                const double dx = fabs(x2 - x1);
                x0 = x1; f0 = f1;
                x1 = x2; f1 = f2;
                if (dx <= 1e-13 * fabs(x2) || fabs(f2) <= 1e-14 * alpha_n || (hi - lo) <= 1e-14 * hi ||
                    (dx <= 1e-10 * fabs(x2) && fabs(f2) <= 1e-11 * alpha_n)) {
                    r.status = ST_OK;
                    break;
                }
                ++it;
            }
            // loop header of the secant iteration
            if (it >= 60) break;
            r.iters = it + 1;
            if (f1 == 0.0) { r.status = ST_OK; break; }
            x2 = (f1 != f0) ? x1 - f1 * (x1 - x0) / (f1 - f0) : 0.5 * (lo + hi);
            if (!(x2 > lo && x2 < hi)) x2 = 0.5 * (lo + hi);
            rstate = R_IT;
            if (forward_begin(x2, r.sol_type)) return true;
            root_fail(f.status);
            return false;
        }
        r.alpha_plus = x1;
        rstate = R_NONE;
        return false;
    }

    // ---- sound_shell_alpha_plus, resumable ------------------------------------------------------------------------
    PTT_HD void profile_init() {
        p.status = ST_OK; p.off = 0; p.npts = 0; p.n_wall = 0; p.xi_sh = v_wall; p.wn_raw = NAN;
        nf = 0; nb = 0;
        xif0 = v_wall + 1.0 / (double)n_xi;
        wf = 1.0;
    }
    // the wall->shock evaluation `f` at r.alpha_plus is known: emit the forward part
    PTT_HD void emit_forward_job() {
        if (f.status != ST_OK) { finish_failed(f.status); return; }
        job_y0.v = ws.vfp_p; job_y0.w = 1.0; job_y0.xi = v_wall;
        job_te = f.t_end2; job_n = n_xi;
        vis.kind = BJ_EMIT_FWD;
        vis.fe.row = row; vis.fe.zero_hit = false; vis.fe.found = -1; vis.fe.last_i = -1;
        phase = PH_EMIT_FWD;
    }
    PTT_HD void after_forward_part() {
        row.set(ROW_BACK_CAP + nf, 0.0, wf, xif0);
        row.set(ROW_BACK_CAP + nf + 1, 0.0, wf, 1.0);
        if (r.sol_type != SOL_SUB_DEF) {
            const double wm = 1.0 / enthalpy_ratio(ws.vfm_w, ws.vfp_w);
            job_y0.v = ws.vfm_p; job_y0.w = wm; job_y0.xi = v_wall;
            job_te = -T_END_DEFAULT; job_n = N_XI_DEFAULT;
            vis.kind = BJ_EMIT_BACK;
            vis.be.row = row; vis.be.cs2b = 1.0 / 3.0; vis.be.hit = -1; vis.be.last_i = -1;
            vis.be.n_start = (r.sol_type == SOL_DETON) ? 0 : 1;
            phase = PH_EMIT_BACK;
            return;
        }
        const double dxi = 1.0 / (double)n_xi;
        const double wm = 1.0 / enthalpy_ratio(ws.vfm_w, ws.vfp_w);
        finalize(fmin(cs0_root(), v_wall) - dxi, wm);
    }
    PTT_HD void finalize(const double xib0, const double wb) {
        row.set(ROW_BACK_CAP - nb - 1, 0.0, wb, xib0);
        row.set(ROW_BACK_CAP - nb - 2, 0.0, wb, 0.0);
        p.off = ROW_BACK_CAP - nb - 2;
        p.npts = 2 + nb + nf + 2;
        p.wn_raw = wf;
        const double scale = w_n / wf;
        for (int k = p.off; k < p.off + p.npts; ++k) row.w[k] = row.w[k] * scale;
        int nw = 0;
        for (int k = 0; k < p.npts; ++k) {
            if (row.xi[p.off + k] >= v_wall) { nw = k; break; }
        }
        p.n_wall = nw;
        phase = PH_DONE;
    }
    PTT_HD void profile_begin() {
        profile_init();
        if (r.sol_type == SOL_DETON) {
            ws = fluid_speeds_at_wall(v_wall, r.alpha_plus, r.sol_type);
            after_forward_part();
            return;
        }
        emit_forward_job();
    }

    // ---- entry points ---------------------------------------------------------------------------------------------
    // root (+ profile): find alpha_+ for (v_wall, alpha_n)
    PTT_HD void start_root(const double v_wall_, const double alpha_n_, const int n_xi_, const double an_max_defl,
                           const bool want_profile_, const double w_n_, const RowOut& row_) {
        v_wall = v_wall_; alpha_n = alpha_n_; n_xi = n_xi_; want_profile = want_profile_; w_n = w_n_; row = row_;
        phase = PH_DONE; rstate = R_NONE; forward_only = false;
        p.status = ST_OK; p.off = 0; p.npts = 0; p.n_wall = 0; p.xi_sh = v_wall; p.wn_raw = NAN;
        r.iters = 0;
        r.alpha_plus = NAN;
        r.sol_type = identify_solution_type_bag(v_wall, alpha_n, an_max_defl);
        if (r.sol_type == SOL_ERROR) { r.status = ST_NO_SOLUTION; return; }
        if (r.sol_type == SOL_DETON) {
            r.status = ST_OK; r.alpha_plus = alpha_n;
            if (want_profile) profile_begin();
            return;
        }
        double guess;
        if (alpha_n < 0.05) {
            guess = alpha_n;
        } else {
            const double ap_min = alpha_plus_min_hybrid(v_wall);
            const double an_min = alpha_plus_max_detonation(v_wall);
            const double slope = (1.0 / 3.0 - ap_min) / (an_max_defl - an_min);
            guess = ap_min + slope * (alpha_n - an_min);
        }
        const double hi_lim = 1.0 / 3.0 - 1.0e-10;    // the reference's own upper probe (alpha.py:47)
        lo = 0.0; hi = hi_lim;                         // residual < 0 at lo, > 0 at hi (alpha_n < an_max_defl)
        if (guess <= lo) guess = 1e-3 * alpha_n;
        if (guess >= hi) guess = 0.5 * (lo + hi);
        x0 = guess;
        rstate = R_F0;
        if (!forward_begin(x0, r.sol_type)) { root_fail(f.status); phase = PH_DONE; }
    }
    // profile only: sound_shell_alpha_plus(v_wall, alpha_plus, sol_type)
    PTT_HD void start_profile(const double v_wall_, const double alpha_plus, const int sol_type, const int n_xi_,
                              const double w_n_, const RowOut& row_) {
        v_wall = v_wall_; alpha_n = NAN; n_xi = n_xi_; want_profile = true; w_n = w_n_; row = row_;
        phase = PH_DONE; rstate = R_NONE; forward_only = false;
        r.status = ST_OK; r.sol_type = sol_type; r.alpha_plus = alpha_plus; r.iters = 0;
        if (sol_type == SOL_DETON) { profile_begin(); return; }
        profile_init();
        if (!forward_begin(alpha_plus, sol_type)) finish_failed(f.status);
    }

    // alpha_n_max_deflagration_bag(v_wall) (alpha.py:44-50): the wall->shock passes of the alpha_+ = 1/3 - 1e-10 shell on
    // the default grid; the result is anmax() once phase == PH_DONE.
    PTT_HD void start_anmax(const double v_wall_) {
        v_wall = v_wall_; alpha_n = NAN; n_xi = N_XI_DEFAULT; want_profile = false; w_n = 1.0;
        phase = PH_DONE; rstate = R_NONE; forward_only = true;
        r.sol_type = (v_wall > cs0()) ? SOL_HYBRID : SOL_SUB_DEF;
        forward_begin(1.0 / 3.0 - 1.0e-10, r.sol_type);
    }
    PTT_HD double anmax(int& status) const {
        status = f.status;
        return (f.status == ST_OK) ? (f.w1 / f.wn_raw) * (1.0 / 3.0) : NAN;
    }

    // Consumes the integration that was pending; afterwards another one is pending or phase == PH_DONE.
    PTT_HD void advance(const int rc) {
        if (phase == PH_FWD1 || phase == PH_FWD2) {
            if (forward_after(rc)) return;
            // evaluation complete
            if (forward_only) { phase = PH_DONE; return; }
            if (rstate != R_NONE) {
                if (root_on_eval()) return;
                // root search over
                if (r.status != ST_OK || !want_profile) { phase = PH_DONE; return; }
                profile_init();          // f is the evaluation at r.alpha_plus (the last one made)
            }
            emit_forward_job();
            return;
        }
        if (rc < 0) { finish_failed(ST_INTEGRATION_FAILED); return; }
        if (phase == PH_EMIT_FWD) {
            nf = (vis.fe.found >= 0) ? vis.fe.found + 1 : n_xi - 1;
            if (nf >= 2) {
                State l, q;
                const int kl = ROW_BACK_CAP + nf - 1;
                l.v = row.v[kl]; l.w = row.w[kl]; l.xi = row.xi[kl];
                q.v = row.v[kl - 1]; q.w = row.w[kl - 1]; q.xi = row.xi[kl - 1];
                shock_zoom(q, l);
                row.set(kl, l.v, l.w, l.xi);
            }
            const int kl = ROW_BACK_CAP + nf - 1;
            const double vfp_s = row.xi[kl];
            const double vfm_s = 1.0 / (3.0 * vfp_s);
            wf = row.w[kl] * enthalpy_ratio(vfm_s, vfp_s);
            xif0 = vfp_s;
            p.xi_sh = vfp_s;
            after_forward_part();
            return;
        }
        // PH_EMIT_BACK
        int n_stop;
        if (vis.be.hit < 0) n_stop = N_XI_DEFAULT - 2 + vis.be.n_start;       // slice [: -2] (+1 for hybrids)
        else n_stop = ((vis.be.hit == 0) ? 1 : vis.be.hit) + vis.be.n_start;
        nb = n_stop - vis.be.n_start;
        if (nb < 1) { finish_failed(ST_INTEGRATION_FAILED); return; }
        finalize(cs0_root(), row.w[ROW_BACK_CAP - nb]);
    }

    PTT_HD void run() {
        while (phase != PH_DONE) {
            const int rc = integrate_grid(1.0 / 3.0, job_y0, job_te, job_n, vis);
            advance(rc);
        }
    }
};

}  // namespace ptt
