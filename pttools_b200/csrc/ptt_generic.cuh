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
    int first_above;    // first sample with xi > cs_n (-1: none so far)
    bool turned;
    State cur, prev, prev2, close_prev;     // close_prev: the sample before first_close

    PTT_HD void reset(const EosDev& m_, const double v_wall_, const double cs_n_, const bool full_) {
        m = m_; v_wall = v_wall_; cs_n = cs_n_; full = full_;
        found = 0; last_i = -1; first_close = -1; first_above = -1; turned = false;
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
        // (near the fixed point (cs, 0) every sample is looked at: the index at which rounding lifts xi to cs there -
        // it fixes v_sh = cs to 1e-14, as the reference's result has it - cannot be found any other way)
        return !full && !behind(y_end) && !near_cs(y_end) && (last_i < 0 || y_end.xi > prev.xi);
    }
    // The first sample behind the shock curve inside one integrator step, by bisection instead of the ordered walk:
    // on thin shells a single step holds thousands of tau-grid samples (the step is set by the accuracy of the curve,
    // the grid by n_xi).  Taken only in the clean case - the curve moves outwards over the step, away from the fixed
    // point (cs, 0), the last sample of the step is behind the shock curve and the sample before the step is not - and
    // with the assumption skip() already makes: within one step the curve crosses the shock curve once.  Leaves the
    // locator as the ordered walk would (found, cur, prev, prev2, last_i).
    template <class Sampler>
    PTT_HD int search(const int i_lo, const int i_hi, const Sampler& sample) {
        if (full || last_i != i_lo - 1) return -1;
        if (m.bag_name && first_close >= 0) {
            // Bag model, the curve creeping into the fixed point (cs, 0) after its first close sample has been noted.
            // Nothing but a sample behind the shock curve is looked for any more, and that needs xi >= cs (v_shock is
            // NaN below, shock.py:63-66).  The index that is finally taken is the one at which the noise of the
            // continuous extension lifts xi to cs (it fixes v_sh = cs to 1e-13, as the reference's result has it): every
            // sample has to be looked at, but only its xi - and not even that in a step that stays below cs.
            const double cs = cs0();
            if (sample.xi_upper_bound() < cs - 1e-14) return -2;
            for (int k = i_lo; k <= i_hi; ++k) {
                if (sample.xi(k) < cs) continue;
                const State y = sample(k);
                if (behind(y)) {
                    if (k - 1 >= i_lo) prev = sample(k - 1);
                    found = k; cur = y; last_i = k - 1;
                    return k;
                }
            }
            prev = sample(i_hi);
            last_i = i_hi;
            return -3;
        }
        const double v_clear = 1e-6;                                        // near_cs() needs |v| <= 1e-8
        const State y_hi = sample(i_hi);
        if (i_lo == 1 && skip(y_hi)) return -2;                             // (the walk applies skip() from the second step on)
        if (!(prev.v > v_clear) || !(y_hi.v > v_clear) || !(y_hi.xi > prev.xi)) return -1;
        if (!behind(y_hi) || behind(prev)) return -1;
        int lo = i_lo - 1, hi = i_hi;                                       // not behind / behind
        State y_found = y_hi;
        while (hi - lo > 1) {
            const int mid = lo + ((hi - lo) >> 1);
            const State y = sample(mid);
            if (!(y.v > v_clear)) return -1;
            if (behind(y)) { hi = mid; y_found = y; } else { lo = mid; }
        }
        // samples hi - 2, hi - 1 as the walk would hold them (first_above is not read once the shock is found)
        if (hi - 1 >= i_lo) {
            const State y1 = sample(hi - 1);
            const State y2 = (hi - 2 >= i_lo) ? sample(hi - 2) : prev;
            if (!m.bag_name && !(y1.xi >= y2.xi && y_found.xi >= y1.xi)) return -1;        // a turning curve: ordered walk
            prev2 = y2;
            prev = y1;
        }
        found = hi; cur = y_found; last_i = hi - 1;
        return hi;
    }
    PTT_HD bool visit(const int i, const State& y) {
        if (first_above < 0 && y.xi > cs_n) first_above = i;                // i_cs_n = argmax(xi > cs_n), shock.py:112
        if (i > 0) {
            if (first_close < 0 && near_cs(y)) {
                first_close = i; close_prev = prev;
                // shock.py:109-121: the first sample close to (cs_n, 0) IS the shock index when the curve has not been
                // above cs_n before it -- decided before the shock curve is looked at, whatever follows (xi that
                // wiggles by an ulp around the fixed point, v that rounds to -1e-17)
                if (!m.bag_name && (first_above < 0 || first_above >= i)) { found = i; cur = y; return false; }
            }
            if (behind(y)) { found = i; cur = y; return false; }
            if (!m.bag_name && y.xi < prev.xi) {
                // the curve turned back before reaching the shock curve: i_right = argmax(xi) (shock.py:134-141)
                turned = true;
                found = last_i;
                cur = prev;
                prev = prev2;
                return false;
            }
        }
        prev2 = prev;
        prev = y;
        last_i = i;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// The solver as a resumable machine.
//
// sound_shell_generic is a nest of loops (root finder -> residual -> thin-shell attempts -> ODE integration ->
// samples).  Written as nested calls, every lane of a warp sits in its own copy of the integration loop at its own
// place in the nest and the warp serialises (measured: 4.7 of 32 lanes active).  Here the nest is turned inside out:
// the ONLY integration call site is in the driver loop of generic_solve_point, and everything around it is a state
// machine that consumes the result of one integration and prepares the next one.  All lanes of a warp are then in
// the same integration loop, whatever phase of the algorithm each of them is in.  The arithmetic and its order are
// those of the nested formulation (fluid.py:121-288, 606-669, 697-843, 848-1052).

enum JobKind : int { JOB_LOCATE = 0, JOB_EMIT = 1, JOB_CSTRIM = 2 };

struct OffsetEmitter {
    RowOut row;
    int base;       // column of sample 0
    int dir;        // +1 forward, -1 reversed
    int limit;      // stop after this many samples (forward part: i_shock)
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
    PTT_HD bool visit(const int i, const State& y) {
        last_i = i;
        row.set(base - i, y.v, y.w, y.xi);
        if ((y.v <= 0.0) || (y.xi * y.xi <= cs2b)) { hit = i; return false; }
        return true;
    }
};

// One visitor type for every kind of integration, so that integrate_grid is instantiated (and entered) once.
struct JobVisitor {
    int kind;
    GenericShockLocator loc;
    OffsetEmitter oe;
    CsTrimEmitter ce;
    PTT_HD bool skip(const State& y_end) const { return kind == JOB_LOCATE && loc.skip(y_end); }
    template <class Sampler>
    PTT_HD int search(const int i_lo, const int i_hi, const Sampler& sample) {
        return (kind == JOB_LOCATE) ? loc.search(i_lo, i_hi, sample) : -1;
    }
    PTT_HD bool visit(const int i, const State& y) {
        if (kind == JOB_LOCATE) return loc.visit(i, y);
        if (kind == JOB_EMIT) return oe.visit(i, y);
        return ce.visit(i, y);
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

struct GenericGuess {
    double vp, vp_tilde, wp, wm;    // starting guesses (fluid reference table scaled by wn, fluid.py:925-949)
};

struct GenericOut {
    int status;
    int off, npts;
    int solver_failed;
    int n_resid;        // residual evaluations of the root search (diagnostics)
    // the reference's 17-tuple scalars (fluid.py:1052)
    double vp, vm, vp_tilde, vm_tilde, v_sh, vm_sh, vm_tilde_sh, wp, wm, wm_sh, v_cj, wn, wn_estimate;
};

// Row layout for the generic solver: the part behind the wall (2 still-fluid samples + up to n_xi samples) ends at
// column gen_back_cap(n_xi); the wall -> shock part starts there.
PTT_HD int gen_back_cap(const int n_xi) { return n_xi + 2; }
PTT_HD int gen_row_capacity(const int n_xi) { return 2 * n_xi + 4; }

struct GenericMachine {
    // what comes back from "advance": an integration to run, or nothing more to do
    enum Phase : int { PH_LOCATE = 0, PH_EMIT_FWD = 1, PH_EMIT_BACK = 2, PH_CSTRIM = 3, PH_DONE = 4, PH_HOLD = 5 };
    // where the root finder waits for a residual
    enum Root : int { RF_A0 = 0, RF_AH = 1, RF_A1 = 2, RF_IT = 3, RF_FINAL = 4, RF_H0 = 5, RF_HJ = 6, RF_HT = 7 };
    enum PassKind : int { PK_SECANT = 0, PK_NEWTON = 1, PK_HYBRD = 2 };

    // --- constants of the point
    EosDev m;
    double v_wall, wn, cs_n, v_cj, t_end, wn_rtol;
    int sol_type, n_xi, thin;
    GenericGuess g;
    RowOut row;
    // --- pending integration
    int phase;
    double job_cs2, job_te;
    State job_y0;
    JobVisitor vis;
    // --- residual evaluation in flight (sound_shell_deflagration_common, fluid.py:121-288)
    DeflagResult d;
    double vm_tilde, vpt_guess;
    int attempt, i_shock;
    State sh;
    double te;
    // --- root finder (fluid.py:606-669, 697-843)
    int rf, pass, it;
    bool newton_start, converged, solver_failed;
    bool hold;                  // candidate of the retry search: park (PH_HOLD) when the root search ends, before the final evaluation
    double p0, p1, q0, q1, p, h_fd;
    // MINPACK hybrd in one dimension (scipy fsolve: xtol 1.49012e-8, factor 100, maxfev 400, forward differences)
    double hx, hf, hJ, hdiag, hdelta, hxnorm, hfnorm, hp, hpnorm;
    int hiter, ncsuc, ncfail, nslow1, nslow2, nfev;
    int n_resid;                // residual evaluations started (diagnostics: scalars[14])
    bool jeval;
    // --- result
    GenericOut o;
    int nb, nf;

    PTT_HD void fail(const int status) { o.status = status; o.n_resid = n_resid; phase = PH_DONE; }

    // ---- residual -------------------------------------------------------------------------------------------
    // Starts wn_estimate(wm) - wn.  Returns true if an integration is pending, false if the residual is already
    // known (NaN, d.status says why).
    PTT_HD bool resid_begin(const double wm) {
        ++n_resid;
        d.status = ST_SOLVER_FAILED;
        d.vp = d.vp_tilde = d.wp = d.wn_estimate = NAN;
        d.i_shock = 0; d.t_end_final = -t_end;
        d.sh.v = d.sh.w = d.sh.xi = NAN;
        if (!(wm >= 0.0)) return false;
        if (!(v_wall >= 0.0 && v_wall <= 1.0) || !(vm_tilde >= 0.0 && vm_tilde <= 1.0) || !(wn >= 0.0) ||
            !(vpt_guess >= 0.0 && vpt_guess <= 1.0) || v_wall > v_cj)
            return false;
        // wall junction, broken -> symmetric: a deflagration accelerates the flow through the wall, v+ < v-
        // (root below cs_s)
        const Junction j = solve_junction(m, vm_tilde, wm, 1, 0, vpt_guess, -1);
        if (!j.ok) return false;
        d.vp_tilde = j.v2;
        d.wp = j.w2;
        d.vp = lorentz(v_wall, j.v2);        // -lorentz(vp_tilde, v_wall)
        if (!(d.vp >= 0.0) || !(d.wp >= 0.0)) return false;
        job_y0.v = d.vp; job_y0.w = d.wp; job_y0.xi = v_wall;
        job_cs2 = m.css2;
        te = -t_end;
        attempt = 0;
        i_shock = 0;
        locate_job();
        return true;
    }
    PTT_HD void locate_job() {
        vis.kind = JOB_LOCATE;
        vis.loc.reset(m, v_wall, cs_n, false);
        job_te = te;
        phase = PH_LOCATE;
    }
    // After a locate integration: first pass + up to 5 thin-shell refinements (fluid.py:223-271).
    // Returns true if another integration is pending, false if the residual is complete.
    PTT_HD bool resid_after(const int rc) {
        if (rc < 0) { d.status = ST_INTEGRATION_FAILED; return false; }
        int i_new = vis.loc.found;
        State before = vis.loc.prev;
        if (i_new == 0 && vis.loc.first_close > 0) {
            // The curve ran into the fixed point (xi, v) = (cs, 0) without ever getting behind the shock curve: a shock
            // of vanishing strength at xi = cs+.  With the reference's integrator the index is then decided by LSODA
            // noise (xi scatters above cs, where v <= v_shock ~ 1e-9 is met at once); the first sample that
            // np.isclose()s the fixed point - the criterion shock.py:114-121 uses for the other models - is taken.
            i_new = vis.loc.first_close;
            before = vis.loc.close_prev;
        }
        bool more = false;
        if (attempt == 0) {
            if (i_new == 0 || i_new + 1 >= n_xi) return false;
        }
        if (attempt > 0 && (i_new == 0 || i_new + 20 >= n_xi)) {
            // keep the previous solution
        } else {
            i_shock = i_new;
            sh = before;
            d.t_end_final = te;
            if (i_shock < thin && attempt < 5) {
                te = grid_time(te, n_xi, i_shock + 20);
                more = true;
            }
        }
        if (more) {
            ++attempt;
            locate_job();
            return true;
        }
        if (i_shock <= 1) return false;
        d.i_shock = i_shock;
        d.sh = sh;
        const double vm_tilde_sh = lorentz(sh.xi, sh.v);
        d.wn_estimate = sh.w * gamma2(vm_tilde_sh) * vm_tilde_sh / (gamma2(sh.xi) * sh.xi);   // w2_junction, boundary.py:638
        d.status = ST_OK;
        return false;
    }

    // ---- root finder ----------------------------------------------------------------------------------------
    // Subsonic deflagrations: scipy's secant (root_scalar, x0 = 0.99 wm_guess, x1 = 1.01 wm_guess, atol 1.48e-8),
    // mirrored step by step; if it fails, the reference's fallback (fsolve from wm_guess, fluid.py:629-640).
    // Hybrids: fsolve from wm_guess (fluid.py:697-720).  fsolve is MINPACK's hybrd; in one dimension it is restated
    // here in full (forward-difference Jacobian, trust region with the dogleg step clipped to it, Broyden updates,
    // Jacobian recomputed after two failures, hybrd's five stopping tests): on the thin shells whose residual jumps
    // with the tau-grid index of the shock, hybrd shrinks its trust region onto the jump and reports success, and so
    // must this.
    // Each function returns true if an integration is pending; otherwise the residual it asked for is already
    // known and the caller feeds it back with root_on_q.
    // Passes.  Subsonic deflagrations: secant, then the fsolve fallback.  Hybrids: fsolve.  fsolve is run in two
    // stages: first its smooth-case behaviour (finite-difference Newton step, then secant updates - what hybrd does
    // when every step succeeds; 6-8 evaluations), and only if that does not converge, hybrd proper (up to ~50).
    PTT_HD int pass_kind() const {
        const int k = (sol_type == SOL_SUB_DEF) ? pass : pass + 1;
        return (k > 2) ? 2 : k;
    }
    PTT_HD bool pass_begin() {
        const int kind = pass_kind();
        newton_start = kind != PK_SECANT;
        if (kind == PK_HYBRD) {
            hx = g.wm;
            p = hx;
            rf = RF_H0;
            return resid_begin(hx);
        }
        p0 = newton_start ? g.wm : 0.99 * g.wm;
        rf = RF_A0;
        return resid_begin(p0);
    }
    // hybrd: forward-difference Jacobian at hx.  Returns 1 = integration pending, 0 = residual known without one.
    PTT_HD int hybrd_jac() {
        jeval = true;
        h_fd = 1.4901161193847656e-08 * fabs(hx);
        if (h_fd == 0.0) h_fd = 1.4901161193847656e-08;
        rf = RF_HJ;
        return resid_begin(hx + h_fd) ? 1 : 0;
    }
    // hybrd: dogleg step (in one dimension: the Newton step clipped to the trust region) and its trial evaluation
    PTT_HD int hybrd_trial() {
        const double pgn = (hJ != 0.0) ? -hf / hJ : ((hf >= 0.0) ? INFINITY : -INFINITY);
        const double qnorm = hdiag * fabs(pgn);
        hp = (qnorm <= hdelta) ? pgn : copysign(hdelta / hdiag, pgn);
        hpnorm = hdiag * fabs(hp);
        if (hiter == 1) hdelta = fmin(hdelta, hpnorm);
        rf = RF_HT;
        return resid_begin(hx + hp) ? 1 : 0;
    }
    // hybrd: consumes the trial residual q.  Returns 1 / 0 as above, -1 when the pass is over (converged set).
    PTT_HD int hybrd_after_trial(const double q) {
        const double xtol = 1.49012e-08, epsmch = 2.220446049250313e-16;
        ++nfev;
        const double fnorm1 = (q == q) ? fabs(q) : INFINITY;
        double actred = -1.0;
        if (fnorm1 < hfnorm) actred = 1.0 - (fnorm1 / hfnorm) * (fnorm1 / hfnorm);
        const double temp = fabs(hf + hJ * hp);
        double prered = 0.0;
        if (temp < hfnorm) prered = 1.0 - (temp / hfnorm) * (temp / hfnorm);
        const double ratio = (prered > 0.0) ? actred / prered : 0.0;
        if (ratio < 0.1) {
            ncsuc = 0; ++ncfail;
            hdelta *= 0.5;
        } else {
            ncfail = 0; ++ncsuc;
            if (ratio >= 0.5 || ncsuc > 1) hdelta = fmax(hdelta, hpnorm / 0.5);
            if (fabs(ratio - 1.0) <= 0.1) hdelta = hpnorm / 0.5;
        }
        const double f_old = hf;
        if (ratio >= 1e-4) {
            hx += hp; hf = q;
            hxnorm = hdiag * fabs(hx);
            hfnorm = fnorm1;
            ++hiter;
        }
        p = hx;
        ++nslow1;
        if (actred >= 0.001) nslow1 = 0;
        if (jeval) ++nslow2;
        if (actred >= 0.1) nslow2 = 0;
        if (hdelta <= xtol * hxnorm || hfnorm == 0.0) { converged = true; return -1; }          // info = 1
        if (nfev >= 400 || 0.1 * fmax(0.1 * hdelta, hpnorm) <= epsmch * hxnorm || nslow2 == 5 || nslow1 == 10)
            return -1;                                                                           // info = 2 .. 5
        if (ncfail == 2) return hybrd_jac();
        if (q == q) hJ = (q - f_old) / hp;           // Broyden's rank-one update is the secant slope in one dimension
        jeval = false;
        return hybrd_trial();
    }
    // secant iteration header; returns 1 = integration pending, 0 = residual known without one, -1 = pass over
    PTT_HD int iter_head() {
        const double atol = 1.48e-8;
        const double rtol_h = 1.4901161193847656e-08;     // hybrd xtol (relative) for the Newton-started pass
        if (it >= (newton_start ? 12 : 50)) return -1;
        if (q1 == q0) {
            p = 0.5 * (p1 + p0);
            converged = (p1 == p0);
            return -1;
        }
        if (fabs(q1) > fabs(q0)) p = (-q0 / q1 * p1 + p0) / (1.0 - q0 / q1);
        else p = (-q1 / q0 * p0 + p1) / (1.0 - q1 / q0);
        const double tol = newton_start ? rtol_h * fabs(p) : atol;
        if (fabs(p - p1) <= tol) { converged = true; return -1; }
        p0 = p1; q0 = q1;
        p1 = p;
        rf = RF_IT;
        return resid_begin(p1) ? 1 : 0;
    }
    // end of a pass (and possibly of the root search); same return convention as iter_head, -1 never
    PTT_HD int pass_end() {
        if (!converged && pass_kind() != PK_HYBRD) {
            ++pass;
            return pass_begin() ? 1 : 0;
        }
        if (hold) { phase = PH_HOLD; return 1; }
        if (!(p == p) || !(p > 0.0)) { fail(ST_ROOT_NOT_CONVERGED); return 1; }
        // final evaluation at the root (fluid.py:642-653, 797-811)
        rf = RF_FINAL;
        return resid_begin(p) ? 1 : 0;
    }
    // Consumes a completed residual.  Returns true when the machine has a pending integration or is done.
    PTT_HD bool root_on_q() {
        const double q = d.wn_estimate - wn;      // NaN if the evaluation failed
        int r;
        switch (rf) {
        case RF_A0:
            q0 = q;
            if (!newton_start) {
                p1 = 1.01 * g.wm;
                rf = RF_A1;
                return resid_begin(p1);
            }
            h_fd = 1.4901161193847656e-08 * fabs(p0);
            rf = RF_AH;
            return resid_begin(p0 + h_fd);
        case RF_AH: {
            const double J = (q - q0) / h_fd;
            p1 = p0 - q0 / J;
            if (!(p1 > 0.0) || !(p1 == p1)) p1 = 1.01 * p0;
            // hybrd's first trust region is factor * |x| = 100 |x|: no clipping in practice
            rf = RF_A1;
            return resid_begin(p1);
        }
        case RF_H0:
            hf = q;
            nfev = 1;
            hfnorm = fabs(hf);
            hiter = 1; ncsuc = ncfail = nslow1 = nslow2 = 0;
            if (!(hf == hf)) r = -1;
            else r = hybrd_jac();
            if (r < 0) r = pass_end();
            return r == 1;
        case RF_HJ:
            ++nfev;
            if (!(q == q)) {
                r = -1;
            } else {
                hJ = (q - hf) / h_fd;
                if (hiter == 1) {
                    hdiag = fabs(hJ);
                    if (hdiag == 0.0) hdiag = 1.0;
                    hxnorm = hdiag * fabs(hx);
                    hdelta = 100.0 * hxnorm;
                    if (hdelta == 0.0) hdelta = 100.0;
                }
                hdiag = fmax(hdiag, fabs(hJ));
                r = hybrd_trial();
            }
            if (r < 0) r = pass_end();
            return r == 1;
        case RF_HT:
            r = hybrd_after_trial(q);
            if (r < 0) r = pass_end();
            return r == 1;
        case RF_A1:
            q1 = q;
            p = p1;
            if (q0 == q0 && q1 == q1) {
                if (fabs(q1) < fabs(q0)) {
                    double t = p0; p0 = p1; p1 = t;
                    t = q0; q0 = q1; q1 = t;
                }
                it = 0;
                r = iter_head();
            } else {
                r = -1;
            }
            if (r < 0) r = pass_end();
            return r == 1;
        case RF_IT:
            q1 = q;
            if (!(q1 == q1)) {
                r = -1;
            } else {
                ++it;
                r = iter_head();
            }
            if (r < 0) r = pass_end();
            return r == 1;
        default:    // RF_FINAL
            o.wm = p;
            if (d.status != ST_OK) { fail(d.status); return true; }
            {
                const bool close = fabs(d.wn_estimate - wn) <= 1e-8 + wn_rtol * fabs(wn);     // np.isclose(.., rtol)
                solver_failed = !(converged && close);
            }
            emit_fwd_job();
            return true;
        }
    }

    // ---- profile materialisation ------------------------------------------------------------------------------
    PTT_HD void emit_fwd_job() {
        // wall -> shock samples [0, i_shock) of the accepted integration
        job_y0.v = d.vp; job_y0.w = d.wp; job_y0.xi = v_wall;
        job_cs2 = m.css2;
        job_te = d.t_end_final;
        vis.kind = JOB_EMIT;
        vis.oe.row = row; vis.oe.base = gen_back_cap(n_xi); vis.oe.dir = 1; vis.oe.limit = d.i_shock;
        phase = PH_EMIT_FWD;
    }
    PTT_HD void finish(const double w_first) {
        const int bc = gen_back_cap(n_xi);
        const double dxi = 1.0 / (double)n_xi;
        const double xi_first = (nb > 0) ? row.xi[bc - nb] : row.xi[bc];
        const double w_center = fmin(o.wm, w_first);                   // fluid.py:1039
        row.set(bc - nb - 1, 0.0, w_center, xi_first - dxi);
        row.set(bc - nb - 2, 0.0, w_center, 0.0);
        o.off = bc - nb - 2;
        o.npts = 2 + nb + nf + 2;
        o.n_resid = n_resid;
        phase = PH_DONE;
    }

    PTT_HD void start(const EosDev& m_, const double v_wall_, const int sol_type_, const double wn_, const double v_cj_,
                      const GenericGuess& g_, const int n_xi_, const double t_end_, const int thin_,
                      const double wn_rtol_, const RowOut& row_, const bool hold_ = false) {
        m = m_; v_wall = v_wall_; sol_type = sol_type_; wn = wn_; v_cj = v_cj_; g = g_; n_xi = n_xi_; t_end = t_end_;
        thin = thin_; wn_rtol = wn_rtol_; row = row_; hold = hold_;
        cs_n = sqrt(m.css2);
        o.status = ST_OK; o.off = 0; o.npts = 0; o.solver_failed = 1;
        o.vp = o.vm = o.vp_tilde = o.vm_tilde = o.v_sh = o.vm_sh = o.vm_tilde_sh = o.wp = o.wm = o.wm_sh = NAN;
        o.v_cj = v_cj; o.wn = wn; o.wn_estimate = NAN; o.n_resid = 0;
        nb = nf = 0;
        n_resid = 0;
        solver_failed = true; converged = false;
        p = NAN;
        if (sol_type == SOL_DETON) {
            // fluid.py:351-478.  Weak detonation: exit speed above the sound speed of the broken phase (root >= cs_b).
            const Junction j = solve_junction(m, v_wall, wn, 0, 1, 0.0, +1);
            if (!j.ok) { fail(ST_SOLVER_FAILED); return; }
            const double vmt = j.v2, wm = j.w2;
            const double vm = lorentz(v_wall, vmt);
            job_y0.v = vm; job_y0.w = wm; job_y0.xi = v_wall;
            job_cs2 = m.csb2;
            job_te = -t_end;
            vis.kind = JOB_CSTRIM;
            vis.ce.row = row; vis.ce.base = gen_back_cap(n_xi) - 1; vis.ce.cs2b = m.csb2; vis.ce.hit = -1;
            vis.ce.last_i = -1;
            o.vp = 0.0; o.vm = vm; o.vp_tilde = v_wall; o.vm_tilde = vmt; o.v_sh = v_wall; o.vm_sh = vm;
            o.vm_tilde_sh = vmt; o.wp = wn; o.wm = wm; o.wm_sh = wm;
            phase = PH_CSTRIM;
            return;
        }
        vm_tilde = (sol_type == SOL_HYBRID) ? sqrt(m.csb2) : v_wall;
        double vp_guess = g.vp;
        vpt_guess = g.vp_tilde;
        if (sol_type == SOL_SUB_DEF) {
            if (vp_guess > v_wall) vp_guess = 0.95 * v_wall;           // fluid.py:613-617
            vpt_guess = lorentz(v_wall, vp_guess);                     // -lorentz(vp_guess, v_wall), fluid.py:79
        }
        // a retry candidate is one fsolve call from g.wm: no secant pass for subsonic deflagrations
        pass = (hold && sol_type == SOL_SUB_DEF) ? 1 : 0;
        bool pending = pass_begin();
        while (!pending) pending = root_on_q();
    }
    // A parked candidate goes on to the final evaluation at its root and the profile (fluid.py:642-653, 797-811).
    PTT_HD void resume_final() {
        hold = false;
        if (!(p == p) || !(p > 0.0)) { fail(ST_ROOT_NOT_CONVERGED); return; }
        rf = RF_FINAL;
        bool pending = resid_begin(p);
        while (!pending) pending = root_on_q();
    }

    // Consumes the integration that was pending; afterwards either another one is pending or phase == PH_DONE.
    PTT_HD void advance(const int rc) {
        const int bc = gen_back_cap(n_xi);
        const double dxi = 1.0 / (double)n_xi;
        if (phase == PH_LOCATE) {
            bool pending = resid_after(rc);
            while (!pending) pending = root_on_q();
            return;
        }
        if (rc < 0) { fail(ST_INTEGRATION_FAILED); return; }
        if (phase == PH_CSTRIM) {
            const int hit = vis.ce.hit;
            nb = (hit < 0) ? n_xi - 2 : ((hit == 0) ? 1 : hit);
            if (nb < 1) { fail(ST_INTEGRATION_FAILED); return; }
            o.solver_failed = 0;
            nf = 0;
            const double w_first = row.w[bc - nb];
            row.set(bc, 0.0, wn, v_wall + dxi);
            row.set(bc + 1, 0.0, wn, 1.0);
            finish(w_first);
            return;
        }
        if (phase == PH_EMIT_FWD) {
            nf = d.i_shock;
            const double xi_last = row.xi[bc + nf - 1];
            row.set(bc + nf, 0.0, wn, xi_last + dxi);
            row.set(bc + nf + 1, 0.0, wn, 1.0);
            o.vp = d.vp; o.vp_tilde = d.vp_tilde; o.vm_tilde = vm_tilde; o.vm = lorentz(vm_tilde, v_wall);
            o.v_sh = d.sh.xi; o.vm_sh = d.sh.v; o.vm_tilde_sh = lorentz(d.sh.xi, d.sh.v);
            o.wp = d.wp; o.wm_sh = d.sh.w; o.wn_estimate = d.wn_estimate;
            o.solver_failed = solver_failed ? 1 : 0;
            if (sol_type == SOL_HYBRID) {
                // rarefaction tail behind the wall, n_xi samples, NOT trimmed (fluid.py:831-841)
                o.vm = lorentz(v_wall, sqrt(m.csb2));
                job_y0.v = o.vm; job_y0.w = o.wm; job_y0.xi = v_wall;
                job_cs2 = m.csb2;
                job_te = -t_end;
                vis.kind = JOB_EMIT;
                vis.oe.row = row; vis.oe.base = bc - 1; vis.oe.dir = -1; vis.oe.limit = n_xi;
                phase = PH_EMIT_BACK;
                return;
            }
            nb = 0;
            finish(row.w[bc]);
            return;
        }
        // PH_EMIT_BACK
        nb = n_xi;
        finish(row.w[bc - nb]);
    }
};

// What the reference tries after its first root search has failed, in its order (fsolve_vary, speedup/solvers.py:12-55;
// backup scan of the hybrid solver, fluid.py:731-795):
//   candidates 0, 1:   fsolve from wm_guess (1 +- 0.01) +- 1e-3
//   candidates 2..21:  hybrids only - fsolve from wm_i = linspace(0.3 wm_guess, 3 wm_guess, 20)[i - 2], for the wm_i whose
//                      v_plus lies above the shock curve at the wall
// The first candidate whose fsolve reports convergence gives wm.  Returns false if candidate c does not exist.
constexpr int GEN_RETRY_CANDIDATES = 22;
// s_coef_s / s_coef_b: entropy density s(w) = s_coef * w^(1 - 1/mu) of the two phases (both models; s = w / T).
PTT_HD bool generic_retry_start(const EosDev& m, const double v_wall, const int sol_type, const double cs_n,
                                const double wm_guess, const double vp_tilde_guess, const double s_coef_s,
                                const double s_coef_b, const int c, double& x0) {
    if (c == 0) { x0 = wm_guess * 1.01 + 1e-3; return true; }
    if (c == 1) { x0 = wm_guess * 0.99 - 1e-3; return true; }
    if (sol_type != SOL_HYBRID || c >= GEN_RETRY_CANDIDATES) return false;
    const double lo = 0.3 * wm_guess, hi = 3.0 * wm_guess;
    const int i = c - 2;
    x0 = (i == 19) ? hi : lo + (double)i * ((hi - lo) / 19.0);            // np.linspace(lo, hi, 20)
    // v_plus_hybrid (boundary.py:600-618) against v_shock at the wall (shock.py:389-408)
    const double v1 = sqrt(m.csb2);
    const Junction j = solve_junction(m, v1, x0, 1, 0, vp_tilde_guess, -1);
    if (!j.ok) return false;
    // In the scan the wall junction is solved with allow_negative_entropy_flux_change = False (fluid.py:745-750): a
    // start whose entropy flux gamma v s drops from the broken to the symmetric side gives NaN there and is skipped
    // (check_entropy_fluxes, boundary.py:65-85, 339-352).
    const double f1 = sqrt(gamma2(v1)) * v1 * s_coef_b * pow(x0, 1.0 - 1.0 / m.mu_b);
    const double f2 = sqrt(gamma2(j.v2)) * j.v2 * s_coef_s * pow(j.w2, 1.0 - 1.0 / m.mu_s);
    if (!(f1 >= 0.0) || !(f2 >= 0.0) || f2 - f1 < 0.0) return false;
    const double vp = lorentz(v_wall, j.v2);
    return vp > v_shock_generic(m, v_wall, cs_n);
}

// sound_shell_generic (fluid.py:848-1052) for one point, default arguments (no user guesses, no bag/Giese solver).
PTT_HD GenericOut generic_solve_point(const EosDev& m, const double v_wall, const double alpha_n, const int sol_type,
                                      const double wn, const double v_cj, const GenericGuess& g, const int n_xi,
                                      const double t_end, const int thin, const double wn_rtol, const RowOut& row) {
    (void)alpha_n;
    GenericMachine mach;
    mach.start(m, v_wall, sol_type, wn, v_cj, g, n_xi, t_end, thin, wn_rtol, row);
    while (mach.phase != GenericMachine::PH_DONE) {
        const int rc = integrate_grid(mach.job_cs2, mach.job_y0, mach.job_te, n_xi, mach.vis);
        mach.advance(rc);
    }
    return mach.o;
}

}  // namespace ptt
