// Batched FP64 adaptive Runge-Kutta core for the self-similar fluid equations.
//
// Replaces scipy.integrate.odeint (ODEPACK LSODA) at the reference call site pttools/bubble/integrate.py:180 and
// the RHS built at integrate.py:38-74.  One parameter point per thread: every lane owns its own step-size control.
// The output semantics are those of the reference: the solution is sampled on t = linspace(0, t_end, n) and events
// (shock, sound-speed crossing) are located BY GRID INDEX, not by root finding (trim.py:50-56, :94-99).
//
// The code is host/device portable (PTT_HD) so the same source can be compiled with g++ into a test-only emulation
// library (tests/hostemu); the product only ever runs the nvcc build.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define PTT_HD __host__ __device__ __forceinline__
#else
#define PTT_HD inline
#endif

// test-only hook of the host emulation (step statistics); empty in the product
#ifndef PTT_ODE_STEP_HOOK
#define PTT_ODE_STEP_HOOK(accepted)
#endif
#ifndef PTT_ODE_SAMPLE_HOOK
#define PTT_ODE_SAMPLE_HOOK(searched)
#endif

namespace ptt {

// Equation of state seen by the fluid equations: both in-scope models have a piecewise-constant sound speed,
// cs2(w, phase) = phase*csb2 + (1-phase)*css2  (models/bag.py:21-27, models/const_cs.py:553-557).
struct Eos {
    double css2;   // symmetric phase (ahead of the wall)
    double csb2;   // broken phase (behind the wall)
};

struct State {
    double v, w, xi;
};

PTT_HD void rhs(const double cs2, const State& y, State& d) {
    // integrate.py:62-73
    const double xiv = y.xi * y.v;
    const double xi_v = y.xi - y.v;
    const double v2 = y.v * y.v;
    const double one_xiv = 1.0 - xiv;
    d.v = 2.0 * y.v * cs2 * (1.0 - v2) * one_xiv;
    // dw = w/(1-v^2) * (xi-v)/(1-xi v) * (1/cs2 + 1) * dv; with dv = 2 v cs2 (1-v^2)(1-xi v) the two quotients cancel
    // exactly: dw = 2 v w (xi - v) (1 + cs2)  -- no division on the hot path (same value up to rounding)
    d.w = 2.0 * y.v * y.w * xi_v * (1.0 + cs2);
    d.xi = y.xi * (xi_v * xi_v - cs2 * one_xiv * one_xiv);
}

// numpy.linspace(0, t_end, n)[i]: i*step with the last sample forced to the end point.
PTT_HD double grid_time(const double t_end, const int n, const int i) {
    if (i >= n - 1) return t_end;
    return (double)i * (t_end / (double)(n - 1));
}

// Dormand-Prince 5(4) with the 4th-order continuous extension (Hairer, Norsett & Wanner, "dopri5").
// Tolerances are far tighter than LSODA's 1.5e-8 defaults so that index decisions agree with the reference
// except on razor edges.
struct Dopri5 {
    static constexpr double RTOL = 1e-10;
    static constexpr double ATOL = 1e-13;

    double cs2;
    double t, h, t_end;
    State y, k1;                 // current state and its derivative (FSAL)
    // dense-output data of the last accepted step [t0, t0 + hd]
    double t0, hd;
    State r1, r2, r3, r4, r5;
    int nsteps, nrej;
    bool failed;

    PTT_HD void init(const double cs2_, const State& y0, const double t_end_) {
        cs2 = cs2_;
        t = 0.0;
        t_end = t_end_;
        y = y0;
        rhs(cs2, y, k1);
        nsteps = 0;
        nrej = 0;
        failed = false;
        // Initial step: small fraction of the interval, bounded by the local time scale.
        const double sc_v = ATOL + RTOL * fabs(y.v), sc_w = ATOL + RTOL * fabs(y.w), sc_x = ATOL + RTOL * fabs(y.xi);
        const double d0 = sqrt((y.v * y.v / (sc_v * sc_v) + y.w * y.w / (sc_w * sc_w) + y.xi * y.xi / (sc_x * sc_x)) / 3.0);
        const double d1 = sqrt((k1.v * k1.v / (sc_v * sc_v) + k1.w * k1.w / (sc_w * sc_w) + k1.xi * k1.xi / (sc_x * sc_x)) / 3.0);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        const double span = fabs(t_end);
        if (h0 > span) h0 = span;
        if (h0 < 1e-12 * span) h0 = 1e-12 * span;
        h = (t_end < 0.0) ? -h0 : h0;
        t0 = 0.0;
        hd = 0.0;
    }

    PTT_HD bool done() const { return failed || t == t_end; }

    // One accepted step (internally retries rejected ones). Returns false on failure.
    PTT_HD bool step() {
        const double c2 = 0.2, c3 = 0.3, c4 = 0.8, c5 = 8.0 / 9.0;
        const double a21 = 0.2;
        const double a31 = 3.0 / 40.0, a32 = 9.0 / 40.0;
        const double a41 = 44.0 / 45.0, a42 = -56.0 / 15.0, a43 = 32.0 / 9.0;
        const double a51 = 19372.0 / 6561.0, a52 = -25360.0 / 2187.0, a53 = 64448.0 / 6561.0, a54 = -212.0 / 729.0;
        const double a61 = 9017.0 / 3168.0, a62 = -355.0 / 33.0, a63 = 46732.0 / 5247.0, a64 = 49.0 / 176.0,
                     a65 = -5103.0 / 18656.0;
        const double a71 = 35.0 / 384.0, a73 = 500.0 / 1113.0, a74 = 125.0 / 192.0, a75 = -2187.0 / 6784.0,
                     a76 = 11.0 / 84.0;
        const double e1 = 71.0 / 57600.0, e3 = -71.0 / 16695.0, e4 = 71.0 / 1920.0, e5 = -17253.0 / 339200.0,
                     e6 = 22.0 / 525.0, e7 = -1.0 / 40.0;
        const double d1 = -12715105075.0 / 11282082432.0, d3 = 87487479700.0 / 32700410799.0,
                     d4 = -10690763975.0 / 1880347072.0, d5 = 701980252875.0 / 199316789632.0,
                     d6 = -1453857185.0 / 822651844.0, d7 = 69997945.0 / 29380423.0;
        (void)c2; (void)c3; (void)c4; (void)c5;   // autonomous system: stage times are not needed

        for (int attempt = 0; attempt < 60; ++attempt) {
            bool last = false;
            double hs = h;
            if ((t + 1.01 * hs - t_end) * ((t_end < 0.0) ? -1.0 : 1.0) >= 0.0) {
                hs = t_end - t;
                last = true;
            }
            State k2, k3, k4, k5, k6, k7, yt, y1;
            yt.v = y.v + hs * a21 * k1.v; yt.w = y.w + hs * a21 * k1.w; yt.xi = y.xi + hs * a21 * k1.xi;
            rhs(cs2, yt, k2);
            yt.v = y.v + hs * (a31 * k1.v + a32 * k2.v);
            yt.w = y.w + hs * (a31 * k1.w + a32 * k2.w);
            yt.xi = y.xi + hs * (a31 * k1.xi + a32 * k2.xi);
            rhs(cs2, yt, k3);
            yt.v = y.v + hs * (a41 * k1.v + a42 * k2.v + a43 * k3.v);
            yt.w = y.w + hs * (a41 * k1.w + a42 * k2.w + a43 * k3.w);
            yt.xi = y.xi + hs * (a41 * k1.xi + a42 * k2.xi + a43 * k3.xi);
            rhs(cs2, yt, k4);
            yt.v = y.v + hs * (a51 * k1.v + a52 * k2.v + a53 * k3.v + a54 * k4.v);
            yt.w = y.w + hs * (a51 * k1.w + a52 * k2.w + a53 * k3.w + a54 * k4.w);
            yt.xi = y.xi + hs * (a51 * k1.xi + a52 * k2.xi + a53 * k3.xi + a54 * k4.xi);
            rhs(cs2, yt, k5);
            yt.v = y.v + hs * (a61 * k1.v + a62 * k2.v + a63 * k3.v + a64 * k4.v + a65 * k5.v);
            yt.w = y.w + hs * (a61 * k1.w + a62 * k2.w + a63 * k3.w + a64 * k4.w + a65 * k5.w);
            yt.xi = y.xi + hs * (a61 * k1.xi + a62 * k2.xi + a63 * k3.xi + a64 * k4.xi + a65 * k5.xi);
            rhs(cs2, yt, k6);
            y1.v = y.v + hs * (a71 * k1.v + a73 * k3.v + a74 * k4.v + a75 * k5.v + a76 * k6.v);
            y1.w = y.w + hs * (a71 * k1.w + a73 * k3.w + a74 * k4.w + a75 * k5.w + a76 * k6.w);
            y1.xi = y.xi + hs * (a71 * k1.xi + a73 * k3.xi + a74 * k4.xi + a75 * k5.xi + a76 * k6.xi);
            rhs(cs2, y1, k7);

            const double ev = hs * (e1 * k1.v + e3 * k3.v + e4 * k4.v + e5 * k5.v + e6 * k6.v + e7 * k7.v);
            const double ew = hs * (e1 * k1.w + e3 * k3.w + e4 * k4.w + e5 * k5.w + e6 * k6.w + e7 * k7.w);
            const double ex = hs * (e1 * k1.xi + e3 * k3.xi + e4 * k4.xi + e5 * k5.xi + e6 * k6.xi + e7 * k7.xi);
            const double sv = ATOL + RTOL * fmax(fabs(y.v), fabs(y1.v));
            const double sw = ATOL + RTOL * fmax(fabs(y.w), fabs(y1.w));
            const double sx = ATOL + RTOL * fmax(fabs(y.xi), fabs(y1.xi));
            // err^2 (the RMS norm squared): the accept test and the step-size factor need no square root
            const double err2 = ((ev / sv) * (ev / sv) + (ew / sw) * (ew / sw) + (ex / sx) * (ex / sx)) * (1.0 / 3.0);
            if (!(err2 == err2)) {         // NaN: the trajectory left the domain (v -> 1 or xi*v -> 1)
                failed = true;
                return false;
            }
            // step-size factor 0.9 err^(-1/5): single precision is plenty for a step-size heuristic
            double fac = (err2 > 0.0) ? 0.9 * (double)powf((float)fmin(fmax(err2, 1e-30), 1e30), -0.1f) : 5.0;
            if (fac < 0.2) fac = 0.2;
            if (fac > 5.0) fac = 5.0;
            if (err2 <= 1.0) {
                // accept; build the continuous extension (Hairer's contd5)
                t0 = t;
                hd = hs;
                r1 = y;
                r2.v = y1.v - y.v; r2.w = y1.w - y.w; r2.xi = y1.xi - y.xi;
                r3.v = hs * k1.v - r2.v; r3.w = hs * k1.w - r2.w; r3.xi = hs * k1.xi - r2.xi;
                r4.v = r2.v - hs * k7.v - r3.v; r4.w = r2.w - hs * k7.w - r3.w; r4.xi = r2.xi - hs * k7.xi - r3.xi;
                r5.v = hs * (d1 * k1.v + d3 * k3.v + d4 * k4.v + d5 * k5.v + d6 * k6.v + d7 * k7.v);
                r5.w = hs * (d1 * k1.w + d3 * k3.w + d4 * k4.w + d5 * k5.w + d6 * k6.w + d7 * k7.w);
                r5.xi = hs * (d1 * k1.xi + d3 * k3.xi + d4 * k4.xi + d5 * k5.xi + d6 * k6.xi + d7 * k7.xi);
                y = y1;
                k1 = k7;
                t = last ? t_end : t + hs;
                if (nrej > 0 && fac > 1.0) fac = 1.0;   // no growth right after a rejection
                nrej = 0;
                h = hs * fac;
                ++nsteps;
                PTT_ODE_STEP_HOOK(true)
                return true;
            }
            // reject
            ++nrej;
            PTT_ODE_STEP_HOOK(false)
            h = hs * fac;
            if (fabs(h) < 1e-15 * fmax(1.0, fabs(t))) {
                failed = true;
                return false;
            }
        }
        failed = true;
        return false;
    }

    // Continuous extension inside the last accepted step.
    // inv_hd = 1 / hd, computed once per step by the caller
    PTT_HD State dense(const double tq, const double inv_hd) const {
        const double th = (tq - t0) * inv_hd;
        const double th1 = 1.0 - th;
        State o;
        o.v = r1.v + th * (r2.v + th1 * (r3.v + th * (r4.v + th1 * r5.v)));
        o.w = r1.w + th * (r2.w + th1 * (r3.w + th * (r4.w + th1 * r5.w)));
        o.xi = r1.xi + th * (r2.xi + th1 * (r3.xi + th * (r4.xi + th1 * r5.xi)));
        return o;
    }
    // xi alone (the same expression as in dense())
    PTT_HD double dense_xi(const double tq, const double inv_hd) const {
        const double th = (tq - t0) * inv_hd;
        const double th1 = 1.0 - th;
        return r1.xi + th * (r2.xi + th1 * (r3.xi + th * (r4.xi + th1 * r5.xi)));
    }
    // Upper bound of xi over the last accepted step: theta (1 - theta) <= 1/4 on [0, 1].
    PTT_HD double xi_upper_bound() const {
        return r1.xi + fmax(r2.xi, 0.0) + 0.25 * (fabs(r3.xi) + fabs(r4.xi) + fabs(r5.xi));
    }
};

// Grid samples inside the last accepted step: sample(k) is what the ordered walk below hands to visit(k, .).
struct StepSampler {
    const Dopri5& s;
    double dt, t_end, inv_hd;
    int n;
    PTT_HD State operator()(const int k) const {
        const double tk = (k >= n - 1) ? t_end : (double)k * dt;
        return (tk == s.t) ? s.y : s.dense(tk, inv_hd);
    }
    PTT_HD double xi(const int k) const {
        const double tk = (k >= n - 1) ? t_end : (double)k * dt;
        return (tk == s.t) ? s.y.xi : s.dense_xi(tk, inv_hd);
    }
    PTT_HD double xi_upper_bound() const { return s.xi_upper_bound(); }
};

// Optional member of a visitor:
//   int search(int i_lo, int i_hi, const StepSampler& sample)
// called for a step that holds the grid samples i_lo..i_hi (more than one) and was not skipped.  A visitor that can find
// its stopping sample without seeing every sample in order (bisection) does so and returns its index (it then has the
// state visit() would have left behind); -1 = walk the samples one by one; -2 = only the last one, i_hi, needs to be seen;
// -3 = the visitor has seen i_lo..i_hi itself and goes on.
template <class Visitor>
PTT_HD auto visitor_search(Visitor& v, const int i_lo, const int i_hi, const StepSampler& sm, int)
    -> decltype(v.search(i_lo, i_hi, sm)) {
    return v.search(i_lo, i_hi, sm);
}
template <class Visitor>
PTT_HD int visitor_search(Visitor&, const int, const int, const StepSampler&, long) { return -1; }

// Walks the tau-grid t_i = linspace(0, t_end, n)[i] in order and hands samples to a visitor:
//   bool visit(int i, const State& y)   -- return false to stop early
//   bool skip(const State& y_end)       -- true if no event can lie inside a step that ends in y_end; then only the
//                                          LAST grid sample of that step is visited (so that the sample preceding an
//                                          event in the next step is always known). Return false to see every sample.
// Returns the number of grid samples covered, or -1 if the integrator failed
// (the reference raises RuntimeError("Integration failed"), integrate.py:133).
template <class Visitor>
PTT_HD int integrate_grid(const double cs2, const State& y0, const double t_end, const int n, Visitor& visit) {
    if (!visit.visit(0, y0)) return 1;
    if (n < 2) return 1;
    if (t_end == 0.0) {
        for (int i = 1; i < n; ++i)
            if (!visit.visit(i, y0)) return i + 1;
        return n;
    }
    Dopri5 s;
    s.init(cs2, y0, t_end);
    const double sgn = (t_end < 0.0) ? -1.0 : 1.0;
    const double dt = t_end / (double)(n - 1);          // grid_time(t_end, n, k) = k * dt below the last sample
    const double inv_dt = 1.0 / dt;
    int i = 1;
    while (i < n) {
        if (!s.step()) return -1;
        // last grid index inside (t0, t]: an estimate, then made exact against the grid times themselves
        int i_hi;
        if (s.t == t_end) {
            i_hi = n - 1;
        } else {
            i_hi = (int)(s.t * inv_dt);
            if (i_hi > n - 1) i_hi = n - 1;
            while (i_hi + 1 < n && (((i_hi + 1 >= n - 1) ? t_end : (double)(i_hi + 1) * dt) - s.t) * sgn <= 0.0) ++i_hi;
            while (i_hi >= i && (((i_hi >= n - 1) ? t_end : (double)i_hi * dt) - s.t) * sgn > 0.0) --i_hi;
        }
        if (i_hi < i) continue;
        if (i > 1 && i_hi > i && visit.skip(s.y)) i = i_hi;
        const StepSampler sample{s, dt, t_end, 1.0 / s.hd, n};
#ifndef PTT_NO_STEP_SEARCH      // (test switch: the ordered walk everywhere)
        if (i_hi > i) {
            const int i_stop = visitor_search(visit, i, i_hi, sample, 0);
            if (i_stop >= 0) { PTT_ODE_SAMPLE_HOOK(true) return i_stop + 1; }
            if (i_stop == -2) i = i_hi;
            if (i_stop == -3) { i = i_hi + 1; continue; }
        }
#endif
        for (; i <= i_hi; ++i) {
            PTT_ODE_SAMPLE_HOOK(false)
            if (!visit.visit(i, sample(i))) return i + 1;
        }
    }
    return n;
}

}  // namespace ptt
