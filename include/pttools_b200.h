/*
 * pttools_b200 -- C ABI of the B200-native hot path of PTtools (fluid-shell solve + Sound Shell Model spectrum).
 *
 * Every entry point is a batched, array-in/array-out replacement of one reference operator; the reference
 * interface it replaces is cited (paths relative to the reference repository).  Conventions:
 *   - plain pointers and sizes only; all floating point data is FP64, indices/status codes are int32;
 *   - `mem` says where the caller's buffers live: PTT_MEM_HOST (the library stages them through device memory
 *     and returns after the results are back on the host) or PTT_MEM_DEVICE (pointers are device pointers on the
 *     current CUDA device; the work is enqueued on `stream` (a cudaStream_t, NULL = default stream) and the call
 *     returns without synchronising);
 *   - the return value is 0 on success or a negative PTT_E_* code for launch/argument errors
 *     (ptt_last_error() gives the message); numerical failures are reported PER POINT in `status[]`
 *     (PTT_ST_*), with NaN-filled outputs where the reference returns nan_arr / raises;
 *   - all output buffers are caller-owned.
 */
#ifndef PTTOOLS_B200_H
#define PTTOOLS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTT_MEM_HOST 0
#define PTT_MEM_DEVICE 1

/* return codes */
#define PTT_OK 0
#define PTT_E_ARG (-1)
#define PTT_E_CUDA (-2)
#define PTT_E_NOMEM (-3)

/* per-point status (mirrors the reference's flags: Bubble.solver_failed / no_solution_found, nan_arr returns,
 * RuntimeError("Integration failed"); pttools/bubble/bubble.py:123-131, fluid_bag.py:61-73, integrate.py:133) */
#define PTT_ST_OK 0
#define PTT_ST_NO_SOLUTION 1
#define PTT_ST_INTEGRATION_FAILED 2
#define PTT_ST_ROOT_NOT_CONVERGED 3
#define PTT_ST_INVALID_INPUT 4
#define PTT_ST_SOLVER_FAILED 5

/* solution types (pttools/bubble/boundary.py:37-62) */
#define PTT_SOL_SUB_DEF 0
#define PTT_SOL_HYBRID 1
#define PTT_SOL_DETON 2
#define PTT_SOL_ERROR (-1)

/* Equation of state.  Replaces the C function pointer `double cs2(double w, double phase)` that the reference
 * compiles per model (pttools/type_hints.py:33-37, models/model.py:620-639): both in-scope models have a
 * piecewise-constant sound speed, so a POD is enough.  p(w,phase) = w/mu - V,  e(w,phase) = (1-1/mu) w + V,
 * mu = 1 + 1/cs2  (models/bag.py, models/const_cs.py:585-621). */
typedef struct ptt_eos {
    int32_t kind;      /* 0 = BagModel, 1 = ConstCSModel */
    int32_t reserved;
    double css2, csb2; /* sound speed squared, symmetric / broken phase */
    double a_s, a_b;   /* pressure prefactors */
    double V_s, V_b;   /* potentials */
    double T_ref;
} ptt_eos_t;

const char* ptt_last_error(void);
int ptt_version(void);
/* capacity (in doubles) of one profile row for a given n_xi; the assembled shell is row[off .. off+npts) */
int ptt_profile_row_capacity(int n_xi);

/* --- Seam A: integrator plug-in ------------------------------------------------------------------------------
 * fluid_integrate_param(v0, w0, xi0, phase, t_end, n_xi, df_dtau_ptr, method)   pttools/bubble/integrate.py:81-135
 * Batched: lane i integrates from (v0[i], w0[i], xi0[i]) over t = linspace(0, t_end[i], n_xi);
 * outputs are [n, n_xi] row-major.  status[i] = PTT_ST_INTEGRATION_FAILED where the reference raises. */
int ptt_integrate_batch(int n, const double* v0, const double* w0, const double* xi0, const double* phase,
                        const double* t_end, int n_xi, const ptt_eos_t* eos,
                        double* out_v, double* out_w, double* out_xi, int32_t* status, int mem, void* stream);

/* --- Seam B: bag-model shell operators -----------------------------------------------------------------------
 * alpha_n_max_deflagration_bag(v_wall)                                           pttools/bubble/alpha.py:44-50,71 */
int ptt_bag_alpha_n_max(int n, const double* v_wall, double* out, int32_t* status, int mem, void* stream);

/* identify_solution_type_bag + find_alpha_plus_bag                 transition.py:19-44, alpha.py:305-331,361-388
 * an_max may be NULL (then it is computed per point).  Outputs: alpha_plus[n], sol_type[n], status[n]. */
int ptt_bag_find_alpha_plus(int n, const double* v_wall, const double* alpha_n, int n_xi, const double* an_max,
                            double* alpha_plus, int32_t* sol_type, int32_t* status, int mem, void* stream);

/* sound_shell_alpha_plus(v_wall, alpha_plus, sol_type, n_xi, w_n)                       fluid_bag.py:83-222
 * rows_*: [n, row_stride] with row_stride >= ptt_profile_row_capacity(n_xi); shell i is
 * rows_*[i*row_stride + off[i] .. + npts[i]).  scalars[n,4] = {xi_shock, w_n_raw, n_wall, reserved}. */
int ptt_bag_profiles(int n, const double* v_wall, const double* alpha_plus, const int32_t* sol_type, int n_xi,
                     const double* w_n, double* rows_v, double* rows_w, double* rows_xi, int64_t row_stride,
                     int32_t* off, int32_t* npts, double* scalars, int32_t* status, int mem, void* stream);

/* sound_shell_bag(v_wall, alpha_n, n_xi) for a whole sweep                                fluid_bag.py:33-79
 * = alpha_n_max (once per distinct v_wall is the caller's business: pass an_max or NULL) + root + profile. */
int ptt_sweep_shells_bag(int n, const double* v_wall, const double* alpha_n, int n_xi, const double* an_max,
                         double* rows_v, double* rows_w, double* rows_xi, int64_t row_stride,
                         int32_t* off, int32_t* npts, double* alpha_plus, int32_t* sol_type, double* scalars,
                         int32_t* status, int mem, void* stream);

/* sound_shell_generic(model, v_wall, alpha_n, sol_type, wn, ...) for a whole sweep      pttools/bubble/fluid.py:848-1052
 * One row of pg per point (models may differ from point to point):
 *   pg[n,16] = {v_wall, alpha_n, sol_type, wn, v_cj, guess_vp, guess_vp_tilde, guess_wp, guess_wm,
 *               css2, csb2, V_s, V_b, model_name_is_bag, s_coef_s, s_coef_b}
 *   s_coef: entropy density s(w) = s_coef * w^(1 - 1/mu) of the phase, s_coef = (mu a)^(1/mu) T_ref^(4/mu - 1); read by
 *   ptt_generic_rescue only (entropy-flux test of the scan starts, boundary.py:65-85)
 * (wn, v_cj and the solution type are closed-form host quantities of the model: models/bag.py:316, const_cs.py:633-737,
 * chapman_jouguet.py:151-264; the guesses come from the fluid-reference table, fluid.py:925-949).
 * Outputs: profile rows ([n,row_stride], row_stride >= ptt_generic_row_capacity(n_xi)), off/npts, and
 *   scalars[n,16] = {vp, vm, vp_tilde, vm_tilde, v_sh, vm_sh, vm_tilde_sh, wp, wm, wm_sh, v_cj, wn, wn_estimate,
 *                    solver_failed, n_resid (residual evaluations of the root search, diagnostics), 0}  -- the
 *                    reference's 17-tuple (fluid.py:1052).
 * status[i] = PTT_ST_SOLVER_FAILED where the reference sets solver_failed. */
int ptt_generic_row_capacity(int n_xi);
int ptt_sweep_shells_generic(int n, const double* pg, int n_xi, double t_end, int thin_shell_limit, double wn_rtol,
                             double* rows_v, double* rows_w, double* rows_xi, int64_t row_stride, int32_t* off,
                             int32_t* npts, double* scalars, int32_t* status, int mem, void* stream);

/* The retries of the reference for points whose first root search failed (status PTT_ST_SOLVER_FAILED after
 * ptt_sweep_shells_generic):  fsolve_vary's two varied starts (pttools/speedup/solvers.py:12-55, called at
 * fluid.py:632 and :720) and, for hybrids, the 20-point backup scan (fluid.py:731-795).  idx[n_f] = rows of pg to retry;
 * outputs are COMPACT (entry f belongs to point idx[f]) with the layout of ptt_sweep_shells_generic; rescued[f] = 1
 * where a candidate start converged and a new profile was written (scalars[f,15] = 1 + its index in the reference's
 * order), 0 where the reference, too, keeps the result of its first search.  Device pointers. */
int ptt_generic_rescue(int n_f, const int32_t* idx, const double* pg, int n_xi, double t_end, int thin_shell_limit,
                       double wn_rtol, double* rows_v, double* rows_w, double* rows_xi, int64_t row_stride, int32_t* off,
                       int32_t* npts, double* scalars, int32_t* status, int32_t* rescued, void* stream);

/* Starting guesses: nearest valid entry of the fluid-reference table with the same solution type, for a batch
 * (FluidReference.get -> NearestNDInterpolator, pttools/bubble/fluid_reference.py:60-62, 152-163).  The table entries
 * of set s are cx/cy[set_off[s] .. set_off[s+1]); sel[i] picks the set of query i.  out_idx[i] = index of the nearest
 * entry (squared Euclidean distance in (v_wall, alpha_n), first one on ties), -1 for an empty or unknown set.
 * Device pointers. */
int ptt_nearest_batch(int n, const double* qx, const double* qy, const int32_t* sel, int n_sets, const int32_t* set_off,
                      const double* cx, const double* cy, int32_t* out_idx, void* stream);

/* props.v_and_w_from_solution for a batch of bag shells (pttools/bubble/props.py:52-87): the entries of the
 * fluid-reference starting-guess table (fluid_reference.py:166-196).  Device pointers.
 * out[n,8] = {vp, vm, vp_tilde, vm_tilde, wp, wm, wn, 0}. */
int ptt_shell_wall_values(int n, const double* rows_v, const double* rows_w, const double* rows_xi, int64_t row_stride,
                          const int32_t* off, const int32_t* npts, const double* v_wall, const int32_t* sol_type,
                          double* out, void* stream);

/* --- Seam C: Sound Shell Model array operators (all pointers DEVICE pointers, enqueued on `stream`) -----------
 * Grids that depend only on (y, cs, nt, ...) are built once per sweep by the host mirror (numpy expressions
 * identical to the reference's) and passed in; the kernels do the per-profile arithmetic. */

/* sin_transform(z, xi, f, z_st_thresh)                                     pttools/ssmtools/calculators.py:119-190
 * Batched over P ragged profiles: profile p is xi/f[p*stride + off[p] .. + npts[p]).
 * frac[n_z]: 0 = exact trapezoid only, 1 = asymptotic formula only, in between = cubic blend weight
 * (calculators.py:136-153; depends only on z, so the host computes it).  out[P, n_z]. */
int ptt_sin_transform_batch(int P, const double* xi, const double* f, int64_t stride, const int32_t* off,
                            const int32_t* npts, int n_z, const double* z, const double* frac,
                            double* out, void* stream);

/* Same operator for profiles that live on 0 <= xi <= 1 (fluid shells do) with the exact branch restricted to
 * z <= 51.2 (z_st_thresh = 50 is): the trapezoid sum is evaluated from block moments of weight*f instead of one sine
 * per (z, xi) pair - same sum, O(n_xi + 256 n_z) instead of O(n_xi n_z) per profile.  Rows that violate either
 * precondition come back as NaN. */
int ptt_sin_transform_unit_batch(int P, const double* xi, const double* f, int64_t stride, const int32_t* off,
                            const int32_t* npts, int n_z, const double* z, const double* frac,
                            double* out, void* stream);

typedef struct ptt_a2_plan {
    int32_t n_z;          /* total number of wavenumbers (all lookup grids concatenated) */
    int32_t n_seg;        /* number of grids; gradients never cross a grid boundary */
    const int32_t* seg;   /* [n_seg+1] start index of each grid in z */
    const double* z;      /* [n_z] */
    const double* frac;   /* [n_z] blend fraction, see ptt_sin_transform_batch */
    int32_t nxi;          /* resampling size for lambda (ssm.py:40, default 2000) */
    int32_t phase_rule;   /* 0: xi < v_wall (boundary.get_phase, old bag path); 1: props.find_phase (Bubble path) */
    const double* xi_re;  /* [nxi] = linspace(0, 1-1/nxi, nxi) */
    double z_exact_max;   /* largest z that takes the exact (trapezoid) branch, i.e. with frac < 1 */
    int32_t force_direct; /* 1: one sine per (z, xi) pair (validation); 0: block-moment evaluation of the same sum */
    int32_t reserved;
} ptt_a2_plan_t;

/* a2_e_conserving / a2_e_conserving_bag                                       pttools/ssmtools/ssm.py:35-140
 * per-point params pp[P,8] = {v_wall, cs, e_coef_s, e_coef_b, V_s, V_b, 0, 0} with e(w,phase) = e_coef*w + V.
 * work: [P, work_stride] doubles of scratch with work_stride >= ptt_a2_work_stride(row_stride, nxi, n_z) (weighted
 * profiles, resampled lambda, envelopes, moment pyramids, f and l per wavenumber).  Outputs a2/fp2_2/lam2: [P, n_z] (the
 * last two may be NULL).  Any batch size (batches above 65535 profiles are processed in slices). */
int ptt_a2_batch(const ptt_a2_plan_t* plan, int P, const double* rows_v, const double* rows_w, const double* rows_xi,
                 int64_t row_stride, const int32_t* off, const int32_t* npts, const double* pp,
                 double* work, int64_t work_stride, double* a2, double* fp2_2, double* lam2, void* stream);

typedef struct ptt_sdv_plan {
    int32_t n_q;          /* size of qT_lookup */
    int32_t n_z;          /* evaluation points */
    int32_t nt;           /* T-tilde samples */
    int32_t z_tile;       /* evaluation points per CTA (host-chosen so that the lookup window fits in shared memory) */
    const double* qT;     /* [n_q]  10**arange(...)   spectrum.py:515 */
    const double* z;      /* [n_z] */
    const double* zterm;  /* [n_z]  (log10 z - log10 qT[0]) / dlog10 */
    const double* T;      /* [nt]   logspace(log10 0.01, log10 20, nt)  spectrum.py:473 */
    const double* LT;     /* [nt]   log10(T) / dlog10 */
    const double* W;      /* [nt]   trapezoid weight * T^6 * nu(T) */
    double inv_dlog;      /* 1 / dlog10z */
    int32_t smem_bytes;   /* dynamic shared memory per CTA */
    int32_t reserved;
    const double* omega_ab; /* [3 nt + 1, 2] or NULL: (alpha, beta) of the knot weight omega(x) = alpha + beta X/qT[e] on
                             * every piece between consecutive kinks of {LT_j - 1, LT_j, LT_j + 1} (ptt_spec_den_v_group) */
} ptt_sdv_plan_t;

/* spec_den_v core                                                        pttools/ssmtools/spectrum.py:451-486
 * a2: [P, a2_stride] lookup values on qT; out: [P, n_z]. */
int ptt_spec_den_v_batch(const ptt_sdv_plan_t* plan, int P, const double* a2, int64_t a2_stride,
                         const double* v_wall, double* out, int64_t out_stride, void* stream);

/* spec_den_v for n_p profiles that SHARE v_wall, as a banded FP64 matrix product (same linear-interpolation weights as
 * np.interp in spectrum.py:451-459, summed per knot first).  kw: band width in knots per 128-row tile (multiple of 16),
 * e_base[n_tiles]: first knot of each tile for this v_wall (both computed by the host from the plan),
 * omega: device scratch of n_tiles*128*kw + 2 n_q doubles.  out[p, i], like ptt_spec_den_v_batch. */
int ptt_spec_den_v_group(const ptt_sdv_plan_t* plan, int n_p, const double* a2, int64_t a2_stride, double v_wall,
                         int kw, const int32_t* e_base, double* omega, double* out, int64_t out_stride, void* stream);
/* The same with flags.  PTT_SDV_ALIGNED_BAND: the caller promises that a2 + e_base[t] lies on a 16-byte boundary for every
 * row tile t (e_base[t] may then be -1: a knot before the table, weight 0) and that a2_stride is even; the operands are then
 * moved by tensor-map copies (TMA) into a shared-memory ring instead of through registers. */
#define PTT_SDV_ALIGNED_BAND 1
int ptt_spec_den_v_group_ex(const ptt_sdv_plan_t* plan, int n_p, const double* a2, int64_t a2_stride, double v_wall, int kw,
                            const int32_t* e_base, double* omega, double* out, int64_t out_stride, int flags, void* stream);

typedef struct ptt_gw_plan {
    int32_t n_zl;         /* size of z_lookup */
    int32_t n_y;
    int32_t n_int;        /* integration samples per y (nz_int) */
    int32_t y_tile;
    const double* z_lookup;  /* [n_zl] log-uniform */
    const double* y;      /* [n_y] */
    const double* yterm;  /* [n_y]  (log10 y - log10 z_lookup[0]) / dlog10 */
    const double* L1;     /* [n_int] log10(zeta_k) / dlog10 */
    const double* L2;     /* [n_int] log10(zeta_+ + zeta_- - zeta_k) / dlog10 */
    const double* Z1;     /* [n_int] zeta_k */
    const double* Z2;     /* [n_int] zeta_+ + zeta_- - zeta_k */
    const double* G;      /* [n_int] trapezoid weight * integrand kernel */
    double prefactor;     /* 0.75 Gamma^2 ((1-cs^2)/cs^2)^2 / (4 pi cs) */
    int32_t smem_bytes;
    int32_t reserved;
} ptt_gw_plan_t;

/* spec_den_gw_scaled core + pow_spec                                    spectrum.py:327-368, 230-240
 * pv: [P, pv_stride] P_v on z_lookup; slf[P] source lifetime factor; scale[P] = r_star * F_gw0 * suppression
 * (omgw0_ssm.py:123-131) or NULL.  Outputs [P, n_y]: sd_gw (may be NULL), pow_gw, omgw0 (NULL if scale NULL). */
int ptt_spec_den_gw_batch(const ptt_gw_plan_t* plan, int P, const double* pv, int64_t pv_stride, const double* slf,
                          const double* scale, double* sd_gw, double* pow_gw, double* omgw0, void* stream);

/* spec_den_gw_scaled on an arbitrary ascending z_lookup and any y (np.interp by binary search per sample; the general
 * form of spectrum.py:402-430 -- the log-uniform grids of gen_lookup / np.logspace take the two entries around this one).
 * Z1[n_int] = zeta_k, Z2 = zeta_+ + zeta_- - zeta_k, G = trapezoid weight * integrand kernel, as in ptt_gw_plan_t. */
int ptt_spec_den_gw_general_batch(int n_zl, const double* z_lookup, int n_y, const double* y, int n_int, const double* Z1,
                                  const double* Z2, const double* G, double prefactor, int P, const double* pv,
                                  int64_t pv_stride, const double* slf, const double* scale, double* sd_gw, double* pow_gw,
                                  double* omgw0, void* stream);

/* spec_den_gw_scaled, cell formulation for large batches (same trapezoid sum over the same samples, regrouped by runs of
 * samples that share both np.interp segments and collected per segment of the first lookup; spectrum.py:327-368).
 * The tables depend on the plan only:
 *   cap   = ptt_gw_cells_capacity(n_zl, n_int, lrange), lrange = max(L1, L2) - min(L1, L2) of the plan
 *   ptt_gw_cells_build fills ncell[n_y] (< 0: not usable, fall back to ptt_spec_den_gw_batch), meta[n_y, cap + 2],
 *   coef[n_y, ptt_gw_cells_coef_capacity(cap), 2]                                                   (device memory)
 * table: device scratch of ptt_gw_cells_table_size(n_zl, P) doubles.  Outputs as ptt_spec_den_gw_batch. */
typedef struct ptt_gw_cells {
    int32_t cap;
    int32_t reserved;
    const int32_t* ncell;
    const int32_t* meta;
    const double* coef;
} ptt_gw_cells_t;
int ptt_gw_cells_capacity(int n_zl, int n_int, double lrange);
int ptt_gw_cells_coef_capacity(int cap);
int ptt_gw_cells_build(const ptt_gw_plan_t* plan, int cap, int32_t* ncell, int32_t* meta, double* coef, void* stream);
int64_t ptt_gw_cells_table_size(int n_zl, int P);
int ptt_spec_den_gw_cells_batch(const ptt_gw_plan_t* plan, const ptt_gw_cells_t* cells, int P, const double* pv,
                                int64_t pv_stride, const double* slf, const double* scale, double* table,
                                double* sd_gw, double* pow_gw, double* omgw0, void* stream);

/* pow_spec(z, spec_den) = z^3/(2 pi^2) spec_den, row-wise                         spectrum.py:230-240 */
int ptt_pow_spec_batch(int P, int n_z, const double* z, const double* spec_den, int64_t stride, double* out,
                       void* stream);

/* scratch size (doubles per profile) that ptt_a2_batch needs */
int ptt_a2_work_stride(int row_stride, int nxi, int n_z);

/* Thermodynamic integrals of a batch of shells (trapezoid rule over xi^3)        pttools/bubble/thermo.py:139-221
 * pp2[P,12] = {v_wall, mu_s, mu_b, V_s, V_b, a_s, a_b, T_ref, phase_rule, 0, 0, 0}  (mu = 1 + 1/cs2)
 * out[P,8]  = {int w v^2 gamma^2 dxi^3, int (theta - theta_n) dxi^3, int (w - w_n) dxi^3, int (s - s_n) dxi^3,
 *              wbar, i_wall, n_broken, w[-1]};  device pointers, enqueued on `stream`. */
int ptt_profile_integrals_batch(int P, const double* rows_v, const double* rows_w, const double* rows_xi,
                                int64_t row_stride, const int32_t* off, const int32_t* npts, const double* pp2,
                                double* out, void* stream);

/* --- Seam B, the performance seam: a whole sweep in one call -------------------------------------------------
 * create_spectra(model, v_walls, alpha_ns, spectrum_kwargs) / create_bubbles(...)     pttools/analysis/parallel.py:89-187
 *   -> run_parallel: one future per (v_wall, alpha_n) point, pickled Bubble / Spectrum objects back
 *                                                                                      pttools/speedup/parallel.py:67-196
 * Here: per point Bubble(model, v_wall, alpha_n).solve() (bubble.py:232-352: sound_shell_generic, the reference's
 * retries, the thermodynamic integrals), SSMSpectrum.compute (ssmtools/spectrum.py:96-136) and
 * Spectrum.omgw0 (omgw0/omgw0_ssm.py:123-131), for all n points, chunked against device memory, with the groups of
 * equal v_wall of K6' on side streams and the result tables streamed out per chunk.
 *
 * The plan is a POD of views: the four sub-plans of seam C plus the few host numbers the driver needs to lay out the
 * bands of K6'.  ptt_ssm_plan_create() builds one (grids with the reference's expressions, spectrum.py:169-186,
 * 503-515, 473; calculators.py:100,136-153); a caller that already holds the sub-plans may fill it in itself. */
typedef struct ptt_sweep_plan {
    ptt_a2_plan_t a2;          /* all lookup grids concatenated: [pass 0 | pass 1] */
    int32_t n_pass;            /* 1 or 2; the LAST pass is the z_lookup pass that feeds the GW integral */
    int32_t has_y_pass;        /* pass 0 evaluates spec_den_v on y itself (-> pow_v) */
    ptt_sdv_plan_t sdv[2];
    ptt_gw_plan_t gw;
    ptt_gw_cells_t cells;      /* cells.cap == 0: no cell tables, the per-sample kernel is used */
    const double* h_zterm[2];  /* HOST copies of sdv[i].zterm */
    double lt_first[2], lt_last[2], dl[2];   /* LT[0], LT[nt-1], dlog10 of each pass */
    int32_t seg[4];            /* start of each pass' grid in a2 (seg[n_pass] = a2.n_z) */
} ptt_sweep_plan_t;

/* Plan builder.  mode 0: the SSMSpectrum.compute chain (pass 0 on y -> pow_v and a2, z_lookup = gen_lookup(y, cs,
 * n_z_lookup, eps=1e-8), pass 1 on z_lookup, GW integral on (z_lookup, y)); mode 1: power_gw_scaled_bag
 * (spectrum.py:243-294: one pass on z_lookup).  ptt_ssm_plan_create_host builds the host arrays only (no CUDA call:
 * usable without a device, e.g. to inspect the grids with ptt_ssm_plan_array); ptt_ssm_plan_upload puts them on the
 * current device and builds the cell tables of the GW integral; ptt_ssm_plan_create does both. */
#define PTT_PLAN_ONLY_FIRST_PASS 1   /* spec_den_v on y only (no z_lookup pass, no GW tables) */
#define PTT_PLAN_ONLY_LOOKUP_PASS 2  /* skip the y pass (no pow_v) */
#define PTT_PLAN_FORCE_DIRECT 4      /* exact sine transform with one sine per (z, xi) pair (validation) */
#define PTT_PLAN_NO_GW_CELLS 8       /* GW integral by the per-sample kernel */
typedef struct ptt_ssm_plan ptt_ssm_plan_t;
typedef struct ptt_ssm_plan_desc {
    int32_t n_y, nt, n_z_lookup, nxi;   /* defaults of the reference: 1000, 10000, 10000, 2000 (ssmtools/const.py:10-20) */
    int32_t nuc_type;                   /* 0 exponential, 1 simultaneous (spectrum.py:189-205) */
    int32_t mode;
    int32_t phase_rule;                 /* -1: the mode's default (ptt_a2_plan_t.phase_rule) */
    int32_t nz_int;                     /* 0: n_z_lookup */
    int32_t flags, reserved;
    double cs;                          /* sound speed of the GW integral */
    double z_st_thresh;                 /* 0 -> 50 */
    double nuc_a;                       /* nucleation parameter a (1) */
    double gamma;                       /* 0 -> 4/3 (ssmtools/const.py:46) */
} ptt_ssm_plan_desc_t;
int ptt_ssm_plan_create_host(const ptt_ssm_plan_desc_t* desc, const double* y, ptt_ssm_plan_t** out);
int ptt_ssm_plan_upload(ptt_ssm_plan_t* plan);
int ptt_ssm_plan_create(const ptt_ssm_plan_desc_t* desc, const double* y, ptt_ssm_plan_t** out);
int ptt_ssm_plan_destroy(ptt_ssm_plan_t* plan);
/* the POD view ptt_sweep_spectra and the seam-C operators take (NULL before the upload) */
const ptt_sweep_plan_t* ptt_ssm_plan_view(const ptt_ssm_plan_t* plan);
/* host array by name ("y", "z_lookup", "z_all", "frac_all", "xi_re", "p<i>_{qT,z,zterm,T,LT,W,omega_ab}",
 * "gw_{L1,L2,Z1,Z2,G,yterm}"): copies up to cap doubles into out (may be NULL) and returns the size, -1 if unknown */
int64_t ptt_ssm_plan_array(const ptt_ssm_plan_t* plan, const char* name, double* out, int64_t cap);
/* n_pass, seg[n_pass+1], z_tile[n_pass], win_entries[n_pass], scal[8] = {l0_0, dl_0, l0_1, dl_1, z_exact_max,
 * gw prefactor, gw lrange, gw win_entries}; any pointer may be NULL */
int ptt_ssm_plan_info(const ptt_ssm_plan_t* plan, int32_t* n_pass, int32_t* seg, int32_t* z_tile, int32_t* win, double* scal);

#define PTT_SWEEP_NSCAL 32
/* columns of the per-point result table scalars[n, PTT_SWEEP_NSCAL]:
 *  0..13  vp, vm, vp_tilde, vm_tilde, v_sh, vm_sh, vm_tilde_sh, wp, wm, wm_sh, v_cj, wn, wn_estimate, solver_failed
 *         (the 17-tuple of sound_shell_generic, fluid.py:1052)
 *  14..29 va_kinetic_energy_density, va_trace_anomaly_diff, va_thermal_energy_density_diff, va_entropy_density_diff,
 *         wbar, ebar, kinetic_energy_density, kinetic_energy_fraction, va_thermal_energy_density,
 *         thermal_energy_density, va_enthalpy_density, kappa, omega, ubarf2, mean_adiabatic_index, nu_gdh2024
 *         (cached properties of Bubble, bubble.py:377-632 / thermo.py)
 *  30     source_lifetime_factor (spectrum.py:123-136)      31  number of profile samples */

#define PTT_SWEEP_TIMING 1     /* opts.flags: bracket every stage with CUDA events (ptt_sweep_last_stats) */
#define PTT_SWEEP_NO_RETRY 2   /* opts.flags: report points whose first root search fails as solver failures */

typedef struct ptt_sweep_opts {
    int32_t n_xi;              /* 0 -> 5000 (bubble/const.py:12) */
    int32_t thin_shell_limit;  /* 0 -> 100  (const.py:26) */
    int32_t chunk;             /* points per spectrum chunk, 0 -> 8000 */
    int32_t shell_chunk;       /* points per shell solve, 0 -> from free device memory */
    int32_t flags;
    int32_t group_min;         /* smallest run of equal v_wall that goes through the banded matrix product K6'
                                * (ptt_spec_den_v_group), 0 -> 24; shorter runs use the gather kernel */
    double t_end;              /* 0 -> 50 (const.py:18) */
    double wn_rtol;            /* 0 -> 1e-4 (fluid.py:650) */
    double r_star;             /* <= 0: Spectrum without r_star (lifetime factor 1/(1+2 nu) only) */
    double lifetime_multiplier;/* spectrum.py:123-136; may be +inf */
    /* Optional, PTT_MEM_DEVICE only: HOST copies of v_wall (pg column 0) and of the solution type (pg column 2), [n] each,
     * in the order of pg.  The driver groups the points by wall speed on the host; with these it does not have to fetch pg
     * from the device and wait for `stream` to get there, so a sweep can be queued while the previous one is still running
     * on the same stream (its shell solve then starts the moment the previous sweep ends).  Both or neither. */
    const double* host_v_wall;
    const int32_t* host_sol_type;
} ptt_sweep_opts_t;

typedef struct ptt_sweep_out {
    double* pow_v;             /* [n, n_y] or NULL */
    double* pow_gw;            /* [n, n_y] or NULL */
    double* omgw0;             /* [n, n_y] or NULL (needs scale[n]) */
    double* scalars;           /* [n, PTT_SWEEP_NSCAL] or NULL */
    int32_t* status;           /* [n] PTT_ST_* */
    double *rows_v, *rows_w, *rows_xi;   /* optional profiles [n, row_stride] (device memory only), with off/npts */
    int64_t row_stride;
    int32_t *off, *npts;
    const int64_t* table_rows; /* HOST [n] or NULL (PTT_MEM_DEVICE only): row of point i in pow_v / pow_gw / omgw0, for tables
                                * larger than this call's n -- e.g. the shard of a sweep writing straight into the full
                                * table, which may live on ANOTHER GPU of the box (peer memory opened with ptt_ipc_open:
                                * the rows then travel over NVLink as they are produced, no collective) */
    double* snr;               /* [n, n_snr] or NULL: signal-to-noise ratios of omgw0 (needs omgw0's scale and snr_weight) */
    const double* snr_weight;  /* DEVICE [n_snr, n_y], see ptt_snr_batch */
    int32_t n_snr, reserved;
} ptt_sweep_out_t;

/* Called after the rows [row0, row1) of the result tables have been queued for delivery on `stream` (a cudaStream_t):
 * whatever is ordered after that stream sees them complete.  Lets a caller overlap its own exchange step (e.g. the
 * NCCL gather of a sharded sweep) with the chunks still being computed.  Only issued per chunk when the points come
 * ordered by v_wall (otherwise once, for [0, n)). */
typedef void (*ptt_chunk_cb)(void* user, int64_t row0, int64_t row1, void* stream);

/* pg[n,16]: as ptt_sweep_shells_generic.  pm[n,4] = {a_s, a_b, T_ref, 0} of each point's model.
 * scale[n] = r_star * F_gw0 * suppression(v_wall, alpha_n) (NULL: no omgw0).  Rows may come in any order; results are
 * returned in the caller's order.  With PTT_MEM_HOST the call returns when all results are in the caller's buffers. */
int ptt_sweep_spectra(const ptt_sweep_plan_t* plan, const ptt_sweep_opts_t* opts, int64_t n, const double* pg,
                      const double* pm, const double* scale, const ptt_sweep_out_t* out, ptt_chunk_cb cb, void* user,
                      int mem, void* stream);

#define PTT_STAGE_SHELLS 0
#define PTT_STAGE_RETRY 1
#define PTT_STAGE_THERMO 2
#define PTT_STAGE_A2 3
#define PTT_STAGE_SDV_Y 4
#define PTT_STAGE_SDV 5
#define PTT_STAGE_GW 6
#define PTT_STAGE_EMIT 7
#define PTT_STAGE_RETRY_FOLD 8
#define PTT_N_STAGES 12
typedef struct ptt_sweep_stats {
    double stage_ms[PTT_N_STAGES];     /* device time per stage (events on the launching stream; stages on different
                                        * streams overlap, so the sum can exceed the duration of the call) */
    int64_t stage_calls[PTT_N_STAGES];
    int64_t launches;                  /* kernels launched */
    int64_t retried, rescued;          /* points sent to / solved by the retry search */
    int64_t group_profiles;            /* profiles multiplied in K6' (lookup pass), after trimming no-solution rows */
    int64_t group_count;
    int64_t group_kw_sum;              /* sum over those groups of the band width in knots */
    int64_t group_macs;                /* multiply-adds of those matrix products: sum of n_z * band * profiles */
    int64_t a2_profiles, gw_profiles;  /* profiles that went through A^2 / the GW integral */
    int64_t chunks_before_fold, chunks_after_fold;   /* spectrum chunks queued before / after the retries were folded in */
    double host_wait_retry_ms;         /* host time spent waiting for the retry search (the device may idle meanwhile) */
    double host_wait_shells_ms;        /* host time spent waiting for the status of a shell chunk */
    int64_t timing_serial;             /* number of this sweep in the process (1, 2, ...), see ptt_sweep_timing_collect */
} ptt_sweep_stats_t;
int ptt_sweep_last_stats(ptt_sweep_stats_t* out);
/* With PTT_MEM_DEVICE a timed sweep (PTT_SWEEP_TIMING) returns with its last chunks in flight and stage_ms not yet read
 * from the events.  ptt_sweep_timing_sync waits for every sweep queued so far and completes the stats
 * ptt_sweep_last_stats returns.  ptt_sweep_timing_collect waits for ONE sweep (its timing_serial) and returns its completed
 * stats: a caller that queues sweeps back to back reads sweep k's times while sweep k+1 runs, without waiting for k+1
 * (the stage events alternate between two sets; the last eight completed sweeps are kept). */
int ptt_sweep_timing_sync(void);
int ptt_sweep_timing_collect(int64_t timing_serial, ptt_sweep_stats_t* out);
/* frees the device / pinned workspace the sweep driver keeps between calls */
int ptt_sweep_release(void);
/* Device buffers that other processes of the box can map (CUDA IPC): the result table of a sharded sweep lives on one
 * GPU, every rank's sweep writes its rows into it through ptt_sweep_out_t.table_rows.  handle: 64 bytes to hand to the
 * other processes (any transport); ptt_ipc_open maps it there (peer access over NVLink / NVSwitch is enabled lazily). */
int ptt_ipc_alloc(int64_t bytes, void** ptr, unsigned char* handle64);
int ptt_ipc_open(const unsigned char* handle64, void** ptr);
int ptt_ipc_close(void* ptr);
int ptt_ipc_free(void* ptr);
/* points per shell solve the driver would pick for `free_bytes` of device memory */
int64_t ptt_sweep_shell_chunk(int n_xi, int64_t free_bytes);

/* --- Post-processing epilogues (device pointers, enqueued on `stream`) ---------------------------------------
 * Suppression factor of Omega_gw,0: piecewise-linear interpolation of the simulation-measured table on its Delaunay
 * triangulation, NaN outside the hull (Suppression.suppression -> scipy griddata(method="linear"),
 * pttools/omgw0/suppression/suppression.py:53-96).  tri[n_tri,6] = (x0,y0,x1,y1,x2,y2) of each triangle in
 * (v_wall, alpha_n), val[n_tri,3] the table values at its vertices (the triangulation is built once on the host).
 * mode 0 = SuppressionMethod.NO_EXT, 1 = EXT_CONSTANT (1 below alpha_min). */
int ptt_suppression_batch(int n, const double* v_wall, const double* alpha_n, int n_tri, const double* tri,
                          const double* val, int mode, double alpha_min, double* out, void* stream);
/* signal_to_noise_ratio(f, signal, noise) = sqrt(T_obs * trapezoid(signal^2 / noise^2, f)) for P spectra that share f
 * (pttools/omgw0/noise.py:7-52; Spectrum.signal_to_noise_ratio[_instrument], omgw0_ssm.py:143-147), for n_w <= 4 noise
 * curves at once: weight[n_w, n_y] = T_obs * trapezoid weight(f) / noise^2, out[P, n_w]. */
int ptt_snr_batch(int P, int n_y, const double* signal, int64_t stride, int n_w, const double* weight, double* out,
                  void* stream);

/* FP64 FMA throughput probe used by bench.py for the roofline denominator (not part of the reference). */
int ptt_fp64_peak_probe(int iters, double* out_flops_per_s, void* stream);
/* mode 0: DFMA only, 1: DMMA m8n8k4 only, 2: both interleaved - measures what the FP64 tensor pipe can deliver on this
 * part next to the FMA pipe (roofline documentation). */
int ptt_fp64_pipe_probe(int mode, int iters, double* out_flops_per_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif
