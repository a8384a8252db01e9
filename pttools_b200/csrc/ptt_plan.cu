// Spectrum plans: the grids of the Sound Shell Model chain that depend on (y, cs, sizes, nucleation law) only, built on
// the host with the reference's expressions and uploaded once per sweep.
//   qT_lookup = 10**arange(log10 zmin + log10 0.01, log10 zmax + log10 20, dlog10z)   pttools/ssmtools/spectrum.py:503-515
//   t = logspace(log10 0.01, log10 20, nt), nu(t)                                     spectrum.py:473, 189-205
//   z_lookup = gen_lookup(y, cs, n_z_lookup, eps=1e-8), lookup_limits                 spectrum.py:106, 169-186
//   log-spaced integration variable of spec_den_gw_scaled                             spectrum.py:347-357
//   blend region of sin_transform, xi_re of resample_uniform_xi                       calculators.py:100, 136-153
// Host code only (plus the uploads); mirrors pttools_b200/engine.py's numpy plan, against which tests/ pin it.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "ptt_host.cuh"

namespace ptt {

constexpr double T_TILDE_MIN = 0.01, T_TILDE_MAX = 20.0;      // ssmtools/const.py:35-37
constexpr double DZ_ST_BLEND = M_PI;                          // ssmtools/const.py:44
constexpr size_t SMEM_WINDOW_BYTES = 200 * 1024;              // shared-memory budget of the lookup windows (K6, K7)

typedef std::vector<double> vec;

// np.linspace(a, b, n)
static vec linspace(const double a, const double b, const int n) {
    vec y((size_t)n);
    if (n == 1) { y[0] = a; return y; }
    const double step = (b - a) / (double)(n - 1);
    for (int i = 0; i < n; ++i) y[i] = (double)i * step + a;
    y[n - 1] = b;
    return y;
}

// np.arange(start, stop, delta) for doubles: length ceil((stop - start) / delta), element i = start + i * ((start + delta) - start)
static vec arange(const double start, const double stop, const double delta) {
    const long long len = (long long)ceil((stop - start) / delta);
    vec y((size_t)std::max(len, 0LL));
    const double next = start + delta;
    const double d = next - start;
    for (long long i = 0; i < len; ++i) y[i] = start + (double)i * d;
    if (len > 1) y[1] = next;
    return y;
}

static vec pow10(const vec& x) {
    vec y(x.size());
    for (size_t i = 0; i < x.size(); ++i) y[i] = pow(10.0, x[i]);
    return y;
}

static vec trapz_weights(const vec& x) {
    const size_t n = x.size();
    vec w(n);
    for (size_t i = 1; i + 1 < n; ++i) w[i] = 0.5 * (x[i + 1] - x[i - 1]);
    w[0] = 0.5 * (x[1] - x[0]);
    w[n - 1] = 0.5 * (x[n - 1] - x[n - 2]);
    return w;
}

// Weight of the asymptotic formula in sin_transform per wavenumber (calculators.py:119-158): 0 exact only, 1 asymptotic
// only, cubic blend on (thresh - pi, thresh].
static vec blend_fraction(const vec& z, const double thresh) {
    vec frac(z.size(), 0.0);
    for (size_t i = 0; i < z.size(); ++i)
        if (z[i] > thresh) frac[i] = 1.0;
    if (z.size() <= 1) return frac;
    double zmax = -INFINITY, zmin = INFINITY;
    bool any = false;
    for (const double v : z)
        if (v > thresh - DZ_ST_BLEND && v <= thresh) { zmax = std::max(zmax, v); zmin = std::min(zmin, v); any = true; }
    if (!any) return frac;
    for (size_t i = 0; i < z.size(); ++i)
        if (z[i] > thresh - DZ_ST_BLEND && z[i] <= thresh) {
            const double s = zmax > zmin ? (z[i] - zmin) / (zmax - zmin) : 0.5;
            frac[i] = 3 * s * s - 2 * s * s * s;
        }
    return frac;
}

struct HostPass {
    vec z, qT, T, LT, W, zterm, omega_ab;
    double l0 = 0, dl = 0;
    int nt = 0, z_tile = 0, win_entries = 0;
};

// (alpha, beta) of the knot weight of K6' on every piece between consecutive kinks of {LT_j - 1, LT_j, LT_j + 1}
// (engine._omega_pieces).
static vec omega_pieces(const vec& LT, const vec& T, const vec& W, const double dl) {
    const int nt = (int)LT.size();
    const double D = pow(10.0, dl) - 1.0;
    const int nk = 3 * nt;
    vec kinks((size_t)nk);
    std::vector<int> label((size_t)nk), order((size_t)nk);
    for (int j = 0; j < nt; ++j) {
        kinks[j] = LT[j] - 1.0; label[j] = 0;
        kinks[nt + j] = LT[j]; label[nt + j] = 1;
        kinks[2 * nt + j] = LT[j] + 1.0; label[2 * nt + j] = 2;
    }
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return kinks[a] < kinks[b]; });
    std::vector<long long> cnt[3];
    for (int k = 0; k < 3; ++k) cnt[k].assign((size_t)nk + 1, 0);
    for (int p = 0; p < nk; ++p)
        for (int k = 0; k < 3; ++k) cnt[k][p + 1] = cnt[k][p] + (label[order[p]] == k ? 1 : 0);
    vec out(2 * ((size_t)nk + 1));
    const double p10 = pow(10.0, dl);
    for (int p = 0; p <= nk; ++p) {
        // A = [n_0, n_m1): weight 1 - t;  B = [n_p1, n_0): weight t
        double sA = 0, sAT = 0, sB = 0, sBT = 0;
        for (long long j = cnt[1][p]; j < cnt[0][p]; ++j) { const long long jj = std::min<long long>(j, nt - 1); sA += W[jj]; sAT += W[jj] * T[jj]; }
        for (long long j = cnt[2][p]; j < cnt[1][p]; ++j) { const long long jj = std::min<long long>(j, nt - 1); sB += W[jj]; sBT += W[jj] * T[jj]; }
        out[2 * p] = (1.0 + 1.0 / D) * sA - sB / D;
        out[2 * p + 1] = -sAT / D + (p10 / D) * sBT;
    }
    return out;
}

static int make_pass(const vec& z, const int nt, const int nuc_type, const double a, HostPass& p) {
    for (size_t i = 1; i < z.size(); ++i)
        if (z[i] < z[i - 1]) return set_error(PTT_E_ARG, "ptt_ssm_plan_create: z must be ascending");
    p.z = z;
    p.nt = nt;
    const double zmin = *std::min_element(z.begin(), z.end()), zmax = *std::max_element(z.begin(), z.end());
    const double lmin = log10(zmin), lmax = log10(zmax);
    p.dl = (lmax - lmin) / (double)z.size();                       // spectrum.py:507 (size, not size - 1)
    p.l0 = lmin + log10(T_TILDE_MIN);
    p.qT = pow10(arange(p.l0, lmax + log10(T_TILDE_MAX), p.dl));
    const vec lt = linspace(log10(T_TILDE_MIN), log10(T_TILDE_MAX), nt);
    p.T = pow10(lt);
    p.LT.resize(lt.size());
    for (size_t j = 0; j < lt.size(); ++j) p.LT[j] = lt[j] / p.dl;
    const vec tw = trapz_weights(p.T);
    p.W.resize(lt.size());
    for (size_t j = 0; j < lt.size(); ++j) {
        const double t = p.T[j];
        double nu;
        if (nuc_type == 1) { const double at = a * t; nu = 0.5 * a * (at * at) * exp(-(at * at * at) / 6.0); }   // simultaneous
        else nu = a * exp(-a * t);                                                                            // exponential
        const double t2 = t * t;
        p.W[j] = tw[j] * (t2 * t2 * t2) * nu;
    }
    p.zterm.resize(z.size());
    for (size_t i = 0; i < z.size(); ++i) p.zterm[i] = (log10(z[i]) - p.l0) / p.dl;
    // tile size of the gather kernel K6: the lookup window of a tile must fit in shared memory
    const long long cap_entries = SMEM_WINDOW_BYTES / 16;
    const double t_span = p.LT.back() - p.LT.front();
    const long long nz = (long long)z.size();
    long long tile = 512, best_tile = 0, best_need = 0;
    for (;;) {
        double span = -INFINITY;
        for (long long i = 0; i < nz; i += tile) span = std::max(span, p.zterm[std::min(i + tile - 1, nz - 1)] - p.zterm[i]);
        const long long need = (long long)ceil(span + t_span) + 6;
        if (need <= cap_entries) {
            best_tile = tile; best_need = need;
            if (tile >= nz) break;
            tile += 512;
        } else break;
    }
    if (!best_tile) return set_error(PTT_E_ARG, "ptt_ssm_plan_create: the spec_den_v lookup window does not fit in shared memory");
    p.z_tile = (int)best_tile;
    p.win_entries = (int)best_need;
    std::vector<double> WT;
    p.omega_ab = omega_pieces(p.LT, p.T, p.W, p.dl);
    return PTT_OK;
}

struct HostGw {
    vec L1, L2, Z1, Z2, G, yterm;
    int n_int = 0, y_tile = 32, win_entries = 0;
    double prefactor = 0, lrange = 0;
};

// Tables of the GW integral (spectrum.py:327-368) with the substitution z = y * zeta; z_lookup is log-uniform.
static int make_gw(const vec& z_lookup, const vec& y, const double cs, const double gamma, const int nz_int, HostGw& g) {
    const int n_zl = (int)z_lookup.size();
    g.n_int = nz_int > 0 ? nz_int : n_zl;
    const double zpf = (1 + cs) / (2 * cs), zmf = (1 - cs) / (2 * cs);
    const vec lz = linspace(log10(zmf), log10(zpf), g.n_int);
    g.Z1 = pow10(lz);
    const double lz0 = log10(z_lookup[0]);
    const double dlz = (log10(z_lookup[n_zl - 1]) - lz0) / (double)(n_zl - 1);
    const size_t m = lz.size();
    g.Z2.resize(m); g.L1.resize(m); g.L2.resize(m); g.G.resize(m);
    const vec tw = trapz_weights(g.Z1);
    double lmax = -INFINITY, lmin = INFINITY;
    for (size_t k = 0; k < m; ++k) {
        const double zeta = g.Z1[k];
        const double zeta2 = zpf + zmf - zeta;
        g.Z2[k] = zeta2;
        g.L1[k] = lz[k] / dlz;
        g.L2[k] = log10(zeta2) / dlz;
        const double a = zeta - zpf, b = zeta - zmf;
        const double kern = ((a * a) * (b * b)) / (zeta * zeta2);
        g.G[k] = tw[k] * kern;
        lmax = std::max(lmax, std::max(g.L1[k], g.L2[k]));
        lmin = std::min(lmin, std::min(g.L1[k], g.L2[k]));
    }
    g.yterm.resize(y.size());
    for (size_t i = 0; i < y.size(); ++i) g.yterm[i] = (log10(y[i]) - lz0) / dlz;
    const double c2 = cs * cs;
    const double r = (1 - c2) / c2;
    g.prefactor = 0.75 * (gamma * gamma) * (r * r) / (4 * M_PI * cs);
    g.lrange = lmax - lmin;
    double span = -INFINITY;
    const long long ny = (long long)y.size();
    for (long long i = 0; i < ny; i += g.y_tile) span = std::max(span, g.yterm[std::min(i + g.y_tile - 1, ny - 1)] - g.yterm[i]);
    g.win_entries = (int)ceil(fabs(span) + g.lrange) + 6;
    if ((size_t)g.win_entries * 16 > SMEM_WINDOW_BYTES)
        return set_error(PTT_E_ARG, "ptt_ssm_plan_create: the spec_den_gw lookup window does not fit in shared memory");
    return PTT_OK;
}

}  // namespace ptt

using namespace ptt;

struct ptt_ssm_plan {
    ptt_ssm_plan_desc_t desc;
    vec y, z_lookup, z_all, frac_all, xi_re;
    std::vector<int32_t> seg;
    std::vector<HostPass> passes;
    HostGw gw;
    bool has_gw = false, has_y_pass = false, uploaded = false;
    int phase_rule = 1;
    double z_exact_max = 0;
    std::map<std::string, const vec*> arrays;
    std::vector<void*> dev;
    ptt_sweep_plan_t view;
};

static void register_arrays(ptt_ssm_plan* p) {
    p->arrays.clear();
    p->arrays["y"] = &p->y; p->arrays["z_lookup"] = &p->z_lookup; p->arrays["z_all"] = &p->z_all;
    p->arrays["frac_all"] = &p->frac_all; p->arrays["xi_re"] = &p->xi_re;
    for (size_t i = 0; i < p->passes.size(); ++i) {
        const std::string pre = "p" + std::to_string(i) + "_";
        HostPass& q = p->passes[i];
        p->arrays[pre + "qT"] = &q.qT; p->arrays[pre + "z"] = &q.z; p->arrays[pre + "zterm"] = &q.zterm;
        p->arrays[pre + "T"] = &q.T; p->arrays[pre + "LT"] = &q.LT; p->arrays[pre + "W"] = &q.W;
        p->arrays[pre + "omega_ab"] = &q.omega_ab;
    }
    if (p->has_gw) {
        p->arrays["gw_L1"] = &p->gw.L1; p->arrays["gw_L2"] = &p->gw.L2; p->arrays["gw_Z1"] = &p->gw.Z1;
        p->arrays["gw_Z2"] = &p->gw.Z2; p->arrays["gw_G"] = &p->gw.G; p->arrays["gw_yterm"] = &p->gw.yterm;
    }
}

extern "C" int ptt_ssm_plan_create_host(const ptt_ssm_plan_desc_t* d, const double* y, ptt_ssm_plan_t** out) {
    if (!d || !y || !out || d->n_y < 1 || d->nt < 2 || d->n_z_lookup < 2 || d->nxi < 2 || !(d->cs > 0 && d->cs < 1))
        return set_error(PTT_E_ARG, "ptt_ssm_plan_create: null pointer or bad size");
    if (d->mode != 0 && d->mode != 1) return set_error(PTT_E_ARG, "ptt_ssm_plan_create: mode must be 0 (ssm) or 1 (bag)");
    if (d->nuc_type != 0 && d->nuc_type != 1) return set_error(PTT_E_ARG, "Nucleation type not recognized");
    for (int i = 0; i < d->n_y; ++i)
        if (!(y[i] > 0.0)) return set_error(PTT_E_ARG, "z values must all be positive.");
    ptt_ssm_plan* p = new ptt_ssm_plan();
    p->desc = *d;
    p->y.assign(y, y + d->n_y);
    const double cs = d->cs;
    const double eps = 1e-8;
    const double ymin = *std::min_element(p->y.begin(), p->y.end()), ymax = *std::max_element(p->y.begin(), p->y.end());
    const double lo = ymin * 0.5 * (1.0 - cs) / cs - eps;        // lookup_limits, spectrum.py:183-186
    const double hi = ymax * 0.5 * (1.0 + cs) / cs + eps;
    p->z_lookup = pow10(linspace(log10(lo), log10(hi), d->n_z_lookup));
    const bool only_first = d->flags & PTT_PLAN_ONLY_FIRST_PASS, only_lookup = d->flags & PTT_PLAN_ONLY_LOOKUP_PASS;
    p->has_y_pass = d->mode == 0 && !only_lookup;
    p->phase_rule = d->phase_rule >= 0 ? d->phase_rule : (d->mode == 0 ? 1 : 0);
    int rc = PTT_OK;
    if (p->has_y_pass) {
        p->passes.emplace_back();
        rc = make_pass(p->y, d->nt, d->nuc_type, d->nuc_a, p->passes.back());
    }
    if (rc == PTT_OK && !only_first) {
        p->passes.emplace_back();
        rc = make_pass(p->z_lookup, d->nt, d->nuc_type, d->nuc_a, p->passes.back());
    }
    if (rc == PTT_OK && p->passes.empty()) rc = set_error(PTT_E_ARG, "ptt_ssm_plan_create: no pass left");
    if (rc == PTT_OK && !only_first) {
        rc = make_gw(p->z_lookup, p->y, cs, d->gamma > 0 ? d->gamma : 4.0 / 3.0, d->nz_int, p->gw);
        p->has_gw = rc == PTT_OK;
    }
    if (rc != PTT_OK) { delete p; return rc; }
    p->seg.assign(1, 0);
    const double thresh = d->z_st_thresh > 0 ? d->z_st_thresh : 50.0;
    for (const HostPass& q : p->passes) {
        p->seg.push_back(p->seg.back() + (int32_t)q.qT.size());
        p->z_all.insert(p->z_all.end(), q.qT.begin(), q.qT.end());
        const vec f = blend_fraction(q.qT, thresh);
        p->frac_all.insert(p->frac_all.end(), f.begin(), f.end());
    }
    p->xi_re = linspace(0.0, 1.0 - 1.0 / d->nxi, d->nxi);          // calculators.py:100
    p->z_exact_max = 0.0;
    for (size_t i = 0; i < p->z_all.size(); ++i)
        if (p->frac_all[i] < 1.0) p->z_exact_max = std::max(p->z_exact_max, p->z_all[i]);
    memset(&p->view, 0, sizeof(p->view));
    register_arrays(p);
    *out = p;
    return PTT_OK;
}

extern "C" int64_t ptt_ssm_plan_array(const ptt_ssm_plan_t* p, const char* name, double* out, int64_t cap) {
    if (!p || !name) return -1;
    const auto it = p->arrays.find(name);
    if (it == p->arrays.end()) return -1;
    const vec& v = *it->second;
    if (out) memcpy(out, v.data(), std::min<size_t>(v.size(), (size_t)std::max<int64_t>(cap, 0)) * sizeof(double));
    return (int64_t)v.size();
}

extern "C" int ptt_ssm_plan_info(const ptt_ssm_plan_t* p, int32_t* n_pass, int32_t* seg, int32_t* z_tile, int32_t* win,
                                 double* scal) {
    if (!p) return set_error(PTT_E_ARG, "ptt_ssm_plan_info: null plan");
    if (n_pass) *n_pass = (int32_t)p->passes.size();
    for (size_t i = 0; i < p->seg.size() && seg; ++i) seg[i] = p->seg[i];
    for (size_t i = 0; i < p->passes.size(); ++i) {
        if (z_tile) z_tile[i] = p->passes[i].z_tile;
        if (win) win[i] = p->passes[i].win_entries;
        if (scal) { scal[2 * i] = p->passes[i].l0; scal[2 * i + 1] = p->passes[i].dl; }
    }
    if (scal) { scal[4] = p->z_exact_max; scal[5] = p->has_gw ? p->gw.prefactor : 0.0; scal[6] = p->has_gw ? p->gw.lrange : 0.0;
                scal[7] = p->has_gw ? (double)p->gw.win_entries : 0.0; }
    return PTT_OK;
}

template <class T>
static const T* upload(ptt_ssm_plan* p, const std::vector<T>& v, cudaError_t& err) {
    void* d = nullptr;
    if (err != cudaSuccess) return nullptr;
    err = cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T));
    if (err != cudaSuccess) return nullptr;
    p->dev.push_back(d);
    err = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return (const T*)d;
}

extern "C" int ptt_ssm_plan_upload(ptt_ssm_plan_t* p) {
    if (!p) return set_error(PTT_E_ARG, "ptt_ssm_plan_upload: null plan");
    if (p->uploaded) return PTT_OK;
    cudaError_t err = cudaSuccess;
    ptt_sweep_plan_t& v = p->view;
    memset(&v, 0, sizeof(v));
    v.a2.n_z = (int32_t)p->z_all.size();
    v.a2.n_seg = (int32_t)p->passes.size();
    v.a2.seg = upload(p, p->seg, err);
    v.a2.z = upload(p, p->z_all, err);
    v.a2.frac = upload(p, p->frac_all, err);
    v.a2.nxi = p->desc.nxi;
    v.a2.phase_rule = p->phase_rule;
    v.a2.xi_re = upload(p, p->xi_re, err);
    v.a2.z_exact_max = p->z_exact_max;
    v.a2.force_direct = (p->desc.flags & PTT_PLAN_FORCE_DIRECT) ? 1 : 0;
    v.n_pass = (int32_t)p->passes.size();
    v.has_y_pass = p->has_y_pass ? 1 : 0;
    for (size_t i = 0; i < p->passes.size(); ++i) {
        HostPass& q = p->passes[i];
        ptt_sdv_plan_t& s = v.sdv[i];
        s.n_q = (int32_t)q.qT.size(); s.n_z = (int32_t)q.z.size(); s.nt = q.nt; s.z_tile = q.z_tile;
        s.qT = upload(p, q.qT, err); s.z = upload(p, q.z, err); s.zterm = upload(p, q.zterm, err);
        s.T = upload(p, q.T, err); s.LT = upload(p, q.LT, err); s.W = upload(p, q.W, err);
        s.inv_dlog = 1.0 / q.dl;
        s.smem_bytes = q.win_entries * 16;
        s.omega_ab = upload(p, q.omega_ab, err);
        v.h_zterm[i] = q.zterm.data();
        v.lt_first[i] = q.LT.front(); v.lt_last[i] = q.LT.back(); v.dl[i] = q.dl;
    }
    for (size_t i = 0; i < p->seg.size() && i < 4; ++i) v.seg[i] = p->seg[i];
    const double* d_y = upload(p, p->y, err);
    const double* d_zl = upload(p, p->z_lookup, err);
    if (p->has_gw) {
        ptt_gw_plan_t& g = v.gw;
        g.n_zl = (int32_t)p->z_lookup.size(); g.n_y = (int32_t)p->y.size(); g.n_int = p->gw.n_int; g.y_tile = p->gw.y_tile;
        g.z_lookup = d_zl; g.y = d_y;
        g.yterm = upload(p, p->gw.yterm, err); g.L1 = upload(p, p->gw.L1, err); g.L2 = upload(p, p->gw.L2, err);
        g.Z1 = upload(p, p->gw.Z1, err); g.Z2 = upload(p, p->gw.Z2, err); g.G = upload(p, p->gw.G, err);
        g.prefactor = p->gw.prefactor;
        g.smem_bytes = p->gw.win_entries * 16;
    }
    if (err != cudaSuccess) return set_cuda_error("ptt_ssm_plan_upload", err);
    // cell tables of the GW integral (K7'), built on the device; a plan whose lookups do not move monotonically keeps
    // cells.cap = 0 and runs on the per-sample kernel
    if (p->has_gw && !(p->desc.flags & PTT_PLAN_NO_GW_CELLS)) {
        const int cap = ptt_gw_cells_capacity(v.gw.n_zl, v.gw.n_int, p->gw.lrange);
        if (cap > 0) {
            const int cap_c = ptt_gw_cells_coef_capacity(cap);
            const size_t ny = p->y.size();
            int32_t *ncell = nullptr, *meta = nullptr;
            double* coef = nullptr;
            if (cudaMalloc((void**)&ncell, ny * sizeof(int32_t)) == cudaSuccess) p->dev.push_back(ncell); else ncell = nullptr;
            if (ncell && cudaMalloc((void**)&meta, ny * ((size_t)cap + 2) * sizeof(int32_t)) == cudaSuccess) p->dev.push_back(meta); else meta = nullptr;
            if (meta && cudaMalloc((void**)&coef, ny * (size_t)cap_c * 2 * sizeof(double)) == cudaSuccess) p->dev.push_back(coef); else coef = nullptr;
            if (coef) {
                cudaMemset(meta, 0, ny * ((size_t)cap + 2) * sizeof(int32_t));
                cudaMemset(coef, 0, ny * (size_t)cap_c * 2 * sizeof(double));
                const int rc = ptt_gw_cells_build(&v.gw, cap, ncell, meta, coef, nullptr);
                if (rc != PTT_OK) return rc;
                std::vector<int32_t> h(ny);
                const cudaError_t e = cudaMemcpy(h.data(), ncell, ny * sizeof(int32_t), cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) return set_cuda_error("gw_cells_build_kernel", e);
                if (*std::min_element(h.begin(), h.end()) >= 0) {
                    v.cells.cap = cap; v.cells.ncell = ncell; v.cells.meta = meta; v.cells.coef = coef;
                }
            } else cudaGetLastError();
        }
    }
    p->uploaded = true;
    return PTT_OK;
}

extern "C" int ptt_ssm_plan_create(const ptt_ssm_plan_desc_t* d, const double* y, ptt_ssm_plan_t** out) {
    ptt_ssm_plan_t* p = nullptr;
    int rc = ptt_ssm_plan_create_host(d, y, &p);
    if (rc != PTT_OK) return rc;
    rc = ptt_ssm_plan_upload(p);
    if (rc != PTT_OK) { ptt_ssm_plan_destroy(p); return rc; }
    *out = p;
    return PTT_OK;
}

extern "C" const ptt_sweep_plan_t* ptt_ssm_plan_view(const ptt_ssm_plan_t* p) {
    if (!p || !p->uploaded) return nullptr;
    return &p->view;
}

extern "C" int ptt_ssm_plan_destroy(ptt_ssm_plan_t* p) {
    if (!p) return PTT_OK;
    for (void* d : p->dev) cudaFree(d);
    delete p;
    return PTT_OK;
}
