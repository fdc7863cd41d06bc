// Batched form of the bounded TRF Moffat fit (same iteration as trf.cuh, i.e. scipy's trf_bounds
// as called from /root/reference/src/motes/tracing.py:550-607), cut into phases so that a batch of
// fits can be advanced in lock step by a persistent grid:
//
//   fit_start     start point and bounds from the profile (tracing.py:578-588)      one warp per fit
//   fit_evaluate  residual sweep at the trial point, the accept / reject / stop logic of the inner
//                 loop (trf.py:335-383), and for an accepted point the forward-difference Jacobian,
//                 g = J^T f and the QR of [J_h ; sqrt(diag_h)] with f carried along   one CTA per fit
//   fit_propose   Coleman-Li scaling, stop tests of the outer loop, SVD of the 6 x 6 triangle,
//                 trust-region step and step selection (trf.py:280-333)              one THREAD per fit
//
// The six-parameter logic therefore runs once per fit instead of once per lane, and the
// data-parallel sweeps run with a whole CTA per fit.  Everything between phases lives in FitState.
#pragma once

#include "trf.cuh"

// Unroll factors of the sweeps' point loops (1 = not unrolled: smallest code; see qr_step).
#ifndef MOTES_FIT_UNROLL_RES
#define MOTES_FIT_UNROLL_RES 1
#endif
#ifndef MOTES_FIT_UNROLL_LIN
#define MOTES_FIT_UNROLL_LIN 1
#endif
#ifndef MOTES_FIT_UNROLL_QR
#define MOTES_FIT_UNROLL_QR 1
#endif

namespace motes {

constexpr int kFitUnrollRes = MOTES_FIT_UNROLL_RES, kFitUnrollLin = MOTES_FIT_UNROLL_LIN, kFitUnrollQr = MOTES_FIT_UNROLL_QR;
constexpr int kStatusRunning = -100;      // scipy's termination_status None

struct FitState {
    double x[kParams], lb[kParams], ub[kParams];
    double g[kParams];
    double R[kParams * kParams];          // upper triangle of the QR of the augmented scaled Jacobian
    double z[kParams];                    // Q^T f
    double x_new[kParams];
    double cost, radius, alpha, predicted, step_h_norm, step_norm, x_norm;
    int nfev, njev, status, first;
};

// Scratch of the sweep phases: a [7 m] = six Jacobian columns + the residual column.  The pow cache of the residual
// sweep (unit [m]) shares the storage of column 0: the thread that owns point i reads unit[i] before it writes the
// point's six column entries, and after the QR the columns are dead (R and Q^T f live in FitState).
MOTES_HD size_t fit_sweep_doubles(int m) { return (size_t)7 * m; }

// ---- start point and bounds -------------------------------------------------------------------
// Replicated serial scan of the profile (every lane of the calling scope computes the same values).
MOTES_HD void fit_start(int m, const double* profile, double scale, double alpha0, FitState* st) {
    using namespace trf_detail;
    constexpr int n = kParams;
    double top1 = -INFINITY, top2 = -INFINITY, top3 = -INFINITY;
    int peak = 0, n_nan = 0;
    bool peak_nan = false;
    for (int i = 0; i < m; ++i) {
        double val = profile[i] * scale;
        if (val != val) {
            if (!peak_nan) { peak_nan = true; peak = i; }       // np.argmax returns the first NaN
            ++n_nan;
            continue;
        }
        if (!peak_nan && val > top1) peak = i;
        if (val > top1) { top3 = top2; top2 = top1; top1 = val; }
        else if (val > top2) { top3 = top2; top2 = val; }
        else if (val > top3) { top3 = val; }
    }
    double amp0;
    if (m - n_nan >= 3 && n_nan == 0) amp0 = top2;
    else if (n_nan == 1 && m >= 3) amp0 = (top2 + top1) * 0.5;
    else if (n_nan == 2 && m >= 3) amp0 = top1;
    else amp0 = (m >= 3) ? NAN : top1;
    (void)top3;
    // np.median(concatenate(col[:5], col[-5:]))
    double ends[10];
    int n_ends = 0;
    for (int i = 0; i < 5 && i < m; ++i) ends[n_ends++] = profile[i] * scale;
    for (int i = (m > 5 ? m - 5 : 0); i < m; ++i) ends[n_ends++] = profile[i] * scale;
    bool ends_nan = false;
    for (int i = 1; i < n_ends; ++i) {
        double key = ends[i];
        int j = i - 1;
        while (j >= 0 && ends[j] > key) { ends[j + 1] = ends[j]; --j; }
        ends[j + 1] = key;
    }
    for (int i = 0; i < n_ends; ++i) ends_nan = ends_nan || (ends[i] != ends[i]);
    double level0 = (n_ends % 2) ? ends[n_ends / 2] : (ends[n_ends / 2 - 1] + ends[n_ends / 2]) * 0.5;
    if (ends_nan) level0 = NAN;

    const double x0[n] = {amp0, (double)peak, alpha0, kBetaStart, level0, 0.0};
    const double lb[n] = {0.0, (double)peak - 1.0, 0.0, 0.0, -INFINITY, -INFINITY};
    const double ub[n] = {INFINITY, (double)peak + 1.0, 5.0 * alpha0, 5.0, INFINITY, INFINITY};
    for (int i = 0; i < n; ++i) { st->x[i] = x0[i]; st->lb[i] = lb[i]; st->ub[i] = ub[i]; st->x_new[i] = x0[i]; st->g[i] = 0.0; }
    st->cost = NAN;
    st->radius = 0.0;
    st->alpha = 0.0;
    st->predicted = 0.0;
    st->step_h_norm = 0.0;
    st->step_norm = 0.0;
    st->x_norm = 0.0;
    st->nfev = 0;
    st->njev = 0;
    st->first = 1;
    st->status = kStatusRunning;
    for (int i = 0; i < n; ++i)
        if (!(lb[i] < ub[i])) { st->status = kFitBadBounds; return; }
    if (!inside(st->x, lb, ub)) { st->status = kFitX0Infeasible; return; }
    push_inside(st->x, lb, ub, 1e-10);
}

#if defined(__CUDACC__)
// The same start point with the profile scan spread over a warp (the serial scan above is 400 dependent global
// loads per fit): every lane scans its stride, the (largest, second largest) pairs, the first index of the
// maximum, the first NaN and the NaN count are merged by shuffles.  Identical results: a multiset's two largest
// values and np.argmax's "first occurrence" do not depend on the scan order.
__device__ __forceinline__ void fit_start_warp(int lane, int m, const double* __restrict__ profile, double scale,
                                               double alpha0, FitState* st) {
    using namespace trf_detail;
    constexpr int n = kParams;
    double top1 = -INFINITY, top2 = -INFINITY;
    int peak = 0x7fffffff, first_nan = 0x7fffffff, n_nan = 0;
    for (int i = lane; i < m; i += 32) {
        const double val = profile[i] * scale;
        if (val != val) { ++n_nan; first_nan = min(first_nan, i); continue; }
        if (val > top1) { top2 = top1; top1 = val; peak = i; }
        else if (val > top2) top2 = val;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double o1 = __shfl_xor_sync(0xffffffffu, top1, o), o2 = __shfl_xor_sync(0xffffffffu, top2, o);
        const int op = __shfl_xor_sync(0xffffffffu, peak, o);
        n_nan += __shfl_xor_sync(0xffffffffu, n_nan, o);
        first_nan = min(first_nan, __shfl_xor_sync(0xffffffffu, first_nan, o));
        const double hi = fmax(top1, o1), lo = fmin(top1, o1);
        // first occurrence of the maximum: the smaller index among equal maxima
        peak = (o1 > top1) ? op : (o1 == top1 ? min(peak, op) : peak);
        top2 = fmax(lo, fmax(top2, o2));
        top1 = hi;
    }
    if (n_nan > 0) peak = first_nan;                              // np.argmax returns the first NaN
    if (peak == 0x7fffffff) peak = 0;
    double amp0;
    if (m - n_nan >= 3 && n_nan == 0) amp0 = top2;
    else if (n_nan == 1 && m >= 3) amp0 = (top2 + top1) * 0.5;
    else if (n_nan == 2 && m >= 3) amp0 = top1;
    else amp0 = (m >= 3) ? NAN : top1;
    double ends[10];
    int n_ends = 0;
    for (int i = 0; i < 5 && i < m; ++i) ends[n_ends++] = profile[i] * scale;
    for (int i = (m > 5 ? m - 5 : 0); i < m; ++i) ends[n_ends++] = profile[i] * scale;
    bool ends_nan = false;
    for (int i = 1; i < n_ends; ++i) {
        double key = ends[i];
        int j = i - 1;
        while (j >= 0 && ends[j] > key) { ends[j + 1] = ends[j]; --j; }
        ends[j + 1] = key;
    }
    for (int i = 0; i < n_ends; ++i) ends_nan = ends_nan || (ends[i] != ends[i]);
    double level0 = (n_ends % 2) ? ends[n_ends / 2] : (ends[n_ends / 2 - 1] + ends[n_ends / 2]) * 0.5;
    if (ends_nan) level0 = NAN;
    if (lane != 0) return;
    const double x0[n] = {amp0, (double)peak, alpha0, kBetaStart, level0, 0.0};
    const double lb[n] = {0.0, (double)peak - 1.0, 0.0, 0.0, -INFINITY, -INFINITY};
    const double ub[n] = {INFINITY, (double)peak + 1.0, 5.0 * alpha0, 5.0, INFINITY, INFINITY};
    for (int i = 0; i < n; ++i) { st->x[i] = x0[i]; st->lb[i] = lb[i]; st->ub[i] = ub[i]; st->x_new[i] = x0[i]; st->g[i] = 0.0; }
    st->cost = NAN;
    st->radius = 0.0;
    st->alpha = 0.0;
    st->predicted = 0.0;
    st->step_h_norm = 0.0;
    st->step_norm = 0.0;
    st->x_norm = 0.0;
    st->nfev = 0;
    st->njev = 0;
    st->first = 1;
    st->status = kStatusRunning;
    for (int i = 0; i < n; ++i)
        if (!(lb[i] < ub[i])) { st->status = kFitBadBounds; return; }
    if (!inside(st->x, lb, ub)) { st->status = kFitX0Infeasible; return; }
    push_inside(st->x, lb, ub, 1e-10);
}
#endif

// ---- sweeps -----------------------------------------------------------------------------------
template <class Ctx>
MOTES_HD bool residual_sweep_b(const Ctx& ctx, int m, const double* profile, double scale, const double* xe,
                               double* unit_out, double* out, double* cost) {
    double acc = 0.0;
    int finite = 1;
#pragma unroll kFitUnrollRes
    for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
        double r = (double)i;
        double u = moffat_unit(r, xe[1], xe[2], xe[3]);
        double f = moffat_value(u, r, xe[0], xe[4], xe[5]) - profile[i] * scale;
        unit_out[i] = u;
        out[i] = f;
        acc += f * f;
        finite &= (int)is_finite(f);
    }
    *cost = 0.5 * ctx.sum(acc);
    return ctx.all(finite) != 0;
}

// ---- forward differences without a second pow --------------------------------------------------
// scipy's 2-point Jacobian evaluates the model at x + h e_k with h ~ 1.5e-8 |x_k| (_numdiff.py).  For
// the three parameters inside the pow that is pow(s', -beta') with s' = s (1 + delta), |delta| ~ 1e-6,
// or beta' = beta + h.  Both are the base value u = pow(s, -beta) times exp(small):
//     pow(s', -beta)     = u exp(-beta log1p(delta)),    delta = (q' - q) / s
//     pow(s, -(beta+h))  = u exp(-h log(s))
// evaluated with short series.  The result is within an ulp or two of the shifted pow - the same
// accuracy the reference's libm / SVML pow has - at a fraction of its cost.  Steps that are not small
// (a parameter sitting on a bound makes scipy take the whole feasible interval) use the real pow.
MOTES_HD double expm1_small(double e) {          // |e| < 1e-2: truncation < 2e-17 relative
    return e * (1.0 + e * (0.5 + e * (1.0 / 6.0 + e * (1.0 / 24.0 + e * (1.0 / 120.0 + e * (1.0 / 720.0))))));
}
// -DMOTES_FIT_REAL_POW: every forward difference evaluates the real pow (A/B build for the parity table,
// tools/parity_table.py: how much of the parameter tail is this shortcut and how much the reference's own jitter).
#ifdef MOTES_FIT_REAL_POW
constexpr bool kFitRealPow = true;
#else
constexpr bool kFitRealPow = false;
#endif
MOTES_HD bool shifted_unit(double u, double q_shift, double q, double inv_s, double beta, double* out) {
    if (kFitRealPow) return false;
    const double delta = (q_shift - q) * inv_s;
    if (!(fabs(delta) < 1e-3)) return false;
    const double l = delta * (1.0 - delta * (0.5 - delta * (1.0 / 3.0 - delta * (0.25 - delta * 0.2))));   // log1p
    const double e = -beta * l;
    if (!(fabs(e) < 1e-2)) return false;
    *out = fma(u, expm1_small(e), u);
    return true;
}
MOTES_HD bool shifted_unit_beta(double u, double s, double dbeta, double* out) {
    if (kFitRealPow) return false;
    const double e = -dbeta * log(s);
    if (!(fabs(e) < 1e-2)) return false;
    *out = fma(u, expm1_small(e), u);
    return true;
}
#if defined(__CUDACC__)
__host__ __device__ __noinline__ inline double moffat_unit_call(double r, double c, double alpha, double beta) {
    return moffat_unit(r, c, alpha, beta);
}
#else
inline double moffat_unit_call(double r, double c, double alpha, double beta) { return moffat_unit(r, c, alpha, beta); }
#endif

// Householder step K of the QR of [J_h | f] with the sparse pivot row sqrt(diag_h[K]) (see linearize).
// `part` holds this thread's share of the dot products of column K with columns K .. n (the step's only reduction);
// the pass that applies the reflector accumulates, from the values it has just written, the same dot products for
// column K + 1 - operand for operand what a separate pass over the updated columns would compute - so a step is one
// pass over the slab instead of two (MOTES_FIT_QR_FUSED = 0: the two-pass form).
#ifndef MOTES_FIT_QR_FUSED
#define MOTES_FIT_QR_FUSED 1
#endif
template <int K, class Ctx>
MOTES_HD void qr_step(const Ctx& ctx, int m, double* a, double diag_hk, FitState* st, double (&part)[kParams + 1]) {
    constexpr int n = kParams;
    constexpr int W = n + 1 - K;              // columns K .. n (n = the residual column)
    const bool leader = ctx.lane == 0;
    const double* colk = a + (size_t)K * m;
    double mine[W], c[W];
#if MOTES_FIT_QR_FUSED
#pragma unroll
    for (int j = 0; j < W; ++j) mine[j] = part[j];
#else
#pragma unroll
    for (int j = 0; j < W; ++j) mine[j] = 0.0;
    // (the point loops of the sweeps are not unrolled: six specialisations of this step, unrolled four times each,
    // made the kernel 104 KB of code and instruction fetch its first stall reason - ncu, round 2)
#pragma unroll kFitUnrollQr
    for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
        double aki = colk[i];
#pragma unroll
        for (int j = 0; j < W; ++j) mine[j] = fma(aki, a[(size_t)(K + j) * m + i], mine[j]);
    }
#endif
    ctx.template sum_vec<W>(mine, c);
#pragma unroll
    for (int j = 0; j < n + 1; ++j) part[j] = 0.0;
    double ek = trf_detail::ool_sqrt(diag_hk);
    double norm2 = c[0] + ek * ek;
    if (norm2 > 0.0) {
        double nrm = trf_detail::ool_sqrt(norm2);
        double v_piv = ek + nrm;
        double tau = trf_detail::ool_div(1.0, norm2 + nrm * ek);
        double wj[W];
#pragma unroll
        for (int j = 1; j < W; ++j) wj[j] = c[j] * tau;
        if (leader) {
            st->R[K * n + K] = -nrm;
#pragma unroll
            for (int j = 1; j < W; ++j) {
                double rkj = -wj[j] * v_piv;
                if (K + j < n) st->R[K * n + K + j] = rkj; else st->z[K] = rkj;
            }
        }
        // (the point loops of the sweeps are not unrolled: six specialisations of this step, unrolled four times each,
        // made the kernel 104 KB of code and instruction fetch its first stall reason - ncu, round 2)
#pragma unroll kFitUnrollQr
        for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
            double aki = colk[i];
            double nv[W];
#pragma unroll
            for (int j = 1; j < W; ++j) {
                nv[j] = fma(-wj[j], aki, a[(size_t)(K + j) * m + i]);
                a[(size_t)(K + j) * m + i] = nv[j];
            }
#if MOTES_FIT_QR_FUSED
            if (W > 1) {
#pragma unroll
                for (int j = 1; j < W; ++j) part[j - 1] = fma(nv[1], nv[j], part[j - 1]);
            }
#endif
        }
    } else {
        if (leader) st->z[K] = 0.0;
#if MOTES_FIT_QR_FUSED
        // (a column that is exactly zero: nothing to apply; the next column's products from the slab as it is)
        if (W > 1) {
            for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
                double aki = a[(size_t)(K + 1) * m + i];
#pragma unroll
                for (int j = 1; j < W; ++j) part[j - 1] = fma(aki, a[(size_t)(K + j) * m + i], part[j - 1]);
            }
        }
#endif
    }
    ctx.sync();
}

// Jacobian at x (residual in a[6m..7m), pow cache in unit), g = J^T f, then the QR of
// [J diag(d) ; diag(sqrt(diag_h))] by Householder reflections pivoting on the sparse rows, with f
// as a seventh column (trf.cuh explains the equivalence with scipy's SVD of the augmented matrix).
// g, R, z go to *st, written by the scope's lane 0 only (st may be shared by the scope).
template <class Ctx>
MOTES_HD void linearize(const Ctx& ctx, int m, const double* profile, double scale, const double (&x)[kParams],
                        const double (&lb)[kParams], const double (&ub)[kParams], const double* unit, double* a,
                        FitState* st) {
    using namespace trf_detail;
    constexpr int n = kParams;
    const bool leader = ctx.lane == 0;
    double h[n];
    fd_steps(x, lb, ub, h);
    const double* f0 = a + 6 * (size_t)m;
    double dx[n];
#pragma unroll
    for (int k = 0; k < n; ++k) dx[k] = (x[k] + h[k]) - x[k];
    double inv_dx[n];
#pragma unroll
    for (int k = 0; k < n; ++k) inv_dx[k] = ool_div(1.0, dx[k]);
    const double c1 = x[1] + h[1], a1 = x[2] + h[2];
    const double aa = x[2] * x[2], aa1 = a1 * a1;
    const double inv_aa = ool_div(1.0, aa), inv_aa1 = ool_div(1.0, aa1);
    double gp[n] = {0, 0, 0, 0, 0, 0};
#pragma unroll kFitUnrollLin
    for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
        const double r = (double)i;
        const double yi = profile[i] * scale, fi = f0[i], u0 = unit[i];
        // A, B, m move the value but not the pow argument.  c, alpha, beta move it by a part in ~1e8:
        // the shifted pow is the base pow times (s'/s)^-beta resp. s^-h, evaluated from the small
        // relative change of s = 1 + q (shifted_unit); a full pow only when the step is not small.
        const double off = r - x[1], off1 = r - c1;
        // (reciprocal multiplies: the differences q' - q below only need each q to a relative 1e-16)
        const double q = (off * off) * inv_aa;
        const double s = 1.0 + q;
        const double inv_s = 1.0 / s;
        double uc, ua, ub_;
        if (!shifted_unit(u0, (off1 * off1) * inv_aa, q, inv_s, x[3], &uc)) uc = moffat_unit_call(r, c1, x[2], x[3]);
        if (!shifted_unit(u0, (off * off) * inv_aa1, q, inv_s, x[3], &ua)) ua = moffat_unit_call(r, x[1], a1, x[3]);
        if (!shifted_unit_beta(u0, s, dx[3], &ub_)) ub_ = moffat_unit_call(r, x[1], x[2], x[3] + h[3]);
        double col[n];
        col[0] = ((moffat_value(u0, r, x[0] + h[0], x[4], x[5]) - yi) - fi) * inv_dx[0];
        col[1] = ((moffat_value(uc, r, x[0], x[4], x[5]) - yi) - fi) * inv_dx[1];
        col[2] = ((moffat_value(ua, r, x[0], x[4], x[5]) - yi) - fi) * inv_dx[2];
        col[3] = ((moffat_value(ub_, r, x[0], x[4], x[5]) - yi) - fi) * inv_dx[3];
        col[4] = ((moffat_value(u0, r, x[0], x[4] + h[4], x[5]) - yi) - fi) * inv_dx[4];
        col[5] = ((moffat_value(u0, r, x[0], x[4], x[5] + h[5]) - yi) - fi) * inv_dx[5];
#pragma unroll
        for (int k = 0; k < n; ++k) { a[(size_t)k * m + i] = col[k]; gp[k] += col[k] * fi; }
    }
    double g[n];
    ctx.template sum_vec<n>(gp, g);

    double v[n], dv[n], d[n], diag_h[n];
    coleman_li(x, g, lb, ub, v, dv);
#pragma unroll
    for (int i = 0; i < n; ++i) { d[i] = ool_sqrt(v[i]); diag_h[i] = g[i] * dv[i]; }
    if (leader) {
#pragma unroll
        for (int k = 0; k < n; ++k) st->g[k] = g[k];
        for (int i = 0; i < n * n; ++i) st->R[i] = 0.0;
    }
    // scaled Jacobian J diag(d); with it the products of its first column with every column (qr_step<0>)
    double part[n + 1];
#pragma unroll
    for (int k = 0; k < n + 1; ++k) part[k] = 0.0;
#pragma unroll 1
    for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
        double sc[n];
#pragma unroll
        for (int k = 0; k < n; ++k) { sc[k] = a[(size_t)k * m + i] * d[k]; a[(size_t)k * m + i] = sc[k]; }
#if MOTES_FIT_QR_FUSED
#pragma unroll
        for (int k = 0; k < n; ++k) part[k] = fma(sc[0], sc[k], part[k]);
        part[n] = fma(sc[0], f0[i], part[n]);
#endif
    }
    ctx.sync();
    qr_step<0>(ctx, m, a, diag_h[0], st, part);
    qr_step<1>(ctx, m, a, diag_h[1], st, part);
    qr_step<2>(ctx, m, a, diag_h[2], st, part);
    qr_step<3>(ctx, m, a, diag_h[3], st, part);
    qr_step<4>(ctx, m, a, diag_h[4], st, part);
    qr_step<5>(ctx, m, a, diag_h[5], st, part);
}

// One trial: advances *st exactly as one pass of scipy's inner loop body after the step has been
// chosen (or, for the first call, as the set-up in front of the outer loop).  *st may be shared by
// the scope: every lane reads it, lane 0 writes it, with scope barriers around the writes.
template <class Ctx>
MOTES_HD void fit_evaluate(const Ctx& ctx, int m, const double* profile, double scale, double* slab, FitState* st) {
    using namespace trf_detail;
    constexpr int n = kParams;
    const bool leader = ctx.lane == 0;
    double* unit = slab;
    double* a = slab;
    double* fcol = a + 6 * (size_t)m;
    double x[n], lb[n], ub[n];
    const bool first = st->first != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) { lb[i] = st->lb[i]; ub[i] = st->ub[i]; x[i] = first ? st->x[i] : st->x_new[i]; }
    const double cost_old = st->cost, radius = st->radius, step_h_norm = st->step_h_norm, predicted = st->predicted;
    const double step_norm = st->step_norm, x_norm = st->x_norm, alpha = st->alpha;
    const int nfev = st->nfev, njev = st->njev;
    double cost_new;
    const bool ok = residual_sweep_b(ctx, m, profile, scale, x, unit, fcol, &cost_new);
    ctx.sync();                                   // every lane has read *st; lane 0 may write from here on
    bool relinearize;
    if (first) {
        if (!ok) {
            if (leader) st->status = kFitNonFinite;
            ctx.sync();
            return;
        }
        if (leader) { st->cost = cost_new; st->nfev = 1; st->njev = 1; }
        relinearize = true;
    } else {
        if (!ok) {
            if (leader) { st->nfev = nfev + 1; st->radius = 0.25 * step_h_norm; }
            ctx.sync();
            return;
        }
        const double gain = cost_old - cost_new;
        double ratio;
        if (predicted > 0.0) ratio = gain / predicted;
        else if (predicted == 0.0 && gain == 0.0) ratio = 1.0;
        else ratio = 0.0;
        double radius_new = radius;
        if (ratio < 0.25) radius_new = 0.25 * step_h_norm;
        else if (ratio > 0.75 && step_h_norm > 0.95 * radius) radius_new = radius * 2.0;
        const bool f_ok = (gain < 1e-12 * cost_old) && (ratio > 0.25);
        const bool x_ok = step_norm < 1e-8 * (1e-8 + x_norm);
        int status = kStatusRunning;
        if (f_ok && x_ok) status = kFitBoth;
        else if (f_ok) status = kFitFtol;
        else if (x_ok) status = kFitXtol;
        if (leader) {
            st->nfev = nfev + 1;
            st->status = status;
            if (status == kStatusRunning) {
                st->alpha = alpha * (radius / radius_new);
                st->radius = radius_new;
            }
            if (gain > 0.0) {
#pragma unroll
                for (int i = 0; i < n; ++i) st->x[i] = x[i];
                st->cost = cost_new;
                st->njev = njev + 1;  // scipy evaluates the Jacobian at the accepted point even when it then stops
            }
        }
        relinearize = gain > 0.0 && status == kStatusRunning;
    }
    if (relinearize) linearize(ctx, m, profile, scale, x, lb, ub, unit, a, st);
    ctx.sync();
}

// Top of the outer loop up to the chosen step, in two halves around the SVD of the 6 x 6 triangle so that
// the SVD can be done by one thread (jacobi_svd6) or by an 8-lane team (jacobi_svd6_warp).  Both halves work
// on register copies; `write` says whether this thread stores the results into *st.
struct ProposeVars {
    double x[kParams], lb[kParams], ub[kParams], g[kParams], d[kParams], g_h[kParams];
    double radius, alpha, theta;
    int status;
};

// Returns false when the fit has ended (st->status final).
MOTES_HD bool propose_begin(const FitState* st, ProposeVars& pv, bool write, FitState* out) {
    using namespace trf_detail;
    constexpr int n = kParams;
    const int max_nfev = 100 * n;
    pv.status = st->status;
    if (pv.status != kStatusRunning) return false;
#pragma unroll
    for (int i = 0; i < n; ++i) { pv.x[i] = st->x[i]; pv.lb[i] = st->lb[i]; pv.ub[i] = st->ub[i]; pv.g[i] = st->g[i]; }
    pv.radius = st->radius;
    pv.alpha = st->alpha;
    const int first = st->first, nfev = st->nfev;
    double v[n], dv[n];
    coleman_li(pv.x, pv.g, pv.lb, pv.ub, v, dv);
    if (first) {
        double t[n];
#pragma unroll
        for (int i = 0; i < n; ++i) t[i] = ool_div(pv.x[i], ool_sqrt(v[i]));
        pv.radius = norm6(t);
        if (pv.radius == 0.0) pv.radius = 1.0;
        pv.alpha = 0.0;
    }
    double g_norm = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) g_norm = fmax(g_norm, fabs(pv.g[i] * v[i]));
    if (g_norm < 1e-8) pv.status = kFitGtol;
    else if (nfev >= max_nfev) pv.status = kFitMaxNfev;
    if (pv.status != kStatusRunning) {
        if (write) out->status = pv.status;
        return false;
    }
#pragma unroll
    for (int i = 0; i < n; ++i) { pv.d[i] = ool_sqrt(v[i]); pv.g_h[i] = pv.d[i] * pv.g[i]; }
    pv.theta = fmax(0.995, 1.0 - g_norm);
    return true;
}

// s (descending), suf = s * (U^T z), V (row-major, right singular vectors in columns) from the SVD of st->R.
MOTES_HD void propose_finish(int m, const FitState* st, ProposeVars& pv, const double* s, const double* suf,
                             const double* V, bool write, FitState* out) {
    using namespace trf_detail;
    constexpr int n = kParams;
    double R[n * n];
#pragma unroll
    for (int i = 0; i < n * n; ++i) R[i] = st->R[i];
    double p_h[n], step[n], step_h[n], x_new[n];
    more_step(m, s, suf, V, pv.radius, &pv.alpha, p_h);
    const double predicted = choose_step(pv.x, R, pv.g_h, p_h, pv.d, pv.radius, pv.lb, pv.ub, pv.theta, step, step_h);
#pragma unroll
    for (int i = 0; i < n; ++i) x_new[i] = pv.x[i] + step[i];
    push_inside(x_new, pv.lb, pv.ub, 0.0);
    if (write) {
#pragma unroll
        for (int i = 0; i < n; ++i) out->x_new[i] = x_new[i];
        out->predicted = predicted;
        out->step_h_norm = norm6(step_h);
        out->step_norm = norm6(step);
        out->x_norm = norm6(pv.x);
        out->radius = pv.radius;
        out->alpha = pv.alpha;
        out->first = 0;
    }
}

// One thread per fit (host simulation; also the reference for the team version in kernels.cuh).
MOTES_HD bool fit_propose(int m, FitState& st) {
    constexpr int n = kParams;
    ProposeVars pv;
    if (!propose_begin(&st, pv, true, &st)) return false;
    double s[n], suf[n], V[n * n];
    jacobi_svd6_serial(st.R, st.z, s, suf, V);
    propose_finish(m, &st, pv, s, suf, V, true, &st);
    return true;
}

// The whole fit through the phases (host simulation and the reference for the persistent kernel).
template <class Ctx>
MOTES_HD void trf_fit_moffat_phased(const Ctx& ctx, int m, const double* profile, double scale, double alpha0,
                                    double* slab, FitResult* res) {
    FitState st;
    fit_start(m, profile, scale, alpha0, &st);
    while (st.status == kStatusRunning) {
        fit_evaluate(ctx, m, profile, scale, slab, &st);
        if (!fit_propose(m, st)) break;
    }
    for (int i = 0; i < kParams; ++i) res->x[i] = st.x[i];
    res->cost = st.cost;
    res->nfev = st.nfev;
    res->njev = st.njev;
    res->status = st.status;
}

}  // namespace motes
