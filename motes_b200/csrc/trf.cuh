// Bounded Trust-Region-Reflective Moffat fit, one lane context (one warp) per profile.
//
// Replaces, for the six-parameter Moffat model, the reference's call
//   scipy.optimize.least_squares(moffat_resid, x0, bounds, method="trf", ftol=1e-12)
// (/root/reference/src/motes/tracing.py:550-607) together with the start-point and
// bound construction in front of it (tracing.py:578-588).  The iteration follows
// scipy's trf_bounds (optimize/_lsq/trf.py:206-410, select_step :129-203, helpers in
// _lsq/common.py, finite differences in _numdiff.py) decision for decision; what is
// done differently is how the linear algebra is carried out:
//
//   * the m x 6 Jacobian lives column-major in a per-warp slab, rows strided over the
//     32 lanes; all 6-vector logic is replicated in every lane (reductions are
//     butterfly all-reduces, so the lanes agree bit for bit);
//   * the SVD of the (m+6) x 6 augmented matrix [J_h; diag(sqrt(diag_h))] is obtained as
//     a QR factorisation (Householder on the sparse rows == modified Gram-Schmidt with the
//     residual carried as a 7th column) followed by a one-sided Jacobi SVD of the 6 x 6
//     triangle R; U^T f = U_R^T (Q^T f);
//   * every product with J_h in select_step is taken with R instead:
//     |J_h s|^2 + s.diag_h.s == |R s|^2, so the m-row matrix is not touched again;
//   * the three forward differences in A, B and m reuse the pow() of the base point (the
//     reference recomputes the identical value), so an iteration costs 4 pow sweeps, not 7.
#pragma once

#include "common.cuh"

namespace motes {

struct FitResult {
    double x[kParams];
    double cost;
    int nfev;
    int njev;
    int status;
};

// Per-fit scratch: caller provides it (shared memory on the device).
//   y    [m]     profile being fitted
//   unit [m]     (1+u)^-beta at the current base point
//   a    [7*m]   columns 0..5: Jacobian (then J_h, then destroyed by QR); column 6: residual
//   w, v [36]    Jacobi SVD work (serial variant); vout [36] right singular vectors, row-major
struct FitScratch {
    double* y;
    double* unit;
    double* a;
    double* w;
    double* v;
    double* vout;
};

MOTES_HD size_t fit_scratch_doubles(int m) { return (size_t)9 * m + 108; }

MOTES_HD FitScratch carve_fit_scratch(double* slab, int m) {
    FitScratch ws;
    ws.y = slab;
    ws.unit = ws.y + m;
    ws.a = ws.unit + m;
    ws.w = ws.a + 7 * (size_t)m;
    ws.v = ws.w + 36;
    ws.vout = ws.v + 36;
    return ws;
}

namespace trf_detail {

// The step-selection logic runs once per fit and round on one thread (or one 8-lane team): its dozens of FP64
// divisions and square roots are issued out of line so that the fit kernel's instruction footprint stays within
// the instruction cache (ncu on B200: 62 % hit rate with everything inlined).  Same IEEE operations, same bits.
#if defined(__CUDA_ARCH__)
__device__ __noinline__ double ool_div(double a, double b) { return a / b; }
__device__ __noinline__ double ool_sqrt(double a) { return sqrt(a); }
#else
inline double ool_div(double a, double b) { return a / b; }
inline double ool_sqrt(double a) { return sqrt(a); }
#endif

MOTES_HD double norm6(const double* x) {
    double s = 0.0;
    for (int i = 0; i < kParams; ++i) s += x[i] * x[i];
    return ool_sqrt(s);
}
MOTES_HD double dot6(const double* a, const double* b) {
    double s = 0.0;
    for (int i = 0; i < kParams; ++i) s += a[i] * b[i];
    return s;
}
MOTES_HD bool inside(const double* x, const double* lb, const double* ub) {
    bool ok = true;
    for (int i = 0; i < kParams; ++i) ok = ok && (x[i] >= lb[i]) && (x[i] <= ub[i]);
    return ok;
}

// make_strictly_feasible (_lsq/common.py:440-464); rstep == 0 -> nextafter.
MOTES_HD void push_inside(double* x, const double* lb, const double* ub, double rstep) {
    for (int i = 0; i < kParams; ++i) {
        int active = 0;
        if (rstep == 0.0) {
            if (x[i] <= lb[i]) active = -1;
            if (x[i] >= ub[i]) active = 1;
        } else {
            double below = x[i] - lb[i], above = ub[i] - x[i];
            double tol_lo = rstep * fmax(1.0, fabs(lb[i]));
            double tol_hi = rstep * fmax(1.0, fabs(ub[i]));
            if (is_finite(lb[i]) && below <= fmin(above, tol_lo)) active = -1;
            if (is_finite(ub[i]) && above <= fmin(below, tol_hi)) active = 1;
        }
        double xn = x[i];
        if (active == -1) xn = (rstep == 0.0) ? nextafter(lb[i], ub[i]) : lb[i] + rstep * fmax(1.0, fabs(lb[i]));
        if (active == 1) xn = (rstep == 0.0) ? nextafter(ub[i], lb[i]) : ub[i] - rstep * fmax(1.0, fabs(ub[i]));
        if (xn < lb[i] || xn > ub[i]) xn = 0.5 * (lb[i] + ub[i]);
        x[i] = xn;
    }
}

// CL_scaling_vector (_lsq/common.py:467-508)
MOTES_HD void coleman_li(const double* x, const double* g, const double* lb, const double* ub,
                         double* v, double* dv) {
    for (int i = 0; i < kParams; ++i) {
        v[i] = 1.0;
        dv[i] = 0.0;
        if (g[i] < 0.0 && is_finite(ub[i])) { v[i] = ub[i] - x[i]; dv[i] = -1.0; }
        if (g[i] > 0.0 && is_finite(lb[i])) { v[i] = x[i] - lb[i]; dv[i] = 1.0; }
    }
}

// step_size_to_bound (_lsq/common.py:372-398); hits[i] = sign(s[i]) where the minimum is attained.
MOTES_HD double stride_to_bound(const double* x, const double* s, const double* lb, const double* ub,
                                int* hits) {
    double steps[kParams];
    double t = INFINITY;
    for (int i = 0; i < kParams; ++i) {
        steps[i] = INFINITY;
        if (s[i] != 0.0) steps[i] = fmax(ool_div(lb[i] - x[i], s[i]), ool_div(ub[i] - x[i], s[i]));
        // numpy.min propagates NaN
        if (steps[i] < t || steps[i] != steps[i]) t = steps[i];
    }
    if (hits) {
        for (int i = 0; i < kParams; ++i) {
            int sg = (s[i] > 0.0) - (s[i] < 0.0);
            hits[i] = (steps[i] == t) ? sg : 0;
        }
    }
    return t;
}

// |R s|^2 for upper-triangular R (row-major 6x6) == |J_h s|^2 + s.diag_h.s
MOTES_HD void r_times(const double* R, const double* s, double* out) {
    for (int i = 0; i < kParams; ++i) {
        double acc = 0.0;
        for (int j = i; j < kParams; ++j) acc += R[i * kParams + j] * s[j];
        out[i] = acc;
    }
}

// evaluate_quadratic (_lsq/common.py:325-361)
MOTES_HD double model_value(const double* R, const double* g, const double* s) {
    double rs[kParams];
    r_times(R, s, rs);
    return 0.5 * dot6(rs, rs) + dot6(s, g);
}

// minimize_quadratic_1d (_lsq/common.py:302-322)
MOTES_HD void line_minimum(double a, double b, double lo, double hi, double c, double* t_out, double* y_out) {
    double tb = lo, yb = lo * (a * lo + b) + c;
    double y = hi * (a * hi + b) + c;
    if (y < yb) { yb = y; tb = hi; }
    if (a != 0.0) {
        double vertex = ool_div(-0.5 * b, a);
        if (lo < vertex && vertex < hi) {
            y = vertex * (a * vertex + b) + c;
            if (y < yb) { yb = y; tb = vertex; }
        }
    }
    *t_out = tb;
    *y_out = yb;
}

// solve_lsq_trust_region (_lsq/common.py:57-168).  s: singular values (descending),
// suf = s * (U^T f), V row-major 6x6 with right singular vectors in columns.
MOTES_HD void more_step(int m, const double* s, const double* suf, const double* V, double radius,
                        double* alpha_io, double* p_h) {
    double tmp[kParams];
    bool full_rank = (m >= kParams) && (s[kParams - 1] > kEps * m * s[0]);
    if (full_rank) {
        for (int j = 0; j < kParams; ++j) tmp[j] = ool_div(ool_div(suf[j], s[j]), s[j]);   // uf / s
        for (int i = 0; i < kParams; ++i) {
            double acc = 0.0;
            for (int j = 0; j < kParams; ++j) acc += V[i * kParams + j] * tmp[j];
            p_h[i] = -acc;
        }
        if (norm6(p_h) <= radius) { *alpha_io = 0.0; return; }
    }
    double upper = ool_div(norm6(suf), radius);
    double lower = 0.0;
    if (full_rank) {
        double pn = 0.0, dsum = 0.0;
        for (int j = 0; j < kParams; ++j) {
            double den = s[j] * s[j];
            double q = ool_div(suf[j], den);
            pn += q * q;
            dsum += ool_div(suf[j] * suf[j], den * den * den);
        }
        pn = ool_sqrt(pn);
        double phi = pn - radius, dphi = ool_div(-dsum, pn);
        lower = ool_div(-phi, dphi);
    }
    double alpha = *alpha_io;
    if (!full_rank && alpha == 0.0) alpha = fmax(0.001 * upper, ool_sqrt(lower * upper));
#pragma unroll 1
    for (int it = 0; it < 10; ++it) {
        if (alpha < lower || alpha > upper) alpha = fmax(0.001 * upper, ool_sqrt(lower * upper));
        double pn = 0.0, dsum = 0.0;
        for (int j = 0; j < kParams; ++j) {
            double den = s[j] * s[j] + alpha;
            double q = ool_div(suf[j], den);
            pn += q * q;
            dsum += ool_div(suf[j] * suf[j], den * den * den);
        }
        pn = ool_sqrt(pn);
        double phi = pn - radius, dphi = ool_div(-dsum, pn);
        if (phi < 0.0) upper = alpha;
        double ratio = ool_div(phi, dphi);
        lower = fmax(lower, alpha - ratio);
        alpha -= ool_div((phi + radius) * ratio, radius);
        if (fabs(phi) < 0.01 * radius) break;
    }
    for (int j = 0; j < kParams; ++j) tmp[j] = ool_div(suf[j], s[j] * s[j] + alpha);
    for (int i = 0; i < kParams; ++i) {
        double acc = 0.0;
        for (int j = 0; j < kParams; ++j) acc += V[i * kParams + j] * tmp[j];
        p_h[i] = -acc;
    }
    double scale = ool_div(radius, norm6(p_h));
    for (int i = 0; i < kParams; ++i) p_h[i] *= scale;
    *alpha_io = alpha;
}

// select_step (_lsq/trf.py:129-203).  Returns the predicted reduction; step / step_h out.
MOTES_HD double choose_step(const double* x, const double* R, const double* g_h, const double* p_h_in,
                            const double* d, double radius, const double* lb, const double* ub,
                            double theta, double* step, double* step_h) {
    double p[kParams], p_h[kParams], xp[kParams];
    for (int i = 0; i < kParams; ++i) { p_h[i] = p_h_in[i]; p[i] = d[i] * p_h[i]; xp[i] = x[i] + p[i]; }
    if (inside(xp, lb, ub)) {
        for (int i = 0; i < kParams; ++i) { step[i] = p[i]; step_h[i] = p_h[i]; }
        return -model_value(R, g_h, p_h);
    }
    int hits[kParams];
    double p_stride = stride_to_bound(x, p, lb, ub, hits);
    double r_h[kParams], r[kParams], x_on[kParams];
    for (int i = 0; i < kParams; ++i) {
        r_h[i] = hits[i] ? -p_h[i] : p_h[i];
        r[i] = d[i] * r_h[i];
        p[i] *= p_stride;
        p_h[i] *= p_stride;
        x_on[i] = x[i] + p[i];
    }
    // intersect_trust_region(p_h, r_h, radius): positive root
    double to_tr;
    {
        double a = dot6(r_h, r_h), b = dot6(p_h, r_h), c = dot6(p_h, p_h) - radius * radius;
        double disc = ool_sqrt(b * b - a * c);
        double q = -(b + copysign(disc, b));
        double t1 = ool_div(q, a), t2 = ool_div(c, q);
        to_tr = (t1 < t2) ? t2 : t1;
    }
    double to_bound = stride_to_bound(x_on, r, lb, ub, nullptr);
    double r_stride = fmin(to_bound, to_tr);
    double r_lo, r_hi;
    if (r_stride > 0.0) {
        r_lo = ool_div((1.0 - theta) * p_stride, r_stride);
        r_hi = (r_stride == to_bound) ? theta * to_bound : to_tr;
    } else {
        r_lo = 0.0;
        r_hi = -1.0;
    }
    double r_value = INFINITY;
    if (r_lo <= r_hi) {
        // build_quadratic_1d(J_h, g_h, r_h, s0=p_h, diag=diag_h) with R in place of [J_h; sqrt(diag_h)]
        double vr[kParams], ur[kParams];
        r_times(R, r_h, vr);
        r_times(R, p_h, ur);
        double a = 0.5 * dot6(vr, vr);
        double b = dot6(g_h, r_h) + dot6(ur, vr);
        double c = 0.5 * dot6(ur, ur) + dot6(g_h, p_h);
        double t;
        line_minimum(a, b, r_lo, r_hi, c, &t, &r_value);
        for (int i = 0; i < kParams; ++i) {
            r_h[i] = r_h[i] * t + p_h[i];
            r[i] = r_h[i] * d[i];
        }
    }
    for (int i = 0; i < kParams; ++i) { p[i] *= theta; p_h[i] *= theta; }
    double p_value = model_value(R, g_h, p_h);

    double ag_h[kParams], ag[kParams];
    for (int i = 0; i < kParams; ++i) { ag_h[i] = -g_h[i]; ag[i] = d[i] * ag_h[i]; }
    double ag_to_tr = ool_div(radius, norm6(ag_h));
    double ag_to_bound = stride_to_bound(x, ag, lb, ub, nullptr);
    double ag_stride = (ag_to_bound < ag_to_tr) ? theta * ag_to_bound : ag_to_tr;
    double ag_value;
    {
        double va[kParams];
        r_times(R, ag_h, va);
        double a = 0.5 * dot6(va, va);
        double b = dot6(g_h, ag_h);
        double t;
        line_minimum(a, b, 0.0, ag_stride, 0.0, &t, &ag_value);
        ag_stride = t;
    }
    if (p_value < r_value && p_value < ag_value) {
        for (int i = 0; i < kParams; ++i) { step[i] = p[i]; step_h[i] = p_h[i]; }
        return -p_value;
    }
    if (r_value < p_value && r_value < ag_value) {
        for (int i = 0; i < kParams; ++i) { step[i] = r[i]; step_h[i] = r_h[i]; }
        return -r_value;
    }
    for (int i = 0; i < kParams; ++i) { step[i] = ag[i] * ag_stride; step_h[i] = ag_h[i] * ag_stride; }
    return -ag_value;
}

// _compute_absolute_step + _adjust_scheme_to_bounds('1-sided') (_numdiff.py:147-203, 14-91)
MOTES_HD void fd_steps(const double* x, const double* lb, const double* ub, double* h) {
    for (int i = 0; i < kParams; ++i) {
        double sign = (x[i] >= 0.0) ? 1.0 : -1.0;
        double hi = kFdRelStep * sign * fmax(1.0, fabs(x[i]));
        double below = x[i] - lb[i], above = ub[i] - x[i];
        double trial = x[i] + hi;
        bool violated = (trial < lb[i]) || (trial > ub[i]);
        bool fitting = fabs(hi) <= fmax(below, above);
        if (violated && fitting) hi = -hi;
        if (!fitting) hi = (above >= below) ? above : -below;
        h[i] = hi;
    }
}

}  // namespace trf_detail

// One-sided Jacobi SVD of the upper-triangular 6x6 R (row-major).  On return
// s[] (descending), V (row-major, singular vectors in columns) and suf = s * (U^T z).
// w / v are 36-double work arrays every lane reads and lane 0 writes.
template <class Ctx>
MOTES_HD void jacobi_svd6(const Ctx& ctx, const double* R, const double* z, double* w, double* v,
                          double* s, double* suf, double* V) {
    constexpr int n = kParams;
    if (ctx.lane == 0) {
        for (int i = 0; i < n * n; ++i) { w[i] = R[i]; v[i] = 0.0; }
        for (int i = 0; i < n; ++i) v[i * n + i] = 1.0;
    }
    ctx.sync();
    for (int sweep = 0; sweep < 40; ++sweep) {
        int rotated = 0;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                double a = 0.0, b = 0.0, c = 0.0;
                double wp[n], wq[n], vp[n], vq[n];
                for (int r = 0; r < n; ++r) {
                    wp[r] = w[r * n + p]; wq[r] = w[r * n + q];
                    vp[r] = v[r * n + p]; vq[r] = v[r * n + q];
                    a += wp[r] * wp[r]; b += wq[r] * wq[r]; c += wp[r] * wq[r];
                }
                ctx.sync();
                if (c != 0.0 && fabs(c) > kEps * trf_detail::ool_sqrt(a * b)) {
                    rotated = 1;
                    double zeta = trf_detail::ool_div(b - a, 2.0 * c);
                    double t = trf_detail::ool_div(copysign(1.0, zeta), fabs(zeta) + trf_detail::ool_sqrt(1.0 + zeta * zeta));
                    double cs = trf_detail::ool_div(1.0, trf_detail::ool_sqrt(1.0 + t * t));
                    double sn = cs * t;
                    if (ctx.lane == 0) {
                        for (int r = 0; r < n; ++r) {
                            w[r * n + p] = cs * wp[r] - sn * wq[r];
                            w[r * n + q] = sn * wp[r] + cs * wq[r];
                            v[r * n + p] = cs * vp[r] - sn * vq[r];
                            v[r * n + q] = sn * vp[r] + cs * vq[r];
                        }
                    }
                }
                ctx.sync();
            }
        }
        if (!rotated) break;
    }
    // singular values, s*uf, and a descending selection sort of the column order
    double sv[n], sz[n];
    int order[n];
    for (int j = 0; j < n; ++j) {
        double nn = 0.0, zz = 0.0;
        for (int r = 0; r < n; ++r) { nn += w[r * n + j] * w[r * n + j]; zz += w[r * n + j] * z[r]; }
        sv[j] = trf_detail::ool_sqrt(nn);
        sz[j] = zz;
        order[j] = j;
    }
    for (int i = 0; i < n - 1; ++i)
        for (int j = i + 1; j < n; ++j)
            if (sv[order[j]] > sv[order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
    for (int j = 0; j < n; ++j) {
        s[j] = sv[order[j]];
        suf[j] = sz[order[j]];
        for (int r = 0; r < n; ++r) V[r * n + j] = v[r * n + order[j]];
    }
    ctx.sync();
}

// One-thread form of the warp variant below, operation for operation: the same round-robin pair ordering, the same
// rotation formulas and the same association of the six-term sums (sum8's xor tree with two zero rows:
// ((x0 + x1) + (x2 + x3)) + (x4 + x5)), so that a fit takes the same steps whether its step is chosen by one thread
// (large batches: lock-step rounds) or by an 8-lane team (small batches: the group kernel).  wc / vc: column-major
// 6 x 6 scratch.
MOTES_HD double jacobi_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
MOTES_HD double jacobi_tree6(double x0, double x1, double x2, double x3, double x4, double x5) {
    return ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + 0.0);
}
MOTES_HD double jacobi_dot6(const double* p, const double* q) {
    return jacobi_tree6(p[0] * q[0], p[1] * q[1], p[2] * q[2], p[3] * q[3], p[4] * q[4], p[5] * q[5]);
}
MOTES_HD int jacobi_rotate_cols(double* wp, double* wq, double* vp, double* vq, double a, double b, double c) {
    if (c == 0.0 || !(fabs(c) > kEps * trf_detail::ool_sqrt(a * b))) return 0;
    const double d = b - a;
    const double t = trf_detail::ool_div((2.0 * c) * (d >= 0.0 ? 1.0 : -1.0), fabs(d) + trf_detail::ool_sqrt(d * d + 4.0 * (c * c)));
    const double cs = jacobi_rsqrt(1.0 + t * t);
    const double sn = cs * t;
    for (int r = 0; r < kParams; ++r) {
        const double a_p = wp[r], a_q = wq[r], b_p = vp[r], b_q = vq[r];
        wp[r] = cs * a_p - sn * a_q;
        wq[r] = sn * a_p + cs * a_q;
        vp[r] = cs * b_p - sn * b_q;
        vq[r] = sn * b_p + cs * b_q;
    }
    return 1;
}
MOTES_HD void jacobi_svd6_serial(const double* R, const double* z, double* s, double* suf, double* V) {
    constexpr int n = kParams;
    // the five rounds of three disjoint pairs of jacobi_svd6_warp
    const unsigned char pairs[5][6] = {{0, 5, 1, 4, 2, 3}, {0, 4, 3, 5, 1, 2}, {0, 3, 2, 4, 1, 5}, {0, 2, 1, 3, 4, 5}, {0, 1, 2, 5, 3, 4}};
    double wc[n * n], vc[n * n];
    for (int j = 0; j < n; ++j)
        for (int r = 0; r < n; ++r) { wc[j * n + r] = R[r * n + j]; vc[j * n + r] = (r == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 40; ++sweep) {
        int rot = 0;
        for (int round = 0; round < 5; ++round) {
            double a[3], b[3], c[3];
            for (int k = 0; k < 3; ++k) {
                const double* wp = wc + pairs[round][2 * k] * n;
                const double* wq = wc + pairs[round][2 * k + 1] * n;
                a[k] = jacobi_dot6(wp, wp);
                b[k] = jacobi_dot6(wq, wq);
                c[k] = jacobi_dot6(wp, wq);
            }
            for (int k = 0; k < 3; ++k) {
                const int P = pairs[round][2 * k], Q = pairs[round][2 * k + 1];
                rot |= jacobi_rotate_cols(wc + P * n, wc + Q * n, vc + P * n, vc + Q * n, a[k], b[k], c[k]);
            }
        }
        if (!rot) break;
    }
    double sv[n], sz[n];
    for (int j = 0; j < n; ++j) {
        sv[j] = trf_detail::ool_sqrt(jacobi_dot6(wc + j * n, wc + j * n));
        sz[j] = jacobi_dot6(wc + j * n, z);
    }
    for (int j = 0; j < n; ++j) {
        int rank = 0;
        for (int k = 0; k < n; ++k) rank += (sv[k] > sv[j]) || (sv[k] == sv[j] && k < j);
        s[rank] = sv[j];
        suf[rank] = sz[j];
        for (int r = 0; r < n; ++r) V[r * n + rank] = vc[j * n + r];
    }
}

#if defined(__CUDACC__)
// Warp variant of the same decomposition.  Lane L (mod 8) < 6 owns row L of W (= R V) and of V;
// the 15 column pairs of a sweep are visited in 5 rounds of 3 disjoint pairs (round-robin
// ordering), so a round costs one 9-value reduction over the rows and one rotation latency
// instead of three.  Every lane computes the rotation parameters from the all-reduced sums, so
// all lanes stay in lock step.  Results: s, suf replicated in registers; V written to vout
// (shared memory, row-major, columns sorted by descending singular value).
__device__ __forceinline__ double sum8(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

template <int P, int Q>
__device__ __forceinline__ void jacobi_pair_sums(const double (&w)[kParams], double& a, double& b, double& c) {
    a = sum8(w[P] * w[P]);
    b = sum8(w[Q] * w[Q]);
    c = sum8(w[P] * w[Q]);
}

template <int P, int Q>
__device__ __forceinline__ int jacobi_pair_rotate(double (&w)[kParams], double (&v)[kParams], double a, double b, double c) {
    if (c == 0.0 || !(fabs(c) > kEps * sqrt(a * b))) return 0;
    double d = b - a;
    double t = (2.0 * c) * (d >= 0.0 ? 1.0 : -1.0) / (fabs(d) + sqrt(d * d + 4.0 * (c * c)));
    double cs = rsqrt(1.0 + t * t);
    double sn = cs * t;
    double wp = w[P], wq = w[Q], vp = v[P], vq = v[Q];
    w[P] = cs * wp - sn * wq;
    w[Q] = sn * wp + cs * wq;
    v[P] = cs * vp - sn * vq;
    v[Q] = sn * vp + cs * vq;
    return 1;
}

template <int P0, int Q0, int P1, int Q1, int P2, int Q2>
__device__ __forceinline__ int jacobi_round(double (&w)[kParams], double (&v)[kParams]) {
    double a0, b0, c0, a1, b1, c1, a2, b2, c2;
    jacobi_pair_sums<P0, Q0>(w, a0, b0, c0);
    jacobi_pair_sums<P1, Q1>(w, a1, b1, c1);
    jacobi_pair_sums<P2, Q2>(w, a2, b2, c2);
    int r = jacobi_pair_rotate<P0, Q0>(w, v, a0, b0, c0);
    r |= jacobi_pair_rotate<P1, Q1>(w, v, a1, b1, c1);
    r |= jacobi_pair_rotate<P2, Q2>(w, v, a2, b2, c2);
    return r;
}

// `warm`: vout holds the V of the previous outer iteration and the sweeps start from W = R V_prev
// (any orthogonal start converges to the same decomposition).  Measured on B200 (round 1): no gain
// in fit time and one more stopping-point flip on the faint-source golden case, so the solver
// calls it cold, like scipy's fresh SVD every iteration.
__device__ __forceinline__ void jacobi_svd6_warp(int lane, const double* R, const double* z, double* s, double* suf,
                                                 double* vout, bool warm) {
    constexpr int n = kParams;
    // rows are replicated in the four 8-lane groups so that every lane sees the same sums
    const int row = lane & 7;
    const int rr = row < n ? row : 0;
    double w[n], v[n];
    if (warm) {
#pragma unroll
        for (int j = 0; j < n; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < n; ++k) acc = fma(R[rr * n + k], vout[k * n + j], acc);
            w[j] = (row < n) ? acc : 0.0;
            v[j] = (row < n) ? vout[rr * n + j] : 0.0;
        }
        __syncwarp();
    } else {
#pragma unroll
        for (int j = 0; j < n; ++j) {
            w[j] = (row < n) ? R[rr * n + j] : 0.0;
            v[j] = (row == j) ? 1.0 : 0.0;
        }
    }
    const double zl = (row < n) ? z[rr] : 0.0;
    for (int sweep = 0; sweep < 40; ++sweep) {
        int rot = 0;
        rot |= jacobi_round<0, 5, 1, 4, 2, 3>(w, v);
        rot |= jacobi_round<0, 4, 3, 5, 1, 2>(w, v);
        rot |= jacobi_round<0, 3, 2, 4, 1, 5>(w, v);
        rot |= jacobi_round<0, 2, 1, 3, 4, 5>(w, v);
        rot |= jacobi_round<0, 1, 2, 5, 3, 4>(w, v);
        if (!__any_sync(0xffffffffu, rot)) break;      // the four 8-lane groups may hold different matrices
    }
    double sv[n], sz[n];
#pragma unroll
    for (int j = 0; j < n; ++j) {
        sv[j] = sqrt(sum8(w[j] * w[j]));
        sz[j] = sum8(w[j] * zl);
    }
    // descending order of the singular values (replicated, branch-free selection of registers)
    int order[n];
#pragma unroll
    for (int j = 0; j < n; ++j) {
        int rank = 0;
#pragma unroll
        for (int k = 0; k < n; ++k) rank += (sv[k] > sv[j]) || (sv[k] == sv[j] && k < j);
        order[j] = rank;          // column j goes to position rank
    }
#pragma unroll
    for (int j = 0; j < n; ++j) {
#pragma unroll
        for (int pos = 0; pos < n; ++pos) {
            if (order[j] == pos) {
                s[pos] = sv[j];
                suf[pos] = sz[j];
                if (row < n) vout[row * n + pos] = v[j];
            }
        }
    }
    __syncwarp();
}
#endif

// Residual sweep at parameters xe: writes f into out[] (and pow values into unit_out when given),
// returns 0.5 * f.f through *cost, and whether every residual is finite.
template <class Ctx>
MOTES_HD bool residual_sweep(const Ctx& ctx, int m, const double* y, const double* xe, double* unit_out,
                             double* out, double* cost) {
    double acc = 0.0;
    int finite = 1;
    for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
        double r = (double)i;
        double u = moffat_unit(r, xe[1], xe[2], xe[3]);
        double f = moffat_value(u, r, xe[0], xe[4], xe[5]) - y[i];
        if (unit_out) unit_out[i] = u;
        out[i] = f;
        acc += f * f;
        finite &= (int)is_finite(f);
    }
    *cost = 0.5 * ctx.sum(acc);
    return ctx.all(finite) != 0;
}

// Forward-difference Jacobian at x (residual f in a[6*m..], pow cache in unit[]) into a[0..6*m).
template <class Ctx>
MOTES_HD void fd_jacobian(const Ctx& ctx, int m, const double* y, const double* x, const double* lb,
                          const double* ub, const double* unit, double* a) {
    double h[kParams];
    trf_detail::fd_steps(x, lb, ub, h);
    const double* f0 = a + 6 * (size_t)m;
    for (int k = 0; k < kParams; ++k) {
        double xk = x[k] + h[k];
        double dx = xk - x[k];
        double xe[kParams];
        for (int j = 0; j < kParams; ++j) xe[j] = x[j];
        xe[k] = xk;
        double* col = a + (size_t)k * m;
        bool shape = (k == 1 || k == 2 || k == 3);      // c, alpha, beta change the pow argument
        for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
            double r = (double)i;
            double u = shape ? moffat_unit(r, xe[1], xe[2], xe[3]) : unit[i];
            double f = moffat_value(u, r, xe[0], xe[4], xe[5]) - y[i];
            col[i] = (f - f0[i]) / dx;
        }
    }
    ctx.sync();
}

// The fit.  `profile` may live anywhere readable (global or shared); the fitted data are
// profile[i] * scale (tracing.py:441,513: the caller's data_scaling_factor; pass 1.0 for none);
// alpha0 = seeing / pixel_resolution.
template <class Ctx>
MOTES_HD void trf_fit_moffat(const Ctx& ctx, int m, const double* profile, double scale, double alpha0,
                             const FitScratch& ws, FitResult* res) {
    using namespace trf_detail;
    constexpr int n = kParams;
    double* y = ws.y;
    double* a = ws.a;
    double* fcol = a + 6 * (size_t)m;

    // ---- start point and bounds (tracing.py:578-588); replicated serial scan of the profile
    double top1 = -INFINITY, top2 = -INFINITY, top3 = -INFINITY;
    int peak = 0, n_nan = 0;
    bool peak_nan = false;
    for (int i = ctx.lane; i < m; i += Ctx::kLanes) y[i] = profile[i] * scale;
    ctx.sync();
    for (int i = 0; i < m; ++i) {
        double val = y[i];
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
    // np.median(concatenate(col[:5], col[-5:]))
    double ends[10];
    int n_ends = 0;
    for (int i = 0; i < 5 && i < m; ++i) ends[n_ends++] = y[i];
    for (int i = (m > 5 ? m - 5 : 0); i < m; ++i) ends[n_ends++] = y[i];
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

    double x[n] = {amp0, (double)peak, alpha0, kBetaStart, level0, 0.0};
    double lb[n] = {0.0, (double)peak - 1.0, 0.0, 0.0, -INFINITY, -INFINITY};
    double ub[n] = {INFINITY, (double)peak + 1.0, 5.0 * alpha0, 5.0, INFINITY, INFINITY};

    for (int i = 0; i < n; ++i) res->x[i] = x[i];
    res->cost = NAN;
    res->nfev = 0;
    res->njev = 0;
    for (int i = 0; i < n; ++i)
        if (!(lb[i] < ub[i])) { res->status = kFitBadBounds; return; }
    if (!inside(x, lb, ub)) { res->status = kFitX0Infeasible; return; }
    push_inside(x, lb, ub, 1e-10);

    double cost;
    bool finite = residual_sweep(ctx, m, y, x, ws.unit, fcol, &cost);
    ctx.sync();
    if (!finite) { res->status = kFitNonFinite; return; }
    fd_jacobian(ctx, m, y, x, lb, ub, ws.unit, a);
    int nfev = 1, njev = 1;
    const int max_nfev = 100 * n;

    // g = J^T f
    double g[n];
    {
        double part[n] = {0, 0, 0, 0, 0, 0};
        for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
            double fi = fcol[i];
            for (int k = 0; k < n; ++k) part[k] += a[(size_t)k * m + i] * fi;
        }
        for (int k = 0; k < n; ++k) g[k] = ctx.sum(part[k]);
    }

    double v[n], dv[n];
    coleman_li(x, g, lb, ub, v, dv);
    double radius;
    {
        double t[n];
        for (int i = 0; i < n; ++i) t[i] = x[i] / sqrt(v[i]);
        radius = norm6(t);
        if (radius == 0.0) radius = 1.0;
    }
    double alpha = 0.0;
    int status = -100;          // "None"
    double g_norm = 0.0;
    double x_new[n];
    double cost_new = cost;

    while (true) {
        coleman_li(x, g, lb, ub, v, dv);
        g_norm = 0.0;
        for (int i = 0; i < n; ++i) g_norm = fmax(g_norm, fabs(g[i] * v[i]));
        if (g_norm < 1e-8) status = kFitGtol;
        if (status != -100 || nfev == max_nfev) break;

        double d[n], diag_h[n], g_h[n];
        for (int i = 0; i < n; ++i) {
            d[i] = sqrt(v[i]);
            diag_h[i] = g[i] * dv[i];
            g_h[i] = d[i] * g[i];
        }
        // ---- QR of [J_h | f] with the sparse rows sqrt(diag_h) as pivots
        double R[n * n], z[n];
        for (int i = 0; i < n * n; ++i) R[i] = 0.0;
        for (int k = 0; k < n; ++k) {
            double* colk = a + (size_t)k * m;
            for (int i = ctx.lane; i < m; i += Ctx::kLanes) colk[i] *= d[k];
        }
        ctx.sync();
        for (int k = 0; k < n; ++k) {
            const double* colk = a + (size_t)k * m;
            double part[n + 1];
            for (int j = k; j <= n; ++j) part[j] = 0.0;
            for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
                double aki = colk[i];
                for (int j = k; j <= n; ++j) part[j] = fma(aki, a[(size_t)j * m + i], part[j]);
            }
            double c[n + 1];
            for (int j = k; j <= n; ++j) c[j] = ctx.sum(part[j]);
            double ek = sqrt(diag_h[k]);
            double norm2 = c[k] + ek * ek;
            if (norm2 > 0.0) {
                double nrm = sqrt(norm2);
                double v_piv = ek + nrm;
                double tau = 1.0 / (norm2 + nrm * ek);
                R[k * n + k] = -nrm;
                double wj[n + 1];
                for (int j = k + 1; j <= n; ++j) {
                    wj[j] = c[j] * tau;
                    double rkj = -wj[j] * v_piv;
                    if (j < n) R[k * n + j] = rkj; else z[k] = rkj;
                }
                for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
                    double aki = colk[i];
                    for (int j = k + 1; j <= n; ++j) a[(size_t)j * m + i] = fma(-wj[j], aki, a[(size_t)j * m + i]);
                }
            } else {
                z[k] = 0.0;
            }
            ctx.sync();
        }
        double s[n], suf[n];
        const double* V = ws.vout;
#if defined(__CUDACC__)
        if constexpr (Ctx::kLanes == 32) {
            jacobi_svd6_warp(ctx.lane, R, z, s, suf, ws.vout, false);
        } else
#endif
        {
            jacobi_svd6(ctx, R, z, ws.w, ws.v, s, suf, ws.vout);
        }

        double theta = fmax(0.995, 1.0 - g_norm);
        double gain = -1.0;
        double x_norm = norm6(x);
        while (gain <= 0.0 && nfev < max_nfev) {
            double p_h[n], step[n], step_h[n];
            more_step(m, s, suf, V, radius, &alpha, p_h);
            double predicted = choose_step(x, R, g_h, p_h, d, radius, lb, ub, theta, step, step_h);
            for (int i = 0; i < n; ++i) x_new[i] = x[i] + step[i];
            push_inside(x_new, lb, ub, 0.0);
            bool ok = residual_sweep(ctx, m, y, x_new, ws.unit, fcol, &cost_new);
            ctx.sync();
            ++nfev;
            double step_h_norm = norm6(step_h);
            if (!ok) { radius = 0.25 * step_h_norm; continue; }
            gain = cost - cost_new;
            double ratio;
            if (predicted > 0.0) ratio = gain / predicted;
            else if (predicted == 0.0 && gain == 0.0) ratio = 1.0;
            else ratio = 0.0;
            double radius_new = radius;
            if (ratio < 0.25) radius_new = 0.25 * step_h_norm;
            else if (ratio > 0.75 && step_h_norm > 0.95 * radius) radius_new = radius * 2.0;
            double step_norm = norm6(step);
            bool f_ok = (gain < 1e-12 * cost) && (ratio > 0.25);
            bool x_ok = step_norm < 1e-8 * (1e-8 + x_norm);
            if (f_ok && x_ok) status = kFitBoth;
            else if (f_ok) status = kFitFtol;
            else if (x_ok) status = kFitXtol;
            if (status != -100) break;
            alpha *= radius / radius_new;
            radius = radius_new;
        }
        if (gain > 0.0) {
            for (int i = 0; i < n; ++i) x[i] = x_new[i];
            cost = cost_new;
            // fcol holds f(x_new) and ws.unit its pow cache: both valid for the new base point
            fd_jacobian(ctx, m, y, x, lb, ub, ws.unit, a);
            ++njev;
            double part[n] = {0, 0, 0, 0, 0, 0};
            for (int i = ctx.lane; i < m; i += Ctx::kLanes) {
                double fi = fcol[i];
                for (int k = 0; k < n; ++k) part[k] += a[(size_t)k * m + i] * fi;
            }
            for (int k = 0; k < n; ++k) g[k] = ctx.sum(part[k]);
        }
        // a rejected final trial leaves fcol/unit at the trial point, but then status != None or
        // nfev == max_nfev and the loop ends at the next check without using them.
    }
    if (status == -100) status = kFitMaxNfev;
    for (int i = 0; i < n; ++i) res->x[i] = x[i];
    res->cost = cost;
    res->nfev = nfev;
    res->njev = njev;
    res->status = status;
}

}  // namespace motes
