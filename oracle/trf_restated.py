"""Restatement of scipy's bounded Trust-Region-Reflective least-squares solver.  TEST INFRASTRUCTURE ONLY.

The reference's Moffat fit is ``scipy.optimize.least_squares(method="trf",
bounds=..., ftol=1e-12)`` with every other argument at its default
(``/root/reference/src/motes/tracing.py:591-598``).  scipy is a third-party
dependency of the reference (``/root/reference/poetry.lock:701-702`` pins 1.14.1;
this image carries 1.18.1, whose sources were read for this restatement).  The
algorithm restated here is the specialisation that call selects:

    dense Jacobian by 2-point forward differences, x_scale = 1, linear loss,
    tr_solver = "exact" (SVD of the augmented Jacobian), xtol = gtol = 1e-8,
    max_nfev = 100 * n.

Sources followed (paths under ``site-packages/scipy/optimize``):
``_lsq/least_squares.py:896-946`` (set-up), ``_lsq/trf.py:129-203`` (step
selection), ``_lsq/trf.py:206-410`` (main loop), ``_lsq/common.py:18-54, 57-168,
222-246, 251-361, 372-398, 440-508, 705-717`` and ``_numdiff.py:14-91, 147-203,
683-713`` (finite differences).

``tests/test_trf_restated.py`` pins it: on seeded Moffat profiles it must return
the same ``x``, ``nfev``, ``njev`` and ``status`` as scipy, bit for bit.  Its
purpose is to document, in plain arithmetic, exactly what the CUDA solver in
``motes_b200/csrc/trf.cuh`` has to do, and to provide iteration traces for it.
"""

from __future__ import annotations

from math import copysign
from types import SimpleNamespace

import numpy as np
from numpy.linalg import norm
from scipy.linalg import svd

EPS = np.finfo(float).eps
FD_REL_STEP = EPS ** 0.5


# ------------------------------------------------------------------ bounds helpers
def _inside(x, lb, ub):
    return bool(np.all((x >= lb) & (x <= ub)))


def _active_set(x, lb, ub, rtol):
    """-1 / 0 / +1 per variable (``_lsq/common.py`` find_active_constraints)."""
    active = np.zeros(x.size, dtype=int)
    if rtol == 0:
        active[x <= lb] = -1
        active[x >= ub] = 1
        return active
    below = x - lb
    above = ub - x
    tol_lo = rtol * np.maximum(1, np.abs(lb))
    tol_hi = rtol * np.maximum(1, np.abs(ub))
    active[np.isfinite(lb) & (below <= np.minimum(above, tol_lo))] = -1
    active[np.isfinite(ub) & (above <= np.minimum(below, tol_hi))] = 1
    return active


def _push_inside(x, lb, ub, rstep):
    """``make_strictly_feasible``: rstep > 0 moves by rstep*max(1,|bound|); rstep == 0 uses nextafter."""
    out = x.copy()
    active = _active_set(x, lb, ub, rstep)
    lo, hi = active == -1, active == 1
    if rstep == 0:
        out[lo] = np.nextafter(lb[lo], ub[lo])
        out[hi] = np.nextafter(ub[hi], lb[hi])
    else:
        out[lo] = lb[lo] + rstep * np.maximum(1, np.abs(lb[lo]))
        out[hi] = ub[hi] - rstep * np.maximum(1, np.abs(ub[hi]))
    squeezed = (out < lb) | (out > ub)
    out[squeezed] = 0.5 * (lb[squeezed] + ub[squeezed])
    return out


def _coleman_li(x, g, lb, ub):
    v = np.ones_like(x)
    dv = np.zeros_like(x)
    up = (g < 0) & np.isfinite(ub)
    v[up] = ub[up] - x[up]
    dv[up] = -1
    down = (g > 0) & np.isfinite(lb)
    v[down] = x[down] - lb[down]
    dv[down] = 1
    return v, dv


def _stride_to_bound(x, s, lb, ub):
    """Smallest t >= 0 with x + t s on a bound, and which variables hit (sign of s)."""
    steps = np.full(x.size, np.inf)
    nz = np.nonzero(s)
    with np.errstate(over="ignore", invalid="ignore"):
        steps[nz] = np.maximum((lb - x)[nz] / s[nz], (ub - x)[nz] / s[nz])
    t = np.min(steps)
    return t, np.equal(steps, t) * np.sign(s).astype(int)


def _ball_crossing(x, s, radius):
    """Roots of ||x + t s|| = radius (``intersect_trust_region``)."""
    a = np.dot(s, s)
    if a == 0:
        raise ValueError("`s` is zero.")
    b = np.dot(x, s)
    c = np.dot(x, x) - radius ** 2
    if c > 0:
        raise ValueError("`x` is not within the trust region.")
    d = np.sqrt(b * b - a * c)
    q = -(b + copysign(d, b))
    t1, t2 = q / a, c / q
    return (t1, t2) if t1 < t2 else (t2, t1)


# ------------------------------------------------------------------ quadratic model
def _line_quadratic(J, g, s, diag, s0=None):
    v = J.dot(s)
    a = np.dot(v, v)
    a += np.dot(s * diag, s)
    a *= 0.5
    b = np.dot(g, s)
    if s0 is None:
        return a, b
    u = J.dot(s0)
    b += np.dot(u, v)
    c = 0.5 * np.dot(u, u) + np.dot(g, s0)
    b += np.dot(s0 * diag, s)
    c += 0.5 * np.dot(s0 * diag, s0)
    return a, b, c


def _line_minimum(a, b, lo, hi, c=0):
    t = [lo, hi]
    if a != 0:
        vertex = -0.5 * b / a
        if lo < vertex < hi:
            t.append(vertex)
    t = np.asarray(t)
    y = t * (a * t + b) + c
    k = np.argmin(y)
    return t[k], y[k]


def _model_value(J, g, s, diag):
    Js = J.dot(s)
    q = np.dot(Js, Js)
    q += np.dot(s * diag, s)
    return 0.5 * q + np.dot(s, g)


# ------------------------------------------------------------------ trust-region subproblem
def _more_step(n, m, uf, s, V, radius, alpha0):
    """Moré's root find on the SVD (``solve_lsq_trust_region``).  Returns p, alpha, n_iter."""
    def phi(alpha):
        den = s ** 2 + alpha
        pn = norm(suf / den)
        return pn - radius, -np.sum(suf ** 2 / den ** 3) / pn

    suf = s * uf
    full_rank = (s[-1] > EPS * m * s[0]) if m >= n else False
    if full_rank:
        p = -V.dot(uf / s)
        if norm(p) <= radius:
            return p, 0.0, 0
    upper = norm(suf) / radius
    if full_rank:
        f0, df0 = phi(0.0)
        lower = -f0 / df0
    else:
        lower = 0.0
    if alpha0 is None or (not full_rank and alpha0 == 0):
        alpha = max(0.001 * upper, (lower * upper) ** 0.5)
    else:
        alpha = alpha0
    it = 0
    for it in range(10):
        if alpha < lower or alpha > upper:
            alpha = max(0.001 * upper, (lower * upper) ** 0.5)
        f, df = phi(alpha)
        if f < 0:
            upper = alpha
        ratio = f / df
        lower = max(lower, alpha - ratio)
        alpha -= (f + radius) * ratio / radius
        if np.abs(f) < 0.01 * radius:
            break
    p = -V.dot(suf / (s ** 2 + alpha))
    p *= radius / norm(p)
    return p, alpha, it + 1


def _choose_step(x, J_h, diag_h, g_h, p, p_h, d, radius, lb, ub, theta):
    """``select_step``: plain, reflected or anti-gradient step, whichever models best."""
    if _inside(x + p, lb, ub):
        return p, p_h, -_model_value(J_h, g_h, p_h, diag_h)

    p_stride, hits = _stride_to_bound(x, p, lb, ub)
    r_h = np.copy(p_h)
    r_h[hits.astype(bool)] *= -1
    r = d * r_h

    p = p * p_stride
    p_h = p_h * p_stride
    x_on_bound = x + p

    _, to_tr = _ball_crossing(p_h, r_h, radius)
    to_bound, _ = _stride_to_bound(x_on_bound, r, lb, ub)
    r_stride = min(to_bound, to_tr)
    if r_stride > 0:
        r_lo = (1 - theta) * p_stride / r_stride
        r_hi = theta * to_bound if r_stride == to_bound else to_tr
    else:
        r_lo, r_hi = 0, -1

    if r_lo <= r_hi:
        a, b, c = _line_quadratic(J_h, g_h, r_h, diag_h, s0=p_h)
        r_stride, r_value = _line_minimum(a, b, r_lo, r_hi, c=c)
        r_h = r_h * r_stride
        r_h = r_h + p_h
        r = r_h * d
    else:
        r_value = np.inf

    p = p * theta
    p_h = p_h * theta
    p_value = _model_value(J_h, g_h, p_h, diag_h)

    ag_h = -g_h
    ag = d * ag_h
    to_tr = radius / norm(ag_h)
    to_bound, _ = _stride_to_bound(x, ag, lb, ub)
    ag_stride = theta * to_bound if to_bound < to_tr else to_tr
    a, b = _line_quadratic(J_h, g_h, ag_h, diag_h)
    ag_stride, ag_value = _line_minimum(a, b, 0, ag_stride)
    ag_h = ag_h * ag_stride
    ag = ag * ag_stride

    if p_value < r_value and p_value < ag_value:
        return p, p_h, -p_value
    if r_value < p_value and r_value < ag_value:
        return r, r_h, -r_value
    return ag, ag_h, -ag_value


# ------------------------------------------------------------------ finite differences
def fd_steps(x, lb, ub):
    """Forward-difference steps with the bound adjustment of ``_adjust_scheme_to_bounds`` ('1-sided')."""
    sign = (x >= 0).astype(float) * 2 - 1
    h = FD_REL_STEP * sign * np.maximum(1.0, np.abs(x))
    below = x - lb
    above = ub - x
    trial = x + h
    violated = (trial < lb) | (trial > ub)
    fitting = np.abs(h) <= np.maximum(below, above)
    h = h.copy()
    h[violated & fitting] *= -1
    forward = (above >= below) & ~fitting
    h[forward] = above[forward]
    backward = (above < below) & ~fitting
    h[backward] = -below[backward]
    return h


def fd_jacobian(fun, x, f0, lb, ub):
    h = fd_steps(x, lb, ub)
    Jt = np.empty((x.size, f0.size))           # scipy assembles J transposed and returns the view .T
    for i in range(x.size):
        x1 = x.copy()
        x1[i] = x[i] + h[i]
        dx = (x[i] + h[i]) - x[i]
        Jt[i] = (fun(x1) - f0) / dx
    return Jt.T


# ------------------------------------------------------------------ driver
def trf_bounded(residual, x0, lower, upper, args=(), ftol=1e-8, xtol=1e-8, gtol=1e-8,
                max_nfev=None, trace=None):
    """Same contract as ``least_squares(residual, x0, bounds=(lower, upper), args=args, method='trf', ftol=ftol)``.

    Returns a namespace with ``x, cost, nfev, njev, status`` (status codes as scipy:
    0 max_nfev, 1 gtol, 2 ftol, 3 xtol, 4 ftol and xtol).  ``trace`` (a list)
    receives one dict per outer iteration.
    """
    x0 = np.atleast_1d(x0).astype(float)
    lb = np.asarray(lower, dtype=float)
    ub = np.asarray(upper, dtype=float)
    if np.any(lb >= ub):
        raise ValueError("Each lower bound must be strictly less than each upper bound.")
    if not _inside(x0, lb, ub):
        raise ValueError("Initial guess is outside of provided bounds")
    x0 = _push_inside(x0, lb, ub, 1e-10)

    def fun(x):
        return np.atleast_1d(residual(x, *args))

    f = fun(x0)
    if not np.all(np.isfinite(f)):
        raise ValueError("Residuals are not finite in the initial point.")
    J = fd_jacobian(fun, x0, f, lb, ub)
    m, n = J.shape
    if max_nfev is None:
        max_nfev = 100 * n

    x = x0.copy()
    nfev = njev = 1
    cost = 0.5 * np.dot(f, f)
    g = J.T.dot(f)

    v, dv = _coleman_li(x, g, lb, ub)
    radius = norm(x0 / v ** 0.5)
    if radius == 0:
        radius = 1.0
    g_norm = norm(g * v, ord=np.inf)
    alpha = 0.0
    status = None
    f_aug = np.zeros(m + n)
    J_aug = np.empty((m + n, n))

    while True:
        v, dv = _coleman_li(x, g, lb, ub)
        g_norm = norm(g * v, ord=np.inf)
        if g_norm < gtol:
            status = 1
        if status is not None or nfev == max_nfev:
            break

        d = v ** 0.5
        diag_h = g * dv
        g_h = d * g
        f_aug[:m] = f
        J_aug[:m] = J * d
        J_h = J_aug[:m]
        J_aug[m:] = np.diag(diag_h ** 0.5)
        U, s, Vt = svd(J_aug, full_matrices=False)
        V = Vt.T
        uf = U.T.dot(f_aug)
        theta = max(0.995, 1 - g_norm)

        gain = -1
        inner = []
        while gain <= 0 and nfev < max_nfev:
            p_h, alpha, n_it = _more_step(n, m, uf, s, V, radius, alpha)
            p = d * p_h
            step, step_h, predicted = _choose_step(x, J_h, diag_h, g_h, p, p_h, d, radius,
                                                   lb, ub, theta)
            x_new = _push_inside(x + step, lb, ub, 0)
            f_new = fun(x_new)
            nfev += 1
            step_h_norm = norm(step_h)
            if not np.all(np.isfinite(f_new)):
                radius = 0.25 * step_h_norm
                continue
            cost_new = 0.5 * np.dot(f_new, f_new)
            gain = cost - cost_new
            # update_tr_radius
            if predicted > 0:
                ratio = gain / predicted
            elif predicted == gain == 0:
                ratio = 1
            else:
                ratio = 0
            radius_new = radius
            if ratio < 0.25:
                radius_new = 0.25 * step_h_norm
            elif ratio > 0.75 and step_h_norm > 0.95 * radius:
                radius_new = radius * 2.0
            step_norm = norm(step)
            # check_termination
            f_ok = gain < ftol * cost and ratio > 0.25
            x_ok = step_norm < xtol * (xtol + norm(x))
            inner.append(dict(radius=radius, alpha=alpha, n_it=n_it, gain=gain,
                              predicted=predicted, ratio=ratio, step=step.copy()))
            if f_ok and x_ok:
                status = 4
            elif f_ok:
                status = 2
            elif x_ok:
                status = 3
            if status is not None:
                break
            alpha *= radius / radius_new
            radius = radius_new

        if trace is not None:
            trace.append(dict(x=x.copy(), cost=cost, g_norm=g_norm, s=s.copy(), inner=inner))

        if gain > 0:
            x = x_new
            f = f_new
            cost = cost_new
            J = fd_jacobian(fun, x, f, lb, ub)
            njev += 1
            g = J.T.dot(f)

    if status is None:
        status = 0
    return SimpleNamespace(x=x, cost=cost, nfev=nfev, njev=njev, status=status,
                           optimality=g_norm, fun=f)
