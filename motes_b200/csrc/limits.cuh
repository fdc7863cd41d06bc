// Interpolation / end-extrapolation of per-bin extraction limits to every dispersion pixel
// (tracing.interpolate_extraction_lims and helpers, tracing.py:13-111, 315-400), including the
// two different linear-interpolation formulas scipy's interp1d dispatches to.
#pragma once

#include "common.cuh"
#include "select.cuh"

namespace motes {

// np.interp for xq inside [xp[0], xp[n-1]) (numpy compiled_base.c arr_interp)
MOTES_HD double interp_numpy(double xq, const double* xp, const double* fp, int n) {
    int lo = 0, hi = n - 1;                  // find j with xp[j] <= xq < xp[j+1]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (xp[mid] <= xq) lo = mid; else hi = mid;
    }
    int j = lo;
    if (xq >= xp[n - 1]) return fp[n - 1];
    if (xp[j] == xq) return fp[j];
    double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    double res = slope * (xq - xp[j]) + fp[j];
    if (res != res) {
        res = slope * (xq - xp[j + 1]) + fp[j + 1];
        if (res != res && fp[j] == fp[j + 1]) res = fp[j];
    }
    return res;
}

// Median of values[a:b] (python slice already resolved to [a, b)); np.median: NaN propagates, empty -> NaN.
template <class Scope>
MOTES_HD double window_median(const Scope& sc, const double* values, int a, int b, unsigned long long* keys,
                              uint32_t* hist, unsigned long long* cell, int* icell) {
    int n = b - a;
    int has_nan = 0;
    for (int i = sc.tid; i < n; i += sc.size()) {
        double v = values[a + i];
        has_nan |= (v != v);
        keys[i] = key_of(v);
    }
    sc.sync();
    has_nan = sc.any(has_nan);
    double med = median_of_keys(sc, keys, n, hist, cell);
    (void)icell;
    return has_nan ? NAN : med;
}

// table: 4 rows (bin centre, lower, upper, Moffat centre) of nb entries, row stride `ts`.
// inner: scratch for 2 * (ndisp + 2) doubles.  keys: scratch for 150 keys.
template <class Scope>
MOTES_HD void interpolate_limits(const Scope& sc, const double* table, int ts, int nb, int ndisp, double* out_lo,
                                 double* out_hi, double* inner, unsigned long long* keys, uint32_t* hist,
                                 unsigned long long* cell, int* icell) {
    const double* centres = table;
    if (nb == 1) {
        for (int p = sc.tid; p < ndisp; p += sc.size()) {
            out_lo[p] = table[ts];
            out_hi[p] = table[2 * ts];
        }
        sc.sync();
        return;
    }
    const double first = centres[0], last = centres[nb - 1];
    int n_inner = (int)ceil(last - first);           // len(np.arange(first, last))
    if (n_inner < 0) n_inner = 0;
    const int L = n_inner + 2;
    for (int k = 0; k < 2; ++k) {
        double* cap = inner + (size_t)k * (ndisp + 2);
        const double* vals = table + (size_t)(k + 1) * ts;
        // zeros + interp (tracing.py:395-398)
        for (int i = sc.tid; i < n_inner; i += sc.size())
            cap[1 + i] = 0.0 + interp_numpy(first + (double)i, centres, vals, nb);
        sc.sync();
        const double* mid = cap + 1;
        int a, b;
        python_slice(0, 150, n_inner, &a, &b);
        double ya = window_median(sc, mid, a, b, keys, hist, cell, icell);
        python_slice(150, 300, n_inner, &a, &b);
        double yb = window_median(sc, mid, a, b, keys, hist, cell, icell);
        double g_short = (ya - yb) / (75.0 - 225.0);
        python_slice(-300, -150, n_inner, &a, &b);
        ya = window_median(sc, mid, a, b, keys, hist, cell, icell);
        python_slice(-150, -1, n_inner, &a, &b);
        yb = window_median(sc, mid, a, b, keys, hist, cell, icell);
        double g_long = (ya - yb) / (-225.0 - (-75.5));
        if (sc.tid == 0) {
            cap[0] = mid[0] - (g_short * first);
            cap[L - 1] = mid[n_inner - 1] + (g_long * ((double)ndisp - last));
        }
        sc.sync();
    }
    // second interpolation: interp1d(..., fill_value="extrapolate") two-weight form on
    // x = [0, first, first+1, ..., ndisp]
    for (int p = sc.tid; p < ndisp; p += sc.size()) {
        double xq = (double)p;
        // searchsorted(x, xq, 'left'): first index with x[i] >= xq
        int lo = 0, hi = L;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            double xm = (mid == 0) ? 0.0 : (mid == L - 1 ? (double)ndisp : first + (double)(mid - 1));
            if (xm < xq) lo = mid + 1; else hi = mid;
        }
        int idx = lo < 1 ? 1 : (lo > L - 1 ? L - 1 : lo);
        double x0 = (idx - 1 == 0) ? 0.0 : first + (double)(idx - 2);
        double x1 = (idx == L - 1) ? (double)ndisp : first + (double)(idx - 1);
        for (int k = 0; k < 2; ++k) {
            const double* cap = inner + (size_t)k * (ndisp + 2);
            double val = ((xq - x0) / (x1 - x0)) * cap[idx] + ((x1 - xq) / (x1 - x0)) * cap[idx - 1];
            (k == 0 ? out_lo : out_hi)[p] = 0.0 + val;
        }
    }
    sc.sync();
}

}  // namespace motes
