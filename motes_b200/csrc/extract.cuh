// Horne-style optimal + aperture extraction of one wavelength column by one lane context.
//
// Replaces the body of the per-column loop of extraction.optimal_extraction
// (/root/reference/src/motes/extraction.py:101-206).  The column has already been
// scrubbed (extraction.py:31-34: errs = |errs|, non-finite errs -> 0, non-finite data -> 0)
// and sits in fast memory with a caller-chosen stride.
#pragma once

#include "common.cuh"

namespace motes {

struct ColumnSpectrum {
    double optimal, optimal_err, aperture, aperture_err;
    int lo, hi;          // int(floor(lim_lo)), int(ceil(lim_hi)) + 1 as the reference computes them
};

// Median of the entries of col[0..n) (stride `st`) that are != 0, by rank counting.
// Only used on the rare columns that contain zero errors (extraction.py:139).
template <class Ctx>
MOTES_HD double median_of_nonzero(const Ctx& ctx, const double* col, int st, int n, int count) {
    int k1 = (count - 1) / 2, k2 = count / 2;
    double v1 = 0.0, v2 = 0.0;
    for (int i = ctx.lane; i < n; i += Ctx::kLanes) {
        double c = col[(size_t)i * st];
        if (c == 0.0) continue;
        int less = 0, equal = 0;
        for (int j = 0; j < n; ++j) {
            double o = col[(size_t)j * st];
            if (o == 0.0) continue;
            less += (o < c);
            equal += (o == c);
        }
        if (less <= k1 && k1 < less + equal) v1 = c;
        if (less <= k2 && k2 < less + equal) v2 = c;
    }
    // exactly one distinct value satisfies each condition; every lane holding it holds the same bits
    v1 = ctx.max(fabs(v1));      // errs are non-negative after the scrub
    v2 = ctx.max(fabs(v2));
    return (v1 + v2) * 0.5;
}

// data / errs: scrubbed column, element i at [i*st].  errs is modified in place when zeros
// are replaced by the median of the non-zero errors (as the reference does on its copy).
// keep (optional): nspat bytes, set to 1 for rows inside the aperture that survive the mask.
template <class Ctx>
MOTES_HD void extract_column(const Ctx& ctx, const double* data, double* errs, int st, int nspat,
                             double lim_lo, double lim_hi, const double* bin /* 8-vector */,
                             ColumnSpectrum* out, uint8_t* keep) {
    out->optimal = out->optimal_err = out->aperture = out->aperture_err = 0.0;
    int lo = (int)floor(lim_lo);
    int hi = (int)ceil(lim_hi) + 1;
    out->lo = lo;
    out->hi = hi;
    if (keep)
        for (int i = ctx.lane; i < nspat; i += Ctx::kLanes) keep[i] = 0;

    // extraction.py:133-139
    double nzc = 0.0;
    for (int i = ctx.lane; i < nspat; i += Ctx::kLanes) nzc += (errs[(size_t)i * st] != 0.0) ? 1.0 : 0.0;
    int nonzero = (int)ctx.sum(nzc);
    if (nonzero == 0) return;
    if (nonzero < nspat) {
        double med = median_of_nonzero(ctx, errs, st, nspat, nonzero);
        ctx.sync();
        for (int i = ctx.lane; i < nspat; i += Ctx::kLanes)
            if (errs[(size_t)i * st] == 0.0) errs[(size_t)i * st] = med;
        ctx.sync();
    }

    int start, stop;
    python_slice(lo, hi, nspat, &start, &stop);

    // aperture sums and the un-normalised profile (extraction.py:143-170)
    const double amp = bin[0], alpha = bin[2], beta = bin[3];
    const double centre = (lim_hi + lim_lo) * 0.5;
    double s_flux = 0.0, s_var = 0.0, s_prof = 0.0;
    int all_zero = 1, all_one = 1;
    for (int i = start + ctx.lane; i < stop; i += Ctx::kLanes) {
        double d = data[(size_t)i * st];
        double e = errs[(size_t)i * st];
        s_flux += d;
        s_var += e * e;
        s_prof += moffat_value(moffat_unit((double)i, centre, alpha, beta), (double)i, amp, 0.0, 0.0);
        all_zero &= (d == 0.0);
        all_one &= (d == 1.0);
    }
    double flux = ctx.sum(s_flux);
    double var_sum = ctx.sum(s_var);
    double prof_sum = ctx.sum(s_prof);
    all_zero = ctx.all(all_zero);
    all_one = ctx.all(all_one);
    out->aperture = flux;
    out->aperture_err = sqrt(var_sum);

    // 5-sigma mask (extraction.py:174-184) and the weighted sums (extraction.py:193-203)
    double num = 0.0, den = 0.0, psum = 0.0;       // masked
    double num_all = 0.0, den_all = 0.0, psum_all = 0.0;
    int any_kept = 0;
    for (int i = start + ctx.lane; i < stop; i += Ctx::kLanes) {
        double d = data[(size_t)i * st];
        double e = errs[(size_t)i * st];
        double var = e * e;
        double p = moffat_value(moffat_unit((double)i, centre, alpha, beta), (double)i, amp, 0.0, 0.0) / prof_sum;
        double dev = d - (flux * p);
        bool rejected = fabs(dev * dev) > var * 25.0;
        double w_num = (p * d) / var, w_den = (p * p) / var;
        num_all += w_num; den_all += w_den; psum_all += p;
        if (!rejected) { num += w_num; den += w_den; psum += p; any_kept = 1; }
        if (keep) keep[i] = rejected ? 0 : 1;
    }
    int none_kept = ctx.all(!any_kept);
    if (none_kept) {
        num = num_all; den = den_all; psum = psum_all;
        if (keep)
            for (int i = start + ctx.lane; i < stop; i += Ctx::kLanes) keep[i] = 1;
    }
    num = ctx.sum(num);
    den = ctx.sum(den);
    psum = ctx.sum(psum);
    double f_opt = 0.0, v_opt = 0.0;
    if (!(all_zero || all_one)) {
        f_opt = num / den;
        v_opt = psum / den;
    }
    out->optimal = f_opt;
    out->optimal_err = sqrt(v_opt);
}

}  // namespace motes
