// Sky localisation, bootstrap sky model and subtraction for one wavelength column
// (sky.subtract_sky / sky.sky_median, /root/reference/src/motes/sky.py:17-43, 257-399).
//
// The reference walks the columns in order and, per column, resamples the good sky pixels 100
// times from numpy's legacy global MT19937 (np.random.choice) and takes mean / std of the 100
// medians.  Here the work is split in three:
//   1. sky_prepare_column   (parallel over columns) clamp limits, pick sky pixels, 10-sigma clip,
//                           rank the surviving pixels;
//   2. sky_bootstrap_frame  (one thread scope per frame, columns in order) consumes the MT19937
//                           stream exactly as numpy does (masked rejection, row-major fill) and
//                           turns every bootstrap sample into a rank histogram from which its
//                           median is read off - no resampled array is ever materialised;
//   3. sky_apply_column     (parallel) subtract, propagate the error, re-mark chip gaps.
#pragma once

#include "common.cuh"
#include "select.cuh"
#include "mt19937.cuh"

namespace motes {

enum SkyFlag : int { kSkyModelled = 0, kSkySkipped = 1, kSkyNoPixels = 2 };

constexpr double kSqrt99 = 9.9498743710662;     // 99 ** 0.5 (sky.py:41), 0x1.3e655eefe1367p+3

// np.add.reduce on a contiguous float64 vector (numpy pairwise_sum_DOUBLE): 8 running
// accumulators for blocks of <= 128, recursive halving above.
MOTES_HD double numpy_sum(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return numpy_sum(a, n2) + numpy_sum(a + n2, n - n2);
}

// ---- 1. per-column preparation --------------------------------------------------------------
// col: the column (element i at col[i*st]).  lim_lo / lim_hi: this column's limits, clamped in
// place as sky.py:303-307 does.  keys: nspat 64-bit scratch.  Outputs: sorted[0..n_good) ascending
// good sky values, rank_of[g] = rank of the g-th good pixel (row order), good_row[g] = its row.
template <class Scope>
MOTES_HD void sky_prepare_column(const Scope& sc, const double* col, int st, int nspat, double* lim_lo,
                                 double* lim_hi, int spatial_floor, int* n_sky_out, int* n_good_out,
                                 int* flag_out, double* sorted, uint16_t* rank_of, uint16_t* good_row,
                                 unsigned long long* keys, uint32_t* hist, unsigned long long* cell, int* icell) {
    double lo = *lim_lo, hi = *lim_hi;
    const double top = (double)(nspat - 1) + (double)spatial_floor;      // (spatial_axis + floor)[-1]
    if (lo < 0.0) lo = 0.0;
    if (hi > top) hi = top;
    sc.sync();
    if (sc.tid == 0) { *lim_lo = lo; *lim_hi = hi; }
    const double edge_lo = lo + (double)spatial_floor, edge_hi = hi + (double)spatial_floor;

    // sky pixels, in row order; keys[] receives the non-NaN ones for the median
    int my_sky = 0, my_valid = 0, my_nan = 0;
    double my_min = INFINITY, my_max = -INFINITY;
    for (int i = sc.tid; i < nspat; i += sc.size()) {
        double axis = (double)(i + spatial_floor);
        if (axis < edge_lo || axis > edge_hi) {
            double v = col[(size_t)i * st];
            ++my_sky;
            if (v != v) ++my_nan;
            else { ++my_valid; my_min = fmin(my_min, v); my_max = fmax(my_max, v); }
        }
    }
    int n_sky = sc.isum(my_sky, icell);
    int n_nan = sc.isum(my_nan, icell);
    int n_valid = n_sky - n_nan;
    *n_sky_out = n_sky;
    *n_good_out = 0;
    if (n_sky == 0) { *flag_out = kSkyNoPixels; return; }
    unsigned long long kmin = sc.umin(key_of(my_min), cell);
    unsigned long long kmax = ~sc.umin(~key_of(my_max), cell);
    // len(set(sky_pixels)) == 1 (sky.py:338): one pixel, or all pixels the same number
    // (every NaN is its own set element, so NaNs never collapse)
    bool constant = (n_sky == 1) || (n_nan == 0 && value_of(kmin) == value_of(kmax));
    if (constant) { *flag_out = kSkySkipped; return; }
    *flag_out = kSkyModelled;

    // compact the valid sky values into keys[] (order irrelevant for the median) -- serial prefix
    // per thread stripe is avoided by a second counting pass: thread-strided compaction via scan
    int base = 0;
    for (int start = 0; start < nspat; start += sc.size()) {
        int i = start + sc.tid;
        int take = 0;
        double v = 0.0;
        if (i < nspat) {
            double axis = (double)(i + spatial_floor);
            if (axis < edge_lo || axis > edge_hi) { v = col[(size_t)i * st]; take = (v == v); }
        }
        int total;
        int off = sc.scan_flags(take, &total, hist);
        if (take) keys[base + off] = key_of(v);
        base += total;
    }
    sc.sync();
    double med = median_of_keys(sc, keys, n_valid, hist, cell);       // np.nanmedian
    // np.nanstd (population): mean over valid, then sqrt(mean((x-mean)^2))
    double part = 0.0;
    for (int i = sc.tid; i < n_valid; i += sc.size()) part += value_of(keys[i]);
    double mean = sc.dsum(part, (double*)cell) / (double)n_valid;
    part = 0.0;
    for (int i = sc.tid; i < n_valid; i += sc.size()) { double d = value_of(keys[i]) - mean; part += d * d; }
    double sigma = sqrt(sc.dsum(part, (double*)cell) / (double)n_valid);
    const double cut_lo = med - (10.0 * sigma), cut_hi = med + (10.0 * sigma);

    // good sky pixels in row order -> keys[] now holds their values' keys, good_row their rows
    sc.sync();
    base = 0;
    for (int start = 0; start < nspat; start += sc.size()) {
        int i = start + sc.tid;
        int take = 0;
        double v = 0.0;
        if (i < nspat) {
            double axis = (double)(i + spatial_floor);
            if (axis < edge_lo || axis > edge_hi) { v = col[(size_t)i * st]; take = (v > cut_lo && v < cut_hi); }
        }
        int total;
        int off = sc.scan_flags(take, &total, hist);
        if (take) { keys[base + off] = key_of(v); good_row[base + off] = (uint16_t)i; }
        base += total;
    }
    sc.sync();
    const int n_good = base;
    *n_good_out = n_good;
    // rank by counting (ties broken by position): O(n^2 / threads), n <= nspat
    for (int g = sc.tid; g < n_good; g += sc.size()) {
        unsigned long long k = keys[g];
        int rank = 0;
        for (int j = 0; j < n_good; ++j) {
            unsigned long long o = keys[j];
            rank += (o < k) || (o == k && j < g);
        }
        rank_of[g] = (uint16_t)rank;
        sorted[rank] = value_of(k);
    }
    sc.sync();
}

// ---- 2. bootstrap over the columns of one frame ---------------------------------------------
// mt / pos: the generator (624 shared words + position, numpy's RandomState layout).
// Per column c (stride `cs` between columns in sorted / rank_of):
//   flag != modelled -> nothing drawn; n_good == 0 -> NaN; n_good == 1 -> no draws (numpy's
//   masked rejection with range 0 consumes none).
// hist: 50 * max_good shared words (100 16-bit counters per rank); ranks: max_good shared uint16;
// meds: 100 shared doubles; scan: 33 shared words; ctl: 2 shared ints.
template <class Scope>
MOTES_HD void sky_bootstrap_frame(const Scope& sc, uint32_t* mt, int* pos_io, int ncol, const int* n_good,
                                  const int* flag, const double* sorted, const uint16_t* rank_of, size_t cs,
                                  double* sky, double* sky_err, uint32_t* hist, uint16_t* ranks, double* meds,
                                  uint32_t* scan, int* ctl) {
    int pos = *pos_io;
    for (int c = 0; c < ncol; ++c) {
        if (flag[c] != kSkyModelled) {
            if (sc.tid == 0) { sky[c] = 0.0; sky_err[c] = 0.0; }
            continue;
        }
        const int n = n_good[c];
        const double* srt = sorted + (size_t)c * cs;
        if (n == 0) {
            if (sc.tid == 0) { sky[c] = NAN; sky_err[c] = NAN; }
            continue;
        }
        if (n > 1) {
            for (int i = sc.tid; i < 50 * n; i += sc.size()) hist[i] = 0;
            for (int i = sc.tid; i < n; i += sc.size()) ranks[i] = rank_of[(size_t)c * cs + i];
            sc.sync();
            const int need = n * kBootstrap;
            const uint32_t mask = rejection_mask((uint32_t)n);
            int have = 0;
            while (have < need) {
                if (pos >= kMtN) { mt_twist(sc, mt); pos = 0; }
                int i = pos + sc.tid;
                int acc = 0;
                uint32_t v = 0;
                if (i < kMtN) {
                    v = mt_temper(mt[i]) & mask;
                    acc = (v <= (uint32_t)(n - 1));
                }
                int total;
                int off = sc.scan_flags(acc, &total, scan);
                int ord = have + off;
                if (acc && ord < need) {
                    int s = ord % kBootstrap;                       // row-major (pixel, sample) fill
                    sc.add(&hist[(int)ranks[v] * 50 + (s >> 1)], 1u << (16 * (s & 1)));
                }
                if (have + total >= need) {
                    if (acc && ord == need - 1) ctl[0] = i + 1;      // first unconsumed output
                    sc.sync();
                    pos = ctl[0];
                    have = need;
                    sc.sync();
                } else {
                    have += total;
                    pos = (pos + sc.size() > kMtN) ? kMtN : pos + sc.size();
                }
            }
            // median of every bootstrap sample from its rank histogram
            const int k1 = (n - 1) / 2, k2 = n / 2;
            for (int s = sc.tid; s < kBootstrap; s += sc.size()) {
                int cum = 0, r1 = -1, r2 = -1;
                for (int r = 0; r < n; ++r) {
                    cum += (int)((hist[r * 50 + (s >> 1)] >> (16 * (s & 1))) & 0xffffu);
                    if (r1 < 0 && cum > k1) r1 = r;
                    if (cum > k2) { r2 = r; break; }
                }
                meds[s] = (k1 == k2) ? srt[r1] : (srt[r1] + srt[r2]) * 0.5;
            }
        } else {
            for (int s = sc.tid; s < kBootstrap; s += sc.size()) meds[s] = srt[0];
        }
        sc.sync();
        if (sc.tid == 0) {
            // np.nanmean / np.std of the 100 medians in numpy's summation order
            double mean = numpy_sum(meds, kBootstrap) / (double)kBootstrap;
            double sq[kBootstrap];
            for (int s = 0; s < kBootstrap; ++s) { double d = meds[s] - mean; sq[s] = d * d; }
            double var = numpy_sum(sq, kBootstrap) / (double)kBootstrap;
            sky[c] = mean;
            sky_err[c] = sqrt(var) / kSqrt99;
        }
        sc.sync();
    }
    if (sc.tid == 0) *pos_io = pos;
    sc.sync();
}

// ---- 3. subtraction, error propagation and chip-gap re-marking (one thread per column) ------
// sky.py:375-378 and :396-397.
MOTES_HD void sky_apply_column(double* data, double* errs, int nspat, int ndisp, int c, int flag, double sky,
                               double sky_err) {
    int lt = 0, eq = 0, n = 0;
    double max_lt = -INFINITY, min_gt = INFINITY;
    const double e2 = sky_err * sky_err;
    for (int r = 0; r < nspat; ++r) {
        size_t g = (size_t)r * ndisp + c;
        double d = data[g];
        if (flag == kSkyModelled) {
            d = d - sky;
            data[g] = d;
            double e = errs[g];
            errs[g] = sqrt((e * e) + e2);
        }
        if (d == d) {
            ++n;
            if (d < 0.0) { ++lt; if (d > max_lt) max_lt = d; }
            else if (d == 0.0) ++eq;
            else if (d < min_gt) min_gt = d;
        }
    }
    if (median_equals(0.0, n, lt, eq, max_lt, min_gt))       // np.nanmedian(column) == 0
        for (int r = 0; r < nspat; ++r) data[(size_t)r * ndisp + c] += 1.0;
}

}  // namespace motes
