// S/N-driven binning of the dispersion axis (tracing.get_bins, tracing.py:114-254) and the
// per-column statistics it needs.
//
// The reference re-sums the whole (rows x width) window every time a bin grows by one column
// (O(width^2) work per bin).  Here a first pass reduces every column once to
//   S[c] = nansum(data[row_lo:row_hi, c]),  V[c] = nansum(errs[row_lo:row_hi, c]^2)
// (Python slice semantics for row_lo/row_hi) and the greedy scan then grows windows by adding
// one S/V pair per step.  The "every spatial row must hold more than min_cols good pixels"
// test (tracing.py:281-312) keeps one running counter per row.
#pragma once

#include "common.cuh"

namespace motes {

// Neumaier-compensated accumulator: the reference sums pairwise, we sum sequentially; both stay
// within an ulp or two of the exact sum, so threshold decisions (snr <= limit) agree unless the
// exact value sits on the threshold.
struct Accum {
    double s = 0.0, c = 0.0;
    MOTES_HD void add(double v) {
        double t = s + v;
        if (fabs(s) >= fabs(v)) c += (s - t) + v; else c += (v - t) + s;
        s = t;
    }
    MOTES_HD double value() const { return s + c; }
};

struct ColumnStats {
    double sum_data;     // S[c]
    double sum_var;      // V[c]
    uint8_t gap;         // np.median(data[:, c]) == 1  -> chip-gap column (tracing.py:506)
};

// One thread: statistics of column c of one frame (nspat x ndisp, dispersion fastest).
MOTES_HD ColumnStats column_stats(const double* data, const double* errs, int nspat, int ndisp, int c,
                                  int row_lo, int row_hi) {
    int start, stop;
    python_slice(row_lo, row_hi, nspat, &start, &stop);
    Accum sd, sv;
    int lt = 0, eq = 0, nan = 0;
    double max_lt = -INFINITY, min_gt = INFINITY;
    for (int r = 0; r < nspat; ++r) {
        double d = data[(size_t)r * ndisp + c];
        if (r >= start && r < stop) {
            double e = errs[(size_t)r * ndisp + c];
            double e2 = e * e;
            if (d == d) sd.add(d);
            if (e2 == e2) sv.add(e2);
        }
        if (d != d) ++nan;
        else if (d < 1.0) { ++lt; if (d > max_lt) max_lt = d; }
        else if (d == 1.0) ++eq;
        else if (d < min_gt) min_gt = d;
    }
    ColumnStats out;
    out.sum_data = sd.value();
    out.sum_var = sv.value();
    out.gap = (nan == 0 && median_equals(1.0, nspat, lt, eq, max_lt, min_gt)) ? 1 : 0;
    return out;
}

// Greedy centre-outward scan for one frame by one thread scope.
//   qual      nspat x ndisp, 1 = good, 0 = bad (uint8)
//   S, V      per-column sums from column_stats
//   allgood   optional, per column: 1 when every row of the column has qual == 1.  Such a column adds one to
//             every row's counter, which is kept as a common offset, so the scan touches the nspat counters
//             (and reads a qual column, a stride-ndisp gather) only for columns with a bad pixel
//   counts    scratch, nspat ints shared by the scope;  cell: one shared 64-bit word
//   bins_i    out, [2*maxbins] first/last column of each bin, sorted by first
//   bins_snr  out, [maxbins]
//   tmp_i / tmp_snr  scratch of the same sizes (blue-ward bins are produced in reverse order)
// Returns the number of bins.
struct BinWindow {        // staging area for a run of columns: S, V and the all-rows-good flag
    double* S;
    double* V;
    uint8_t* good;
    double* xfer;         // 10 doubles: scan state handed from the warp that runs ahead to the rest of the scope
    int cap;
};

template <class Scope>
MOTES_HD int snr_bins_scan(const Scope& sc, const uint8_t* qual, const double* S, const double* V, int nspat,
                           int ndisp, double snr_limit, double min_cols, int* counts, int32_t* bins_i,
                           double* bins_snr, int32_t* tmp_i, double* tmp_snr, const uint8_t* allgood = nullptr,
                           unsigned long long* cell = nullptr, const BinWindow* win = nullptr) {
    const int centre = ndisp / 2;
    int n_blue = 0, n_red = 0;
    // Per-column inputs through a staged window (shared memory on the device): the scan is one serial chain per
    // frame, and a global load per column would sit on it.  Every thread sees the same `col`, so the refill
    // (two scope barriers around a cooperative copy) is taken by all of them together.
    int win_lo = 0, win_n = 0;
    const double limit2 = (snr_limit * snr_limit) * (1.0 - 1e-12);      // see the S/N test below
    auto stage = [&](int col, int dir) {
        if (!win || (col >= win_lo && col < win_lo + win_n)) return;
        sc.sync();
        win_lo = (dir == 0) ? (col - win->cap + 1 > 0 ? col - win->cap + 1 : 0) : col;
        win_n = (ndisp - win_lo < win->cap) ? ndisp - win_lo : win->cap;
        for (int i = sc.tid; i < win_n; i += sc.size()) {
            win->S[i] = S[win_lo + i];
            win->V[i] = V[win_lo + i];
            win->good[i] = allgood ? allgood[win_lo + i] : (uint8_t)0;
        }
        sc.sync();
    };
    for (int dir = 0; dir < 2; ++dir) {
        int edge = centre, width = 0;
        while (dir == 0 ? (edge - width > 0) : (edge + width < ndisp)) {
            double snr = 0.0;
            bool pending = false;               // the last window's S/N was passed over (see below)
            double s_pending = 0.0, v_pending = 0.0;
            Accum sum_s, sum_v;
            for (int r = sc.tid; r < nspat; r += sc.size()) counts[r] = 0;
            int common = 0, least = 0;          // counter of row r = common + counts[r]; least = min over r of counts[r]
            sc.sync();
            while (snr <= snr_limit) {
                // Runs of all-good columns inside the staged window need no other thread: one warp takes them alone
                // (the arithmetic is a serial chain; a dozen warps repeating it only compete for issue slots) and hands
                // the state back through win->xfer.  It stops in front of the frame edge, the window's end and any
                // column with a bad pixel, which the code below handles.
                if (win) {
                    const int w1 = width + 1;
                    const int next = (dir == 0) ? edge - w1 : edge + w1 - 1;
                    const bool inside = (dir == 0) ? (edge - w1 >= 0) : (edge + w1 <= ndisp);
                    if (inside && next >= win_lo && next < win_lo + win_n && win->good[next - win_lo]) {
                        if (sc.tid < 32) {
                            while (true) {
                                const int w2 = width + 1;
                                if (dir == 0 ? (edge - w2 < 0) : (edge + w2 > ndisp)) break;
                                const int col = (dir == 0) ? edge - w2 : edge + w2 - 1;
                                if (col < win_lo || col >= win_lo + win_n || !win->good[col - win_lo]) break;
                                width = w2;
                                common += 1;
                                sum_s.add(win->S[col - win_lo]);
                                sum_v.add(win->V[col - win_lo]);
                                if ((double)(common + least) <= min_cols) continue;
                                const double s_now = sum_s.value(), v_now = sum_v.value();
                                if (snr_limit >= 0.0 && v_now > 0.0 && (s_now <= 0.0 || s_now * s_now < limit2 * v_now)) {
                                    pending = true;
                                    s_pending = s_now;
                                    v_pending = v_now;
                                    continue;
                                }
                                pending = false;
                                snr = s_now / sqrt(v_now);
                                if (!(snr <= snr_limit)) break;
                            }
                            if (sc.tid == 0) {
                                double* x = win->xfer;
                                x[0] = (double)width; x[1] = (double)common; x[2] = snr; x[3] = pending ? 1.0 : 0.0;
                                x[4] = s_pending; x[5] = v_pending; x[6] = sum_s.s; x[7] = sum_s.c; x[8] = sum_v.s; x[9] = sum_v.c;
                            }
                        }
                        sc.sync();
                        {
                            const double* x = win->xfer;
                            width = (int)x[0]; common = (int)x[1]; snr = x[2]; pending = x[3] != 0.0;
                            s_pending = x[4]; v_pending = x[5]; sum_s.s = x[6]; sum_s.c = x[7]; sum_v.s = x[8]; sum_v.c = x[9];
                        }
                        sc.sync();
                        continue;
                    }
                }
                width += 1;
                if (dir == 0 ? (edge - width < 0) : (edge + width > ndisp)) {
                    width = (dir == 0) ? edge : ndisp - edge;
                    break;
                }
                int col = (dir == 0) ? edge - width : edge + width - 1;
                stage(col, dir);
                const bool col_good = win ? (win->good[col - win_lo] != 0) : (allgood && allgood[col]);
                const double s_col = win ? win->S[col - win_lo] : S[col], v_col = win ? win->V[col - win_lo] : V[col];
                bool any_short;
                if (col_good) {
                    common += 1;
                    any_short = ((double)(common + least) <= min_cols);
                } else {
                    int is_short = 0, mine = 0x7fffffff;
                    for (int r = sc.tid; r < nspat; r += sc.size()) {
                        int cnt = counts[r] + (int)qual[(size_t)r * ndisp + col];
                        counts[r] = cnt;
                        mine = cnt < mine ? cnt : mine;
                        is_short |= ((double)(common + cnt) <= min_cols);
                    }
                    if (allgood) least = (int)sc.umin((unsigned long long)(unsigned)mine, cell);
                    any_short = sc.any(is_short) != 0;
                }
                sum_s.add(s_col);
                sum_v.add(v_col);
                if (any_short) continue;
                // tracing.py:272-278: snr = sum / sqrt(sum of variances), tested against the limit.  The division and
                // the square root sit on the serial chain, so a window that is below the limit by a clear margin
                // (1e-12 relative, ten thousand times the rounding of either side) is passed over without them; the
                // decision, and the S/N stored with the bin, still come from the exact expression.
                const double s_now = sum_s.value(), v_now = sum_v.value();
                if (snr_limit >= 0.0 && v_now > 0.0 &&
                    (s_now <= 0.0 || s_now * s_now < limit2 * v_now)) {
                    pending = true;             // below the limit: the value itself is only needed if the frame ends here
                    s_pending = s_now;
                    v_pending = v_now;
                    continue;
                }
                pending = false;
                snr = s_now / sqrt(v_now);
            }
            if (pending) snr = s_pending / sqrt(v_pending);      // the bin ended at the frame edge
            if (sc.tid == 0) {
                if (dir == 0) {
                    tmp_i[2 * n_blue] = edge - width;
                    tmp_i[2 * n_blue + 1] = edge;
                    tmp_snr[n_blue] = snr;
                } else {
                    bins_i[2 * n_red] = edge;          // parked at the front, shifted below
                    bins_i[2 * n_red + 1] = edge + width;
                    bins_snr[n_red] = snr;
                }
            }
            if (dir == 0) { ++n_blue; edge -= width; } else { ++n_red; edge += width; }
            width = 0;
            sc.sync();
        }
    }
    // bin_locations.sort(key=first): blue-ward bins reversed, then the red-ward ones
    sc.sync();
    // (simple two-step copy through the scratch arrays keeps this free of overlap hazards)
    for (int i = sc.tid; i < n_red; i += sc.size()) {
        tmp_i[2 * (n_blue + i)] = bins_i[2 * i];
        tmp_i[2 * (n_blue + i) + 1] = bins_i[2 * i + 1];
        tmp_snr[n_blue + i] = bins_snr[i];
    }
    sc.sync();
    for (int i = sc.tid; i < n_blue + n_red; i += sc.size()) {
        int src = (i < n_blue) ? (n_blue - 1 - i) : i;
        bins_i[2 * i] = tmp_i[2 * src];
        bins_i[2 * i + 1] = tmp_i[2 * src + 1];
        bins_snr[i] = tmp_snr[src];
    }
    sc.sync();
    return n_blue + n_red;
}

}  // namespace motes
