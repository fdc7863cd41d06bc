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
// The generator is kept as a ring of the last 4096 state words indexed by absolute stream
// position n (slot n & 4095): word n = mix(word n-624, word n-623, word n-397+... i.e. n-227),
// so any 227 consecutive new words are independent of each other and are produced by one
// parallel sweep.  A numpy RandomState (624 words + position) maps to words [0, 624) with
// `used` = position, and back (the block of 624 that holds the last consumed word).
//
// One loop iteration = one barrier: every thread (a) produces a word of the next sweep and
// (b) tests one pending output of the current column; after the barrier the per-warp accept
// counts give every accepted draw its ordinal k in the stream order, hence its place in numpy's
// row-major (n, 100) index array: pixel slot k / 100 of bootstrap sample k % 100.  Every slot is
// written exactly once, with the RANK of the drawn pixel, by a plain 16-bit shared-memory store
// (no atomics, nothing to clear between columns).  A sample's median is then the median of its n
// ranks, found by a bitwise binary search with warp-wide counts, and looked up in the sorted
// good-pixel values.
constexpr int kRing = 4096;          // > 624 (block) + kSlabs*256 (round) + 2*227 (look-ahead), power of two
constexpr int kSlabs = 4;            // outputs tested per thread and barrier
constexpr int kPackWords = 8;        // packed median search: up to 64 * 8 = 512 ranks per sample
constexpr int kSweep = kMtN - kMtM;      // 227

struct BootShared {
    uint32_t* ring;      // kRing words
    uint32_t* counts;    // 2 x 32 bytes: per-(slab, warp) accept counts, double buffered (16-byte aligned)
    int* ctl;            // 4 ints
    uint32_t* hist;      // 100 rows x row_words words: per sample, the ranks of its n drawn pixels (uint16)
    uint16_t* ranks;     // max_good
    double* meds;        // 100
    double* sq;          // 100
};

MOTES_HD int boot_row_words(int n) { return ((n + 1) / 2) | 1; }     // odd -> conflict-free row walks

// sum of 100 doubles in numpy's order, cooperatively when a warp is available
template <class Scope>
MOTES_HD double numpy_sum100(const Scope& sc, const double* a) {
#if defined(__CUDA_ARCH__)
    if (sc.size() >= 32) {
        double res = 0.0;
        if (sc.tid < 32) {
            int lane = sc.tid;
            double v = 0.0;
            if (lane < 8) {
                v = a[lane];
                for (int i = 1; i < 12; ++i) v += a[8 * i + lane];
            }
            double t = v + __shfl_down_sync(0xffffffffu, v, 1);
            double u = t + __shfl_down_sync(0xffffffffu, t, 2);
            double w = u + __shfl_down_sync(0xffffffffu, u, 4);
            res = ((w + a[96]) + a[97]) + a[98];
            res = res + a[99];
        }
        return res;            // valid in thread 0
    }
#endif
    return numpy_sum(a, kBootstrap);
}

template <class Scope>
MOTES_HD void sky_bootstrap_frame(const Scope& sc, uint32_t* state /* 625 words, in/out */, int ncol,
                                  const int* n_good, const int* flag, const double* sorted,
                                  const uint16_t* rank_of, size_t cs, double* sky, double* sky_err,
                                  const BootShared& sh) {
    uint32_t* ring = sh.ring;
    for (int i = sc.tid; i < kMtN; i += sc.size()) ring[i] = state[i];
    // stream positions modulo 2^32 (only differences and the low 10 bits are ever used)
    uint32_t gen = kMtN;
    uint32_t used = state[kMtN];
    unsigned long long used_total = state[kMtN];     // absolute, advanced once per column
    sc.sync();
    int parity = 0;
#if defined(__CUDA_ARCH__)
    const int lane = sc.tid & 31, warp = sc.tid >> 5;
    unsigned char* counts8 = reinterpret_cast<unsigned char*>(sh.counts);     // 2 x 8 bytes used
    const unsigned long long below_mask = (warp >= 8) ? ~0ull : ((1ull << (8 * warp)) - 1ull);
#endif

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
            const int rw = boot_row_words(n);
            uint16_t* slots = reinterpret_cast<uint16_t*>(sh.hist);
            for (int i = sc.tid; i < n; i += sc.size()) sh.ranks[i] = rank_of[(size_t)c * cs + i];
            sc.sync();
            const int need = n * kBootstrap;
            const uint32_t mask = rejection_mask((uint32_t)n);
            const uint32_t top = (uint32_t)(n - 1);
            const uint32_t used_at_start = used;
            int have = 0;
            while (have < need) {
                // (a) generation: kSlabs outputs per thread must be pending; one barrier per 227-word sweep
                const uint32_t want = (uint32_t)(kSlabs * sc.size());
                while ((uint32_t)(gen - used) < want) {
                    for (int j = sc.tid; j < kSweep; j += sc.size()) {
                        const uint32_t w = gen + (uint32_t)j;
                        ring[w & (kRing - 1)] = mt_mix(ring[(w - kMtN) & (kRing - 1)], ring[(w - kMtN + 1) & (kRing - 1)],
                                                       ring[(w - kSweep) & (kRing - 1)]);
                    }
                    gen += (uint32_t)kSweep;
                    sc.sync();
                }
                // (b) test kSlabs pending outputs per thread; stream order is slab-major, then thread
                int acc[kSlabs], off[kSlabs];
                uint32_t v[kSlabs];
#pragma unroll
                for (int k = 0; k < kSlabs; ++k) {
                    v[k] = mt_temper(ring[(used + (uint32_t)(k * sc.size() + sc.tid)) & (kRing - 1)]) & mask;
                    acc[k] = (v[k] <= top);
                }
                int total = 0;
#if defined(__CUDA_ARCH__)
                if (sc.size() > 1) {
                    // per-(slab, warp) accept counts as bytes: 32 bytes give every thread all of them
                    unsigned ballot[kSlabs];
#pragma unroll
                    for (int k = 0; k < kSlabs; ++k) {
                        ballot[k] = __ballot_sync(0xffffffffu, acc[k]);
                        if (lane == 0) counts8[parity * 32 + k * 8 + warp] = (unsigned char)__popc(ballot[k]);
                    }
                    sc.sync();
                    const uint4 c4 = *reinterpret_cast<const uint4*>(counts8 + parity * 32);      // slabs 0,1
                    const uint4 c5 = *reinterpret_cast<const uint4*>(counts8 + parity * 32 + 16); // slabs 2,3
                    const unsigned lo_w[kSlabs] = {c4.x, c4.z, c5.x, c5.z}, hi_w[kSlabs] = {c4.y, c4.w, c5.y, c5.w};
                    const unsigned lanes_below = (1u << lane) - 1u;
#pragma unroll
                    for (int k = 0; k < kSlabs; ++k) {
                        const unsigned long long all = ((unsigned long long)hi_w[k] << 32) | lo_w[k];
                        const unsigned long long low = all & below_mask;
                        off[k] = total + (int)(__vsadu4((unsigned)low, 0u) + __vsadu4((unsigned)(low >> 32), 0u)) +
                                 __popc(ballot[k] & lanes_below);
                        total += (int)(__vsadu4(lo_w[k], 0u) + __vsadu4(hi_w[k], 0u));
                    }
                    parity ^= 1;
                } else
#endif
                {
                    for (int k = 0; k < kSlabs; ++k) { off[k] = total; total += acc[k]; }
                }
                // (c) every accepted draw fills its slot of numpy's row-major (pixel, sample) array
                int last_at = 0;
#pragma unroll
                for (int k = 0; k < kSlabs; ++k) {
                    const int ord = have + off[k];
                    if (acc[k] && ord < need) {
                        const int pix = ord / kBootstrap, smp = ord - pix * kBootstrap;
                        slots[smp * (2 * rw) + pix] = sh.ranks[v[k]];
                        if (ord == need - 1) last_at = k * sc.size() + sc.tid + 1;
                    }
                }
                if (have + total >= need) {
                    // the accept with ordinal need-1 is the column's last draw; later outputs of this
                    // round belong to the next column and are tested again under its (mask, n)
                    if (last_at) sh.ctl[0] = last_at;
                    sc.sync();
                    used += (uint32_t)sh.ctl[0];
                    have = need;
                    sc.sync();
                } else {
                    have += total;
                    used += want;
                }
            }
            used_total += (unsigned long long)(used - used_at_start);
            // median of every bootstrap sample: bitwise binary search over its n ranks
            sc.sync();
            {
                const int k1 = (n - 1) / 2, k2 = n / 2;
                int nbits = 1;
                while ((1 << nbits) < n) ++nbits;
                bool done = false;
#if defined(__CUDA_ARCH__)
                if (sc.size() >= 32 && n <= 64 * kPackWords) {
                    // packed path: a warp owns a sample, each lane keeps up to kPackWords rank pairs in
                    // registers; "rank >= cand" for both halves of a word is bit 15 of (half + 0x8000 - cand)
                    done = true;
                    const int nwords = (n + 1) >> 1;
                    for (int smp = warp; smp < kBootstrap; smp += (sc.size() >> 5)) {
                        const uint32_t* row = sh.hist + smp * rw;
                        uint32_t reg[kPackWords];
#pragma unroll
                        for (int t = 0; t < kPackWords; ++t) {
                            const int wd = lane + 32 * t;
                            uint32_t wv = 0x7fff7fffu;                      // padding: larger than any rank
                            if (wd < nwords) {
                                wv = row[wd];
                                if (2 * wd + 1 >= n) wv = (wv & 0xffffu) | 0x7fff0000u;
                            }
                            reg[t] = wv;
                        }
                        const int slots_total = 64 * kPackWords;
                        int x1 = 0;
                        for (int bit = nbits - 1; bit >= 0; --bit) {
                            const int cand = x1 | (1 << bit);
                            const uint32_t kk = (uint32_t)(0x8000 - cand) * 0x00010001u;
                            uint32_t ge = 0;
#pragma unroll
                            for (int t = 0; t < kPackWords; ++t) ge += ((reg[t] + kk) >> 15) & 0x00010001u;
                            const int ge_all = __reduce_add_sync(0xffffffffu, (int)((ge & 0xffffu) + (ge >> 16)));
                            if (slots_total - ge_all <= k1) x1 = cand;
                        }
                        int x2 = x1;
                        if (k2 != k1) {
                            const uint32_t kk = (uint32_t)(0x8000 - (x1 + 1)) * 0x00010001u;
                            uint32_t ge = 0;
                            int nxt = 0x7fff;
#pragma unroll
                            for (int t = 0; t < kPackWords; ++t) {
                                ge += ((reg[t] + kk) >> 15) & 0x00010001u;
                                const int h0 = (int)(reg[t] & 0xffffu), h1 = (int)(reg[t] >> 16);
                                if (h0 > x1 && h0 < nxt) nxt = h0;
                                if (h1 > x1 && h1 < nxt) nxt = h1;
                            }
                            const int le = slots_total - __reduce_add_sync(0xffffffffu, (int)((ge & 0xffffu) + (ge >> 16)));
                            nxt = __reduce_min_sync(0xffffffffu, nxt);
                            if (le <= k2) x2 = nxt;
                        }
                        if (lane == 0) sh.meds[smp] = (k1 == k2) ? srt[x1] : (srt[x1] + srt[x2]) * 0.5;
                    }
                }
                const bool wide = sc.size() >= 32;
                const int g_id = wide ? warp : 0, g_n = wide ? (sc.size() >> 5) : 1;
                const int l_id = wide ? lane : 0, l_n = wide ? 32 : 1;
#else
                const int g_id = 0, g_n = 1, l_id = 0, l_n = 1;
#endif
                for (int smp = g_id; smp < kBootstrap && !done; smp += g_n) {
                    const uint16_t* row = slots + smp * (2 * rw);
                    int x1 = 0;
                    for (int bit = nbits - 1; bit >= 0; --bit) {
                        const int cand = x1 | (1 << bit);
                        int cnt = 0;
                        for (int j = l_id; j < n; j += l_n) cnt += ((int)row[j] < cand);
#if defined(__CUDA_ARCH__)
                        if (wide) cnt = __reduce_add_sync(0xffffffffu, cnt);
#endif
                        if (cnt <= k1) x1 = cand;
                    }
                    int x2 = x1;
                    if (k2 != k1) {
                        int le = 0, nxt = 0x7fffffff;
                        for (int j = l_id; j < n; j += l_n) {
                            const int r = (int)row[j];
                            le += (r <= x1);
                            if (r > x1 && r < nxt) nxt = r;
                        }
#if defined(__CUDA_ARCH__)
                        if (wide) { le = __reduce_add_sync(0xffffffffu, le); nxt = __reduce_min_sync(0xffffffffu, nxt); }
#endif
                        if (le <= k2) x2 = nxt;
                    }
                    if (l_id == 0) sh.meds[smp] = (k1 == k2) ? srt[x1] : (srt[x1] + srt[x2]) * 0.5;
                }
            }
        } else {
            for (int smp = sc.tid; smp < kBootstrap; smp += sc.size()) sh.meds[smp] = srt[0];
        }
        sc.sync();
        // np.nanmean / np.std of the 100 medians in numpy's summation order (sky.py:40-41)
        double mean = numpy_sum100(sc, sh.meds) / (double)kBootstrap;
        if (sc.tid == 0) sh.sq[0] = mean;
        sc.sync();
        mean = sh.sq[0];
        sc.sync();
        for (int smp = sc.tid; smp < kBootstrap; smp += sc.size()) { double d = sh.meds[smp] - mean; sh.sq[smp] = d * d; }
        sc.sync();
        double var = numpy_sum100(sc, sh.sq) / (double)kBootstrap;
        if (sc.tid == 0) {
            sky[c] = mean;
            sky_err[c] = sqrt(var) / kSqrt99;
        }
        sc.sync();
    }

    // back to numpy's layout: the 624-word block that holds the last consumed word.  `block` is the
    // absolute index of its first word; ring indices are recovered relative to `used`.
    const unsigned long long block = (used_total > 0) ? ((used_total - 1) / kMtN) * kMtN : 0;
    const uint32_t into_block = (uint32_t)(used_total - block);          // numpy's position, 1..624 (or 0)
    const uint32_t block32 = used - into_block;                           // ring index of the block start
    while ((uint32_t)(gen - block32) < (uint32_t)kMtN) {
        for (int j = sc.tid; j < kSweep; j += sc.size()) {
            const uint32_t w = gen + (uint32_t)j;
            ring[w & (kRing - 1)] = mt_mix(ring[(w - kMtN) & (kRing - 1)], ring[(w - kMtN + 1) & (kRing - 1)],
                                           ring[(w - kSweep) & (kRing - 1)]);
        }
        gen += (uint32_t)kSweep;
        sc.sync();
    }
    sc.sync();
    for (int i = sc.tid; i < kMtN; i += sc.size()) state[i] = ring[(block32 + (uint32_t)i) & (kRing - 1)];
    if (sc.tid == 0) state[kMtN] = into_block;
    sc.sync();
}

// ---- 2b. bootstrap polynomial sky (sky.sky_polynomial, sky.py:177-254) ------------------------
// Per column: every good sky pixel j is replaced 100 times by N(pixel_j, err_j) drawn with numpy's
// legacy polar Gaussian (np.random.normal: two 53-bit uniforms per attempt, accept 0 < r2 < 1,
// return f*x2 and cache f*x1), a polynomial of the requested order is fitted to each of the 100
// perturbed sky vectors, and the model / uncertainty of every spatial pixel are the median / std
// of the 100 polynomial values there.
//
// The generator words are consumed four at a time (one attempt); accepted attempts are numbered
// in stream order: attempt k feeds pixel k / 50, samples 2 (k % 50) and 2 (k % 50) + 1.
// The 100 least-squares fits share one design matrix: a QR (modified Gram-Schmidt) of the
// column-normalised Vandermonde matrix is built once per column and every sample is a
// projection + back-substitution.  np.polyfit solves the same scaled system with LAPACK gelsd;
// the two agree to ~1e-13 in the fitted values, i.e. this mode is
// tolerance-parity, not bit-parity (CUDA's log() also differs from libm's in the last ulp).
constexpr int kMaxPoly = 6;              // order <= 5

struct PolyShared {
    uint32_t* ring;      // kRing
    uint32_t* counts;    // 16 words
    int* ctl;            // 4
    double* q;           // max_good * kMaxPoly : Vandermonde -> Q
    double* rmat;        // kMaxPoly * kMaxPoly
    double* colscale;    // kMaxPoly
    double* coef;        // 100 * kMaxPoly (ascending powers)
    double* loc;         // max_good
    double* scl;         // max_good
    double* red;         // 40 doubles: block reduction scratch
    double* samples;     // per warp 100 values + the +inf the padding indices point at (4 doubles of room)
    uint16_t* rows;      // max_good
};

template <class Scope>
MOTES_HD double scope_dsum(const Scope& sc, double v, double* red) {
#if defined(__CUDA_ARCH__)
    if (sc.size() > 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int lane = sc.tid & 31, warp = sc.tid >> 5, nw = (sc.size() + 31) >> 5;
        if (lane == 0) red[warp] = v;
        sc.sync();
        double t = 0.0;
        for (int k = 0; k < nw; ++k) t += red[k];       // fixed order: every thread gets the same bits
        sc.sync();
        return t;
    }
#endif
    (void)red;
    return v;
}

// The fit half of sky.sky_polynomial for one column, once the 100 perturbed sky vectors y[j * 100 + s] exist:
// QR of the column-normalised Vandermonde matrix of the good pixels (sh.rows), one projection + back-substitution
// per sample, then per spatial pixel the median / std of the 100 polynomial values (sky.py:226-239).
template <class Scope>
MOTES_HD void sky_poly_fit_column(const Scope& sc, const PolyShared& sh, int n, int p, int nspat, int ndisp, int c,
                                  int spatial_floor, const double* y, double* model, double* model_err) {
#if defined(__CUDA_ARCH__)
    const int lane = sc.tid & 31, warp = sc.tid >> 5;
#endif
    // ---- QR of the column-normalised Vandermonde matrix (np.polyfit: lhs /= sqrt((lhs*lhs).sum(0)))
    // column k of the design holds x^(order-k) as np.vander does
    for (int g = sc.tid; g < n; g += sc.size()) {
        const double x = (double)((int)sh.rows[g] + spatial_floor);
        double pw = 1.0;
        for (int k = p - 1; k >= 0; --k) { sh.q[(size_t)g * kMaxPoly + k] = pw; pw *= x; }
    }
    sc.sync();
    for (int k = 0; k < p; ++k) {
        double part = 0.0;
        for (int g = sc.tid; g < n; g += sc.size()) { double t = sh.q[(size_t)g * kMaxPoly + k]; part += t * t; }
        double nrm = sqrt(scope_dsum(sc, part, sh.red));
        if (sc.tid == 0) sh.colscale[k] = nrm;
        for (int g = sc.tid; g < n; g += sc.size()) sh.q[(size_t)g * kMaxPoly + k] /= nrm;
        sc.sync();
    }
    for (int k = 0; k < p; ++k) {
        double part = 0.0;
        for (int g = sc.tid; g < n; g += sc.size()) { double t = sh.q[(size_t)g * kMaxPoly + k]; part += t * t; }
        double rkk = sqrt(scope_dsum(sc, part, sh.red));
        if (sc.tid == 0) sh.rmat[k * kMaxPoly + k] = rkk;
        for (int g = sc.tid; g < n; g += sc.size()) sh.q[(size_t)g * kMaxPoly + k] /= rkk;
        sc.sync();
        for (int j2 = k + 1; j2 < p; ++j2) {
            part = 0.0;
            for (int g = sc.tid; g < n; g += sc.size()) part += sh.q[(size_t)g * kMaxPoly + k] * sh.q[(size_t)g * kMaxPoly + j2];
            double rkj = scope_dsum(sc, part, sh.red);
            if (sc.tid == 0) sh.rmat[k * kMaxPoly + j2] = rkj;
            for (int g = sc.tid; g < n; g += sc.size()) sh.q[(size_t)g * kMaxPoly + j2] -= rkj * sh.q[(size_t)g * kMaxPoly + k];
            sc.sync();
        }
    }
    // ---- one projection + back-substitution per bootstrap sample; the rows are split between as many groups of 100
    // threads as the scope has (partial sums of the second group go through sh.samples)
    const int groups = sc.size() >= 2 * kBootstrap ? 2 : 1;
    for (int item = sc.tid; item < groups * kBootstrap; item += sc.size()) {
        const int grp = item / kBootstrap, smp = item % kBootstrap;
        const int g0 = (int)(((long long)n * grp) / groups), g1 = (int)(((long long)n * (grp + 1)) / groups);
        double z[kMaxPoly];
        for (int k = 0; k < p; ++k) z[k] = 0.0;
#pragma unroll 4
        for (int g = g0; g < g1; ++g) {
            const double yv = y[(size_t)g * kBootstrap + smp];
            for (int k = 0; k < p; ++k) z[k] = fma(sh.q[(size_t)g * kMaxPoly + k], yv, z[k]);
        }
        double* part = (grp == 0 ? sh.coef : sh.samples) + smp * kMaxPoly;
        for (int k = 0; k < p; ++k) part[k] = z[k];
    }
    sc.sync();
    for (int smp = sc.tid; smp < kBootstrap; smp += sc.size()) {
        double z[kMaxPoly], cf[kMaxPoly];
        for (int k = 0; k < p; ++k) z[k] = sh.coef[smp * kMaxPoly + k] + (groups > 1 ? sh.samples[smp * kMaxPoly + k] : 0.0);
        for (int k = p - 1; k >= 0; --k) {
            double acc2 = z[k];
            for (int j2 = k + 1; j2 < p; ++j2) acc2 -= sh.rmat[k * kMaxPoly + j2] * cf[j2];
            cf[k] = acc2 / sh.rmat[k * kMaxPoly + k];
        }
        // un-scale and flip to ascending powers (sky.py:230)
        for (int k = 0; k < p; ++k) sh.coef[smp * kMaxPoly + (p - 1 - k)] = cf[k] / sh.colscale[k];
    }
    sc.sync();
    // ---- per spatial pixel: median and std over the 100 polynomial values (sky.py:231-239)
#if defined(__CUDA_ARCH__)
    if (sc.size() > 1) {
        // One warp per run of neighbouring pixels.  The 100 polynomials are smooth, so their order at one pixel is
        // almost their order at the next: the warp keeps the sample indices in sorted order (four per lane, padded to
        // 128 with +inf) and repairs the order with odd-even transposition passes - a full bitonic sort only for the
        // first pixel of the run or when the repair takes long.  The selection is exact either way.
        const int nwarp = (sc.size() + 31) >> 5;
        const int per = (nspat + nwarp - 1) / nwarp;
        const int px0 = warp * per, px1 = min(nspat, px0 + per);
        double* vals = sh.samples + (size_t)warp * (kBootstrap + 4);
        int id[4];
        double sv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) id[k] = (4 * lane + k < kBootstrap) ? 4 * lane + k : kBootstrap;
        if (lane == 0) vals[kBootstrap] = INFINITY;          // where the padding indices point
        bool have_order = false;
        for (int px = px0; px < px1; ++px) {
            const double x = (double)(px + spatial_floor);
            for (int smp = lane; smp < kBootstrap; smp += 32) {
                double acc2 = 0.0, pw = 1.0;
                for (int k = 0; k < p; ++k) { acc2 += pw * sh.coef[smp * kMaxPoly + k]; pw *= x; }
                vals[smp] = acc2;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; ++k) sv[k] = vals[id[k]];
            // (value, index) order; the index breaks ties so that both sides of an exchange agree
            auto before = [](double va, int ia, double vb, int ib) { return va < vb || (va == vb && ia < ib); };
            auto exchange = [&](int ka, int kb) {           // ascending inside the lane
                if (before(sv[kb], id[kb], sv[ka], id[ka])) {
                    const double tv = sv[ka]; sv[ka] = sv[kb]; sv[kb] = tv;
                    const int ti = id[ka]; id[ka] = id[kb]; id[kb] = ti;
                }
            };
            bool sorted_now = false;
            if (have_order) {
                for (int pass = 0; pass < 12; ++pass) {
                    const double nv = __shfl_down_sync(0xffffffffu, sv[0], 1);
                    const int ni = __shfl_down_sync(0xffffffffu, id[0], 1);
                    const bool ok = !before(sv[1], id[1], sv[0], id[0]) && !before(sv[2], id[2], sv[1], id[1]) &&
                                    !before(sv[3], id[3], sv[2], id[2]) && (lane == 31 || !before(nv, ni, sv[3], id[3]));
                    if (__all_sync(0xffffffffu, ok)) { sorted_now = true; break; }
                    exchange(0, 1);
                    exchange(2, 3);
                    exchange(1, 2);
                    // slot 3 of this lane against slot 0 of the next one, both from the values before the exchange
                    const double dv = __shfl_down_sync(0xffffffffu, sv[0], 1), uv = __shfl_up_sync(0xffffffffu, sv[3], 1);
                    const int di = __shfl_down_sync(0xffffffffu, id[0], 1), ui = __shfl_up_sync(0xffffffffu, id[3], 1);
                    if (lane < 31 && before(dv, di, sv[3], id[3])) { sv[3] = dv; id[3] = di; }
                    if (lane > 0 && before(sv[0], id[0], uv, ui)) { sv[0] = uv; id[0] = ui; }
                }
            }
            if (!sorted_now) {
                // warp bitonic sort of the 128 (value, index) pairs
#pragma unroll
                for (int size = 2; size <= 128; size <<= 1) {
#pragma unroll
                    for (int d = size >> 1; d > 0; d >>= 1) {
                        if (d < 4) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if ((k & d) == 0) {
                                    const bool up = (size < 4) ? ((k & size) == 0) : ((lane & (size / 4)) == 0);
                                    const bool wrong = up ? before(sv[k | d], id[k | d], sv[k], id[k]) : before(sv[k], id[k], sv[k | d], id[k | d]);
                                    if (wrong) {
                                        const double tv = sv[k]; sv[k] = sv[k | d]; sv[k | d] = tv;
                                        const int ti = id[k]; id[k] = id[k | d]; id[k | d] = ti;
                                    }
                                }
                            }
                        } else {
                            const int dl = d / 4;
                            const bool up = (lane & (size / 4)) == 0 || size == 128;
                            const bool keep_min = (((lane & dl) == 0) == up);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const double ov = __shfl_xor_sync(0xffffffffu, sv[k], dl);
                                const int oi = __shfl_xor_sync(0xffffffffu, id[k], dl);
                                const bool other_first = before(ov, oi, sv[k], id[k]);
                                if (other_first == keep_min) { sv[k] = ov; id[k] = oi; }
                            }
                        }
                    }
                }
                have_order = true;
            }
            // np.median of 100: mean of the order statistics 49 and 50 (lane 12, slots 1 and 2)
            const double m1 = __shfl_sync(0xffffffffu, sv[1], 12), m2 = __shfl_sync(0xffffffffu, sv[2], 12);
            double psum = 0.0;
            for (int e = lane; e < kBootstrap; e += 32) psum += vals[e];
            for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
            const double mean = psum / (double)kBootstrap;
            double sq = 0.0;
            for (int e = lane; e < kBootstrap; e += 32) { const double d = vals[e] - mean; sq += d * d; }
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            if (lane == 0) {
                model[(size_t)px * ndisp + c] = (m1 + m2) * 0.5;
                model_err[(size_t)px * ndisp + c] = sqrt(sq / (double)kBootstrap);
            }
            __syncwarp();
        }
        sc.sync();
        return;
    }
#endif
    // one lane (host build, one-thread scopes): rank counting
    double* vals = sh.samples;
    for (int px = 0; px < nspat; ++px) {
        const double x = (double)(px + spatial_floor);
        for (int smp = 0; smp < kBootstrap; ++smp) {
            double acc2 = 0.0, pw = 1.0;
            for (int k = 0; k < p; ++k) { acc2 += pw * sh.coef[smp * kMaxPoly + k]; pw *= x; }
            vals[smp] = acc2;
        }
        double m1 = 0.0, m2 = 0.0, psum = 0.0;
        for (int e = 0; e < kBootstrap; ++e) {
            const double v = vals[e];
            int rank = 0;
            for (int j2 = 0; j2 < kBootstrap; ++j2) { const double o = vals[j2]; rank += (o < v) || (o == v && j2 < e); }
            if (rank == 49) m1 = v;
            if (rank == 50) m2 = v;
            psum += v;
        }
        const double mean = psum / (double)kBootstrap;
        double sq = 0.0;
        for (int e = 0; e < kBootstrap; ++e) { const double d = vals[e] - mean; sq += d * d; }
        model[(size_t)px * ndisp + c] = (m1 + m2) * 0.5;
        model_err[(size_t)px * ndisp + c] = sqrt(sq / (double)kBootstrap);
    }
    sc.sync();
}

// y: global scratch of 100 * max_good doubles for this frame, pixel-major (y[j * 100 + s]).
// model / model_err: (nspat, ndisp) frames receiving this column's result at [pixel * ndisp + c].
template <class Scope>
MOTES_HD void sky_bootstrap_poly_frame(const Scope& sc, uint32_t* state, int ncol, int nspat, int ndisp, int order,
                                       int spatial_floor, const double* errs /* frame */, const int* n_good,
                                       const int* flag, const double* sorted, const uint16_t* rank_of,
                                       const uint16_t* good_row, size_t cs, double* y, double* model,
                                       double* model_err, const PolyShared& sh) {
    uint32_t* ring = sh.ring;
    for (int i = sc.tid; i < kMtN; i += sc.size()) ring[i] = state[i];
    uint32_t gen = kMtN;
    uint32_t used = state[kMtN];
    unsigned long long used_total = state[kMtN];
    sc.sync();
    const int p = order + 1;
    int parity = 0;
#if defined(__CUDA_ARCH__)
    const int lane = sc.tid & 31, warp = sc.tid >> 5;
    unsigned char* counts8 = reinterpret_cast<unsigned char*>(sh.counts);
    const unsigned long long below_mask = (warp >= 8) ? ~0ull : ((1ull << (8 * warp)) - 1ull);
#endif

    for (int c = 0; c < ncol; ++c) {
        if (flag[c] != kSkyModelled) continue;
        const int n = n_good[c];
        if (n < p) {          // np.polyfit would warn (rank deficient); the reference result is undefined here
            for (int i = sc.tid; i < nspat; i += sc.size()) { model[(size_t)i * ndisp + c] = NAN; model_err[(size_t)i * ndisp + c] = NAN; }
            continue;
        }
        // good pixels of this column in row order: value, error, axis
        for (int g = sc.tid; g < n; g += sc.size()) {
            const int row = (int)good_row[(size_t)c * cs + g];
            sh.rows[g] = (uint16_t)row;
            sh.loc[g] = sorted[(size_t)c * cs + rank_of[(size_t)c * cs + g]];
            sh.scl[g] = errs[(size_t)row * ndisp + c];
        }
        sc.sync();

        // ---- 100 Gaussian-perturbed copies of the sky vector (sky.py:218-225)
        const int need = n * (kBootstrap / 2);          // accepted attempts (pairs of deviates)
        const uint32_t used_at_start = used;
        int have = 0;
        while (have < need) {
            const uint32_t pending = gen - used;
            const bool sweep = pending <= (uint32_t)(kSweep + 3);
            if (sweep) {
                for (int j = sc.tid; j < kSweep; j += sc.size()) {
                    const uint32_t w = gen + (uint32_t)j;
                    ring[w & (kRing - 1)] = mt_mix(ring[(w - kMtN) & (kRing - 1)], ring[(w - kMtN + 1) & (kRing - 1)],
                                                   ring[(w - kSweep) & (kRing - 1)]);
                }
            }
            const int attempts = (int)(pending / 4u);
            const int round = attempts < sc.size() ? attempts : sc.size();
            int acc = 0;
            double x1 = 0.0, x2 = 0.0, r2 = 0.0;
            if (sc.tid < round) {
                const uint32_t base = used + 4u * (uint32_t)sc.tid;
                const uint32_t a0 = mt_temper(ring[base & (kRing - 1)]), b0 = mt_temper(ring[(base + 1) & (kRing - 1)]);
                const uint32_t a1 = mt_temper(ring[(base + 2) & (kRing - 1)]), b1 = mt_temper(ring[(base + 3) & (kRing - 1)]);
                const double u1 = ((double)(a0 >> 5) * 67108864.0 + (double)(b0 >> 6)) / 9007199254740992.0;
                const double u2 = ((double)(a1 >> 5) * 67108864.0 + (double)(b1 >> 6)) / 9007199254740992.0;
                x1 = 2.0 * u1 - 1.0;
                x2 = 2.0 * u2 - 1.0;
                r2 = x1 * x1 + x2 * x2;
                acc = (r2 < 1.0 && r2 != 0.0);
            }
            int off, total;
#if defined(__CUDA_ARCH__)
            if (sc.size() > 1) {
                const unsigned ballot = __ballot_sync(0xffffffffu, acc);
                if (lane == 0) counts8[parity * 8 + warp] = (unsigned char)__popc(ballot);
                sc.sync();
                const unsigned long long all = *reinterpret_cast<const unsigned long long*>(counts8 + parity * 8);
                const unsigned long long low = all & below_mask;
                total = (int)(__vsadu4((unsigned)all, 0u) + __vsadu4((unsigned)(all >> 32), 0u));
                off = (int)(__vsadu4((unsigned)low, 0u) + __vsadu4((unsigned)(low >> 32), 0u)) +
                      __popc(ballot & ((1u << lane) - 1u));
                parity ^= 1;
            } else
#endif
            {
                sc.sync();
                off = 0;
                total = acc;
            }
            if (sweep) gen += (uint32_t)kSweep;
            const int ord = have + off;
            if (acc && ord < need) {
                const double f = sqrt(-2.0 * log(r2) / r2);
                const int j = ord / (kBootstrap / 2), smp = 2 * (ord % (kBootstrap / 2));
                y[(size_t)j * kBootstrap + smp] = sh.loc[j] + sh.scl[j] * (f * x2);         // returned first
                y[(size_t)j * kBootstrap + smp + 1] = sh.loc[j] + sh.scl[j] * (f * x1);     // the cached deviate
            }
            if (have + total >= need) {
                if (acc && ord == need - 1) sh.ctl[0] = sc.tid + 1;
                sc.sync();
                used += 4u * (uint32_t)sh.ctl[0];
                have = need;
                sc.sync();
            } else {
                have += total;
                used += 4u * (uint32_t)round;
            }
        }
        used_total += (unsigned long long)(used - used_at_start);
        sc.sync();

        sky_poly_fit_column(sc, sh, n, p, nspat, ndisp, c, spatial_floor, y, model, model_err);
    }

    const unsigned long long block = (used_total > 0) ? ((used_total - 1) / kMtN) * kMtN : 0;
    const uint32_t into_block = (uint32_t)(used_total - block);
    const uint32_t block32 = used - into_block;
    while ((uint32_t)(gen - block32) < (uint32_t)kMtN) {
        for (int j = sc.tid; j < kSweep; j += sc.size()) {
            const uint32_t w = gen + (uint32_t)j;
            ring[w & (kRing - 1)] = mt_mix(ring[(w - kMtN) & (kRing - 1)], ring[(w - kMtN + 1) & (kRing - 1)],
                                           ring[(w - kSweep) & (kRing - 1)]);
        }
        gen += (uint32_t)kSweep;
        sc.sync();
    }
    sc.sync();
    for (int i = sc.tid; i < kMtN; i += sc.size()) state[i] = ring[(block32 + (uint32_t)i) & (kRing - 1)];
    if (sc.tid == 0) state[kMtN] = into_block;
    sc.sync();
}

// sky.py:375-378 for a column with a per-pixel model (order > 0), then the chip-gap re-marking
MOTES_HD void sky_apply_column_poly(double* data, double* errs, int nspat, int ndisp, int c, int flag,
                                    const double* model, const double* model_err) {
    int lt = 0, eq = 0, n = 0;
    double max_lt = -INFINITY, min_gt = INFINITY;
    for (int r = 0; r < nspat; ++r) {
        size_t g = (size_t)r * ndisp + c;
        double d = data[g];
        if (flag == kSkyModelled) {
            d = d - model[g];
            data[g] = d;
            double e = errs[g], me = model_err[g];
            errs[g] = sqrt((e * e) + (me * me));
        }
        if (d == d) {
            ++n;
            if (d < 0.0) { ++lt; if (d > max_lt) max_lt = d; }
            else if (d == 0.0) ++eq;
            else if (d < min_gt) min_gt = d;
        }
    }
    if (median_equals(0.0, n, lt, eq, max_lt, min_gt))
        for (int r = 0; r < nspat; ++r) data[(size_t)r * ndisp + c] += 1.0;
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
