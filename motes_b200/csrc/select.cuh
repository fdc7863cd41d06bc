// Exact order statistics (np.median / np.nanmedian semantics) by radix selection on
// order-preserving 64-bit keys.  Used for: the full-frame median spatial profile
// (core.py:80), per-bin median profiles (tracing.py:505-507), the window medians of the
// limit extrapolation (tracing.py:101-102), sky-pixel medians (sky.py:341) and the
// non-zero error median (extraction.py:139).
#pragma once

#include "common.cuh"

namespace motes {

MOTES_HD unsigned long long key_of(double v) {
    unsigned long long b;
#if defined(__CUDA_ARCH__)
    b = (unsigned long long)__double_as_longlong(v);
#else
    memcpy(&b, &v, sizeof(b));
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

MOTES_HD double value_of(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
#if defined(__CUDA_ARCH__)
    v = __longlong_as_double((long long)b);
#else
    memcpy(&v, &b, sizeof(v));
#endif
    return v;
}

constexpr unsigned long long kKeyExcluded = ~0ull;     // sorts above every number (NaN / masked entries)

MOTES_HD unsigned long long key_or_excluded(double v) { return (v != v) ? kKeyExcluded : key_of(v); }

// Key source backed by an array.
struct ArrayKeys {
    const unsigned long long* p;
    MOTES_HD unsigned long long operator()(int i) const { return p[i]; }
};

// The digit of the k-th key among the 256 counters of one radix pass: cum = keys in the digits below it, equal = keys in
// it.  Generic form: every thread scans the counters.  A 256-thread CTA does it with one block-wide prefix sum instead
// (the scan was eight times the work of counting 16 keys per thread: row_median_kernel 3.7 -> 1.6 ms per 128 frames).
template <class Scope>
MOTES_HD void pick_digit(const Scope&, const uint32_t* hist, int k, int* digit, int* cum_out, int* equal) {
    int cum = 0;
    *digit = 255;
    for (int d = 0; d < 256; ++d) {
        int c = (int)hist[d];
        if (cum + c > k) { *digit = d; *equal = c; break; }
        cum += c;
    }
    *cum_out = cum;
}
#if defined(__CUDACC__)
__device__ __forceinline__ void pick_digit(const BlockScope& sc, const uint32_t* hist, int k, int* digit, int* cum_out, int* equal) {
    if (blockDim.x != 256) { pick_digit<BlockScope>(sc, hist, k, digit, cum_out, equal); return; }
    __shared__ int warp_sum[8];
    __shared__ int pick[3];
    const int t = sc.tid, lane = t & 31, warp = t >> 5;
    const int c = (int)hist[t];
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_sum[warp] = incl;
    if (t == 0) pick[0] = -1;
    __syncthreads();
    int total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int ws = warp_sum[w];
        if (w < warp) incl += ws;
        total += ws;
    }
    if (incl - c <= k && k < incl) { pick[0] = t; pick[1] = incl - c; pick[2] = c; }
    __syncthreads();
    if (pick[0] >= 0) { *digit = pick[0]; *cum_out = pick[1]; *equal = pick[2]; }
    else { *digit = 255; *cum_out = total; }
    __syncthreads();
}
#endif

// k-th smallest (0-based) of keys(0..n-1).  hist: 256 shared 32-bit counters.
// Returns the key; *less / *equal receive the number of keys below it and equal to it.
template <class Scope, class Keys>
MOTES_HD unsigned long long select_kth(const Scope& sc, const Keys& keys, int n, int k,
                                       uint32_t* hist, int* less_out, int* equal_out) {
    unsigned long long prefix = 0, mask = 0;
    int less = 0, equal = n;
    for (int pass = 7; pass >= 0; --pass) {
        const int shift = pass * 8;
        for (int d = sc.tid; d < 256; d += sc.size()) hist[d] = 0;
        sc.sync();
        for (int i = sc.tid; i < n; i += sc.size()) {
            unsigned long long key = keys(i);
            if ((key & mask) == prefix) sc.inc(&hist[(key >> shift) & 255]);
        }
        sc.sync();
        int cum = 0, digit = 255;
        pick_digit(sc, hist, k, &digit, &cum, &equal);
        k -= cum;
        less += cum;
        prefix |= (unsigned long long)digit << shift;
        mask |= 255ull << shift;
        sc.sync();
    }
    if (less_out) *less_out = less;
    if (equal_out) *equal_out = equal;
    return prefix;
}

// Median of the n keys (all of them valid numbers).  n == 0 -> NaN.  Even n -> (a+b)/2 of
// the two middle order statistics, as np.median does.  cell: one shared 64-bit word.
// With n_valid < n the top n - n_valid keys must be kKeyExcluded: the median is then taken over the
// n_valid real ones (np.nanmedian).
template <class Scope, class Keys>
MOTES_HD double median_of(const Scope& sc, const Keys& keys, int n_total, int n, uint32_t* hist,
                          unsigned long long* cell) {
    if (n <= 0) return NAN;
    int k1 = (n - 1) / 2, k2 = n / 2;
    int less, equal;
    unsigned long long a = select_kth(sc, keys, n_total, k1, hist, &less, &equal);
    double va = value_of(a);
    if (k2 == k1) return va;
    unsigned long long b = a;
    if (k2 >= less + equal) {
        unsigned long long best = ~0ull;
        for (int i = sc.tid; i < n_total; i += sc.size()) {
            unsigned long long key = keys(i);
            if (key > a && key < best) best = key;
        }
        b = sc.umin(best, cell);
    }
    return (va + value_of(b)) * 0.5;
}

template <class Scope>
MOTES_HD double median_of_keys(const Scope& sc, const unsigned long long* keys, int n, uint32_t* hist,
                               unsigned long long* cell) {
    return median_of(sc, ArrayKeys{keys}, n, n, hist, cell);
}

}  // namespace motes
