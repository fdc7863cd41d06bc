// Shared definitions for the motes_b200 device code.
//
// Every numerical routine is written once, as a template over a "lane context":
//   * WarpCtx  - the real thing: 32 lanes of one warp, reductions by warp shuffles;
//   * SeqCtx   - one lane, reductions are the identity.  It exists so that
//                tests/hostsim can compile the *same* solver source with g++ and
//                check its logic against scipy on a machine without a GPU.  It is
//                test infrastructure: the product library never instantiates it.
#pragma once

#include <cmath>
#include <cstdint>
#include <cfloat>

#if defined(__CUDACC__)
#define MOTES_HD __host__ __device__ __forceinline__
#define MOTES_D __device__ __forceinline__
#else
#define MOTES_HD inline
#define MOTES_D inline
#endif

namespace motes {

constexpr int kParams = 6;            // A, c, alpha, beta, B, m   (common.py:82-87)
constexpr double kEps = 2.220446049250313e-16;
constexpr double kFdRelStep = 1.4901161193847656e-08;   // sqrt(eps): scipy _numdiff 2-point step
constexpr double kBetaStart = 4.765;  // tracing.py:581
constexpr int kBootstrap = 100;       // sky.py:36,216

// status codes shared by every stage (see include/motes_b200.h)
enum FitStatus : int {
    kFitMaxNfev = 0, kFitGtol = 1, kFitFtol = 2, kFitXtol = 3, kFitBoth = 4,
    kFitX0Infeasible = -1,   // scipy: "Initial guess is outside of provided bounds" (ValueError)
    kFitNonFinite = -2,      // scipy: "Residuals are not finite in the initial point." (ValueError)
    kFitBadBounds = -3,      // scipy: "Each lower bound must be strictly less than each upper bound."
};

struct SeqCtx {
    static constexpr int kLanes = 1;
    int lane = 0;
    MOTES_HD double sum(double v) const { return v; }
    template <int N>
    MOTES_HD void sum_vec(const double (&in)[N], double (&out)[N]) const {
        for (int j = 0; j < N; ++j) out[j] = in[j];
    }
    MOTES_HD double max(double v) const { return v; }
    MOTES_HD int all(int p) const { return p; }
    MOTES_HD void sync() const {}
};

#if defined(__CUDACC__)
struct WarpCtx {
    static constexpr int kLanes = 32;
    int lane;
    // Butterfly all-reduce.  IEEE addition is commutative, so every lane ends with the
    // same bits and the replicated scalar logic of the solver stays in lock step.
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ double max(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    template <int N>
    __device__ __forceinline__ void sum_vec(const double (&in)[N], double (&out)[N]) const {
#pragma unroll
        for (int j = 0; j < N; ++j) out[j] = sum(in[j]);
    }
    __device__ __forceinline__ int all(int p) const { return __all_sync(0xffffffffu, p); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};

// A whole CTA as one lane context: `lane` is the thread index, reductions combine the warps'
// butterfly sums in warp order through a double-buffered shared scratch (one barrier per call).
template <int THREADS>
struct BlockCtx {
    static constexpr int kLanes = THREADS;
    static constexpr int kWarps = THREADS / 32;
    static constexpr int kRedDoubles = 2 * kWarps * 8;
    int lane;
    double* red;            // kRedDoubles of shared memory
    mutable int flip;
    // N sums at once.  Instead of N full butterflies, the lanes first split the values between them
    // (each halving stage exchanges half of what a lane still holds), so 8 values cost 9 shuffles
    // instead of 40; the lane that ends up with value j is lane j << (5 - log2 P).
    template <int H>
    __device__ __forceinline__ static void halve(double* t, int l, int mask) {
        const bool upper = (l & mask) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const double send = upper ? t[i] : t[i + H];
            const double keep = upper ? t[i + H] : t[i];
            t[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
    }
    template <int N>
    __device__ __forceinline__ void sum_vec(const double (&in)[N], double (&out)[N]) const {
        static_assert(N <= 8, "at most 8 values per reduction");
        constexpr int P = N <= 1 ? 1 : N <= 2 ? 2 : N <= 4 ? 4 : 8;
        constexpr int LOG = P == 1 ? 0 : P == 2 ? 1 : P == 4 ? 2 : 3;
        double t[P];
#pragma unroll
        for (int j = 0; j < P; ++j) t[j] = j < N ? in[j] : 0.0;
        const int l = lane & 31;
        if (P >= 8) halve<4>(t, l, 16);
        if (P >= 4) halve<(P >= 8 ? 2 : 2)>(t, l, P >= 8 ? 8 : 16);
        if (P >= 2) halve<1>(t, l, P >= 8 ? 4 : P >= 4 ? 8 : 16);
#pragma unroll
        for (int o = 16 >> LOG; o > 0; o >>= 1) t[0] += __shfl_xor_sync(0xffffffffu, t[0], o);
        double* buf = red + flip * (kWarps * 8);
        flip ^= 1;
        const int j = l >> (5 - LOG);
        if ((l & ((1 << (5 - LOG)) - 1)) == 0 && j < N) buf[(lane >> 5) * 8 + j] = t[0];
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < N; ++jj) {
            double acc = buf[jj];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) acc += buf[w * 8 + jj];
            out[jj] = acc;
        }
    }
    __device__ __forceinline__ double sum(double v) const {
        const double in[1] = {v};
        double out[1];
        sum_vec<1>(in, out);
        return out[0];
    }
    __device__ __forceinline__ int all(int p) const { return __syncthreads_and(p); }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
#endif

// ---------------------------------------------------------------------------------------
// Thread scopes for the cooperative (non-solver) routines: the same source runs on a warp,
// on a whole CTA, or on one host thread (tests/hostsim).
//   tid / size      this thread's index and the number of cooperating threads
//   sync()          barrier over the scope (also orders shared-memory traffic)
//   inc(p)          atomic ++ on a 32-bit counter shared by the scope
//   umin(v)/dsum(v) scope-wide reductions; every thread receives the result
// ---------------------------------------------------------------------------------------
struct SeqScope {
    int tid = 0;
    static constexpr int size_c = 1;
    MOTES_HD int size() const { return 1; }
    MOTES_HD void sync() const {}
    MOTES_HD void inc(uint32_t* p) const { ++*p; }
    MOTES_HD void add(uint32_t* p, uint32_t v) const { *p += v; }
    MOTES_HD unsigned long long umin(unsigned long long v, unsigned long long*) const { return v; }
    MOTES_HD double dsum(double v, double*) const { return v; }
    MOTES_HD int isum(int v, int*) const { return v; }
    MOTES_HD int any(int p) const { return p; }
    // exclusive prefix count of `flag` over the scope in thread order; *total = scope-wide count
    MOTES_HD int scan_flags(int flag, int* total, uint32_t*) const { *total = flag; return 0; }
};

#if defined(__CUDACC__)
struct WarpScope {
    int tid;   // lane
    __device__ __forceinline__ int size() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ void inc(uint32_t* p) const { atomicAdd(p, 1u); }
    __device__ __forceinline__ void add(uint32_t* p, uint32_t v) const { atomicAdd(p, v); }
    __device__ __forceinline__ unsigned long long umin(unsigned long long v, unsigned long long*) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
            v = t < v ? t : v;
        }
        return v;
    }
    __device__ __forceinline__ double dsum(double v, double*) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ int isum(int v, int*) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ int any(int p) const { return __any_sync(0xffffffffu, p); }
    __device__ __forceinline__ int scan_flags(int flag, int* total, uint32_t*) const {
        unsigned b = __ballot_sync(0xffffffffu, flag);
        *total = __popc(b);
        return __popc(b & ((1u << tid) - 1u));
    }
};

// Whole CTA.  Reductions go through a caller-provided shared-memory cell.
struct BlockScope {
    int tid;
    __device__ __forceinline__ int size() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void inc(uint32_t* p) const { atomicAdd(p, 1u); }
    __device__ __forceinline__ void add(uint32_t* p, uint32_t v) const { atomicAdd(p, v); }
    __device__ __forceinline__ unsigned long long umin(unsigned long long v, unsigned long long* cell) const {
        if (tid == 0) *cell = ~0ull;
        __syncthreads();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
            v = t < v ? t : v;
        }
        if ((tid & 31) == 0) atomicMin(cell, v);
        __syncthreads();
        unsigned long long r = *cell;
        __syncthreads();
        return r;
    }
    __device__ __forceinline__ int isum(int v, int* cell) const {
        if (tid == 0) *cell = 0;
        __syncthreads();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) atomicAdd(cell, v);
        __syncthreads();
        int r = *cell;
        __syncthreads();
        return r;
    }
    __device__ __forceinline__ int any(int p) const { return __syncthreads_or(p); }
    // scratch: 33 shared words.  Thread order = tid order.
    __device__ __forceinline__ int scan_flags(int flag, int* total, uint32_t* scratch) const {
        unsigned b = __ballot_sync(0xffffffffu, flag);
        int lane = tid & 31, warp = tid >> 5, nwarp = (blockDim.x + 31) >> 5;
        if (lane == 0) scratch[warp] = __popc(b);
        __syncthreads();
        if (warp == 0) {
            uint32_t v = (lane < nwarp) ? scratch[lane] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            scratch[lane] = incl - v;                  // exclusive warp offsets
            if (lane == 31) scratch[32] = incl;        // grand total
        }
        __syncthreads();
        int off = (int)scratch[warp] + __popc(b & ((1u << lane) - 1u));
        *total = (int)scratch[32];
        __syncthreads();
        return off;
    }
};
#endif

MOTES_HD bool is_finite(double x) { return fabs(x) <= DBL_MAX; }   // false for NaN and +-inf

// Python slice [lo:hi] on a sequence of length n -> [start, stop)
MOTES_HD void python_slice(int lo, int hi, int n, int* start, int* stop) {
    int a = lo < 0 ? (lo + n < 0 ? 0 : lo + n) : (lo > n ? n : lo);
    int b = hi < 0 ? (hi + n < 0 ? 0 : hi + n) : (hi > n ? n : hi);
    if (b < a) b = a;
    *start = a;
    *stop = b;
}

// Does the median (np.median over n values, none NaN) equal `thr` exactly?  Decided from counts
// and the two neighbours of thr instead of a sort:  lt / eq = number of values < thr / == thr,
// max_lt = largest value below thr, min_gt = smallest value above (only read when they decide).
MOTES_HD bool median_equals(double thr, int n, int lt, int eq, double max_lt, double min_gt) {
    if (n <= 0) return false;
    int k1 = (n - 1) / 2, k2 = n / 2;
    // class of an order statistic: -1 below thr, 0 equal, +1 above
    int c1 = (k1 < lt) ? -1 : (k1 < lt + eq ? 0 : 1);
    int c2 = (k2 < lt) ? -1 : (k2 < lt + eq ? 0 : 1);
    if (c1 == 0 && c2 == 0) return true;
    if (c1 == c2) return false;                 // both strictly on one side: the mean stays there
    // the two middle values straddle: they are exactly the neighbours of the boundary
    double a = (c1 == -1) ? max_lt : thr;
    double b = (c2 == 1) ? min_gt : thr;
    return (a + b) * 0.5 == thr;
}

// Moffat profile on a linear background, same operation order as common.py:89-92
// (the library is compiled with -fmad=false so nothing here is contracted).
MOTES_HD double moffat_unit(double r, double c, double alpha, double beta) {
    double off = r - c;
    double q = (off * off) / (alpha * alpha);
    return pow(1.0 + q, -beta);
}

MOTES_HD double moffat_value(double unit, double r, double amp, double level, double grad) {
    return (amp * unit + level) + (r * grad);
}

// common.py:112
MOTES_HD double moffat_fwhm(double alpha, double beta) {
    return (2.0 * alpha) * sqrt(pow(2.0, 1.0 / beta) - 1.0);
}

}  // namespace motes
