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
    __device__ __forceinline__ int all(int p) const { return __all_sync(0xffffffffu, p); }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};
#endif

MOTES_HD bool is_finite(double x) { return fabs(x) <= DBL_MAX; }   // false for NaN and +-inf

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
