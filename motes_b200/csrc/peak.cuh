// FP64 roof of the device this library runs on: a DFMA micro-benchmark (SURVEY.md §8d asks for the fit kernels'
// FP64 fraction against a MEASURED peak; MEASURED_PEAKS.json has no FP64 entry).
#pragma once

namespace motes {

constexpr int kPeakChains = 8;        // independent dependency chains per thread (DFMA latency ~ 8 issue slots)
constexpr int kPeakThreads = 256;

__global__ void __launch_bounds__(kPeakThreads)
fp64_peak_kernel(double* __restrict__ sink, int iters, double a, double b) {
    double x[kPeakChains];
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) x[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < kPeakChains; ++k) x[k] = fma(x[k], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k) s += x[k];
    if (s == 123456.789) sink[0] = s;                 // never true: keeps the chains alive
}

}  // namespace motes
