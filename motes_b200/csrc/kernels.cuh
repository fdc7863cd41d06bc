// __global__ kernels of libmotes_b200 (sm_100a).  Every kernel handles a batch of frames;
// blockIdx.y (or a flattened work index) selects the frame.  Frames are (nspat, ndisp) float64,
// dispersion fastest, `frame_stride` elements apart.
#pragma once

#include "common.cuh"
#include "select.cuh"
#include "trfb.cuh"
#include "bins.cuh"
#include "limits.cuh"
#include "sky.cuh"
#include "extract.cuh"

namespace motes {

// Per-frame scalars that flow between the stages (device resident).
struct FrameState {
    double seeing;          // header seeing handed to the first fit [arcsec] (gmosio.py:118-126)
    double pixres;          // arcsec / pixel
    double scale;           // data_scaling_factor (tracing.py:435-437)
    double seeing_px;       // FWHM of the median-profile fit in pixels (tracing.py:449)
    double alpha0_first;    // seeing / pixres
    double alpha0_bins;     // seeing_px / pixres  (quirk: tracing.py:449,580)
    double median_params[6];
    int row_lo, row_hi;     // int(floor(c - 3 FWHM)), int(ceil(c + 3 FWHM))  (core.py:134-135)
    int status;             // 0, MOTES_E_FIT or MOTES_E_NOSKY
    int pad;
};

// ---------------------------------------------------------------------------------------------
// a1: np.nanmedian(data, axis=1) (core.py:80).  One CTA per (row, frame).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
row_median_kernel(const double* __restrict__ data, size_t frame_stride, int nspat, int ndisp,
                  double* __restrict__ out /* [F][nspat] */) {
    extern __shared__ __align__(16) unsigned char rm_smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(rm_smem);
    uint32_t* hist = reinterpret_cast<uint32_t*>(keys + ndisp);
    unsigned long long* cell = reinterpret_cast<unsigned long long*>(hist + 256);
    int* icell = reinterpret_cast<int*>(cell + 1);
    BlockScope sc{(int)threadIdx.x};
    const int row = blockIdx.x, f = blockIdx.y;
    const double* src = data + (size_t)f * frame_stride + (size_t)row * ndisp;
    int nan = 0;
    for (int i = threadIdx.x; i < ndisp; i += blockDim.x) {
        double v = src[i];
        nan += (v != v);
        keys[i] = key_or_excluded(v);
    }
    int n_nan = sc.isum(nan, icell);
    double med = median_of(sc, ArrayKeys{keys}, ndisp, ndisp - n_nan, hist, cell);
    if (threadIdx.x == 0) out[(size_t)f * nspat + row] = med;
}

// The same for dispersion axes too long to stage a row in shared memory (beyond ~28 k columns at 227 KB): the
// radix selection reads the row from global memory in every pass (a row is a few hundred KB: it stays in L2).
__global__ void __launch_bounds__(256)
row_median_long_kernel(const double* __restrict__ data, size_t frame_stride, int nspat, int ndisp,
                       double* __restrict__ out /* [F][nspat] */) {
    __shared__ uint32_t hist[256];
    __shared__ unsigned long long cell;
    __shared__ int icell;
    BlockScope sc{(int)threadIdx.x};
    const int row = blockIdx.x, f = blockIdx.y;
    const double* src = data + (size_t)f * frame_stride + (size_t)row * ndisp;
    int nan = 0;
    for (int i = threadIdx.x; i < ndisp; i += blockDim.x) nan += (src[i] != src[i]);
    int n_nan = sc.isum(nan, &icell);
    struct RowKeys {
        const double* p;
        __device__ unsigned long long operator()(int i) const { return key_or_excluded(p[i]); }
    };
    double med = median_of(sc, RowKeys{src}, ndisp, ndisp - n_nan, hist, &cell);
    if (threadIdx.x == 0) out[(size_t)f * nspat + row] = med;
}

// exact 10^k for the scale factor (10 ** np.abs(np.floor(...)), tracing.py:437)
__device__ __forceinline__ double pow10_int(double k) {
    if (k >= 0.0 && k <= 22.0) {
        double r = 1.0;
        for (int i = 0; i < (int)k; ++i) r *= 10.0;
        return r;
    }
    return pow(10.0, k);
}

// a2 (first half): scale factor of the median profile; one warp per frame.
__global__ void __launch_bounds__(32)
frame_prep_kernel(const double* __restrict__ medprof, int nspat, FrameState* __restrict__ st) {
    __shared__ uint32_t hist[256];
    __shared__ unsigned long long cell;
    WarpScope sc{(int)threadIdx.x};
    const int f = blockIdx.x;
    const double* prof = medprof + (size_t)f * nspat;
    int nan = 0;
    for (int i = threadIdx.x; i < nspat; i += 32) nan += (prof[i] != prof[i]);
    int n_nan = sc.isum(nan, nullptr);
    struct ProfKeys {
        const double* p;
        __device__ unsigned long long operator()(int i) const { return key_or_excluded(p[i]); }
    };
    double med = median_of(sc, ProfKeys{prof}, nspat, nspat - n_nan, hist, &cell);
    if (threadIdx.x == 0) {
        FrameState& s = st[f];
        s.scale = pow10_int(fabs(floor(log10(fabs(med)))));
        s.alpha0_first = s.seeing / s.pixres;
        s.status = 0;
    }
}

// a2 (second half) + a7: after the median-profile fit.  One thread per frame.
__global__ void frame_post_kernel(int nframes, const double* __restrict__ fit_params, const int32_t* __restrict__ fit_status,
                                  FrameState* __restrict__ st) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    FrameState& s = st[f];
    const double* p = fit_params + (size_t)f * kParams;
    if (fit_status[f] < 0) s.status = MOTES_E_FIT;
    double fwhm = moffat_fwhm(p[2], p[3]);
    s.seeing_px = fwhm;                         // header_parameters["seeing"] is now in pixels
    s.alpha0_bins = fwhm / s.pixres;
    for (int k = 0; k < kParams; ++k) s.median_params[k] = p[k];
    s.median_params[0] = p[0] / s.scale;
    s.median_params[4] = p[4] / s.scale;
    s.median_params[5] = p[5] / s.scale;
    double lo = p[1] - (3.0 * fwhm), hi = p[1] + (3.0 * fwhm);
    // a frame whose first fit could not start (the reference raises there) runs no further stage: every later
    // kernel tests status; (int) of a non-finite limit would be undefined
    const bool usable = s.status == 0 && is_finite(lo) && is_finite(hi) && fabs(lo) < 1e9 && fabs(hi) < 1e9;
    if (!usable && s.status == 0) s.status = MOTES_E_FIT;
    s.row_lo = usable ? (int)floor(lo) : 0;
    s.row_hi = usable ? (int)ceil(hi) : 0;
}

// ---------------------------------------------------------------------------------------------
// a5 / a6 support: per-column sums over the binning rows and the chip-gap flag.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
column_stats_kernel(const double* __restrict__ data, const double* __restrict__ errs, size_t frame_stride, int nspat,
                    int ndisp, const FrameState* __restrict__ st, double* __restrict__ S, double* __restrict__ V,
                    uint8_t* __restrict__ gap) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (c >= ndisp) return;
    ColumnStats cs = column_stats(data + (size_t)f * frame_stride, errs + (size_t)f * frame_stride, nspat, ndisp, c,
                                  st[f].row_lo, st[f].row_hi);
    S[(size_t)f * ndisp + c] = cs.sum_data;
    V[(size_t)f * ndisp + c] = cs.sum_var;
    gap[(size_t)f * ndisp + c] = cs.gap;
}

// per column: does every row carry qual == 1?  (thread per column, coalesced over the columns)
__global__ void __launch_bounds__(128)
qual_columns_kernel(const uint8_t* __restrict__ qual, size_t qual_stride, int nspat, int ndisp, uint8_t* __restrict__ allgood) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (c >= ndisp) return;
    const uint8_t* q = qual + (size_t)f * qual_stride + c;
    int good = 1;
    for (int r = 0; r < nspat; ++r) good &= (q[(size_t)r * ndisp] == 1);
    allgood[(size_t)f * ndisp + c] = (uint8_t)good;
}

// a5: greedy S/N binning, one CTA per frame (tracing.py:114-254).
constexpr int kBinWindow = 1024;       // columns staged in shared memory at a time (bins.cuh: BinWindow)
__global__ void __launch_bounds__(1024)
bins_kernel(const uint8_t* __restrict__ qual, size_t qual_stride, const uint8_t* __restrict__ allgood, const double* __restrict__ S,
            const double* __restrict__ V, int nspat, int ndisp, double snr_limit, double min_cols, int cap,
            int32_t* __restrict__ bins_i /* [F][2*cap] */, double* __restrict__ bins_snr /* [F][cap] */,
            int32_t* __restrict__ tmp_i, double* __restrict__ tmp_snr, int32_t* __restrict__ nbins,
            const FrameState* __restrict__ st) {
    extern __shared__ __align__(16) int bins_counts[];      // nspat counters (padded to 16 bytes), then the column window
    __shared__ unsigned long long cell;
    BlockScope sc{(int)threadIdx.x};
    const int f = blockIdx.x;
    if (st && st[f].status != 0) {                          // failed frame: no bins, hence no fits, limits or spectra
        if (threadIdx.x == 0) nbins[f] = 0;
        return;
    }
    BinWindow win;
    win.S = reinterpret_cast<double*>(bins_counts + ((nspat + 3) & ~3));
    win.V = win.S + kBinWindow;
    win.xfer = win.V + kBinWindow;
    win.good = reinterpret_cast<uint8_t*>(win.xfer + 10);
    win.cap = kBinWindow;
    int nb = snr_bins_scan(sc, qual + (size_t)f * qual_stride, S + (size_t)f * ndisp, V + (size_t)f * ndisp, nspat,
                           ndisp, snr_limit, min_cols, bins_counts, bins_i + (size_t)f * 2 * cap,
                           bins_snr + (size_t)f * cap, tmp_i + (size_t)f * 2 * cap, tmp_snr + (size_t)f * cap,
                           allgood + (size_t)f * ndisp, &cell, &win);
    if (threadIdx.x == 0) nbins[f] = nb;
}

// a6 (first half): median spatial profile of every bin, leaving out chip-gap columns
// (tracing.py:505-507).  One warp per (frame, bin, row).  Bins of up to kProfStage columns are
// staged in shared memory and the median is found by rank counting (w^2/32 compares per lane:
// 8 for the usual 16-column bins); wider bins fall back to radix selection straight from the frame.
constexpr int kProfWarps = 8;
constexpr int kProfStage = 512;
constexpr int kProfFast = 16;                   // bin width with a thread-per-median kernel (the default binning)

// Bins of exactly 16 columns (what get_bins produces on any source bright enough for the default S/N
// limits): one THREAD per (bin, row) median - 16 keys in registers, a bitonic network, the middle one or
// two valid keys.  Adjacent threads take adjacent bins of one row, i.e. adjacent 128-byte runs of the frame.
__global__ void __launch_bounds__(128)
bin_profile16_kernel(const double* __restrict__ data, size_t frame_stride, int nspat, int ndisp,
                     const uint8_t* __restrict__ gap, const int32_t* __restrict__ bins_i, const int32_t* __restrict__ nbins,
                     int cap, double* __restrict__ profiles /* [F][cap][nspat] */) {
    const int f = blockIdx.y;
    const int nb = nbins[f];
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= (long long)nb * nspat) return;
    const int row = (int)(item / nb), b = (int)(item % nb);
    const int first = bins_i[((size_t)f * cap + b) * 2], last = bins_i[((size_t)f * cap + b) * 2 + 1];
    if (last - first != kProfFast) return;
    const double* p = data + (size_t)f * frame_stride + (size_t)row * ndisp + first;
    const uint8_t* g = gap + (size_t)f * ndisp + first;
    unsigned long long k[kProfFast];
    int n = 0;
#pragma unroll
    for (int i = 0; i < kProfFast; ++i) {
        k[i] = g[i] ? kKeyExcluded : key_or_excluded(p[i]);
        n += (k[i] != kKeyExcluded);
    }
#pragma unroll
    for (int size = 2; size <= kProfFast; size <<= 1) {
#pragma unroll
        for (int d = size >> 1; d > 0; d >>= 1) {
#pragma unroll
            for (int i = 0; i < kProfFast; ++i) {
                const int j = i ^ d;
                if (j > i) {
                    const bool up = (i & size) == 0;
                    const unsigned long long a = k[i], c = k[j];
                    const bool sw = (a > c) == up;
                    k[i] = sw ? c : a;
                    k[j] = sw ? a : c;
                }
            }
        }
    }
    double med = NAN;
    if (n > 0) {
        const int k1 = (n - 1) / 2, k2 = n / 2;
        unsigned long long a1 = 0, a2 = 0;
#pragma unroll
        for (int i = 0; i < kProfFast; ++i) {
            if (i == k1) a1 = k[i];
            if (i == k2) a2 = k[i];
        }
        med = (k1 == k2) ? value_of(a1) : (value_of(a1) + value_of(a2)) * 0.5;
    }
    profiles[((size_t)f * cap + b) * nspat + row] = med;
}

__global__ void __launch_bounds__(kProfWarps * 32)
bin_profile_kernel(const double* __restrict__ data, size_t frame_stride, int nspat, int ndisp,
                   const uint8_t* __restrict__ gap, const int32_t* __restrict__ bins_i, const int32_t* __restrict__ nbins,
                   int cap, double* __restrict__ profiles /* [F][cap][nspat] */) {
    __shared__ unsigned long long staged[kProfWarps][kProfStage];
    __shared__ uint32_t hist_all[kProfWarps][256];
    __shared__ unsigned long long cells[kProfWarps];
    const int warp = threadIdx.x >> 5;
    WarpScope sc{(int)(threadIdx.x & 31)};
    const int f = blockIdx.y;
    const int b = blockIdx.x;                   // one CTA per bin, its warps stride over the rows
    if (b >= nbins[f]) return;
    const int first = bins_i[((size_t)f * cap + b) * 2], last = bins_i[((size_t)f * cap + b) * 2 + 1];
    const int w = last - first;
    if (w == kProfFast) return;                 // done by bin_profile16_kernel
    struct BinKeys {
        const double* p;
        const uint8_t* g;
        __device__ unsigned long long operator()(int i) const { return g[i] ? kKeyExcluded : key_or_excluded(p[i]); }
    };
    for (int row = warp; row < nspat; row += kProfWarps) {
    BinKeys keys{data + (size_t)f * frame_stride + (size_t)row * ndisp + first, gap + (size_t)f * ndisp + first};
    double med;
    if (w <= kProfStage) {
        unsigned long long* sk = staged[warp];
        int valid = 0;
        for (int i = sc.tid; i < w; i += 32) {
            unsigned long long k = keys(i);
            sk[i] = k;
            valid += (k != kKeyExcluded);
        }
        const int n = sc.isum(valid, nullptr);
        __syncwarp();
        const int k1 = (n - 1) / 2, k2 = n / 2;
        unsigned long long a1 = kKeyExcluded, a2 = kKeyExcluded;
        for (int e = sc.tid; e < w; e += 32) {
            const unsigned long long k = sk[e];
            if (k == kKeyExcluded) continue;
            int rank = 0;
            for (int j = 0; j < w; ++j) {
                const unsigned long long o = sk[j];
                rank += (o < k) || (o == k && j < e);
            }
            if (rank == k1) a1 = k;
            if (rank == k2) a2 = k;
        }
        a1 = sc.umin(a1, nullptr);
        a2 = sc.umin(a2, nullptr);
        med = (n <= 0) ? NAN : (k1 == k2 ? value_of(a1) : (value_of(a1) + value_of(a2)) * 0.5);
    } else {
        int valid = 0;
        for (int i = sc.tid; i < w; i += 32) valid += (keys(i) != kKeyExcluded);
        int n_valid = sc.isum(valid, nullptr);
        med = median_of(sc, keys, w, n_valid, hist_all[warp], &cells[warp]);
    }
    if (sc.tid == 0) profiles[((size_t)f * cap + b) * nspat + row] = med;
    __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// a3: one warp per Moffat fit; slab in shared memory; work handed out by an atomic counter.
// Work item i -> frame i / bound, slot i % bound; slots >= nbins[frame] are skipped.
// ---------------------------------------------------------------------------------------------
constexpr int kFitMaxWarps = 8;

struct FitBatch {
    const double* profiles;     // [(frame * cap + slot) * m]
    int m;
    int nframes, bound, cap;
    const int32_t* nbins;       // per frame, or nullptr: every slot < bound is a fit
    const double* alpha0;       // per fit [(frame*cap+slot)] or per frame when alpha0_frame_stride > 0
    size_t alpha0_frame_stride; // in bytes between frames when alpha0 lives inside FrameState; 0 = per-fit array
    const double* scale;        // per frame (same addressing as alpha0), or nullptr = 1
    size_t scale_frame_stride;
    double* params;             // [(frame*cap+slot) * 6]
    double* cost;
    int32_t* nfev;
    int32_t* njev;
    int32_t* status;
};

__global__ void __launch_bounds__(kFitMaxWarps * 32)
moffat_fit_kernel(FitBatch fb, unsigned int* work_counter) {
    extern __shared__ __align__(16) double fit_smem[];
    const int warp = threadIdx.x >> 5;
    WarpCtx ctx{(int)(threadIdx.x & 31)};
    const int m = fb.m;
    const size_t per_warp = fit_scratch_doubles(m);
    FitScratch ws = carve_fit_scratch(fit_smem + per_warp * warp, m);
    const unsigned int total = (unsigned int)fb.nframes * (unsigned int)fb.bound;
    while (true) {
        unsigned int item = 0;
        if (ctx.lane == 0) item = atomicAdd(work_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        const int f = item / fb.bound, slot = item % fb.bound;
        if (fb.nbins && slot >= fb.nbins[f]) continue;
        const size_t idx = (size_t)f * fb.cap + slot;
        double alpha0 = fb.alpha0_frame_stride
                            ? *reinterpret_cast<const double*>(reinterpret_cast<const char*>(fb.alpha0) + f * fb.alpha0_frame_stride)
                            : fb.alpha0[idx];
        double scale = 1.0;
        if (fb.scale)
            scale = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(fb.scale) + f * fb.scale_frame_stride);
        FitResult r;
        trf_fit_moffat(ctx, m, fb.profiles + idx * m, scale, alpha0, ws, &r);
        if (ctx.lane == 0) {
            for (int k = 0; k < kParams; ++k) fb.params[idx * kParams + k] = r.x[k];
            if (fb.cost) fb.cost[idx] = r.cost;
            if (fb.nfev) fb.nfev[idx] = r.nfev;
            if (fb.njev) fb.njev[idx] = r.njev;
            if (fb.status) fb.status[idx] = r.status;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// a3, batched (trfb.cuh): a CTA takes a group of up to 32 fits from a queue and advances them in
// lock step: sweep phase - the whole CTA works on one fit at a time (residual sweep at the trial
// point, accept / reject / stop, re-linearisation); step phase - thread j chooses the next trial
// step of fit j (SVD of the 6 x 6 triangle, trust-region step, step selection), so the
// six-parameter logic runs once per fit instead of once per lane.  The states of the group live in
// shared memory; nothing but the profiles is read from and the results written to HBM.
// ---------------------------------------------------------------------------------------------
constexpr int kFitGroupMax = 32;
constexpr int kStatusRetired = -200;          // finished (results written) or empty slot

__device__ __forceinline__ void fit_finalize(const FitBatch& fb, size_t idx, const FitState& st) {
    for (int k = 0; k < kParams; ++k) fb.params[idx * kParams + k] = st.x[k];
    if (fb.cost) fb.cost[idx] = st.cost;
    if (fb.nfev) fb.nfev[idx] = st.nfev;
    if (fb.njev) fb.njev[idx] = st.njev;
    if (fb.status) fb.status[idx] = (st.status == kStatusRunning) ? (int)kFitMaxNfev : st.status;
}

__device__ __forceinline__ double fit_frame_scalar(const double* base, size_t stride, int f, size_t idx) {
    return stride ? *reinterpret_cast<const double*>(reinterpret_cast<const char*>(base) + f * stride) : base[idx];
}

// Threads that cooperate on one fit, by profile length - the SAME choice in the group kernel (small batches) and in the
// lock-step rounds (large batches), so that a fit's arithmetic (the order of its reductions) does not depend on the
// size of the batch it arrives in.  Measured on B200 (32.9 k fits, lock-step rounds): one warp wins up to ~224 points
// (80 points: 5.1 ms against 7.4 / 10.5 ms with two / four warps), two warps up to ~576 (400 points: 11.6 against 15.1
// ms with four warps and 15.9 with one), four warps beyond.
constexpr int kFitWarpMax = 224, kFitPairMax = 576;
MOTES_HD int fit_threads_for(int m) { return m <= kFitWarpMax ? 32 : m <= kFitPairMax ? 64 : 128; }

template <int THREADS>
struct FitGroupShape {
    static constexpr int kTeams = THREADS / 8;                 // 8-lane teams of the step phase
    static constexpr int kStepScratch = kTeams * 36 + 48;      // V per team, an identity R and a zero z for idle teams
    static size_t smem(int m, int group) {
        return (fit_sweep_doubles(m) + BlockCtx<THREADS>::kRedDoubles + kStepScratch) * sizeof(double) +
               (size_t)group * sizeof(FitState);
    }
};
// out of line: the step logic is long, runs once per round and must not bloat the sweep loop's registers
__device__ __noinline__ bool fit_propose_call(int m, FitState* st) { return fit_propose(m, *st); }

template <bool TEAM_STEP, int THREADS>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
moffat_fit_group_kernel(FitBatch fb, int group, unsigned int* __restrict__ queue) {
    extern __shared__ __align__(16) double fitb_smem[];
    __shared__ unsigned int s_item;
    using FitBlockCtx = BlockCtx<THREADS>;
    constexpr int kFitTeams = FitGroupShape<THREADS>::kTeams, kFitStepScratch = FitGroupShape<THREADS>::kStepScratch;
    constexpr int kFitThreads = THREADS;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m = fb.m;
    double* slab = fitb_smem;
    FitBlockCtx ctx{tid, fitb_smem + fit_sweep_doubles(m), 0};
    double* vscratch = fitb_smem + fit_sweep_doubles(m) + FitBlockCtx::kRedDoubles;
    double* idle_R = vscratch + kFitTeams * 36;          // identity; idle_z = idle_R + 36 (zeros)
    FitState* states = reinterpret_cast<FitState*>(vscratch + kFitStepScratch);
    for (int i = tid; i < 48; i += THREADS) idle_R[i] = (i < 36 && i % 7 == 0) ? 1.0 : 0.0;
    const int lane = tid & 31, team = tid >> 3, sub = tid & 7;
    const unsigned int total = (unsigned int)fb.nframes * (unsigned int)fb.bound;
    while (true) {
        if (tid == 0) s_item = atomicAdd(queue, 1u);
        __syncthreads();
        const unsigned int base = s_item * (unsigned int)group;
        __syncthreads();
        if (base >= total) break;
        const int count = (int)min((unsigned int)group, total - base);
        // ---- start points and bounds: one warp per fit
        for (int j = warp; j < count; j += kFitThreads / 32) {
            const unsigned int item = base + j;
            const int f = item / fb.bound, slot = item % fb.bound;
            if (fb.nbins && slot >= fb.nbins[f]) {
                states[j].status = kStatusRetired;
                continue;
            }
            const size_t idx = (size_t)f * fb.cap + slot;
            const double alpha0 = fit_frame_scalar(fb.alpha0, fb.alpha0_frame_stride, f, idx);
            const double scale = fb.scale ? fit_frame_scalar(fb.scale, fb.scale_frame_stride, f, 0) : 1.0;
            fit_start_warp(lane, m, fb.profiles + idx * m, scale, alpha0, &states[j]);
        }
        __syncthreads();
        while (true) {
            // ---- sweep phase
            for (int j = 0; j < count; ++j) {
                if (states[j].status != kStatusRunning) continue;              // uniform
                const unsigned int item = base + j;
                const int f = item / fb.bound, slot = item % fb.bound;
                const size_t idx = (size_t)f * fb.cap + slot;
                const double scale = fb.scale ? fit_frame_scalar(fb.scale, fb.scale_frame_stride, f, 0) : 1.0;
                fit_evaluate(ctx, m, fb.profiles + idx * m, scale, slab, &states[j]);
            }
            __syncthreads();
            // ---- step phase.  Small groups (latency matters): 8-lane teams, one fit each; the SVD of the 6 x 6
            // triangle is spread over the team (jacobi_svd6_warp), the scalar logic around it is replicated in
            // the team and stored by its lane 0.  Large groups (throughput matters): one thread per fit.
            int running = 0;
            if constexpr (TEAM_STEP) {
                for (int first_fit = 0; first_fit < count; first_fit += kFitTeams) {
                    const int j = first_fit + team;
                    const bool have = j < count && states[j].status != kStatusRetired;
                    FitState* st = &states[have ? j : 0];
                    ProposeVars pv;
                    const bool go = have && propose_begin(st, pv, sub == 0, st);
                    double sv[kParams], suf[kParams];
                    double* V = vscratch + team * 36;
                    jacobi_svd6_warp(lane, go ? st->R : idle_R, go ? st->z : idle_R + 36, sv, suf, V, false);
                    if (go) {
                        propose_finish(m, st, pv, sv, suf, V, sub == 0, st);
                        running = 1;
                    } else if (have && sub == 0) {
                        const unsigned int item = base + j;
                        const int f = item / fb.bound, slot = item % fb.bound;
                        fit_finalize(fb, (size_t)f * fb.cap + slot, *st);
                        st->status = kStatusRetired;
                    }
                    __syncwarp();
                }
            } else if (tid < count && states[tid].status != kStatusRetired) {
                (void)lane; (void)team; (void)sub;
                if (fit_propose_call(m, &states[tid])) {
                    running = 1;
                } else {
                    const unsigned int item = base + tid;
                    const int f = item / fb.bound, slot = item % fb.bound;
                    fit_finalize(fb, (size_t)f * fb.cap + slot, states[tid]);
                    states[tid].status = kStatusRetired;
                }
            }
            if (!__syncthreads_or(running)) break;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// a3, lock-step rounds (large batches).  The group kernel above keeps a CTA's warps idle while one thread per fit
// chooses the next step (ncu, round 1: 17 % of all stall samples on that barrier) and holds its group's states in
// shared memory, which caps the fits in flight per SM at four.  Here the two phases are separate kernels over ALL
// running fits of the batch, with the states in global memory (L2 resident: 648 B per fit):
//
//   fit_init_kernel    start point and bounds, one warp per fit; running fits enter list 0
//   fit_sweep_kernel   one small CTA per running fit: fit_evaluate (residual sweep, accept / reject / stop,
//                      Jacobian, QR) - no step-phase bubble, only the 7 m slab in shared memory
//   fit_step_kernel    one THREAD per running fit, full warps: fit_propose; finished fits are written out,
//                      the others enter the next round's list
//   fit_finish_kernel  after kFitRounds rounds: the few fits still running, one CTA each, to the end
//
// Every fit sees exactly the arithmetic of the group kernel; the order of the lists does not enter any result.
// ---------------------------------------------------------------------------------------------
constexpr int kFitRoundsMax = 64;
struct FitRounds {                                   // zeroed before every stage_fit call
    unsigned int queue[kFitRoundsMax + 2];           // work counter of the sweep (finish) kernel of round r
    unsigned int count[kFitRoundsMax + 2];           // fits running at the start of round r
};

__global__ void __launch_bounds__(256)
fit_init_kernel(FitBatch fb, FitState* __restrict__ states, unsigned int* __restrict__ list, FitRounds* __restrict__ rounds) {
    const int lane = threadIdx.x & 31;
    const unsigned int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned int total = (unsigned int)fb.nframes * (unsigned int)fb.bound;
    if (item >= total) return;
    const int f = item / fb.bound, slot = item % fb.bound;
    if (fb.nbins && slot >= fb.nbins[f]) return;
    const size_t idx = (size_t)f * fb.cap + slot;
    const double alpha0 = fit_frame_scalar(fb.alpha0, fb.alpha0_frame_stride, f, idx);
    const double scale = fb.scale ? fit_frame_scalar(fb.scale, fb.scale_frame_stride, f, 0) : 1.0;
    FitState* st = &states[item];
    fit_start_warp(lane, fb.m, fb.profiles + idx * fb.m, scale, alpha0, st);
    __syncwarp();
    if (lane == 0) {
        if (st->status == kStatusRunning) list[atomicAdd(&rounds->count[0], 1u)] = item;
        else fit_finalize(fb, idx, *st);
    }
}

#ifndef MOTES_FIT_SWEEP_WARPS
#define MOTES_FIT_SWEEP_WARPS 16
#endif
template <int THREADS>
__global__ void __launch_bounds__(THREADS, MOTES_FIT_SWEEP_WARPS * 32 / THREADS)
fit_sweep_kernel(FitBatch fb, FitState* __restrict__ states, const unsigned int* __restrict__ list,
                 FitRounds* __restrict__ rounds, int round) {
    extern __shared__ __align__(16) double fits_smem[];
    __shared__ unsigned int s_item;
    using Ctx = BlockCtx<THREADS>;
    const int tid = threadIdx.x, m = fb.m;
    Ctx ctx{tid, fits_smem + fit_sweep_doubles(m), 0};
    const unsigned int n = rounds->count[round];
    while (true) {
        if (tid == 0) s_item = atomicAdd(&rounds->queue[round], 1u);
        __syncthreads();
        const unsigned int k = s_item;
        __syncthreads();
        if (k >= n) break;
        const unsigned int item = list[k];
        const int f = item / fb.bound, slot = item % fb.bound;
        const size_t idx = (size_t)f * fb.cap + slot;
        const double scale = fb.scale ? fit_frame_scalar(fb.scale, fb.scale_frame_stride, f, 0) : 1.0;
        fit_evaluate(ctx, m, fb.profiles + idx * m, scale, fits_smem, &states[item]);
    }
}

__global__ void __launch_bounds__(128)
fit_step_kernel(FitBatch fb, FitState* __restrict__ states, const unsigned int* __restrict__ list,
                unsigned int* __restrict__ list_next, FitRounds* __restrict__ rounds, int round) {
    const unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rounds->count[round]) return;
    const unsigned int item = list[t];
    FitState* st = &states[item];
    if (fit_propose_call(fb.m, st)) {
        list_next[atomicAdd(&rounds->count[round + 1], 1u)] = item;
    } else {
        const int f = item / fb.bound, slot = item % fb.bound;
        fit_finalize(fb, (size_t)f * fb.cap + slot, *st);
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
fit_finish_kernel(FitBatch fb, FitState* __restrict__ states, const unsigned int* __restrict__ list,
                  FitRounds* __restrict__ rounds, int round) {
    extern __shared__ __align__(16) double fitf_smem[];
    __shared__ unsigned int s_item;
    using Ctx = BlockCtx<THREADS>;
    const int tid = threadIdx.x, m = fb.m;
    Ctx ctx{tid, fitf_smem + fit_sweep_doubles(m), 0};
    const unsigned int n = rounds->count[round];
    while (true) {
        if (tid == 0) s_item = atomicAdd(&rounds->queue[round], 1u);
        __syncthreads();
        const unsigned int k = s_item;
        __syncthreads();
        if (k >= n) break;
        const unsigned int item = list[k];
        const int f = item / fb.bound, slot = item % fb.bound;
        const size_t idx = (size_t)f * fb.cap + slot;
        const double scale = fb.scale ? fit_frame_scalar(fb.scale, fb.scale_frame_stride, f, 0) : 1.0;
        FitState* st = &states[item];
        while (true) {
            fit_evaluate(ctx, m, fb.profiles + idx * m, scale, fitf_smem, st);
            int running = 0;
            if (tid == 0) running = fit_propose_call(m, st) ? 1 : 0;
            if (!__syncthreads_or(running)) break;
        }
        if (tid == 0) fit_finalize(fb, idx, *st);
    }
}

// a6 (second half): un-scale, extraction limits, limit table row and the 8-vector of a bin
// (tracing.py:517-545).  One thread per (frame, slot).
__global__ void bin_post_kernel(int nframes, int bound, int cap, const int32_t* __restrict__ nbins,
                                const int32_t* __restrict__ bins_i, const double* __restrict__ fit_params,
                                const int32_t* __restrict__ fit_status, FrameState* __restrict__ st,
                                double fwhm_multiplier, int wavelength_start,
                                double* __restrict__ params8 /* [F][cap][8] */, double* __restrict__ table /* [F][4][cap] */) {
    int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= nframes * bound) return;
    int f = item / bound, b = item % bound;
    if (b >= nbins[f]) return;
    size_t idx = (size_t)f * cap + b;
    const double* p = fit_params + idx * kParams;
    if (fit_status[idx] < 0) atomicMin(&st[f].status, (int)MOTES_E_FIT);
    const double scale = st[f].scale;
    double q[kParams];
    for (int k = 0; k < kParams; ++k) q[k] = p[k];
    q[0] = p[0] / scale;
    q[4] = p[4] / scale;
    q[5] = p[5] / scale;
    double fwhm = moffat_fwhm(q[2], q[3]);
    int first = bins_i[idx * 2], last = bins_i[idx * 2 + 1];
    double* t = table + (size_t)f * 4 * cap;
    t[b] = (double)(first + last) * 0.5;
    t[cap + b] = q[1] - (fwhm_multiplier * fwhm);
    t[2 * cap + b] = q[1] + (fwhm_multiplier * fwhm);
    t[3 * cap + b] = q[1];
    double* o = params8 + idx * 8;
    for (int k = 0; k < kParams; ++k) o[k] = q[k];
    o[6] = (double)(first + wavelength_start);
    o[7] = (double)(last + wavelength_start);
}

// a8: limits for every dispersion pixel; one CTA per frame.
__global__ void __launch_bounds__(256)
limits_kernel(const double* __restrict__ table, int cap, const int32_t* __restrict__ nbins, int ndisp,
              double* __restrict__ limits /* [F][2][ndisp] */, double* __restrict__ inner /* [F][2*(ndisp+2)] */) {
    __shared__ unsigned long long keys[160];
    __shared__ uint32_t hist[256];
    __shared__ unsigned long long cell;
    __shared__ int icell;
    BlockScope sc{(int)threadIdx.x};
    const int f = blockIdx.x;
    if (nbins[f] < 1) return;                               // failed frame (bins_kernel)
    interpolate_limits(sc, table + (size_t)f * 4 * cap, cap, nbins[f], ndisp, limits + (size_t)f * 2 * ndisp,
                       limits + (size_t)f * 2 * ndisp + ndisp, inner + (size_t)f * 2 * (ndisp + 2), keys, hist, &cell,
                       &icell);
}

// ---------------------------------------------------------------------------------------------
// a9: sky.  Tile geometry shared with the extraction kernel: a CTA stages kTileCols consecutive
// columns through shared memory with coalesced row-segment loads; row stride is odd in doubles so
// a warp walking down a column is bank-conflict free.
// ---------------------------------------------------------------------------------------------
constexpr int kTileCols = 16;
constexpr int kTileStride = kTileCols + 1;
constexpr int kTileThreads = 256;
constexpr int kTileWarps = kTileThreads / 32;

// Warp-level bitonic sort of EPL * 32 (key, row) pairs held in registers (position = lane * EPL + k),
// ascending by key - the order np.sort gives the sky pixels -, equal keys by row (the order in which the
// rank-by-counting of sky_prepare_column breaks ties).
#ifndef MOTES_PREP_CTAS
#define MOTES_PREP_CTAS 4
#endif
// (key, row) pairs compared as one 96-bit number.  The row makes the order STRICT, which
// the exchange across lanes relies on: both lanes evaluate "mine > partner's" and exactly one of them sees true, so
// exactly one pair moves each way.  Compared by key alone, two EQUAL keys (quantised counts: float32 or integer
// detectors) made both lanes see false, the lane keeping the larger pair took its partner's, and one row was
// duplicated while the other was lost (found by the float32 end-to-end check of bench.py;
// tests/test_gpu_vs_oracle.py::test_equal_sky_values_keep_every_pixel).
// All ones when (ka, ra) > (kb, rb): the borrow of (kb, rb) - (ka, ra) through a three-word subtract-with-carry chain.
__device__ __forceinline__ uint32_t pair_gt_mask(unsigned long long ka, uint32_t ra, unsigned long long kb, uint32_t rb) {
    uint32_t m;
    asm("{\n\t.reg .u32 t;\n\tsub.cc.u32 t, %6, %3;\n\tsubc.cc.u32 t, %5, %2;\n\tsubc.cc.u32 t, %4, %1;\n\tsubc.u32 %0, 0, 0;\n\t}"
        : "=r"(m)
        : "r"((uint32_t)(ka >> 32)), "r"((uint32_t)ka), "r"(ra), "r"((uint32_t)(kb >> 32)), "r"((uint32_t)kb), "r"(rb));
    return m;
}
// compare-exchange inside a lane, without branches or predicates: `up` = ascending
template <int EPL>
__device__ __forceinline__ void pair_cswap(unsigned long long (&key)[EPL], uint32_t (&row)[EPL], int k, int k2, bool up) {
    const uint32_t gt = pair_gt_mask(key[k], row[k], key[k2], row[k2]);
    const uint32_t take = up ? gt : ~gt;                                   // swap when (k > k2) == up
    const unsigned long long take64 = ((unsigned long long)take << 32) | take;
    const unsigned long long dk = (key[k] ^ key[k2]) & take64;
    const uint32_t dr = (row[k] ^ row[k2]) & take;
    key[k] ^= dk; key[k2] ^= dk;
    row[k] ^= dr; row[k2] ^= dr;
}
// The network as loops: the stages inside a lane (register indices are compile-time) are unrolled once - the lane's own
// sort, then one merge tail shared by all larger sizes - and the stages across lanes, identical but for the partner
// distance and the direction, are a run-time loop.  Fully unrolled the kernel was 33 k instructions (0.5 MB) per
// template instance; this form is ~1.5 k.
template <int EPL>
__device__ __forceinline__ void sort_pairs_warp(unsigned long long (&key)[EPL], uint32_t (&row)[EPL], int lane) {
    // sizes 2 .. EPL: inside the lane
#pragma unroll
    for (int size = 2; size <= EPL; size <<= 1) {
#pragma unroll
        for (int d = size >> 1; d > 0; d >>= 1) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                if ((k & d) == 0) {
                    const bool up = (size < EPL) ? ((k & size) == 0) : ((lane & 1) == 0);
                    pair_cswap<EPL>(key, row, k, k | d, up);
                }
            }
        }
    }
    // sizes 2 EPL .. 32 EPL: partner lanes at distance sl / 2 .. 1, then the tail inside the lane
#pragma unroll 1
    for (int sl = 2; sl <= 32; sl <<= 1) {
        const bool up = (lane & sl) == 0 || sl == 32;
#pragma unroll 1
        for (int dl = sl >> 1; dl > 0; dl >>= 1) {
            const bool keep_min = (((lane & dl) == 0) == up);
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key[k], dl);
                const uint32_t orow = __shfl_xor_sync(0xffffffffu, row[k], dl);
                const uint32_t gt = pair_gt_mask(key[k], row[k], ok, orow);       // (strict: the rows differ)
                const uint32_t take = keep_min ? gt : ~gt;                         // take the partner's when (mine > its) == keep_min
                const unsigned long long take64 = ((unsigned long long)take << 32) | take;
                key[k] = (ok & take64) | (key[k] & ~take64);
                row[k] = (orow & take) | (row[k] & ~take);
            }
        }
#pragma unroll
        for (int d = EPL >> 1; d > 0; d >>= 1) {
#pragma unroll
            for (int k = 0; k < EPL; ++k)
                if ((k & d) == 0) pair_cswap<EPL>(key, row, k, k | d, up);
        }
    }
}

// sky_prepare_column for nspat <= 32 EPL, one warp per column, by sorting instead of rank counting.
// skeys: nspat 64-bit words, rank_by_row: nspat 16-bit words (both shared, per warp).
template <int EPL>
__device__ __forceinline__ void sky_prepare_column_sorted(int lane, const double* col, int st, int nspat, double* lim_lo,
                                                          double* lim_hi, int spatial_floor, int* n_sky_out,
                                                          int* n_good_out, int* flag_out, double* __restrict__ sorted,
                                                          uint16_t* __restrict__ rank_of, uint16_t* __restrict__ good_row,
                                                          unsigned long long* skeys, uint16_t* rank_by_row) {
    double lo = *lim_lo, hi = *lim_hi;
    const double top = (double)(nspat - 1) + (double)spatial_floor;
    if (lo < 0.0) lo = 0.0;
    if (hi > top) hi = top;
    __syncwarp();
    if (lane == 0) { *lim_lo = lo; *lim_hi = hi; }
    const double edge_lo = lo + (double)spatial_floor, edge_hi = hi + (double)spatial_floor;
    unsigned long long key[EPL];
    uint32_t row[EPL];
    int my_sky = 0, my_nan = 0;
    unsigned long long my_min = ~0ull, my_max = 0ull;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        const int i = 32 * k + lane;
        key[k] = kKeyExcluded;
        row[k] = (uint32_t)i;
        if (i < nspat) {
            rank_by_row[i] = 0xffffu;
            const double axis = (double)(i + spatial_floor);
            if (axis < edge_lo || axis > edge_hi) {
                const double v = col[(size_t)i * st];
                ++my_sky;
                if (v != v) ++my_nan;
                else {
                    key[k] = key_of(v);
                    my_min = key[k] < my_min ? key[k] : my_min;
                    my_max = key[k] > my_max ? key[k] : my_max;
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_sky += __shfl_xor_sync(0xffffffffu, my_sky, o);
        my_nan += __shfl_xor_sync(0xffffffffu, my_nan, o);
        const unsigned long long tmin = __shfl_xor_sync(0xffffffffu, my_min, o), tmax = __shfl_xor_sync(0xffffffffu, my_max, o);
        my_min = tmin < my_min ? tmin : my_min;
        my_max = tmax > my_max ? tmax : my_max;
    }
    const int n_sky = my_sky, n_valid = my_sky - my_nan;
    *n_sky_out = n_sky;
    *n_good_out = 0;
    if (n_sky == 0) { *flag_out = kSkyNoPixels; return; }
    // len(set(sky_pixels)) == 1 (sky.py:338): one pixel, or all pixels the same number (NaNs never collapse)
    if (n_sky == 1 || (my_nan == 0 && value_of(my_min) == value_of(my_max))) { *flag_out = kSkySkipped; return; }
    *flag_out = kSkyModelled;

    sort_pairs_warp<EPL>(key, row, lane);
    // np.nanmedian / np.nanstd of the valid sky pixels
    double med = NAN;
    if (n_valid > 0) {
        // sorted position p sits in register p % EPL of lane p / EPL
        const int p1 = (n_valid - 1) / 2, p2 = n_valid / 2;
        unsigned long long k1 = 0ull, k2 = 0ull;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            if ((p1 % EPL) == k) k1 = key[k];
            if ((p2 % EPL) == k) k2 = key[k];
        }
        const double a = value_of(__shfl_sync(0xffffffffu, k1, p1 / EPL)), b = value_of(__shfl_sync(0xffffffffu, k2, p2 / EPL));
        med = ((n_valid & 1) ? a : (a + b) * 0.5);
    }
    double part = 0.0;
#pragma unroll
    for (int k = 0; k < EPL; ++k) part += (lane * EPL + k < n_valid) ? value_of(key[k]) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const double mean = part / (double)n_valid;
    part = 0.0;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        const double dd = value_of(key[k]) - mean;
        part += (lane * EPL + k < n_valid) ? dd * dd : 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const double sigma = sqrt(part / (double)n_valid);
    const double cut_lo = med - (10.0 * sigma), cut_hi = med + (10.0 * sigma);
    // single-pass 10-sigma clip (sky.py:342-345): the survivors are a contiguous run [first, last) of the sorted values
    int below = 0, under = 0;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        const bool valid = lane * EPL + k < n_valid;
        const double v = value_of(key[k]);
        below += valid && !(v > cut_lo);
        under += valid && (v < cut_hi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        below += __shfl_xor_sync(0xffffffffu, below, o);
        under += __shfl_xor_sync(0xffffffffu, under, o);
    }
    const int first = below, last = under > below ? under : below;
    const int n_good = last - first;
    *n_good_out = n_good;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        const int p = lane * EPL + k;
        if (p >= first && p < last) {
            sorted[p - first] = value_of(key[k]);
            rank_by_row[row[k]] = (uint16_t)(p - first);
        }
    }
    __syncwarp();
    // good pixels in row order -> their ranks (and rows)
    int base = 0;
    for (int start = 0; start < nspat; start += 32) {
        const int i = start + lane;
        const uint16_t r = (i < nspat) ? rank_by_row[i] : (uint16_t)0xffffu;
        const unsigned bal = __ballot_sync(0xffffffffu, r != 0xffffu);
        if (r != 0xffffu) {
            const int g = base + __popc(bal & ((1u << lane) - 1u));
            rank_of[g] = r;
            if (good_row) good_row[g] = (uint16_t)i;
        }
        base += __popc(bal);
    }
    __syncwarp();
}

// Tile geometry of the sky preparation (MOTES_PREP_COLS columns, MOTES_PREP_THREADS threads, MOTES_PREP_CTAS CTAs per SM).
// Measured on B200, 128 frames of 4096 x 400: 16 columns x 256 threads x 2 CTAs per SM 9.26 ms; 8 x 128 x 4 (the same 16
// warps per SM in finer units) 8.67 ms; 8 x 128 x 5 (96 registers: the sort spills) 20.4 ms.
#ifndef MOTES_PREP_COLS
#define MOTES_PREP_COLS 8
#endif
#ifndef MOTES_PREP_THREADS
#define MOTES_PREP_THREADS 128
#endif
constexpr int kPrepCols = MOTES_PREP_COLS, kPrepStride = kPrepCols + 1, kPrepThreads = MOTES_PREP_THREADS, kPrepWarps = kPrepThreads / 32;
__global__ void __launch_bounds__(kPrepThreads, MOTES_PREP_CTAS)
sky_prepare_kernel(const double* __restrict__ data, size_t frame_stride, int nspat, int ndisp,
                   double* __restrict__ limits /* [F][2][ndisp], clamped in place */, int spatial_floor,
                   int32_t* __restrict__ n_sky, int32_t* __restrict__ n_good, int32_t* __restrict__ flag,
                   double* __restrict__ sorted /* [F][ndisp][nspat] */, uint16_t* __restrict__ rank_of,
                   uint16_t* __restrict__ good_row /* optional, same shape as rank_of */, FrameState* __restrict__ st) {
    extern __shared__ __align__(16) unsigned char sp_smem[];
    double* tile = reinterpret_cast<double*>(sp_smem);
    // (staged keys and digit histograms: only the general path, nspat > 512, has them - sky_prepare_smem)
    const size_t nkeys = nspat > 512 ? (size_t)kPrepWarps * nspat : 0;
    unsigned long long* keys_all = reinterpret_cast<unsigned long long*>(tile + (size_t)nspat * kPrepStride);
    uint32_t* hist_all = reinterpret_cast<uint32_t*>(keys_all + nkeys);
    unsigned long long* cells = reinterpret_cast<unsigned long long*>(hist_all + kPrepWarps * 256);
    uint16_t* rows_all = reinterpret_cast<uint16_t*>(cells + kPrepWarps);
    const int f = blockIdx.y;
    if (st[f].status != 0) return;                          // failed frame: left as it came
    const int col0 = blockIdx.x * kPrepCols;
    const int ncols = min(kPrepCols, ndisp - col0);
    const double* src = data + (size_t)f * frame_stride;
    for (int idx = threadIdx.x; idx < nspat * kPrepCols; idx += blockDim.x) {
        int row = idx / kPrepCols, c = idx % kPrepCols;
        if (c < ncols) tile[row * kPrepStride + c] = src[(size_t)row * ndisp + col0 + c];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    WarpScope sc{(int)(threadIdx.x & 31)};
    for (int c = warp; c < ncols; c += kPrepWarps) {
        const size_t col = (size_t)f * ndisp + col0 + c;
        double* lo = limits + (size_t)f * 2 * ndisp + col0 + c;
        double* hi = lo + ndisp;
        int ns, ng, fl;
        int icell;
        uint16_t* grow = (good_row != nullptr) ? good_row + col * nspat : nullptr;
        if (nspat <= 128) {
            sky_prepare_column_sorted<4>(sc.tid, tile + c, kPrepStride, nspat, lo, hi, spatial_floor, &ns, &ng, &fl,
                                         sorted + col * nspat, rank_of + col * nspat, grow, keys_all + (size_t)warp * nspat,
                                         rows_all + (size_t)warp * nspat);
        } else if (nspat <= 256) {
            sky_prepare_column_sorted<8>(sc.tid, tile + c, kPrepStride, nspat, lo, hi, spatial_floor, &ns, &ng, &fl,
                                         sorted + col * nspat, rank_of + col * nspat, grow, keys_all + (size_t)warp * nspat,
                                         rows_all + (size_t)warp * nspat);
        } else if (nspat <= 512) {
            sky_prepare_column_sorted<16>(sc.tid, tile + c, kPrepStride, nspat, lo, hi, spatial_floor, &ns, &ng, &fl,
                                          sorted + col * nspat, rank_of + col * nspat, grow, keys_all + (size_t)warp * nspat,
                                          rows_all + (size_t)warp * nspat);
        } else {
            sky_prepare_column(sc, tile + c, kPrepStride, nspat, lo, hi, spatial_floor, &ns, &ng, &fl,
                               sorted + col * nspat, rank_of + col * nspat, rows_all + (size_t)warp * nspat,
                               keys_all + (size_t)warp * nspat, hist_all + warp * 256, cells + warp, &icell);
            if (good_row && fl == kSkyModelled)
                for (int g = sc.tid; g < ng; g += 32) good_row[col * nspat + g] = rows_all[(size_t)warp * nspat + g];
        }
        if (sc.tid == 0) {
            n_sky[col] = ns;
            n_good[col] = ng;
            flag[col] = fl;
            if (fl == kSkyNoPixels) atomicMin(&st[f].status, (int)MOTES_E_NOSKY);
        }
        __syncwarp();
    }
}

// One CTA per frame walks the columns in order and consumes the frame's MT19937 stream.
constexpr int kBootThreads = 256;     // >= 227: one thread per word of a generator sweep

__global__ void __launch_bounds__(kBootThreads)
sky_bootstrap_kernel(uint32_t* __restrict__ rng /* [F][625]: 624 state words + position */, int nspat, int ndisp,
                     const int32_t* __restrict__ n_good, const int32_t* __restrict__ flag,
                     const double* __restrict__ sorted, const uint16_t* __restrict__ rank_of,
                     double* __restrict__ sky, double* __restrict__ sky_err, const FrameState* __restrict__ st) {
    extern __shared__ __align__(16) unsigned char sb_smem[];
    BootShared sh;
    sh.meds = reinterpret_cast<double*>(sb_smem);
    sh.sq = sh.meds + kBootstrap;
    sh.ring = reinterpret_cast<uint32_t*>(sh.sq + kBootstrap);
    sh.counts = sh.ring + kRing;
    sh.ctl = reinterpret_cast<int*>(sh.counts + 64);
    sh.hist = reinterpret_cast<uint32_t*>(sh.ctl + 4);
    sh.ranks = reinterpret_cast<uint16_t*>(sh.hist + (size_t)kBootstrap * boot_row_words(nspat));
    BlockScope sc{(int)threadIdx.x};
    const int f = blockIdx.x;
    if (st[f].status != 0) return;
    const size_t base = (size_t)f * ndisp;
    sky_bootstrap_frame(sc, rng + (size_t)f * 625, ndisp, n_good + base, flag + base, sorted + base * nspat,
                        rank_of + base * nspat, (size_t)nspat, sky + base, sky_err + base, sh);
}

// Bootstrap polynomial sky (sky_order > 0): one CTA per frame, columns in order.
__global__ void __launch_bounds__(kBootThreads)
sky_bootstrap_poly_kernel(uint32_t* __restrict__ rng, int nspat, int ndisp, int order, int spatial_floor,
                          const double* __restrict__ errs, size_t frame_stride, const int32_t* __restrict__ n_good,
                          const int32_t* __restrict__ flag, const double* __restrict__ sorted,
                          const uint16_t* __restrict__ rank_of, const uint16_t* __restrict__ good_row,
                          double* __restrict__ yscratch /* [F][100*nspat] */, double* __restrict__ model,
                          double* __restrict__ model_err /* [F][nspat][ndisp] */, const FrameState* __restrict__ st) {
    extern __shared__ __align__(16) unsigned char sp2_smem[];
    PolyShared sh;
    sh.q = reinterpret_cast<double*>(sp2_smem);
    sh.rmat = sh.q + (size_t)nspat * kMaxPoly;
    sh.colscale = sh.rmat + kMaxPoly * kMaxPoly;
    sh.coef = sh.colscale + kMaxPoly;
    sh.loc = sh.coef + kBootstrap * kMaxPoly;
    sh.scl = sh.loc + nspat;
    sh.red = sh.scl + nspat;
    sh.samples = sh.red + 40;
    sh.ring = reinterpret_cast<uint32_t*>(sh.samples + (size_t)(kBootThreads / 32) * (kBootstrap + 4));
    sh.counts = sh.ring + kRing;
    sh.ctl = reinterpret_cast<int*>(sh.counts + 16);
    sh.rows = reinterpret_cast<uint16_t*>(sh.ctl + 4);
    BlockScope sc{(int)threadIdx.x};
    const int f = blockIdx.x;
    if (st[f].status != 0) return;
    const size_t base = (size_t)f * ndisp;
    sky_bootstrap_poly_frame(sc, rng + (size_t)f * 625, ndisp, nspat, ndisp, order, spatial_floor,
                             errs + (size_t)f * frame_stride, n_good + base, flag + base, sorted + base * nspat,
                             rank_of + base * nspat, good_row + base * nspat, (size_t)nspat,
                             yscratch + (size_t)f * kBootstrap * nspat, model + (size_t)f * frame_stride,
                             model_err + (size_t)f * frame_stride, sh);
}

__global__ void __launch_bounds__(128)
sky_apply_poly_kernel(double* __restrict__ data, double* __restrict__ errs, size_t frame_stride, int nspat, int ndisp,
                      const int32_t* __restrict__ flag, const double* __restrict__ model,
                      const double* __restrict__ model_err, const FrameState* __restrict__ st) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (c >= ndisp || st[f].status != 0) return;
    sky_apply_column_poly(data + (size_t)f * frame_stride, errs + (size_t)f * frame_stride, nspat, ndisp, c,
                          flag[(size_t)f * ndisp + c], model + (size_t)f * frame_stride,
                          model_err + (size_t)f * frame_stride);
}

__global__ void __launch_bounds__(128)
sky_apply_kernel(double* __restrict__ data, double* __restrict__ errs, size_t frame_stride, int nspat, int ndisp,
                 const int32_t* __restrict__ flag, const double* __restrict__ sky, const double* __restrict__ sky_err,
                 const FrameState* __restrict__ st) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
    if (c >= ndisp || st[f].status != 0) return;
    const size_t col = (size_t)f * ndisp + c;
    sky_apply_column(data + (size_t)f * frame_stride, errs + (size_t)f * frame_stride, nspat, ndisp, c, flag[col],
                     sky[col], sky_err[col]);
}

// np.random.seed(k) for every frame: state words + position 624 (regenerate before the first draw).
__global__ void seed_states_kernel(const uint32_t* __restrict__ seeds, int nframes, uint32_t* __restrict__ rng) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    mt_seed(rng + (size_t)f * 625, seeds[f]);
    rng[(size_t)f * 625 + 624] = kMtN;
}

// ---------------------------------------------------------------------------------------------
// a10: extraction (extraction.py:39-210).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
optimal_extract_kernel(double* __restrict__ data, const double* __restrict__ errs, size_t frame_stride, int nspat,
                       int ndisp, const double* __restrict__ limits /* [F][2][ndisp] */,
                       const double* __restrict__ bin_params /* [F][cap][8] */, const int32_t* __restrict__ nbins,
                       int cap, int wavelength_start, double* __restrict__ spectra /* [F][4][ndisp] */,
                       int32_t* __restrict__ lo_idx, int32_t* __restrict__ hi_idx, uint8_t* __restrict__ crmask,
                       const FrameState* __restrict__ st) {
    extern __shared__ __align__(16) double ext_smem[];
    double* sd = ext_smem;
    double* se = ext_smem + (size_t)nspat * kTileStride;
    const int f = blockIdx.y;
    const int col0 = blockIdx.x * kTileCols;
    const int ncols = min(kTileCols, ndisp - col0);
    if (st && st[f].status != 0) {      // failed frame: the reference raised before it got here - zero spectra, frame untouched
        for (int i = threadIdx.x; i < 4 * ncols; i += blockDim.x)
            spectra[(size_t)f * 4 * ndisp + (size_t)(i / ncols) * ndisp + col0 + i % ncols] = 0.0;
        return;
    }
    double* fdata = data + (size_t)f * frame_stride;
    const double* ferrs = errs + (size_t)f * frame_stride;
    // stage + scrub (extraction.py:31-34, 86)
    for (int idx = threadIdx.x; idx < nspat * kTileCols; idx += blockDim.x) {
        int row = idx / kTileCols, c = idx % kTileCols;
        if (c < ncols) {
            size_t g = (size_t)row * ndisp + col0 + c;
            double d = fdata[g];
            double e = fabs(ferrs[g]);
            if (!is_finite(e)) e = 0.0;
            if (!is_finite(d)) { d = 0.0; fdata[g] = 0.0; }
            sd[row * kTileStride + c] = d;
            se[row * kTileStride + c] = e;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    WarpCtx ctx{(int)(threadIdx.x & 31)};
    const int nb = nbins[f];
    const double* bp = bin_params + (size_t)f * cap * 8;
    const double* lim_lo = limits + (size_t)f * 2 * ndisp;
    const double* lim_hi = lim_lo + ndisp;
    for (int c = warp; c < ncols; c += kTileWarps) {
        int col = col0 + c;
        // the bin whose parameters the reference holds when it reaches this column
        // (extraction.py:104-112): bins tile the axis in order, so it is the last bin whose
        // first column is <= this column.
        int which = -1;
        {
            int lo_b = 0, hi_b = nb - 1;
            double key = (double)(col + wavelength_start);
            while (lo_b <= hi_b) {
                int mid = (lo_b + hi_b) >> 1;
                if (bp[(size_t)mid * 8 + 6] <= key) { which = mid; lo_b = mid + 1; }
                else hi_b = mid - 1;
            }
        }
        double zero_bin[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const double* bin = (which >= 0) ? bp + (size_t)which * 8 : zero_bin;
        ColumnSpectrum out;
        extract_column(ctx, sd + c, se + c, kTileStride, nspat, lim_lo[col], lim_hi[col], bin, &out,
                       crmask ? crmask + ((size_t)f * ndisp + col) * nspat : nullptr);
        if (ctx.lane == 0) {
            double* sp = spectra + (size_t)f * 4 * ndisp;
            sp[col] = out.optimal;
            sp[(size_t)ndisp + col] = out.optimal_err;
            sp[(size_t)2 * ndisp + col] = out.aperture;
            sp[(size_t)3 * ndisp + col] = out.aperture_err;
            if (lo_idx) lo_idx[(size_t)f * ndisp + col] = out.lo;
            if (hi_idx) hi_idx[(size_t)f * ndisp + col] = out.hi;
        }
    }
}

// float64 {0,1} quality frame -> uint8 (harvesters hand out float64, gmosio.py:50-52)
__global__ void qual_pack_kernel(const double* __restrict__ q, size_t n, uint8_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (q[i] != 0.0) ? 1 : 0;
}

// float32 planes as the GMOS harvester hands them over (big-endian float32 straight from the FITS file,
// io/gmosio.py:40-52) -> float64 frames: byte swap (optional) and exact widening, four pixels per thread.
__global__ void __launch_bounds__(256)
widen_f32_kernel(const uint32_t* __restrict__ src, size_t n, int swap, double* __restrict__ dst) {
    const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;   // odd frame sizes
    if (i4 + 3 < n && aligned) {
        const uint4 w = *reinterpret_cast<const uint4*>(src + i4);
        uint32_t v[4] = {w.x, w.y, w.z, w.w};
        double o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = (double)__uint_as_float(swap ? __byte_perm(v[k], 0u, 0x0123) : v[k]);
        *reinterpret_cast<double2*>(dst + i4) = make_double2(o[0], o[1]);
        *reinterpret_cast<double2*>(dst + i4 + 2) = make_double2(o[2], o[3]);
    } else {
        for (size_t i = i4; i < n && i < i4 + 4; ++i) {
            const uint32_t v = src[i];
            dst[i] = (double)__uint_as_float(swap ? __byte_perm(v, 0u, 0x0123) : v);
        }
    }
}

}  // namespace motes
