// libmotes_b200.so: stage launchers and the C ABI declared in include/motes_b200.h.
#include "handle.cuh"
#include "kernels.cuh"
#include "skyseg.cuh"
#include "walktab.cuh"
#include "mtjump.hpp"
#include "peak.cuh"
#include <utility>

namespace motes {

// Work-area slots inside motes_handle::work
enum WorkSlot { W_STATE = 0, W_F64, W_I32, W_U8, W_SORTED, W_RANKS, W_PROFILES, W_GOODROW, W_YSCRATCH, W_MODEL, W_MODELERR,
                W_STREAM, W_WINDOWS, W_SNAPS, W_SEGI, W_SUBS, W_YPOOL, W_FITSTATE, W_FITLIST, W_WALKTAB };

// Device-side working set of one batch.  All pointers are carved out of handle scratch.
struct Batch {
    int F, nspat, ndisp, cap;
    FrameState* st;
    double *medprof, *fit0_params, *S, *V;
    int32_t* fit0_status;
    uint8_t* gap;
    uint8_t* allgood;           // per column: every row has qual == 1
    // two bin sets: [0] sky localisation, [1] extraction
    int32_t* bins_i[2];
    double* bins_snr[2];
    int32_t* nbins[2];
    int32_t* tmp_i;
    double* tmp_snr;
    double* profiles;
    double *fit_params, *fit_cost;
    int32_t *fit_nfev, *fit_njev, *fit_status;
    double* params8[2];
    double* table;
    double* limits[2];
    double* inner;
    int32_t *n_sky, *n_good, *sky_flag;
    double *sorted, *sky, *sky_err;
    uint16_t* rank_of;
    double* spectra;
    // sky_order > 0 only
    uint16_t* good_row;
    double *yscratch, *model, *model_err;
};

struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <class T>
    T* take(size_t count) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

static size_t carve_f64(Batch& b, void* base) {
    Carver c(base);
    const size_t F = b.F, ns = b.nspat, nd = b.ndisp, cap = b.cap;
    b.medprof = c.take<double>(F * ns);
    b.fit0_params = c.take<double>(F * kParams);
    b.S = c.take<double>(F * nd);
    b.V = c.take<double>(F * nd);
    for (int k = 0; k < 2; ++k) b.bins_snr[k] = c.take<double>(F * cap);
    b.tmp_snr = c.take<double>(F * cap);
    b.fit_params = c.take<double>(F * cap * kParams);
    b.fit_cost = c.take<double>(F * cap);
    for (int k = 0; k < 2; ++k) b.params8[k] = c.take<double>(F * cap * 8);
    b.table = c.take<double>(F * 4 * cap);
    for (int k = 0; k < 2; ++k) b.limits[k] = c.take<double>(F * 2 * nd);
    b.inner = c.take<double>(F * 2 * (nd + 2));
    b.sky = c.take<double>(F * nd);
    b.sky_err = c.take<double>(F * nd);
    b.spectra = c.take<double>(F * 4 * nd);
    return c.off + 256;
}

static size_t carve_i32(Batch& b, void* base) {
    Carver c(base);
    const size_t F = b.F, nd = b.ndisp, cap = b.cap;
    b.fit0_status = c.take<int32_t>(F);
    for (int k = 0; k < 2; ++k) b.bins_i[k] = c.take<int32_t>(F * 2 * cap);
    for (int k = 0; k < 2; ++k) b.nbins[k] = c.take<int32_t>(F);
    b.tmp_i = c.take<int32_t>(F * 2 * cap);
    b.fit_nfev = c.take<int32_t>(F * cap);
    b.fit_njev = c.take<int32_t>(F * cap);
    b.fit_status = c.take<int32_t>(F * cap);
    b.n_sky = c.take<int32_t>(F * nd);
    b.n_good = c.take<int32_t>(F * nd);
    b.sky_flag = c.take<int32_t>(F * nd);
    return c.off + 256;
}

static int make_batch(motes_handle* h, Batch& b, int F, int nspat, int ndisp, int cap, bool need_sky, bool need_poly = false) {
    b = Batch();
    b.F = F; b.nspat = nspat; b.ndisp = ndisp; b.cap = cap;
    MOTES_TRY(h->work[W_STATE].ensure(sizeof(FrameState) * (size_t)F));
    b.st = h->work[W_STATE].as<FrameState>();
    MOTES_TRY(h->work[W_F64].ensure(carve_f64(b, nullptr)));
    carve_f64(b, h->work[W_F64].ptr);
    MOTES_TRY(h->work[W_I32].ensure(carve_i32(b, nullptr)));
    carve_i32(b, h->work[W_I32].ptr);
    MOTES_TRY(h->work[W_U8].ensure((size_t)2 * F * ndisp + 512));
    b.gap = h->work[W_U8].as<uint8_t>();
    b.allgood = b.gap + (((size_t)F * ndisp + 255) & ~(size_t)255);
    MOTES_TRY(h->work[W_PROFILES].ensure((size_t)F * cap * nspat * sizeof(double)));
    b.profiles = h->work[W_PROFILES].as<double>();
    if (need_sky) {
        MOTES_TRY(h->work[W_SORTED].ensure((size_t)F * ndisp * nspat * sizeof(double)));
        b.sorted = h->work[W_SORTED].as<double>();
        MOTES_TRY(h->work[W_RANKS].ensure((size_t)F * ndisp * nspat * sizeof(uint16_t)));
        b.rank_of = h->work[W_RANKS].as<uint16_t>();
    }
    if (need_sky && need_poly) {
        MOTES_TRY(h->work[W_GOODROW].ensure((size_t)F * ndisp * nspat * sizeof(uint16_t)));
        b.good_row = h->work[W_GOODROW].as<uint16_t>();
        MOTES_TRY(h->work[W_YSCRATCH].ensure((size_t)F * kBootstrap * nspat * sizeof(double)));
        b.yscratch = h->work[W_YSCRATCH].as<double>();
        MOTES_TRY(h->work[W_MODEL].ensure((size_t)F * ndisp * nspat * sizeof(double)));
        b.model = h->work[W_MODEL].as<double>();
        MOTES_TRY(h->work[W_MODELERR].ensure((size_t)F * ndisp * nspat * sizeof(double)));
        b.model_err = h->work[W_MODELERR].as<double>();
    }
    return MOTES_OK;
}

#define MOTES_LAUNCHED(h)               \
    do {                                \
        MOTES_CUDA(cudaGetLastError()); \
        (h)->launches += 1;             \
    } while (0)

// ------------------------------------------------------------------------------------------ stages
static int stage_row_median(motes_handle* h, const double* data, size_t frame_stride, int F, int nspat, int ndisp,
                            double* out) {
    StageTimer timer(h, ST_ROW_MEDIAN);
    size_t smem = (size_t)ndisp * 8 + 256 * 4 + 32;
    if (smem > h->smem_optin) {          // long merged axes: select straight from global memory / L2
        row_median_long_kernel<<<dim3(nspat, F), 256, 0, h->stream>>>(data, frame_stride, nspat, ndisp, out);
        MOTES_LAUNCHED(h);
        return MOTES_OK;
    }
    MOTES_CUDA(cudaFuncSetAttribute(row_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    row_median_kernel<<<dim3(nspat, F), 256, smem, h->stream>>>(data, frame_stride, nspat, ndisp, out);
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

// One warp per fit, whole iteration inside the warp (round-1 kernel; kept for A/B: MOTES_B200_FIT=warp).
static int stage_fit_warp(motes_handle* h, const FitBatch& fb) {
    size_t per_warp = fit_scratch_doubles(fb.m) * sizeof(double);
    size_t budget = h->smem_optin > 2048 ? h->smem_optin - 1024 : h->smem_optin;
    int warps = (int)(budget / per_warp);
    if (warps < 1) return fail_arg("nspat too large for the shared-memory fit slab");
    if (warps > kFitMaxWarps) warps = kFitMaxWarps;
    size_t smem = per_warp * warps;
    MOTES_CUDA(cudaFuncSetAttribute(moffat_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = (int)(budget / smem);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    long long items = (long long)fb.nframes * fb.bound;
    long long grid = (long long)h->sm_count * ctas_per_sm;
    long long needed = (items + warps - 1) / warps;
    if (grid > needed) grid = needed;
    MOTES_TRY(h->counters.ensure(256));
    MOTES_CUDA(cudaMemsetAsync(h->counters.ptr, 0, 256, h->stream));
    moffat_fit_kernel<<<(int)grid, warps * 32, smem, h->stream>>>(fb, h->counters.as<unsigned int>());
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

// Groups of fits advanced in lock step by one CTA each (kernels.cuh: moffat_fit_group_kernel).
//
// Measured on B200 (128 frames of 4096 x 400, 65.7 k fits, 76 ms) and NOT adopted: 5 or 6 CTAs per SM (96 / 80
// registers, group state in global memory): 76 ms; groups of 64 / 128 fits with one step-phase thread each on all
// four warps: 75-76 ms; one warp per group of 8 / 16 / 32 fits with warp-wide sweeps and no CTA barriers: 96-106 ms;
// a single Householder body instead of six specialisations: 81 ms.  ncu: instruction-cache hit rate 62 %, FP64 pipe
// 19 %, issue 25 % - the kernel streams ~110 KB of mostly straight-line FP64 code per fit and round.
// Lock-step rounds (kernels.cuh: fit_init / fit_sweep / fit_step / fit_finish): the default for large batches.
template <int THREADS>
static int stage_fit_rounds_t(motes_handle* h, const FitBatch& fb) {
    const long long items = (long long)fb.nframes * fb.bound;
    const size_t smem = (fit_sweep_doubles(fb.m) + BlockCtx<THREADS>::kRedDoubles) * sizeof(double);
    if (smem > h->smem_optin) return fail_arg("nspat too large for the shared-memory fit slab");
    MOTES_TRY(h->work[W_FITSTATE].ensure((size_t)items * sizeof(FitState) + sizeof(FitRounds) + 256));
    MOTES_TRY(h->work[W_FITLIST].ensure((size_t)items * 2 * sizeof(unsigned int) + 256));
    FitRounds* rounds = h->work[W_FITSTATE].as<FitRounds>();
    FitState* states = reinterpret_cast<FitState*>(reinterpret_cast<char*>(rounds) + ((sizeof(FitRounds) + 255) & ~(size_t)255));
    unsigned int* list[2] = {h->work[W_FITLIST].as<unsigned int>(), h->work[W_FITLIST].as<unsigned int>() + items};
    MOTES_CUDA(cudaMemsetAsync(rounds, 0, sizeof(FitRounds), h->stream));
    auto sweep = fit_sweep_kernel<THREADS>;
    auto finish = fit_finish_kernel<THREADS>;
    MOTES_CUDA(cudaFuncSetAttribute(sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MOTES_CUDA(cudaFuncSetAttribute(finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MOTES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    if (h->fit_ctas > 0 && per_sm > h->fit_ctas) per_sm = h->fit_ctas;      // leave room for a kernel on the other stream
    long long grid = (long long)h->sm_count * per_sm;
    if (grid > items) grid = items;
    fit_init_kernel<<<(unsigned)((items + 7) / 8), 256, 0, h->stream>>>(fb, states, list[0], rounds);
    MOTES_LAUNCHED(h);
    const int nrounds = h->fit_rounds < kFitRoundsMax ? h->fit_rounds : kFitRoundsMax;
    for (int r = 0; r < nrounds; ++r) {
        sweep<<<(unsigned)grid, THREADS, smem, h->stream>>>(fb, states, list[r & 1], rounds, r);
        MOTES_LAUNCHED(h);
        fit_step_kernel<<<(unsigned)((items + 127) / 128), 128, 0, h->stream>>>(fb, states, list[r & 1], list[(r + 1) & 1], rounds, r);
        MOTES_LAUNCHED(h);
    }
    finish<<<(unsigned)grid, THREADS, smem, h->stream>>>(fb, states, list[nrounds & 1], rounds, nrounds);
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

template <int THREADS>
static int stage_fit_group_t(motes_handle* h, const FitBatch& fb) {
    const long long items = (long long)fb.nframes * fb.bound;
    // enough groups to give every SM several CTAs; at most 32 fits (one per lane of the step phase) per group
    long long group = items / (8ll * h->sm_count);
    if (group < 1) group = 1;
    if (group > kFitGroupMax) group = kFitGroupMax;
    // small groups are latency-bound (few fits in flight): spread each step over an 8-lane team; large groups are
    // throughput-bound: one thread per step costs an eighth of the instructions (B200, 128 frames: 76 vs 102 ms)
    const bool team_step = group < 16;
    const size_t smem = FitGroupShape<THREADS>::smem(fb.m, (int)group);
    if (smem > h->smem_optin) return fail_arg("nspat too large for the shared-memory fit slab");
    auto kernel = team_step ? moffat_fit_group_kernel<true, THREADS> : moffat_fit_group_kernel<false, THREADS>;
    MOTES_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MOTES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const long long groups = (items + group - 1) / group;
    long long grid = (long long)h->sm_count * per_sm;
    if (grid > groups) grid = groups;
    MOTES_TRY(h->counters.ensure(256));
    MOTES_CUDA(cudaMemsetAsync(h->counters.ptr, 0, 256, h->stream));
    kernel<<<(unsigned)grid, THREADS, smem, h->stream>>>(fb, (int)group, h->counters.as<unsigned int>());
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

// Large batches: lock-step rounds; small ones (a single frame: latency matters, few fits in flight): the group kernel.
// Both with fit_threads_for(nspat) threads per fit, so a fit's result does not depend on the batch it is part of.
static int stage_fit(motes_handle* h, const FitBatch& fb) {
    if (fb.nframes <= 0 || fb.bound <= 0) return MOTES_OK;
    if (h->fit_mode == 1) return stage_fit_warp(h, fb);
    const long long items = (long long)fb.nframes * fb.bound;
    const int threads = h->fit_threads ? h->fit_threads : fit_threads_for(fb.m);
    const bool rounds = h->fit_mode == 0 && items >= h->fit_rounds_min;
    if (threads >= 128) return rounds ? stage_fit_rounds_t<128>(h, fb) : stage_fit_group_t<128>(h, fb);
    if (threads <= 32) return rounds ? stage_fit_rounds_t<32>(h, fb) : stage_fit_group_t<32>(h, fb);
    return rounds ? stage_fit_rounds_t<64>(h, fb) : stage_fit_group_t<64>(h, fb);
}

// median profile -> scale -> fit -> FWHM / binning rows  (core.py:80-100)
static int stage_median_fit(motes_handle* h, Batch& b, const double* data, size_t frame_stride) {
    MOTES_TRY(stage_row_median(h, data, frame_stride, b.F, b.nspat, b.ndisp, b.medprof));
    StageTimer timer(h, ST_MEDIAN_FIT);
    frame_prep_kernel<<<b.F, 32, 0, h->stream>>>(b.medprof, b.nspat, b.st);
    MOTES_LAUNCHED(h);
    FitBatch fb{};
    fb.profiles = b.medprof; fb.m = b.nspat; fb.nframes = b.F; fb.bound = 1; fb.cap = 1; fb.nbins = nullptr;
    fb.alpha0 = &b.st->alpha0_first; fb.alpha0_frame_stride = sizeof(FrameState);
    fb.scale = &b.st->scale; fb.scale_frame_stride = sizeof(FrameState);
    fb.params = b.fit0_params; fb.status = b.fit0_status;
    MOTES_TRY(stage_fit(h, fb));
    frame_post_kernel<<<(b.F + 127) / 128, 128, 0, h->stream>>>(b.F, b.fit0_params, b.fit0_status, b.st);
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

static int stage_column_stats(motes_handle* h, Batch& b, const double* data, const double* errs, size_t frame_stride) {
    StageTimer timer(h, ST_COLUMN_STATS);
    column_stats_kernel<<<dim3((b.ndisp + 127) / 128, b.F), 128, 0, h->stream>>>(data, errs, frame_stride, b.nspat,
                                                                                b.ndisp, b.st, b.S, b.V, b.gap);
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

static int stage_bins(motes_handle* h, Batch& b, const uint8_t* qual, size_t qual_stride, double snr_limit,
                      double min_cols, int set) {
    StageTimer timer(h, ST_BINS);
    int threads = ((b.nspat + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    size_t smem = (size_t)((b.nspat + 3) & ~3) * sizeof(int) + (size_t)kBinWindow * (2 * sizeof(double) + 1) + 10 * sizeof(double);
    qual_columns_kernel<<<dim3((b.ndisp + 127) / 128, b.F), 128, 0, h->stream>>>(qual, qual_stride, b.nspat, b.ndisp, b.allgood);
    MOTES_LAUNCHED(h);
    bins_kernel<<<b.F, threads, smem, h->stream>>>(qual, qual_stride, b.allgood, b.S, b.V, b.nspat, b.ndisp, snr_limit, min_cols,
                                                   b.cap, b.bins_i[set], b.bins_snr[set], b.tmp_i, b.tmp_snr,
                                                   b.nbins[set], b.st);
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

static int stage_bin_profiles(motes_handle* h, Batch& b, const double* data, size_t frame_stride, int set) {
    StageTimer timer(h, ST_BIN_PROFILES);
    long long items = (long long)b.cap * b.nspat;
    bin_profile16_kernel<<<dim3((unsigned)((items + 127) / 128), b.F), 128, 0, h->stream>>>(
        data, frame_stride, b.nspat, b.ndisp, b.gap, b.bins_i[set], b.nbins[set], b.cap, b.profiles);
    MOTES_LAUNCHED(h);
    bin_profile_kernel<<<dim3(b.cap, b.F), kProfWarps * 32, 0, h->stream>>>(data, frame_stride, b.nspat, b.ndisp, b.gap,
                                                                         b.bins_i[set], b.nbins[set], b.cap, b.profiles);
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

// bins -> profiles -> fits -> limit table -> per-pixel limits  (sky.py:96-144 / core.py:200-252)
static int stage_localise(motes_handle* h, Batch& b, const double* data, const double* errs, size_t frame_stride,
                          const uint8_t* qual, size_t qual_stride, double snr_limit, const motes_params& p, int set) {
    MOTES_TRY(stage_column_stats(h, b, data, errs, frame_stride));
    MOTES_TRY(stage_bins(h, b, qual, qual_stride, snr_limit, p.minimum_column_limit, set));
    MOTES_TRY(stage_bin_profiles(h, b, data, frame_stride, set));
    FitBatch fb{};
    fb.profiles = b.profiles; fb.m = b.nspat; fb.nframes = b.F; fb.bound = b.cap; fb.cap = b.cap;
    fb.nbins = b.nbins[set];
    fb.alpha0 = &b.st->alpha0_bins; fb.alpha0_frame_stride = sizeof(FrameState);
    fb.scale = &b.st->scale; fb.scale_frame_stride = sizeof(FrameState);
    fb.params = b.fit_params; fb.cost = b.fit_cost; fb.nfev = b.fit_nfev; fb.njev = b.fit_njev; fb.status = b.fit_status;
    {
        StageTimer timer(h, ST_BIN_FITS);
        MOTES_TRY(stage_fit(h, fb));
    }
    StageTimer timer(h, ST_LIMITS);
    int items = b.F * b.cap;
    bin_post_kernel<<<(items + 127) / 128, 128, 0, h->stream>>>(b.F, b.cap, b.cap, b.nbins[set], b.bins_i[set],
                                                                b.fit_params, b.fit_status, b.st, p.sky_fwhm_multiplier,
                                                                p.wavelength_start, b.params8[set], b.table);
    MOTES_LAUNCHED(h);
    limits_kernel<<<b.F, 256, 0, h->stream>>>(b.table, b.cap, b.nbins[set], b.ndisp, b.limits[set], b.inner);
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

static size_t sky_prepare_smem(int nspat) {
    const size_t keys = nspat > 512 ? (size_t)kPrepWarps * nspat * 8 : 0;      // staged keys: the general path only
    return (size_t)nspat * kPrepStride * 8 + keys + kPrepWarps * 256 * 4 + kPrepWarps * 8 + (size_t)kPrepWarps * nspat * 2 + 64;
}
static size_t sky_boot_smem(int nspat) {
    return (size_t)2 * kBootstrap * 8 + kRing * 4 + 64 * 4 + 4 * 4 + (size_t)kBootstrap * boot_row_words(nspat) * 4 +
           (size_t)nspat * 2 + 64;
}

static size_t sky_poly_smem(int nspat) {
    return ((size_t)nspat * kMaxPoly + kMaxPoly * kMaxPoly + kMaxPoly + kBootstrap * kMaxPoly + 2 * (size_t)nspat + 40 +
            (size_t)(kBootThreads / 32) * (kBootstrap + 4)) * 8 + kRing * 4 + 16 * 4 + 4 * 4 + (size_t)nspat * 2 + 64;
}


// Order-0 bootstrap with the frame's generator stream split across the GPU (skyseg.cuh).
// *done = false when the layout does not fit (stream positions beyond 32 bits, no memory for even one
// frame): the caller then falls back to the one-CTA-per-frame kernel, which has no such limits.
static size_t sky_col_smem(int nspat) {
    return (size_t)2 * kBootstrap * 8 + (size_t)kBootstrap * boot_row_words(nspat) * 4 + (size_t)nspat * 2 + 64;
}
static size_t sky_colh_smem(int nspat, bool packed) {
    if (!packed) return (size_t)2 * kBootstrap * 8 + (size_t)kH32HistWords * 4 + (size_t)kH32Entries * 2 + 64;
    return (size_t)2 * kBootstrap * 8 + (size_t)(kBootstrap * HistLayout<true>::kStride) * 4 + (size_t)nspat * 2 + 64;
}

// Layout of the segmented replay for a batch: frames per wave, segment stride, buffers (all from the a-priori bound
// 100 * ndisp * 2^ceil(log2 nspat) words per frame, so it can be fixed before the sky pixels are known).
// plan->ok = false: the layout does not fit (the caller falls back to the one-CTA-per-frame replay).
static int boot_plan(motes_handle* h, const Batch& b, BootPlan* plan, int ctas_per_sm = 8) {
    plan->ok = false;
    plan->hoisted = false;
    const JumpTable& table = jump_table();
    if (!table.ok) return MOTES_OK;
    unsigned long long p2 = 1;
    while (p2 < (unsigned long long)b.nspat) p2 <<= 1;
    const unsigned long long bound = 624ull + 100ull * (unsigned long long)b.ndisp * p2 + kSegSlack;
    if (bound >= 0xff000000ull) return MOTES_OK;
    if (sky_col_smem(b.nspat) > h->smem_optin) return MOTES_OK;
    // per-frame footprint is dominated by the 16-bit stream: 2 bytes per generator word
    size_t budget = h->stream_budget;
    if (budget == 0) {
        size_t free_b = 0, total_b = 0;
        MOTES_CUDA(cudaMemGetInfo(&free_b, &total_b));
        free_b += h->work[W_STREAM].bytes + h->work[W_SNAPS].bytes + h->work[W_WINDOWS].bytes + h->work[W_SUBS].bytes;
        budget = free_b / 2;
    }
    const size_t per_frame_est = (size_t)(bound * 2ull + bound / 26 + bound / 512 + (1u << 20));
    int wave = (int)(budget / per_frame_est);
    if (wave < 1) return MOTES_OK;
    if (wave > b.F) wave = b.F;
    if (!h->jump_ready) {
        MOTES_TRY(h->jump_polys.ensure(table.words.size() * sizeof(uint32_t)));
        MOTES_CUDA(cudaMemcpyAsync(h->jump_polys.ptr, table.words.data(), table.words.size() * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice, h->stream));
        MOTES_CUDA(cudaStreamSynchronize(h->stream));
        h->jump_ready = true;
    }
    SegLayout lay{};
    int s = kSnapLog2;
    while (true) {
        const int waves = (b.F + wave - 1) / wave;
        wave = (b.F + waves - 1) / waves;
        // segment stride: enough CTAs to fill the machine, not more jumps than needed
        const long long target = (long long)ctas_per_sm * h->sm_count;
        s = kSnapLog2;
        while (s < 30 && (long long)wave * (long long)((bound + (1ull << s) - 1) >> s) > target) ++s;
        if (s + 10 > kJumpMaxLog2) return MOTES_OK;
        lay.log2_stride = s;
        lay.stride = 1u << s;
        lay.segs = (int)((bound + lay.stride - 1) >> s);
        const unsigned long long words = (unsigned long long)lay.segs << s;
        lay.stream_len = words + 3ull * kWalkChunk;
        lay.sub_len = (words >> kSubLog2) + 16;
        lay.snaps = (int)(words >> kSnapLog2) + 1;
        int rc = h->work[W_STREAM].ensure((size_t)wave * lay.stream_len * sizeof(uint16_t));
        if (rc == MOTES_OK) rc = h->work[W_WINDOWS].ensure((size_t)wave * lay.segs * 624 * sizeof(uint32_t));
        if (rc == MOTES_OK) rc = h->work[W_SNAPS].ensure((size_t)wave * lay.snaps * 624 * sizeof(uint32_t));
        if (rc == MOTES_OK) rc = h->work[W_SUBS].ensure((size_t)wave * lay.sub_len * sizeof(uint32_t));
        if (rc == MOTES_OK)
            rc = h->work[W_SEGI].ensure(((size_t)wave * 2 + (size_t)wave * b.ndisp * 3 + 64) * sizeof(uint32_t) + 256);
        if (rc == MOTES_OK) break;
        if (rc != MOTES_E_NOMEM) return rc;
        // the device could not give that much in one piece: half as many frames per wave, down to one
        cudaGetLastError();
        if (wave == 1) return MOTES_OK;                 // the caller falls back to the one-CTA-per-frame replay
        wave = (wave + 1) / 2;
    }
    plan->lay = lay;
    plan->wave = wave;
    plan->log2_stride = s;
    plan->ok = true;
    return MOTES_OK;
}

// Generator part of the replay for frames [f0, f0 + nf) on stream `cs`: which segments are needed (`b` = nullptr: all
// of them - the sky pixels are not known yet), the segment windows by jump-ahead, the 16-bit stream and its snapshots.
static int boot_generate(motes_handle* h, cudaStream_t cs, const Batch* b, int ndisp, const BootPlan& plan, uint32_t* rng_w,
                         int f0, int nf, bool timed) {
    const SegLayout& lay = plan.lay;
    uint16_t* stream = h->work[W_STREAM].as<uint16_t>();
    uint32_t* windows = h->work[W_WINDOWS].as<uint32_t>();
    uint32_t* snaps = h->work[W_SNAPS].as<uint32_t>();
    int32_t* seg_need = h->work[W_SEGI].as<int32_t>();
    const uint32_t* polys = h->jump_polys.as<uint32_t>();
    const size_t jump_smem = (size_t)(kJumpStreamWords + 624) * sizeof(uint32_t);
    MOTES_CUDA(cudaFuncSetAttribute(mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jump_smem));
    {
        StageTimer t_jump(h, timed ? ST_BOOT_JUMP : -1);
        if (b) {
            const size_t c0 = (size_t)f0 * ndisp;
            seg_plan_kernel<<<nf, 256, 0, cs>>>(rng_w, ndisp, b->n_good + c0, b->sky_flag + c0, lay, seg_need);
        } else {
            seg_plan_all_kernel<<<(nf + 255) / 256, 256, 0, cs>>>(nf, lay, seg_need);
        }
        MOTES_LAUNCHED(h);
        MOTES_CUDA(cudaMemsetAsync(windows, 0, (size_t)nf * lay.segs * 624 * sizeof(uint32_t), cs));
        for (int r = 0; (1 << r) < lay.segs; ++r) {
            const int count = ((1 << (r + 1)) <= lay.segs ? (1 << r) : lay.segs - (1 << r));
            mt_jump_kernel<<<dim3(count, nf, kJumpParts), kJumpThreads, jump_smem, cs>>>(
                rng_w, windows, lay.segs, r, polys + (size_t)(plan.log2_stride + r) * kJumpWords32, seg_need);
            MOTES_LAUNCHED(h);
        }
    }
    StageTimer t_stream(h, timed ? ST_BOOT_STREAM : -1);
    mt_stream_kernel<uint16_t><<<dim3(lay.segs, nf), 256, 0, cs>>>(rng_w, windows, lay, seg_need, stream, snaps);
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

// Start the generator of a batch on the second stream before anything is known about its sky pixels: the stream only
// depends on the frames' generator states, and its kernels (integer issue bound) share the SMs with the latency-bound
// localisation stages that run meanwhile on the main stream.  Only when one wave covers the batch.
// Generator CTAs per SM (MOTES_B200_HOIST_CTAS), 128 frames of 4096 x 400 on B200: 8 -> 182.4 ms per step, 4 -> 185.5,
// 3 -> 186.3, 2 -> 185.8, 1 -> 214.1 (un-hoisted 189.4): leaving the fit kernel more room does not pay, the
// generator just takes longer.
static int boot_hoist(motes_handle* h, const Batch& b, uint32_t* rng) {
    h->boot.ok = false;
    h->boot.hoisted = false;
    if (!h->hoist || (h->profiling && !h->profile_hoist) || h->boot_mode != 0 || h->chunks > 1 || !rng) return MOTES_OK;
    BootPlan plan;
    MOTES_TRY(boot_plan(h, b, &plan, h->hoist_ctas));
    if (!plan.ok || plan.wave != b.F) return MOTES_OK;
    if (!h->ev_hoist[0]) {
        MOTES_CUDA(cudaEventCreateWithFlags(&h->ev_hoist[0], cudaEventDisableTiming));
        MOTES_CUDA(cudaEventCreateWithFlags(&h->ev_hoist[1], cudaEventDisableTiming));
    }
    // The second stream already waits for the start-of-call event of run_frames_chunked, which follows the previous
    // call's work and the upload of the generator states - but not the upload of the frames: from host buffers the
    // generator runs while the frames are still on their way.  With the frames already on the device it is released
    // together with the first kernels of the main stream instead, so that the two share the SMs from the start
    // (released earlier it takes every SM first and the call gets slower, 185 against 175 ms per 128 frames).
    if (!h->upload_pending) {
        MOTES_CUDA(cudaEventRecord(h->ev_hoist[0], h->stream));
        MOTES_CUDA(cudaStreamWaitEvent(h->lane1_stream, h->ev_hoist[0], 0));
    }
    MOTES_TRY(boot_generate(h, h->lane1_stream, nullptr, b.ndisp, plan, rng, 0, b.F, false));
    MOTES_CUDA(cudaEventRecord(h->ev_hoist[1], h->lane1_stream));
    plan.hoisted = true;
    h->boot = plan;
    return MOTES_OK;
}

static int stage_bootstrap_segmented(motes_handle* h, Batch& b, uint32_t* rng, bool* done) {
    *done = false;
    BootPlan plan = h->boot;
    if (plan.hoisted) {
        MOTES_CUDA(cudaStreamWaitEvent(h->stream, h->ev_hoist[1], 0));
        h->boot.hoisted = false;
    } else {
        MOTES_TRY(boot_plan(h, b, &plan));
        if (!plan.ok) return MOTES_OK;
    }
    const SegLayout lay = plan.lay;
    const int wave = plan.wave;
    const size_t col_smem = sky_col_smem(b.nspat);
    uint16_t* stream = h->work[W_STREAM].as<uint16_t>();
    uint32_t* snaps = h->work[W_SNAPS].as<uint32_t>();
    uint32_t* subs = h->work[W_SUBS].as<uint32_t>();
    int32_t* seg_need = h->work[W_SEGI].as<int32_t>();
    uint32_t* end_pos = reinterpret_cast<uint32_t*>(seg_need + wave);
    uint32_t* col_start = end_pos + wave;
    uint32_t* col_end = col_start + (size_t)wave * b.ndisp;
    uint32_t* retry_list = col_end + (size_t)wave * b.ndisp;          // columns the histogram kernel hands to the exact one
    uint32_t* retry_count = retry_list + (size_t)wave * b.ndisp;
    const int col_mode = h->column_window;      // 0: exact kernel only; < 0: the layout's whole window; else that window
    // n <= nspat <= 512: the median rank of a bootstrap sample stays within 64 (5.7 sigma at n = 512, 6.4 at 400; a
    // sample outside sends its column to the exact kernel) of n / 2 -> plain 32-bit
    // bins over a window of 128 ranks; wider frames: packed 8-bit bins over 256 ranks
    const bool packed = b.nspat > kH32Entries || col_mode > HistLayout<false>::kBins;
    const size_t colh_smem = sky_colh_smem(b.nspat, packed);
    auto hist_kernel = packed ? sky_column_hist_kernel<true> : sky_column_hist32_kernel;
    MOTES_CUDA(cudaFuncSetAttribute(hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)colh_smem));
    MOTES_CUDA(cudaFuncSetAttribute(sky_column_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)col_smem));
    const size_t walk_smem = (size_t)kWalkStages * kWalkChunkBytes;
    MOTES_CUDA(cudaFuncSetAttribute(sky_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)walk_smem));

    for (int f0 = 0; f0 < b.F; f0 += wave) {
        const int nf = (b.F - f0 < wave) ? b.F - f0 : wave;
        const size_t c0 = (size_t)f0 * b.ndisp;
        uint32_t* rng_w = rng + (size_t)f0 * 625;
        if (!plan.hoisted) MOTES_TRY(boot_generate(h, h->stream, &b, b.ndisp, plan, rng_w, f0, nf, true));
        StageTimer t_walk(h, ST_BOOT_WALK);
        // Few frames in flight, or long ones: the per-column latency of the streaming walk (~2 us) is what costs, so the
        // counting goes to a table built by every SM and one warp per frame walks the table (walktab.cuh).  Many frames:
        // the streaming walk already reads the stream at the HBM roofline and needs no second pass.
        const WalkTabInfo* tab_info = nullptr;
        {
            const double stream_s = (double)nf * (double)lay.stream_len * 2.0 / 5.6e12;
            // (measured on B200: 2.1 us per column for the streaming walk below its bandwidth limit; the table walk by runs:
            // 0.26 us per column of a 4096 x 400 frame, 0.15 us for 24000 x 80 - 1.25 us one column at a time; the table pass
            // reads the stream at 4.2 TB/s.  128 frames of 3000 x 200: 5.9 ms streaming, 6.5 ms by table)
            const double t_stream = stream_s > b.ndisp * 2.1e-6 ? stream_s : b.ndisp * 2.1e-6;
            const double t_table = 1.4 * stream_s + b.ndisp * (h->walk_mode == 3 ? 1.3e-6 : 0.5e-6) + 50e-6;
            const bool use_table = h->walk_mode == 1 || h->walk_mode == 3 || (h->walk_mode == 0 && t_table < t_stream);
            const size_t pieces = (size_t)(lay.stream_len >> kSubLog2) + 1;
            if (use_table && h->work[W_WALKTAB].ensure((size_t)nf * pieces * kTabPiece * sizeof(uint16_t) + (size_t)nf * sizeof(WalkTabInfo) + 512) == MOTES_OK) {
                WalkTabInfo* info = h->work[W_WALKTAB].as<WalkTabInfo>();
                uint16_t* table = reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(info) + (((size_t)nf * sizeof(WalkTabInfo) + 255) & ~(size_t)255));
                walk_range_kernel<<<nf, 256, 0, h->stream>>>(b.ndisp, b.n_good + c0, b.sky_flag + c0, b.st + f0, info);
                MOTES_LAUNCHED(h);
                walk_table_kernel<<<dim3((unsigned)((pieces + 7) / 8), nf), 256, 0, h->stream>>>(stream, lay, rng_w, seg_need, info, pieces, table);
                MOTES_LAUNCHED(h);
                if (h->walk_mode == 3)                                                        // (one column at a time)
                    walk_serial_kernel<<<nf, 32, 0, h->stream>>>(stream, lay, rng_w, seg_need, b.ndisp, b.n_good + c0, b.sky_flag + c0, info, pieces,
                                                                 table, col_start, col_end, subs, end_pos, b.st + f0);
                else
                    walk_runs_kernel<<<nf, kRunWarps * 32, 0, h->stream>>>(stream, lay, rng_w, seg_need, b.ndisp, b.n_good + c0, b.sky_flag + c0, info,
                                                                           pieces, table, col_start, col_end, subs, end_pos, b.st + f0);
                MOTES_LAUNCHED(h);
                tab_info = info;
            } else if (use_table) {
                cudaGetLastError();
            }
        }
        // (frames the table does not serve - a wide range of sky-pixel counts - and every frame in streaming mode)
        sky_walk_kernel<<<nf, kWalkThreads, walk_smem, h->stream>>>(stream, lay, rng_w, seg_need, b.ndisp, b.n_good + c0, b.sky_flag + c0,
                                                   col_start, col_end, subs, end_pos, b.st + f0, tab_info);
        MOTES_LAUNCHED(h);
        t_walk.stop();
        StageTimer t_col(h, ST_BOOT_COLUMN);
        if (col_mode != 0) {
            const int bins = packed ? HistLayout<true>::kBins : HistLayout<false>::kBins;
            const int window = (col_mode < 0 || col_mode > bins) ? bins : col_mode;
            MOTES_CUDA(cudaMemsetAsync(retry_count, 0, sizeof(uint32_t), h->stream));
            hist_kernel<<<dim3(b.ndisp, nf), packed ? kColThreads : h32_threads(b.nspat), colh_smem, h->stream>>>(
                stream, lay, b.nspat, b.ndisp, window, b.n_good + c0, b.sky_flag + c0, b.sorted + c0 * b.nspat,
                b.rank_of + c0 * b.nspat, col_start, col_end, subs, b.sky + c0, b.sky_err + c0, b.st + f0, 0u, retry_list,
                retry_count);
            MOTES_LAUNCHED(h);
            sky_column_kernel<<<dim3(2 * h->sm_count, 1), kColThreads, col_smem, h->stream>>>(
                stream, lay, b.nspat, b.ndisp, b.n_good + c0, b.sky_flag + c0, b.sorted + c0 * b.nspat, b.rank_of + c0 * b.nspat,
                col_start, col_end, subs, b.sky + c0, b.sky_err + c0, b.st + f0, retry_list, retry_count);
            MOTES_LAUNCHED(h);
        } else {
            sky_column_kernel<<<dim3(b.ndisp, nf), kColThreads, col_smem, h->stream>>>(
                stream, lay, b.nspat, b.ndisp, b.n_good + c0, b.sky_flag + c0, b.sorted + c0 * b.nspat, b.rank_of + c0 * b.nspat,
                col_start, col_end, subs, b.sky + c0, b.sky_err + c0, b.st + f0, nullptr, nullptr);
            MOTES_LAUNCHED(h);
        }
        mt_export_kernel<<<nf, 256, 0, h->stream>>>(rng_w, snaps, lay, end_pos, b.st + f0);
        MOTES_LAUNCHED(h);
    }
    *done = true;
    return MOTES_OK;
}

// Polynomial sky (sky_order > 0) over the segmented generator (skyseg.cuh).  *done = false: the layout does not fit,
// the caller runs the one-CTA-per-frame kernel.
static int stage_poly_segmented(motes_handle* h, Batch& b, const double* errs, size_t frame_stride, const motes_params& p,
                                uint32_t* rng, bool* done) {
    *done = false;
    const JumpTable& table = jump_table();
    if (!table.ok || h->boot_mode != 0) return MOTES_OK;
    const unsigned long long accepts_max = (unsigned long long)(kBootstrap / 2) * b.nspat * b.ndisp;
    const unsigned long long bound = 624ull + 4ull * (unsigned long long)((double)accepts_max * 1.2733) + kSegSlack;
    if (bound >= 0xff000000ull) return MOTES_OK;
    const size_t poly_smem = sky_poly_smem(b.nspat);
    if (poly_smem > h->smem_optin) return MOTES_OK;
    size_t budget = h->stream_budget;
    if (budget == 0) {
        size_t free_b = 0, total_b = 0;
        MOTES_CUDA(cudaMemGetInfo(&free_b, &total_b));
        free_b += h->work[W_STREAM].bytes + h->work[W_SNAPS].bytes + h->work[W_WINDOWS].bytes + h->work[W_SUBS].bytes;
        budget = free_b / 2;
    }
    const size_t per_frame_est = (size_t)(bound * 4ull + bound / 16 + (1u << 20));
    int wave = (int)(budget / per_frame_est);
    if (wave < 1) return MOTES_OK;
    if (wave > b.F) wave = b.F;
    if (!h->jump_ready) {
        MOTES_TRY(h->jump_polys.ensure(table.words.size() * sizeof(uint32_t)));
        MOTES_CUDA(cudaMemcpyAsync(h->jump_polys.ptr, table.words.data(), table.words.size() * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice, h->stream));
        MOTES_CUDA(cudaStreamSynchronize(h->stream));
        h->jump_ready = true;
    }
    MOTES_CUDA(cudaFuncSetAttribute(sky_poly_column_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)poly_smem));
    int per_sm = 0;
    MOTES_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sky_poly_column_kernel, kBootThreads, poly_smem));
    if (per_sm < 1) return MOTES_OK;
    const int col_grid = h->sm_count * per_sm;
    SegLayout lay{};
    int s = kSnapLog2;
    size_t bitmap_len = 0, block_len = 0;
    while (true) {
        const int waves = (b.F + wave - 1) / wave;
        wave = (b.F + waves - 1) / waves;
        const long long target = 8ll * h->sm_count;
        s = kSnapLog2;
        while (s < 30 && (long long)wave * (long long)((bound + (1ull << s) - 1) >> s) > target) ++s;
        if (s + 10 > kJumpMaxLog2) return MOTES_OK;
        lay.log2_stride = s;
        lay.stride = 1u << s;
        lay.segs = (int)((bound + lay.stride - 1) >> s);
        const unsigned long long words = (unsigned long long)lay.segs << s;
        lay.stream_len = words + 1024;
        lay.sub_len = 0;
        lay.snaps = (int)(words >> kSnapLog2) + 1;
        bitmap_len = (size_t)(words / 128) + 64;                       // one bit per attempt of four words
        block_len = (size_t)(words / (4 * kAttemptBlock)) + 8;
        int rc = h->work[W_STREAM].ensure((size_t)wave * lay.stream_len * sizeof(uint32_t));
        if (rc == MOTES_OK) rc = h->work[W_WINDOWS].ensure((size_t)wave * lay.segs * 624 * sizeof(uint32_t));
        if (rc == MOTES_OK) rc = h->work[W_SNAPS].ensure((size_t)wave * lay.snaps * 624 * sizeof(uint32_t));
        if (rc == MOTES_OK) rc = h->work[W_SUBS].ensure((size_t)wave * (bitmap_len + block_len) * sizeof(uint32_t));
        if (rc == MOTES_OK)
            rc = h->work[W_SEGI].ensure((size_t)wave * 16 + (size_t)wave * b.ndisp * 2 * sizeof(unsigned long long) + 512);
        if (rc == MOTES_OK) rc = h->work[W_YPOOL].ensure((size_t)col_grid * kBootstrap * b.nspat * sizeof(double));
        if (rc == MOTES_OK) break;
        if (rc != MOTES_E_NOMEM) return rc;
        cudaGetLastError();
        if (wave == 1) return MOTES_OK;
        wave = (wave + 1) / 2;
    }
    uint32_t* stream = h->work[W_STREAM].as<uint32_t>();
    uint32_t* windows = h->work[W_WINDOWS].as<uint32_t>();
    uint32_t* snaps = h->work[W_SNAPS].as<uint32_t>();
    uint32_t* bitmap = h->work[W_SUBS].as<uint32_t>();
    uint32_t* block_cnt = bitmap + (size_t)wave * bitmap_len;
    unsigned long long* col_first = h->work[W_SEGI].as<unsigned long long>();
    unsigned long long* col_last = col_first + (size_t)wave * b.ndisp;
    int32_t* seg_need = reinterpret_cast<int32_t*>(col_last + (size_t)wave * b.ndisp);
    uint32_t* end_pos = reinterpret_cast<uint32_t*>(seg_need + wave);
    double* yscratch = h->work[W_YPOOL].as<double>();
    const uint32_t* polys = h->jump_polys.as<uint32_t>();
    const size_t jump_smem = (size_t)(kJumpStreamWords + 624) * sizeof(uint32_t);
    MOTES_CUDA(cudaFuncSetAttribute(mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jump_smem));
    MOTES_TRY(h->counters.ensure(256));
    const unsigned long long attempts_max = ((unsigned long long)lay.segs << s) / 4ull + 1;

    for (int f0 = 0; f0 < b.F; f0 += wave) {
        const int nf = (b.F - f0 < wave) ? b.F - f0 : wave;
        const size_t c0 = (size_t)f0 * b.ndisp;
        uint32_t* rng_w = rng + (size_t)f0 * 625;
        StageTimer t_jump(h, ST_BOOT_JUMP);
        seg_plan_poly_kernel<<<nf, 256, 0, h->stream>>>(rng_w, b.ndisp, p.sky_order, b.n_good + c0, b.sky_flag + c0, lay, seg_need);
        MOTES_LAUNCHED(h);
        MOTES_CUDA(cudaMemsetAsync(windows, 0, (size_t)nf * lay.segs * 624 * sizeof(uint32_t), h->stream));
        for (int r = 0; (1 << r) < lay.segs; ++r) {
            const int count = ((1 << (r + 1)) <= lay.segs ? (1 << r) : lay.segs - (1 << r));
            mt_jump_kernel<<<dim3(count, nf, kJumpParts), kJumpThreads, jump_smem, h->stream>>>(
                rng_w, windows, lay.segs, r, polys + (size_t)(s + r) * kJumpWords32, seg_need);
            MOTES_LAUNCHED(h);
        }
        t_jump.stop();
        StageTimer t_stream(h, ST_BOOT_STREAM);
        mt_stream_kernel<uint32_t><<<dim3(lay.segs, nf), 256, 0, h->stream>>>(rng_w, windows, lay, seg_need, stream, snaps);
        MOTES_LAUNCHED(h);
        t_stream.stop();
        StageTimer t_walk(h, ST_BOOT_WALK);
        MOTES_CUDA(cudaMemsetAsync(block_cnt, 0, (size_t)nf * block_len * sizeof(uint32_t), h->stream));
        poly_attempts_kernel<<<dim3((unsigned)((attempts_max + 255) / 256), nf), 256, 0, h->stream>>>(
            stream, lay, rng_w, seg_need, bitmap, bitmap_len, block_cnt, block_len);
        MOTES_LAUNCHED(h);
        poly_bounds_kernel<<<nf, 1024, 0, h->stream>>>(lay, rng_w, seg_need, b.ndisp, p.sky_order, b.n_good + c0, b.sky_flag + c0,
                                                       bitmap, bitmap_len, block_cnt, block_len, col_first, col_last, end_pos,
                                                       b.st + f0);
        MOTES_LAUNCHED(h);
        t_walk.stop();
        StageTimer t_col(h, ST_BOOT_COLUMN);
        MOTES_CUDA(cudaMemsetAsync(h->counters.ptr, 0, 256, h->stream));
        sky_poly_column_kernel<<<col_grid, kBootThreads, poly_smem, h->stream>>>(
            stream, lay, rng_w, nf, b.nspat, b.ndisp, p.sky_order, p.spatial_floor, errs + (size_t)f0 * frame_stride, frame_stride,
            b.n_good + c0, b.sky_flag + c0, b.sorted + c0 * b.nspat, b.rank_of + c0 * b.nspat, b.good_row + c0 * b.nspat, col_first,
            col_last, yscratch, b.model + (size_t)f0 * frame_stride, b.model_err + (size_t)f0 * frame_stride, b.st + f0,
            h->counters.as<unsigned int>());
        MOTES_LAUNCHED(h);
        mt_export_kernel<<<nf, 256, 0, h->stream>>>(rng_w, snaps, lay, end_pos, b.st + f0);
        MOTES_LAUNCHED(h);
    }
    *done = true;
    return MOTES_OK;
}

static int stage_sky(motes_handle* h, Batch& b, double* data, double* errs, size_t frame_stride, double* limits,
                     const motes_params& p, uint32_t* rng) {
    if (p.sky_order < 0 || p.sky_order >= kMaxPoly) return fail_arg("sky_order must be 0..5");
    if (b.nspat > 65535) return fail_arg("nspat too large");
    const bool poly = p.sky_order > 0;
    size_t smem = sky_prepare_smem(b.nspat);
    if (smem > h->smem_optin || sky_boot_smem(b.nspat) > h->smem_optin || sky_poly_smem(b.nspat) > h->smem_optin)
        return fail_arg("nspat too large for the shared-memory sky kernels");
    MOTES_CUDA(cudaFuncSetAttribute(sky_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        StageTimer timer(h, ST_SKY_PREPARE);
        sky_prepare_kernel<<<dim3((b.ndisp + kPrepCols - 1) / kPrepCols, b.F), kPrepThreads, smem, h->stream>>>(
            data, frame_stride, b.nspat, b.ndisp, limits, p.spatial_floor, b.n_sky, b.n_good, b.sky_flag, b.sorted,
            b.rank_of, poly ? b.good_row : nullptr, b.st);
        MOTES_LAUNCHED(h);
    }
    if (!poly) {
        {
            bool done = false;
            if (h->boot_mode == 0) MOTES_TRY(stage_bootstrap_segmented(h, b, rng, &done));
            if (!done) {
                StageTimer timer(h, ST_SKY_BOOTSTRAP);
                smem = sky_boot_smem(b.nspat);
                MOTES_CUDA(cudaFuncSetAttribute(sky_bootstrap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                sky_bootstrap_kernel<<<b.F, kBootThreads, smem, h->stream>>>(rng, b.nspat, b.ndisp, b.n_good, b.sky_flag,
                                                                             b.sorted, b.rank_of, b.sky, b.sky_err, b.st);
                MOTES_LAUNCHED(h);
            }
        }
        StageTimer timer(h, ST_SKY_APPLY);
        sky_apply_kernel<<<dim3((b.ndisp + 127) / 128, b.F), 128, 0, h->stream>>>(data, errs, frame_stride, b.nspat,
                                                                                 b.ndisp, b.sky_flag, b.sky, b.sky_err, b.st);
        MOTES_LAUNCHED(h);
        return MOTES_OK;
    }
    if (!b.good_row) return fail_arg("internal: polynomial sky scratch missing");
    smem = sky_poly_smem(b.nspat);
    MOTES_CUDA(cudaFuncSetAttribute(sky_bootstrap_poly_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MOTES_CUDA(cudaMemsetAsync(b.model, 0, sizeof(double) * (size_t)b.F * frame_stride, h->stream));
    MOTES_CUDA(cudaMemsetAsync(b.model_err, 0, sizeof(double) * (size_t)b.F * frame_stride, h->stream));
    bool poly_done = false;
    MOTES_TRY(stage_poly_segmented(h, b, errs, frame_stride, p, rng, &poly_done));
    if (!poly_done) {
        StageTimer timer(h, ST_SKY_BOOTSTRAP);
        sky_bootstrap_poly_kernel<<<b.F, kBootThreads, smem, h->stream>>>(
            rng, b.nspat, b.ndisp, p.sky_order, p.spatial_floor, errs, frame_stride, b.n_good, b.sky_flag, b.sorted,
            b.rank_of, b.good_row, b.yscratch, b.model, b.model_err, b.st);
        MOTES_LAUNCHED(h);
    }
    StageTimer timer(h, ST_SKY_APPLY);
    sky_apply_poly_kernel<<<dim3((b.ndisp + 127) / 128, b.F), 128, 0, h->stream>>>(data, errs, frame_stride, b.nspat,
                                                                                  b.ndisp, b.sky_flag, b.model, b.model_err, b.st);
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

static int stage_extract(motes_handle* h, double* data, const double* errs, size_t frame_stride, int F, int nspat,
                         int ndisp, const double* limits, const double* bin_params, const int32_t* nbins, int cap,
                         int wavelength_start, double* spectra, int32_t* lo_idx, int32_t* hi_idx, uint8_t* crmask,
                         const FrameState* st = nullptr) {
    StageTimer timer(h, ST_EXTRACT);
    size_t smem = (size_t)2 * nspat * kTileStride * sizeof(double);
    if (smem > h->smem_optin) return fail_arg("nspat too large for the extraction tile");
    MOTES_CUDA(cudaFuncSetAttribute(optimal_extract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    optimal_extract_kernel<<<dim3((ndisp + kTileCols - 1) / kTileCols, F), kTileThreads, smem, h->stream>>>(
        data, errs, frame_stride, nspat, ndisp, limits, bin_params, nbins, cap, wavelength_start, spectra, lo_idx,
        hi_idx, crmask, st);
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

template <class T>
static int upload(motes_handle* h, Scratch& s, const T* host, size_t count, T** dev) {
    MOTES_TRY(s.ensure(count * sizeof(T) + 16));
    *dev = s.as<T>();
    if (host && count) MOTES_CUDA(cudaMemcpyAsync(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return MOTES_OK;
}
template <class T>
static int download(motes_handle* h, T* host, const T* dev, size_t count) {
    if (host && dev && count) MOTES_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    return MOTES_OK;
}
template <class T>
static int copy_dd(motes_handle* h, T* dst, const T* src, size_t count) {
    if (dst && src && count) MOTES_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToDevice, h->stream));
    return MOTES_OK;
}

// The whole path on device-resident frames.  `emit` copies results into the caller's block
// (device-to-device for the _dev entry point, device-to-host for _host).
template <class Emit>
static int run_frames_core(motes_handle* h, Batch& b, double* data, double* errs, const uint8_t* qual,
                           const motes_params& p, const double* seeing_dev, const double* pixres_dev, uint32_t* rng,
                           Emit emit) {
    const size_t frame_stride = (size_t)b.nspat * b.ndisp;
    // seed FrameState with seeing / pixres
    MOTES_CUDA(cudaMemsetAsync(b.st, 0, sizeof(FrameState) * (size_t)b.F, h->stream));
    MOTES_CUDA(cudaMemcpy2DAsync(&b.st->seeing, sizeof(FrameState), seeing_dev, sizeof(double), sizeof(double), b.F,
                                 cudaMemcpyDeviceToDevice, h->stream));
    MOTES_CUDA(cudaMemcpy2DAsync(&b.st->pixres, sizeof(FrameState), pixres_dev, sizeof(double), sizeof(double), b.F,
                                 cudaMemcpyDeviceToDevice, h->stream));
    // a hoisted generator belongs to this call only, whichever way the call ends
    struct HoistScope { motes_handle* h; ~HoistScope() { h->boot.hoisted = false; } } hoist_scope{h};
    h->boot.hoisted = false;
    if (p.subtract_sky && p.sky_order == 0) MOTES_TRY(boot_hoist(h, b, rng));
    MOTES_TRY(stage_median_fit(h, b, data, frame_stride));
    if (p.subtract_sky) {
        MOTES_TRY(stage_localise(h, b, data, errs, frame_stride, qual, frame_stride, p.sky_snr_limit, p, 0));
        MOTES_TRY(stage_sky(h, b, data, errs, frame_stride, b.limits[0], p, rng));
    } else {
        MOTES_CUDA(cudaMemsetAsync(b.nbins[0], 0, sizeof(int32_t) * (size_t)b.F, h->stream));
    }
    MOTES_TRY(stage_localise(h, b, data, errs, frame_stride, qual, frame_stride, p.extraction_snr_limit, p, 1));
    MOTES_TRY(stage_extract(h, data, errs, frame_stride, b.F, b.nspat, b.ndisp, b.limits[1], b.params8[1], b.nbins[1],
                            b.cap, p.wavelength_start, b.spectra, nullptr, nullptr, nullptr, b.st));
    return emit();
}

// Copies one sub-batch's results into the caller's block at frame offset f0.
template <cudaMemcpyKind KIND>
static int emit_outputs(motes_handle* h, const Batch& b, const motes_params& p, const motes_outputs& o, int f0) {
    const size_t F = b.F, nd = b.ndisp, cp = b.cap, ns = b.nspat, at = (size_t)f0;
    cudaStream_t st = h->stream;
#define MOTES_PUT(dst, src, per_frame)                                                                              \
    do {                                                                                                            \
        if ((dst) && (src))                                                                                         \
            MOTES_CUDA(cudaMemcpyAsync((dst) + at * (per_frame), (src), F * (per_frame) * sizeof(*(dst)), KIND, st)); \
    } while (0)
    MOTES_PUT(o.spectra, b.spectra, 4 * nd);
    MOTES_PUT(o.ext_limits, b.limits[1], 2 * nd);
    MOTES_PUT(o.ext_params, b.params8[1], cp * 8);
    MOTES_PUT(o.ext_bins, b.bins_i[1], cp * 2);
    MOTES_PUT(o.ext_bin_snr, b.bins_snr[1], cp);
    MOTES_PUT(o.n_ext_bins, b.nbins[1], (size_t)1);
    MOTES_PUT(o.n_sky_bins, b.nbins[0], (size_t)1);
    if (p.subtract_sky) {
        MOTES_PUT(o.sky_limits, b.limits[0], 2 * nd);
        MOTES_PUT(o.sky_params, b.params8[0], cp * 8);
        MOTES_PUT(o.sky_bins, b.bins_i[0], cp * 2);
        MOTES_PUT(o.sky_bin_snr, b.bins_snr[0], cp);
        if (p.sky_order > 0) {
            MOTES_PUT(o.sky_model, b.model, nd * ns);
            MOTES_PUT(o.sky_uncertainty, b.model_err, nd * ns);
        } else {
            MOTES_PUT(o.sky_model, b.sky, nd);
            MOTES_PUT(o.sky_uncertainty, b.sky_err, nd);
        }
        MOTES_PUT(o.sky_flag, b.sky_flag, nd);
    }
#undef MOTES_PUT
    if (o.median_params) MOTES_CUDA(cudaMemcpy2DAsync(o.median_params + at * 6, 6 * sizeof(double), b.st->median_params, sizeof(FrameState), 6 * sizeof(double), F, KIND, st));
    if (o.scale) MOTES_CUDA(cudaMemcpy2DAsync(o.scale + at, sizeof(double), &b.st->scale, sizeof(FrameState), sizeof(double), F, KIND, st));
    if (o.seeing_px) MOTES_CUDA(cudaMemcpy2DAsync(o.seeing_px + at, sizeof(double), &b.st->seeing_px, sizeof(FrameState), sizeof(double), F, KIND, st));
    if (o.frame_status) MOTES_CUDA(cudaMemcpy2DAsync(o.frame_status + at, sizeof(int32_t), &b.st->status, sizeof(FrameState), sizeof(int32_t), F, KIND, st));
    return MOTES_OK;
}

static void swap_lanes(motes_handle* h) {
    std::swap(h->stream, h->lane1_stream);
    for (int i = 0; i < 24; ++i) std::swap(h->work[i], h->lane1_work[i]);
    std::swap(h->counters, h->lane1_counters);
}

static int get_event(motes_handle* h, size_t i, cudaEvent_t* ev) {
    while (h->events.size() <= i) {
        cudaEvent_t e;
        MOTES_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->events.push_back(e);
    }
    *ev = h->events[i];
    return MOTES_OK;
}

// Sub-batches per call (MOTES_B200_CHUNKS, default 1).  Measured on B200 with 128 frames of 4096 x 400 (round-1
// final kernels): 1 -> 182.4 ms, 2 -> 194.6 ms per step (earlier in the round 253 / 253 / 308 / 395 ms for 1 / 2 / 4 / 8):
// the throughput-bound kernels fill every SM, so the other lane's latency-bound kernels queue behind them instead of
// running beside them, while every per-frame serial stage is paid once per sub-batch.  The mechanism stays for
// callers whose batches exceed device memory.
static int chunk_frames(const motes_handle* h, int nframes) {
    int chunks = h->chunks > 0 ? h->chunks : 1;
    if (chunks > nframes) chunks = nframes;
    return (nframes + chunks - 1) / chunks;
}

// The whole path over device-resident frames [0, F), in sub-batches on alternating lanes.  `upload`
// (optional) brings a sub-batch's frames to the device on the copy stream first; `emit` returns its
// results.  On return the work is queued: lane 0 (h->stream) waits for lane 1.
template <class Upload, class Emit>
static int run_frames_chunked(motes_handle* h, double* data, double* errs, const uint8_t* qual, int nframes, int nspat,
                              int ndisp, int cap, const motes_params& p, const double* seeing_dev,
                              const double* pixres_dev, uint32_t* rng, Upload upload, Emit emit) {
    const int per = chunk_frames(h, nframes);
    const size_t frame_px = (size_t)nspat * ndisp;
    cudaEvent_t ev_start, ev_done;
    MOTES_TRY(get_event(h, 0, &ev_start));
    MOTES_TRY(get_event(h, 1, &ev_done));
    MOTES_CUDA(cudaEventRecord(ev_start, h->stream));
    MOTES_CUDA(cudaStreamWaitEvent(h->lane1_stream, ev_start, 0));
    MOTES_CUDA(cudaStreamWaitEvent(h->copy_stream, ev_start, 0));
    int rc = MOTES_OK;
    bool used_lane1 = false;
    for (int f0 = 0, k = 0; f0 < nframes && rc == MOTES_OK; f0 += per, ++k) {
        const int nf = (nframes - f0 < per) ? nframes - f0 : per;
        cudaEvent_t ev_up;
        rc = get_event(h, 2 + (size_t)k, &ev_up);
        if (rc != MOTES_OK) break;
        rc = upload(f0, nf, ev_up);                       // on h->copy_stream; records ev_up (or does nothing)
        if (rc != MOTES_OK) break;
        const bool lane1 = (k & 1) != 0 && nframes > per;
        if (lane1) { swap_lanes(h); used_lane1 = true; }
        rc = [&]() -> int {
            MOTES_CUDA(cudaStreamWaitEvent(h->stream, ev_up, 0));
            Batch b;
            MOTES_TRY(make_batch(h, b, nf, nspat, ndisp, cap, p.subtract_sky != 0, p.sky_order > 0));
            MOTES_TRY(run_frames_core(h, b, data + f0 * frame_px, errs + f0 * frame_px, qual + f0 * frame_px, p,
                                      seeing_dev + f0, pixres_dev + f0, rng ? rng + (size_t)f0 * 625 : nullptr,
                                      [&]() -> int { return emit(b, f0, nf); }));
            return MOTES_OK;
        }();
        if (lane1) swap_lanes(h);
    }
    if (used_lane1) {
        MOTES_CUDA(cudaEventRecord(ev_done, h->lane1_stream));
        MOTES_CUDA(cudaStreamWaitEvent(h->stream, ev_done, 0));
    }
    return rc;
}

}  // namespace motes

using namespace motes;

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char* motes_last_error(void) { return last_error_string().c_str(); }
const char* motes_version(void) { return "motes_b200 0.2 (sm_100a)"; }

int motes_create(int device, motes_handle** out) {
    if (!out) return fail_arg("motes_create: out is NULL");
    *out = nullptr;
    int count = 0;
    MOTES_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail_arg("motes_create: no such CUDA device");
    MOTES_CUDA(cudaSetDevice(device));
    motes_handle* h = new motes_handle();
    h->device = device;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete h; return fail_cuda(e, "cudaGetDeviceProperties", __FILE__, __LINE__); }
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    // A/B and cross-check switch: "classic" = one CTA walks a frame's whole stream (the literal replay)
    const char* boot = getenv("MOTES_B200_BOOTSTRAP");
    h->boot_mode = (boot && strcmp(boot, "classic") == 0) ? 1 : 0;
    const char* colw = getenv("MOTES_B200_COLUMN");
    if (colw) h->column_window = atoi(colw);
    const char* fitm = getenv("MOTES_B200_FIT");              // "warp": round-1 one-warp-per-fit kernel
    h->fit_mode = (fitm && strcmp(fitm, "warp") == 0) ? 1 : (fitm && strcmp(fitm, "group") == 0) ? 2 : 0;
    const char* fit_threads = getenv("MOTES_B200_FIT_THREADS");   // threads per fit of the lock-step sweep kernel: 32, 64, 128
    if (fit_threads && atoi(fit_threads) > 0) h->fit_threads = atoi(fit_threads);
    const char* fit_rounds = getenv("MOTES_B200_FIT_ROUNDS");     // lock-step rounds before the finishing kernel
    if (fit_rounds && atoi(fit_rounds) > 0) h->fit_rounds = atoi(fit_rounds);
    const char* fit_ctas = getenv("MOTES_B200_FIT_CTAS");         // cap on the sweep kernel's CTAs per SM (A/B: sharing the SMs with the generator)
    if (fit_ctas && atoi(fit_ctas) > 0) h->fit_ctas = atoi(fit_ctas);
    const char* fit_min = getenv("MOTES_B200_FIT_MIN");           // fits per call from which the lock-step rounds are used
    if (fit_min && atoll(fit_min) >= 0) h->fit_rounds_min = atoll(fit_min);
    const char* budget = getenv("MOTES_B200_STREAM_GB");
    if (budget && atof(budget) > 0.0) h->stream_budget = (size_t)(atof(budget) * 1073741824.0);
    const char* hoist = getenv("MOTES_B200_HOIST");            // "0": generate the bootstrap stream where it is consumed
    h->hoist = !(hoist && atoi(hoist) == 0);
    const char* hoist_ctas = getenv("MOTES_B200_HOIST_CTAS");
    if (hoist_ctas && atoi(hoist_ctas) > 0) h->hoist_ctas = atoi(hoist_ctas);
    const char* prof_hoist = getenv("MOTES_B200_PROFILE_HOIST");   // 1: per-stage timers leave the generator hoisted (main-stream stage times under overlap)
    h->profile_hoist = prof_hoist && atoi(prof_hoist) != 0;
    const char* walk = getenv("MOTES_B200_WALK");             // "table" / "stream": force one form of the position walk
    h->walk_mode = (walk && strcmp(walk, "table") == 0) ? 1 : (walk && strcmp(walk, "stream") == 0) ? 2 :
                   (walk && strcmp(walk, "serial") == 0) ? 3 : 0;
    const char* chunks = getenv("MOTES_B200_CHUNKS");
    if (chunks && atoi(chunks) > 0) h->chunks = atoi(chunks);
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->lane1_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return fail_cuda(e, "cudaStreamCreate", __FILE__, __LINE__); }
    *out = h;
    return MOTES_OK;
}

int motes_destroy(motes_handle* h) {
    if (!h) return MOTES_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto& s : h->stage) s.release();
    if (h->lane1_stream) cudaStreamSynchronize(h->lane1_stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    for (auto& s : h->work) s.release();
    for (auto& s : h->lane1_work) s.release();
    h->counters.release();
    h->lane1_counters.release();
    h->jump_polys.release();
    for (auto& ev : h->events) cudaEventDestroy(ev);
    if (h->lane1_stream) cudaStreamDestroy(h->lane1_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->owns_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MOTES_OK;
}

void* motes_stream(motes_handle* h) { return h ? (void*)h->stream : nullptr; }

int motes_set_stream(motes_handle* h, void* cuda_stream) {
    if (!h) return fail_arg("null handle");
    if (h->owns_stream && h->stream) {
        MOTES_CUDA(cudaStreamSynchronize(h->stream));
        MOTES_CUDA(cudaStreamDestroy(h->stream));
    }
    h->stream = (cudaStream_t)cuda_stream;
    h->owns_stream = false;
    return MOTES_OK;
}

int motes_synchronize(motes_handle* h) {
    if (!h) return fail_arg("null handle");
    MOTES_CUDA(cudaStreamSynchronize(h->copy_stream));
    MOTES_CUDA(cudaStreamSynchronize(h->lane1_stream));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    return MOTES_OK;
}

long long motes_launch_count(motes_handle* h) { return h ? h->launches : 0; }

int motes_profile_enable(motes_handle* h, int on) {
    if (!h) return fail_arg("null handle");
    h->profiling = on != 0;
    return MOTES_OK;
}

int motes_profile_stage_count(void) { return ST_COUNT; }

const char* motes_profile_stage_name(int stage) {
    static const char* names[ST_COUNT] = {"row_median", "median_fit", "column_stats", "snr_bins", "bin_profiles",
                                          "bin_fits", "limits", "sky_prepare", "sky_bootstrap", "sky_apply",
                                          "extract", "boot_jump", "boot_stream", "boot_walk", "boot_column"};
    return (stage >= 0 && stage < ST_COUNT) ? names[stage] : "";
}

int motes_profile_read(motes_handle* h, double* ms, long long* launches, int reset) {
    if (!h || !ms) return fail_arg("motes_profile_read: null argument");
    MOTES_CUDA(cudaSetDevice(h->device));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    for (auto& s : h->spans) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) h->stage_ms[s.stage] += t;
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    h->spans.clear();
    for (int i = 0; i < ST_COUNT; ++i) {
        ms[i] = h->stage_ms[i];
        if (launches) launches[i] = h->stage_launches[i];
        if (reset) { h->stage_ms[i] = 0; h->stage_launches[i] = 0; }
    }
    return MOTES_OK;
}

int motes_fp64_peak(motes_handle* h, double* tflops) {
    if (!h || !tflops) return fail_arg("motes_fp64_peak: null argument");
    MOTES_CUDA(cudaSetDevice(h->device));
    MOTES_TRY(h->counters.ensure(256));
    const int iters = 4096, ctas = h->sm_count * 8;
    const double flop = 2.0 * 4.0 * kPeakChains * (double)iters * (double)kPeakThreads * (double)ctas;
    cudaEvent_t a, b;
    MOTES_CUDA(cudaEventCreate(&a));
    MOTES_CUDA(cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        MOTES_CUDA(cudaEventRecord(a, h->stream));
        fp64_peak_kernel<<<ctas, kPeakThreads, 0, h->stream>>>(h->counters.as<double>(), iters, 0.999999, 1e-9);
        MOTES_LAUNCHED(h);
        MOTES_CUDA(cudaEventRecord(b, h->stream));
        MOTES_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        MOTES_CUDA(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms > 0.f && flop / (ms * 1e9) > best) best = flop / (ms * 1e9);    // first launch is warm-up
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return MOTES_OK;
}

int motes_bin_capacity(int ndisp, double minimum_column_limit) {
    if (ndisp < 1) return 0;
    if (!(minimum_column_limit >= 1.0)) return ndisp;
    long long cap = (long long)(ndisp / (floor(minimum_column_limit) + 1.0)) + 2;
    return (int)(cap < ndisp ? cap : ndisp);
}

int motes_seed_states(const uint32_t* seeds, int nframes, uint32_t* states) {
    if (!seeds || !states || nframes < 0) return fail_arg("motes_seed_states: bad argument");
    for (int f = 0; f < nframes; ++f) {
        mt_seed(states + (size_t)f * 625, seeds[f]);
        states[(size_t)f * 625 + 624] = kMtN;
    }
    return MOTES_OK;
}

int motes_qual_pack_dev(motes_handle* h, const double* qual, size_t count, uint8_t* out) {
    if (!h || !qual || !out) return fail_arg("motes_qual_pack_dev: null argument");
    MOTES_CUDA(cudaSetDevice(h->device));
    if (count == 0) return MOTES_OK;
    qual_pack_kernel<<<(unsigned)((count + 255) / 256), 256, 0, h->stream>>>(qual, count, out);
    MOTES_LAUNCHED(h);
    return MOTES_OK;
}

static int check_run_args(motes_handle* h, const void* data, const void* errs, const void* qual, int nframes, int nspat,
                          int ndisp, int cap, const motes_params* p, const void* seeing, const void* pixres,
                          const void* rng, const motes_outputs* out) {
    if (!h || !data || !errs || !qual || !p || !seeing || !pixres || !out) return fail_arg("motes_run_frames: null argument");
    if (nframes < 1 || nspat < 10 || ndisp < 2) return fail_arg("motes_run_frames: bad size");
    if (p->subtract_sky && !rng) return fail_arg("motes_run_frames: rng_states required when subtract_sky is set");
    if (cap < motes_bin_capacity(ndisp, p->minimum_column_limit)) return fail_arg("motes_run_frames: cap below motes_bin_capacity()");
    return MOTES_OK;
}

int motes_run_frames_dev(motes_handle* h, double* data, double* errs, const uint8_t* qual, int nframes, int nspat,
                         int ndisp, int cap, const motes_params* params, const double* seeing, const double* pixres,
                         uint32_t* rng_states, const motes_outputs* out) {
    MOTES_TRY(check_run_args(h, data, errs, qual, nframes, nspat, ndisp, cap, params, seeing, pixres, rng_states, out));
    MOTES_CUDA(cudaSetDevice(h->device));
    const motes_params p = *params;
    const motes_outputs o = *out;
    auto upload = [&](int, int, cudaEvent_t ev) -> int {
        MOTES_CUDA(cudaEventRecord(ev, h->copy_stream));      // nothing to upload: the event is immediately reachable
        return MOTES_OK;
    };
    auto emit = [&](const Batch& b, int f0, int) -> int { return emit_outputs<cudaMemcpyDeviceToDevice>(h, b, p, o, f0); };
    return run_frames_chunked(h, data, errs, qual, nframes, nspat, ndisp, cap, p, seeing, pixres, rng_states, upload, emit);
}

// `f32`: data / errs are float32 planes (`swap`: big-endian) that are widened on the device after a 4-byte-per-pixel
// upload; otherwise float64.
static int run_frames_host_enqueue(motes_handle* h, void* data_v, void* errs_v, const uint8_t* qual, int nframes, int nspat,
                                   int ndisp, int cap, const motes_params* params, const double* seeing, const double* pixres,
                                   uint32_t* rng_states, int copy_back, const motes_outputs* out, bool f32 = false,
                                   bool swap = false) {
    MOTES_TRY(check_run_args(h, data_v, errs_v, qual, nframes, nspat, ndisp, cap, params, seeing, pixres, rng_states, out));
    if (f32 && copy_back) return fail_arg("motes_run_frames_host_f32: copy_back is not available for float32 frames");
    MOTES_CUDA(cudaSetDevice(h->device));
    const motes_params p = *params;
    const motes_outputs o = *out;
    const size_t frame_px = (size_t)nspat * ndisp;
    const size_t npix = (size_t)nframes * frame_px;
    double* data = static_cast<double*>(data_v);
    double* errs = static_cast<double*>(errs_v);
    double *d_data, *d_errs, *d_seeing, *d_pixres;
    uint8_t* d_qual;
    uint32_t* d_rng = nullptr;
    uint32_t *d_data32 = nullptr, *d_errs32 = nullptr;
    // frames go up sub-batch by sub-batch on the copy stream (below); the small per-frame arrays go now
    MOTES_TRY(upload<double>(h, h->stage[0], nullptr, npix, &d_data));
    MOTES_TRY(upload<double>(h, h->stage[1], nullptr, npix, &d_errs));
    MOTES_TRY(upload<uint8_t>(h, h->stage[2], nullptr, npix, &d_qual));
    MOTES_TRY(upload(h, h->stage[3], seeing, (size_t)nframes, &d_seeing));
    MOTES_TRY(upload(h, h->stage[4], pixres, (size_t)nframes, &d_pixres));
    if (rng_states) MOTES_TRY(upload(h, h->stage[5], rng_states, (size_t)nframes * 625, &d_rng));
    if (f32) {
        MOTES_TRY(upload<uint32_t>(h, h->stage[8], nullptr, npix, &d_data32));
        MOTES_TRY(upload<uint32_t>(h, h->stage[9], nullptr, npix, &d_errs32));
    }
    auto upload_chunk = [&](int f0, int nf, cudaEvent_t ev) -> int {
        const size_t off = (size_t)f0 * frame_px, cnt = (size_t)nf * frame_px;
        if (f32) {
            const uint32_t* h_data = static_cast<const uint32_t*>(data_v);
            const uint32_t* h_errs = static_cast<const uint32_t*>(errs_v);
            MOTES_CUDA(cudaMemcpyAsync(d_data32 + off, h_data + off, cnt * 4, cudaMemcpyHostToDevice, h->copy_stream));
            widen_f32_kernel<<<(unsigned)((cnt / 4 + 256) / 256), 256, 0, h->copy_stream>>>(d_data32 + off, cnt, swap ? 1 : 0, d_data + off);
            MOTES_LAUNCHED(h);
            MOTES_CUDA(cudaMemcpyAsync(d_errs32 + off, h_errs + off, cnt * 4, cudaMemcpyHostToDevice, h->copy_stream));
            widen_f32_kernel<<<(unsigned)((cnt / 4 + 256) / 256), 256, 0, h->copy_stream>>>(d_errs32 + off, cnt, swap ? 1 : 0, d_errs + off);
            MOTES_LAUNCHED(h);
        } else {
            MOTES_CUDA(cudaMemcpyAsync(d_data + off, data + off, cnt * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
            MOTES_CUDA(cudaMemcpyAsync(d_errs + off, errs + off, cnt * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
        }
        MOTES_CUDA(cudaMemcpyAsync(d_qual + off, qual + off, cnt, cudaMemcpyHostToDevice, h->copy_stream));
        MOTES_CUDA(cudaEventRecord(ev, h->copy_stream));
        return MOTES_OK;
    };
    auto emit = [&](const Batch& b, int f0, int nf) -> int {
        MOTES_TRY(emit_outputs<cudaMemcpyDeviceToHost>(h, b, p, o, f0));
        if (p.subtract_sky && rng_states)
            MOTES_CUDA(cudaMemcpyAsync(rng_states + (size_t)f0 * 625, d_rng + (size_t)f0 * 625, (size_t)nf * 625 * sizeof(uint32_t),
                                       cudaMemcpyDeviceToHost, h->stream));
        if (copy_back) {
            const size_t off = (size_t)f0 * frame_px, cnt = (size_t)nf * frame_px;
            MOTES_CUDA(cudaMemcpyAsync(data + off, d_data + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            MOTES_CUDA(cudaMemcpyAsync(errs + off, d_errs + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        }
        return MOTES_OK;
    };
    h->upload_pending = true;
    const int rc = run_frames_chunked(h, d_data, d_errs, d_qual, nframes, nspat, ndisp, cap, p, d_seeing, d_pixres, d_rng,
                                      upload_chunk, emit);
    h->upload_pending = false;
    return rc;
}

static int finish_host_call(motes_handle* h, int rc) {
    // the caller's buffers must not be touched after return, whatever happened
    if (h) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->lane1_stream);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (rc == MOTES_OK && e != cudaSuccess) return fail_cuda(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    }
    return rc;
}

int motes_run_frames_host_f32(motes_handle* h, const float* data, const float* errs, const uint8_t* qual, int nframes,
                              int nspat, int ndisp, int cap, const motes_params* params, const double* seeing,
                              const double* pixres, uint32_t* rng_states, int big_endian, int wait, const motes_outputs* out) {
    const int rc = run_frames_host_enqueue(h, const_cast<float*>(data), const_cast<float*>(errs), qual, nframes, nspat, ndisp, cap,
                                           params, seeing, pixres, rng_states, 0, out, true, big_endian != 0);
    return wait ? finish_host_call(h, rc) : rc;
}

int motes_run_frames_host(motes_handle* h, double* data, double* errs, const uint8_t* qual, int nframes, int nspat,
                          int ndisp, int cap, const motes_params* params, const double* seeing, const double* pixres,
                          uint32_t* rng_states, int copy_back, const motes_outputs* out) {
    return finish_host_call(h, run_frames_host_enqueue(h, data, errs, qual, nframes, nspat, ndisp, cap, params, seeing, pixres,
                                                       rng_states, copy_back, out));
}

int motes_run_frames_host_async(motes_handle* h, double* data, double* errs, const uint8_t* qual, int nframes, int nspat,
                                int ndisp, int cap, const motes_params* params, const double* seeing, const double* pixres,
                                uint32_t* rng_states, int copy_back, const motes_outputs* out) {
    return run_frames_host_enqueue(h, data, errs, qual, nframes, nspat, ndisp, cap, params, seeing, pixres, rng_states,
                                   copy_back, out);
}

// ------------------------------------------------------------------------------------------ stage entry points
int motes_row_median_host(motes_handle* h, const double* data, int nspat, int ndisp, double* out) {
    if (!h || !data || !out || nspat < 1 || ndisp < 1) return fail_arg("motes_row_median: bad argument");
    MOTES_CUDA(cudaSetDevice(h->device));
    double *d_data, *d_out;
    MOTES_TRY(upload(h, h->stage[0], data, (size_t)nspat * ndisp, &d_data));
    MOTES_TRY(upload<double>(h, h->stage[1], nullptr, (size_t)nspat, &d_out));
    MOTES_TRY(stage_row_median(h, d_data, (size_t)nspat * ndisp, 1, nspat, ndisp, d_out));
    MOTES_TRY(download(h, out, d_out, (size_t)nspat));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    return MOTES_OK;
}

int motes_moffat_fit_batch_dev(motes_handle* h, const double* profiles, int nfit, int nspat, const double* alpha0,
                               double* params, double* cost, int32_t* nfev, int32_t* njev, int32_t* status) {
    if (!h || !profiles || !alpha0 || !params) return fail_arg("motes_moffat_fit_batch: null argument");
    if (nfit < 0 || nspat < 1) return fail_arg("motes_moffat_fit_batch: bad size");
    MOTES_CUDA(cudaSetDevice(h->device));
    FitBatch fb{};
    fb.profiles = profiles; fb.m = nspat; fb.nframes = 1; fb.bound = nfit; fb.cap = nfit; fb.alpha0 = alpha0;
    fb.params = params; fb.cost = cost; fb.nfev = nfev; fb.njev = njev; fb.status = status;
    return stage_fit(h, fb);
}

int motes_moffat_fit_batch_host(motes_handle* h, const double* profiles, int nfit, int nspat, const double* alpha0,
                                double* params, double* cost, int32_t* nfev, int32_t* njev, int32_t* status) {
    if (!h || !profiles || !alpha0 || !params) return fail_arg("motes_moffat_fit_batch: null argument");
    if (nfit < 0 || nspat < 1) return fail_arg("motes_moffat_fit_batch: bad size");
    if (nfit == 0) return MOTES_OK;
    MOTES_CUDA(cudaSetDevice(h->device));
    double *d_prof, *d_a0, *d_par, *d_cost;
    int32_t* d_int;
    MOTES_TRY(upload(h, h->stage[0], profiles, (size_t)nfit * nspat, &d_prof));
    MOTES_TRY(upload(h, h->stage[1], alpha0, (size_t)nfit, &d_a0));
    MOTES_TRY(upload<double>(h, h->stage[2], nullptr, (size_t)nfit * kParams, &d_par));
    MOTES_TRY(upload<double>(h, h->stage[3], nullptr, (size_t)nfit, &d_cost));
    MOTES_TRY(upload<int32_t>(h, h->stage[4], nullptr, (size_t)nfit * 3, &d_int));
    MOTES_TRY(motes_moffat_fit_batch_dev(h, d_prof, nfit, nspat, d_a0, d_par, d_cost, d_int, d_int + nfit,
                                         d_int + 2 * (size_t)nfit));
    MOTES_TRY(download(h, params, d_par, (size_t)nfit * kParams));
    MOTES_TRY(download(h, cost, d_cost, (size_t)nfit));
    MOTES_TRY(download(h, nfev, d_int, (size_t)nfit));
    MOTES_TRY(download(h, njev, d_int + nfit, (size_t)nfit));
    MOTES_TRY(download(h, status, d_int + 2 * (size_t)nfit, (size_t)nfit));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    return MOTES_OK;
}

int motes_get_bins_host(motes_handle* h, const double* data, const double* errs, const uint8_t* qual, int nspat,
                        int ndisp, int row_lo, int row_hi, double snr_limit, double minimum_column_limit, int cap,
                        int32_t* bins, double* snr, int32_t* nbins, uint8_t* gap) {
    if (!h || !data || !errs || !qual || !bins || !snr || !nbins) return fail_arg("motes_get_bins: null argument");
    if (nspat < 1 || ndisp < 1) return fail_arg("motes_get_bins: bad size");
    if (cap < motes_bin_capacity(ndisp, minimum_column_limit)) return fail_arg("motes_get_bins: cap below motes_bin_capacity()");
    MOTES_CUDA(cudaSetDevice(h->device));
    const size_t npix = (size_t)nspat * ndisp;
    Batch b;
    MOTES_TRY(make_batch(h, b, 1, nspat, ndisp, cap, false));
    double *d_data, *d_errs;
    uint8_t* d_qual;
    MOTES_TRY(upload(h, h->stage[0], data, npix, &d_data));
    MOTES_TRY(upload(h, h->stage[1], errs, npix, &d_errs));
    MOTES_TRY(upload(h, h->stage[2], qual, npix, &d_qual));
    FrameState st{};
    st.row_lo = row_lo;
    st.row_hi = row_hi;
    MOTES_CUDA(cudaMemcpyAsync(b.st, &st, sizeof(st), cudaMemcpyHostToDevice, h->stream));
    MOTES_TRY(stage_column_stats(h, b, d_data, d_errs, npix));
    MOTES_TRY(stage_bins(h, b, d_qual, npix, snr_limit, minimum_column_limit, 0));
    MOTES_TRY(download(h, nbins, b.nbins[0], 1));
    MOTES_TRY(download(h, bins, b.bins_i[0], (size_t)2 * cap));
    MOTES_TRY(download(h, snr, b.bins_snr[0], (size_t)cap));
    MOTES_TRY(download(h, gap, b.gap, (size_t)ndisp));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    return MOTES_OK;
}

int motes_bin_profiles_host(motes_handle* h, const double* data, int nspat, int ndisp, const int32_t* bins, int nbins,
                            double* profiles) {
    if (!h || !data || !bins || !profiles) return fail_arg("motes_bin_profiles: null argument");
    if (nspat < 1 || ndisp < 1 || nbins < 1) return fail_arg("motes_bin_profiles: bad size");
    MOTES_CUDA(cudaSetDevice(h->device));
    const size_t npix = (size_t)nspat * ndisp;
    Batch b;
    MOTES_TRY(make_batch(h, b, 1, nspat, ndisp, nbins, false));
    double* d_data;
    MOTES_TRY(upload(h, h->stage[0], data, npix, &d_data));
    FrameState st{};
    MOTES_CUDA(cudaMemcpyAsync(b.st, &st, sizeof(st), cudaMemcpyHostToDevice, h->stream));
    MOTES_TRY(stage_column_stats(h, b, d_data, d_data, npix));   // only the chip-gap flags are used
    MOTES_CUDA(cudaMemcpyAsync(b.bins_i[0], bins, sizeof(int32_t) * 2 * (size_t)nbins, cudaMemcpyHostToDevice, h->stream));
    int32_t nb = nbins;
    MOTES_CUDA(cudaMemcpyAsync(b.nbins[0], &nb, sizeof(nb), cudaMemcpyHostToDevice, h->stream));
    MOTES_TRY(stage_bin_profiles(h, b, d_data, npix, 0));
    MOTES_TRY(download(h, profiles, b.profiles, (size_t)nbins * nspat));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    return MOTES_OK;
}

int motes_interp_limits_host(motes_handle* h, const double* table, int nbins, int ndisp, double* lim_lo, double* lim_hi) {
    if (!h || !table || !lim_lo || !lim_hi || nbins < 1 || ndisp < 1) return fail_arg("motes_interp_limits: bad argument");
    MOTES_CUDA(cudaSetDevice(h->device));
    Batch b;
    MOTES_TRY(make_batch(h, b, 1, 16, ndisp, nbins, false));
    MOTES_CUDA(cudaMemcpyAsync(b.table, table, sizeof(double) * 4 * (size_t)nbins, cudaMemcpyHostToDevice, h->stream));
    int32_t nb = nbins;
    MOTES_CUDA(cudaMemcpyAsync(b.nbins[0], &nb, sizeof(nb), cudaMemcpyHostToDevice, h->stream));
    limits_kernel<<<1, 256, 0, h->stream>>>(b.table, nbins, b.nbins[0], ndisp, b.limits[0], b.inner);
    MOTES_LAUNCHED(h);
    MOTES_TRY(download(h, lim_lo, b.limits[0], (size_t)ndisp));
    MOTES_TRY(download(h, lim_hi, b.limits[0] + ndisp, (size_t)ndisp));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    return MOTES_OK;
}

int motes_sky_subtract_host(motes_handle* h, double* data, double* errs, int nspat, int ndisp, double* lim_lo,
                            double* lim_hi, int spatial_floor, int sky_order, uint32_t* rng_state, double* sky,
                            double* sky_err, int32_t* flag, int32_t* n_sky, int32_t* n_good) {
    if (!h || !data || !errs || !lim_lo || !lim_hi || !rng_state) return fail_arg("motes_sky_subtract: null argument");
    if (nspat < 1 || ndisp < 1) return fail_arg("motes_sky_subtract: bad size");
    MOTES_CUDA(cudaSetDevice(h->device));
    const size_t npix = (size_t)nspat * ndisp;
    Batch b;
    MOTES_TRY(make_batch(h, b, 1, nspat, ndisp, 1, true, sky_order > 0));
    double *d_data, *d_errs;
    uint32_t* d_rng;
    MOTES_TRY(upload(h, h->stage[0], data, npix, &d_data));
    MOTES_TRY(upload(h, h->stage[1], errs, npix, &d_errs));
    MOTES_TRY(upload(h, h->stage[5], rng_state, (size_t)625, &d_rng));
    MOTES_CUDA(cudaMemcpyAsync(b.limits[0], lim_lo, sizeof(double) * ndisp, cudaMemcpyHostToDevice, h->stream));
    MOTES_CUDA(cudaMemcpyAsync(b.limits[0] + ndisp, lim_hi, sizeof(double) * ndisp, cudaMemcpyHostToDevice, h->stream));
    FrameState st{};
    MOTES_CUDA(cudaMemcpyAsync(b.st, &st, sizeof(st), cudaMemcpyHostToDevice, h->stream));
    motes_params p{};
    p.sky_order = sky_order;
    p.spatial_floor = spatial_floor;
    MOTES_TRY(stage_sky(h, b, d_data, d_errs, npix, b.limits[0], p, d_rng));
    MOTES_CUDA(cudaMemcpyAsync(&st, b.st, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
    MOTES_TRY(download(h, lim_lo, b.limits[0], (size_t)ndisp));
    MOTES_TRY(download(h, lim_hi, b.limits[0] + ndisp, (size_t)ndisp));
    MOTES_TRY(download(h, n_sky, b.n_sky, (size_t)ndisp));
    MOTES_TRY(download(h, n_good, b.n_good, (size_t)ndisp));
    MOTES_TRY(download(h, flag, b.sky_flag, (size_t)ndisp));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    if (st.status == MOTES_E_NOSKY) {
        last_error_string() = "No pixels contained inside sky region.";
        return MOTES_E_NOSKY;
    }
    MOTES_TRY(download(h, data, d_data, npix));
    MOTES_TRY(download(h, errs, d_errs, npix));
    if (sky_order > 0) {
        MOTES_TRY(download(h, sky, b.model, npix));
        MOTES_TRY(download(h, sky_err, b.model_err, npix));
    } else {
        MOTES_TRY(download(h, sky, b.sky, (size_t)ndisp));
        MOTES_TRY(download(h, sky_err, b.sky_err, (size_t)ndisp));
    }
    MOTES_TRY(download(h, rng_state, d_rng, (size_t)625));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    return MOTES_OK;
}

int motes_optimal_extract_host(motes_handle* h, double* data, const double* errs, int nspat, int ndisp,
                               const double* lim_lo, const double* lim_hi, const double* bin_params, int nbins,
                               int wavelength_start, double* spectra, int32_t* lo_idx, int32_t* hi_idx, uint8_t* crmask) {
    if (!h || !data || !errs || !lim_lo || !lim_hi || !bin_params || !spectra)
        return fail_arg("motes_optimal_extract: null argument");
    if (nspat < 1 || ndisp < 1 || nbins < 0) return fail_arg("motes_optimal_extract: bad size");
    MOTES_CUDA(cudaSetDevice(h->device));
    size_t npix = (size_t)nspat * ndisp;
    double *d_data, *d_errs, *d_lim, *d_bins, *d_spec;
    int32_t* d_idx;
    uint8_t* d_mask;
    MOTES_TRY(upload(h, h->stage[0], data, npix, &d_data));
    MOTES_TRY(upload(h, h->stage[1], errs, npix, &d_errs));
    MOTES_TRY(upload<double>(h, h->stage[2], nullptr, (size_t)2 * ndisp, &d_lim));
    MOTES_CUDA(cudaMemcpyAsync(d_lim, lim_lo, sizeof(double) * ndisp, cudaMemcpyHostToDevice, h->stream));
    MOTES_CUDA(cudaMemcpyAsync(d_lim + ndisp, lim_hi, sizeof(double) * ndisp, cudaMemcpyHostToDevice, h->stream));
    MOTES_TRY(upload(h, h->stage[4], bin_params, (size_t)nbins * 8, &d_bins));
    MOTES_TRY(upload<double>(h, h->stage[5], nullptr, (size_t)4 * ndisp, &d_spec));
    MOTES_TRY(upload<int32_t>(h, h->stage[6], nullptr, (size_t)2 * ndisp + 1, &d_idx));
    MOTES_TRY(upload<uint8_t>(h, h->stage[7], nullptr, crmask ? npix : 0, &d_mask));
    int32_t nb = nbins;
    MOTES_CUDA(cudaMemcpyAsync(d_idx + 2 * (size_t)ndisp, &nb, sizeof(nb), cudaMemcpyHostToDevice, h->stream));
    MOTES_TRY(stage_extract(h, d_data, d_errs, npix, 1, nspat, ndisp, d_lim, d_bins, d_idx + 2 * (size_t)ndisp, nbins,
                            wavelength_start, d_spec, d_idx, d_idx + ndisp, crmask ? d_mask : nullptr));
    MOTES_TRY(download(h, data, d_data, npix));
    MOTES_TRY(download(h, spectra, d_spec, (size_t)4 * ndisp));
    MOTES_TRY(download(h, lo_idx, d_idx, (size_t)ndisp));
    MOTES_TRY(download(h, hi_idx, d_idx + ndisp, (size_t)ndisp));
    MOTES_TRY(download(h, crmask, d_mask, crmask ? npix : 0));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    return MOTES_OK;
}

}  // extern "C"
