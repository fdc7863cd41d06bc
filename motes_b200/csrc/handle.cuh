// Library handle: device, stream, growable device scratch, launch accounting.
#pragma once

#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/motes_b200.h"

namespace motes {

inline std::string& last_error_string() {
    static thread_local std::string s;
    return s;
}

inline int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    last_error_string() = buf;
    return MOTES_E_CUDA;
}

#define MOTES_CUDA(call)                                                              \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) return ::motes::fail_cuda(e_, #call, __FILE__, __LINE__); \
    } while (0)

#define MOTES_TRY(call)              \
    do {                             \
        int rc_ = (call);            \
        if (rc_ != MOTES_OK) return rc_; \
    } while (0)

inline int fail_arg(const char* msg) {
    last_error_string() = msg;
    return MOTES_E_ARG;
}

// A device buffer that only ever grows; contents are scratch.
struct Scratch {
    void* ptr = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return MOTES_OK;
        if (ptr) {
            MOTES_CUDA(cudaFree(ptr));
            ptr = nullptr;
            bytes = 0;
        }
        size_t slack = need / 4;                      // room to grow without reallocating - but not on multi-GB buffers
        if (slack > ((size_t)64 << 20)) slack = (size_t)64 << 20;
        size_t want = need + slack + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e != cudaSuccess) {
            ptr = nullptr;
            last_error_string() = std::string("cudaMalloc failed: ") + cudaGetErrorString(e);
            return MOTES_E_NOMEM;
        }
        bytes = want;
        return MOTES_OK;
    }
    template <class T>
    T* as() const { return static_cast<T*>(ptr); }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
};

}  // namespace motes

namespace motes {
// Optional per-stage timing with CUDA events on the handle's stream (bench.py's roofline leg).
enum Stage { ST_ROW_MEDIAN = 0, ST_MEDIAN_FIT, ST_COLUMN_STATS, ST_BINS, ST_BIN_PROFILES, ST_BIN_FITS, ST_LIMITS,
             ST_SKY_PREPARE, ST_SKY_BOOTSTRAP, ST_SKY_APPLY, ST_EXTRACT,
             ST_BOOT_JUMP, ST_BOOT_STREAM, ST_BOOT_WALK, ST_BOOT_COLUMN, ST_COUNT };
struct StageSpan { cudaEvent_t a, b; int stage; };
}  // namespace motes

namespace motes {
struct SegLayout {                 // segmented generator replay (skyseg.cuh): per call, host-computed
    int log2_stride;
    uint32_t stride;               // words per segment
    int segs;                      // segments allocated per frame (K)
    unsigned long long stream_len; // 16-bit words per frame (padded for the walk's prefetch)
    unsigned long long sub_len;    // running-count entries per frame
    int snaps;                     // snapshots per frame
};
struct BootPlan {                  // how a batch's order-0 bootstrap is laid out (motes_b200.cu: boot_plan)
    bool ok = false;
    bool hoisted = false;          // the generator part is already running on the second stream
    SegLayout lay{};
    int wave = 0;                  // frames per wave
    int log2_stride = 0;
};
}  // namespace motes

struct motes_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    int sm_count = 148;
    size_t smem_optin = 0;
    long long launches = 0;
    // scratch slots: host-API staging and per-stage work areas
    motes::Scratch stage[12];
    motes::Scratch work[24];
    motes::Scratch counters;    // small zero-initialised control block
    // second lane: sub-batches of one call alternate between two streams (each with its own work areas) so
    // that the latency-bound stages of one overlap the throughput-bound stages of the other
    cudaStream_t lane1_stream = nullptr;
    motes::Scratch lane1_work[24];
    motes::Scratch lane1_counters;
    cudaStream_t copy_stream = nullptr;     // host-to-device uploads of the next sub-batch
    std::vector<cudaEvent_t> events;
    int chunks = 0;              // MOTES_B200_CHUNKS: sub-batches per call (0 = auto)
    bool profiling = false;
    bool profile_hoist = false;  // MOTES_B200_PROFILE_HOIST
    int fit_mode = 0;            // MOTES_B200_FIT: 0 auto (lock-step rounds for large batches, else the group kernel), 1 "warp", 2 "group"
    int fit_threads = 0;         // MOTES_B200_FIT_THREADS: threads per fit of the lock-step sweep kernel (0 = by nspat)
    int fit_ctas = 0;            // MOTES_B200_FIT_CTAS: cap on the sweep kernel's CTAs per SM (0 = what fits)
    int fit_rounds = 40;         // MOTES_B200_FIT_ROUNDS: lock-step rounds before the finishing kernel takes the stragglers
    long long fit_rounds_min = 2048;   // MOTES_B200_FIT_MIN: fits per call from which the lock-step rounds are used
    int column_window = -1;      // MOTES_B200_COLUMN: histogram window of the bootstrap column kernel (-1 = auto, 0 = exact slot kernel only)
    int walk_mode = 0;           // MOTES_B200_WALK: 0 auto, 1 "table" (walktab.cuh, by runs), 2 "stream" (sky_walk_kernel), 3 "serial" (table, one column at a time)
    int boot_mode = 0;           // MOTES_B200_BOOTSTRAP: 0 segmented (default), 1 classic (one CTA per frame)
    motes::Scratch jump_polys;   // device copy of the MT19937 jump polynomial table (mtjump.hpp)
    bool jump_ready = false;
    size_t stream_budget = 0;    // bytes the segmented bootstrap may use for generator streams (0 = auto)
    bool hoist = true;           // MOTES_B200_HOIST=0: generate the bootstrap stream where it is consumed
    int hoist_ctas = 8;          // MOTES_B200_HOIST_CTAS: generator CTAs per SM when it runs beside the localisation stages
    motes::BootPlan boot;
    cudaEvent_t ev_hoist[2] = {nullptr, nullptr};
    bool upload_pending = false; // the frames of the running call are still being uploaded (host entry points)
    std::vector<motes::StageSpan> spans;
    double stage_ms[motes::ST_COUNT] = {0};
    long long stage_launches[motes::ST_COUNT] = {0};
};

namespace motes {
struct StageTimer {
    motes_handle* h;
    int idx = -1;
    StageTimer(motes_handle* handle, int stage) : h(handle) {
        if (!h->profiling || stage < 0) return;
        StageSpan s;
        s.stage = stage;
        if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
        cudaEventRecord(s.a, h->stream);
        h->spans.push_back(s);
        idx = (int)h->spans.size() - 1;
        h->stage_launches[stage] += 1;
    }
    void stop() {
        if (idx >= 0) cudaEventRecord(h->spans[idx].b, h->stream);
        idx = -1;
    }
    ~StageTimer() { stop(); }
};
}  // namespace motes
