// Library handle: device, stream, growable device scratch, launch accounting.
#pragma once

#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
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
        size_t want = need + need / 4 + 256;
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

struct motes_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    int sm_count = 148;
    size_t smem_optin = 0;
    long long launches = 0;
    // scratch slots: host-API staging and per-stage work areas
    motes::Scratch stage[12];
    motes::Scratch work[12];
    motes::Scratch counters;    // small zero-initialised control block
};
