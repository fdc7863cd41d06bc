// libmotes_b200.so: CUDA kernels (sm_100a) and the C ABI declared in include/motes_b200.h.
#include "handle.cuh"
#include "trf.cuh"
#include "extract.cuh"

namespace motes {

// ------------------------------------------------------------------------------------------
// a3: one warp per Moffat fit, slab in shared memory, work handed out by an atomic counter.
// ------------------------------------------------------------------------------------------
constexpr int kFitMaxWarps = 8;

__global__ void __launch_bounds__(kFitMaxWarps * 32)
moffat_fit_kernel(const double* __restrict__ profiles, int nfit, int m, const double* __restrict__ alpha0,
                  double* __restrict__ params, double* __restrict__ cost, int32_t* __restrict__ nfev,
                  int32_t* __restrict__ njev, int32_t* __restrict__ status, unsigned int* work_counter) {
    extern __shared__ __align__(16) double fit_smem[];
    const int warp = threadIdx.x >> 5;
    WarpCtx ctx{(int)(threadIdx.x & 31)};
    const size_t per_warp = fit_scratch_doubles(m);
    double* slab = fit_smem + per_warp * warp;
    FitScratch ws;
    ws.y = slab;
    ws.unit = ws.y + m;
    ws.a = ws.unit + m;
    ws.w = ws.a + 7 * (size_t)m;
    ws.v = ws.w + 36;
    while (true) {
        unsigned int f = 0;
        if (ctx.lane == 0) f = atomicAdd(work_counter, 1u);
        f = __shfl_sync(0xffffffffu, f, 0);
        if (f >= (unsigned int)nfit) break;
        FitResult r;
        trf_fit_moffat(ctx, m, profiles + (size_t)f * m, alpha0[f], ws, &r);
        if (ctx.lane == 0) {
            for (int k = 0; k < kParams; ++k) params[(size_t)f * kParams + k] = r.x[k];
            if (cost) cost[f] = r.cost;
            if (nfev) nfev[f] = r.nfev;
            if (njev) njev[f] = r.njev;
            if (status) status[f] = r.status;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// a10: extraction.  A CTA stages a tile of kExtCols consecutive columns (all nspat rows of data
// and errs) through shared memory with coalesced row-segment loads, then each warp extracts
// columns of the tile.  Row stride in shared memory is kExtCols+1 doubles (odd -> no bank
// conflicts when a warp walks down a column).
// ------------------------------------------------------------------------------------------
constexpr int kExtCols = 16;
constexpr int kExtStride = kExtCols + 1;
constexpr int kExtThreads = 256;

__global__ void __launch_bounds__(kExtThreads)
optimal_extract_kernel(double* __restrict__ data, const double* __restrict__ errs, int nspat, int ndisp,
                       const double* __restrict__ lim_lo, const double* __restrict__ lim_hi,
                       const double* __restrict__ bin_params, int nbins, int wavelength_start,
                       double* __restrict__ spectra, int32_t* __restrict__ lo_idx, int32_t* __restrict__ hi_idx,
                       uint8_t* __restrict__ crmask) {
    extern __shared__ __align__(16) double ext_smem[];
    double* sd = ext_smem;
    double* se = ext_smem + (size_t)nspat * kExtStride;
    const int col0 = blockIdx.x * kExtCols;
    const int ncols = min(kExtCols, ndisp - col0);
    // stage + scrub (extraction.py:31-34, 86)
    for (int idx = threadIdx.x; idx < nspat * kExtCols; idx += blockDim.x) {
        int row = idx / kExtCols, c = idx % kExtCols;
        if (c < ncols) {
            size_t g = (size_t)row * ndisp + col0 + c;
            double d = data[g];
            double e = fabs(errs[g]);
            if (!is_finite(e)) e = 0.0;
            if (!is_finite(d)) { d = 0.0; data[g] = 0.0; }
            sd[row * kExtStride + c] = d;
            se[row * kExtStride + c] = e;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    WarpCtx ctx{(int)(threadIdx.x & 31)};
    for (int c = warp; c < ncols; c += kExtThreads / 32) {
        int col = col0 + c;
        // the bin whose parameters the reference holds when it reaches this column
        // (extraction.py:104-112): last bin b with first_col(b) <= col, stepping only through
        // consecutive bins; bins tile the axis, so that is the bin containing the column.
        int which = -1;
        {
            int lo_b = 0, hi_b = nbins - 1;
            double key = (double)(col + wavelength_start);
            while (lo_b <= hi_b) {
                int mid = (lo_b + hi_b) >> 1;
                if (bin_params[(size_t)mid * 8 + 6] <= key) { which = mid; lo_b = mid + 1; }
                else hi_b = mid - 1;
            }
        }
        double zero_bin[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const double* bin = (which >= 0) ? bin_params + (size_t)which * 8 : zero_bin;
        ColumnSpectrum out;
        extract_column(ctx, sd + c, se + c, kExtStride, nspat, lim_lo[col], lim_hi[col], bin, &out,
                       crmask ? crmask + (size_t)col * nspat : nullptr);
        if (ctx.lane == 0) {
            spectra[col] = out.optimal;
            spectra[(size_t)ndisp + col] = out.optimal_err;
            spectra[(size_t)2 * ndisp + col] = out.aperture;
            spectra[(size_t)3 * ndisp + col] = out.aperture_err;
            if (lo_idx) lo_idx[col] = out.lo;
            if (hi_idx) hi_idx[col] = out.hi;
        }
    }
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
static int launch_fit(motes_handle* h, const double* profiles, int nfit, int m, const double* alpha0,
                      double* params, double* cost, int32_t* nfev, int32_t* njev, int32_t* status) {
    if (nfit <= 0) return MOTES_OK;
    size_t per_warp = fit_scratch_doubles(m) * sizeof(double);
    size_t budget = h->smem_optin > 2048 ? h->smem_optin - 1024 : h->smem_optin;
    int warps = (int)(budget / per_warp);
    if (warps < 1) return fail_arg("nspat too large for the shared-memory fit slab");
    if (warps > kFitMaxWarps) warps = kFitMaxWarps;
    size_t smem = per_warp * warps;
    MOTES_CUDA(cudaFuncSetAttribute(moffat_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = (int)(budget / smem);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    int grid = h->sm_count * ctas_per_sm;
    int needed = (nfit + warps - 1) / warps;
    if (grid > needed) grid = needed;
    MOTES_TRY(h->counters.ensure(256));
    MOTES_CUDA(cudaMemsetAsync(h->counters.ptr, 0, 256, h->stream));
    moffat_fit_kernel<<<grid, warps * 32, smem, h->stream>>>(profiles, nfit, m, alpha0, params, cost, nfev,
                                                             njev, status, h->counters.as<unsigned int>());
    MOTES_CUDA(cudaGetLastError());
    h->launches += 1;
    return MOTES_OK;
}

static int launch_extract(motes_handle* h, double* data, const double* errs, int nspat, int ndisp,
                          const double* lim_lo, const double* lim_hi, const double* bin_params, int nbins,
                          int wavelength_start, double* spectra, int32_t* lo_idx, int32_t* hi_idx,
                          uint8_t* crmask) {
    size_t smem = (size_t)2 * nspat * kExtStride * sizeof(double);
    if (smem > h->smem_optin) return fail_arg("nspat too large for the extraction tile");
    MOTES_CUDA(cudaFuncSetAttribute(optimal_extract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (ndisp + kExtCols - 1) / kExtCols;
    optimal_extract_kernel<<<grid, kExtThreads, smem, h->stream>>>(data, errs, nspat, ndisp, lim_lo, lim_hi,
                                                                   bin_params, nbins, wavelength_start, spectra,
                                                                   lo_idx, hi_idx, crmask);
    MOTES_CUDA(cudaGetLastError());
    h->launches += 1;
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
    if (host && count) MOTES_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    return MOTES_OK;
}

}  // namespace motes

using namespace motes;

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char* motes_last_error(void) { return last_error_string().c_str(); }
const char* motes_version(void) { return "motes_b200 0.1 (sm_100a)"; }

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
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return fail_cuda(e, "cudaStreamCreate", __FILE__, __LINE__); }
    *out = h;
    return MOTES_OK;
}

int motes_destroy(motes_handle* h) {
    if (!h) return MOTES_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto& s : h->stage) s.release();
    for (auto& s : h->work) s.release();
    h->counters.release();
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
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    return MOTES_OK;
}

long long motes_launch_count(motes_handle* h) { return h ? h->launches : 0; }

int motes_moffat_fit_batch_dev(motes_handle* h, const double* profiles, int nfit, int nspat,
                               const double* alpha0, double* params, double* cost, int32_t* nfev,
                               int32_t* njev, int32_t* status) {
    if (!h || !profiles || !alpha0 || !params) return fail_arg("motes_moffat_fit_batch: null argument");
    if (nfit < 0 || nspat < 1) return fail_arg("motes_moffat_fit_batch: bad size");
    MOTES_CUDA(cudaSetDevice(h->device));
    return launch_fit(h, profiles, nfit, nspat, alpha0, params, cost, nfev, njev, status);
}

int motes_moffat_fit_batch_host(motes_handle* h, const double* profiles, int nfit, int nspat,
                                const double* alpha0, double* params, double* cost, int32_t* nfev,
                                int32_t* njev, int32_t* status) {
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
    MOTES_TRY(launch_fit(h, d_prof, nfit, nspat, d_a0, d_par, d_cost, d_int, d_int + nfit, d_int + 2 * (size_t)nfit));
    MOTES_TRY(download(h, params, d_par, (size_t)nfit * kParams));
    MOTES_TRY(download(h, cost, d_cost, (size_t)nfit));
    MOTES_TRY(download(h, nfev, d_int, (size_t)nfit));
    MOTES_TRY(download(h, njev, d_int + nfit, (size_t)nfit));
    MOTES_TRY(download(h, status, d_int + 2 * (size_t)nfit, (size_t)nfit));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    return MOTES_OK;
}

int motes_optimal_extract_host(motes_handle* h, double* data, const double* errs, int nspat, int ndisp,
                               const double* lim_lo, const double* lim_hi, const double* bin_params,
                               int nbins, int wavelength_start, double* spectra, int32_t* lo_idx,
                               int32_t* hi_idx, uint8_t* crmask) {
    if (!h || !data || !errs || !lim_lo || !lim_hi || !bin_params || !spectra)
        return fail_arg("motes_optimal_extract: null argument");
    if (nspat < 1 || ndisp < 1 || nbins < 0) return fail_arg("motes_optimal_extract: bad size");
    MOTES_CUDA(cudaSetDevice(h->device));
    size_t npix = (size_t)nspat * ndisp;
    double *d_data, *d_errs, *d_lo, *d_hi, *d_bins, *d_spec;
    int32_t* d_idx;
    uint8_t* d_mask;
    MOTES_TRY(upload(h, h->stage[0], data, npix, &d_data));
    MOTES_TRY(upload(h, h->stage[1], errs, npix, &d_errs));
    MOTES_TRY(upload(h, h->stage[2], lim_lo, (size_t)ndisp, &d_lo));
    MOTES_TRY(upload(h, h->stage[3], lim_hi, (size_t)ndisp, &d_hi));
    MOTES_TRY(upload(h, h->stage[4], bin_params, (size_t)nbins * 8, &d_bins));
    MOTES_TRY(upload<double>(h, h->stage[5], nullptr, (size_t)4 * ndisp, &d_spec));
    MOTES_TRY(upload<int32_t>(h, h->stage[6], nullptr, (size_t)2 * ndisp, &d_idx));
    MOTES_TRY(upload<uint8_t>(h, h->stage[7], nullptr, crmask ? npix : 0, &d_mask));
    MOTES_TRY(launch_extract(h, d_data, d_errs, nspat, ndisp, d_lo, d_hi, d_bins, nbins, wavelength_start, d_spec,
                             d_idx, d_idx + ndisp, crmask ? d_mask : nullptr));
    MOTES_TRY(download(h, data, d_data, npix));
    MOTES_TRY(download(h, spectra, d_spec, (size_t)4 * ndisp));
    MOTES_TRY(download(h, lo_idx, d_idx, (size_t)ndisp));
    MOTES_TRY(download(h, hi_idx, d_idx + ndisp, (size_t)ndisp));
    MOTES_TRY(download(h, crmask, d_mask, crmask ? npix : 0));
    MOTES_CUDA(cudaStreamSynchronize(h->stream));
    return MOTES_OK;
}

}  // extern "C"
