/* motes_b200 - C ABI of the B200-native MOTES extraction hot path.
 *
 * The reference (tseccull/motes) is pure Python and has no FFI of its own; each entry
 * point below replaces one reference function (or the numpy/scipy call inside it) and
 * is what a binding for that function would call.  Citations are to
 * /root/reference/src/motes/<file>:<lines>.  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - every function returns 0 on success or a negative MOTES_E_* code; nothing throws;
 *   - all arrays are caller-owned, C-contiguous, little-endian; frames are
 *     (nspat, ndisp) with the dispersion axis fastest, exactly as motesio.data_harvest
 *     builds them (io/motesio.py:132-136);
 *   - "_host" entry points take host pointers and do their own H2D/D2H copies on the
 *     handle's stream and synchronise before returning; "_dev" entry points take device
 *     pointers, enqueue on the handle's stream and do NOT synchronise;
 *   - there is no CPU fallback: without a usable CUDA device every call fails with
 *     MOTES_E_CUDA.
 */
#ifndef MOTES_B200_H
#define MOTES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOTES_OK 0
#define MOTES_E_CUDA (-1)      /* CUDA runtime error; see motes_last_error() */
#define MOTES_E_ARG (-2)       /* bad argument (null pointer, size out of range) */
#define MOTES_E_NOMEM (-3)     /* device allocation failed */
#define MOTES_E_NOSKY (-4)     /* a column has no sky pixel: reference raises RuntimeError (sky.py:323-335) */
#define MOTES_E_FIT (-5)       /* a Moffat fit could not start: reference raises ValueError (scipy x0/bounds/residual checks) */

/* per-fit status written by the fit kernels: scipy's termination codes 0..4, or */
#define MOTES_FIT_X0_INFEASIBLE (-1)
#define MOTES_FIT_NONFINITE (-2)
#define MOTES_FIT_BAD_BOUNDS (-3)

typedef struct motes_handle motes_handle;

/* Library / device management ------------------------------------------------------ */
int motes_create(int device, motes_handle** out);
int motes_destroy(motes_handle* h);
const char* motes_last_error(void);
const char* motes_version(void);
/* cudaStream_t the handle enqueues on (as void*), and a way to adopt an external one. */
void* motes_stream(motes_handle* h);
int motes_set_stream(motes_handle* h, void* cuda_stream);
int motes_synchronize(motes_handle* h);
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
long long motes_launch_count(motes_handle* h);

/* a3: batched bounded-TRF Moffat fit.
 * Replaces tracing.moffat_least_squares (tracing.py:550-607) including the scipy
 * least_squares(method="trf", ftol=1e-12) call, for nfit profiles of nspat points on the
 * spatial axis 0..nspat-1 (io/motesio.py:94).  alpha0[i] = seeing / pixel_resolution.
 * params: nfit x 6 (A, c, alpha, beta, B, m).  cost/nfev/njev/status may be NULL. */
int motes_moffat_fit_batch_host(motes_handle* h, const double* profiles, int nfit, int nspat,
                                const double* alpha0, double* params, double* cost,
                                int32_t* nfev, int32_t* njev, int32_t* status);
int motes_moffat_fit_batch_dev(motes_handle* h, const double* profiles, int nfit, int nspat,
                               const double* alpha0, double* params, double* cost,
                               int32_t* nfev, int32_t* njev, int32_t* status);

/* a10: Horne optimal + aperture extraction with the single-pass 5-sigma mask.
 * Replaces extraction.optimal_extraction (extraction.py:39-210) for one frame.
 *   data, errs   (nspat, ndisp) frames; data is scrubbed in place as filter_data does
 *                (extraction.py:31-34), errs is not modified (the reference works on a copy)
 *   lim_lo/hi    ndisp fine extraction limits (tracing.interpolate_extraction_lims output)
 *   bin_params   nbins x 8 rows [A, c, alpha, beta, B, m, first_col+wavstart, last_col+wavstart]
 *   spectra      4 x ndisp: optimal, optimal_err, aperture, aperture_err
 *   lo_idx/hi_idx  (optional) ndisp int32: int(floor(lo)), int(ceil(hi))+1
 *   crmask       (optional) ndisp x nspat uint8: 1 where the pixel was kept by the mask */
int motes_optimal_extract_host(motes_handle* h, double* data, const double* errs, int nspat, int ndisp,
                               const double* lim_lo, const double* lim_hi, const double* bin_params,
                               int nbins, int wavelength_start, double* spectra, int32_t* lo_idx,
                               int32_t* hi_idx, uint8_t* crmask);

#ifdef __cplusplus
}
#endif
#endif /* MOTES_B200_H */
