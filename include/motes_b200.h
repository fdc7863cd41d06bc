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
 *   - sizes: nspat up to ~1000 rows (shared-memory slabs of the fit, sky and extraction kernels), any ndisp (rows
 *     longer than ~28 000 columns take a global-memory median path), 100 * ndisp * 2^ceil(log2 nspat) < 2^32 for the
 *     parallel bootstrap (beyond: one CTA per frame);
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
#define MOTES_E_INTERNAL (-6)  /* a device-side invariant failed (e.g. the planned generator stream was too short) */

/* per-fit status written by the fit kernels: scipy's termination codes 0..4, or */
#define MOTES_FIT_X0_INFEASIBLE (-1)
#define MOTES_FIT_NONFINITE (-2)
#define MOTES_FIT_BAD_BOUNDS (-3)

typedef struct motes_handle motes_handle;

/* Library / device management ------------------------------------------------------ */
/* motes_create reads these environment switches (A/B timing and cross-checks in tests/; none changes results):
 *   MOTES_B200_BOOTSTRAP=classic  one CTA replays a frame's whole generator stream (the literal form of sky.py:34-38)
 *   MOTES_B200_COLUMN=0|<window>  exact rank-slot kernel only / histogram window of the bootstrap column kernel
 *   MOTES_B200_HOIST=0            generate the bootstrap stream where it is consumed instead of at the start of the call
 *   MOTES_B200_WALK=table|stream|serial  position walk: by runs over the accept-count table / streaming / table, one column at a time
 *   MOTES_B200_HOIST_CTAS=<n>     generator CTAs per SM while it runs beside the localisation stages (default 8)
 *   MOTES_B200_STREAM_GB=<x>      device memory the generator streams may take (default: half of what is free)
 *   MOTES_B200_FIT=warp           round-1 one-warp-per-fit kernel
 *   MOTES_B200_CHUNKS=<n>         sub-batches per call on two streams (default 1) */
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

/* FP64 roof of the handle's device: a DFMA micro-benchmark (8 independent chains per thread, every SM full),
 * best of four timed launches, in TFLOP/s counting an FMA as two.  The denominator of bench.py's roofline.fp64
 * (the reference has no counterpart). */
int motes_fp64_peak(motes_handle* h, double* tflops);

/* Optional per-stage timing with CUDA events on the handle's stream.  Stages: see
 * motes_profile_stage_name().  ms / launches: motes_profile_stage_count() entries, accumulated since
 * the last reset; motes_profile_read synchronises the stream. */
int motes_profile_enable(motes_handle* h, int on);
int motes_profile_stage_count(void);
const char* motes_profile_stage_name(int stage);
int motes_profile_read(motes_handle* h, double* ms, long long* launches, int reset);

/* ---------------------------------------------------------------------------------------
 * Whole path for a batch of frames (core.motes stage order, core.py:80-275).
 * ------------------------------------------------------------------------------------- */
typedef struct motes_params {
    double minimum_column_limit;   /* -m   (xmotes.py), tracing.py:165 */
    double extraction_snr_limit;   /* -xs  tracing.py:162 */
    double sky_snr_limit;          /* -ks  tracing.py:160 */
    double sky_fwhm_multiplier;    /* -kf  used for BOTH the sky and the extraction limits (core.py:208) */
    int32_t subtract_sky;          /* -k   */
    int32_t sky_order;             /* -ko  0 = bootstrap median (sky.py:17-43), 1..5 = bootstrap polynomial (sky.py:177-254) */
    int32_t spatial_floor;         /* axes_dict["data_spatial_floor"] */
    int32_t wavelength_start;      /* axes_dict["wavelength_start"] */
} motes_params;

/* Output block.  Every pointer may be NULL (that output is then not produced / copied).
 * F = nframes, cap = bin capacity passed to the call. */
typedef struct motes_outputs {
    double* spectra;          /* [F][4][ndisp] optimal, optimal_err, aperture, aperture_err (extraction.py:208) */
    double* ext_limits;       /* [F][2][ndisp] final extraction limits (core.py:248-252) */
    double* sky_limits;       /* [F][2][ndisp] sky limits after the in-place clamp (sky.py:303-307) */
    double* ext_params;       /* [F][cap][8]  per-bin [A,c,alpha,beta,B,m,first+wavstart,last+wavstart] (tracing.py:544-545) */
    double* sky_params;       /* [F][cap][8] */
    int32_t* ext_bins;        /* [F][cap][2]  bin first/last column (tracing.py:207,246) */
    int32_t* sky_bins;        /* [F][cap][2] */
    double* ext_bin_snr;      /* [F][cap] */
    double* sky_bin_snr;      /* [F][cap] */
    int32_t* n_ext_bins;      /* [F] */
    int32_t* n_sky_bins;      /* [F] */
    double* median_params;    /* [F][6]  fit of the full-frame median profile (tracing.py:403-456) */
    double* scale;            /* [F]  data_scaling_factor */
    double* seeing_px;        /* [F]  header_parameters["seeing"] after tracing.py:449 */
    double* sky_model;        /* sky_order 0: [F][ndisp] per-column sky level; sky_order > 0: [F][nspat][ndisp]; 0 where not modelled */
    double* sky_uncertainty;  /* same shape as sky_model */
    int32_t* sky_flag;        /* [F][ndisp] 0 modelled, 1 skipped as constant (sky.py:338-339), 2 no sky pixels */
    int32_t* frame_status;    /* [F] 0, MOTES_E_NOSKY or MOTES_E_FIT: what the reference would have raised for that frame */
} motes_outputs;

/* Upper bound on the number of bins get_bins can produce; use as `cap`. */
int motes_bin_capacity(int ndisp, double minimum_column_limit);

/* np.random.seed(seed[i]) for every frame: fills states[i*625 .. i*625+624] with the 624 MT19937
 * state words followed by the position (624).  Pure host code. */
int motes_seed_states(const uint32_t* seeds, int nframes, uint32_t* states);

/* data / errs are modified in place on the device exactly as the reference modifies frame_dict
 * (sky subtraction sky.py:375-397, scrub extraction.py:31-34); with copy_back != 0 the host
 * version copies them back.  qual: uint8, 1 = good.  seeing / pixres: per frame.
 * rng_states: [F][625] in/out, the numpy legacy RandomState of each frame (see motes_seed_states). */
int motes_run_frames_host(motes_handle* h, double* data, double* errs, const uint8_t* qual, int nframes,
                          int nspat, int ndisp, int cap, const motes_params* params, const double* seeing,
                          const double* pixres, uint32_t* rng_states, int copy_back, const motes_outputs* out);
int motes_run_frames_dev(motes_handle* h, double* data, double* errs, const uint8_t* qual, int nframes,
                         int nspat, int ndisp, int cap, const motes_params* params, const double* seeing,
                         const double* pixres, uint32_t* rng_states, const motes_outputs* out);
/* motes_run_frames_host without the final wait: uploads, kernels and downloads are queued on the handle's
 * streams and the call returns.  Every host buffer (inputs, rng_states, outputs; pinned memory for real
 * overlap) must stay untouched until motes_synchronize(h) returns.  A caller with a stream of batches
 * alternates between two handles, so that the upload of batch k+1 runs under the kernels of batch k
 * (the reference has no counterpart: core.motes processes one file at a time, core.py:28-275). */
int motes_run_frames_host_async(motes_handle* h, double* data, double* errs, const uint8_t* qual, int nframes,
                                int nspat, int ndisp, int cap, const motes_params* params, const double* seeing,
                                const double* pixres, uint32_t* rng_states, int copy_back, const motes_outputs* out);
/* The same path for frames held as float32 planes - what the GMOS harvester hands over: big-endian float32 straight
 * from the FITS file (io/gmosio.py:40-52; big_endian != 0) or native float32.  The planes cross PCIe at 4 bytes per
 * pixel and are widened to float64 on the device (exact), so the results are those of motes_run_frames_host on the
 * same values converted on the host; nothing is written back into the caller's frames.  wait != 0: return when the
 * results are in the output block (motes_run_frames_host semantics); wait == 0: motes_run_frames_host_async semantics. */
int motes_run_frames_host_f32(motes_handle* h, const float* data, const float* errs, const uint8_t* qual, int nframes,
                              int nspat, int ndisp, int cap, const motes_params* params, const double* seeing,
                              const double* pixres, uint32_t* rng_states, int big_endian, int wait, const motes_outputs* out);
/* float64 {0,1} quality frame -> uint8 on the device (harvesters produce float64, io/gmosio.py:50-52) */
int motes_qual_pack_dev(motes_handle* h, const double* qual, size_t count, uint8_t* out);

/* ---------------------------------------------------------------------------------------
 * Stage entry points (one frame, host pointers): what each reference function would bind.
 * ------------------------------------------------------------------------------------- */

/* a1: np.nanmedian(frame_dict["data"], axis=1) (core.py:80) -> out[nspat] */
int motes_row_median_host(motes_handle* h, const double* data, int nspat, int ndisp, double* out);

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

/* a5: tracing.get_bins (tracing.py:114-254).  bins: [cap][2], snr: [cap]; returns the count in *nbins.
 * gap (optional, ndisp): 1 where np.median(data[:, c]) == 1 (chip-gap columns, tracing.py:506). */
int motes_get_bins_host(motes_handle* h, const double* data, const double* errs, const uint8_t* qual,
                        int nspat, int ndisp, int row_lo, int row_hi, double snr_limit,
                        double minimum_column_limit, int cap, int32_t* bins, double* snr, int32_t* nbins,
                        uint8_t* gap);

/* a6: per-bin median spatial profiles (tracing.py:505-507) -> profiles[nbins][nspat] */
int motes_bin_profiles_host(motes_handle* h, const double* data, int nspat, int ndisp, const int32_t* bins,
                            int nbins, double* profiles);

/* a8: tracing.interpolate_extraction_lims (tracing.py:315-375).  table: [4][nbins]. */
int motes_interp_limits_host(motes_handle* h, const double* table, int nbins, int ndisp, double* lim_lo,
                             double* lim_hi);

/* a9: sky.subtract_sky (sky.py:257-399).  data / errs / lim_lo / lim_hi are updated in place as the
 * reference does; rng_state[625] is numpy's legacy RandomState (no cached Gaussian), advanced by exactly
 * the draws the reference consumes.  sky_order 0 = bootstrap median (sky.py:17-43): sky / sky_err hold
 * ndisp values; sky_order 1..5 = bootstrap polynomial (sky.py:177-254): they hold (nspat, ndisp) frames.
 * Returns MOTES_E_NOSKY where the reference raises RuntimeError. */
int motes_sky_subtract_host(motes_handle* h, double* data, double* errs, int nspat, int ndisp, double* lim_lo,
                            double* lim_hi, int spatial_floor, int sky_order, uint32_t* rng_state,
                            double* sky, double* sky_err, int32_t* flag, int32_t* n_sky, int32_t* n_good);

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
