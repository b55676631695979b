/*
 * apgpu.h -- C ABI of the B200 (sm_100a) FITS-reduction kernels.
 *
 * This is the drop-in boundary for ONE hot path of DaveStrickland/AstroPhotography
 * (v0.5.1): master-frame combination, science-frame calibration and bad-pixel
 * repair.  The reference is pure Python/numpy and has no FFI of its own; each
 * entry point below replaces the numpy / ccdproc arithmetic at the cited
 * reference location and is what a binding added to the reference would call
 * (ctypes stub: see INTEGRATION.md).
 *
 * Conventions
 *   - every function returns 0 (APGPU_OK) or a non-zero error code;
 *     apgpu_last_error() returns a thread-local message for the last failure
 *     (the Python side raises RuntimeError(msg), the reference's own error
 *     convention, core/ApCalibrate.py:123-125);
 *   - all image pointers are DEVICE pointers to row-major [H][W] images
 *     (numpy shape (NAXIS2, NAXIS1), core/ApCalibrate.py:280-281) owned by the
 *     caller; inputs are never modified, outputs must not alias inputs;
 *   - every call is stream-ordered on `stream` (a cudaStream_t, NULL = default
 *     stream), asynchronous, re-entrant, and allocates nothing persistent;
 *   - there is no CPU fallback anywhere behind this header.
 */
#ifndef APGPU_H
#define APGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APGPU_ABI_VERSION 2

#define APGPU_OK 0
#define APGPU_ERR_ARG 1         /* bad argument (null pointer, bad size, bad enum) */
#define APGPU_ERR_CUDA 2        /* a CUDA runtime call failed */
#define APGPU_ERR_UNSUPPORTED 3 /* valid request outside implemented limits */

typedef void* apgpu_stream_t;   /* cudaStream_t */

int apgpu_abi_version(void);
const char* apgpu_last_error(void);
/* Number of kernel launches issued through this library by the calling
 * process since load (bench.py's "gpu_launches" evidence). */
uint64_t apgpu_launch_count(void);

/* ------------------------------------------------------------------------
 * Frame-stack reducer.  Replaces ccdproc.combine(...) as called at
 * scripts/ap_combine_darks.py:411-420 (settings :394-399) and generalises it
 * to the iterative kappa-sigma clip of astropy.stats.sigma_clip.
 *
 * frames      HOST array of N DEVICE pointers, each an H x W float32 frame
 * row0,nrows  row band to reduce (multi-GPU row-band sharding); outputs have the
 *             frame geometry and only rows [row0,row0+nrows) are written
 * method      APGPU_METHOD_*  (ccdproc method='average' | 'median'; min/max extra)
 * k_lo,k_hi   clip thresholds in units of the deviation (reference: 5, 5)
 * maxiters    0 = no clipping, >0 = at most that many clip passes
 *             (reference / ccdproc: 1), <0 = iterate to convergence
 * cen, dev    APGPU_CEN_* / APGPU_DEV_* (reference: MEDIAN / MAD_STD)
 * out_data    H x W float32 (out_is_f64=0) or float64 (=1) combined image
 * out_nrej    H x W count of samples not used (clipped or NaN/non-finite),
 *             uint8 (nrej_is_u16=0, requires N<=255) or uint16; may be NULL
 * out_uncert  same dtype as out_data, or NULL.  average/min/max:
 *             std(kept)/sqrt(n_kept); median: 1.4826*MAD(kept)/sqrt(n_kept)
 * out_allmasked  uint8, 1 where no sample survived (data = NaN there); may be NULL
 * flags       APGPU_STACK_FORCE_GENERIC: use the exact float64 generic kernel
 *             even where a fast kernel exists; APGPU_STACK_PREFER_REGISTERS /
 *             APGPU_STACK_PREFER_SHARED: override the default choice between the
 *             register-resident and shared-memory-resident fast kernels
 * Limits: 1 <= N <= 1024.
 * One call is a short sequence of launches on `stream`: the kernel(s) of the chosen
 * family, then -- unless the generic kernel did everything -- a scan launch that redoes
 * the few pixels a fast kernel could not finish (non-finite samples, float32 guard-band
 * hits).  Between the two, those pixels hold a marker (a payload NaN in out_data, all
 * ones in out_nrej): outputs are complete when the work queued on `stream` is.
 * ---------------------------------------------------------------------- */
enum { APGPU_METHOD_MEDIAN = 0, APGPU_METHOD_AVERAGE = 1, APGPU_METHOD_MIN = 2, APGPU_METHOD_MAX = 3 };
enum { APGPU_CEN_MEAN = 0, APGPU_CEN_MEDIAN = 1 };
enum { APGPU_DEV_STD = 0, APGPU_DEV_MAD_STD = 1 };
#define APGPU_STACK_FORCE_GENERIC 1
#define APGPU_STACK_PREFER_REGISTERS 2   /* tuning/tests: register-resident kernel where one exists */
#define APGPU_STACK_PREFER_SHARED 4      /* tuning/tests: shared-memory-resident kernel where one exists */
#define APGPU_STACK_USE_TMA 8            /* tuning/tests: CTA-wide TMA bulk-copy staging */
#define APGPU_STACK_DIRECT_LOADS 16      /* tuning/tests: direct global loads, no shared-memory staging */
#define APGPU_STACK_USE_CPASYNC 32       /* tuning/tests: warp-granular cp.async staging pipeline */
#define APGPU_STACK_USE_TENSORMAP 64     /* tuning/tests: warp-granular tensor-map TMA staging (equally spaced frames) */
#define APGPU_STACK_MAX_FRAMES 1024

int apgpu_stack_reduce_f32(const float* const* frames, int N, int64_t H, int64_t W,
                           int64_t row0, int64_t nrows, int method,
                           double k_lo, double k_hi, int maxiters, int cen, int dev,
                           void* out_data, int out_is_f64,
                           void* out_nrej, int nrej_is_u16,
                           void* out_uncert, uint8_t* out_allmasked,
                           int flags, apgpu_stream_t stream);

/* ------------------------------------------------------------------------
 * The same reducer on raw 16-bit frames: the integer -> float32 conversion of
 * _read_fits (core/ApCalibrate.py:303-307; raw frames are BITPIX 16 / BZERO 32768,
 * doc/fits_metadata.md:70-76) is fused into the load phase, so a frame costs
 * 2 bytes per pixel over PCIe and HBM.  Results are those of apgpu_stack_reduce_f32
 * on float32(frames).
 *
 * u16_format  APGPU_U16_NATIVE      host-order uint16 samples (what astropy hands out with uint=True)
 *             APGPU_U16_FITS_BZERO  the data unit of a BITPIX=16, BZERO=32768 FITS image as it is on
 *                                   disk: big-endian int16 s, value = s + 32768
 * PEDESTAL is not applied here: like ccdproc.combine the master keeps the keyword.
 * ---------------------------------------------------------------------- */
enum { APGPU_U16_NATIVE = 0, APGPU_U16_FITS_BZERO = 1 };
int apgpu_stack_reduce_u16(const uint16_t* const* frames, int u16_format, int N, int64_t H, int64_t W,
                           int64_t row0, int64_t nrows, int method,
                           double k_lo, double k_hi, int maxiters, int cen, int dev,
                           void* out_data, int out_is_f64,
                           void* out_nrej, int nrej_is_u16,
                           void* out_uncert, uint8_t* out_allmasked,
                           int flags, apgpu_stream_t stream);

/* Name of the kernel family apgpu_stack_reduce_f32 would pick for these
 * settings ("generic", "meanclip<NB>", "sorted<NB>", ...); for logs and tests. */
const char* apgpu_stack_kernel_name(int N, int method, double k_lo, double k_hi,
                                    int maxiters, int cen, int dev,
                                    int want_uncert, int out_is_f64, int flags);

/* How the last apgpu_stack_reduce_f32 call of the calling thread fed its kernel
 * (tests and the benchmark's bookkeeping):
 *   -1 generic kernel
 *    0 direct global loads (any frame pointers)
 *    1 CTA-wide copies: bulk copies (meanclip, on request) / one tensor-map TMA box per
 *      256-pixel tile (sorted kernels, equally spaced frames)
 *    2 warp-granular cp.async pipeline
 *    3 warp-granular tensor-map TMA pipeline (equally spaced frames; meanclip default)
 *    4 lane-split cp.async (512 < N <= 1024)
 *    5 warp-cooperative 128B-swizzled tensor-map TMA (100 < N <= 512)
 * A requested staging falls back to 0 when its layout / alignment preconditions do not hold. */
int apgpu_stack_last_staging(void);

/* ------------------------------------------------------------------------
 * Flat normalisation.  Replaces ApCalibrate._generate_flat,
 * core/ApCalibrate.py:178-190:  norm = np.nanmean(flat); out = flat / norm.
 *
 * apgpu_flat_norm_f32 reproduces numpy's float32 pairwise summation tree
 * bit-for-bit (NaN -> 0, count of non-NaN) and writes
 * norm = float32(float64(sum)/count) to *norm_out (device).  `workspace` is
 * device scratch of at least apgpu_flat_norm_workspace_bytes(npix) bytes.
 * ---------------------------------------------------------------------- */
size_t apgpu_flat_norm_workspace_bytes(int64_t npix);
int apgpu_flat_norm_f32(const float* flat, int64_t npix, void* workspace,
                        size_t workspace_bytes, float* norm_out, apgpu_stream_t stream);
/* normflat[i] = flat[i] / *norm  (float32 division, IEEE round-to-nearest). */
int apgpu_flat_divide_f32(const float* flat, const float* norm, float* normflat,
                          int64_t npix, apgpu_stream_t stream);

/* ------------------------------------------------------------------------
 * Fused science-frame calibration.  Replaces the numpy arithmetic of
 * ApCalibrate.calibrate, core/ApCalibrate.py:439-474, in one pass:
 *     t   = raw - bias
 *     d   = dark_still_biased ? dark - bias : dark
 *     out = t - float32(exp_ratio) * d
 *     out = (normflat != 0) ? out / normflat : out          (normflat may be NULL)
 * every operation separately rounded in float32 (no FMA contraction).
 * The _u16 variant fuses _read_fits' conversion, core/ApCalibrate.py:303-326:
 * raw = float32(raw_u16) (+ pedestal when has_pedestal).
 * ---------------------------------------------------------------------- */
int apgpu_calibrate_f32(const float* raw, const float* bias, const float* dark,
                        const float* normflat, float exp_ratio, int dark_still_biased,
                        float* out, int64_t npix, apgpu_stream_t stream);
int apgpu_calibrate_u16(const uint16_t* raw, float pedestal, int has_pedestal,
                        const float* bias, const float* dark, const float* normflat,
                        float exp_ratio, int dark_still_biased,
                        float* out, int64_t npix, apgpu_stream_t stream);

/* ------------------------------------------------------------------------
 * Bad-pixel repair.  Replaces the per-bad-pixel Python loop of
 * ApFixBadPixels.fix_bad_pixels, core/ApFixBadPixels.py:380-419: every pixel
 * whose mask is non-zero is replaced by np.median of the good pixels of the
 * ORIGINAL data in the (2*deltapix+1)^2 window clipped to the image, when at
 * least min_valid (reference: 4) of them exist; otherwise it is left unchanged.
 *
 * data, mask  `band_rows` rows starting at image row `band_row0` of an
 *             H_image x W image (single GPU: band_row0=0, band_rows=H_image);
 *             the band must contain the deltapix halo rows that exist
 * row0,nrows  image rows to produce; out has `nrows` rows (row0 first)
 * mask_dtype  APGPU_MASK_* element type of `mask`; only (!= 0) is used
 * counts      device int64[2]: += {number of bad pixels, number repaired}
 *             over the produced rows (caller zeroes it)
 * Limits: 1 <= deltapix <= 7.
 * ---------------------------------------------------------------------- */
enum { APGPU_MASK_U8 = 0, APGPU_MASK_I16 = 1, APGPU_MASK_I32 = 2, APGPU_MASK_F32 = 3, APGPU_MASK_F64 = 4 };

int apgpu_fix_badpix_f32(const float* data, const void* mask, int mask_dtype,
                         int64_t H_image, int64_t W, int64_t band_row0, int64_t band_rows,
                         int64_t row0, int64_t nrows, int deltapix, int min_valid,
                         float* out, int64_t* counts, apgpu_stream_t stream);

/* ------------------------------------------------------------------------
 * Fused calibration + bad-pixel repair of one science frame (the batch driver's per-frame launch):
 * ApCalibrate.calibrate, core/ApCalibrate.py:439-479 -- the arithmetic of apgpu_calibrate_* followed by
 * apgpu_fix_badpix_f32 on its result, bit for bit, without the intermediate image (a bad pixel's donors are
 * the calibrated values of its neighbours, recomputed on the fly).
 *
 * raw, raw_kind   H x W frame: float32, host-order uint16, or the data unit of a BITPIX=16 / BZERO=32768
 *                 FITS image as stored on disk (big-endian int16 + 32768)
 * mask            uint8, non-zero = bad; NULL = calibrate only (counts may then be NULL too)
 * out_big_endian  non-zero: write big-endian float32, i.e. the data unit of a BITPIX=-32 FITS image
 * counts          device int64[2]: += {bad pixels, repaired} (caller zeroes it)
 * ---------------------------------------------------------------------- */
enum { APGPU_RAW_F32 = 0, APGPU_RAW_U16 = 1, APGPU_RAW_U16_FITS = 2 };
int apgpu_calibrate_repair(const void* raw, int raw_kind, float pedestal, int has_pedestal,
                           const float* bias, const float* dark, const float* normflat,
                           float exp_ratio, int dark_still_biased,
                           const uint8_t* mask, int64_t H, int64_t W, int deltapix, int min_valid,
                           float* out, int out_big_endian, int64_t* counts, apgpu_stream_t stream);

/* ------------------------------------------------------------------------
 * Image arithmetic.  Replaces np.add / np.subtract / np.multiply / np.divide(data1, data2, out=result) of
 * ApImArith.process_files, core/ApImArith.py:321-333, for a float32 image `a`:
 *   b_kind 0  scalar (rounded to float32 first, as numpy treats a Python float), float32 arithmetic
 *   b_kind 1  `b` is a float32 image, float32 arithmetic
 *   b_kind 2  `b` is a float64 image: float64 arithmetic, result cast to float32
 * Division by zero gives inf / NaN as in numpy.
 * ---------------------------------------------------------------------- */
enum { APGPU_OP_ADD = 0, APGPU_OP_SUB = 1, APGPU_OP_MUL = 2, APGPU_OP_DIV = 3 };
int apgpu_imarith_f32(const float* a, const void* b, int b_kind, double scalar, int op,
                      float* out, int64_t npix, apgpu_stream_t stream);

/* ------------------------------------------------------------------------
 * Whole-image sigma-clipped statistics and threshold mask: the arithmetic of the
 * mask producer ApFindBadPixels._generate_sigmaclip_mask,
 * core/ApFindBadPixels.py:191-209.
 *
 * apgpu_sigma_clipped_stats_f32 restates astropy.stats.sigma_clipped_stats(data,
 * sigma) with its defaults (median centre, population std, maxiters clip rounds)
 * over the finite pixels and writes {mean, median, std, count} of the survivors
 * to out4 (device, 4 doubles).  The median is exact (radix select).
 * apgpu_threshold_mask_f32 writes mask = (data < lo) | (data > hi) as uint8
 * (comparison in float64, as numpy promotes) and adds the number of set pixels
 * to *nbad (device int64, caller zeroes it).
 * ---------------------------------------------------------------------- */
size_t apgpu_image_stats_workspace_bytes(int64_t npix);
int apgpu_sigma_clipped_stats_f32(const float* data, int64_t npix, double sigma, int maxiters,
                                  void* workspace, size_t workspace_bytes, double* out4,
                                  apgpu_stream_t stream);
int apgpu_threshold_mask_f32(const float* data, int64_t npix, double lo, double hi,
                             uint8_t* mask, int64_t* nbad, apgpu_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* APGPU_H */
