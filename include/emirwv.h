/*
 * emirwv — C ABI of the B200-native rectification + wavelength-calibration path.
 *
 * This is the drop-in boundary for ONE hot path of guaix-ucm/pyemir:
 *   emirdrp.processing.wavecal.apply_rectwv_coeff
 *   (reference: src/emirdrp/processing/wavecal/apply_rectwv_coeff.py:35-287).
 * The reference is pure Python; the arithmetic it delegates to numina
 * (apply_integer_offsets :74-78, Slitlet2D.rectify -> rectify2d slitlet2d.py:421-423,
 * resample_image2d_flux :152-159, find_pix_borders :190) is what this library
 * replaces.  A Python caller binds it with ctypes (pyemir_b200/_cabi.py); the
 * reference-side stub is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C: pointers, sizes, POD structs; no C++ or torch types cross the boundary;
 *  - every function returns 0 on success or a negative EMIRWV_E_* code;
 *    emirwv_last_error() gives the message of the calling thread's last failure;
 *  - "d_" arguments are device pointers on the CUDA device that was current when the
 *    plan was created, "h_" arguments are host pointers; the caller owns all of them;
 *  - emirwv_run is asynchronous on the given stream and may be called concurrently
 *    on different streams (a plan is immutable after creation);
 *  - there is no CPU fallback: without a CUDA device every compute entry point
 *    fails with EMIRWV_E_CUDA.
 */
#ifndef EMIRWV_H
#define EMIRWV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMIRWV_VERSION 200          /* 0.2.0 */

#define EMIRWV_MAX_SLITLETS 55      /* EMIR_NBARS, core/__init__.py:39 */
#define EMIRWV_MAX_COEF 21          /* 2-D polynomial of order 5 */
#define EMIRWV_MAX_WPOLY 16         /* wavelength polynomial coefficients */
#define EMIRWV_ROWS_SCANNED 39      /* 38 pasted rows + 1 (apply_rectwv_coeff.py:189) */

/* error codes; the Python shim maps them to the exceptions of the reference */
#define EMIRWV_OK 0
#define EMIRWV_E_VALUE (-1)         /* invalid argument            -> ValueError  */
#define EMIRWV_E_INDEX (-2)         /* row outside the bounding box -> IndexError */
#define EMIRWV_E_CUDA (-3)          /* CUDA runtime failure         -> RuntimeError */
#define EMIRWV_E_NOMEM (-4)         /* allocation failure           -> MemoryError */
#define EMIRWV_E_LIMIT (-5)         /* geometry exceeds kernel limits -> RuntimeError */

/* element type of the frames, of the arithmetic and of the output */
#define EMIRWV_F32 0                /* float32 in, float32 arithmetic, float32 out */
#define EMIRWV_F64 1                /* float64 in, float64 arithmetic, float64 out */

/* resampling: the reference's args_resampling (slitlet2d.py:388-389) */
#define EMIRWV_NEAREST 1            /* nearest neighbour                          */
#define EMIRWV_AREA 2               /* flux preserving (exact pixel-overlap area) */
#define EMIRWV_BILINEAR 3           /* extension, not in the reference: bilinear  */

/* One slitlet of RectWaveCoeff.contents (slitlet2d.py:149-215). */
typedef struct emirwv_slitlet_desc {
    int32_t valid;                  /* 0: listed in missing_slitlets (rows stay 0) */
    int32_t bb_nc1_orig;            /* bounding box, 1-based inclusive             */
    int32_t bb_nc2_orig;
    int32_t bb_ns1_orig;
    int32_t bb_ns2_orig;
    int32_t min_row_rectified;      /* 0-based row of the rectified slitlet        */
    int32_t max_row_rectified;
    int32_t ncoef;                  /* len(ttd_aij) = (order+1)(order+2)/2         */
    int32_t nwpoly;                 /* len(wpoly_coeff)                            */
    int32_t reserved;
    double aij[EMIRWV_MAX_COEF];    /* ttd_aij (tti_aij for the inverse map)       */
    double bij[EMIRWV_MAX_COEF];    /* ttd_bij                                     */
    double wpoly[EMIRWV_MAX_WPOLY]; /* wavelength vs 1-based pixel, ascending powers */
} emirwv_slitlet_desc;

/* Everything apply_rectwv_coeff reads from RectWaveCoeff + set_wv_parameters. */
typedef struct emirwv_plan_desc {
    int32_t naxis1;                 /* frame columns (EMIR_NAXIS1 = 2048)          */
    int32_t naxis2;                 /* frame rows    (EMIR_NAXIS2 = 2048)          */
    int32_t nslitlets;              /* EMIR_NBARS = 55                             */
    int32_t rows_per_slitlet;       /* EMIR_NPIXPERSLIT_RECTIFIED = 38             */
    int32_t offx;                   /* global_integer_offset_x_pix                 */
    int32_t offy;                   /* global_integer_offset_y_pix                 */
    int32_t resampling;             /* EMIRWV_NEAREST / _AREA / _BILINEAR          */
    int32_t dtype;                  /* EMIRWV_F32 / EMIRWV_F64                     */
    int32_t max_order;              /* numina's cap is 4; 5 is accepted on request */
    int32_t naxis1_out;             /* naxis1_enlarged                             */
    double crpix1;                  /* crpix1_enlarged (1.0)                       */
    double crval1;                  /* crval1_enlarged                             */
    double cdelt1;                  /* cdelt1_enlarged                             */
    int32_t tile_k;                 /* tuning, 0 = default (248): output columns per tile,
                                       a multiple of 4; narrowed automatically when the
                                       distortion needs a larger source window           */
    int32_t stages;                 /* tuning, 0 = default: rectified rows whose table records and
                                       source rows are in flight per CTA (2 ... 4; default 4 for
                                       float32, 2 for float64 plans) */
    emirwv_slitlet_desc slitlet[EMIRWV_MAX_SLITLETS];
} emirwv_plan_desc;

typedef struct emirwv_plan emirwv_plan;   /* opaque */

typedef struct emirwv_plan_info {
    int32_t dtype;
    int32_t resampling;
    int32_t footprint;              /* K: every output pixel reads K x K source pixels */
    int32_t naxis1_out;
    int32_t naxis2_out;             /* nslitlets * rows_per_slitlet                */
    int32_t ntiles;                 /* work items per frame group                  */
    int32_t ntiles_empty;           /* tiles outside the wavelength coverage       */
    int32_t tile_k;
    int32_t stages;                 /* rows in flight per CTA (copy pipeline depth) */
    int32_t window_rows;            /* staged source window (max over tiles)       */
    int32_t window_cols;
    int32_t max_span;               /* max whole source pixels inside one output pixel */
    int32_t frames_per_cta;         /* G chosen for the next launch                */
    int32_t threads;                /* threads per CTA of that launch              */
    int32_t smem_bytes;             /* dynamic shared memory of that launch        */
    int32_t ctas_per_sm;            /* resident CTAs per SM of that launch (occupancy API) */
    int64_t table_bytes;            /* device memory held by the plan              */
    int64_t in_frame_bytes;
    int64_t out_frame_bytes;
} emirwv_plan_info;

int emirwv_version(void);
const char *emirwv_last_error(void);

/* Number of CUDA devices visible, or a negative error code. */
int emirwv_device_count(void);

/* Select the CUDA device used by subsequent emirwv_plan_create calls of this thread
 * (one process per GPU: pass LOCAL_RANK). */
int emirwv_set_device(int device);

/* The calling thread's current CUDA device, or a negative error code. */
int emirwv_get_device(void);

/*
 * Build the frame-independent tables of a RectWaveCoeff on the current device.
 * Replaces the per-frame work of Slitlet2D.__init__ (slitlet2d.py:145-218), fmap /
 * the corner mapping and polygon clipping inside rectify2d, and the wavelength
 * border bookkeeping inside resample_image2d_flux.
 */
int emirwv_plan_create(const emirwv_plan_desc *desc, emirwv_plan **plan);
void emirwv_plan_destroy(emirwv_plan *plan);
int emirwv_plan_get_info(const emirwv_plan *plan, emirwv_plan_info *info);

/* Launch tuning for subsequent runs (0 keeps the default): frames per CTA (1, 2 or 4). */
int emirwv_plan_set_tuning(emirwv_plan *plan, int frames_per_cta);

/*
 * Rectify + wavelength-calibrate nframes device-resident frames.
 *   d_in      [nframes][naxis2][naxis1]            plan dtype
 *   d_out     [nframes][naxis2_out][naxis1_out]    plan dtype (every element written)
 *   d_jminmax [nframes][nslitlets][2] int32: 0-based first / last channel with a
 *             non-zero value over the 39 scanned rows of each slitlet, or
 *             (INT32_MAX, -1) when there is none; may be NULL.
 * Replaces the slitlet loop of apply_rectwv_coeff.py:136-204.
 * One call enqueues a 4-byte memset (the item counter of this launch, taken from a pool of
 * 1024 so that launches on different streams never share one) and three kernels.
 */
int emirwv_run(const emirwv_plan *plan, const void *d_in, void *d_out, int nframes,
               int32_t *d_jminmax, void *cuda_stream);

/*
 * emirwv_run with options.
 *   EMIRWV_RUN_OUT_ZEROED: the caller guarantees that d_out already holds zeros wherever the
 *   plan has no wavelength coverage (cudaMemset once, or a buffer filled by an earlier call
 *   with the same plan): the zero-fill kernel is skipped, only the hot kernel runs.  The
 *   result in d_out is the same as without the flag.
 */
#define EMIRWV_RUN_OUT_ZEROED 1u
int emirwv_run_ex(const emirwv_plan *plan, const void *d_in, void *d_out, int nframes,
                  int32_t *d_jminmax, void *cuda_stream, uint32_t flags);

/*
 * Same from host buffers: host->device copies, kernels and device->host copies are
 * pipelined over internal streams; returns when h_out / h_jminmax are complete.
 * Pinned buffers (emirwv_host_alloc, cudaHostAlloc, torch pin_memory) copy
 * asynchronously; pageable buffers work but serialise.
 */
int emirwv_run_host(emirwv_plan *plan, const void *h_in, void *h_out, int nframes,
                    int32_t *h_jminmax);

/*
 * emirwv_run_host with options.
 *   EMIRWV_HOST_OUT_ZEROED: the caller guarantees that h_out already holds zeros wherever the
 *   plan has no wavelength coverage (a zero-initialised buffer, or one filled by an earlier
 *   call with the same plan -- the usual loop over exposures).  Only the covered columns of
 *   every slitlet (emirwv_plan_get_coverage) are copied back: about 62 % of the product for
 *   the J/H/K grisms -- by one kernel per pipeline chunk that writes them straight into a pinned
 *   h_out, by one strided copy per slitlet into a pageable one.  The result in h_out is the
 *   same as without the flag.
 */
#define EMIRWV_HOST_OUT_ZEROED 1u
int emirwv_run_host_ex(emirwv_plan *plan, const void *h_in, void *h_out, int nframes,
                       int32_t *h_jminmax, uint32_t flags);

/*
 * The recipes run `flow(newimg)` on every raw frame before the path (recipes/spec/abba.py:371,
 * core/recipe.py:25-61: bad-pixel mask, bias, dark, flat).  A plan can hold the calibration
 * images of that flow (copied to its device; NULL skips a stage, all NULL clears them):
 *   h_bpm [naxis2][naxis1] uint8 (non-zero = bad); h_bias, h_dark, h_flat [naxis2][naxis1] of
 *   cal_dtype (EMIRWV_F32 / EMIRWV_F64).
 * emirwv_run_host_raw is emirwv_run_host_ex for RAW frames (raw_dtype EMIRWV_U16 -- detector
 * counts, half the bytes over PCIe --, EMIRWV_F32 or EMIRWV_F64): every pipeline chunk goes
 * host -> device, through emirwv_prereduce_ex with the plan's calibration and through the hot
 * kernel.  The plan must be float32 (the flow rounds to float32).
 */
int emirwv_plan_set_calibration(emirwv_plan *plan, const uint8_t *h_bpm, const void *h_bias,
                                const void *h_dark, const void *h_flat, int cal_dtype);
int emirwv_run_host_raw(emirwv_plan *plan, const void *h_raw, int raw_dtype, void *h_out,
                        int nframes, int32_t *h_jminmax, uint32_t flags);

/*
 * The ABBA product from host frames in one call (recipes/spec/abba.py:351-381 rectification of
 * every frame, :545-558 per-frame vertical offsets, :579-606 combination of the A and of the B
 * images and their float32 difference): only the combined image travels back.
 *   h_in      [nframes][naxis2][naxis1]: reduced frames of the plan dtype, or raw frames of
 *             in_dtype with EMIRWV_HOST_RAW (the plan's calibration flow runs first)
 *   h_labels  [nframes] int8: +1 = A, -1 = B, 0 = not used
 *   h_offset  [nframes] the recipe's list_offsets (a frame is shifted by
 *             shift_image2d(data, yoffset=-offset).astype("float32") when its offset is not 0),
 *             or NULL
 *   method    EMIRWV_MEAN / EMIRWV_MEDIAN / EMIRWV_SIGMACLIP (low, high: clipping limits);
 *             median and sigma clipping keep the products on the device: at most 64 images per
 *             nodding position
 *   h_out     [naxis2_out][naxis1_out] float32
 * Copies and kernels are pipelined over internal streams; the float64 sums of the mean run in
 * frame order.  Returns when h_out is complete.
 */
#define EMIRWV_MEAN 0
#define EMIRWV_MEDIAN 1
#define EMIRWV_SIGMACLIP 2
#define EMIRWV_HOST_RAW 2u
int emirwv_abba_host(emirwv_plan *plan, const void *h_in, int in_dtype, int nframes,
                     const int8_t *h_labels, const double *h_offset, int method, double low,
                     double high, float *h_out, uint32_t flags);

/* h_cover [nslitlets][2] int32: output columns [lo, hi) of each slitlet that can hold flux
 * (0, 0 for a missing slitlet); everything outside is exactly zero in every product. */
int emirwv_plan_get_coverage(const emirwv_plan *plan, int32_t *h_cover);

void *emirwv_host_alloc(int64_t nbytes);   /* pinned host memory, NULL on failure */
void emirwv_host_free(void *ptr);

/*
 * Kernel 1 alone: the rectified slitlet rows before the wavelength rebin,
 * d_rect [39][bbw] for rows min_row_rectified .. max_row_rectified+1 of slitlet
 * islitlet (1-based).  Serves Slitlet2D.rectify (slitlet2d.py:381-431).
 */
int emirwv_rectify_slitlet(const emirwv_plan *plan, const void *d_frame, int islitlet,
                           void *d_rect, void *cuda_stream);

/*
 * rectify2d on an arbitrary float64 image with an explicit polynomial map (rectified ->
 * original, term order 1, x, y, x^2, xy, y^2, ...): output has the shape of the input.
 * Replaces numina.array.distortion.rectify2d as called by Slitlet2D.rectify
 * (slitlet2d.py:413-423) with ttd_* (direct) or tti_* (inverse=True) coefficients.
 */
int emirwv_rectify2d_host(const double *h_image, int naxis2, int naxis1, const double *aij,
                          const double *bij, int ncoef, int resampling, int max_order,
                          double *h_out);

/*
 * median_slitlets_rectified, modes 0 and 1 (median_slitlets_rectified.py:31-164, the consumer
 * right behind the path: stare.py:144-153, refine_rectwv_coeff.py:109-112): for every
 * slitlet and column the median over its rows_per_slitlet (38) rows of the rectified
 * product, with numpy.median's semantics (mean of the two middle values, NaN if any NaN).
 *   d_in  [nframes][nslitlets * rows_per_slitlet][naxis1]
 *   mode 0: d_out has the shape of d_in, the median spectrum repeated over the slitlet rows
 *   mode 1: d_out [nframes][nslitlets][naxis1]
 * Asynchronous on the given stream.  (Mode 2 of the reference is mode 1 followed by a masked
 * median of at most 55 values per column, done by the Python shim.)
 */
int emirwv_median_slitlets(const void *d_in, void *d_out, int nframes, int nslitlets,
                           int rows_per_slitlet, int naxis1, int mode, int dtype,
                           void *cuda_stream);

/* Same from host buffers (copies in, runs, copies out, returns when h_out is complete). */
int emirwv_median_slitlets_host(const void *h_in, void *h_out, int nframes, int nslitlets,
                                int rows_per_slitlet, int naxis1, int mode, int dtype);

/*
 * ABBA combination with method "mean" (recipes/spec/abba.py:579-606): per-rank partial sums
 * of the rectified frames per nodding position, float64, in frame order.
 *   d_frames [nframes][npix] plan dtype; d_labels [nframes] int8: +1 = A, -1 = B, 0 = unused
 *   d_sum_a, d_sum_b [npix] float64; accumulate != 0 adds to their current contents
 * The sums of all ranks are added by a collective (NCCL reduce, pyemir_b200/abba.py), then
 * emirwv_abba_finish forms float32(mean A) - float32(mean B) like abba.py:604-610.
 */
int emirwv_abba_partial(const void *d_frames, const int8_t *d_labels, int nframes, int64_t npix,
                        int dtype, double *d_sum_a, double *d_sum_b, int accumulate,
                        void *cuda_stream);
int emirwv_abba_finish(const double *d_sum_a, const double *d_sum_b, int count_a, int count_b,
                       int64_t npix, float *d_out, void *cuda_stream);

/*
 * Pre-reduction in front of the path: the `flow(newimg)` of the recipes (recipes/spec/abba.py:371,
 * core/recipe.py:25-61) for the element-wise correctors -- bias and dark subtraction (numina
 * BiasCorrector / DarkCorrector, restated) and pyemir's FlatFieldCorrector
 * (processing/flatfield.py:27-70: flat values <= 0 count as 1.0; data / flat, cast to float32).
 *   d_in  [nframes][npix] EMIRWV_F32 / EMIRWV_F64 / EMIRWV_U16 (raw detector counts)
 *   d_out [nframes][npix] float32 (may alias d_in when in_dtype is EMIRWV_F32)
 *   d_bias, d_dark, d_flat [npix] of cal_dtype (EMIRWV_F32 / EMIRWV_F64); NULL skips the stage
 * Stages in the order of the flow (bias, dark, flat), each rounded to float32, with numpy's
 * type promotion (a float64 operand makes the stage float64).  Asynchronous on the stream.
 * Runs on the device that owns d_in.
 */
#define EMIRWV_U16 2
int emirwv_prereduce(const void *d_in, int in_dtype, float *d_out, int nframes, int64_t npix,
                     const void *d_bias, const void *d_dark, const void *d_flat, int cal_dtype,
                     void *cuda_stream);

/*
 * The same with the first stage of the flow in front: the bad-pixel correction (numina
 * BadPixelCorrector built by get_corrector_p, core/correctors.py:23-42; restated as
 * numina.array.bpm.process_bpm_median with its 5 x 5 window).
 *   d_bpm [naxis2][naxis1] uint8, non-zero = bad pixel, NULL skips the stage: a bad pixel becomes
 *   the median of the good pixels of the raw frame within +-2 rows and columns (clipped at the
 *   detector edges; 0 when there is none), then goes through bias, dark and flat like any other.
 * The element-wise pass reads every frame once; a second small kernel looks at the mask only and
 * rewrites the bad pixels (their windows come from L2).  d_out must not alias d_in when d_bpm is
 * given.
 */
int emirwv_prereduce_ex(const void *d_in, int in_dtype, float *d_out, int nframes, int naxis2,
                        int naxis1, const uint8_t *d_bpm, const void *d_bias, const void *d_dark,
                        const void *d_flat, int cal_dtype, void *cuda_stream);

/*
 * ABBA combination with method "median" (recipes/spec/abba.py:132-137,579-601; numina's
 * combine.median restated): per pixel the median over the selected images of a stack, with
 * numpy.median's semantics (mean of the two middle values, NaN if any sample is NaN).
 *   d_frames [nframes][npix] plan dtype; h_index [nsel] HOST array: the images that take part
 *   (the A or the B images of the sequence), 1 <= nsel <= 64; d_out [npix] float64.
 * emirwv_abba_finish(med_a, med_b, 1, 1, ...) then forms float32(A) - float32(B).  Frames that
 * are sharded over ranks are first exchanged by row slabs (pyemir_b200/abba.py).
 */
int emirwv_stack_median(const void *d_frames, int nframes, const int32_t *h_index, int nsel,
                        int64_t npix, int dtype, double *d_out, void *cuda_stream);

/* Both medians and the difference in one pass: d_out [npix] float32 =
 * float32(median of the h_index_a images) - float32(median of the h_index_b images); the stack
 * is read once and no float64 intermediate is written (1 <= na, nb <= 64). */
int emirwv_abba_median(const void *d_frames, int nframes, const int32_t *h_index_a, int na,
                       const int32_t *h_index_b, int nb, int64_t npix, int dtype, float *d_out,
                       void *cuda_stream);

/*
 * ABBA combination with method "sigmaclip" (recipes/spec/abba.py:132-137,579-601; numina's
 * combine.sigmaclip restated): per pixel, repeat { mean and unbiased standard deviation of the
 * samples still in; drop the samples outside [mean - low*std, mean + high*std] } until nothing
 * is dropped; the result is the last mean (float64, sums in frame order).  Same stack layout,
 * selection arrays and limits (1 <= nsel <= 64) as emirwv_stack_median / emirwv_abba_median.
 */
int emirwv_stack_sigmaclip(const void *d_frames, int nframes, const int32_t *h_index, int nsel,
                           int64_t npix, int dtype, double low, double high, double *d_out,
                           void *cuda_stream);
int emirwv_abba_sigmaclip(const void *d_frames, int nframes, const int32_t *h_index_a, int na,
                          const int32_t *h_index_b, int nb, int64_t npix, int dtype, double low,
                          double high, float *d_out, void *cuda_stream);

/*
 * The per-frame vertical offsets of the ABBA recipe (recipes/spec/abba.py:545-558:
 * `shift_image2d(data, yoffset=-offset).astype("float32")` for every frame whose offset is not
 * zero; numina's shift_image2d = rectify2d with a translation map, area-overlap resampling):
 * all frames of a device-resident stack in one launch.
 *   d_in  [nframes][naxis2][naxis1] in_dtype;  d_out the same shape in out_dtype (not d_in)
 *   h_yoffset [nframes] HOST array: the yoffset argument of shift_image2d for every frame (the
 *   pixel at row y of the result covers rows y - yoffset -+ 0.5 of the input); a frame whose
 *   offset is 0 is copied (cast).
 * float64 arithmetic with the weights of the general clipping, bit for bit.  Replaces one
 * emirwv_rectify2d_host call (host image in, two allocations, host image out) per frame.
 */
int emirwv_shift_rows(const void *d_in, int in_dtype, void *d_out, int out_dtype, int nframes,
                      int naxis2, int naxis1, const double *h_yoffset, void *cuda_stream);

/*
 * The exchanges of a sharded sequence for a caller that owns an NCCL communicator (one rank per
 * GPU; recipes/spec/abba.py:351-381 sharded over ranks, :579-606 combined on one).  NCCL is NOT
 * linked: the entry points (ncclGroupStart/End, ncclSend, ncclRecv, ncclReduce) are looked up in
 * the calling process, else in libnccl.so.2 (EMIRWV_NCCL_LIB names another file);
 * EMIRWV_E_CUDA when there is none.  `nccl_comm` is the caller's ncclComm_t.  Asynchronous on
 * the stream; every rank of the communicator must make the call.
 *
 * emirwv_gather: the products of all ranks on `root`, in rank order.
 *   d_local [counts[rank]] frames of frame_bytes bytes each (this rank's products)
 *   d_root  [sum(counts)] frames on `root` (ignored elsewhere)
 * emirwv_abba_reduce: method "mean" -- the float64 sums of emirwv_abba_partial of every rank
 *   added in place on `root` (emirwv_abba_finish then forms float32(mean A) - float32(mean B)).
 */
int emirwv_gather(const void *d_local, const int32_t *counts, int world, int rank,
                  int64_t frame_bytes, void *d_root, void *nccl_comm, int root,
                  void *cuda_stream);
int emirwv_abba_reduce(double *d_sum_a, double *d_sum_b, int64_t npix, void *nccl_comm, int root,
                       void *cuda_stream);

/*
 * ABBA combination, method "mean", FUSED with its exchange over NVLink peer memory (one process
 * per GPU on one box; no collective library on the data path).  Every rank creates a workspace
 * (one device allocation), the 64-byte IPC handles of all ranks are exchanged by the caller (any
 * channel: MPI, torch.distributed, a file) and connected; then, per sequence:
 *   earlier passes over a rank's rectified frames: emirwv_abba_partial (local float64 sums);
 *   the LAST pass: emirwv_abba_mean_push -- the kernel that adds the rank's frames pushes every
 *   sum straight into the receive buffer of the rank that OWNS the pixel (pixels are dealt out in
 *   `world` slabs), flags travel with system-scope release / acquire, every owner adds the
 *   `world` contributions of its slab in rank order, forms float32(mean A) - float32(mean B)
 *   (abba.py:604-606) and stores it into the image on `root`.
 *   d_sum_a / d_sum_b: the sums of earlier passes (both NULL: none);  count_a / count_b: A and B
 *   frames of the WHOLE sequence;  *d_image: on `root` the [npix] float32 result inside the
 *   workspace (valid in stream order; overwritten by the next call), NULL
 *   elsewhere.  Every rank must make the call, in the same order.
 * emirwv_peer_timed_out: 1 when a wait for a peer gave up (a rank missing): synchronises.
 */
typedef struct emirwv_peer emirwv_peer;   /* opaque */
int emirwv_peer_create(int64_t npix, int rank, int world, emirwv_peer **peer, void *h_handle64);
int emirwv_peer_connect(emirwv_peer *peer, const void *h_handles /* [world][64] */);
/* Shutting down: every rank disconnects (closes the workspaces of the others), the ranks meet at
 * a barrier of the caller's, then every rank destroys its own workspace -- memory that another
 * process still has open must not be freed. */
int emirwv_peer_disconnect(emirwv_peer *peer);
void emirwv_peer_destroy(emirwv_peer *peer);
int emirwv_peer_timed_out(emirwv_peer *peer);
int emirwv_abba_mean_push(emirwv_peer *peer, const void *d_frames, const int8_t *d_labels,
                          int nframes, int dtype, const double *d_sum_a, const double *d_sum_b,
                          int count_a, int count_b, int root, float **d_image, void *cuda_stream);

/*
 * Parity hooks (tests): the plan's geometry for one slitlet, copied to the host.
 *   h_row, h_col [39][bbw]       source pixel of tap (0,0) in bounding-box array
 *                                coordinates (for EMIRWV_NEAREST: the index maps
 *                                iyyy, ixxx of rectify2d)
 *   h_ok         [39][bbw]       EMIRWV_NEAREST: 1 when the source pixel exists
 *   h_weights    [39][bbw][K][K] float64 weights as used by the kernels (NULL ok)
 */
int emirwv_debug_geometry(const emirwv_plan *plan, int islitlet, int32_t *h_row,
                          int32_t *h_col, uint8_t *h_ok, double *h_weights);

/*
 *   h_idx [naxis1_out+1] source interval of every output border
 *   h_t   [naxis1_out+1] offset inside that interval (wavelength units)
 */
int emirwv_debug_borders(const emirwv_plan *plan, int islitlet, int32_t *h_idx,
                         double *h_t);

#ifdef __cplusplus
}
#endif
#endif /* EMIRWV_H */
