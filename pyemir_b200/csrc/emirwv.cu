// emirwv.cu — sm_100a implementation of the apply_rectwv_coeff hot path.
//
// Reference path: emirdrp/processing/wavecal/apply_rectwv_coeff.py:35-287 and the
// numina routines it calls.  Design:
//
//   plan  (once per RectWaveCoeff, float64, on the GPU + a little host code)
//     k_extent / k_build_geometry : 2-D polynomial map of every output pixel
//         (its centre, or its four corners) and the exact overlap areas with the
//         source pixels  ->  frame-independent records  tab[s][y][j] = {w[K*K], col, slots}
//     build_border_tables (host)  : for every border of the linear wavelength grid
//         the source interval it falls in and the fractional position inside it
//     build_tiles (host)          : work items (slitlet, run of output columns cut into one
//         piece per warp) with the source window each one needs
//
//   run   (per batch of frames; one fused kernel, nothing intermediate touches HBM)
//     k_rectwv<T,K,G>: three CTAs per SM; a copy warp streams the source window of G frames
//       (TMA tensor copies: one box of 256 columns x 1 row x G frames per source row), the
//       staged table rows (bulk copies) and the zero runs outside the wavelength coverage
//       (bulk stores) through shared memory, mbarriers; every compute warp owns a piece of
//       the tile and marches down its 39 rows:
//         A: rectified flux = K*K weighted gather            (kernel 1 of the spec)
//         C: Steffen-limited end slopes of every source pixel (kernel 2)
//         D: cumulative-flux difference at the output borders (kernel 2), coalesced
//            store, first/last non-zero channel per slitlet   (kernel 3)
//
// No tensor cores (there is no dense contraction), no CPU fallback.
//
// One translation unit; the parts live next to this file and are included below:
//   geom.cuh            2-D polynomial map and polygon clipping (shared with the CPU hostcheck)
//   pixel_ops.cuh       pre-reduction stages, stack-median networks (shared with the hostcheck)
//   plan_kernels.cuh    k_extent, k_build_geometry, k_rectify2d, parity-hook helpers
//   device_utils.cuh    TMA bulk copies, mbarriers, packed float32 pairs
//   side_kernels.cuh    zero fill, median_slitlets_rectified, ABBA mean / median / sigma clipping,
//                       per-frame vertical offsets, pre-reduction with its bad-pixel stage
//   rectwv_kernel.cuh   the hot kernel k_rectwv
//   peer_kernels.cuh    ABBA mean fused with its exchange over NVLink peer memory
//   plan_build.inl      host side of a plan (launch configuration, border tables, tiles)
// and this file keeps the constants, the device structs, the plan object and the C ABI.

#include <cuda.h>               // CUtensorMap and its enums (types only: libcuda is not linked)
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only: nothing is linked

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <dlfcn.h>
#include <string>
#include <thread>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <type_traits>
#include <vector>

#include "emirwv.h"

#ifndef EMIRWV_NSTAGE
#define EMIRWV_NSTAGE 4
#endif
#ifndef EMIRWV_WIN_PITCH
#define EMIRWV_WIN_PITCH 256
#endif
#ifndef EMIRWV_PXCAP
#define EMIRWV_PXCAP 248
#endif
#ifndef EMIRWV_WIN_ROWS
#define EMIRWV_WIN_ROWS 11
#endif
#ifndef EMIRWV_NOFILL
#define EMIRWV_NOFILL 0        // 1: the zero runs are always written by k_zero_fill (experiment)
#endif
#ifndef EMIRWV_PREFETCH
#define EMIRWV_PREFETCH 0      // rows of L2 prefetch ahead of the copy pipeline (0: none)
#endif
#ifndef EMIRWV_RECPF
#define EMIRWV_RECPF 0         // 1: prefetch the next row's full records in the rare-footprint path
#endif
#ifndef EMIRWV_DIAG
#define EMIRWV_DIAG 0          // timing diagnostics that break the result (scripts/build_variants.py)
#endif
#ifndef EMIRWV_MINBLOCKS
#define EMIRWV_MINBLOCKS 3     // resident CTAs per SM the float32 kernels are compiled for (72 registers)
#endif
#ifndef EMIRWV_MINBLOCKS_F64
#define EMIRWV_MINBLOCKS_F64 3 // the same for the float64 kernels
#endif
#include "geom.cuh"
#include "pixel_ops.cuh"

namespace {

using namespace emirwv;

constexpr int NROWS = EMIRWV_ROWS_SCANNED;   // 39 rectified rows per slitlet
constexpr int W = 2048;                      // EMIR_NAXIS1: pitch of per-column tables
constexpr int MAXK = 4;                      // largest supported footprint
constexpr int NWARPS = 8;                     // compute warps: two per SM sub-partition and CTA
constexpr int NTHREADS = 32 * (NWARPS + 1);   // threads per CTA: compute warps + one copy warp
constexpr int ZFILL_BYTES = 2048;             // zeros in shared memory: source of the bulk stores that
                                              // write the runs without wavelength coverage
constexpr int TILE_K_MAX = 31 * NWARPS;       // output columns per tile (a warp does 31)
constexpr int WIN_ROWS = EMIRWV_WIN_ROWS;                  // rolling source window: rows per frame
constexpr int RING_BIAS = 12 * 1024;          // makes (row + bias) % WIN_ROWS non-negative
constexpr int PXCAP = EMIRWV_PXCAP;            // rectified columns a CTA can stage (8 x 29 + 3 at most,
                                              // rounded out to multiples of 4: 16-byte copies)
constexpr int WIN_PITCH = EMIRWV_WIN_PITCH;                // source window columns, per frame: a multiple of
                                              // 32 banks, so lanes on two ring rows never collide
constexpr int WARP_PX = 30;                   // rectified columns a warp owns
constexpr int WARP_BORDERS = 31;              // output columns a warp evaluates
constexpr int NSTAGE_MAX = EMIRWV_NSTAGE;                 // rectified rows in flight: a row's copies are
                                              // requested NSTAGE rows ahead of its use (a kernel
                                              // template parameter: NSTAGE_MAX, or 2 for float64
                                              // plans on request -- 96-byte records, see smem_bytes)
constexpr int BAR_BYTES = ((2 * NSTAGE_MAX * 8) + 63) / 64 * 64;   // "full" and "done" mbarriers
// one staged table row of a tile: PXCAP block-weight quadruples, then PXCAP meta words
__host__ __device__ constexpr int stage_bytes(int esize) { return PXCAP * (4 * esize + 4); }
// Tables per rectified pixel (s, y, j), frame independent:
//   quad  [4]   the weights of the 2 x 2 block of source pixels the pixel is made of (K <= 2, and
//               the footprints of a larger K that fit such a block: ~88 % of them), zeros else
//   meta        packed uint32: bits 0-15 source column of the block (int16, frame coordinates),
//               16-19 / 20-23 ring slots of its two source rows in the hot kernel's rolling
//               window, bit 24 "the 2 x 2 block is the whole footprint"
//   w [RS]      the full record: K*K weights followed by the packed source coordinates, the ring
//               slots and the flags, padded to a multiple of 4 elements (16-byte aligned).
// The hot kernel stages quad + meta (20 / 36 bytes) in shared memory; a warp that holds a pixel
// whose footprint does not fit 2 x 2 reads the full records from global memory instead.
__host__ __device__ constexpr int rec_size(int K) { return (K * K + 1 + 3) & ~3; }
constexpr uint32_t META_COMPACT = 1u << 24;
constexpr int HOST_SLOTS = 3;                // pipeline depth of emirwv_run_host
constexpr int SCHED_POOL = 1024;             // item counters: launches of one plan in flight
constexpr int SMEM_LIMIT = 227 * 1024;

thread_local std::string g_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CUDA_TRY(call)                                                              \
    do {                                                                            \
        cudaError_t err_ = (call);                                                  \
        if (err_ != cudaSuccess)                                                    \
            return fail(EMIRWV_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(err_)); \
    } while (0)

// ------------------------------------------------------------------ device structs
struct SlitDev {
    int valid;
    int ns1, nc1;       // 1-based origin of the bounding box in the frame
    int bbh, bbw;
    int min_row;
    int order;
    int pad;
    double a[EMIRWV_MAX_COEF];
    double b[EMIRWV_MAX_COEF];
};

// Work item of the hot kernel: a block of output columns of one slitlet, all rows.
struct Tile {
    int slit;           // 0-based slitlet
    int k0, nk;         // output columns [k0, k0+nk)
    int jlo, npx;       // rectified columns needed: [jlo, jlo+npx)
    int c0, wcols;      // source window columns (multiples of 4: 16-byte copies)
    int empty;          // no wavelength coverage: the tile is exactly zero
    int nchan;          // bbw of the slitlet
    uint8_t kw[12];     // kw[w] .. kw[w+1]-1: output columns (relative to k0) of warp w
    int32_t span[40];   // per scanned row: source rows needed, lo | hi << 16 (int16 each),
                        // monotone envelopes (rolling window)
};
static_assert(sizeof(Tile) % 16 == 0, "tile descriptors are copied in 16-byte chunks");

struct GeomParams {
    const SlitDev *slit;
    int nsl;
    int naxis1, naxis2;
    int offx, offy;
    int mode, K;
    int32_t *base;
    uint32_t *meta;         // per pixel: packed block column, ring slots, flag
    uint8_t *blockrow;      // per pixel: first source row with weight, relative to base (host)
    int *kmax;              // k_extent result
};

// Per source pixel i of a slitlet (frame independent), one 16/32-byte load:
//   qm = h_i / h_{i-1},  qp = h_i / h_{i+1}   (h: width of the pixel in wavelength)
//   r0 = h_i / (h_{i-1} + h_i),  r1 = h_{i+1} / (h_i + h_{i+1})
template <typename T>
struct alignas(4 * sizeof(T)) PixGeom {
    T qm, qp, r0, r1;
};

// Per output border k of a slitlet: fractional position tau in source pixel idx.
template <typename T>
struct alignas(2 * sizeof(T)) Border {
    T tau;
    int32_t idx;
};

// Per source pixel and frame (shared memory): its flux and the Steffen-limited
// slopes of the cumulative flux at its two ends, in flux-per-pixel units.
template <typename T>
struct alignas(4 * sizeof(T)) Knot {
    T rect;   // flux of the pixel = y_{i+1} - y_i of the cumulative flux
    T ya;     // h_i * y'(left knot)
    T yb;     // h_i * y'(right knot)
    T pad;
};

template <typename T>
struct RunParams {
    const T *in;
    T *out;
    int32_t *jmm;
    int nframes;
    const Tile *tiles;
    const int32_t *base;
    const T *w;             // full records (global memory path)
    const char *tilepack;   // per live tile and scanned row: the staged table row (stage_bytes)
    const Border<T> *border;
    const PixGeom<T> *pix;
    int naxis2;
    int nout, NB;           // naxis1_out, naxis1_out + 1
    int rps, nrows_out;     // 38, 55*38
    int nsl;
    int vec_ok;             // rows of the output are 16-byte aligned
    int nlive_tiles;        // tiles [0, nlive_tiles) have wavelength coverage
    int ndead_tiles;        // the tiles after them are exactly zero
    int ngroups;            // ceil(nframes / G)
    const int2 *fillruns;   // zero runs of one frame: {first element, bytes}, or NULL
    int nfillruns;          // dead tiles x output rows
    int *sched;             // dynamic work distribution: item counter of this launch (or NULL)
    int early;              // first rows of the next item requested at the end of the current one
};

// flags in the top byte of a table record's slot word
constexpr uint32_t REC_COMPACT = 1u;          // bits 1-2: a0, bits 3-4: b0 of the 2 x 2 block

// weight of tap (a, b) of a table record (host and plan kernels; the hot kernel has its own
// unpacking)
template <typename T>
__host__ __device__ inline T record_weight(const T *rec, int K, int a, int b) {
    const uint32_t flags = *reinterpret_cast<const uint32_t *>(rec + K * K + 1) >> 24;
    if (!(flags & REC_COMPACT)) return rec[a * K + b];
    const int da = a - (int)((flags >> 1) & 3u), db = b - (int)((flags >> 3) & 3u);
    return (da >= 0 && da < 2 && db >= 0 && db < 2) ? rec[da * 2 + db] : T(0);
}

__host__ __device__ inline int32_t pack_base(int r, int c) {
    r = max(-30000, min(30000, r));
    c = max(-30000, min(30000, c));
    return (int32_t)(((uint32_t)(r & 0xffff) << 16) | (uint32_t)(c & 0xffff));
}
__host__ __device__ inline int base_row(int32_t p) { return (int)(int16_t)((uint32_t)p >> 16); }
__host__ __device__ inline int base_col(int32_t p) { return (int)(int16_t)((uint32_t)p & 0xffffu); }

// ring slot of a frame row in the hot kernel's rolling window (rows far outside the detector
// carry no weight: any valid slot will do)
__host__ __device__ inline uint32_t ring_slot(int r) {
    r = max(-RING_BIAS, min(30000, r));
    return (uint32_t)((r + RING_BIAS) % WIN_ROWS);
}

__host__ __device__ inline int span_lo(int32_t v) { return (int)(int16_t)((uint32_t)v & 0xffffu); }
__host__ __device__ inline int span_hi(int32_t v) { return (int)(int16_t)((uint32_t)v >> 16); }

__device__ inline double clampd(double v) { return fmin(fmax(v, -1.0e6), 1.0e6); }

// corners of output pixel (j, i) in numina's order:
// (j-.5,i-.5) (j-.5,i+.5) (j+.5,i+.5) (j+.5,i-.5)
__device__ inline void map_corners(const SlitDev &sd, int j, int row, double *qu, double *qv) {
    const double dj[4] = {-0.5, -0.5, 0.5, 0.5};
    const double di[4] = {-0.5, 0.5, 0.5, -0.5};
    for (int c = 0; c < 4; ++c) {
        double u, v;
        fmap(sd.order, sd.a, sd.b, (double)j + dj[c], (double)row + di[c], u, v);
        qu[c] = clampd(u);
        qv[c] = clampd(v);
    }
}

#include "plan_kernels.cuh"
#include "device_utils.cuh"
#include "side_kernels.cuh"
#include "rectwv_kernel.cuh"
#include "peer_kernels.cuh"

}  // namespace

// =================================================================== plan object
struct emirwv_plan {
    emirwv_plan_desc desc;
    int device = 0;
    int dtype = 0, mode = 0, K = 1;
    int nsl = 0, rps = 0, nout = 0, NB = 0;
    size_t esize = 4;

    SlitDev *d_slit = nullptr;
    int32_t *d_base = nullptr;
    void *d_w = nullptr, *d_quad = nullptr;
    uint32_t *d_meta = nullptr;
    char *d_tilepack = nullptr;        // [live tile][39][stage_bytes]: what the hot kernel stages
    void *d_border = nullptr, *d_pix = nullptr;
    Tile *d_tiles = nullptr;
    int2 *d_fillruns = nullptr;        // [ntiles_empty * rps] zero runs of one frame
    int2 *d_cover = nullptr;           // [nsl] covered output columns (sparse copy-back kernel)

    std::vector<SlitDev> h_slit;
    std::vector<int32_t> h_base;
    std::vector<uint8_t> h_blockrow;   // bits 0-3 first used source row - base row, bit 4 compact
    std::vector<int32_t> h_idx;
    std::vector<double> h_t;
    std::vector<Tile> h_tiles;

    int ntiles = 0, ntiles_empty = 0;
    int tile_k = 0;
    int wrows_max = 0, wcols_max = 0, npx_max = 0, max_span = 0;
    std::vector<int32_t> h_cover;      // [nsl][2]: live output columns [lo, hi) of a slitlet
    int *d_sched = nullptr;            // [SCHED_POOL] item counters, one per launch in flight
    mutable std::atomic<unsigned> sched_next{0};
    int dynamic_sched = 1;
    int early_rows = 1;
    int nstage = NSTAGE_MAX;           // stages of the hot kernel's row pipeline
    int tune_G = 0;
    int num_sms = 148;
    int64_t table_bytes = 0;

    // side stream for the zero-fill kernel (memory bound) so that it overlaps with the
    // hot kernel (issue bound); fork/join with events around every run
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    mutable std::mutex aux_mu;

    // calibration images of the flow in front of the path (emirwv_plan_set_calibration)
    uint8_t *d_bpm = nullptr;
    void *d_bias = nullptr, *d_dark = nullptr, *d_flat = nullptr;
    int cal_dtype = EMIRWV_F32;

    std::mutex mu;
    struct Slot {
        void *d_in = nullptr;
        void *d_out = nullptr;
        void *d_raw = nullptr;         // raw frames before the flow (allocated on first use)
        int32_t *d_jmm = nullptr;
        cudaStream_t stream = nullptr;
        cudaEvent_t ev_ready = nullptr;   // the slot's products are complete
        cudaEvent_t ev_free = nullptr;    // the consumer stream is through with them
    } slots[HOST_SLOTS];
    int ws_chunk = 0;
    size_t ws_raw_bytes = 0;           // bytes per frame the d_raw buffers were sized for
    // ABBA host entry: consumer stream, float64 sums, result image, frame stack (median /
    // sigma clipping), shifted chunk
    cudaStream_t acc = nullptr;
    double *d_sums = nullptr;
    float *d_image = nullptr;
    void *d_stack = nullptr;
    size_t stack_bytes = 0;
    void *d_shift = nullptr;
    int shift_chunk = 0;
};

namespace {

#include "plan_build.inl"

// The entry points either side of the path take no plan: they run on the device that owns the
// buffer they are given (not on whatever device happens to be current in the calling thread).
int device_of(const void *ptr) {
    cudaPointerAttributes attr;
    if (ptr != nullptr && cudaPointerGetAttributes(&attr, ptr) == cudaSuccess &&
        (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged))
        return attr.device;
    cudaGetLastError();
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}

int sm_count(int dev) {
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        cudaGetLastError();
        sms = 148;
    }
    return sms;
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

int emirwv_version(void) { return EMIRWV_VERSION; }

const char *emirwv_last_error(void) { return g_error.c_str(); }

int emirwv_device_count(void) {
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess)
        return fail(EMIRWV_E_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(err));
    return n;
}

int emirwv_get_device(void) {
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess)
        return fail(EMIRWV_E_CUDA, "cudaGetDevice failed: %s", cudaGetErrorString(err));
    return dev;
}

int emirwv_set_device(int device) {
    CUDA_TRY(cudaSetDevice(device));
    return EMIRWV_OK;
}

int emirwv_plan_create(const emirwv_plan_desc *desc, emirwv_plan **out) {
    if (desc == nullptr || out == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    *out = nullptr;
    int rc = validate(desc);
    if (rc) return rc;
    emirwv_plan *pl = new (std::nothrow) emirwv_plan();
    if (pl == nullptr) return fail(EMIRWV_E_NOMEM, "out of host memory");
    pl->desc = *desc;
    rc = build_plan(pl);
    if (rc == EMIRWV_OK) {
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess)
            rc = fail(EMIRWV_E_CUDA, "plan build failed: %s", cudaGetErrorString(err));
    }
    if (rc) {
        const std::string keep = g_error;
        emirwv_plan_destroy(pl);
        g_error = keep;
        return rc;
    }
    *out = pl;
    return EMIRWV_OK;
}

void emirwv_plan_destroy(emirwv_plan *pl) {
    if (pl == nullptr) return;
    DeviceGuard guard(pl->device);
    release_workspace(pl);
    if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
    if (pl->ev_join) cudaEventDestroy(pl->ev_join);
    if (pl->aux) cudaStreamDestroy(pl->aux);
    cudaFree(pl->d_slit);
    cudaFree(pl->d_base);
    cudaFree(pl->d_w);
    cudaFree(pl->d_quad);
    cudaFree(pl->d_meta);
    cudaFree(pl->d_tilepack);
    cudaFree(pl->d_border);
    cudaFree(pl->d_pix);
    cudaFree(pl->d_tiles);
    cudaFree(pl->d_fillruns);
    cudaFree(pl->d_cover);
    cudaFree(pl->d_sched);
    cudaFree(pl->d_bpm);
    cudaFree(pl->d_bias);
    cudaFree(pl->d_dark);
    cudaFree(pl->d_flat);
    cudaFree(pl->d_sums);
    cudaFree(pl->d_image);
    cudaFree(pl->d_stack);
    cudaFree(pl->d_shift);
    if (pl->acc) cudaStreamDestroy(pl->acc);
    delete pl;
}

int emirwv_plan_get_info(const emirwv_plan *pl, emirwv_plan_info *info) {
    if (pl == nullptr || info == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    memset(info, 0, sizeof(*info));
    const LaunchCfg cfg = choose_cfg(pl, 1 << 20);
    info->dtype = pl->dtype;
    info->resampling = pl->mode;
    info->footprint = pl->K;
    info->naxis1_out = pl->nout;
    info->naxis2_out = pl->nsl * pl->rps;
    info->ntiles = pl->ntiles;
    info->ntiles_empty = pl->ntiles_empty;
    info->tile_k = pl->tile_k;
    info->stages = pl->nstage;
    info->window_rows = pl->wrows_max;
    info->window_cols = pl->wcols_max;
    info->max_span = pl->max_span;
    info->frames_per_cta = cfg.G;
    info->threads = cfg.threads;
    info->smem_bytes = (int32_t)cfg.smem;
    {
        DeviceGuard guard(pl->device);
        int per_sm = 0;
        const int rc = pl->dtype == EMIRWV_F32
                           ? launch_t<float>(pl, pl->d_w, pl->d_w, 1 << 20, nullptr, nullptr, &per_sm)
                           : launch_t<double>(pl, pl->d_w, pl->d_w, 1 << 20, nullptr, nullptr, &per_sm);
        if (rc) return rc;
        info->ctas_per_sm = per_sm;
    }
    info->table_bytes = pl->table_bytes;
    info->in_frame_bytes = (int64_t)pl->desc.naxis1 * pl->desc.naxis2 * pl->esize;
    info->out_frame_bytes = (int64_t)pl->nsl * pl->rps * pl->nout * pl->esize;
    return EMIRWV_OK;
}

int emirwv_plan_set_tuning(emirwv_plan *pl, int frames_per_cta) {
    if (pl == nullptr) return fail(EMIRWV_E_VALUE, "null plan");
    if (frames_per_cta < 0 || frames_per_cta > 4)
        return fail(EMIRWV_E_VALUE, "frames_per_cta must be 0 (default), 1, 2 or 4");
    pl->tune_G = frames_per_cta;
    return EMIRWV_OK;
}

int emirwv_run(const emirwv_plan *pl, const void *d_in, void *d_out, int nframes,
               int32_t *d_jminmax, void *cuda_stream) {
    return emirwv_run_ex(pl, d_in, d_out, nframes, d_jminmax, cuda_stream, 0u);
}

int emirwv_run_ex(const emirwv_plan *pl, const void *d_in, void *d_out, int nframes,
                  int32_t *d_jminmax, void *cuda_stream, uint32_t flags) {
    if (pl == nullptr || d_in == nullptr || d_out == nullptr)
        return fail(EMIRWV_E_VALUE, "null argument");
    if (nframes < 0) return fail(EMIRWV_E_VALUE, "negative frame count");
    if (flags & ~(uint32_t)EMIRWV_RUN_OUT_ZEROED) return fail(EMIRWV_E_VALUE, "unknown flag");
    if (nframes == 0) return EMIRWV_OK;
    DeviceGuard guard(pl->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const bool skip_fill = (flags & EMIRWV_RUN_OUT_ZEROED) != 0;
    if (pl->dtype == EMIRWV_F32)
        return launch_t<float>(pl, d_in, d_out, nframes, d_jminmax, st, nullptr, skip_fill);
    return launch_t<double>(pl, d_in, d_out, nframes, d_jminmax, st, nullptr, skip_fill);
}

int emirwv_plan_get_coverage(const emirwv_plan *pl, int32_t *h_cover) {
    if (pl == nullptr || h_cover == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    memcpy(h_cover, pl->h_cover.data(), pl->h_cover.size() * sizeof(int32_t));
    return EMIRWV_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ host-buffer entries
namespace {

// NVTX ranges around plan build / copies / kernels of the host entries (header-only NVTX 3: no
// library is linked, the calls are no-ops unless a tool is attached)
struct nvtx_range {
    bool open = true;
    explicit nvtx_range(const char *name) { nvtxRangePushA(name); }
    void end() {
        if (open) nvtxRangePop();
        open = false;
    }
    ~nvtx_range() { end(); }
};

// The covered columns of every slitlet of n products, device -> the caller's pinned output
// buffer through its device alias: a warp per (frame, output row) writes the row's covered run
// with 16-byte stores that go out over PCIe; ONE launch per chunk instead of one strided copy
// per slitlet (55 per chunk).
template <typename T>
__global__ void __launch_bounds__(256) k_copy_cover(const T *__restrict__ src, T *__restrict__ dst,
                                                    const int2 *__restrict__ cover, int nsl, int rps,
                                                    int nout, int nframes, int vec_ok) {
    using V = typename Vec16<T>::type;
    constexpr int VN = Vec16<T>::N;
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long units = (long long)nframes * nsl * rps;
    for (long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units;
         u += nwarps) {
        const int2 cv = cover[(int)((u / rps) % nsl)];
        const size_t row = (size_t)u * nout;
        if (vec_ok) {
            for (int c = (cv.x / VN + lane) * VN; c < cv.y; c += 32 * VN) {
                if (c >= cv.x && c + VN <= cv.y) {
                    *reinterpret_cast<V *>(dst + row + c) = __ldcs(reinterpret_cast<const V *>(src + row + c));
                } else {
#pragma unroll
                    for (int j = 0; j < VN; ++j)
                        if (c + j >= cv.x && c + j < cv.y) dst[row + c + j] = src[row + c + j];
                }
            }
        } else {
            for (int c = cv.x + lane; c < cv.y; c += 32) dst[row + c] = src[row + c];
        }
    }
}

int launch_copy_cover(emirwv_plan *pl, const void *d_src, void *d_host_alias, int n, cudaStream_t st) {
    if (pl->d_cover == nullptr) {
        CUDA_TRY(cudaMalloc(&pl->d_cover, pl->h_cover.size() * sizeof(int32_t)));
        CUDA_TRY(cudaMemcpy(pl->d_cover, pl->h_cover.data(), pl->h_cover.size() * sizeof(int32_t),
                            cudaMemcpyHostToDevice));
    }
    const int vec_ok = ((size_t)pl->nout * pl->esize) % 16 == 0 &&
                       (reinterpret_cast<uintptr_t>(d_host_alias) & 15u) == 0;
    const int grid = std::max(1, env_int("EMIRWV_PACK_BLOCKS", pl->num_sms));
    if (pl->dtype == EMIRWV_F32)
        k_copy_cover<float><<<grid, 256, 0, st>>>(static_cast<const float *>(d_src),
                                                  static_cast<float *>(d_host_alias), pl->d_cover,
                                                  pl->nsl, pl->rps, pl->nout, n, vec_ok);
    else
        k_copy_cover<double><<<grid, 256, 0, st>>>(static_cast<const double *>(d_src),
                                                   static_cast<double *>(d_host_alias), pl->d_cover,
                                                   pl->nsl, pl->rps, pl->nout, n, vec_ok);
    CUDA_TRY(cudaGetLastError());
    return EMIRWV_OK;
}

size_t dtype_size(int dtype) { return dtype == EMIRWV_F64 ? 8 : (dtype == EMIRWV_U16 ? 2 : 4); }

// device address of a pinned host buffer the GPU can write to directly, or NULL
void *mapped_alias(void *h_ptr) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, h_ptr) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return attr.type == cudaMemoryTypeHost ? attr.devicePointer : nullptr;
}

// one chunk of frames into a slot and through the kernels: host -> device, the flow of the
// recipes when the frames are raw (bad pixels, bias, dark, flat: one pass), the hot kernel.
// The products land in d_dst (the slot's own buffer or a place in a stack).
int enqueue_chunk(emirwv_plan *pl, emirwv_plan::Slot &sl, const void *h_src, int in_dtype,
                  bool raw, int n, void *d_dst, int32_t *d_jmm) {
    const size_t npix = (size_t)pl->desc.naxis1 * pl->desc.naxis2;
    nvtx_range r_h2d("emirwv: host -> device");
    if (!raw) {
        CUDA_TRY(cudaMemcpyAsync(sl.d_in, h_src, npix * pl->esize * n, cudaMemcpyHostToDevice,
                                 sl.stream));
    } else {
        CUDA_TRY(cudaMemcpyAsync(sl.d_raw, h_src, npix * dtype_size(in_dtype) * n,
                                 cudaMemcpyHostToDevice, sl.stream));
        const int rc = emirwv_prereduce_ex(sl.d_raw, in_dtype, static_cast<float *>(sl.d_in), n,
                                           pl->desc.naxis2, pl->desc.naxis1, pl->d_bpm, pl->d_bias,
                                           pl->d_dark, pl->d_flat, pl->cal_dtype, sl.stream);
        if (rc) return rc;
    }
    r_h2d.end();
    nvtx_range r_k("emirwv: k_rectwv");
    if (pl->dtype == EMIRWV_F32) return launch_t<float>(pl, sl.d_in, d_dst, n, d_jmm, sl.stream);
    return launch_t<double>(pl, sl.d_in, d_dst, n, d_jmm, sl.stream);
}

int ensure_raw(emirwv_plan *pl, int chunk, int in_dtype) {
    const size_t per_frame = (size_t)pl->desc.naxis1 * pl->desc.naxis2 * dtype_size(in_dtype);
    if (pl->ws_raw_bytes >= per_frame && pl->slots[0].d_raw != nullptr) return EMIRWV_OK;
    for (auto &sl : pl->slots) {
        if (sl.d_raw) cudaFree(sl.d_raw);
        sl.d_raw = nullptr;
        CUDA_TRY(cudaMalloc(&sl.d_raw, per_frame * pl->ws_chunk));
    }
    pl->ws_raw_bytes = per_frame;
    (void)chunk;
    return EMIRWV_OK;
}

int check_raw(const emirwv_plan *pl, int in_dtype, bool raw) {
    if (!raw) {
        if (in_dtype != pl->dtype)
            return fail(EMIRWV_E_VALUE, "reduced frames must have the dtype of the plan");
        return EMIRWV_OK;
    }
    if (in_dtype != EMIRWV_F32 && in_dtype != EMIRWV_F64 && in_dtype != EMIRWV_U16)
        return fail(EMIRWV_E_VALUE, "unknown image dtype");
    if (pl->dtype != EMIRWV_F32)
        return fail(EMIRWV_E_VALUE, "raw frames need a float32 plan (the flow rounds to float32)");
    return EMIRWV_OK;
}

// The zeros outside the wavelength coverage (38 % of a J/H/K product) written where they are
// needed instead of shipped: host threads of the library fill the uncovered columns of the
// caller's buffer while the covered ones arrive over PCIe.  The device -> host direction is the
// narrow one (one GPU: the link; eight GPUs of one box: the host's ingest, ~80 GB/s in all,
// profiles/r02_pcie_probe_8gpu.json), so every byte that does not travel counts.
// zeros with non-temporal stores: no read-for-ownership of lines that are overwritten whole, no
// cache pollution next to the DMA traffic into the same buffer
inline void zero_stream(char *p, size_t n) {
#if defined(__x86_64__) && defined(__SSE2__)
    while (n > 0 && (reinterpret_cast<uintptr_t>(p) & 15u)) {
        *p++ = 0;
        --n;
    }
    const __m128i z = _mm_setzero_si128();
    for (; n >= 64; n -= 64, p += 64) {
        _mm_stream_si128(reinterpret_cast<__m128i *>(p), z);
        _mm_stream_si128(reinterpret_cast<__m128i *>(p + 16), z);
        _mm_stream_si128(reinterpret_cast<__m128i *>(p + 32), z);
        _mm_stream_si128(reinterpret_cast<__m128i *>(p + 48), z);
    }
    for (; n >= 16; n -= 16, p += 16) _mm_stream_si128(reinterpret_cast<__m128i *>(p), z);
#endif
    if (n > 0) memset(p, 0, n);
}

void host_zero_fill(const emirwv_plan *pl, char *h_out, int f_begin, int f_end, int f_step) {
    const size_t pitch = (size_t)pl->nout * pl->esize;
    const size_t out_b = (size_t)pl->nsl * pl->rps * pitch;
    for (int f = f_begin; f < f_end; f += f_step)
        for (int s = 0; s < pl->nsl; ++s) {
            const size_t lo = (size_t)pl->h_cover[2 * s] * pl->esize;
            const size_t hi = (size_t)pl->h_cover[2 * s + 1] * pl->esize;
            char *row = h_out + (size_t)f * out_b + (size_t)s * pl->rps * pitch;
            for (int y = 0; y < pl->rps; ++y, row += pitch) {
                if (hi <= lo) {
                    zero_stream(row, pitch);
                } else {
                    if (lo > 0) zero_stream(row, lo);
                    if (hi < pitch) zero_stream(row + hi, pitch - hi);
                }
            }
        }
#if defined(__x86_64__) && defined(__SSE2__)
    _mm_sfence();
#endif
}

int run_host_impl(emirwv_plan *pl, const void *h_in, int in_dtype, bool raw, void *h_out,
                  int nframes, int32_t *h_jminmax, bool sparse) {
    DeviceGuard guard(pl->device);
    std::lock_guard<std::mutex> lock(pl->mu);
    nvtx_range r_all("emirwv_run_host");
    // A batch into a page-locked buffer without the caller's promise: ship the covered columns
    // only and let host threads write the zeros (EMIRWV_HOST_ZERO_THREADS, default 3; 0: every
    // byte travels).  Single frames and pageable buffers take the plain copy.
    std::vector<std::thread> zero_team;
    if (!sparse && nframes >= 8 && mapped_alias(h_out) != nullptr) {
        const int nthreads = std::min(env_int("EMIRWV_HOST_ZERO_THREADS", 3), nframes);
        if (nthreads > 0) {
            sparse = true;
            for (int t = 0; t < nthreads; ++t)
                zero_team.emplace_back(host_zero_fill, pl, static_cast<char *>(h_out), t, nframes,
                                       nthreads);
        }
    }
    struct Joiner {
        std::vector<std::thread> &team;
        ~Joiner() {
            for (auto &th : team) th.join();
        }
    } joiner{zero_team};
    // frames per pipeline slot: 4 keeps the ramp over PCIe short; the sparse copy-back takes 8
    const int chunk = std::max(1, std::min(nframes, env_int("EMIRWV_HOST_CHUNK", sparse ? 8 : 4)));
    int rc = ensure_workspace(pl, chunk);
    if (rc == EMIRWV_OK && raw) rc = ensure_raw(pl, chunk, in_dtype);
    if (rc) return rc;
    const size_t in_b = (size_t)pl->desc.naxis1 * pl->desc.naxis2 * (raw ? dtype_size(in_dtype) : pl->esize);
    const size_t out_b = (size_t)pl->nsl * pl->rps * pl->nout * pl->esize;
    const size_t jmm_b = (size_t)pl->nsl * 2 * sizeof(int32_t);
    // the covered columns can be written straight into a pinned output buffer by a kernel (one
    // launch per chunk); a pageable buffer takes one strided copy per slitlet
    char *alias = sparse && env_int("EMIRWV_HOST_PACK", 1) ? static_cast<char *>(mapped_alias(h_out))
                                                           : nullptr;
    int slot = 0;
    // (an error inside the loop must not return while other slots still copy from / into the
    // caller's buffers: break, then wait for every stream below)
    auto step = [&](int f0, emirwv_plan::Slot &sl) -> int {
        const int n = std::min(chunk, nframes - f0);
        int rc2 = enqueue_chunk(pl, sl, static_cast<const char *>(h_in) + (size_t)f0 * in_b,
                                in_dtype, raw, n, sl.d_out, sl.d_jmm);
        if (rc2) return rc2;
        nvtx_range r_d2h("emirwv: device -> host");
        char *dst = static_cast<char *>(h_out) + (size_t)f0 * out_b;
        if (!sparse) {
            CUDA_TRY(cudaMemcpyAsync(dst, sl.d_out, out_b * n, cudaMemcpyDeviceToHost, sl.stream));
        } else if (alias != nullptr) {
            rc2 = launch_copy_cover(pl, sl.d_out, alias + (size_t)f0 * out_b, n, sl.stream);
            if (rc2) return rc2;
        } else {
            // only the columns with wavelength coverage travel: one strided copy per slitlet
            // (its live columns x 38 rows x n frames); the rest of h_out already holds the
            // zeros the reference writes there
            const size_t pitch = (size_t)pl->nout * pl->esize;
            const size_t rows = (size_t)pl->nsl * pl->rps;
            for (int s = 0; s < pl->nsl; ++s) {
                const int lo = pl->h_cover[2 * s], hi = pl->h_cover[2 * s + 1];
                if (hi <= lo) continue;
                cudaMemcpy3DParms cp;
                memset(&cp, 0, sizeof(cp));
                cp.srcPtr = make_cudaPitchedPtr(sl.d_out, pitch, pitch, rows);
                cp.dstPtr = make_cudaPitchedPtr(dst, pitch, pitch, rows);
                cp.srcPos = cp.dstPos = make_cudaPos((size_t)lo * pl->esize, (size_t)s * pl->rps, 0);
                cp.extent = make_cudaExtent((size_t)(hi - lo) * pl->esize, (size_t)pl->rps, (size_t)n);
                cp.kind = cudaMemcpyDeviceToHost;
                CUDA_TRY(cudaMemcpy3DAsync(&cp, sl.stream));
            }
        }
        if (h_jminmax != nullptr)
            CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char *>(h_jminmax) + (size_t)f0 * jmm_b,
                                     sl.d_jmm, jmm_b * n, cudaMemcpyDeviceToHost, sl.stream));
        return EMIRWV_OK;
    };
    for (int f0 = 0; f0 < nframes && rc == EMIRWV_OK; f0 += chunk, slot = (slot + 1) % HOST_SLOTS)
        rc = step(f0, pl->slots[slot]);
    const std::string keep = g_error;
    for (auto &sl : pl->slots) {
        cudaError_t err = cudaStreamSynchronize(sl.stream);
        if (err != cudaSuccess && rc == EMIRWV_OK)
            rc = fail(EMIRWV_E_CUDA, "stream failed: %s", cudaGetErrorString(err));
    }
    if (rc != EMIRWV_OK && !keep.empty()) g_error = keep;
    return rc;
}

}  // namespace

extern "C" {

int emirwv_run_host(emirwv_plan *pl, const void *h_in, void *h_out, int nframes,
                    int32_t *h_jminmax) {
    return emirwv_run_host_ex(pl, h_in, h_out, nframes, h_jminmax, 0u);
}

int emirwv_run_host_ex(emirwv_plan *pl, const void *h_in, void *h_out, int nframes,
                       int32_t *h_jminmax, uint32_t flags) {
    if (pl == nullptr || h_in == nullptr || h_out == nullptr)
        return fail(EMIRWV_E_VALUE, "null argument");
    if (nframes < 0) return fail(EMIRWV_E_VALUE, "negative frame count");
    if (flags & ~(uint32_t)EMIRWV_HOST_OUT_ZEROED) return fail(EMIRWV_E_VALUE, "unknown flag");
    if (nframes == 0) return EMIRWV_OK;
    return run_host_impl(pl, h_in, pl->dtype, false, h_out, nframes, h_jminmax,
                         (flags & EMIRWV_HOST_OUT_ZEROED) != 0);
}

int emirwv_plan_set_calibration(emirwv_plan *pl, const uint8_t *h_bpm, const void *h_bias,
                                const void *h_dark, const void *h_flat, int cal_dtype) {
    if (pl == nullptr) return fail(EMIRWV_E_VALUE, "null plan");
    if (cal_dtype != EMIRWV_F32 && cal_dtype != EMIRWV_F64)
        return fail(EMIRWV_E_VALUE, "unknown calibration dtype");
    DeviceGuard guard(pl->device);
    std::lock_guard<std::mutex> lock(pl->mu);
    const size_t npix = (size_t)pl->desc.naxis1 * pl->desc.naxis2;
    const size_t csize = cal_dtype == EMIRWV_F64 ? 8 : 4;
    auto put = [&](void **d_ptr, const void *h_ptr, size_t nbytes) -> int {
        if (*d_ptr) cudaFree(*d_ptr);
        *d_ptr = nullptr;
        if (h_ptr == nullptr) return EMIRWV_OK;
        if (cudaMalloc(d_ptr, nbytes) != cudaSuccess) {
            cudaGetLastError();
            return fail(EMIRWV_E_NOMEM, "out of device memory");
        }
        CUDA_TRY(cudaMemcpy(*d_ptr, h_ptr, nbytes, cudaMemcpyHostToDevice));
        return EMIRWV_OK;
    };
    for (auto &sl : pl->slots) cudaStreamSynchronize(sl.stream);
    int rc = put(reinterpret_cast<void **>(&pl->d_bpm), h_bpm, npix);
    if (rc == EMIRWV_OK) rc = put(&pl->d_bias, h_bias, npix * csize);
    if (rc == EMIRWV_OK) rc = put(&pl->d_dark, h_dark, npix * csize);
    if (rc == EMIRWV_OK) rc = put(&pl->d_flat, h_flat, npix * csize);
    pl->cal_dtype = cal_dtype;
    return rc;
}

int emirwv_run_host_raw(emirwv_plan *pl, const void *h_raw, int raw_dtype, void *h_out,
                        int nframes, int32_t *h_jminmax, uint32_t flags) {
    if (pl == nullptr || h_raw == nullptr || h_out == nullptr)
        return fail(EMIRWV_E_VALUE, "null argument");
    if (nframes < 0) return fail(EMIRWV_E_VALUE, "negative frame count");
    if (flags & ~(uint32_t)EMIRWV_HOST_OUT_ZEROED) return fail(EMIRWV_E_VALUE, "unknown flag");
    const int rc = check_raw(pl, raw_dtype, true);
    if (rc) return rc;
    if (nframes == 0) return EMIRWV_OK;
    return run_host_impl(pl, h_raw, raw_dtype, true, h_out, nframes, h_jminmax,
                         (flags & EMIRWV_HOST_OUT_ZEROED) != 0);
}

int emirwv_abba_host(emirwv_plan *pl, const void *h_in, int in_dtype, int nframes,
                     const int8_t *h_labels, const double *h_offset, int method, double low,
                     double high, float *h_out, uint32_t flags) {
    if (pl == nullptr || h_in == nullptr || h_labels == nullptr || h_out == nullptr)
        return fail(EMIRWV_E_VALUE, "null argument");
    if (nframes < 1 || nframes > 32768) return fail(EMIRWV_E_VALUE, "size out of range");
    if (flags & ~(uint32_t)EMIRWV_HOST_RAW) return fail(EMIRWV_E_VALUE, "unknown flag");
    if (method != EMIRWV_MEAN && method != EMIRWV_MEDIAN && method != EMIRWV_SIGMACLIP)
        return fail(EMIRWV_E_VALUE, "Unexpected method");
    const bool raw = (flags & EMIRWV_HOST_RAW) != 0;
    int rc = check_raw(pl, in_dtype, raw);
    if (rc) return rc;
    std::vector<int32_t> ia, ib;
    for (int f = 0; f < nframes; ++f) {
        if (h_labels[f] > 0) ia.push_back(f);
        else if (h_labels[f] < 0) ib.push_back(f);
    }
    if (ia.empty() || ib.empty()) return fail(EMIRWV_E_VALUE, "both A and B images are needed");
    const bool stacked = method != EMIRWV_MEAN;
    if (stacked && (ia.size() > (size_t)STACK_MAX || ib.size() > (size_t)STACK_MAX))
        return fail(EMIRWV_E_LIMIT, "median / sigmaclip of more than %d images per position", STACK_MAX);
    // the recipe passes yoffset = -offset and leaves frames with a zero offset alone (abba.py:554-556)
    std::vector<double> yoff(nframes, 0.0);
    bool any_shift = false;
    if (h_offset != nullptr)
        for (int f = 0; f < nframes; ++f) {
            if (!std::isfinite(h_offset[f])) return fail(EMIRWV_E_VALUE, "offset of frame %d is not finite", f);
            yoff[f] = h_offset[f] != 0.0 ? -h_offset[f] : 0.0;
            any_shift |= h_offset[f] != 0.0;
        }

    DeviceGuard guard(pl->device);
    std::lock_guard<std::mutex> lock(pl->mu);
    nvtx_range r_all("emirwv_abba_host");
    const int chunk = std::max(1, std::min(nframes, env_int("EMIRWV_HOST_CHUNK", 8)));
    rc = ensure_workspace(pl, chunk);
    if (rc == EMIRWV_OK && raw) rc = ensure_raw(pl, chunk, in_dtype);
    if (rc) return rc;
    const size_t npix_out = (size_t)pl->nsl * pl->rps * pl->nout;
    const size_t in_b = (size_t)pl->desc.naxis1 * pl->desc.naxis2 * (raw ? dtype_size(in_dtype) : pl->esize);
    // the shifted frames are float32 like the recipe's `.astype("float32")`
    const size_t shift_esize = 4;
    const int comb_dtype = any_shift ? EMIRWV_F32 : pl->dtype;
    const size_t comb_esize = any_shift ? shift_esize : pl->esize;
    if (pl->acc == nullptr) CUDA_TRY(cudaStreamCreateWithFlags(&pl->acc, cudaStreamNonBlocking));
    if (pl->d_image == nullptr) CUDA_TRY(cudaMalloc(&pl->d_image, npix_out * sizeof(float)));
    if (!stacked && pl->d_sums == nullptr) CUDA_TRY(cudaMalloc(&pl->d_sums, 2 * npix_out * sizeof(double)));
    if (stacked && pl->stack_bytes < (size_t)nframes * npix_out * comb_esize) {
        cudaFree(pl->d_stack);
        pl->d_stack = nullptr;
        pl->stack_bytes = 0;
        if (cudaMalloc(&pl->d_stack, (size_t)nframes * npix_out * comb_esize) != cudaSuccess) {
            cudaGetLastError();
            return fail(EMIRWV_E_NOMEM, "out of device memory for the stack of %d products", nframes);
        }
        pl->stack_bytes = (size_t)nframes * npix_out * comb_esize;
    }
    if (any_shift && !stacked && pl->shift_chunk < chunk) {
        cudaFree(pl->d_shift);
        pl->d_shift = nullptr;
        CUDA_TRY(cudaMalloc(&pl->d_shift, (size_t)chunk * npix_out * shift_esize));
        pl->shift_chunk = chunk;
    }
    int8_t *d_labels = nullptr;
    CUDA_TRY(cudaMallocAsync(&d_labels, (size_t)nframes, pl->acc));

    // Slot streams: host -> device, flow, hot kernel.  Consumer stream `acc`, chunk after chunk
    // in frame order: the offsets of the chunk's frames, then the float64 sums (mean) -- so the
    // sums run in frame order whatever the slots do -- or nothing (stack: the frames are
    // rectified or shifted straight into their place).
    // (an error must not return while streams still read the caller's buffers: the steps are
    // lambdas, the streams are drained below whatever happened)
    auto step = [&](int f0, int nchunk, emirwv_plan::Slot &sl) -> int {
        const int n = std::min(chunk, nframes - f0);
        if (nchunk >= HOST_SLOTS) CUDA_TRY(cudaStreamWaitEvent(sl.stream, sl.ev_free, 0));
        char *place = stacked ? static_cast<char *>(pl->d_stack) + (size_t)f0 * npix_out * comb_esize : nullptr;
        void *d_dst = (stacked && !any_shift) ? static_cast<void *>(place) : sl.d_out;
        int rc2 = enqueue_chunk(pl, sl, static_cast<const char *>(h_in) + (size_t)f0 * in_b, in_dtype,
                                raw, n, d_dst, nullptr);
        if (rc2) return rc2;
        CUDA_TRY(cudaEventRecord(sl.ev_ready, sl.stream));
        CUDA_TRY(cudaStreamWaitEvent(pl->acc, sl.ev_ready, 0));
        const void *d_comb = d_dst;
        if (any_shift) {
            void *d_sh = stacked ? static_cast<void *>(place) : pl->d_shift;
            rc2 = emirwv_shift_rows(sl.d_out, pl->dtype, d_sh, EMIRWV_F32, n, pl->nsl * pl->rps,
                                    pl->nout, yoff.data() + f0, pl->acc);
            if (rc2) return rc2;
            d_comb = d_sh;
        }
        if (!stacked) {
            rc2 = emirwv_abba_partial(d_comb, d_labels + f0, n, (int64_t)npix_out, comb_dtype,
                                      pl->d_sums, pl->d_sums + npix_out, f0 > 0 ? 1 : 0, pl->acc);
            if (rc2) return rc2;
        }
        CUDA_TRY(cudaEventRecord(sl.ev_free, pl->acc));
        return EMIRWV_OK;
    };
    {
        cudaError_t e0 = cudaMemcpyAsync(d_labels, h_labels, (size_t)nframes, cudaMemcpyHostToDevice, pl->acc);
        if (e0 != cudaSuccess) rc = fail(EMIRWV_E_CUDA, "copy failed: %s", cudaGetErrorString(e0));
    }
    int slot = 0, nchunk = 0;
    for (int f0 = 0; f0 < nframes && rc == EMIRWV_OK; f0 += chunk, slot = (slot + 1) % HOST_SLOTS, ++nchunk)
        rc = step(f0, nchunk, pl->slots[slot]);
    if (rc == EMIRWV_OK) {
        if (method == EMIRWV_MEAN)
            rc = emirwv_abba_finish(pl->d_sums, pl->d_sums + npix_out, (int)ia.size(), (int)ib.size(),
                                    (int64_t)npix_out, pl->d_image, pl->acc);
        else if (method == EMIRWV_MEDIAN)
            rc = emirwv_abba_median(pl->d_stack, nframes, ia.data(), (int)ia.size(), ib.data(),
                                    (int)ib.size(), (int64_t)npix_out, comb_dtype, pl->d_image, pl->acc);
        else
            rc = emirwv_abba_sigmaclip(pl->d_stack, nframes, ia.data(), (int)ia.size(), ib.data(),
                                       (int)ib.size(), (int64_t)npix_out, comb_dtype, low, high,
                                       pl->d_image, pl->acc);
    }
    if (rc == EMIRWV_OK) {
        cudaError_t err = cudaMemcpyAsync(h_out, pl->d_image, npix_out * sizeof(float),
                                          cudaMemcpyDeviceToHost, pl->acc);
        if (err != cudaSuccess) rc = fail(EMIRWV_E_CUDA, "copy failed: %s", cudaGetErrorString(err));
    }
    cudaFreeAsync(d_labels, pl->acc);
    const std::string keep = g_error;
    for (auto &sl : pl->slots) {
        cudaError_t err = cudaStreamSynchronize(sl.stream);
        if (err != cudaSuccess && rc == EMIRWV_OK)
            rc = fail(EMIRWV_E_CUDA, "stream failed: %s", cudaGetErrorString(err));
    }
    cudaError_t err = cudaStreamSynchronize(pl->acc);
    if (err != cudaSuccess && rc == EMIRWV_OK)
        rc = fail(EMIRWV_E_CUDA, "stream failed: %s", cudaGetErrorString(err));
    if (rc != EMIRWV_OK && !keep.empty()) g_error = keep;
    return rc;
}

void *emirwv_host_alloc(int64_t nbytes) {
    void *ptr = nullptr;
    if (nbytes <= 0) return nullptr;
    if (cudaHostAlloc(&ptr, (size_t)nbytes, cudaHostAllocDefault) != cudaSuccess) {
        fail(EMIRWV_E_NOMEM, "cudaHostAlloc(%lld) failed", (long long)nbytes);
        cudaGetLastError();
        return nullptr;
    }
    return ptr;
}

void emirwv_host_free(void *ptr) {
    if (ptr != nullptr) cudaFreeHost(ptr);
}

int emirwv_rectify_slitlet(const emirwv_plan *pl, const void *d_frame, int islitlet,
                           void *d_rect, void *cuda_stream) {
    if (pl == nullptr || d_frame == nullptr || d_rect == nullptr)
        return fail(EMIRWV_E_VALUE, "null argument");
    if (islitlet < 1 || islitlet > pl->nsl || !pl->h_slit[islitlet - 1].valid)
        return fail(EMIRWV_E_VALUE, "slitlet %d is not available", islitlet);
    DeviceGuard guard(pl->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int s = islitlet - 1;
    const int bbw = pl->h_slit[s].bbw;
    const dim3 grid((bbw + 127) / 128, NROWS);
    if (pl->dtype == EMIRWV_F32)
        k_rectify_only<float><<<grid, 128, 0, st>>>(
            static_cast<const float *>(d_frame), static_cast<float *>(d_rect), pl->d_base,
            static_cast<const float *>(pl->d_w), s, pl->K, bbw, pl->desc.naxis1,
            pl->desc.naxis2);
    else
        k_rectify_only<double><<<grid, 128, 0, st>>>(
            static_cast<const double *>(d_frame), static_cast<double *>(d_rect), pl->d_base,
            static_cast<const double *>(pl->d_w), s, pl->K, bbw, pl->desc.naxis1,
            pl->desc.naxis2);
    CUDA_TRY(cudaGetLastError());
    return EMIRWV_OK;
}

int emirwv_rectify2d_host(const double *h_image, int naxis2, int naxis1, const double *aij,
                          const double *bij, int ncoef, int resampling, int max_order,
                          double *h_out) {
    if (h_image == nullptr || aij == nullptr || bij == nullptr || h_out == nullptr)
        return fail(EMIRWV_E_VALUE, "null argument");
    if (naxis1 < 1 || naxis2 < 1) return fail(EMIRWV_E_VALUE, "empty image");
    if (resampling < 1 || resampling > 3)
        return fail(EMIRWV_E_VALUE, "Sorry, resampling method must be 1 or 2");
    const int cap = max_order > 0 ? max_order : 4;
    SlitDev sd;
    memset(&sd, 0, sizeof(sd));
    sd.order = order_from_ncoef(ncoef);
    if (sd.order < 0 || sd.order > cap) return fail(EMIRWV_E_VALUE, "order > %d not implemented", cap);
    sd.valid = 1;
    memcpy(sd.a, aij, sizeof(double) * ncoef);
    memcpy(sd.b, bij, sizeof(double) * ncoef);
    const size_t nbytes = (size_t)naxis1 * naxis2 * sizeof(double);
    double *d_in = nullptr, *d_out = nullptr;
    cudaError_t err = cudaMalloc(&d_in, nbytes);
    if (err == cudaSuccess) err = cudaMalloc(&d_out, nbytes);
    if (err == cudaSuccess) err = cudaMemcpy(d_in, h_image, nbytes, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) {
        const dim3 grid((naxis1 + 127) / 128, naxis2);
        k_rectify2d<<<grid, 128>>>(d_in, naxis2, naxis1, sd, resampling, d_out, naxis2, naxis1);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaMemcpy(h_out, d_out, nbytes, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (err != cudaSuccess)
        return fail(EMIRWV_E_CUDA, "rectify2d failed: %s", cudaGetErrorString(err));
    return EMIRWV_OK;
}

int emirwv_median_slitlets(const void *d_in, void *d_out, int nframes, int nslitlets,
                           int rows_per_slitlet, int naxis1, int mode, int dtype,
                           void *cuda_stream) {
    if (d_in == nullptr || d_out == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (mode != 0 && mode != 1) return fail(EMIRWV_E_VALUE, "mode must be 0 or 1");
    if (dtype != EMIRWV_F32 && dtype != EMIRWV_F64) return fail(EMIRWV_E_VALUE, "unknown dtype");
    if (rows_per_slitlet != NROWS - 1)
        return fail(EMIRWV_E_VALUE, "rows_per_slitlet must be %d", NROWS - 1);
    if (nframes < 0 || nslitlets < 1 || nslitlets > 65535 || naxis1 < 1 || nframes > 65535)
        return fail(EMIRWV_E_VALUE, "size out of range");
    if (nframes == 0) return EMIRWV_OK;
    DeviceGuard guard(device_of(d_in));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const dim3 grid((naxis1 + 127) / 128, nslitlets, nframes);
    if (dtype == EMIRWV_F32)
        k_median_slitlets<float, NROWS - 1><<<grid, 128, 0, st>>>(
            static_cast<const float *>(d_in), static_cast<float *>(d_out), naxis1, nslitlets, mode);
    else
        k_median_slitlets<double, NROWS - 1><<<grid, 128, 0, st>>>(
            static_cast<const double *>(d_in), static_cast<double *>(d_out), naxis1, nslitlets, mode);
    CUDA_TRY(cudaGetLastError());
    return EMIRWV_OK;
}

int emirwv_median_slitlets_host(const void *h_in, void *h_out, int nframes, int nslitlets,
                                int rows_per_slitlet, int naxis1, int mode, int dtype) {
    if (h_in == nullptr || h_out == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (mode != 0 && mode != 1) return fail(EMIRWV_E_VALUE, "mode must be 0 or 1");
    if (dtype != EMIRWV_F32 && dtype != EMIRWV_F64) return fail(EMIRWV_E_VALUE, "unknown dtype");
    if (nframes < 0 || nslitlets < 1 || naxis1 < 1 || rows_per_slitlet < 1)
        return fail(EMIRWV_E_VALUE, "size out of range");
    const size_t esize = dtype == EMIRWV_F32 ? 4 : 8;
    const size_t nin = (size_t)nframes * nslitlets * rows_per_slitlet * naxis1 * esize;
    const size_t nout = mode == 0 ? nin : (size_t)nframes * nslitlets * naxis1 * esize;
    if (nframes == 0) return EMIRWV_OK;
    void *d_in = nullptr, *d_out = nullptr;
    cudaError_t err = cudaMalloc(&d_in, nin);
    if (err == cudaSuccess) err = cudaMalloc(&d_out, nout);
    if (err == cudaSuccess) err = cudaMemcpy(d_in, h_in, nin, cudaMemcpyHostToDevice);
    int rc = EMIRWV_OK;
    if (err == cudaSuccess)
        rc = emirwv_median_slitlets(d_in, d_out, nframes, nslitlets, rows_per_slitlet, naxis1,
                                    mode, dtype, nullptr);
    if (err == cudaSuccess && rc == EMIRWV_OK)
        err = cudaMemcpy(h_out, d_out, nout, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (rc != EMIRWV_OK) return rc;
    if (err != cudaSuccess)
        return fail(EMIRWV_E_CUDA, "median_slitlets failed: %s", cudaGetErrorString(err));
    return EMIRWV_OK;
}

int emirwv_abba_partial(const void *d_frames, const int8_t *d_labels, int nframes, int64_t npix,
                        int dtype, double *d_sum_a, double *d_sum_b, int accumulate,
                        void *cuda_stream) {
    if (d_sum_a == nullptr || d_sum_b == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (nframes < 0 || nframes > 32768 || npix < 1) return fail(EMIRWV_E_VALUE, "size out of range");
    if (nframes > 0 && (d_frames == nullptr || d_labels == nullptr))
        return fail(EMIRWV_E_VALUE, "null argument");
    if (dtype != EMIRWV_F32 && dtype != EMIRWV_F64) return fail(EMIRWV_E_VALUE, "unknown dtype");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int dev = device_of(d_sum_a);
    DeviceGuard guard(dev);
    const int sms = sm_count(dev);
    const int grid = (int)std::min<int64_t>((npix + 255) / 256, (int64_t)sms * 32);
    if (dtype == EMIRWV_F32)
        k_abba_partial<float><<<grid, 256, (size_t)nframes, st>>>(static_cast<const float *>(d_frames), d_labels,
                                                    nframes, (size_t)npix, d_sum_a, d_sum_b,
                                                    accumulate);
    else
        k_abba_partial<double><<<grid, 256, (size_t)nframes, st>>>(static_cast<const double *>(d_frames), d_labels,
                                                     nframes, (size_t)npix, d_sum_a, d_sum_b,
                                                     accumulate);
    CUDA_TRY(cudaGetLastError());
    return EMIRWV_OK;
}

int emirwv_prereduce(const void *d_in, int in_dtype, float *d_out, int nframes, int64_t npix,
                     const void *d_bias, const void *d_dark, const void *d_flat, int cal_dtype,
                     void *cuda_stream) {
    return emirwv_prereduce_ex(d_in, in_dtype, d_out, nframes, 1, npix > INT_MAX ? -1 : (int)npix,
                               nullptr, d_bias, d_dark, d_flat, cal_dtype, cuda_stream);
}

int emirwv_prereduce_ex(const void *d_in, int in_dtype, float *d_out, int nframes, int naxis2,
                        int naxis1, const uint8_t *d_bpm, const void *d_bias, const void *d_dark,
                        const void *d_flat, int cal_dtype, void *cuda_stream) {
    if (d_in == nullptr || d_out == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (nframes < 0 || naxis1 < 1 || naxis2 < 1 || (d_bpm != nullptr && nframes > 65535))
        return fail(EMIRWV_E_VALUE, "size out of range");
    if (in_dtype != EMIRWV_F32 && in_dtype != EMIRWV_F64 && in_dtype != EMIRWV_U16)
        return fail(EMIRWV_E_VALUE, "unknown image dtype");
    if (cal_dtype != EMIRWV_F32 && cal_dtype != EMIRWV_F64)
        return fail(EMIRWV_E_VALUE, "unknown calibration dtype");
    if (d_bpm != nullptr && static_cast<const void *>(d_in) == static_cast<const void *>(d_out))
        return fail(EMIRWV_E_VALUE, "the bad-pixel stage reads neighbours: d_out must not alias d_in");
    if (nframes == 0) return EMIRWV_OK;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int dev = device_of(d_in);
    DeviceGuard guard(dev);
    const int grid = sm_count(dev);  // the launcher sizes the grid (scalar or 4-pixel threads)
    const size_t np = (size_t)naxis1 * naxis2;
    const BpmArgs bpm{d_bpm, naxis1, naxis2};
    const bool c64 = cal_dtype == EMIRWV_F64;
    if (in_dtype == EMIRWV_F32) {
        if (c64) launch_prereduce<float, double>(d_in, d_out, nframes, np, d_bias, d_dark, d_flat, grid, st, bpm);
        else launch_prereduce<float, float>(d_in, d_out, nframes, np, d_bias, d_dark, d_flat, grid, st, bpm);
    } else if (in_dtype == EMIRWV_F64) {
        if (c64) launch_prereduce<double, double>(d_in, d_out, nframes, np, d_bias, d_dark, d_flat, grid, st, bpm);
        else launch_prereduce<double, float>(d_in, d_out, nframes, np, d_bias, d_dark, d_flat, grid, st, bpm);
    } else {
        if (c64) launch_prereduce<uint16_t, double>(d_in, d_out, nframes, np, d_bias, d_dark, d_flat, grid, st, bpm);
        else launch_prereduce<uint16_t, float>(d_in, d_out, nframes, np, d_bias, d_dark, d_flat, grid, st, bpm);
    }
    CUDA_TRY(cudaGetLastError());
    return EMIRWV_OK;
}

int emirwv_stack_sigmaclip(const void *d_frames, int nframes, const int32_t *h_index, int nsel,
                           int64_t npix, int dtype, double low, double high, double *d_out,
                           void *cuda_stream) {
    if (d_frames == nullptr || d_out == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (npix < 1 || nframes < 1) return fail(EMIRWV_E_VALUE, "size out of range");
    if (dtype != EMIRWV_F32 && dtype != EMIRWV_F64) return fail(EMIRWV_E_VALUE, "unknown dtype");
    if (!(low >= 0.0) || !(high >= 0.0)) return fail(EMIRWV_E_VALUE, "low and high must be >= 0");
    StackSel sel;
    const int rc = fill_stack_sel(sel, h_index, nsel, nframes);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int dev = device_of(d_frames);
    DeviceGuard guard(dev);
    const int sms = sm_count(dev);
    if (dtype == EMIRWV_F32)
        launch_sigmaclip(static_cast<const float *>(d_frames), sel, nullptr, (size_t)npix, low, high, d_out, sms, st);
    else
        launch_sigmaclip(static_cast<const double *>(d_frames), sel, nullptr, (size_t)npix, low, high, d_out, sms, st);
    CUDA_TRY(cudaGetLastError());
    return EMIRWV_OK;
}

int emirwv_abba_sigmaclip(const void *d_frames, int nframes, const int32_t *h_index_a, int na,
                          const int32_t *h_index_b, int nb, int64_t npix, int dtype, double low,
                          double high, float *d_out, void *cuda_stream) {
    if (d_frames == nullptr || d_out == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (npix < 1 || nframes < 1) return fail(EMIRWV_E_VALUE, "size out of range");
    if (dtype != EMIRWV_F32 && dtype != EMIRWV_F64) return fail(EMIRWV_E_VALUE, "unknown dtype");
    if (na < 1 || nb < 1) return fail(EMIRWV_E_VALUE, "both A and B images are needed");
    if (!(low >= 0.0) || !(high >= 0.0)) return fail(EMIRWV_E_VALUE, "low and high must be >= 0");
    StackSel sa, sb;
    int rc = fill_stack_sel(sa, h_index_a, na, nframes);
    if (rc) return rc;
    rc = fill_stack_sel(sb, h_index_b, nb, nframes);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int dev = device_of(d_frames);
    DeviceGuard guard(dev);
    const int sms = sm_count(dev);
    if (dtype == EMIRWV_F32)
        launch_sigmaclip(static_cast<const float *>(d_frames), sa, &sb, (size_t)npix, low, high, d_out, sms, st);
    else
        launch_sigmaclip(static_cast<const double *>(d_frames), sa, &sb, (size_t)npix, low, high, d_out, sms, st);
    CUDA_TRY(cudaGetLastError());
    return EMIRWV_OK;
}

int emirwv_shift_rows(const void *d_in, int in_dtype, void *d_out, int out_dtype, int nframes,
                      int naxis2, int naxis1, const double *h_yoffset, void *cuda_stream) {
    if (d_in == nullptr || d_out == nullptr || h_yoffset == nullptr)
        return fail(EMIRWV_E_VALUE, "null argument");
    if (d_in == d_out) return fail(EMIRWV_E_VALUE, "d_out must not alias d_in");
    if (nframes < 0 || nframes > 65535 || naxis1 < 1 || naxis2 < 1)
        return fail(EMIRWV_E_VALUE, "size out of range");
    if ((in_dtype != EMIRWV_F32 && in_dtype != EMIRWV_F64) ||
        (out_dtype != EMIRWV_F32 && out_dtype != EMIRWV_F64))
        return fail(EMIRWV_E_VALUE, "unknown dtype");
    for (int f = 0; f < nframes; ++f)
        if (!std::isfinite(h_yoffset[f]) || std::fabs(h_yoffset[f]) > 1.0e6)
            return fail(EMIRWV_E_VALUE, "offset of frame %d is not a usable number", f);
    if (nframes == 0) return EMIRWV_OK;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    DeviceGuard guard(device_of(d_in));
    // the offsets ride along in stream order (allocated, filled and freed on the stream)
    double *d_off = nullptr;
    CUDA_TRY(cudaMallocAsync(&d_off, sizeof(double) * nframes, st));
    cudaError_t err = cudaMemcpyAsync(d_off, h_yoffset, sizeof(double) * nframes,
                                      cudaMemcpyHostToDevice, st);
    if (err == cudaSuccess) {
        const dim3 grid((naxis1 + 255) / 256, (naxis2 + SHIFT_STRIP - 1) / SHIFT_STRIP, nframes);
        if (grid.y > 65535u) {
            cudaFreeAsync(d_off, st);
            return fail(EMIRWV_E_VALUE, "size out of range");
        }
#define SHIFT(TI, TO)                                                                          \
    k_shift_rows<TI, TO><<<grid, 256, 0, st>>>(static_cast<const TI *>(d_in),                  \
                                               static_cast<TO *>(d_out), naxis2, naxis1, d_off)
        if (in_dtype == EMIRWV_F32 && out_dtype == EMIRWV_F32) SHIFT(float, float);
        else if (in_dtype == EMIRWV_F32) SHIFT(float, double);
        else if (out_dtype == EMIRWV_F32) SHIFT(double, float);
        else SHIFT(double, double);
#undef SHIFT
        err = cudaGetLastError();
    }
    cudaFreeAsync(d_off, st);
    if (err != cudaSuccess)
        return fail(EMIRWV_E_CUDA, "shift_rows failed: %s", cudaGetErrorString(err));
    return EMIRWV_OK;
}

int emirwv_stack_median(const void *d_frames, int nframes, const int32_t *h_index, int nsel,
                        int64_t npix, int dtype, double *d_out, void *cuda_stream) {
    if (d_frames == nullptr || d_out == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (npix < 1 || nframes < 1) return fail(EMIRWV_E_VALUE, "size out of range");
    if (dtype != EMIRWV_F32 && dtype != EMIRWV_F64) return fail(EMIRWV_E_VALUE, "unknown dtype");
    StackSel sel;
    const int rc = fill_stack_sel(sel, h_index, nsel, nframes);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int dev = device_of(d_frames);
    DeviceGuard guard(dev);
    const int sms = sm_count(dev);
    if (dtype == EMIRWV_F32)
        launch_stack_median(static_cast<const float *>(d_frames), sel, (size_t)npix, d_out, sms, st);
    else
        launch_stack_median(static_cast<const double *>(d_frames), sel, (size_t)npix, d_out, sms, st);
    CUDA_TRY(cudaGetLastError());
    return EMIRWV_OK;
}

int emirwv_abba_median(const void *d_frames, int nframes, const int32_t *h_index_a, int na,
                       const int32_t *h_index_b, int nb, int64_t npix, int dtype, float *d_out,
                       void *cuda_stream) {
    if (d_frames == nullptr || d_out == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (npix < 1 || nframes < 1) return fail(EMIRWV_E_VALUE, "size out of range");
    if (dtype != EMIRWV_F32 && dtype != EMIRWV_F64) return fail(EMIRWV_E_VALUE, "unknown dtype");
    if (na < 1 || nb < 1) return fail(EMIRWV_E_VALUE, "both A and B images are needed");
    StackSel sa, sb;
    int rc = fill_stack_sel(sa, h_index_a, na, nframes);
    if (rc) return rc;
    rc = fill_stack_sel(sb, h_index_b, nb, nframes);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int dev = device_of(d_frames);
    DeviceGuard guard(dev);
    const int sms = sm_count(dev);
    if (dtype == EMIRWV_F32)
        launch_abba_median(static_cast<const float *>(d_frames), sa, sb, (size_t)npix, d_out, sms, st);
    else
        launch_abba_median(static_cast<const double *>(d_frames), sa, sb, (size_t)npix, d_out, sms, st);
    CUDA_TRY(cudaGetLastError());
    return EMIRWV_OK;
}

int emirwv_abba_finish(const double *d_sum_a, const double *d_sum_b, int count_a, int count_b,
                       int64_t npix, float *d_out, void *cuda_stream) {
    if (d_sum_a == nullptr || d_sum_b == nullptr || d_out == nullptr)
        return fail(EMIRWV_E_VALUE, "null argument");
    if (count_a < 1 || count_b < 1) return fail(EMIRWV_E_VALUE, "both A and B images are needed");
    if (npix < 1) return fail(EMIRWV_E_VALUE, "size out of range");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int dev = device_of(d_sum_a);
    DeviceGuard guard(dev);
    const int sms = sm_count(dev);
    const int grid = (int)std::min<int64_t>((npix + 255) / 256, (int64_t)sms * 16);
    k_abba_finish<<<grid, 256, 0, st>>>(d_sum_a, d_sum_b, (double)count_a, (double)count_b,
                                        (size_t)npix, d_out);
    CUDA_TRY(cudaGetLastError());
    return EMIRWV_OK;
}

// ------------------------------------------------------------------ peer-memory exchange
}  // extern "C"

// Workspace of the fused ABBA mean combination: ONE device allocation per rank, opened by every
// other rank through CUDA IPC.  Layout (the same on every rank, from npix and world):
//   recv  [2 parities][world sources][2 (A, B)][slab]  float64
//   image [npix] float32   (the result; the destination rank's copy is the one that is written)
//   flags [4][PEER_MAX] int32: pushed (parity 0, 1), slab finished (parity 0, 1);  [1] time-out
struct emirwv_peer {
    int device = 0, rank = 0, world = 1;
    size_t npix = 0, slab = 0;
    size_t off_image = 0, off_flags = 0, nbytes = 0;
    char *local = nullptr;
    char *peer[PEER_MAX] = {};
    int epoch = 0;
    bool connected = false;
};

extern "C" {

int emirwv_peer_create(int64_t npix, int rank, int world, emirwv_peer **out, void *h_handle) {
    if (out == nullptr || h_handle == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    *out = nullptr;
    if (npix < 1 || world < 1 || world > PEER_MAX || rank < 0 || rank >= world)
        return fail(EMIRWV_E_VALUE, "invalid rank / world size (at most %d ranks)", PEER_MAX);
    emirwv_peer *p = new (std::nothrow) emirwv_peer();
    if (p == nullptr) return fail(EMIRWV_E_NOMEM, "out of host memory");
    CUDA_TRY(cudaGetDevice(&p->device));
    p->rank = rank;
    p->world = world;
    p->npix = (size_t)npix;
    p->slab = ((size_t)npix + world - 1) / world;
    const size_t recv_b = (size_t)2 * world * 2 * p->slab * sizeof(double);
    p->off_image = (recv_b + 255) & ~(size_t)255;
    p->off_flags = (p->off_image + (size_t)npix * sizeof(float) + 255) & ~(size_t)255;
    p->nbytes = p->off_flags + (4 * PEER_MAX + 4) * sizeof(int);
    cudaError_t err = cudaMalloc(&p->local, p->nbytes);
    if (err == cudaSuccess) err = cudaMemset(p->local, 0, p->nbytes);
    cudaIpcMemHandle_t handle;
    if (err == cudaSuccess) err = cudaIpcGetMemHandle(&handle, p->local);
    if (err != cudaSuccess) {
        cudaFree(p->local);
        delete p;
        return fail(EMIRWV_E_CUDA, "peer workspace: %s", cudaGetErrorString(err));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(h_handle, &handle, sizeof(handle));
    *out = p;
    return EMIRWV_OK;
}

int emirwv_peer_connect(emirwv_peer *p, const void *h_handles) {
    if (p == nullptr || h_handles == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    DeviceGuard guard(p->device);
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) {
            p->peer[r] = p->local;
            continue;
        }
        cudaIpcMemHandle_t handle;
        memcpy(&handle, static_cast<const char *>(h_handles) + (size_t)r * 64, 64);
        void *ptr = nullptr;
        cudaError_t err = cudaIpcOpenMemHandle(&ptr, handle, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess)
            return fail(EMIRWV_E_CUDA, "cannot open the workspace of rank %d: %s", r,
                        cudaGetErrorString(err));
        p->peer[r] = static_cast<char *>(ptr);
    }
    p->connected = true;
    return EMIRWV_OK;
}

int emirwv_peer_disconnect(emirwv_peer *p) {
    if (p == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    DeviceGuard guard(p->device);
    CUDA_TRY(cudaDeviceSynchronize());
    for (int r = 0; r < p->world; ++r) {
        if (r != p->rank && p->peer[r] != nullptr) cudaIpcCloseMemHandle(p->peer[r]);
        p->peer[r] = nullptr;
    }
    p->connected = false;
    return EMIRWV_OK;
}

void emirwv_peer_destroy(emirwv_peer *p) {
    if (p == nullptr) return;
    emirwv_peer_disconnect(p);
    DeviceGuard guard(p->device);
    cudaFree(p->local);
    delete p;
}

int emirwv_peer_timed_out(emirwv_peer *p) {
    if (p == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    DeviceGuard guard(p->device);
    int flag = 0;
    CUDA_TRY(cudaMemcpy(&flag, p->local + p->off_flags + 4 * PEER_MAX * sizeof(int), sizeof(int),
                        cudaMemcpyDeviceToHost));
    return flag;
}

int emirwv_abba_mean_push(emirwv_peer *p, const void *d_frames, const int8_t *d_labels,
                          int nframes, int dtype, const double *d_sum_a, const double *d_sum_b,
                          int count_a, int count_b, int root, float **d_image, void *cuda_stream) {
    if (p == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (!p->connected) return fail(EMIRWV_E_VALUE, "the peer workspace is not connected");
    if (nframes < 0 || nframes > 32768) return fail(EMIRWV_E_VALUE, "size out of range");
    if (nframes > 0 && (d_frames == nullptr || d_labels == nullptr))
        return fail(EMIRWV_E_VALUE, "null argument");
    if ((d_sum_a == nullptr) != (d_sum_b == nullptr)) return fail(EMIRWV_E_VALUE, "null argument");
    if (dtype != EMIRWV_F32 && dtype != EMIRWV_F64) return fail(EMIRWV_E_VALUE, "unknown dtype");
    if (count_a < 1 || count_b < 1) return fail(EMIRWV_E_VALUE, "both A and B images are needed");
    if (root < 0 || root >= p->world) return fail(EMIRWV_E_VALUE, "invalid rank / world size");
    DeviceGuard guard(p->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int epoch = ++p->epoch, parity = epoch & 1;
    PeerPtrs peers;
    memset(&peers, 0, sizeof(peers));
    for (int r = 0; r < p->world; ++r) {
        peers.recv[r] = reinterpret_cast<double *>(p->peer[r]);
        peers.flags[r] = reinterpret_cast<int *>(p->peer[r] + p->off_flags);
    }
    int *flags = reinterpret_cast<int *>(p->local + p->off_flags);
    int *timeout = flags + 4 * PEER_MAX;
    const long long max_spins = 4000000;                       // a few seconds, then give up
    const int sms = sm_count(p->device);
    const int grid = (int)std::min<size_t>((p->npix + 255) / 256, (size_t)sms * 32);
    const int acc = d_sum_a != nullptr ? 1 : 0;
    // 1. this rank's sums, pushed to the owners of the pixels
    if (dtype == EMIRWV_F32)
        k_abba_partial_push<float><<<grid, 256, (size_t)nframes, st>>>(
            static_cast<const float *>(d_frames), d_labels, nframes, p->npix, d_sum_a, d_sum_b, acc,
            peers, p->rank, p->world, p->slab, parity);
    else
        k_abba_partial_push<double><<<grid, 256, (size_t)nframes, st>>>(
            static_cast<const double *>(d_frames), d_labels, nframes, p->npix, d_sum_a, d_sum_b, acc,
            peers, p->rank, p->world, p->slab, parity);
    // 2. tell every owner; 3. wait for every rank's pushes into this rank's slab
    k_peer_signal<<<1, PEER_MAX, 0, st>>>(peers, p->rank, p->world, parity, epoch, 0, -1);
    k_peer_wait<<<1, PEER_MAX, 0, st>>>(flags + parity * PEER_MAX, p->world, epoch, max_spins, timeout);
    // 4. this rank's slab of the result, stored into the destination rank's image
    const size_t first = (size_t)p->rank * p->slab;
    const double *recv = reinterpret_cast<const double *>(p->local) + (size_t)parity * p->world * 2 * p->slab;
    float *image_root = reinterpret_cast<float *>(p->peer[root] + p->off_image);
    const int fgrid = (int)std::max<size_t>(1, std::min<size_t>((p->slab + 255) / 256, (size_t)sms * 8));
    k_abba_finish_slab<<<fgrid, 256, 0, st>>>(recv, p->world, p->slab, first, p->npix, (double)count_a,
                                              (double)count_b, image_root);
    k_peer_signal<<<1, PEER_MAX, 0, st>>>(peers, p->rank, p->world, parity, epoch, 2 * PEER_MAX, root);
    // 5. the destination waits for the slabs of all ranks
    if (p->rank == root)
        k_peer_wait<<<1, PEER_MAX, 0, st>>>(flags + (2 + parity) * PEER_MAX, p->world, epoch, max_spins,
                                            timeout);
    CUDA_TRY(cudaGetLastError());
    if (d_image != nullptr)
        *d_image = p->rank == root ? reinterpret_cast<float *>(p->local + p->off_image) : nullptr;
    return EMIRWV_OK;
}

// ------------------------------------------------------------------ NCCL exchanges
}  // extern "C"

namespace {

// NCCL is not linked: the library must load on a box without it.  A caller that owns an
// ncclComm_t has NCCL in its process already; the few entry points used here are looked up in
// it (or in libnccl.so.2) on first use.  Enum values are NCCL's stable ABI.
struct NcclApi {
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
constexpr int NCCL_UINT8 = 1, NCCL_FLOAT64 = 8, NCCL_SUM = 0;

const NcclApi &nccl_api() {
    static NcclApi api = []() {
        NcclApi a;
        void *h = dlopen(nullptr, RTLD_NOW);                       // already in the process?
        if (h == nullptr || dlsym(h, "ncclSend") == nullptr) {
            const char *name = getenv("EMIRWV_NCCL_LIB");
            h = dlopen(name && *name ? name : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        }
        if (h == nullptr) return a;
        a.GroupStart = reinterpret_cast<int (*)()>(dlsym(h, "ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<int (*)()>(dlsym(h, "ncclGroupEnd"));
        a.Send = reinterpret_cast<decltype(a.Send)>(dlsym(h, "ncclSend"));
        a.Recv = reinterpret_cast<decltype(a.Recv)>(dlsym(h, "ncclRecv"));
        a.Reduce = reinterpret_cast<decltype(a.Reduce)>(dlsym(h, "ncclReduce"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        a.ok = a.GroupStart && a.GroupEnd && a.Send && a.Recv && a.Reduce;
        return a;
    }();
    return api;
}

int nccl_fail(const NcclApi &api, int rc, const char *what) {
    return fail(EMIRWV_E_CUDA, "%s failed: %s", what,
                api.GetErrorString ? api.GetErrorString(rc) : "NCCL error");
}

}  // namespace

extern "C" {

int emirwv_gather(const void *d_local, const int32_t *counts, int world, int rank,
                  int64_t frame_bytes, void *d_root, void *nccl_comm, int root,
                  void *cuda_stream) {
    if (counts == nullptr || nccl_comm == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (world < 1 || rank < 0 || rank >= world || root < 0 || root >= world || frame_bytes < 1)
        return fail(EMIRWV_E_VALUE, "invalid rank / world size");
    for (int r = 0; r < world; ++r)
        if (counts[r] < 0) return fail(EMIRWV_E_VALUE, "negative frame count");
    if (counts[rank] > 0 && d_local == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (rank == root && d_root == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    const NcclApi &api = nccl_api();
    if (!api.ok) return fail(EMIRWV_E_CUDA, "NCCL is not available in this process (libnccl.so.2)");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    DeviceGuard guard(device_of(rank == root ? d_root : d_local));
    if (rank == root && counts[rank] > 0) {        // the root's own products: a local copy
        size_t off = 0;
        for (int r = 0; r < rank; ++r) off += (size_t)counts[r] * (size_t)frame_bytes;
        char *dst = static_cast<char *>(d_root) + off;
        if (dst != d_local)
            CUDA_TRY(cudaMemcpyAsync(dst, d_local, (size_t)counts[rank] * (size_t)frame_bytes,
                                     cudaMemcpyDeviceToDevice, st));
    }
    int rc = api.GroupStart();
    if (rc) return nccl_fail(api, rc, "ncclGroupStart");
    if (rank == root) {
        size_t off = 0;
        for (int r = 0; r < world && rc == 0; ++r) {
            const size_t nbytes = (size_t)counts[r] * (size_t)frame_bytes;
            char *dst = static_cast<char *>(d_root) + off;
            off += nbytes;
            if (nbytes == 0 || r == rank) continue;
            rc = api.Recv(dst, nbytes, NCCL_UINT8, r, nccl_comm, st);
        }
    } else if (counts[rank] > 0) {
        rc = api.Send(d_local, (size_t)counts[rank] * (size_t)frame_bytes, NCCL_UINT8, root,
                      nccl_comm, st);
    }
    const int rc_end = api.GroupEnd();
    if (rc) return nccl_fail(api, rc, "ncclSend / ncclRecv");
    if (rc_end) return nccl_fail(api, rc_end, "ncclGroupEnd");
    return EMIRWV_OK;
}

int emirwv_abba_reduce(double *d_sum_a, double *d_sum_b, int64_t npix, void *nccl_comm, int root,
                       void *cuda_stream) {
    if (d_sum_a == nullptr || d_sum_b == nullptr || nccl_comm == nullptr)
        return fail(EMIRWV_E_VALUE, "null argument");
    if (npix < 1 || root < 0) return fail(EMIRWV_E_VALUE, "size out of range");
    const NcclApi &api = nccl_api();
    if (!api.ok) return fail(EMIRWV_E_CUDA, "NCCL is not available in this process (libnccl.so.2)");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    DeviceGuard guard(device_of(d_sum_a));
    int rc = api.GroupStart();
    if (rc) return nccl_fail(api, rc, "ncclGroupStart");
    rc = api.Reduce(d_sum_a, d_sum_a, (size_t)npix, NCCL_FLOAT64, NCCL_SUM, root, nccl_comm, st);
    if (rc == 0)
        rc = api.Reduce(d_sum_b, d_sum_b, (size_t)npix, NCCL_FLOAT64, NCCL_SUM, root, nccl_comm, st);
    const int rc_end = api.GroupEnd();
    if (rc) return nccl_fail(api, rc, "ncclReduce");
    if (rc_end) return nccl_fail(api, rc_end, "ncclGroupEnd");
    return EMIRWV_OK;
}

int emirwv_debug_geometry(const emirwv_plan *pl, int islitlet, int32_t *h_row, int32_t *h_col,
                          uint8_t *h_ok, double *h_weights) {
    if (pl == nullptr) return fail(EMIRWV_E_VALUE, "null plan");
    if (islitlet < 1 || islitlet > pl->nsl || !pl->h_slit[islitlet - 1].valid)
        return fail(EMIRWV_E_VALUE, "slitlet %d is not available", islitlet);
    DeviceGuard guard(pl->device);
    const int s = islitlet - 1;
    const SlitDev &sd = pl->h_slit[s];
    const int KK = pl->K * pl->K;
    for (int y = 0; y < NROWS; ++y)
        for (int j = 0; j < sd.bbw; ++j) {
            const int32_t pk = pl->h_base[((size_t)s * NROWS + y) * W + j];
            const int r = base_row(pk) - (sd.ns1 - 1 - pl->desc.offy);
            const int c = base_col(pk) - (sd.nc1 - 1 - pl->desc.offx);
            if (h_row) h_row[(size_t)y * sd.bbw + j] = r;
            if (h_col) h_col[(size_t)y * sd.bbw + j] = c;
            if (h_ok) h_ok[(size_t)y * sd.bbw + j] = (r >= 0 && r < sd.bbh && c >= 0 && c < sd.bbw);
        }
    if (h_weights) {
        const int RS = rec_size(pl->K);
        const size_t n = (size_t)NROWS * W * RS;
        std::vector<unsigned char> raw(n * pl->esize);
        CUDA_TRY(cudaMemcpy(raw.data(),
                            static_cast<const char *>(pl->d_w) + (size_t)s * n * pl->esize,
                            n * pl->esize, cudaMemcpyDeviceToHost));
        for (int y = 0; y < NROWS; ++y)
            for (int tap = 0; tap < KK; ++tap)
                for (int j = 0; j < sd.bbw; ++j) {
                    const size_t rec = ((size_t)y * W + j) * RS;
                    const int K = pl->K;
                    const double v =
                        pl->dtype == EMIRWV_F32
                            ? (double)record_weight(reinterpret_cast<const float *>(raw.data()) + rec,
                                                    K, tap / K, tap % K)
                            : record_weight(reinterpret_cast<const double *>(raw.data()) + rec, K,
                                            tap / K, tap % K);
                    h_weights[((size_t)y * sd.bbw + j) * KK + tap] = v;
                }
    }
    return EMIRWV_OK;
}

int emirwv_debug_borders(const emirwv_plan *pl, int islitlet, int32_t *h_idx, double *h_t) {
    if (pl == nullptr) return fail(EMIRWV_E_VALUE, "null plan");
    if (islitlet < 1 || islitlet > pl->nsl || !pl->h_slit[islitlet - 1].valid)
        return fail(EMIRWV_E_VALUE, "slitlet %d is not available", islitlet);
    const size_t off = (size_t)(islitlet - 1) * pl->NB;
    for (int k = 0; k < pl->NB; ++k) {
        if (h_idx) h_idx[k] = pl->h_idx[off + k];
        if (h_t) h_t[k] = pl->h_t[off + k];
    }
    return EMIRWV_OK;
}

}  // extern "C"
