// emirwv.cu — sm_100a implementation of the apply_rectwv_coeff hot path.
//
// Reference path: emirdrp/processing/wavecal/apply_rectwv_coeff.py:35-287 and the
// numina routines it calls.  Design:
//
//   plan  (once per RectWaveCoeff, float64, on the GPU + a little host code)
//     k_extent / k_build_geometry : 2-D polynomial map of every output pixel
//         (its centre, or its four corners) and the exact overlap areas with the
//         source pixels  ->  frame-independent tables  base[s][y][j], w[s][y][tap][j]
//     build_border_tables (host)  : for every border of the linear wavelength grid
//         the source interval it falls in and the fractional position inside it
//     build_tiles (host)          : work items (slitlet, row block, column block)
//         with the source window each one needs
//
//   run   (per batch of frames; one fused kernel, nothing intermediate touches HBM)
//     k_rectwv<T,K,G>: stage the source window of G frames in shared memory, then
//       every warp owns whole rectified rows of the tile:
//         A: rectified flux = K*K weighted gather            (kernel 1 of the spec)
//         C: Steffen-limited end slopes of every source pixel (kernel 2)
//         D: cumulative-flux difference at the output borders (kernel 2), coalesced
//            store, first/last non-zero channel per slitlet   (kernel 3)
//
// No tensor cores (there is no dense contraction), no CPU fallback.

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "emirwv.h"
#include "geom.cuh"

namespace {

using namespace emirwv;

constexpr int NROWS = EMIRWV_ROWS_SCANNED;   // 39 rectified rows per slitlet
constexpr int W = 2048;                      // EMIR_NAXIS1: pitch of per-column tables
constexpr int MAXK = 4;                      // largest supported footprint
constexpr int MAXTHREADS = 512;
constexpr int HOST_SLOTS = 3;                // pipeline depth of emirwv_run_host
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

struct Tile {
    int slit;           // 0-based slitlet
    int y0, ny;         // rows (of the 39 scanned) handled by this tile
    int k0, nk;         // output columns [k0, k0+nk)
    int jlo, npx;       // rectified columns needed: [jlo, jlo+npx)
    int r0, c0;         // frame coordinates of the staged source window origin
    int wrows, wcols;   // c0 and wcols are multiples of 4 (16-byte vector loads)
    int empty;          // no wavelength coverage: the tile is exactly zero
    int nchan;          // bbw of the slitlet
    int pad[3];
};

struct GeomParams {
    const SlitDev *slit;
    int nsl;
    int naxis1, naxis2;
    int offx, offy;
    int mode, K;
    int32_t *base;
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
    const T *w;
    const Border<T> *border;
    const PixGeom<T> *pix;
    int naxis2;
    int nout, NB;           // naxis1_out, naxis1_out + 1
    int rps, nrows_out;     // 38, 55*38
    int nsl;
    int wpitch, wrows;      // shared window: row pitch and rows per frame
    int npxp;               // scratch pitch (>= max npx)
    int max_span;           // max whole source pixels inside one output pixel
    int vec_ok;             // rows of the output are 16-byte aligned
    int nlive_tiles;        // tiles [0, nlive_tiles) have wavelength coverage
    int ngroups;            // ceil(nframes / G)
    int stage_bytes;        // one pipeline stage in shared memory
    int off_border, off_pix;  // byte offsets of the tables inside a stage
};

__host__ __device__ inline int32_t pack_base(int r, int c) {
    r = max(-30000, min(30000, r));
    c = max(-30000, min(30000, c));
    return (int32_t)(((uint32_t)(r & 0xffff) << 16) | (uint32_t)(c & 0xffff));
}
__host__ __device__ inline int base_row(int32_t p) { return (int)(int16_t)((uint32_t)p >> 16); }
__host__ __device__ inline int base_col(int32_t p) { return (int)(int16_t)((uint32_t)p & 0xffffu); }

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

// ---------------------------------------------------------------- plan kernels
// Largest footprint (in source pixels) of any mapped output pixel, mode 2.
__global__ void k_extent(GeomParams p) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int s = blockIdx.z;
    const SlitDev &sd = p.slit[s];
    int need = 1;
    if (sd.valid && j < sd.bbw) {
        double qu[4], qv[4];
        map_corners(sd, j, sd.min_row + y, qu, qv);
        int lo, hi;
        quad_extent(qu, lo, hi);
        need = hi - lo + 1;
        quad_extent(qv, lo, hi);
        need = max(need, hi - lo + 1);
    }
    need = __reduce_max_sync(0xffffffffu, need);
    if ((threadIdx.x & 31) == 0) atomicMax(p.kmax, need);
}

// base[s][y][j] (frame coordinates of tap (0,0)) and w[s][y][tap][j].
template <typename T>
__global__ void k_build_geometry(GeomParams p, T *w) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int s = blockIdx.z;
    const SlitDev &sd = p.slit[s];
    if (j >= W) return;
    const int K = p.K, KK = K * K;
    const size_t rowtab = (size_t)s * NROWS + y;
    T *wp = w + rowtab * KK * W + j;
    if (!sd.valid || j >= sd.bbw) {
        p.base[rowtab * W + j] = pack_base(0, 0);
        for (int tap = 0; tap < KK; ++tap) wp[(size_t)tap * W] = T(0);
        return;
    }
    const int row = sd.min_row + y;
    double wt[MAXK * MAXK];
    int br, bc;
    if (p.mode == EMIRWV_AREA) {
        double qu[4], qv[4];
        map_corners(sd, j, row, qu, qv);
        int hr, hc;
        quad_extent(qu, bc, hc);
        quad_extent(qv, br, hr);
        // absolute bounding-box coordinates, like the reference's clipping
        for (int a = 0; a < K; ++a)
            for (int b = 0; b < K; ++b) {
                double area = 0.0;
                if (br + a <= hr && bc + b <= hc)
                    area = quad_box_area(qu, qv, (double)(bc + b) - 0.5, (double)(bc + b) + 0.5,
                                         (double)(br + a) - 0.5, (double)(br + a) + 0.5);
                wt[a * K + b] = area;
            }
    } else {
        double u, v;
        fmap(sd.order, sd.a, sd.b, (double)j, (double)row, u, v);
        u = clampd(u);
        v = clampd(v);
        if (p.mode == EMIRWV_NEAREST) {
            bc = (int)rint(u);     // round half to even, like numpy.rint
            br = (int)rint(v);
            wt[0] = 1.0;
        } else {                   // bilinear extension
            const double fu0 = floor(u), fv0 = floor(v);
            bc = (int)fu0;
            br = (int)fv0;
            const double fu = u - fu0, fv = v - fv0;
            wt[0] = (1.0 - fv) * (1.0 - fu);
            wt[1] = (1.0 - fv) * fu;
            wt[2] = fv * (1.0 - fu);
            wt[3] = fv * fu;
        }
    }
    // frame coordinates: bounding-box crop (slitlet2d.py:366-369) and the global
    // integer offsets (apply_rectwv_coeff.py:74-78) folded into the address
    const int fr0 = br + sd.ns1 - 1 - p.offy;
    const int fc0 = bc + sd.nc1 - 1 - p.offx;
    for (int a = 0; a < K; ++a)
        for (int b = 0; b < K; ++b) {
            const int r = br + a, c = bc + b;
            const int fr = fr0 + a, fc = fc0 + b;
            const bool inside = r >= 0 && r < sd.bbh && c >= 0 && c < sd.bbw && fr >= 0 &&
                                fr < p.naxis2 && fc >= 0 && fc < p.naxis1;
            wp[(size_t)(a * K + b) * W] = inside ? (T)wt[a * K + b] : T(0);
        }
    p.base[rowtab * W + j] = pack_base(fr0, fc0);
}

__global__ void k_init_jmm(int32_t *jmm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) jmm[i] = (i & 1) ? -1 : INT_MAX;
}

// Kernel 1 alone, straight from global memory (parity hook / Slitlet2D.rectify).
template <typename T>
__global__ void k_rectify_only(const T *frame, T *rect, const int32_t *base, const T *w,
                               int s, int K, int bbw, int naxis1, int naxis2) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (j >= bbw) return;
    const size_t rowtab = (size_t)s * NROWS + y;
    const int32_t pk = base[rowtab * W + j];
    const int r0 = base_row(pk), c0 = base_col(pk);
    const T *wp = w + rowtab * (K * K) * W + j;
    T acc = T(0);
    for (int a = 0; a < K; ++a)
        for (int b = 0; b < K; ++b) {
            const T wt = wp[(size_t)(a * K + b) * W];
            const int r = min(max(r0 + a, 0), naxis2 - 1);
            const int c = min(max(c0 + b, 0), naxis1 - 1);
            acc = fma(wt, frame[(size_t)r * naxis1 + c], acc);
        }
    rect[(size_t)y * bbw + j] = acc;
}

// ---------------------------------------------------------------- the hot kernel
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <>
struct Vec16<double> {
    using type = double2;
    static constexpr int N = 2;
};

// 16-byte asynchronous global -> shared copy (LDGSTS); zero fill when !valid
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int src_bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src),
                 "r"(src_bytes)
                 : "memory");
}
// 8- or 16-byte asynchronous copy of one table record
template <int BYTES>
__device__ __forceinline__ void cp_async_small(void *smem_dst, const void *gmem_src) {
    static_assert(BYTES == 8 || BYTES == 16, "record size");
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(dst), "l"(gmem_src),
                 "n"(BYTES)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// 8/16/32-byte table records through the read-only path (ld.global.nc)
__device__ __forceinline__ PixGeom<float> ldg_struct(const PixGeom<float> *p) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    return PixGeom<float>{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ PixGeom<double> ldg_struct(const PixGeom<double> *p) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return PixGeom<double>{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ Border<float> ldg_struct(const Border<float> *p) {
    const int2 v = __ldg(reinterpret_cast<const int2 *>(p));
    Border<float> b;
    b.tau = __int_as_float(v.x);
    b.idx = v.y;
    return b;
}
__device__ __forceinline__ Border<double> ldg_struct(const Border<double> *p) {
    const int4 v = __ldg(reinterpret_cast<const int4 *>(p));
    Border<double> b;
    b.tau = __hiloint2double(v.y, v.x);
    b.idx = v.z;
    return b;
}

// Make a loop-invariant value opaque so that the compiler keeps it in a register
// instead of recomputing it from kernel parameters inside the hot loops.
__device__ __forceinline__ void pin(int &x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void pin(unsigned &x) { asm volatile("" : "+r"(x)); }
template <typename P>
__device__ __forceinline__ void pin(P *&x) { asm volatile("" : "+l"(x)); }

// Zero-coverage tiles (outside the wavelength range of a slitlet, or a missing
// slitlet): exact zeros, written with 16-byte stores where the row alignment allows.
template <typename T>
__global__ void __launch_bounds__(256) k_zero_fill(const RunParams<T> p, int first_tile) {
    using V = typename Vec16<T>::type;
    constexpr int VN = Vec16<T>::N;
    const Tile tile = p.tiles[first_tile + blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const size_t frame_out = (size_t)p.nrows_out * p.nout;
    const int ylast = min(tile.y0 + tile.ny, p.rps);
    const bool vec = p.vec_ok && ((tile.k0 | tile.nk) % VN) == 0;
    for (int f = blockIdx.y; f < p.nframes; f += gridDim.y)
        for (int y = tile.y0 + warp; y < ylast; y += nwarps) {
            T *orow = p.out + (size_t)f * frame_out + (size_t)(tile.slit * p.rps + y) * p.nout + tile.k0;
            if (vec) {
                V zero;
                memset(&zero, 0, sizeof(zero));
                for (int k = lane; k < tile.nk / VN; k += 32) __stcs(reinterpret_cast<V *>(orow) + k, zero);
            } else {
                for (int k = lane; k < tile.nk; k += 32) __stcs(orow + k, T(0));
            }
        }
}

// One pipeline stage in shared memory: everything a work item (tile x frame group)
// needs that does not depend on the rectified row.
//   [G source windows][Border table of the tile][PixGeom table of the tile]
// Copied with cp.async while the previous item is being computed.
template <typename T, int G>
__device__ __forceinline__ void issue_stage(const RunParams<T> &p, const Tile &t, int grp,
                                            unsigned char *stage, int wp1, int tid, int nthr) {
    constexpr int VN = Vec16<T>::N;
    T *win = reinterpret_cast<T *>(stage);
    const int wsz = p.wrows * wp1;
    const int vcols = t.wcols / VN;
    const int nvec = t.wrows * vcols;
    const float inv = 1.0f / (float)vcols;
    const size_t frame_in = (size_t)W * p.naxis2;
    const int nlive = min(G, p.nframes - grp * G);
    for (int g = 0; g < nlive; ++g) {
        const T *fr = p.in + (size_t)(grp * G + g) * frame_in;
        T *wg = win + (size_t)g * wsz;
        for (int e = tid; e < nvec; e += nthr) {
            const int rr = (int)(((float)e + 0.5f) * inv);     // e / vcols, exact here
            const int vc = e - rr * vcols;
            const int r = t.r0 + rr, c = t.c0 + vc * VN;
            const bool ok = r >= 0 && r < p.naxis2 && c >= 0 && c < W;
            cp_async_16(wg + rr * wp1 + vc * VN, fr + (ok ? (size_t)r * W + c : 0), ok);
        }
    }
    Border<T> *bs = reinterpret_cast<Border<T> *>(stage + p.off_border);
    const Border<T> *bsrc = p.border + (size_t)t.slit * p.NB + t.k0;
    for (int e = tid; e <= t.nk; e += nthr) cp_async_small<sizeof(Border<T>)>(bs + e, bsrc + e);
    PixGeom<T> *ps = reinterpret_cast<PixGeom<T> *>(stage + p.off_pix);
    const PixGeom<T> *psrc = p.pix + (size_t)t.slit * W + t.jlo;
    for (int e = tid; e < t.npx * (int)(sizeof(PixGeom<T>) / 16); e += nthr)
        cp_async_16(reinterpret_cast<char *>(ps) + e * 16,
                    reinterpret_cast<const char *>(psrc) + e * 16, true);
}

// NPXP_C / WP_C: compile-time scratch pitch and window pitch (0 = take them from the
// plan at run time).  SPANLOOP: an output pixel may contain more than one whole
// source pixel (coarse output grids); the EMIR grids need at most one.
//
// Persistent kernel: CTA b processes work items b, b + gridDim.x, ...; item =
// (live tile, frame group), frame group fastest so that CTAs running at the same
// time share the tile's tables in L2.  Shared memory:
//   [tile descriptor ring: 3][stage 0][stage 1][per warp: G x (Knot[npxp], rect[npxp+4])]
template <typename T, int K, int G, int NPXP_C, int WP_C, bool SPANLOOP>
__global__ void __launch_bounds__(MAXTHREADS) k_rectwv(const RunParams<T> p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int npxp = NPXP_C ? NPXP_C : p.npxp;
    const int wp1 = WP_C ? WP_C : p.wpitch;
    const int wsz = p.wrows * wp1;
    const int rpitch = npxp + 4;
    const size_t frame_out = (size_t)p.nrows_out * p.nout;
    const int nwork = p.nlive_tiles * p.ngroups;

    Tile *ring = reinterpret_cast<Tile *>(smem_raw);                       // [3]
    unsigned char *stage0 = smem_raw + 3 * sizeof(Tile);
    T *scr = reinterpret_cast<T *>(stage0 + 2 * (size_t)p.stage_bytes);
    int woff = warp * (G * (npxp * 4 + rpitch));                // this warp's scratch
    pin(woff);
    Knot<T> *knots = reinterpret_cast<Knot<T> *>(scr + woff);   // [G][npxp]
    T *rect = scr + woff + G * npxp * 4 + 1;                    // [G][rpitch], halo at -1

    int w = blockIdx.x;
    if (w >= nwork) return;

    // prologue: stage of the first item, descriptors of the first two
    {
        const Tile t0 = p.tiles[w / p.ngroups];
        issue_stage<T, G>(p, t0, w % p.ngroups, stage0, wp1, tid, nthr);
        if (tid < 4) cp_async_16(reinterpret_cast<char *>(&ring[0]) + tid * 16,
                                 reinterpret_cast<const char *>(p.tiles + w / p.ngroups) + tid * 16, true);
        const int w1 = w + gridDim.x;
        if (w1 < nwork && tid >= 4 && tid < 8)
            cp_async_16(reinterpret_cast<char *>(&ring[1]) + (tid - 4) * 16,
                        reinterpret_cast<const char *>(p.tiles + w1 / p.ngroups) + (tid - 4) * 16, true);
    }

    unsigned fo = (unsigned)frame_out;                         // < 2^32 elements per frame
    pin(fo);

    for (int it = 0;; ++it) {
        cp_async_wait_all();
        __syncthreads();                     // stage(it), descriptors it and it+1 have landed

        unsigned char *stage = stage0 + (size_t)(it & 1) * p.stage_bytes;
        const int wnext = w + gridDim.x;
        if (wnext < nwork) {
            // prefetch the next item's stage; its descriptor arrived one iteration ago
            const Tile tn = ring[(it + 1) % 3];
            issue_stage<T, G>(p, tn, wnext % p.ngroups,
                              stage0 + (size_t)((it + 1) & 1) * p.stage_bytes, wp1, tid, nthr);
            const int w2 = wnext + gridDim.x;
            if (w2 < nwork && tid < 4)
                cp_async_16(reinterpret_cast<char *>(&ring[(it + 2) % 3]) + tid * 16,
                            reinterpret_cast<const char *>(p.tiles + w2 / p.ngroups) + tid * 16, true);
        }

        // ---------------------------------------------------------------- item `w`
        const Tile tile = ring[it % 3];
        const int grp = w % p.ngroups;
        const int g0 = grp * G;
        const int s = tile.slit;
        const int npx = tile.npx;
        const int nlive = min(G, p.nframes - g0);
        const T *win = reinterpret_cast<const T *>(stage);
        const Border<T> *border = reinterpret_cast<const Border<T> *>(stage + p.off_border);
        const PixGeom<T> *pix = reinterpret_cast<const PixGeom<T> *>(stage + p.off_pix);
        const Knot<T> *kread = knots - tile.jlo;               // indexed by source pixel
        T *obase = p.out + (size_t)g0 * frame_out + (size_t)s * p.rps * p.nout;
        int wb = -(tile.r0 * wp1 + tile.c0);                   // window addressed by frame coords
        pin(wb);

        if (lane < G) {
            rect[lane * rpitch - 1] = T(0);
            rect[lane * rpitch + npx] = T(0);
        }
        int jmin[G], jmax[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            jmin[g] = INT_MAX;
            jmax[g] = -1;
        }

        // every warp owns whole rows of the tile: no block-level barrier inside
        for (int yy = warp; yy < tile.ny; yy += nwarps) {
            const int y = tile.y0 + yy;
            const size_t rowtab = (size_t)s * NROWS + y;
            const int32_t *brow = p.base + rowtab * W + tile.jlo + lane;
            const T *wrow = p.w + rowtab * (K * K) * W + tile.jlo + lane;
            T *orow = obase + (size_t)y * p.nout;
            pin(orow);

            // ---- A: rectified flux, K*K weighted gather from the staged window.
            //         The table entries of the next 32 pixels are requested before the
            //         current ones are used (the tables are padded past their end).
            {
                int32_t npk = __ldg(brow);
                T nwt[K * K];
#pragma unroll
                for (int tap = 0; tap < K * K; ++tap) nwt[tap] = __ldg(wrow + tap * W);
                for (int p0 = 0; p0 < npx; p0 += 32) {
                    const int32_t pk = npk;
                    T wt[K * K];
#pragma unroll
                    for (int tap = 0; tap < K * K; ++tap) wt[tap] = nwt[tap];
                    if (p0 + 32 < npx) {
                        npk = __ldg(brow + p0 + 32);
#pragma unroll
                        for (int tap = 0; tap < K * K; ++tap)
                            nwt[tap] = __ldg(wrow + p0 + 32 + tap * W);
                    }
                    if (p0 + lane < npx) {
                        const T *src = win + (wb + base_row(pk) * wp1 + base_col(pk));
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            const T *sg = src + g * wsz;
                            T acc = T(0);
#pragma unroll
                            for (int a = 0; a < K; ++a)
#pragma unroll
                                for (int b = 0; b < K; ++b)
                                    acc = fma(wt[a * K + b], sg[a * wp1 + b], acc);
                            rect[g * rpitch + p0 + lane] = acc;
                        }
                    }
                }
            }
            __syncwarp();

            // ---- C: Steffen-limited slopes at both ends of every source pixel, scaled
            //         by the pixel width (flux units).  Branch-free: the table holds
            //         qm = 0 for the first pixel of the spectrum and qp = 0 for the last
            //         one, which gives the zero end slopes; the halo of `rect` is zero.
            for (int p0 = 0; p0 < npx; p0 += 32) {
                if (p0 + lane < npx) {
                    const PixGeom<T> pg = pix[p0 + lane];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const T *rv = rect + g * rpitch + p0 + lane;
                        const T c = rv[0];
                        const T cl = rv[-1] * pg.qm;
                        const T cr = rv[1] * pg.qp;
                        Knot<T> kn;
                        kn.rect = c;
                        kn.ya = steffen_slope(cl, c, pg.r0);
                        kn.yb = steffen_slope(c, cr, pg.r1);
                        kn.pad = T(0);
                        knots[g * npxp + p0 + lane] = kn;
                    }
                }
            }
            __syncwarp();

            // ---- D: cumulative flux at the output borders (32 borders = 31 pixels per
            //         warp pass), difference, coalesced store, non-zero channel range.
            //         Lanes past the tile's last border re-read that border: their source
            //         pixel stays inside this warp's scratch and their result is dropped.
            {
                const bool store = y < p.rps;
                for (int q0 = 0; q0 < tile.nk; q0 += 31) {
                    const int q = q0 + lane;                   // border, relative to the tile
                    const Border<T> bd = border[min(q, tile.nk)];
                    const int ib = bd.idx;
                    const int ibn = __shfl_down_sync(0xffffffffu, ib, 1);
                    const bool col_ok = lane < 31 && q < tile.nk;
                    const int kb = tile.k0 + q;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const Knot<T> *kr = kread + g * npxp;
                        const Knot<T> kn = kr[ib];
                        const T P = steffen_partial_flux(kn.rect, kn.ya, kn.yb, bd.tau);
                        const T Pn = __shfl_down_sync(0xffffffffu, P, 1);
                        T whole = (ibn > ib) ? kn.rect : T(0);
                        if (ibn > ib + 1) whole += kr[ib + 1].rect;
                        if (SPANLOOP)
                            for (int i = ib + 2; i < ibn; ++i) whole += kr[i].rect;
                        const T o = whole + (Pn - P);
                        const bool okg = col_ok && g < nlive;
                        if (okg && store) __stcs(orow + ((unsigned)kb + (unsigned)g * fo), o);
                        if (okg && o != T(0)) {
                            jmin[g] = min(jmin[g], kb);
                            jmax[g] = max(jmax[g], kb);
                        }
                    }
                }
            }
            __syncwarp();
        }

        // ---- first / last non-zero channel of the slitlet (JMNSLTnn / JMXSLTnn)
        if (p.jmm != nullptr) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int lo = __reduce_min_sync(0xffffffffu, jmin[g]);
                const int hi = __reduce_max_sync(0xffffffffu, jmax[g]);
                if (lane == 0 && hi >= 0 && g < nlive) {
                    int32_t *dst = p.jmm + ((size_t)(g0 + g) * p.nsl + s) * 2;
                    atomicMin(dst, lo);
                    atomicMax(dst + 1, hi);
                }
            }
        }

        if (wnext >= nwork) break;
        w = wnext;
        __syncthreads();          // stage (it & 1) is free again for the prefetch of it + 2
    }
}

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
    void *d_w = nullptr;
    void *d_border = nullptr, *d_pix = nullptr;
    Tile *d_tiles = nullptr;

    std::vector<SlitDev> h_slit;
    std::vector<int32_t> h_base;
    std::vector<int32_t> h_idx;
    std::vector<double> h_t;
    std::vector<Tile> h_tiles;

    int ntiles = 0, ntiles_empty = 0;
    int tile_k = 0, tile_rows = 0;
    int wrows_max = 0, wcols_max = 0, wpitch = 0, npxp = 0, max_span = 0;
    int tune_G = 0, tune_threads = 0;
    bool fast_pitch = false;     // npxp / wpitch equal the compile-time constants
    int num_sms = 148;
    int64_t table_bytes = 0;

    std::mutex mu;
    struct Slot {
        void *d_in = nullptr;
        void *d_out = nullptr;
        int32_t *d_jmm = nullptr;
        cudaStream_t stream = nullptr;
    } slots[HOST_SLOTS];
    int ws_chunk = 0;
};

namespace {

int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}

// launch configuration for a given number of frames per CTA
struct LaunchCfg {
    int G, threads;
    int stage_bytes, off_border, off_pix;
    size_t smem;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// shared memory of the persistent kernel:
//   [3 tile descriptors][2 stages: G windows | border table | pixel table][warp scratch]
void smem_layout(const emirwv_plan *pl, int G, int threads, LaunchCfg &cfg) {
    const size_t es = pl->esize;
    const size_t win = (size_t)G * pl->wrows_max * pl->wpitch * es;
    const size_t border = align_up((size_t)(pl->tile_k + 1) * 2 * es, 16);
    const size_t pix = (size_t)pl->npxp * 4 * es;
    cfg.off_border = (int)align_up(win, 16);
    cfg.off_pix = (int)(cfg.off_border + border);
    cfg.stage_bytes = (int)align_up(cfg.off_pix + pix, 128);
    const size_t scratch = (size_t)(threads / 32) * G * (pl->npxp * 5 + 4) * es;
    cfg.smem = 3 * sizeof(Tile) + 2 * (size_t)cfg.stage_bytes + scratch;
}

LaunchCfg choose_cfg(const emirwv_plan *pl, int nframes) {
    LaunchCfg cfg;
    cfg.threads = pl->tune_threads > 0 ? pl->tune_threads : env_int("EMIRWV_THREADS", 256);
    cfg.threads = std::max(32, std::min(MAXTHREADS, (cfg.threads / 32) * 32));
    int G = pl->tune_G > 0 ? pl->tune_G : env_int("EMIRWV_G", 2);
    G = (G >= 4) ? 4 : (G >= 2 ? 2 : 1);
    smem_layout(pl, G, cfg.threads, cfg);
    while (G > 1 && (G > nframes || cfg.smem > (size_t)SMEM_LIMIT)) {
        G >>= 1;
        smem_layout(pl, G, cfg.threads, cfg);
    }
    while (cfg.threads > 32 && cfg.smem > (size_t)SMEM_LIMIT) {
        cfg.threads -= 32;
        smem_layout(pl, G, cfg.threads, cfg);
    }
    cfg.G = G;
    return cfg;
}

constexpr int FAST_NPXP = 144;   // scratch pitch of the specialised kernels
constexpr int FAST_WP = 160;     // window pitch of the specialised kernels

template <typename T, int K, int G, int NPXP_C, int WP_C, bool SPANLOOP>
int launch_final(const emirwv_plan *pl, RunParams<T> rp, const LaunchCfg &cfg, cudaStream_t st) {
    auto kern = k_rectwv<T, K, G, NPXP_C, WP_C, SPANLOOP>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)cfg.smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, cfg.threads, cfg.smem));
    if (per_sm < 1) return fail(EMIRWV_E_LIMIT, "kernel does not fit on an SM");
    rp.ngroups = (rp.nframes + G - 1) / G;
    const long long nwork = (long long)rp.nlive_tiles * rp.ngroups;
    if (nwork > 0) {
        if (nwork > INT_MAX) return fail(EMIRWV_E_LIMIT, "too many frames in one call");
        const int grid = (int)std::min<long long>(nwork, (long long)pl->num_sms * per_sm);
        kern<<<grid, cfg.threads, cfg.smem, st>>>(rp);
        CUDA_TRY(cudaGetLastError());
    }
    return EMIRWV_OK;
}

template <typename T, int K, int G>
int launch_one(const emirwv_plan *pl, const RunParams<T> &rp, const LaunchCfg &cfg,
               cudaStream_t st) {
    const bool span = pl->max_span > 2;
    if (pl->fast_pitch) {
        if (span) return launch_final<T, K, G, FAST_NPXP, FAST_WP, true>(pl, rp, cfg, st);
        return launch_final<T, K, G, FAST_NPXP, FAST_WP, false>(pl, rp, cfg, st);
    }
    if (span) return launch_final<T, K, G, 0, 0, true>(pl, rp, cfg, st);
    return launch_final<T, K, G, 0, 0, false>(pl, rp, cfg, st);
}

template <typename T, int K>
int launch_k(const emirwv_plan *pl, const RunParams<T> &rp, const LaunchCfg &cfg,
             cudaStream_t st) {
    switch (cfg.G) {
        case 4: return launch_one<T, K, 4>(pl, rp, cfg, st);
        case 2: return launch_one<T, K, 2>(pl, rp, cfg, st);
        default: return launch_one<T, K, 1>(pl, rp, cfg, st);
    }
}

template <typename T>
int launch_t(const emirwv_plan *pl, const void *d_in, void *d_out, int nframes,
             int32_t *d_jmm, cudaStream_t st) {
    const LaunchCfg cfg = choose_cfg(pl, nframes);
    if (cfg.smem > (size_t)SMEM_LIMIT)
        return fail(EMIRWV_E_LIMIT, "tile needs %zu bytes of shared memory", cfg.smem);
    if ((reinterpret_cast<uintptr_t>(d_in) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u))
        return fail(EMIRWV_E_VALUE, "device buffers must be 16-byte aligned");
    RunParams<T> rp;
    rp.in = static_cast<const T *>(d_in);
    rp.out = static_cast<T *>(d_out);
    rp.jmm = d_jmm;
    rp.nframes = nframes;
    rp.tiles = pl->d_tiles;
    rp.base = pl->d_base;
    rp.w = static_cast<const T *>(pl->d_w);
    rp.border = static_cast<const Border<T> *>(pl->d_border);
    rp.pix = static_cast<const PixGeom<T> *>(pl->d_pix);
    rp.naxis2 = pl->desc.naxis2;
    rp.nout = pl->nout;
    rp.NB = pl->NB;
    rp.rps = pl->rps;
    rp.nrows_out = pl->nsl * pl->rps;
    rp.nsl = pl->nsl;
    rp.wpitch = pl->wpitch;
    rp.wrows = pl->wrows_max;
    rp.npxp = pl->npxp;
    rp.max_span = pl->max_span;
    rp.vec_ok = ((size_t)pl->nout * sizeof(T)) % 16 == 0;
    rp.nlive_tiles = pl->ntiles - pl->ntiles_empty;
    rp.ngroups = 0;
    rp.stage_bytes = cfg.stage_bytes;
    rp.off_border = cfg.off_border;
    rp.off_pix = cfg.off_pix;
    if (d_jmm != nullptr) {
        const int n = nframes * pl->nsl * 2;
        k_init_jmm<<<(n + 255) / 256, 256, 0, st>>>(d_jmm, n);
        CUDA_TRY(cudaGetLastError());
    }
    if (pl->ntiles_empty > 0) {
        const dim3 grid(pl->ntiles_empty, std::min(nframes, 32));
        k_zero_fill<T><<<grid, 256, 0, st>>>(rp, rp.nlive_tiles);
        CUDA_TRY(cudaGetLastError());
    }
    switch (pl->K) {
        case 1: return launch_k<T, 1>(pl, rp, cfg, st);
        case 2: return launch_k<T, 2>(pl, rp, cfg, st);
        case 3: return launch_k<T, 3>(pl, rp, cfg, st);
        case 4: return launch_k<T, 4>(pl, rp, cfg, st);
    }
    return fail(EMIRWV_E_LIMIT, "unsupported footprint %d", pl->K);
}

// numpy.polynomial.polynomial.polyval: Horner, one rounding per operation
double polyval_np(double x, const double *c, int n) {
    double c0 = c[n - 1];
    for (int m = n - 2; m >= 0; --m) c0 = add_rn(c[m], mul_rn(c0, x));
    return c0;
}

template <typename T>
int upload_tables(emirwv_plan *pl, const std::vector<double> &tau, const std::vector<double> &h) {
    const int nsl = pl->nsl, NB = pl->NB;
    // both tables carry 64 extra entries: the kernel requests the row of its next
    // pass before it knows whether there is one
    std::vector<Border<T>> border((size_t)nsl * NB + 64);
    memset(border.data(), 0, border.size() * sizeof(Border<T>));
    for (size_t k = 0; k < (size_t)nsl * NB; ++k) {
        border[k].tau = (T)tau[k];
        border[k].idx = pl->h_idx[k];
    }
    std::vector<PixGeom<T>> pix((size_t)nsl * W + 64);
    memset(pix.data(), 0, pix.size() * sizeof(PixGeom<T>));
    for (int s = 0; s < nsl; ++s) {
        const emirwv_slitlet_desc &sl = pl->desc.slitlet[s];
        if (!sl.valid) continue;
        const int nchan = sl.bb_nc2_orig - sl.bb_nc1_orig + 1;
        const double *hs = h.data() + (size_t)s * W;
        for (int i = 0; i < nchan; ++i) {
            PixGeom<T> &pg = pix[(size_t)s * W + i];
            pg.qm = i > 0 ? (T)(hs[i] / hs[i - 1]) : T(0);
            pg.qp = i + 1 < nchan ? (T)(hs[i] / hs[i + 1]) : T(0);
            pg.r0 = i > 0 ? (T)(hs[i] / (hs[i - 1] + hs[i])) : T(0);
            pg.r1 = i + 1 < nchan ? (T)(hs[i + 1] / (hs[i] + hs[i + 1])) : T(0);
        }
    }
    CUDA_TRY(cudaMalloc(&pl->d_border, border.size() * sizeof(Border<T>)));
    CUDA_TRY(cudaMemcpy(pl->d_border, border.data(), border.size() * sizeof(Border<T>),
                        cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&pl->d_pix, pix.size() * sizeof(PixGeom<T>)));
    CUDA_TRY(cudaMemcpy(pl->d_pix, pix.data(), pix.size() * sizeof(PixGeom<T>),
                        cudaMemcpyHostToDevice));
    pl->table_bytes += (int64_t)(border.size() * sizeof(Border<T>) + pix.size() * sizeof(PixGeom<T>));
    return EMIRWV_OK;
}

// Border tables of every slitlet; float64 with numpy's operation order so that the
// interval of every border is the one numpy.searchsorted finds in the reference's
// resample_image2d_flux / SteffenInterpolator (apply_rectwv_coeff.py:152-159).
int build_border_tables(emirwv_plan *pl, std::vector<uint8_t> &covered) {
    const emirwv_plan_desc &d = pl->desc;
    const int nout = pl->nout, NB = pl->NB, nsl = pl->nsl;
    std::vector<double> nb(NB), wl(nout);
    for (int k = 0; k < nout; ++k) wl[k] = add_rn(d.crval1, mul_rn(d.cdelt1, (double)k));
    for (int k = 1; k < nout; ++k) nb[k] = mul_rn(0.5, add_rn(wl[k], wl[k - 1]));
    nb[0] = add_rn(mul_rn(2.0, wl[0]), -nb[1]);
    nb[nout] = add_rn(mul_rn(2.0, wl[nout - 1]), -nb[nout - 1]);

    pl->h_idx.assign((size_t)nsl * NB, 0);
    pl->h_t.assign((size_t)nsl * NB, 0.0);
    std::vector<double> h_tau((size_t)nsl * NB, 0.0);
    std::vector<double> h_h((size_t)nsl * W, 0.0);
    covered.assign((size_t)nsl * NB, 0);
    pl->max_span = 0;

    std::vector<double> xw(W + 1);
    for (int s = 0; s < nsl; ++s) {
        const emirwv_slitlet_desc &sl = d.slitlet[s];
        if (!sl.valid) continue;
        const int nchan = sl.bb_nc2_orig - sl.bb_nc1_orig + 1;
        double *h = h_h.data() + (size_t)s * W;
        for (int i = 0; i <= nchan; ++i) {
            const double xb = add_rn(add_rn(-0.5, (double)i), d.crpix1);
            xw[i] = polyval_np(xb, sl.wpoly, sl.nwpoly);
        }
        for (int i = 0; i < nchan; ++i) {
            h[i] = xw[i + 1] - xw[i];
            if (!(h[i] > 0.0))
                return fail(EMIRWV_E_VALUE,
                            "slitlet %d: wavelength polynomial is not increasing at pixel %d",
                            s + 1, i + 1);
        }
        int prev = 0;
        for (int k = 0; k < NB; ++k) {
            const double x = nb[k];
            int idx;
            double t, tau;
            uint8_t cov = 1;
            if (x < xw[0]) {                       // extrapolate='border': y_0
                idx = 0;
                t = 0.0;
                tau = 0.0;
                cov = 0;
            } else if (x > xw[nchan]) {            // extrapolate='border': y_N
                idx = nchan - 1;
                t = h[nchan - 1];
                tau = 1.0;
                cov = 0;
            } else {                               // searchsorted(side='left').clip(1, N) - 1
                const int pos = (int)(std::lower_bound(xw.begin(), xw.begin() + nchan + 1, x) -
                                      xw.begin());
                idx = std::min(std::max(pos, 1), nchan) - 1;
                t = x - xw[idx];
                tau = t / h[idx];
            }
            pl->h_idx[(size_t)s * NB + k] = idx;
            pl->h_t[(size_t)s * NB + k] = t;
            h_tau[(size_t)s * NB + k] = tau;
            covered[(size_t)s * NB + k] = cov;
            if (k > 0) pl->max_span = std::max(pl->max_span, idx - prev);
            prev = idx;
        }
    }
    if (pl->dtype == EMIRWV_F32) return upload_tables<float>(pl, h_tau, h_h);
    return upload_tables<double>(pl, h_tau, h_h);
}

// Work items: (slitlet, block of rectified rows, block of output columns).
int build_tiles(emirwv_plan *pl, const std::vector<uint8_t> &covered) {
    const emirwv_plan_desc &d = pl->desc;
    const int nout = pl->nout, NB = pl->NB, K = pl->K;
    const int TK = pl->tile_k, TR = pl->tile_rows;
    std::vector<Tile> live, dead;
    pl->wrows_max = pl->wcols_max = 1;
    int npx_max = 1;
    for (int s = 0; s < pl->nsl; ++s) {
        const emirwv_slitlet_desc &sl = d.slitlet[s];
        const int nchan = sl.valid ? sl.bb_nc2_orig - sl.bb_nc1_orig + 1 : 0;
        for (int y0 = 0; y0 < NROWS; y0 += TR) {
            const int ny = std::min(TR, NROWS - y0);
            for (int k0 = 0; k0 < nout; k0 += TK) {
                Tile t;
                memset(&t, 0, sizeof(t));
                t.slit = s;
                t.y0 = y0;
                t.ny = ny;
                t.k0 = k0;
                t.nk = std::min(TK, nout - k0);
                t.nchan = nchan;
                bool any = false;
                if (sl.valid)
                    for (int k = k0; k <= k0 + t.nk; ++k) any |= covered[(size_t)s * NB + k] != 0;
                // a pixel whose two borders lie on either side of the spectrum also counts
                if (sl.valid && !any)
                    any = pl->h_idx[(size_t)s * NB + k0] != pl->h_idx[(size_t)s * NB + k0 + t.nk];
                if (!any) {
                    t.empty = 1;
                    if (y0 >= pl->rps) continue;           // the ghost row stores nothing
                    // neighbouring empty tiles become one wide run of zeros
                    if (!dead.empty() && dead.back().slit == s && dead.back().y0 == y0 &&
                        dead.back().k0 + dead.back().nk == k0 && dead.back().nk + t.nk <= 1024)
                        dead.back().nk += t.nk;
                    else
                        dead.push_back(t);
                    continue;
                }
                const int ilo = pl->h_idx[(size_t)s * NB + k0];
                const int ihi = pl->h_idx[(size_t)s * NB + k0 + t.nk];
                t.jlo = std::max(0, ilo - 1);
                const int jhi = std::min(nchan - 1, ihi + 1);
                t.npx = jhi - t.jlo + 1;
                int r0 = INT_MAX, r1 = INT_MIN, c0 = INT_MAX, c1 = INT_MIN;
                for (int y = y0; y < y0 + ny; ++y) {
                    const int32_t *row = pl->h_base.data() + ((size_t)s * NROWS + y) * W;
                    for (int j = t.jlo; j <= jhi; ++j) {
                        const int br = base_row(row[j]), bc = base_col(row[j]);
                        r0 = std::min(r0, br);
                        r1 = std::max(r1, br + K - 1);
                        c0 = std::min(c0, bc);
                        c1 = std::max(c1, bc + K - 1);
                    }
                }
                c0 = (int)(std::floor(c0 / 4.0) * 4.0);    // 16-byte aligned window rows
                t.r0 = r0;
                t.c0 = c0;
                t.wrows = r1 - r0 + 1;
                t.wcols = ((c1 - c0 + 1) + 3) & ~3;
                pl->wrows_max = std::max(pl->wrows_max, t.wrows);
                pl->wcols_max = std::max(pl->wcols_max, t.wcols);
                npx_max = std::max(npx_max, t.npx);
                live.push_back(t);
            }
        }
    }
    pl->ntiles_empty = (int)dead.size();
    pl->h_tiles = live;
    pl->h_tiles.insert(pl->h_tiles.end(), dead.begin(), dead.end());
    pl->ntiles = (int)pl->h_tiles.size();
    pl->wpitch = pl->wcols_max;                       // multiple of 4
    if ((pl->wpitch & 31) == 0) pl->wpitch += 4;      // rows start in different banks
    pl->npxp = (npx_max + 3) & ~3;
    pl->fast_pitch = pl->npxp <= FAST_NPXP && pl->wcols_max <= FAST_WP;
    if (pl->fast_pitch) {
        pl->npxp = FAST_NPXP;
        pl->wpitch = FAST_WP;
    }
    if (pl->ntiles > 65535) return fail(EMIRWV_E_LIMIT, "too many tiles (%d)", pl->ntiles);
    LaunchCfg probe;
    smem_layout(pl, 1, 32, probe);
    if (probe.smem > (size_t)SMEM_LIMIT)
        return fail(EMIRWV_E_LIMIT,
                    "distortion too strong: a %d x %d source window does not fit in shared memory",
                    pl->wrows_max, pl->wcols_max);
    CUDA_TRY(cudaMalloc(&pl->d_tiles, pl->h_tiles.size() * sizeof(Tile)));
    CUDA_TRY(cudaMemcpy(pl->d_tiles, pl->h_tiles.data(), pl->h_tiles.size() * sizeof(Tile),
                        cudaMemcpyHostToDevice));
    pl->table_bytes += (int64_t)(pl->h_tiles.size() * sizeof(Tile));
    return EMIRWV_OK;
}

int validate(const emirwv_plan_desc *d) {
    // Slitlet2D.extract_slitlet2d accepts 2048 x 2048 frames only (slitlet2d.py:359-363)
    if (d->naxis1 != W) return fail(EMIRWV_E_VALUE, "Unexpected naxis1");
    if (d->naxis2 < 2 || d->naxis2 > 16384) return fail(EMIRWV_E_VALUE, "Unexpected naxis2");
    if (d->nslitlets < 1 || d->nslitlets > EMIRWV_MAX_SLITLETS)
        return fail(EMIRWV_E_VALUE, "nslitlets=%d out of range", d->nslitlets);
    if (d->rows_per_slitlet != NROWS - 1)
        return fail(EMIRWV_E_VALUE, "rows_per_slitlet must be %d", NROWS - 1);
    if (d->resampling < 1 || d->resampling > 3)
        return fail(EMIRWV_E_VALUE, "Unexpected resampling value=%d", d->resampling);
    if (d->dtype != EMIRWV_F32 && d->dtype != EMIRWV_F64)
        return fail(EMIRWV_E_VALUE, "unknown dtype %d", d->dtype);
    if (d->naxis1_out < 2 || d->naxis1_out > 65536)
        return fail(EMIRWV_E_VALUE, "naxis1_out=%d out of range", d->naxis1_out);
    if (!(d->cdelt1 > 0.0)) return fail(EMIRWV_E_VALUE, "cdelt1 must be positive");
    const int cap = d->max_order > 0 ? d->max_order : 4;
    for (int s = 0; s < d->nslitlets; ++s) {
        const emirwv_slitlet_desc &sl = d->slitlet[s];
        if (!sl.valid) continue;
        const int order = order_from_ncoef(sl.ncoef);
        if (order < 0 || order > cap)
            return fail(EMIRWV_E_VALUE, "order > %d not implemented", cap);
        if (sl.nwpoly < 1 || sl.nwpoly > EMIRWV_MAX_WPOLY)
            return fail(EMIRWV_E_VALUE, "slitlet %d: %d wavelength coefficients", s + 1, sl.nwpoly);
        if (sl.bb_nc1_orig < 1 || sl.bb_nc2_orig > d->naxis1 || sl.bb_nc2_orig < sl.bb_nc1_orig + 1 ||
            sl.bb_ns1_orig < 1 || sl.bb_ns2_orig > d->naxis2 || sl.bb_ns2_orig < sl.bb_ns1_orig)
            return fail(EMIRWV_E_VALUE, "slitlet %d: bounding box outside the frame", s + 1);
        const int bbh = sl.bb_ns2_orig - sl.bb_ns1_orig + 1;
        if (sl.max_row_rectified - sl.min_row_rectified + 1 != d->rows_per_slitlet)
            return fail(EMIRWV_E_VALUE,
                        "could not broadcast input array from shape (%d,%d) into shape (%d,%d)",
                        sl.max_row_rectified - sl.min_row_rectified + 1, d->naxis1_out,
                        d->rows_per_slitlet, d->naxis1_out);
        if (sl.min_row_rectified < 0)
            return fail(EMIRWV_E_VALUE, "slitlet %d: negative min_row_rectified", s + 1);
        if (sl.max_row_rectified + 1 >= bbh)
            return fail(EMIRWV_E_INDEX, "index %d is out of bounds for axis 0 with size %d",
                        sl.max_row_rectified + 1, bbh);
    }
    return EMIRWV_OK;
}

int build_plan(emirwv_plan *pl) {
    const emirwv_plan_desc &d = pl->desc;
    CUDA_TRY(cudaGetDevice(&pl->device));
    CUDA_TRY(cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, pl->device));
    pl->dtype = d.dtype;
    pl->mode = d.resampling;
    pl->esize = d.dtype == EMIRWV_F32 ? 4 : 8;
    pl->nsl = d.nslitlets;
    pl->rps = d.rows_per_slitlet;
    pl->nout = d.naxis1_out;
    pl->NB = d.naxis1_out + 1;
    pl->tile_k = d.tile_k > 0 ? d.tile_k : env_int("EMIRWV_TILE_K", 124);
    pl->tile_rows = d.tile_rows > 0 ? d.tile_rows : env_int("EMIRWV_TILE_ROWS", NROWS);
    pl->tile_k = std::max(31, std::min(pl->tile_k, 1024));
    pl->tile_rows = std::max(1, std::min(pl->tile_rows, NROWS));

    pl->h_slit.resize(pl->nsl);
    for (int s = 0; s < pl->nsl; ++s) {
        const emirwv_slitlet_desc &sl = d.slitlet[s];
        SlitDev &sd = pl->h_slit[s];
        memset(&sd, 0, sizeof(sd));
        sd.valid = sl.valid ? 1 : 0;
        if (!sl.valid) continue;
        sd.ns1 = sl.bb_ns1_orig;
        sd.nc1 = sl.bb_nc1_orig;
        sd.bbh = sl.bb_ns2_orig - sl.bb_ns1_orig + 1;
        sd.bbw = sl.bb_nc2_orig - sl.bb_nc1_orig + 1;
        sd.min_row = sl.min_row_rectified;
        sd.order = order_from_ncoef(sl.ncoef);
        memcpy(sd.a, sl.aij, sizeof(double) * sl.ncoef);
        memcpy(sd.b, sl.bij, sizeof(double) * sl.ncoef);
    }
    CUDA_TRY(cudaMalloc(&pl->d_slit, pl->nsl * sizeof(SlitDev)));
    CUDA_TRY(cudaMemcpy(pl->d_slit, pl->h_slit.data(), pl->nsl * sizeof(SlitDev),
                        cudaMemcpyHostToDevice));

    GeomParams gp;
    gp.slit = pl->d_slit;
    gp.nsl = pl->nsl;
    gp.naxis1 = d.naxis1;
    gp.naxis2 = d.naxis2;
    gp.offx = d.offx;
    gp.offy = d.offy;
    gp.mode = pl->mode;
    gp.kmax = nullptr;
    const dim3 block(128);
    const dim3 grid((W + 127) / 128, NROWS, pl->nsl);

    if (pl->mode == EMIRWV_AREA) {
        int *d_kmax = nullptr;
        int kmax = 1;
        CUDA_TRY(cudaMalloc(&d_kmax, sizeof(int)));
        CUDA_TRY(cudaMemcpy(d_kmax, &kmax, sizeof(int), cudaMemcpyHostToDevice));
        gp.kmax = d_kmax;
        gp.K = 0;
        gp.base = nullptr;
        k_extent<<<grid, block>>>(gp);
        cudaError_t err = cudaGetLastError();
        if (err == cudaSuccess)
            err = cudaMemcpy(&kmax, d_kmax, sizeof(int), cudaMemcpyDeviceToHost);
        cudaFree(d_kmax);
        if (err != cudaSuccess)
            return fail(EMIRWV_E_CUDA, "k_extent failed: %s", cudaGetErrorString(err));
        if (kmax > MAXK)
            return fail(EMIRWV_E_LIMIT,
                        "distortion too strong: an output pixel covers %d source pixels "
                        "along one axis (limit %d)", kmax, MAXK);
        pl->K = kmax;
    } else {
        pl->K = pl->mode == EMIRWV_NEAREST ? 1 : 2;
    }

    const size_t nrowtab = (size_t)pl->nsl * NROWS;
    CUDA_TRY(cudaMalloc(&pl->d_base, nrowtab * W * sizeof(int32_t)));
    CUDA_TRY(cudaMalloc(&pl->d_w, nrowtab * pl->K * pl->K * W * pl->esize));
    pl->table_bytes += (int64_t)(nrowtab * W * sizeof(int32_t));
    pl->table_bytes += (int64_t)(nrowtab * pl->K * pl->K * W * pl->esize);
    gp.K = pl->K;
    gp.base = pl->d_base;
    gp.kmax = nullptr;
    if (pl->dtype == EMIRWV_F32)
        k_build_geometry<float><<<grid, block>>>(gp, static_cast<float *>(pl->d_w));
    else
        k_build_geometry<double><<<grid, block>>>(gp, static_cast<double *>(pl->d_w));
    CUDA_TRY(cudaGetLastError());
    pl->h_base.resize(nrowtab * W);
    CUDA_TRY(cudaMemcpy(pl->h_base.data(), pl->d_base, pl->h_base.size() * sizeof(int32_t),
                        cudaMemcpyDeviceToHost));

    std::vector<uint8_t> covered;
    int rc = build_border_tables(pl, covered);
    if (rc) return rc;
    return build_tiles(pl, covered);
}

void release_workspace(emirwv_plan *pl) {
    for (auto &sl : pl->slots) {
        if (sl.d_in) cudaFree(sl.d_in);
        if (sl.d_out) cudaFree(sl.d_out);
        if (sl.d_jmm) cudaFree(sl.d_jmm);
        if (sl.stream) cudaStreamDestroy(sl.stream);
        sl = emirwv_plan::Slot();
    }
    pl->ws_chunk = 0;
}

int ensure_workspace(emirwv_plan *pl, int chunk) {
    if (pl->ws_chunk >= chunk) return EMIRWV_OK;
    release_workspace(pl);
    const size_t in_b = (size_t)pl->desc.naxis1 * pl->desc.naxis2 * pl->esize;
    const size_t out_b = (size_t)pl->nsl * pl->rps * pl->nout * pl->esize;
    for (auto &sl : pl->slots) {
        CUDA_TRY(cudaMalloc(&sl.d_in, in_b * chunk));
        CUDA_TRY(cudaMalloc(&sl.d_out, out_b * chunk));
        CUDA_TRY(cudaMalloc(&sl.d_jmm, (size_t)chunk * pl->nsl * 2 * sizeof(int32_t)));
        CUDA_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    }
    pl->ws_chunk = chunk;
    return EMIRWV_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

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
    cudaFree(pl->d_slit);
    cudaFree(pl->d_base);
    cudaFree(pl->d_w);
    cudaFree(pl->d_border);
    cudaFree(pl->d_pix);
    cudaFree(pl->d_tiles);
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
    info->tile_rows = pl->tile_rows;
    info->window_rows = pl->wrows_max;
    info->window_cols = pl->wcols_max;
    info->max_span = pl->max_span;
    info->frames_per_cta = cfg.G;
    info->threads = cfg.threads;
    info->smem_bytes = (int32_t)cfg.smem;
    info->table_bytes = pl->table_bytes;
    info->in_frame_bytes = (int64_t)pl->desc.naxis1 * pl->desc.naxis2 * pl->esize;
    info->out_frame_bytes = (int64_t)pl->nsl * pl->rps * pl->nout * pl->esize;
    return EMIRWV_OK;
}

int emirwv_plan_set_tuning(emirwv_plan *pl, int frames_per_cta, int threads) {
    if (pl == nullptr) return fail(EMIRWV_E_VALUE, "null plan");
    pl->tune_G = frames_per_cta;
    pl->tune_threads = threads;
    return EMIRWV_OK;
}

int emirwv_run(const emirwv_plan *pl, const void *d_in, void *d_out, int nframes,
               int32_t *d_jminmax, void *cuda_stream) {
    if (pl == nullptr || d_in == nullptr || d_out == nullptr)
        return fail(EMIRWV_E_VALUE, "null argument");
    if (nframes < 0) return fail(EMIRWV_E_VALUE, "negative frame count");
    if (nframes == 0) return EMIRWV_OK;
    DeviceGuard guard(pl->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (pl->dtype == EMIRWV_F32) return launch_t<float>(pl, d_in, d_out, nframes, d_jminmax, st);
    return launch_t<double>(pl, d_in, d_out, nframes, d_jminmax, st);
}

int emirwv_run_host(emirwv_plan *pl, const void *h_in, void *h_out, int nframes,
                    int32_t *h_jminmax) {
    if (pl == nullptr || h_in == nullptr || h_out == nullptr)
        return fail(EMIRWV_E_VALUE, "null argument");
    if (nframes < 0) return fail(EMIRWV_E_VALUE, "negative frame count");
    if (nframes == 0) return EMIRWV_OK;
    DeviceGuard guard(pl->device);
    std::lock_guard<std::mutex> lock(pl->mu);
    const int chunk = std::max(1, std::min(nframes, env_int("EMIRWV_HOST_CHUNK", 8)));
    int rc = ensure_workspace(pl, chunk);
    if (rc) return rc;
    const size_t in_b = (size_t)pl->desc.naxis1 * pl->desc.naxis2 * pl->esize;
    const size_t out_b = (size_t)pl->nsl * pl->rps * pl->nout * pl->esize;
    const size_t jmm_b = (size_t)pl->nsl * 2 * sizeof(int32_t);
    int slot = 0;
    for (int f0 = 0; f0 < nframes; f0 += chunk, slot = (slot + 1) % HOST_SLOTS) {
        const int n = std::min(chunk, nframes - f0);
        emirwv_plan::Slot &sl = pl->slots[slot];
        CUDA_TRY(cudaMemcpyAsync(sl.d_in, static_cast<const char *>(h_in) + (size_t)f0 * in_b,
                                 in_b * n, cudaMemcpyHostToDevice, sl.stream));
        if (pl->dtype == EMIRWV_F32)
            rc = launch_t<float>(pl, sl.d_in, sl.d_out, n, sl.d_jmm, sl.stream);
        else
            rc = launch_t<double>(pl, sl.d_in, sl.d_out, n, sl.d_jmm, sl.stream);
        if (rc) break;
        CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(h_out) + (size_t)f0 * out_b, sl.d_out,
                                 out_b * n, cudaMemcpyDeviceToHost, sl.stream));
        if (h_jminmax != nullptr)
            CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char *>(h_jminmax) + (size_t)f0 * jmm_b,
                                     sl.d_jmm, jmm_b * n, cudaMemcpyDeviceToHost, sl.stream));
    }
    for (auto &sl : pl->slots) {
        cudaError_t err = cudaStreamSynchronize(sl.stream);
        if (err != cudaSuccess && rc == EMIRWV_OK)
            rc = fail(EMIRWV_E_CUDA, "stream failed: %s", cudaGetErrorString(err));
    }
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
        const size_t n = (size_t)NROWS * KK * W;
        std::vector<unsigned char> raw(n * pl->esize);
        CUDA_TRY(cudaMemcpy(raw.data(),
                            static_cast<const char *>(pl->d_w) + (size_t)s * n * pl->esize,
                            n * pl->esize, cudaMemcpyDeviceToHost));
        for (int y = 0; y < NROWS; ++y)
            for (int tap = 0; tap < KK; ++tap)
                for (int j = 0; j < sd.bbw; ++j) {
                    const size_t src = ((size_t)y * KK + tap) * W + j;
                    const double v = pl->dtype == EMIRWV_F32
                                         ? (double)reinterpret_cast<const float *>(raw.data())[src]
                                         : reinterpret_cast<const double *>(raw.data())[src];
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
