// Kernels either side of the hot path: zero fill, median_slitlets_rectified, the ABBA product
// (mean, median and sigma-clip combinations, per-frame vertical offsets), pre-reduction (bad
// pixels, bias, dark, flat).
// Included by emirwv.cu inside its anonymous namespace (one translation unit).

// Zero-coverage runs (outside the wavelength range of a slitlet, or a missing slitlet):
// exact zeros.  Only used when the copy warps of the hot kernel cannot write them (output rows
// that are not 16-byte aligned, or EMIRWV_FILL=kernel).  A small persistent grid: every warp takes rows (dead tile, frame, row) in a
// grid-stride loop and writes them with 16-byte stores (scalar head and tail where the run
// is not aligned), so the fill trickles along next to the hot kernel instead of bursting.
template <typename T>
__global__ void __launch_bounds__(128) k_zero_fill(const RunParams<T> p, int first_tile) {
    using V = typename Vec16<T>::type;
    constexpr int VN = Vec16<T>::N;
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long units = (long long)p.ndead_tiles * p.nframes * p.rps;
    V zero;
    memset(&zero, 0, sizeof(zero));
    for (long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < units;
         u += nwarps) {
        const int y = (int)(u % p.rps);
        const long long tf = u / p.rps;
        const int f = (int)(tf % p.nframes);
        const Tile &tile = p.tiles[first_tile + (int)(tf / p.nframes)];
        const int k0 = tile.k0, k1 = k0 + tile.nk;
        T *orow = p.out + ((size_t)f * p.nrows_out + (size_t)tile.slit * p.rps + y) * p.nout;
        if (p.vec_ok) {
            for (int c = (k0 / VN + lane) * VN; c < k1; c += 32 * VN) {
                if (c >= k0 && c + VN <= k1) {
                    __stcs(reinterpret_cast<V *>(orow + c), zero);
                } else {
#pragma unroll
                    for (int j = 0; j < VN; ++j)
                        if (c + j >= k0 && c + j < k1) __stcs(orow + c + j, T(0));
                }
            }
        } else {
            for (int c = k0 + lane; c < k1; c += 32) __stcs(orow + c, T(0));
        }
    }
}

// ---------------------------------------------------------------------------------
// median_slitlets_rectified, modes 0 and 1 (median_slitlets_rectified.py:85-97): the
// consumer right behind the path.  For every slitlet and column the median over the R = 38
// rows of the rectified product, i.e. numpy.median(axis=0) of an even number of values:
// the mean of the two middle ones, NaN when any value is NaN.
//
// One thread per column holds its 38 values in registers and runs a pruned selection network
// (compare-exchanges only, no data-dependent branches) that leaves the two middle values on
// two known wires.  Loads and stores are coalesced along the columns; the kernel reads the
// product once.
template <typename T, int R>
__global__ void __launch_bounds__(128) k_median_slitlets(const T *__restrict__ in,
                                                        T *__restrict__ out, int naxis1,
                                                        int nsl, int mode) {
    static_assert(R % 2 == 0 && R >= 4, "even number of rows per slitlet");
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int slit = blockIdx.y, f = blockIdx.z;
    if (j >= naxis1) return;
    const T *src = in + ((size_t)f * nsl + slit) * R * naxis1 + j;
    bool anynan = false;
    T lo_mid, hi_mid;                                          // the two middle values
    if constexpr (R == 38) {
        // selection network: 223 compare-exchanges bring the two middle values of the 38
        // inputs to wires 18 and 19 (scripts/gen_median_net.py)
        T w[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            w[i] = __ldcs(src + (size_t)i * naxis1);
            anynan |= w[i] != w[i];
        }
#define CE(a, b) order2(w[a], w[b]);
#include "median_net38.inc"
#undef CE
        lo_mid = w[R / 2 - 1];
        hi_mid = w[R / 2];
    } else {
        // any even R: a working set of R/2 + 2 values; its smallest and largest element
        // cannot be one of the two middle values of the whole column, so both are dropped
        // and the next value of the column takes a free place
        constexpr int M = R / 2 + 2;
        T w[M];
#pragma unroll
        for (int i = 0; i < M; ++i) {
            w[i] = __ldcs(src + (size_t)i * naxis1);
            anynan |= w[i] != w[i];
        }
#pragma unroll
        for (int step = 0; step <= R - M; ++step) {
            const int m = M - step;                            // current working-set size
#pragma unroll
            for (int i = 0; i + 1 < m; ++i) order2(w[i], w[i + 1]);      // maximum -> w[m-1]
#pragma unroll
            for (int i = m - 2; i >= 1; --i) order2(w[i - 1], w[i]);     // minimum -> w[0]
            if (step < R - M) {                                // drop both, take the next value
                w[0] = __ldcs(src + (size_t)(M + step) * naxis1);
                anynan |= w[0] != w[0];
            }
        }
        lo_mid = w[1];            // 4 values are left: minimum, the two middle ones, maximum
        hi_mid = w[2];
    }
    T med;
    if (sizeof(T) == 4)
        med = (T)__fmul_rn(__fadd_rn((float)lo_mid, (float)hi_mid), 0.5f);
    else
        med = (T)__dmul_rn(__dadd_rn((double)lo_mid, (double)hi_mid), 0.5);
    if (anynan) med = (T)NAN;
    if (mode == 1) {
        out[((size_t)f * nsl + slit) * naxis1 + j] = med;
    } else {
        T *dst = out + ((size_t)f * nsl + slit) * R * naxis1 + j;
#pragma unroll 2
        for (int i = 0; i < R; ++i) __stcs(dst + (size_t)i * naxis1, med);
    }
}

// ---------------------------------------------------------------------------------
// ABBA combination, method "mean" (recipes/spec/abba.py:579-606): the rectified frames of a
// rank are summed per nodding position in float64, in frame order, one pass over the stack
// (labels: +1 = A, -1 = B, 0 = not used).  The sums of all ranks are added by a collective;
// k_abba_finish forms mean(A) and mean(B), casts both to float32 as the reference does, and
// subtracts.
template <typename T>
__global__ void __launch_bounds__(256) k_abba_partial(const T *__restrict__ frames,
                                                      const int8_t *__restrict__ labels, int nframes,
                                                      size_t npix, double *__restrict__ sum_a,
                                                      double *__restrict__ sum_b, int accumulate) {
    constexpr int U = 8;                                       // frames in flight per thread
    extern __shared__ int8_t lab_s[];
    for (int f = threadIdx.x; f < nframes; f += blockDim.x) lab_s[f] = labels[f];
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        double a = accumulate ? sum_a[i] : 0.0, b = accumulate ? sum_b[i] : 0.0;
        int f0 = 0;
        for (; f0 + U <= nframes; f0 += U) {                   // branch-free: U loads in flight
            T v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldcs(frames + (size_t)(f0 + u) * npix + i);
#pragma unroll
            for (int u = 0; u < U; ++u) {                      // frame order, one rounding each
                const int lab = lab_s[f0 + u];                 // (adding +0.0 changes nothing)
                a = __dadd_rn(a, lab > 0 ? (double)v[u] : 0.0);
                b = __dadd_rn(b, lab < 0 ? (double)v[u] : 0.0);
            }
        }
        for (; f0 < nframes; ++f0) {
            const int lab = lab_s[f0];
            const double v = (double)__ldcs(frames + (size_t)f0 * npix + i);
            a = __dadd_rn(a, lab > 0 ? v : 0.0);
            b = __dadd_rn(b, lab < 0 ? v : 0.0);
        }
        sum_a[i] = a;
        sum_b[i] = b;
    }
}

__global__ void __launch_bounds__(256) k_abba_finish(const double *__restrict__ sum_a,
                                                     const double *__restrict__ sum_b, double na,
                                                     double nb, size_t npix, float *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const float ma = (float)__ddiv_rn(sum_a[i], na), mb = (float)__ddiv_rn(sum_b[i], nb);
        out[i] = __fsub_rn(ma, mb);
    }
}

// Pre-reduction in front of the path (recipes/spec/abba.py:371 `flow(newimg)`, core/recipe.py:25-61):
// bias subtraction, dark subtraction (numina BiasCorrector / DarkCorrector, restated: image
// minus calibration) and the flat field of pyemir's own FlatFieldCorrector
// (processing/flatfield.py:43,58-59: flat values <= 0 count as 1.0, result = data / flat cast to
// float32).  One pass over the frames, calibrations (L2 resident: 3 x 16.8 MB) shared by all
// frames; every stage rounds to float32 like the correctors' dtype="float32".  numpy's type
// promotion is kept: float32 image with float32 calibrations computes in float32, any float64
// operand makes the stage float64.
// four consecutive elements with the widest loads the type allows (the streaming variants for
// the frames, the cached ones for the calibrations that every frame reuses)
template <bool STREAM>
__device__ __forceinline__ void load4(const float *p, float (&v)[4]) {
    const float4 t = STREAM ? __ldcs(reinterpret_cast<const float4 *>(p))
                            : __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
template <bool STREAM>
__device__ __forceinline__ void load4(const double *p, double (&v)[4]) {
    const double2 a = STREAM ? __ldcs(reinterpret_cast<const double2 *>(p))
                             : __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = STREAM ? __ldcs(reinterpret_cast<const double2 *>(p) + 1)
                             : __ldg(reinterpret_cast<const double2 *>(p) + 1);
    v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
}
template <bool STREAM>
__device__ __forceinline__ void load4(const uint16_t *p, uint16_t (&v)[4]) {
    const ushort4 t = STREAM ? __ldcs(reinterpret_cast<const ushort4 *>(p))
                             : __ldg(reinterpret_cast<const ushort4 *>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}

// VEC = 4: a thread owns four consecutive pixels (16-byte loads and stores, npix % 4 == 0 and
// aligned buffers); VEC = 1: any size and alignment.  Frames are taken four at a time so that
// four independent loads are in flight per thread.
//
// The bad-pixel stage comes first in the flow (BadPixelCorrector, core/correctors.py:23-42,
// core/recipe.py:26-41).  It touches a fraction of a per cent of the pixels, so it does not ride
// in k_prereduce (whose threads would all pay for the registers of the window median): a second
// small kernel, k_bpm_fix, looks at the mask only and, for a masked pixel of a frame, replaces
// what k_prereduce wrote by the value that starts from the median of the unmasked RAW pixels of
// the 5 x 5 window (bpm_window_median, pixel_ops.cuh) and goes through the same bias / dark /
// flat stages.  The frames are still read once (the windows come from L2).
struct BpmArgs {
    const uint8_t *mask;    // [naxis2][naxis1], non-zero = bad; NULL: no bad-pixel stage
    int naxis1, naxis2;
};

// grid: x over groups of 4 pixels (one mask word per thread), y over frames.  A warp looks at
// 128 pixels of the mask; for every masked one among them ALL its lanes work on the fix: lane
// l < 25 holds window pixel (l / 5, l % 5), the rank of every value among the others comes from
// 25 shuffles, the two middle values from two ballots -- about a hundred warp instructions per
// bad pixel instead of a few thousand in one lane of a divergent warp.
template <typename TIn, typename TCal>
__global__ void __launch_bounds__(128) k_bpm_fix(const TIn *__restrict__ in, float *__restrict__ out,
                                                 size_t npix, const TCal *__restrict__ bias,
                                                 const TCal *__restrict__ dark,
                                                 const TCal *__restrict__ flat, const BpmArgs bpm,
                                                 int words_ok) {
    using A = typename std::conditional<sizeof(TIn) == 8 || sizeof(TCal) == 8, double, float>::type;
    // uint16 and float32 pixels compare exactly as float32
    using Key = typename std::conditional<sizeof(TIn) == 8, double, float>::type;
    constexpr int HW = 2, S = 2 * HW + 1;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    uint32_t m4 = 0;
    if (i0 < npix) {
        if (words_ok && i0 + 4 <= npix) {
            m4 = __ldg(reinterpret_cast<const uint32_t *>(bpm.mask + i0));
        } else {
            for (int e = 0; e < 4 && i0 + e < npix; ++e)
                m4 |= (uint32_t)(bpm.mask[i0 + e] != 0) << (8 * e);
        }
    }
    if (__ballot_sync(full, m4 != 0) == 0) return;
    const TIn *frame = in + (size_t)blockIdx.y * npix;
    float *dst = out + (size_t)blockIdx.y * npix;
    const bool hb = bias != nullptr, hd = dark != nullptr, hf = flat != nullptr;
    const int wa = lane / S - HW, wb = lane % S - HW;          // this lane's place in the window
#pragma unroll 1
    for (int e = 0; e < 4; ++e) {
        unsigned todo = __ballot_sync(full, (m4 >> (8 * e)) & 0xffu);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const size_t i = ((size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31) + src) * 4 + e;
            const int col = (int)(i % (size_t)bpm.naxis1), row = (int)(i / (size_t)bpm.naxis1);
            const int r = row + wa, c = col + wb;
            bool ok = lane < S * S && r >= 0 && r < bpm.naxis2 && c >= 0 && c < bpm.naxis1;
            Key val = (Key)INFINITY;
            if (ok) {
                const size_t k = (size_t)r * bpm.naxis1 + c;
                ok = bpm.mask[k] == 0;
                if (ok) val = (Key)frame[k];
            }
            const unsigned good = __ballot_sync(full, ok);
            const bool has_nan = __ballot_sync(full, ok && val != val) != 0;
            val = val != val ? (Key)INFINITY : val;
            int rank = 0;
#pragma unroll
            for (int j = 0; j < S * S; ++j) {
                const Key vj = __shfl_sync(full, val, j);
                rank += (vj < val || (j < lane && vj == val)) ? 1 : 0;
            }
            const int n = __popc(good);
            double med = 0.0;                                  // `fill` when the window is empty
            if (n > 0) {
                const unsigned in_win = (1u << (S * S)) - 1u;
                const unsigned at_lo = __ballot_sync(full, rank == ((n - 1) >> 1)) & in_win;
                const unsigned at_hi = __ballot_sync(full, rank == (n >> 1)) & in_win;
                const double lo = (double)__shfl_sync(full, val, __ffs(at_lo) - 1);
                const double hi = (double)__shfl_sync(full, val, __ffs(at_hi) - 1);
                med = (n & 1) ? lo : mul_rn(add_rn(lo, hi), 0.5);
                if (has_nan) med = quiet_nan();
            }
            if (lane == 0) {
                const A b = hb ? (A)bias[i] : A(0), d = hd ? (A)dark[i] : A(0);
                A f = hf ? (A)flat[i] : A(1);
                f = f <= A(0) ? A(1) : f;
                dst[i] = prereduce_from<TIn, A>(med, hb, hd, hf, b, d, f);
            }
        }
    }
}

template <typename TIn, typename TCal, int VEC>
__global__ void __launch_bounds__(256) k_prereduce(const TIn *__restrict__ in, float *__restrict__ out,
                                                   int nframes, size_t npix,
                                                   const TCal *__restrict__ bias,
                                                   const TCal *__restrict__ dark,
                                                   const TCal *__restrict__ flat) {
    // float32 unless an operand is float64 (uint16 raw frames promote like numpy: with
    // float32 calibrations the arithmetic is float32)
    using A = typename std::conditional<sizeof(TIn) == 8 || sizeof(TCal) == 8, double, float>::type;
    constexpr int U = 4;
    const bool hb = bias != nullptr, hd = dark != nullptr, hf = flat != nullptr;
    const size_t stride = (size_t)gridDim.x * blockDim.x * VEC;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < npix; i += stride) {
        A b[VEC], d[VEC], f[VEC];
        if constexpr (VEC == 4) {
            TCal t[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) b[e] = d[e] = A(0), f[e] = A(1);
            if (hb) {
                load4<false>(bias + i, t);
#pragma unroll
                for (int e = 0; e < 4; ++e) b[e] = (A)t[e];
            }
            if (hd) {
                load4<false>(dark + i, t);
#pragma unroll
                for (int e = 0; e < 4; ++e) d[e] = (A)t[e];
            }
            if (hf) {
                load4<false>(flat + i, t);
#pragma unroll
                for (int e = 0; e < 4; ++e) f[e] = (A)t[e];
            }
        } else {
            b[0] = hb ? (A)bias[i] : A(0);
            d[0] = hd ? (A)dark[i] : A(0);
            f[0] = hf ? (A)flat[i] : A(1);
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            f[e] = f[e] <= A(0) ? A(1) : f[e];   // flatfield.py:43 (NaN stays NaN, like numpy)
        int n = 0;
        for (; n + U <= nframes; n += U) {
            TIn raw[U][VEC];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if constexpr (VEC == 4) load4<true>(in + (size_t)(n + u) * npix + i, raw[u]);
                else raw[u][0] = __ldcs(in + (size_t)(n + u) * npix + i);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float r[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    r[e] = prereduce_px<TIn, A>(raw[u][e], hb, hd, hf, b[e], d[e], f[e]);
                float *dst = out + (size_t)(n + u) * npix + i;
                if constexpr (VEC == 4) __stcs(reinterpret_cast<float4 *>(dst), make_float4(r[0], r[1], r[2], r[3]));
                else __stcs(dst, r[0]);
            }
        }
        for (; n < nframes; ++n) {
            TIn raw[VEC];
            if constexpr (VEC == 4) load4<true>(in + (size_t)n * npix + i, raw);
            else raw[0] = __ldcs(in + (size_t)n * npix + i);
            float r[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                r[e] = prereduce_px<TIn, A>(raw[e], hb, hd, hf, b[e], d[e], f[e]);
            float *dst = out + (size_t)n * npix + i;
            if constexpr (VEC == 4) __stcs(reinterpret_cast<float4 *>(dst), make_float4(r[0], r[1], r[2], r[3]));
            else __stcs(dst, r[0]);
        }
    }
}

template <typename TIn, typename TCal>
void launch_prereduce(const void *in, float *out, int nframes, size_t npix, const void *bias,
                      const void *dark, const void *flat, int sms, cudaStream_t st,
                      const BpmArgs bpm = BpmArgs{nullptr, 0, 0}) {
    auto aligned = [](const void *ptr, size_t a) {
        return ptr == nullptr || (reinterpret_cast<uintptr_t>(ptr) & (a - 1)) == 0;
    };
    const bool vec = npix % 4 == 0 && aligned(in, 4 * sizeof(TIn) > 16 ? 16 : 4 * sizeof(TIn)) &&
                     aligned(out, 16) && aligned(bias, 16) && aligned(dark, 16) &&
                     aligned(flat, 16);
    const TIn *tin = static_cast<const TIn *>(in);
    const TCal *tb = static_cast<const TCal *>(bias), *td = static_cast<const TCal *>(dark),
               *tf = static_cast<const TCal *>(flat);
    if (vec) {
        const int grid = (int)std::min<size_t>((npix / 4 + 255) / 256, (size_t)sms * 16);
        k_prereduce<TIn, TCal, 4><<<grid, 256, 0, st>>>(tin, out, nframes, npix, tb, td, tf);
    } else {
        const int grid = (int)std::min<size_t>((npix + 255) / 256, (size_t)sms * 16);
        k_prereduce<TIn, TCal, 1><<<grid, 256, 0, st>>>(tin, out, nframes, npix, tb, td, tf);
    }
    if (bpm.mask != nullptr) {
        const dim3 grid((unsigned)((npix / 4 + 128) / 128), (unsigned)nframes);
        k_bpm_fix<TIn, TCal><<<grid, 128, 0, st>>>(tin, out, npix, tb, td, tf, bpm,
                                                  (int)aligned(bpm.mask, 4));
    }
}

// ABBA combination, method "median" (abba.py:579-601 with numina's combine.median, restated):
// per pixel the median over the selected frames of a stack (numpy.median semantics: mean of
// the two middle values, NaN if any sample is NaN), float64 result like numina's combine.
// One thread per pixel (or four): its samples sit in registers (N = stack size rounded up to a
// power of two, padded with -inf / +inf) and go through a selection network, fully unrolled;
// all N loads of a thread are in flight together.
constexpr int STACK_MAX = 64;
struct StackSel {
    int n;
    int idx[STACK_MAX];
};

// (median_of_registers, the padding rule and the selection networks: pixel_ops.cuh)

// medians of VEC consecutive pixels starting at i over the selected images (VEC = 4: 16-byte
// loads, npix % 4 == 0 and aligned frames)
template <typename T, int N, int VEC>
__device__ __forceinline__ void stack_median_px(const T *__restrict__ frames, const StackSel &sel,
                                                size_t npix, size_t i, double (&med)[VEC]) {
    const T inf = (T)INFINITY;
    const int nneg = (N - sel.n) >> 1;                          // pads below the samples
    T v[VEC][N];
#pragma unroll
    for (int u = 0; u < N; ++u) {
        if (u < sel.n) {
            const T *src = frames + (size_t)sel.idx[u] * npix + i;
            if constexpr (VEC == 4) {
                T t[4];
                load4<true>(src, t);
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e][u] = t[e];
            } else {
                v[0][u] = __ldcs(src);
            }
        } else {
            const T pad = u < sel.n + nneg ? -inf : inf;        // (= stack_pad, pixel_ops.cuh)
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[e][u] = pad;
        }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) med[e] = median_of_registers<T, N>(v[e], sel.n);
}

template <typename T, int N, int VEC>
__global__ void __launch_bounds__(256) k_stack_median(const T *__restrict__ frames,
                                                      const StackSel sel, size_t npix,
                                                      double *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * VEC;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < npix; i += stride) {
        double med[VEC];
        stack_median_px<T, N, VEC>(frames, sel, npix, i, med);
#pragma unroll
        for (int e = 0; e < VEC; ++e) out[i + e] = med[e];
    }
}

// median(A images) and median(B images) of a pixel one after the other in the same registers,
// then float32(A) - float32(B) like abba.py:604-606: the stack is read once and only the
// float32 difference is written.
template <typename T, int N, int VEC>
__global__ void __launch_bounds__(256) k_abba_median(const T *__restrict__ frames,
                                                     const StackSel sel_a, const StackSel sel_b,
                                                     size_t npix, float *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * VEC;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < npix; i += stride) {
        double ma[VEC], mb[VEC];
        stack_median_px<T, N, VEC>(frames, sel_a, npix, i, ma);
        stack_median_px<T, N, VEC>(frames, sel_b, npix, i, mb);
        float r[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) r[e] = __fsub_rn((float)ma[e], (float)mb[e]);
        if constexpr (VEC == 4)
            __stcs(reinterpret_cast<float4 *>(out + i), make_float4(r[0], r[1], r[2], r[3]));
        else
            out[i] = r[0];
    }
}

// four pixels per thread while the samples of four pixels fit comfortably in registers
template <typename T, int N>
constexpr bool stack_vec4() { return N * sizeof(T) <= 64; }

template <typename T, int N>
void launch_abba_median_n(const T *frames, const StackSel &sa, const StackSel &sb, size_t npix,
                          float *out, int sms, bool vec_ok, cudaStream_t st) {
    if constexpr (stack_vec4<T, N>()) {
        if (vec_ok) {
            const int grid = (int)std::min<size_t>((npix / 4 + 255) / 256, (size_t)sms * 8);
            k_abba_median<T, N, 4><<<grid, 256, 0, st>>>(frames, sa, sb, npix, out);
            return;
        }
    }
    const int grid = (int)std::min<size_t>((npix + 255) / 256, (size_t)sms * 8);
    k_abba_median<T, N, 1><<<grid, 256, 0, st>>>(frames, sa, sb, npix, out);
}

template <typename T, int N>
void launch_stack_median_n(const T *frames, const StackSel &sel, size_t npix, double *out,
                           int sms, bool vec_ok, cudaStream_t st) {
    if constexpr (stack_vec4<T, N>()) {
        if (vec_ok) {
            const int grid = (int)std::min<size_t>((npix / 4 + 255) / 256, (size_t)sms * 8);
            k_stack_median<T, N, 4><<<grid, 256, 0, st>>>(frames, sel, npix, out);
            return;
        }
    }
    const int grid = (int)std::min<size_t>((npix + 255) / 256, (size_t)sms * 8);
    k_stack_median<T, N, 1><<<grid, 256, 0, st>>>(frames, sel, npix, out);
}

int fill_stack_sel(StackSel &sel, const int32_t *h_index, int nsel, int nframes) {
    memset(&sel, 0, sizeof(sel));
    if (h_index == nullptr) return fail(EMIRWV_E_VALUE, "null argument");
    if (nsel < 1) return fail(EMIRWV_E_VALUE, "no image selected");
    if (nsel > STACK_MAX)
        return fail(EMIRWV_E_LIMIT, "median of more than %d images per call", STACK_MAX);
    sel.n = nsel;
    for (int u = 0; u < nsel; ++u) {
        if (h_index[u] < 0 || h_index[u] >= nframes)
            return fail(EMIRWV_E_VALUE, "image index %d outside the stack", h_index[u]);
        sel.idx[u] = h_index[u];
    }
    return EMIRWV_OK;
}

template <typename T>
void launch_abba_median(const T *frames, const StackSel &sa, const StackSel &sb, size_t npix,
                        float *out, int sms, cudaStream_t st) {
    const int n = std::max(sa.n, sb.n);
    const bool vec_ok = npix % 4 == 0 && (reinterpret_cast<uintptr_t>(frames) & 15u) == 0 &&
                        (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    if (n <= 4) launch_abba_median_n<T, 4>(frames, sa, sb, npix, out, sms, vec_ok, st);
    else if (n <= 8) launch_abba_median_n<T, 8>(frames, sa, sb, npix, out, sms, vec_ok, st);
    else if (n <= 16) launch_abba_median_n<T, 16>(frames, sa, sb, npix, out, sms, vec_ok, st);
    else if (n <= 32) launch_abba_median_n<T, 32>(frames, sa, sb, npix, out, sms, vec_ok, st);
    else launch_abba_median_n<T, 64>(frames, sa, sb, npix, out, sms, vec_ok, st);
}

template <typename T>
void launch_stack_median(const T *frames, const StackSel &sel, size_t npix, double *out, int sms,
                         cudaStream_t st) {
    const bool vec_ok = npix % 4 == 0 && (reinterpret_cast<uintptr_t>(frames) & 15u) == 0 &&
                        (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    if (sel.n <= 4) launch_stack_median_n<T, 4>(frames, sel, npix, out, sms, vec_ok, st);
    else if (sel.n <= 8) launch_stack_median_n<T, 8>(frames, sel, npix, out, sms, vec_ok, st);
    else if (sel.n <= 16) launch_stack_median_n<T, 16>(frames, sel, npix, out, sms, vec_ok, st);
    else if (sel.n <= 32) launch_stack_median_n<T, 32>(frames, sel, npix, out, sms, vec_ok, st);
    else launch_stack_median_n<T, 64>(frames, sel, npix, out, sms, vec_ok, st);
}

// ---------------------------------------------------------------------------------
// ABBA combination, method "sigmaclip" (abba.py:132-137,579-601 with numina's
// combine.sigmaclip, restated): per pixel the sigma-clipped mean of the selected frames of a
// stack (sigmaclip_of_registers, pixel_ops.cuh), the samples of a pixel in registers, all N
// loads of a thread in flight together; the stack is read once.
template <typename T, int N>
__device__ __forceinline__ double stack_sigmaclip_px(const T *__restrict__ frames,
                                                     const StackSel &sel, size_t npix, size_t i,
                                                     double low, double high) {
    T v[N];
#pragma unroll
    for (int u = 0; u < N; ++u) v[u] = u < sel.n ? __ldcs(frames + (size_t)sel.idx[u] * npix + i) : T(0);
    return sigmaclip_of_registers<T, N>(v, sel.n, low, high);
}

template <typename T, int N>
__global__ void __launch_bounds__(256) k_stack_sigmaclip(const T *__restrict__ frames,
                                                         const StackSel sel, size_t npix,
                                                         double low, double high,
                                                         double *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride)
        out[i] = stack_sigmaclip_px<T, N>(frames, sel, npix, i, low, high);
}

// float32(clipped mean of the A images) - float32(clipped mean of the B images), abba.py:604-606
template <typename T, int N>
__global__ void __launch_bounds__(256) k_abba_sigmaclip(const T *__restrict__ frames,
                                                        const StackSel sel_a, const StackSel sel_b,
                                                        size_t npix, double low, double high,
                                                        float *__restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const double ma = stack_sigmaclip_px<T, N>(frames, sel_a, npix, i, low, high);
        const double mb = stack_sigmaclip_px<T, N>(frames, sel_b, npix, i, low, high);
        __stcs(out + i, __fsub_rn((float)ma, (float)mb));
    }
}

template <typename T>
void launch_sigmaclip(const T *frames, const StackSel &sa, const StackSel *sb, size_t npix,
                      double low, double high, void *out, int sms, cudaStream_t st) {
    const int n = sb != nullptr ? std::max(sa.n, sb->n) : sa.n;
    const int grid = (int)std::min<size_t>((npix + 255) / 256, (size_t)sms * 8);
#define SIGCLIP_BUCKET(N)                                                                      \
    do {                                                                                       \
        if (sb != nullptr)                                                                     \
            k_abba_sigmaclip<T, N><<<grid, 256, 0, st>>>(frames, sa, *sb, npix, low, high,     \
                                                         static_cast<float *>(out));           \
        else                                                                                   \
            k_stack_sigmaclip<T, N><<<grid, 256, 0, st>>>(frames, sa, npix, low, high,         \
                                                          static_cast<double *>(out));         \
    } while (0)
    if (n <= 4) SIGCLIP_BUCKET(4);
    else if (n <= 8) SIGCLIP_BUCKET(8);
    else if (n <= 16) SIGCLIP_BUCKET(16);
    else if (n <= 32) SIGCLIP_BUCKET(32);
    else SIGCLIP_BUCKET(64);
#undef SIGCLIP_BUCKET
}

// ---------------------------------------------------------------------------------
// The per-frame vertical offset of the ABBA recipe (abba.py:545-558:
// `shift_image2d(data, yoffset=-offset).astype("float32")` for every frame whose offset is not
// zero), all frames of a stack in one launch, device resident.  A thread owns one column of a
// strip of SHIFT_STRIP output rows of one frame; an output pixel is the sum over the (at most
// three, normally two) source rows it overlaps of area x value in float64, the areas with the
// bits of the general clipping (shift_row / shift_area, pixel_ops.cuh); consecutive output rows
// share a source row, which the thread keeps in a register.  HBM bound: the stack is read once
// and written once.
constexpr int SHIFT_STRIP = 8;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) k_shift_rows(const TIn *__restrict__ in, TOut *__restrict__ out,
                                                    int naxis2, int naxis1,
                                                    const double *__restrict__ yoffset) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int ys = blockIdx.y * SHIFT_STRIP;
    const int f = blockIdx.z;
    if (x >= naxis1) return;
    const size_t fo = (size_t)f * naxis2 * naxis1;
    const TIn *img = in + fo;
    TOut *dst = out + fo;
    const double yo = yoffset[f];
    if (yo == 0.0) {                                   // abba.py:554 `if offset != 0`
        for (int y = ys; y < min(ys + SHIFT_STRIP, naxis2); ++y)
            __stcs(dst + (size_t)y * naxis1 + x, (TOut)__ldcs(img + (size_t)y * naxis1 + x));
        return;
    }
    const double xa = (double)x - 0.5, xb = (double)x + 0.5;
    int have = INT_MIN;                                // source row held in `val`
    double val = 0.0;
    for (int y = ys; y < min(ys + SHIFT_STRIP, naxis2); ++y) {
        const ShiftRow r = shift_row(yo, y);
        double acc = 0.0;
        for (int ii = max(r.lo, 0); ii <= min(r.hi, naxis2 - 1); ++ii) {
            const double area = shift_area(xa, xb, r, ii);
            if (ii != have) {
                val = (double)__ldg(img + (size_t)ii * naxis1 + x);
                have = ii;
            }
            if (area != 0.0) acc = add_rn(acc, mul_rn(area, val));
        }
        __stcs(dst + (size_t)y * naxis1 + x, (TOut)acc);
    }
}
