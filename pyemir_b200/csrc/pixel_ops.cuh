// Per-pixel arithmetic of the kernels either side of the hot path, shared by the CUDA kernels
// (side_kernels.cuh) and by the host-side self check (hostcheck.cpp), so that the CPU test
// suite runs the very functions the GPU runs: the stages of the pre-reduction and the
// selection networks of the stack medians.
#pragma once

#include <math.h>
#include <stdint.h>

#include "geom.cuh"

namespace emirwv {

EMIRWV_HD double quiet_nan() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return (double)NAN;
#endif
}

// compare-exchange of a selection network: two min/max instructions, no branch
template <typename T>
EMIRWV_HD void order2(T &a, T &b) {
    const T lo = fmin(a, b), hi = fmax(a, b);
    a = lo;
    b = hi;
}

// One pixel of one frame through the stages of the pre-reduction (bias, dark, flat), each
// rounded to float32; A is the arithmetic type numpy's promotion gives the stage.
template <typename TIn, typename A>
EMIRWV_HD float prereduce_px(TIn raw, bool hb, bool hd, bool hf, A b, A d, A f) {
    if constexpr (sizeof(TIn) == 8) {
        double x = raw;
        if (hb) x = (double)(float)(x - (double)b);
        if (hd) x = (double)(float)(x - (double)d);
        if (hf) x = (double)(float)(x / (double)f);
        return (float)x;
    } else {
        float x = (float)raw;
        if (hb) x = (float)((A)x - b);
        if (hd) x = (float)((A)x - d);
        if (hf) x = (float)((A)x / f);
        return x;
    }
}

// the same stages from a float64 starting value (the window median of the bad-pixel stage:
// exact in float32 for uint16 and float32 frames, kept in float64 for float64 frames)
template <typename TIn, typename A>
EMIRWV_HD float prereduce_from(double x0, bool hb, bool hd, bool hf, A b, A d, A f) {
    if constexpr (sizeof(TIn) == 8) {
        double x = x0;
        if (hb) x = (double)(float)(x - (double)b);
        if (hd) x = (double)(float)(x - (double)d);
        if (hf) x = (double)(float)(x / (double)f);
        return (float)x;
    } else {
        float x = (float)x0;
        if (hb) x = (float)((A)x - b);
        if (hd) x = (float)((A)x - d);
        if (hf) x = (float)((A)x / f);
        return x;
    }
}

#include "stack_median_nets.inc"

// padding of a stack of n samples held in N registers: floor((N-n)/2) times -inf right after
// the samples, +inf above
template <typename T>
EMIRWV_HD T stack_pad(int u, int n, int N) {
    const T inf = (T)INFINITY;
    return u < n + ((N - n) >> 1) ? -inf : inf;
}

// Median of the n samples held in v[0..n), float64, NaN if any sample is NaN.  The padding
// v[n..N) (stack_pad) puts the median of the samples on the wires N/2-1 and N/2 for every n
// (odd n: wire N/2-1 alone), so the network only has to bring the two middle values of N
// wires home: Batcher's odd-even merge sort pruned to those two outputs
// (scripts/gen_stack_median_nets.py), two min/max instructions per exchange, every index a
// compile-time constant: the samples never leave the registers.
template <typename T, int N>
EMIRWV_HD double median_of_registers(T (&v)[N], int n) {
    const T inf = (T)INFINITY;
    bool has_nan = false;
#pragma unroll
    for (int u = 0; u < N; ++u) {
        has_nan |= v[u] != v[u];
        v[u] = v[u] != v[u] ? inf : v[u];
    }
#define CE(a, b) order2(v[a], v[b]);
    if constexpr (N == 4) {
        STACK_NET_4(CE)
    } else if constexpr (N == 8) {
        STACK_NET_8(CE)
    } else if constexpr (N == 16) {
        STACK_NET_16(CE)
    } else if constexpr (N == 32) {
        STACK_NET_32(CE)
    } else {
        static_assert(N == 4 || N == 8 || N == 16 || N == 32 || N == 64, "stack size bucket");
        STACK_NET_64(CE)
    }
#undef CE
    const T a = v[N / 2 - 1], b = (n & 1) ? a : v[N / 2];
    const double med = mul_rn(add_rn((double)a, (double)b), 0.5);
    return has_nan ? quiet_nan() : med;
}

}  // namespace emirwv

// ------------------------------------------------------------------------------------------
// The pieces behind the ABBA combination beyond mean and median (recipes/spec/abba.py:545-610)
// and the bad-pixel stage of the pre-reduction (core/correctors.py:23-42).
// ------------------------------------------------------------------------------------------
namespace emirwv {

EMIRWV_HD double sqrt_rn(double a) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(a);
#else
    volatile double r = sqrt(a);
    return r;
#endif
}

EMIRWV_HD int popcount64(unsigned long long m) {
#if defined(__CUDA_ARCH__)
    return __popcll(m);
#else
    return __builtin_popcountll(m);
#endif
}

// Sigma-clipped mean of the n samples held in v[0..n) (numina combine.sigmaclip, restated:
// abba.py:132-137 `method`, used at :579-601).  Repeat { mean and unbiased standard deviation
// of the samples still in; throw out what lies outside [mean - low*std, mean + high*std] }
// until nothing is thrown out (or nothing would be left); the mean of the last pass is the
// result.  float64, sums in frame order over the samples still in, one rounding per
// operation -- the same sequence as the numpy statement in oracle/pipeline.py.
template <typename T, int N>
EMIRWV_HD double sigmaclip_of_registers(const T (&v)[N], int n, double low, double high) {
    unsigned long long keep = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
    double mean = 0.0;
    for (;;) {
        const int cnt = popcount64(keep);
        // (a dropped sample enters as +0.0: the select sits in front of the addition, off the
        // chain of dependent additions, and changes nothing but the sign of a zero sum)
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < N; ++u) s = add_rn(s, ((keep >> u) & 1ull) ? (double)v[u] : 0.0);
        mean = div_rn(s, (double)cnt);
        double ss = 0.0;
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const double d = sub_rn((double)v[u], mean);
            ss = add_rn(ss, ((keep >> u) & 1ull) ? mul_rn(d, d) : 0.0);
        }
        const double sd = sqrt_rn(cnt > 1 ? div_rn(ss, (double)(cnt - 1)) : 0.0);
        const double lo = sub_rn(mean, mul_rn(sd, low)), hi = add_rn(mean, mul_rn(sd, high));
        unsigned long long next = 0ull;
#pragma unroll
        for (int u = 0; u < N; ++u)
            if (((keep >> u) & 1ull) && (double)v[u] >= lo && (double)v[u] <= hi) next |= 1ull << u;
        const int ncnt = popcount64(next);
        if (ncnt == cnt || ncnt == 0) break;
        keep = next;
    }
    return mean;
}

// shift_image2d(data, yoffset=...) of the ABBA recipe (abba.py:556; numina: rectify2d with the
// translation map aij = [-0, 1, 0], bij = [-yoffset, 0, 1], area-overlap resampling).  With no
// shift along x the mapped corners of output pixel (x, y) are (x -+ .5, v0 | v1) with
//     v0 = fl(b0 + (y - .5)),  v1 = fl(b0 + (y + .5)),  b0 = -yoffset
// and the clipping against source pixel (x, ii) leaves the rectangle [x-.5, x+.5] x [ya, yb],
// ya = max(v0, ii-.5), yb = min(v1, ii+.5), always as the vertex cycle (xa,ya) (xa,yb) (xb,yb)
// (xb,ya): the shoelace sum below is that of quad_box_area for this cycle, operation by
// operation, so the weight has the same bits as the general clipping -- without clipping.
struct ShiftRow {
    double v0, v1;
    int lo, hi;          // source rows with a non-degenerate overlap
};

EMIRWV_HD ShiftRow shift_row(double yoffset, int y) {
    ShiftRow r;
    const double b0 = add_rn(0.0, -yoffset);           // fmap's accumulator starts at 0.0
    r.v0 = add_rn(b0, (double)y - 0.5);
    r.v1 = add_rn(b0, (double)y + 0.5);
    r.lo = (int)floor(r.v0 + 0.5);                     // quad_extent
    r.hi = (int)ceil(r.v1 - 0.5);
    if (r.hi < r.lo) r.hi = r.lo;
    return r;
}

EMIRWV_HD double shift_area(double xa, double xb, const ShiftRow &r, int ii) {
    const double y0 = (double)ii - 0.5, y1 = (double)ii + 0.5;
    const double ya = r.v0 >= y0 ? r.v0 : y0;
    const double yb = r.v1 <= y1 ? r.v1 : y1;
    if (!(yb > ya)) return 0.0;
    const double pa = mul_rn(xa, ya), pb = mul_rn(xa, yb), pc = mul_rn(xb, ya), pd = mul_rn(xb, yb);
    double twice = sub_rn(pb, pa);
    twice = add_rn(twice, sub_rn(pb, pd));
    twice = add_rn(twice, sub_rn(pc, pd));
    twice = add_rn(twice, sub_rn(pc, pa));
    return 0.5 * fabs(twice);
}

// Bad-pixel stage of the flow (numina BadPixelCorrector -> array.bpm.process_bpm_median,
// restated; built by get_corrector_p, core/correctors.py:23-42): a masked pixel becomes the
// median of the unmasked pixels of the (2*HW+1)^2 window around it, clipped at the detector
// edges (values of the INPUT image: corrected pixels are not reused), `fill` when the window
// holds none.  numpy.median semantics (mean of the two middle values), float64.
template <typename TIn, int HW>
EMIRWV_HD double bpm_window_median(const TIn *img, const uint8_t *mask, int naxis1, int naxis2,
                                   int col, int row, double fill) {
    // the window in registers (every index below is a compile-time constant once unrolled):
    // pixels that do not count -- masked, or outside the detector -- hold +inf
    constexpr int S = 2 * HW + 1, M = S * S;
    double w[M];
    int n = 0;
    bool has_nan = false;
#pragma unroll
    for (int a = 0; a < S; ++a)
#pragma unroll
        for (int b = 0; b < S; ++b) {
            const int r = row - HW + a, c = col - HW + b;
            double val = (double)INFINITY;
            if (r >= 0 && r < naxis2 && c >= 0 && c < naxis1) {
                const long long i = (long long)r * naxis1 + c;
                if (mask[i] == 0) {
                    val = (double)img[i];
                    has_nan |= val != val;
                    ++n;
                }
            }
            w[a * S + b] = val != val ? (double)INFINITY : val;
        }
    if (n == 0) return fill;
    if (has_nan) return quiet_nan();
    // selection by rank (ties broken by position): no sort, no indexed array
    const int klo = (n - 1) >> 1, khi = n >> 1;
    double lo = 0.0, hi = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        int rank = 0;
#pragma unroll
        for (int j = 0; j < M; ++j)
            rank += (w[j] < w[i] || (j < i && w[j] == w[i])) ? 1 : 0;
        lo = rank == klo ? w[i] : lo;
        hi = rank == khi ? w[i] : hi;
    }
    return (n & 1) ? lo : mul_rn(add_rn(lo, hi), 0.5);
}

}  // namespace emirwv
