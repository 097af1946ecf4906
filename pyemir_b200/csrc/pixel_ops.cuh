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
