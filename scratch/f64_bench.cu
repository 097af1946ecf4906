// which part of "float32 stack -> float64 sums" is slow on sm_100a?
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ double widen(float x) {
    unsigned u = __float_as_uint(x), e = (u >> 23) & 0xffu;
    if (e == 0u || e == 0xffu) return (double)x;
    return __hiloint2double((int)((u & 0x80000000u) | ((e + 896u) << 20) | ((u >> 3) & 0xfffffu)), (int)(u << 29));
}
template <int MODE> __global__ void k(const float *x, int nf, size_t npix, double *out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        double a = 0; float fa = 0;
        for (int f0 = 0; f0 < nf; f0 += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(x + (size_t)(f0 + u) * npix + i);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (MODE == 0) fa += v[u];
                if (MODE == 1) a += (double)v[u];
                if (MODE == 2) a += widen(v[u]);
                if (MODE == 3) fa += (float)(double)v[u];
            }
        }
        out[i] = a + fa;
    }
}
int main() {
    const int nf = 64; const size_t npix = 2090 * 3400;
    float *x; double *o; cudaMalloc(&x, nf * npix * 4); cudaMalloc(&o, npix * 8); cudaMemset(x, 0x3f, nf * npix * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 4; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 32, 256>>>(x, nf, npix, o);
        if (mode == 1) k<1><<<148 * 32, 256>>>(x, nf, npix, o);
        if (mode == 2) k<2><<<148 * 32, 256>>>(x, nf, npix, o);
        if (mode == 3) k<3><<<148 * 32, 256>>>(x, nf, npix, o);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("mode %d (%s): %.3f ms\n", mode, mode == 0 ? "float add" : mode == 1 ? "F2F + DADD" : mode == 2 ? "int widen + DADD" : "F2F both ways + FADD", ms);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
