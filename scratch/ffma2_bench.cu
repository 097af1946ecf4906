// microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c){ u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c){ float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template<int MODE> __global__ void k(float *out, int iters, float s){
    float a[8]; u64 b[8];
    for (int i=0;i<8;++i){ a[i]=threadIdx.x*0.001f+i; float2 t=make_float2(a[i],a[i]+1); b[i]=*reinterpret_cast<u64*>(&t);}
    float2 s2=make_float2(s,s); u64 sp=*reinterpret_cast<u64*>(&s2);
    for (int it=0; it<iters; ++it){
        if (MODE==0){
#pragma unroll
            for (int i=0;i<8;++i) a[i]=ffma1(a[i],s,s);
        } else {
#pragma unroll
            for (int i=0;i<8;++i) b[i]=ffma2(b[i],sp,sp);
        }
    }
    float r=0; for(int i=0;i<8;++i){ r+=a[i]; float2 t=*reinterpret_cast<float2*>(&b[i]); r+=t.x+t.y; }
    out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
int main(){
    float *o; cudaMalloc(&o, 148*8*1024*4);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters=20000;
    for (int mode=0; mode<2; ++mode) for (int rep=0; rep<2; ++rep){
        cudaEventRecord(e0);
        if(mode==0) k<0><<<148*4,512>>>(o,iters,0.999f); else k<1><<<148*4,512>>>(o,iters,0.999f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms,e0,e1);
        double winstr = 148.0*4*16*iters*8;
        printf("mode %d: %.3f ms, %.1f warp-instr/clk/SM (at 1.965 GHz)\n", mode, ms, winstr/(ms*1e-3)/1.965e9/148);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
