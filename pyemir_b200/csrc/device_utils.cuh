// Device helpers of the streaming kernels: 16-byte vectors, TMA bulk copies and mbarriers,
// table loads, packed float32 pairs.
// Included by emirwv.cu inside its anonymous namespace (one translation unit).

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
// ---- TMA bulk copies (cp.async.bulk) completed through an mbarrier transaction count
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared, 16-byte aligned addresses and size; completes `bytes` on the barrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// Orders this thread's earlier generic-proxy accesses to shared memory (ordinary st.shared)
// before later async-proxy accesses (cp.async.bulk writes) to the same locations.
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// shared -> global bulk store (TMA), 16-byte aligned addresses and size; completion is
// tracked per thread in bulk async-groups
__device__ __forceinline__ void bulk_s2g(void *dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"(src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// TMA tensor copy: one box of a 3-D tensor (column, row, frame) -> shared memory; elements of
// the box outside the tensor arrive as zeros; completes the box's bytes on the barrier
__device__ __forceinline__ void tensor_g2s_3d(unsigned dst, const void *tmap, int x, int y, int z,
                                              unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void *p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
__device__ __forceinline__ unsigned smem_addr_of(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
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

// ---- packed float32 pairs (SASS FFMA2 / FADD2 / FMUL2): one issue slot does the same
// IEEE operation for two frames.  Every packed expression below keeps the operation
// order and the roundings of its scalar counterpart, so both paths give the same bits.
__device__ __forceinline__ unsigned long long f2_bits(float2 a) {
    return *reinterpret_cast<unsigned long long *>(&a);
}
__device__ __forceinline__ float2 f2_from(unsigned long long a) {
    return *reinterpret_cast<float2 *>(&a);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return f2_from(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
}
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }

// The G values of one knot (one per frame of the item), moved as one vector.
template <typename T, int G>
struct alignas((sizeof(T) * G >= 16) ? 16 : sizeof(T) * G) KnotVec {
    T v[G];
};

// non-zero test that accumulates with one logic instruction: bits other than the sign
__device__ __forceinline__ unsigned nz_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }
__device__ __forceinline__ unsigned nz_bits(double v) {
    return ((unsigned)__double2hiint(v) & 0x7fffffffu) | (unsigned)__double2loint(v);
}
