// ABBA "mean" combination fused with its exchange over NVLink peer memory (one process per GPU,
// peer buffers opened through CUDA IPC): recipes/spec/abba.py:579-606 for a sequence sharded
// over ranks, without a collective library call on the data path.
// Included by emirwv.cu inside its anonymous namespace (one translation unit).
//
// Pixels are dealt out in `world` slabs; rank j OWNS slab j.
//   k_abba_partial_push : the last pass of a rank over its rectified frames (float64 sums of its
//                         A and B frames, in frame order, on top of the sums of earlier passes)
//                         does not write the sums locally: every sum goes straight into the
//                         receive buffer of the pixel's owner -- stores over NVLink that ride
//                         along with the streaming reads of the frames.
//   k_peer_signal       : (next in the stream: the pushes have landed) raises this rank's flag
//                         in every owner's flag array, system scope.
//   k_peer_wait         : an owner waits until the flags of all ranks show this call's epoch.
//   k_abba_finish_slab  : the owner adds the world contributions of its slab in rank order,
//                         forms float32(mean A) - float32(mean B) (abba.py:604-606) and stores it
//                         straight into the image on the destination rank (peer store), then
//                         raises its slab flag there; the destination waits for the world flags.
// Buffers are used alternately by epoch parity: a rank can start call e+2 only after every rank
// has pushed e+1, i.e. finished reading the buffers of call e.

constexpr int PEER_MAX = 16;

struct PeerPtrs {
    double *recv[PEER_MAX];     // recv[j]: rank j's receive buffer [parity][src][2][slab]
    int *flags[PEER_MAX];       // flags[j]: rank j's flag array [2][PEER_MAX] (pushes), then
                                //           [2][PEER_MAX] (finished slabs)
};

__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <typename T>
__global__ void __launch_bounds__(256) k_abba_partial_push(
    const T *__restrict__ frames, const int8_t *__restrict__ labels, int nframes, size_t npix,
    const double *__restrict__ sum_a, const double *__restrict__ sum_b, int accumulate,
    const PeerPtrs peers, int rank, int world, size_t slab, int parity) {
    constexpr int U = 8;
    extern __shared__ int8_t lab_s[];
    for (int f = threadIdx.x; f < nframes; f += blockDim.x) lab_s[f] = labels[f];
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        double a = accumulate ? sum_a[i] : 0.0, b = accumulate ? sum_b[i] : 0.0;
        int f0 = 0;
        for (; f0 + U <= nframes; f0 += U) {
            T v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldcs(frames + (size_t)(f0 + u) * npix + i);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int lab = lab_s[f0 + u];
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
        const int owner = (int)(i / slab);
        const size_t o = i - (size_t)owner * slab;
        double *dst = peers.recv[owner] + (((size_t)parity * world + rank) * 2) * slab + o;
        dst[0] = a;             // over NVLink when owner != rank
        dst[slab] = b;
    }
}

// flag_base: 0 = "pushed", 2 * PEER_MAX = "slab finished"
__global__ void k_peer_signal(const PeerPtrs peers, int rank, int world, int parity, int epoch,
                              int flag_base, int only) {
    const int j = threadIdx.x;
    if (j >= world || (only >= 0 && j != only)) return;
    __threadfence_system();
    st_release_sys(peers.flags[j] + flag_base + parity * PEER_MAX + rank, epoch);
}

// spins (bounded) until the `world` flags of this parity hold `epoch`; *timeout is set when not
__global__ void k_peer_wait(const int *flags, int world, int epoch, long long max_spins,
                            int *timeout) {
    const int j = threadIdx.x;
    if (j >= world) return;
    long long spins = 0;
    while (ld_acquire_sys(flags + j) != epoch) {
        if (++spins > max_spins) {
            atomicExch(timeout, 1);
            break;
        }
        __nanosleep(200);
    }
}

__global__ void __launch_bounds__(256) k_abba_finish_slab(const double *__restrict__ recv, int world,
                                                          size_t slab, size_t first, size_t npix,
                                                          double na, double nb,
                                                          float *__restrict__ out_root) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n = first < npix ? min(slab, npix - first) : 0;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += stride) {
        double a = 0.0, b = 0.0;
        for (int src = 0; src < world; ++src) {               // rank order
            a = __dadd_rn(a, recv[((size_t)src * 2) * slab + o]);
            b = __dadd_rn(b, recv[((size_t)src * 2 + 1) * slab + o]);
        }
        const float ma = (float)__ddiv_rn(a, na), mb = (float)__ddiv_rn(b, nb);
        out_root[first + o] = __fsub_rn(ma, mb);              // peer store when this is not the root
    }
}
