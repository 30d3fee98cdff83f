// shard_p2p.cu -- the two exchanges of the endpoint-sharded mode over NVLink PEER MEMORY instead of
// NCCL: every rank exposes one buffer {flags, presence masks, best records} through CUDA IPC; the consumers read the
// peers' copies directly with P2P loads and reduce on the fly (bitwise OR of the masks -- a reduction NCCL does not
// have -- and the max / lowest-slot / tie-sum merge of the records).  Cross-rank ordering is two monotonically
// increasing flags per rank and batch half, written with st.release.sys after the producing kernel and polled with
// ld.acquire.sys; everything is stream-ordered -- the two halves of a batch on the engine's two streams, so that a rank
// computes on one half while it waits for its peers on the other -- and the host only waits at the end of the batch.
//
// Buffer reuse needs no second buffer: a rank overwrites its masks of batch k+1 only after its own merge of batch k,
// which waited for every peer's flag2(k), which each peer raises after it has finished reading the masks of batch k;
// and it overwrites its records of batch k+1 only after every peer's flag1(k+1), raised after that peer's merge(k).
#include "kernels.h"

namespace epp {

namespace {
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void k_p2p_signal(unsigned long long *flag, unsigned long long epoch) {
    __threadfence_system();
    st_release_sys(flag, epoch);
}

// The same, unless an earlier wait of this batch failed (*err != 0): records computed from an incomplete OR of the masks
// must never be published.
__global__ void k_p2p_signal_unless(unsigned long long *flag, unsigned long long epoch, const int *err) {
    if (*err) return;
    __threadfence_system();
    st_release_sys(flag, epoch);
}

// One thread per peer: wait until its flag reaches `epoch`.  A peer that never arrives must not hang the GPU: after
// timeout_ns the wait gives up and reports the rank in *err (the host turns it into an error code).
__global__ void k_p2p_wait(unsigned char *const *peers, int n, size_t flag_off, unsigned long long epoch, int *err,
                           unsigned long long timeout_ns) {
    const int g = threadIdx.x;
    if (g >= n) return;
    const unsigned long long *f = reinterpret_cast<const unsigned long long *>(peers[g] + flag_off);
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        const unsigned long long v = ld_acquire_sys(f);
        if (v == ~0ull) {                              // the peer poisoned its flags: its exchange is broken
            atomicExch(err, 64 + g + 1);
            return;
        }
        if (v >= epoch) return;
        __nanosleep(256);
        if (globaltimer_ns() - t0 > timeout_ns) {
            atomicExch(err, g + 1);
            return;
        }
    }
}

// out[i] = OR over the ranks of their mask word i (16 bytes per thread).
__global__ void k_p2p_or_masks(unsigned char *const *peers, int n, size_t masks_off, unsigned long long n_vec,
                               uint4 *out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vec) return;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int g = 0; g < n; g++) {
        const uint4 v = __ldcv(reinterpret_cast<const uint4 *>(peers[g] + masks_off) + i);
        acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
    }
    out[i] = acc;
}

// out[g][:] = rank g's best records (8 bytes per thread).
__global__ void k_p2p_gather(unsigned char *const *peers, int n, size_t best_off, unsigned long long n_u64,
                             unsigned long long *out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_u64 * (unsigned long long)n) return;
    const int g = (int)(i / n_u64);
    const unsigned long long k = i % n_u64;
    out[i] = __ldcv(reinterpret_cast<const unsigned long long *>(peers[g] + best_off) + k);
}
}  // namespace

cudaError_t launch_p2p_signal(unsigned long long *flag, unsigned long long epoch, cudaStream_t s) {
    k_p2p_signal<<<1, 1, 0, s>>>(flag, epoch);
    return cudaGetLastError();
}
cudaError_t launch_p2p_signal_unless(unsigned long long *flag, unsigned long long epoch, const int *err_dev, cudaStream_t s) {
    k_p2p_signal_unless<<<1, 1, 0, s>>>(flag, epoch, err_dev);
    return cudaGetLastError();
}
cudaError_t launch_p2p_wait(unsigned char *const *peers_dev, int n, size_t flag_off, unsigned long long epoch, int *err_dev,
                            unsigned long long timeout_ns, cudaStream_t s) {
    k_p2p_wait<<<1, 64, 0, s>>>(peers_dev, n, flag_off, epoch, err_dev, timeout_ns);
    return cudaGetLastError();
}
cudaError_t launch_p2p_or_masks(unsigned char *const *peers_dev, int n, size_t masks_off, unsigned long long n_bytes,
                                void *out, cudaStream_t s) {
    const unsigned long long n_vec = n_bytes / 16;
    if (!n_vec) return cudaSuccess;
    k_p2p_or_masks<<<(unsigned)((n_vec + 255) / 256), 256, 0, s>>>(peers_dev, n, masks_off, n_vec, reinterpret_cast<uint4 *>(out));
    return cudaGetLastError();
}
cudaError_t launch_p2p_gather(unsigned char *const *peers_dev, int n, size_t best_off, unsigned long long n_bytes,
                              void *out, cudaStream_t s) {
    const unsigned long long n_u64 = n_bytes / 8;
    if (!n_u64) return cudaSuccess;
    const unsigned long long total = n_u64 * (unsigned long long)n;
    k_p2p_gather<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(peers_dev, n, best_off, n_u64, reinterpret_cast<unsigned long long *>(out));
    return cudaGetLastError();
}

}  // namespace epp
