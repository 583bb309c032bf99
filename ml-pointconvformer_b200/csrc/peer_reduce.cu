// Small all-reduce (sum) over NVLink / NVSwitch peer memory for the SyncBatchNorm statistics exchange (sm_100a).
//
// SyncBatchNorm (the reference trains with sync_bn: True, configs/configPCF_Opt_10cm.yaml; torch.nn.SyncBatchNorm ->
// /root/reference/train_ScanNet_DDP_WarmUP.py:190-195) makes ~540 all-reduces of a few hundred bytes per training step: pure
// latency.  An NCCL all-reduce inside the captured step costs ~14 us each at 2 GPUs (7.6 ms of a 45 ms step).  Here the
// exchange is ONE single-CTA kernel: every rank PUSHES its <= 2048 floats into a slot of every peer's symmetric buffer
// (plain stores through the peer mapping), publishes a release flag carrying the call's epoch, spins on its own flags
// (acquire) and then sums the slots in rank order -- the same order on every rank, so the result is bit-identical
// everywhere and deterministic.  No reset between calls: the epoch (a per-rank device counter) only grows and the data
// slots are double buffered by epoch parity (a peer cannot reach call e+2 before this rank has published call e+1, i.e.
// finished reading call e).
//
// Symmetric buffer layout per rank (bytes): [0,256) flags[world] (uint32, written by the peers) | [256,260) epoch counter |
// [1024, ...) data[2][world][PR_MAXN] floats.
#include "common.cuh"

namespace pcfb {

constexpr int PR_MAXN = 2048;
constexpr int PR_MAXW = 16;
constexpr int PR_DATA_OFF = 1024;

__device__ __forceinline__ void pr_st_release(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t pr_ld_acquire(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float pr_ld_cv(const float *p) {        // never from L1: the slot is written by the peers
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];\n" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pr_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const float *__restrict__ in, float *__restrict__ out, int n, const unsigned long long *__restrict__ bases,
                      int rank, int world)
{
    __shared__ uint32_t ep_s;
    __shared__ unsigned char *base_s[PR_MAXW];
    const int tid = threadIdx.x;
    if (tid < world) base_s[tid] = reinterpret_cast<unsigned char *>(bases[tid]);
    if (tid == 0) {
        uint32_t *ctr = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(bases[rank]) + 256);
        ep_s = *ctr + 1;
        *ctr = ep_s;
    }
    __syncthreads();
    const uint32_t ep = ep_s;
    const size_t slot_off = PR_DATA_OFF + ((size_t)(ep & 1u) * world + rank) * PR_MAXN * sizeof(float);
    for (int i = tid; i < n; i += blockDim.x) {                     // push: my values into my slot of every rank's buffer
        const float v = in[i];
        for (int p = 0; p < world; ++p) reinterpret_cast<float *>(base_s[p] + slot_off)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) {
        pr_st_release(reinterpret_cast<uint32_t *>(base_s[tid]) + rank, ep);          // publish to rank `tid`
        const uint32_t *mine = reinterpret_cast<const uint32_t *>(base_s[rank]) + tid; // and wait for rank `tid`'s data
        const unsigned long long t0 = pr_now_ns();
        while ((int32_t)(pr_ld_acquire(mine) - ep) < 0) {
            if (pr_now_ns() - t0 > 20000000000ull) {                 // 20 s: a rank is missing from the collective
                printf("pcfb peer_allreduce: rank %d timed out waiting for rank %d (epoch %u)\n", rank, tid, ep);
                __trap();
            }
        }
    }
    __syncthreads();
    const float *data = reinterpret_cast<const float *>(base_s[rank] + PR_DATA_OFF + (size_t)(ep & 1u) * world * PR_MAXN * sizeof(float));
    for (int i = tid; i < n; i += blockDim.x) {
        float s = 0.f;
        for (int q = 0; q < world; ++q) s += pr_ld_cv(data + (size_t)q * PR_MAXN + i);   // rank order: identical on every rank
        out[i] = s;
    }
}

}  // namespace pcfb

using namespace pcfb;

extern "C" size_t pcfb_peer_buffer_bytes(int world)
{
    if (world < 1 || world > PR_MAXW) return 0;
    return PR_DATA_OFF + (size_t)2 * world * PR_MAXN * sizeof(float);
}

extern "C" int pcfb_peer_max_floats(void) { return PR_MAXN; }

extern "C" int pcfb_peer_allreduce(const float *in, float *out, int n, const void *peer_bases, int rank, int world, void *stream)
{
    PCFB_REQUIRE(in && out && peer_bases, "pcfb_peer_allreduce: null pointer");
    PCFB_REQUIRE(n >= 0 && n <= PR_MAXN, "pcfb_peer_allreduce: n = %d outside [0, %d]", n, PR_MAXN);
    PCFB_REQUIRE(world >= 1 && world <= PR_MAXW && rank >= 0 && rank < world, "pcfb_peer_allreduce: bad rank %d / world %d", rank, world);
    peer_allreduce_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out, n, static_cast<const unsigned long long *>(peer_bases), rank, world);
    return check_launch("peer_allreduce_kernel");
}
