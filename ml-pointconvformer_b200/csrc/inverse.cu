// kNN inverse map = CSR transpose of an [n_out, K] edge table (sm_100a).
//
// Replaces pcf_cuda.compute_knn_inverse (/root/reference/cpp_wrappers/cpp_pcf_kernel/src/knn.cu:24-168):
//   reference: 1 CTA per point histogram, a single-thread serial prefix sum (knn.cu:44-56,139), then an
//   atomic slot-claim fill whose order inside a segment is nondeterministic (knn.cu:76-83).
//   here:      (1) grid-stride histogram, (2) single-pass decoupled look-back exclusive scan,
//              (3) atomic slot-claim fill of the packed edge id e = n*K + k into a scratch array,
//              (4) per-segment rank sort of the edge ids -> ascending (n, k): deterministic output,
//                  identical to the stable-sort oracle.
// HBM-bound integer work: 8E (nei, read twice) + 4E scratch w/r + 5E out + 4(N+1)*3 bytes.
#include "common.cuh"
#include "scan.cuh"

namespace pcfb {

__global__ void inv_count_kernel(const int64_t *__restrict__ nei, int64_t n_edges, int total,
                                 int32_t *__restrict__ counts)
{
    pdl_wait();
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = nei[e];
        if (p >= 0 && p < total) atomicAdd(&counts[p], 1);
    }
}

__global__ void inv_fill_kernel(const int64_t *__restrict__ nei, int64_t n_edges, int total,
                                const int32_t *__restrict__ inv_idx, int32_t *__restrict__ cursor,
                                int32_t *__restrict__ scratch)
{
    pdl_wait();
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = nei[e];
        if (p >= 0 && p < total) {
            const int pos = atomicAdd(&cursor[p], 1);
            scratch[inv_idx[p] + pos] = (int32_t)e;
        }
    }
}

// one warp per segment: rank sort of the (distinct) edge ids, then split into (n, k)
__global__ void inv_sort_kernel(const int32_t *__restrict__ scratch, const int32_t *__restrict__ inv_idx,
                                int total, int K, int32_t *__restrict__ inv_neighbors,
                                uint8_t *__restrict__ inv_k)
{
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < total; p += gridDim.x * warps_per_block) {
        const int beg = inv_idx[p], end = inv_idx[p + 1], len = end - beg;
        if (len <= 32) {
            const int mine = lane < len ? scratch[beg + lane] : 0x7fffffff;
            int rank = 0;
#pragma unroll 8
            for (int j = 0; j < 32; ++j) {
                const int other = __shfl_sync(0xffffffffu, mine, j);
                rank += (other < mine) ? 1 : 0;
            }
            if (lane < len) {
                inv_neighbors[beg + rank] = mine / K;
                inv_k[beg + rank] = (uint8_t)(mine % K);
            }
        } else {
            for (int i = lane; i < len; i += 32) {
                const int mine = scratch[beg + i];
                int rank = 0;
                for (int j = 0; j < len; ++j) rank += (scratch[beg + j] < mine) ? 1 : 0;
                inv_neighbors[beg + rank] = mine / K;
                inv_k[beg + rank] = (uint8_t)(mine % K);
            }
        }
    }
}

struct InvWorkspace {
    int32_t *counts;       // [total]  (histogram, then reused as fill cursor)
    int32_t *scratch;      // [n_out*K]
    unsigned long long *state;   // [tiles]
    unsigned int *ticket;  // [1]
    size_t bytes;
};

static InvWorkspace carve_inv(void *ws, int n_out, int K, int total) {
    Carver c(ws);
    InvWorkspace w;
    // zero-initialised region first (counts, state, ticket are contiguous -> one memset)
    w.counts = c.take<int32_t>((size_t)total + 1);
    w.state = c.take<unsigned long long>((size_t)ceil_div(total + 1, SCAN_TILE) + 1);
    w.ticket = c.take<unsigned int>(4);
    w.scratch = c.take<int32_t>((size_t)n_out * K + 1);
    w.bytes = align_up(c.off, 256);
    return w;
}

}  // namespace pcfb

extern "C" size_t pcfb_knn_inverse_workspace(int n_out, int K, int total)
{
    return pcfb::carve_inv(nullptr, n_out, K, total).bytes;
}

extern "C" int pcfb_knn_inverse(const int64_t *nei, int n_out, int K, int total, int32_t *inv_neighbors,
                                uint8_t *inv_k, int32_t *inv_idx, void *workspace, size_t workspace_bytes,
                                void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(K >= 1 && K <= 255, "pcfb_knn_inverse: K=%d outside [1,255] (inv_k is uint8)", K);
    PCFB_REQUIRE(n_out >= 0 && total >= 0, "pcfb_knn_inverse: bad sizes");
    PCFB_REQUIRE((int64_t)n_out * K < (1ll << 31), "pcfb_knn_inverse: n_out*K overflows int32");
    PCFB_REQUIRE(inv_neighbors && inv_k && inv_idx && workspace, "pcfb_knn_inverse: null pointer");
    InvWorkspace w = carve_inv(workspace, n_out, K, total);
    if (workspace_bytes < w.bytes) {
        set_error("pcfb_knn_inverse: workspace %zu < %zu", workspace_bytes, w.bytes);
        return PCFB_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t E = (int64_t)n_out * K;
    const size_t zero_bytes = (size_t)((char *)w.scratch - (char *)w.counts);
    PCFB_CUDA(cudaMemsetAsync(w.counts, 0, zero_bytes, st));
    PCFB_CUDA(cudaMemsetAsync(inv_neighbors, 0, sizeof(int32_t) * (size_t)E, st));
    PCFB_CUDA(cudaMemsetAsync(inv_k, 0, (size_t)E, st));
    int rc;
    const int64_t eb64 = (E + 255) / 256;
    const int eb = (int)(eb64 < (int64_t)kNumSMs * 16 ? eb64 : (int64_t)kNumSMs * 16);
    if (E > 0) {
        launch_k(inv_count_kernel, eb, 256, 0, st, nei, E, total, w.counts);
        if ((rc = check_launch("inv_count_kernel"))) return rc;
    }
    const int tiles = ceil_div(total + 1, SCAN_TILE);
    // scan over total+1 entries (the extra trailing zero yields inv_idx[total] = grand total)
    launch_k(inv_scan_kernel, tiles, SCAN_THREADS, 0, st, w.counts, total, inv_idx, w.state, w.ticket);
    if ((rc = check_launch("inv_scan_kernel"))) return rc;
    if (E > 0) {
        PCFB_CUDA(cudaMemsetAsync(w.counts, 0, sizeof(int32_t) * (size_t)total, st));
        launch_k(inv_fill_kernel, eb, 256, 0, st, nei, E, total, inv_idx, w.counts, w.scratch);
        if ((rc = check_launch("inv_fill_kernel"))) return rc;
        const int sb = min(ceil_div(total, 8), kNumSMs * 8);
        launch_k(inv_sort_kernel, max(sb, 1), 256, 0, st, w.scratch, inv_idx, total, K, inv_neighbors, inv_k);
        if ((rc = check_launch("inv_sort_kernel"))) return rc;
    }
    return PCFB_OK;
}
