// Neighbour gather family (sm_100a): index_points (/root/reference/layer_utils.py:13-30), its
// backward as a CSR segment sum over the kNN inverse map (no atomics; the reference's torch path
// uses index_put_(accumulate=True)), and the strided-shortcut max-pool gather
// (/root/reference/layers.py:403-408, 728-733) with its backward.
// All HBM-bound: rows are moved as float4 when C % 4 == 0, one sub-warp group per row.
#include "common.cuh"

namespace pcfb {

template <typename VT>
__global__ void gather_kernel(const VT *__restrict__ feats, const int64_t *__restrict__ nei,
                              int n_in, int64_t n_edges, int CV, VT *__restrict__ out)
{
    pdl_wait();
    const int64_t total = n_edges * CV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / CV;
        const int c = (int)(i - e * CV);
        const int64_t p = nei[e];
        VT v{};
        if (p >= 0 && p < n_in) v = feats[p * CV + c];
        out[i] = v;
    }
}

// grad_feats[p, :] = sum_{e in seg(p)} grad_out[(inv_n[e]*K + inv_k[e]), :]
template <typename VT>
__device__ __forceinline__ void vadd(VT &a, const VT &b);
template <> __device__ __forceinline__ void vadd<float>(float &a, const float &b) { a += b; }
template <> __device__ __forceinline__ void vadd<float4>(float4 &a, const float4 &b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

template <typename VT>
__global__ void gather_bwd_kernel(const VT *__restrict__ grad_out, const int32_t *__restrict__ inv_n,
                                  const uint8_t *__restrict__ inv_k, const int32_t *__restrict__ inv_idx,
                                  int n_in, int K, int CV, VT *__restrict__ grad_feats)
{
    pdl_wait();
    const int64_t total = (int64_t)n_in * CV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i / CV);
        const int c = (int)(i - (int64_t)p * CV);
        const int beg = inv_idx[p], end = inv_idx[p + 1];
        VT acc{};
        for (int e = beg; e < end; ++e) {
            const int64_t row = (int64_t)inv_n[e] * K + inv_k[e];
            vadd(acc, grad_out[row * CV + c]);
        }
        grad_feats[i] = acc;
    }
}

__global__ void gather_max_kernel(const float *__restrict__ feats, const int64_t *__restrict__ nei,
                                  int n_in, int n_out, int K, int C, float *__restrict__ out,
                                  uint8_t *__restrict__ arg)
{
    pdl_wait();
    const int64_t total = (int64_t)n_out * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / C);
        const int c = (int)(i - (int64_t)m * C);
        float best = -__int_as_float(0x7f800000);
        int bk = 0;
        for (int k = 0; k < K; ++k) {
            const int64_t p = nei[(int64_t)m * K + k];
            if (p < 0 || p >= n_in) continue;
            const float v = feats[p * C + c];
            if (v > best) { best = v; bk = k; }          // strict: first maximum wins on ties
        }
        out[i] = best;
        if (arg) arg[i] = (uint8_t)bk;
    }
}

__global__ void gather_max_bwd_kernel(const float *__restrict__ grad_out, const uint8_t *__restrict__ arg,
                                      const int32_t *__restrict__ inv_n, const uint8_t *__restrict__ inv_k,
                                      const int32_t *__restrict__ inv_idx, int n_in, int C,
                                      float *__restrict__ grad_feats)
{
    pdl_wait();
    const int64_t total = (int64_t)n_in * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i / C);
        const int c = (int)(i - (int64_t)p * C);
        float acc = 0.f;
        for (int e = inv_idx[p]; e < inv_idx[p + 1]; ++e) {
            const int64_t o = (int64_t)inv_n[e] * C + c;
            if (arg[o] == inv_k[e]) acc += grad_out[o];
        }
        grad_feats[i] = acc;
    }
}

static inline int grid_for(int64_t work, int threads) {
    int64_t b = (work + threads - 1) / threads;
    const int64_t cap = (int64_t)kNumSMs * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace pcfb

extern "C" int pcfb_gather(const float *feats, const int64_t *nei, int n_in, int n_out, int K, int C,
                           float *out, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(n_in >= 0 && n_out >= 0 && K >= 1 && C >= 1, "pcfb_gather: bad sizes");
    if ((int64_t)n_out * K * C == 0) return PCFB_OK;
    PCFB_REQUIRE(feats && nei && out, "pcfb_gather: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t E = (int64_t)n_out * K;
    if (C % 4 == 0 && ((uintptr_t)feats % 16 == 0) && ((uintptr_t)out % 16 == 0))
        launch_k(gather_kernel<float4>, grid_for(E * (C / 4), 256), 256, 0, st, reinterpret_cast<const float4 *>(feats), nei, n_in, E, C / 4, reinterpret_cast<float4 *>(out));
    else
        launch_k(gather_kernel<float>, grid_for(E * C, 256), 256, 0, st, feats, nei, n_in, E, C, out);
    return check_launch("pcfb_gather");
}

extern "C" int pcfb_gather_backward(const float *grad_out, const int32_t *inv_neighbors, const uint8_t *inv_k,
                                    const int32_t *inv_idx, int n_in, int n_out, int K, int C,
                                    float *grad_feats, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(n_in >= 0 && n_out >= 0 && K >= 1 && C >= 1, "pcfb_gather_backward: bad sizes");
    if ((int64_t)n_in * C == 0) return PCFB_OK;
    PCFB_REQUIRE(grad_out && inv_neighbors && inv_k && inv_idx && grad_feats, "pcfb_gather_backward: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (C % 4 == 0 && ((uintptr_t)grad_out % 16 == 0) && ((uintptr_t)grad_feats % 16 == 0))
        launch_k(gather_bwd_kernel<float4>, grid_for((int64_t)n_in * (C / 4), 256), 256, 0, st, reinterpret_cast<const float4 *>(grad_out), inv_neighbors, inv_k, inv_idx, n_in, K, C / 4,
            reinterpret_cast<float4 *>(grad_feats));
    else
        launch_k(gather_bwd_kernel<float>, grid_for((int64_t)n_in * C, 256), 256, 0, st, grad_out, inv_neighbors, inv_k, inv_idx, n_in, K, C, grad_feats);
    return check_launch("pcfb_gather_backward");
}

extern "C" int pcfb_gather_max(const float *feats, const int64_t *nei, int n_in, int n_out, int K, int C,
                               float *out, uint8_t *arg, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(n_in >= 0 && n_out >= 0 && K >= 1 && K <= 255 && C >= 1, "pcfb_gather_max: bad sizes");
    if ((int64_t)n_out * C == 0) return PCFB_OK;
    PCFB_REQUIRE(feats && nei && out, "pcfb_gather_max: null pointer");
    launch_k(gather_max_kernel, grid_for((int64_t)n_out * C, 256), 256, 0, static_cast<cudaStream_t>(stream), feats, nei, n_in, n_out, K, C, out, arg);
    return check_launch("pcfb_gather_max");
}

extern "C" int pcfb_gather_max_backward(const float *grad_out, const uint8_t *arg, const int32_t *inv_neighbors,
                                        const uint8_t *inv_k, const int32_t *inv_idx, int n_in, int n_out,
                                        int K, int C, float *grad_feats, void *stream)
{
    using namespace pcfb;
    (void)n_out; (void)K;
    PCFB_REQUIRE(n_in >= 0 && C >= 1, "pcfb_gather_max_backward: bad sizes");
    if ((int64_t)n_in * C == 0) return PCFB_OK;
    PCFB_REQUIRE(grad_out && arg && inv_neighbors && inv_k && inv_idx && grad_feats, "pcfb_gather_max_backward: null pointer");
    launch_k(gather_max_bwd_kernel, grid_for((int64_t)n_in * C, 256), 256, 0, static_cast<cudaStream_t>(stream), grad_out, arg, inv_neighbors, inv_k, inv_idx, n_in, C, grad_feats);
    return check_launch("pcfb_gather_max_backward");
}
