// C_mid == 1 fast path of the PointConv contraction (sm_100a).
//
// With a single weightnet output (mid_dim_back = 1: every PointConvTransposePE of the decoder,
// /root/reference/model_architecture.py:370-388, layers.py:1086-1092) the contraction degenerates to a
// weighted neighbour sum  P[m,c] = sum_k G[m,k,c] * w[m,k]  -- no K x C_mid matrix to keep on chip, so it is a
// pure HBM/L2 streaming kernel, and the Linear becomes an ordinary [M, C_cat] x [C_cat, C_out] product
// (pcfb_gemm_nt on tcgen05).  Backward: dP = dY W (gemm_nt), dw[m,k] = <dP[m,:], G[m,k,:]> (warp dot
// products), dadd = dP[m, C_in:] * w[m,k], and dx[p,c] = sum over the inverse segment of p of dP[n,c] w[n,k]
// gathered straight from dP (no per-edge gradient tensor, no atomics).
#include "common.cuh"

namespace pcfb {

// thread <-> (point m, 4 channels); coalesced along channels, 16 independent row loads in flight
__global__ void mid1_fwd_kernel(const float *__restrict__ feats, const int64_t *__restrict__ nei,
                                const float *__restrict__ w, const float *__restrict__ add, int n_in, int n_out, int K,
                                int C_in, int C_add, float *__restrict__ P)
{
    pdl_wait();
    const int C_cat = C_in + C_add, G4 = C_cat / 4, I4 = C_in / 4;
    const int64_t total = (int64_t)n_out * G4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / G4), c4 = (int)(i - (int64_t)m * G4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t *nm = nei + (size_t)m * K;
        const float *wm = w + (size_t)m * K;
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c4 < I4) {
                const int64_t q = nm[k];
                if (q >= 0 && q < n_in) v = __ldg(reinterpret_cast<const float4 *>(feats + (size_t)q * C_in) + c4);
            } else {
                v = __ldg(reinterpret_cast<const float4 *>(add + ((size_t)m * K + k) * C_add) + (c4 - I4));
            }
            const float wk = wm[k];
            acc.x = fmaf(v.x, wk, acc.x); acc.y = fmaf(v.y, wk, acc.y); acc.z = fmaf(v.z, wk, acc.z); acc.w = fmaf(v.w, wk, acc.w);
        }
        reinterpret_cast<float4 *>(P)[i] = acc;
    }
}

// one warp per output point: dw[m,k] = sum_c dP[m,c] G[m,k,c];  dadd[m,k,c'] = dP[m,C_in+c'] w[m,k]
__global__ void mid1_bwd_point_kernel(const float *__restrict__ dP, const float *__restrict__ feats,
                                      const int64_t *__restrict__ nei, const float *__restrict__ w,
                                      const float *__restrict__ add, int n_in, int n_out, int K, int C_in, int C_add,
                                      float *__restrict__ dw, float *__restrict__ dadd)
{
    pdl_wait();
    const int C_cat = C_in + C_add;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < n_out; m += warps) {
        const float *dpm = dP + (size_t)m * C_cat;
        for (int k = 0; k < K; ++k) {
            const int64_t q = nei[(size_t)m * K + k];
            const bool valid = q >= 0 && q < n_in;
            const float wk = w[(size_t)m * K + k];
            float part = 0.f;
            for (int c = lane; c < C_cat; c += 32) {
                const float d = dpm[c];
                float gv;
                if (c < C_in) gv = valid ? __ldg(feats + (size_t)q * C_in + c) : 0.f;
                else {
                    gv = __ldg(add + ((size_t)m * K + k) * C_add + (c - C_in));
                    if (dadd) dadd[((size_t)m * K + k) * C_add + (c - C_in)] = d * wk;
                }
                part = fmaf(d, gv, part);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (dw && lane == 0) dw[(size_t)m * K + k] = part;
        }
    }
}

// thread <-> (input point p, 4 channels): dx[p,c] = sum_{e in seg(p)} dP[n_e, c] * w[n_e, k_e]
__global__ void mid1_bwd_input_kernel(const float *__restrict__ dP, const float *__restrict__ w,
                                      const int32_t *__restrict__ inv_n, const uint8_t *__restrict__ inv_k,
                                      const int32_t *__restrict__ inv_idx, int n_in, int K, int C_in, int C_cat,
                                      float *__restrict__ dx)
{
    pdl_wait();
    const int I4 = C_in / 4;
    const int64_t total = (int64_t)n_in * I4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i / I4), c4 = (int)(i - (int64_t)p * I4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int e1 = inv_idx[p + 1];
        for (int e = inv_idx[p]; e < e1; ++e) {
            const int n = inv_n[e];
            const float wk = w[(size_t)n * K + inv_k[e]];
            const float4 v = __ldg(reinterpret_cast<const float4 *>(dP + (size_t)n * C_cat) + c4);
            acc.x = fmaf(v.x, wk, acc.x); acc.y = fmaf(v.y, wk, acc.y); acc.z = fmaf(v.z, wk, acc.z); acc.w = fmaf(v.w, wk, acc.w);
        }
        reinterpret_cast<float4 *>(dx)[i] = acc;
    }
}

static inline int m1_blocks(int64_t work, int threads) {
    int64_t b = (work + threads - 1) / threads;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// C_mid in {2, 3, 4} (mid_dim_back = 3: every PointConvTransposePE of configPCF_2cm_PTF2; mid_dim = 4 with guidance: the
// unfused PointConvFormer layers of configPCF_10cm_lite): the same streaming structure with
// CM accumulators per channel -- P[m, c*CM + j] = sum_k G[m,k,c] * w[m,k,j]; thread <-> (point, 4 channels), its 4 x CM
// results are 16 x CM contiguous bytes of P.  The generic CUDA-core contraction kernel these layers used before spends
// 6 ms on the 250 k-point level of that config (59 % of an inference pass, scripts/profile_infer.py); this one is bound by
// the L2 gather like mid1_fwd_kernel.
template <int CM>
__global__ void midn_fwd_kernel(const float *__restrict__ feats, const int64_t *__restrict__ nei,
                                const float *__restrict__ w, const float *__restrict__ add, const float *__restrict__ gd, int H,
                                int n_in, int n_out, int K, int C_in, int C_add, float *__restrict__ P)
{
    pdl_wait();
    const int C_cat = C_in + C_add, G4 = C_cat / 4, I4 = C_in / 4;
    const int64_t total = (int64_t)n_out * G4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / G4), c4 = (int)(i - (int64_t)m * G4);
        float acc[4][CM];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < CM; ++j) acc[a][j] = 0.f;
        const int64_t *nm = nei + (size_t)m * K;
        const float *wm = w + (size_t)m * K * CM;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c4 < I4) {
                const int64_t q = nm[k];
                if (q >= 0 && q < n_in) v = __ldg(reinterpret_cast<const float4 *>(feats + (size_t)q * C_in) + c4);
                if (gd) {                                          // guidance: channel c of the gathered row times head c % H
                    const float *gk = gd + ((size_t)m * K + k) * H;
                    const int c = c4 * 4;
                    v.x *= __ldg(gk + c % H); v.y *= __ldg(gk + (c + 1) % H); v.z *= __ldg(gk + (c + 2) % H); v.w *= __ldg(gk + (c + 3) % H);
                }
            } else {
                v = __ldg(reinterpret_cast<const float4 *>(add + ((size_t)m * K + k) * C_add) + (c4 - I4));
            }
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < CM; ++j) {
                const float wk = __ldg(wm + k * CM + j);
#pragma unroll
                for (int a = 0; a < 4; ++a) acc[a][j] = fmaf(vv[a], wk, acc[a][j]);
            }
        }
        float flat[4 * CM];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < CM; ++j) flat[a * CM + j] = acc[a][j];
        float4 *dst = reinterpret_cast<float4 *>(P + ((size_t)m * C_cat + (size_t)c4 * 4) * CM);
#pragma unroll
        for (int u = 0; u < CM; ++u) dst[u] = make_float4(flat[4 * u], flat[4 * u + 1], flat[4 * u + 2], flat[4 * u + 3]);
    }
}

bool pconv_midn_supported(const pcfb_pconv_shape *s) {
    return s->C_mid >= 2 && s->C_mid <= 4 && s->H >= 0 && s->C_in >= 4 && s->C_in % 4 == 0 && s->C_add % 4 == 0;
}

int pconv_midn_forward_p(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                         const float *additional, const float *guidance, float *P, cudaStream_t st)
{
    PCFB_REQUIRE((s->H > 0) == (guidance != nullptr), "pconv_midn_forward_p: H and guidance disagree");
    const float *gd = guidance;
    const int H = s->H;
    PCFB_REQUIRE(pconv_midn_supported(s), "pconv_midn_forward_p: unsupported shape");
    PCFB_REQUIRE(((uintptr_t)feats % 16 == 0) && ((uintptr_t)P % 16 == 0) && (s->C_add == 0 || (uintptr_t)additional % 16 == 0),
                 "pcfb_pconv: the small-C_mid path needs 16-byte aligned feats/additional/P");
    if (s->n_out == 0) return PCFB_OK;
    const int64_t work = (int64_t)s->n_out * ((s->C_in + s->C_add) / 4);
    const int blocks = m1_blocks(work, 256);
    if (s->C_mid == 2) launch_k(midn_fwd_kernel<2>, blocks, 256, 0, st, feats, nei, weights, additional, gd, H, s->n_in, s->n_out, s->K, s->C_in, s->C_add, P);
    else if (s->C_mid == 3) launch_k(midn_fwd_kernel<3>, blocks, 256, 0, st, feats, nei, weights, additional, gd, H, s->n_in, s->n_out, s->K, s->C_in, s->C_add, P);
    else launch_k(midn_fwd_kernel<4>, blocks, 256, 0, st, feats, nei, weights, additional, gd, H, s->n_in, s->n_out, s->K, s->C_in, s->C_add, P);
    return check_launch("midn_fwd_kernel");
}

bool pconv_mid1_supported(const pcfb_pconv_shape *s) {
    return s->C_mid == 1 && s->H == 0 && s->C_in % 4 == 0 && s->C_add % 4 == 0;
}

int pconv_mid1_forward_p(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                         const float *additional, float *P, cudaStream_t st)
{
    PCFB_REQUIRE(((uintptr_t)feats % 16 == 0) && ((uintptr_t)P % 16 == 0) && (s->C_add == 0 || (uintptr_t)additional % 16 == 0),
                 "pcfb_pconv: C_mid=1 path needs 16-byte aligned feats/additional/P");
    if (s->n_out == 0) return PCFB_OK;
    const int64_t work = (int64_t)s->n_out * ((s->C_in + s->C_add) / 4);
    launch_k(mid1_fwd_kernel, m1_blocks(work, 256), 256, 0, st, feats, nei, weights, additional, s->n_in, s->n_out, s->K, s->C_in, s->C_add, P);
    return check_launch("mid1_fwd_kernel");
}

int pconv_mid1_backward(const pcfb_pconv_shape *s, const float *dP, const float *feats, const int64_t *nei,
                        const int32_t *inv_n, const uint8_t *inv_k, const int32_t *inv_idx, const float *weights,
                        const float *additional, float *grad_feats, float *grad_weights, float *grad_additional,
                        cudaStream_t st)
{
    int rc;
    if (s->n_out > 0 && (grad_weights || grad_additional)) {
        launch_k(mid1_bwd_point_kernel, m1_blocks((int64_t)s->n_out * 32, 256), 256, 0, st, dP, feats, nei, weights, additional, s->n_in, s->n_out, s->K, s->C_in, s->C_add, grad_weights, grad_additional);
        if ((rc = check_launch("mid1_bwd_point_kernel"))) return rc;
    }
    if (grad_feats) {
        PCFB_REQUIRE(inv_n && inv_k && inv_idx, "pcfb_pconv_backward: grad_feats needs the inverse map");
        PCFB_REQUIRE(((uintptr_t)dP % 16 == 0) && ((uintptr_t)grad_feats % 16 == 0), "pcfb_pconv_backward: unaligned buffers");
        launch_k(mid1_bwd_input_kernel, m1_blocks((int64_t)s->n_in * (s->C_in / 4), 256), 256, 0, st, dP, weights, inv_n, inv_k, inv_idx, s->n_in, s->K, s->C_in, s->C_in + s->C_add, grad_feats);
        if ((rc = check_launch("mid1_bwd_input_kernel"))) return rc;
    }
    return PCFB_OK;
}

}  // namespace pcfb
