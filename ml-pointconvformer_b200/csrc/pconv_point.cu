// Contraction kernels for the COARSE levels of the pyramid: one CTA per output point (sm_100a).
//
// 19 of PCF_Normal's 25 PointConvFormer layers live on levels of 5 k / 1 k / 184 points.  The tiled kernels (pconv_ws.cu:
// 88 points per CTA, pconv_bwd2.cu: 16 or 64 points per CTA with a thread per (point, 4 neighbours)) run those sizes as
// 3 .. 80 CTAs whose threads walk a point's 48 .. 96 channels serially: 55 .. 80 us per launch for a few MFLOP, all of it
// on the critical path of the step (profiles/step_timeline_r02.txt: 34 launches, 2.05 ms).  Here a point's work is spread
// over the 256 threads of its own CTA -- a few hundred FMAs per thread, every operand staged once in shared memory -- and
// the grid is n_out CTAs, so even 184 points put 184 SMs' worth of CTAs in flight.
//
//   forward  (P only; the Linear then runs on the tensor cores, gemm.cu):  P[c*16 + j] = sum_k G'[k][c] * w[k][j]
//   backward (/root/reference/cpp_wrappers/cpp_pcf_kernel/src/pconv_ops.cu:390-536 semantics, rows a10 / a12):
//       T[k][c]  = sum_j dP[c*16 + j] * w[k][j]
//       dE[k][c] = T[k][c] * g[k][c % H]      (c <  C_in: per-edge gradient of the gathered input, summed per input point
//                                              by pcfb_gather_backward through the kNN inverse map -- no atomics)
//       dadd[k][c - C_in] = T[k][c]           (c >= C_in)
//       dg[k][h] = sum_{c < C_in, c % H == h} T[k][c] * G[k][c]
//       dw[k][j] = sum_c dP[c*16 + j] * G'[k][c]
//   with G[k][:] = cat(feats[nei[m][k]], additional[m][k]) and G' = G with the first C_in channels times g[k][c % H].
// K = 16 neighbours, C_mid = 16, C_in + C_add <= 128, H in {0, 1, 2, 4, 8, 16}.  Exact fp32 (fmaf, fixed order).
#include "common.cuh"
#include <stdlib.h>

namespace pcfb {

constexpr int PP_K = 16, PP_MID = 16, PP_CMAX = 128, PP_THREADS = 256, PP_DPS = 17;

struct PointArgs {
    pcfb_pconv_shape s;
    const float *dP, *feats, *weights, *additional, *guidance;
    const int64_t *nei;
    float *grad_weights, *grad_additional, *grad_guidance, *grad_edge, *P;
};

// stage one point's operands: G (raw), Gg (guided), w, g
__device__ __forceinline__ void pp_stage(const PointArgs &a, int m, float *G_s, float *Gg_s, float *w_s, float *g_s, int C_cat)
{
    const int t = threadIdx.x, C_in = a.s.C_in, C_add = a.s.C_add, H = a.s.H;
    w_s[t] = __ldg(a.weights + (size_t)m * PP_K * PP_MID + t);                  // 16 x 16 = one per thread
    if (t < PP_K * H) g_s[t] = __ldg(a.guidance + (size_t)m * PP_K * H + t);
    for (int i = t; i < PP_K * C_cat; i += PP_THREADS) {
        const int k = i / C_cat, c = i - k * C_cat;
        float v;
        if (c < C_in) {
            const long long q = __ldg(a.nei + (size_t)m * PP_K + k);
            v = (q >= 0 && q < a.s.n_in) ? __ldg(a.feats + (size_t)q * C_in + c) : 0.f;   // padding rows gather zeros
        } else {
            v = __ldg(a.additional + ((size_t)m * PP_K + k) * C_add + (c - C_in));
        }
        G_s[i] = v;
    }
    __syncthreads();
    for (int i = t; i < PP_K * C_cat; i += PP_THREADS) {
        const int k = i / C_cat, c = i - k * C_cat;
        Gg_s[i] = (H > 0 && c < C_in) ? G_s[i] * g_s[k * H + (c % H)] : G_s[i];
    }
}

__global__ void __launch_bounds__(PP_THREADS)
pconv_point_fwd_p_kernel(PointArgs a)
{
    pdl_wait();
    __shared__ float G_s[PP_K * PP_CMAX], Gg_s[PP_K * PP_CMAX], w_s[PP_K * PP_MID], g_s[PP_K * 16];
    const int m = blockIdx.x, t = threadIdx.x;
    const int C_cat = a.s.C_in + a.s.C_add;
    pp_stage(a, m, G_s, Gg_s, w_s, g_s, C_cat);
    __syncthreads();
    const int j = t & 15, cs = t >> 4;
    float wk[PP_K];
#pragma unroll
    for (int k = 0; k < PP_K; ++k) wk[k] = w_s[k * PP_MID + j];
    for (int c = cs; c < C_cat; c += 16) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < PP_K; ++k) acc = fmaf(Gg_s[k * C_cat + c], wk[k], acc);
        a.P[(size_t)m * C_cat * PP_MID + (size_t)c * PP_MID + j] = acc;
    }
}

__global__ void __launch_bounds__(PP_THREADS)
pconv_point_bwd_kernel(PointArgs a)
{
    pdl_wait();
    __shared__ float G_s[PP_K * PP_CMAX], Gg_s[PP_K * PP_CMAX], w_s[PP_K * PP_MID], g_s[PP_K * 16];
    __shared__ float dP_s[PP_CMAX * PP_DPS];                     // rows padded to 17 floats: conflict-free column walks
    const int m = blockIdx.x, t = threadIdx.x;
    const int C_in = a.s.C_in, C_add = a.s.C_add, H = a.s.H, C_cat = C_in + C_add;
    {
        const float4 *src = reinterpret_cast<const float4 *>(a.dP + (size_t)m * C_cat * PP_MID);
        for (int i = t; i < C_cat * PP_MID / 4; i += PP_THREADS) {
            const float4 v = __ldg(src + i);
            const int c = i >> 2, j = (i & 3) * 4;
            float *d = dP_s + c * PP_DPS + j;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    }
    pp_stage(a, m, G_s, Gg_s, w_s, g_s, C_cat);
    __syncthreads();
    const int k = t >> 4, lo = t & 15;
    // ---- T[k][c], dE / dadd / dg: thread = (neighbour k, channel slice lo) ----
    {
        float wj[PP_MID];
#pragma unroll
        for (int j = 0; j < PP_MID; ++j) wj[j] = w_s[k * PP_MID + j];
        float dg = 0.f;
        for (int c = lo; c < C_cat; c += 16) {
            const float *dp = dP_s + c * PP_DPS;
            float T = 0.f;
#pragma unroll
            for (int j = 0; j < PP_MID; ++j) T = fmaf(dp[j], wj[j], T);
            if (c < C_in) {
                if (H > 0) {
                    dg = fmaf(T, G_s[k * C_cat + c], dg);
                    T *= g_s[k * H + (c % H)];
                }
                if (a.grad_edge) a.grad_edge[((size_t)m * PP_K + k) * C_in + c] = T;
            } else if (a.grad_additional) {
                a.grad_additional[((size_t)m * PP_K + k) * C_add + (c - C_in)] = T;
            }
        }
        if (H > 0 && a.grad_guidance) {
            // every channel of this thread has head lo % H (16 % H == 0); lanes of one k with the same head sit H apart
            for (int off = H; off < 16; off <<= 1) dg += __shfl_xor_sync(0xffffffffu, dg, off);
            if (lo < H) a.grad_guidance[((size_t)m * PP_K + k) * H + lo] = dg;
        }
    }
    // ---- dw[k][j]: thread = (neighbour k, weight column j = lo) ----
    if (a.grad_weights) {
        float acc = 0.f;
        const float *gg = Gg_s + k * C_cat;
        for (int c = 0; c < C_cat; ++c) acc = fmaf(dP_s[c * PP_DPS + lo], gg[c], acc);
        a.grad_weights[((size_t)m * PP_K + k) * PP_MID + lo] = acc;
    }
}

bool pconv_point_supported(const pcfb_pconv_shape *s) {
    if (s->K != PP_K || s->C_mid != PP_MID) return false;
    if (s->C_in + s->C_add > PP_CMAX || s->C_in < 1) return false;
    if (s->H != 0 && !(s->H == 1 || s->H == 2 || s->H == 4 || s->H == 8 || s->H == 16)) return false;
    return true;
}

// below this many output points the per-point kernels beat the tiled ones (measured: profiles/README.md, round 2)
static int g_point_max = -1;
int pconv_point_max_points() {
    if (g_point_max < 0) {
        const char *e = getenv("PCFB_POINT_KERNEL_MAX");
        g_point_max = e ? atoi(e) : 4000;            // 10 cm pyramid: 3 000 -> 20.22, 12 000 -> 20.28, 30 000 -> 20.39 ms/step
    }
    return g_point_max;
}

int pconv_point_bwd(const pcfb_pconv_shape *s, const float *dP, const float *feats, const int64_t *nei, const float *weights,
                    const float *additional, const float *guidance, float *grad_weights, float *grad_additional,
                    float *grad_guidance, float *grad_edge, cudaStream_t st)
{
    PCFB_REQUIRE(pconv_point_supported(s), "pconv_point_bwd: unsupported shape");
    PCFB_REQUIRE(((uintptr_t)dP % 16 == 0), "pconv_point_bwd: unaligned dP");
    if (s->n_out == 0) return PCFB_OK;
    PointArgs a{};
    a.s = *s; a.dP = dP; a.feats = feats; a.nei = nei; a.weights = weights; a.additional = additional; a.guidance = guidance;
    a.grad_weights = grad_weights; a.grad_additional = grad_additional; a.grad_guidance = grad_guidance; a.grad_edge = grad_edge;
    launch_k(pconv_point_bwd_kernel, s->n_out, PP_THREADS, 0, st, a);
    return check_launch("pconv_point_bwd_kernel");
}

int pconv_point_fwd_p(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                      const float *additional, const float *guidance, float *P, cudaStream_t st)
{
    PCFB_REQUIRE(pconv_point_supported(s), "pconv_point_fwd_p: unsupported shape");
    if (s->n_out == 0) return PCFB_OK;
    PointArgs a{};
    a.s = *s; a.feats = feats; a.nei = nei; a.weights = weights; a.additional = additional; a.guidance = guidance; a.P = P;
    launch_k(pconv_point_fwd_p_kernel, s->n_out, PP_THREADS, 0, st, a);
    return check_launch("pconv_point_fwd_p_kernel");
}

}  // namespace pcfb

extern "C" int pcfb_set_point_kernel_max(int n_points)
{
    const int old = pcfb::pconv_point_max_points();
    pcfb::g_point_max = n_points < 0 ? 0 : n_points;
    return old;
}
