// P[m, c*C_mid + j] = sum_k G[m,k,c] * w[m,k,j] for small outputs (a few hundred to a few thousand points at the deep
// levels): one thread per (point, channel, 4 weights), so even 184 points fill the machine.  The tiled kernels above
// put whole points on a CTA and leave 2-3 CTAs busy for ~0.4 ms at that size.  Exact fp32 (fmaf in k order).
#include "common.cuh"

namespace pcfb {

__global__ void pconv_p_small_kernel(pcfb_pconv_shape s, const float *__restrict__ feats, const int64_t *__restrict__ nei,
                                     const float *__restrict__ weights, const float *__restrict__ additional,
                                     const float *__restrict__ guidance, float *__restrict__ P)
{
    pdl_wait();
    const int C_cat = s.C_in + s.C_add, JQ = s.C_mid / 4;
    const int64_t total = (int64_t)s.n_out * C_cat * JQ;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int jq = (int)(i % JQ);
        const int c = (int)((i / JQ) % C_cat);
        const int m = (int)(i / ((int64_t)JQ * C_cat));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < s.K; ++k) {
            float x;
            if (c < s.C_in) {
                const long long q = nei[(size_t)m * s.K + k];
                x = (q >= 0 && q < s.n_in) ? __ldg(feats + (size_t)q * s.C_in + c) : 0.f;
                if (s.H > 0) x *= __ldg(guidance + ((size_t)m * s.K + k) * s.H + (c % s.H));
            } else {
                x = __ldg(additional + ((size_t)m * s.K + k) * s.C_add + (c - s.C_in));
            }
            const float4 w = __ldg(reinterpret_cast<const float4 *>(weights + ((size_t)m * s.K + k) * s.C_mid) + jq);
            acc.x = fmaf(x, w.x, acc.x); acc.y = fmaf(x, w.y, acc.y); acc.z = fmaf(x, w.z, acc.z); acc.w = fmaf(x, w.w, acc.w);
        }
        *reinterpret_cast<float4 *>(P + (size_t)m * C_cat * s.C_mid + (size_t)c * s.C_mid + jq * 4) = acc;
    }
}

bool pconv_p_small_supported(const pcfb_pconv_shape *s, const float *weights, const float *P) {
    return s->C_mid % 4 == 0 && ((uintptr_t)weights % 16 == 0) && ((uintptr_t)P % 16 == 0);
}

int pconv_p_small(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                  const float *additional, const float *guidance, float *P, cudaStream_t st)
{
    if (s->n_out == 0) return PCFB_OK;
    const int64_t total = (int64_t)s->n_out * (s->C_in + s->C_add) * (s->C_mid / 4);
    const int64_t blocks = (total + 255) / 256;
    launch_k(pconv_p_small_kernel, (int)(blocks < 8 * kNumSMs ? blocks : 8 * kNumSMs), 256, 0, st, *s, feats, nei, weights, additional, guidance, P);
    return check_launch("pconv_p_small_kernel");
}

}  // namespace pcfb
