// C-ABI entry points of the fused PointConv / PointConvFormer contraction: variant dispatch.
#include "common.cuh"

namespace pcfb {
// pconv_simt.cu
int pconv_forward_simt(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                       const float *additional, const float *guidance, const float *lin_w, const float *lin_b,
                       float *out_y, float *out_p, cudaStream_t st);
size_t pconv_backward_simt_workspace(const pcfb_pconv_shape *s);
int pconv_backward_simt(const pcfb_pconv_shape *s, const float *grad_y, const float *grad_p, const float *feats,
                        const int64_t *nei, const int32_t *inv_n, const uint8_t *inv_k, const int32_t *inv_idx,
                        const float *weights, const float *additional, const float *guidance, const float *lin_w,
                        const float *pconv_out, float *grad_feats, float *grad_weights, float *grad_additional,
                        float *grad_guidance, float *grad_lin_w, float *grad_lin_b, void *workspace,
                        size_t workspace_bytes, cudaStream_t st);
// pconv_umma.cu
bool pconv_forward_umma_supported(const pcfb_pconv_shape *s, bool has_lin);
size_t pconv_forward_umma_workspace(const pcfb_pconv_shape *s);
int pconv_forward_umma(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                       const float *additional, const float *guidance, const float *lin_w, const float *lin_b,
                       float *out_y, float *out_p, void *workspace, size_t workspace_bytes, cudaStream_t st);
// pconv_umma2.cu (pipelined tcgen05 forward)
bool pconv_forward_umma2_supported(const pcfb_pconv_shape *s, bool has_lin);
size_t pconv_forward_umma2_workspace(const pcfb_pconv_shape *s);
int pconv_forward_umma2(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                        const float *additional, const float *guidance, const float *lin_w, const float *lin_b,
                        float *out_y, float *out_p, void *workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace pcfb

extern "C" int pcfb_pconv_forward_supported(const pcfb_pconv_shape *s, int variant)
{
    if (!s) return 0;
    if (variant == 2 || variant == 3)
        return (pcfb::pconv_forward_umma_supported(s, s->C_out > 0) || pcfb::pconv_forward_umma2_supported(s, s->C_out > 0)) ? 1 : 0;
    return 1;
}

extern "C" size_t pcfb_pconv_forward_workspace(const pcfb_pconv_shape *s, int variant)
{
    if (!s) return 0;
    if (variant == 1) return 0;
    if (variant != 3 && pcfb::pconv_forward_umma2_supported(s, s->C_out > 0)) return pcfb::pconv_forward_umma2_workspace(s);
    return pcfb::pconv_forward_umma_supported(s, s->C_out > 0) ? pcfb::pconv_forward_umma_workspace(s) : 0;
}

extern "C" int pcfb_pconv_forward(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei,
                                  const float *weights, const float *additional, const float *guidance,
                                  const float *lin_w, const float *lin_b, float *out_y, float *out_p,
                                  void *workspace, size_t workspace_bytes, int variant, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(s != nullptr, "pcfb_pconv_forward: null shape");
    PCFB_REQUIRE(feats && nei && weights, "pcfb_pconv_forward: null pointer");
    PCFB_REQUIRE(s->C_add == 0 || additional, "pcfb_pconv_forward: C_add=%d but additional is NULL", s->C_add);
    PCFB_REQUIRE((s->H > 0) == (guidance != nullptr), "pcfb_pconv_forward: H and guidance disagree");
    PCFB_REQUIRE(lin_w ? out_y != nullptr : out_p != nullptr, "pcfb_pconv_forward: no output requested");
    PCFB_REQUIRE(variant >= 0 && variant <= 3, "pcfb_pconv_forward: unknown variant %d", variant);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // variants: 0 auto, 1 exact-fp32 SIMT, 2 tcgen05 (pipelined kernel when the tile fits, else the simple one),
    // 3 tcgen05 simple kernel only (kept for bisecting)
    const bool u1_ok = lin_w && pconv_forward_umma_supported(s, true);
    const bool u2_ok = lin_w && variant != 3 && pconv_forward_umma2_supported(s, true);
    if ((variant == 2 || variant == 3) && !u1_ok && !u2_ok) {
        set_error("pcfb_pconv_forward: tcgen05 variant does not support this shape");
        return PCFB_ERR_UNSUPPORTED;
    }
    if (variant != 1 && u2_ok)
        return pconv_forward_umma2(s, feats, nei, weights, additional, guidance, lin_w, lin_b, out_y, out_p,
                                   workspace, workspace_bytes, st);
    if (variant != 1 && u1_ok)
        return pconv_forward_umma(s, feats, nei, weights, additional, guidance, lin_w, lin_b, out_y, out_p,
                                  workspace, workspace_bytes, st);
    return pconv_forward_simt(s, feats, nei, weights, additional, guidance, lin_w, lin_b, out_y, out_p, st);
}

extern "C" size_t pcfb_pconv_backward_workspace(const pcfb_pconv_shape *s, int variant)
{
    (void)variant;
    return s ? pcfb::pconv_backward_simt_workspace(s) : 0;
}

extern "C" int pcfb_pconv_backward(const pcfb_pconv_shape *s, const float *grad_y, const float *grad_p,
                                   const float *feats, const int64_t *nei, const int32_t *inv_neighbors,
                                   const uint8_t *inv_k, const int32_t *inv_idx, const float *weights,
                                   const float *additional, const float *guidance, const float *lin_w,
                                   const float *pconv_out, float *grad_feats, float *grad_weights,
                                   float *grad_additional, float *grad_guidance, float *grad_lin_w,
                                   float *grad_lin_b, void *workspace, size_t workspace_bytes, int variant,
                                   void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(s != nullptr, "pcfb_pconv_backward: null shape");
    PCFB_REQUIRE(feats && nei && weights, "pcfb_pconv_backward: null pointer");
    PCFB_REQUIRE(s->C_add == 0 || additional, "pcfb_pconv_backward: C_add=%d but additional is NULL", s->C_add);
    PCFB_REQUIRE((s->H > 0) == (guidance != nullptr), "pcfb_pconv_backward: H and guidance disagree");
    PCFB_REQUIRE(variant >= 0 && variant <= 2, "pcfb_pconv_backward: unknown variant %d", variant);
    return pconv_backward_simt(s, grad_y, grad_p, feats, nei, inv_neighbors, inv_k, inv_idx, weights, additional,
                               guidance, lin_w, pconv_out, grad_feats, grad_weights, grad_additional, grad_guidance,
                               grad_lin_w, grad_lin_b, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
