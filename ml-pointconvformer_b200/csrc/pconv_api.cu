// C-ABI entry points of the fused PointConv / PointConvFormer contraction: variant dispatch.
#include "common.cuh"
#include <stdlib.h>

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
// pconv_ws.cu (warp-specialised tcgen05 forward, C_mid == 16)
bool pconv_forward_ws_supported(const pcfb_pconv_shape *s, bool has_lin);
size_t pconv_forward_ws_workspace(const pcfb_pconv_shape *s);
int pconv_forward_ws(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                     const float *additional, const float *guidance, const float *lin_w, const float *lin_b,
                     float *out_y, float *out_p, void *workspace, size_t workspace_bytes, cudaStream_t st);
// pconv_small.cu (P for small outputs, one thread per (point, channel, 4 weights))
bool pconv_p_small_supported(const pcfb_pconv_shape *s, const float *weights, const float *P);
int pconv_p_small(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                  const float *additional, const float *guidance, float *P, cudaStream_t st);
// pconv_mid1.cu (C_mid == 1: weighted neighbour sum + tensor-core Linear)
bool pconv_mid1_supported(const pcfb_pconv_shape *s);
bool pconv_midn_supported(const pcfb_pconv_shape *s);
int pconv_midn_forward_p(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                         const float *additional, const float *guidance, float *P, cudaStream_t st);
int pconv_mid1_forward_p(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                         const float *additional, float *P, cudaStream_t st);
int pconv_mid1_backward(const pcfb_pconv_shape *s, const float *dP, const float *feats, const int64_t *nei,
                        const int32_t *inv_n, const uint8_t *inv_k, const int32_t *inv_idx, const float *weights,
                        const float *additional, float *grad_feats, float *grad_weights, float *grad_additional,
                        cudaStream_t st);
// pconv_bwd2.cu (pipelined contraction backward on dP)
bool pconv_bwd2_supported(const pcfb_pconv_shape *s);
int pconv_bwd2(const pcfb_pconv_shape *s, const float *dP, const float *feats, const int64_t *nei, const float *weights,
               const float *additional, const float *guidance, float *grad_weights, float *grad_additional,
               float *grad_guidance, float *grad_edge, cudaStream_t st);
// pconv_point.cu (one CTA per output point: the coarse levels of the pyramid)
bool pconv_point_supported(const pcfb_pconv_shape *s);
int pconv_point_max_points();
int pconv_point_bwd(const pcfb_pconv_shape *s, const float *dP, const float *feats, const int64_t *nei, const float *weights,
                    const float *additional, const float *guidance, float *grad_weights, float *grad_additional,
                    float *grad_guidance, float *grad_edge, cudaStream_t st);
int pconv_point_fwd_p(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                      const float *additional, const float *guidance, float *P, cudaStream_t st);
}  // namespace pcfb

// coarse levels: per-point CTAs (pconv_point.cu) instead of the tiled kernels
static bool point_path(const pcfb_pconv_shape *s) {
    return s->n_out > 0 && s->n_out <= pcfb::pconv_point_max_points() && pcfb::pconv_point_supported(s);
}
// the contraction backward on dP: per-point kernel on the coarse levels, the pipelined tiled kernel otherwise
static int contraction_backward(const pcfb_pconv_shape *s, const float *dP, const float *feats, const int64_t *nei,
                                const float *weights, const float *additional, const float *guidance, float *grad_weights,
                                float *grad_additional, float *grad_guidance, float *grad_edge, cudaStream_t st) {
    if (point_path(s))
        return pcfb::pconv_point_bwd(s, dP, feats, nei, weights, additional, guidance, grad_weights, grad_additional, grad_guidance, grad_edge, st);
    return pcfb::pconv_bwd2(s, dP, feats, nei, weights, additional, guidance, grad_weights, grad_additional, grad_guidance, grad_edge, st);
}

static bool mid1_path(const pcfb_pconv_shape *s, int variant) {
    return variant != 1 && variant != 3 && variant != 4 && s->C_out >= 1 && s->C_out <= 256 && pcfb::pconv_mid1_supported(s);
}

// auto only: shapes that neither pipelined tcgen05 kernel holds (wide C_out at the deep levels, a few hundred points):
// P by the CUDA-core contraction kernel, Y = P W^T + b as a column-block tensor-core GEMM.  The simple tcgen05 kernel
// these shapes used before keeps 2-3 CTAs busy for ~0.4 ms.
static bool midn_compose() {                        // PCFB_MIDN_COMPOSE=0: A/B switch
    static int on = -1;
    if (on < 0) { const char *e = getenv("PCFB_MIDN_COMPOSE"); on = (e && e[0] == '0') ? 0 : 1; }
    return on != 0;
}
static bool compose_path(const pcfb_pconv_shape *s) {
    if (s->C_out > 0 && s->C_mid > 1 && point_path(s) && s->n_out <= pcfb::pconv_point_max_points() / 2) return true;
    // C_mid 2..4 on a large level: the streaming weighted-sum kernel + the tensor-core Linear beat the 64-point-tile fused
    // kernel (configPCF_10cm_lite, 305 k points: 580 us per layer, 10 % of the HBM roofline)
    if (midn_compose() && s->C_out > 0 && s->n_out > 16384 && pcfb::pconv_midn_supported(s)) return true;
    return s->C_out > 0 && s->C_mid > 1 && !pcfb::pconv_forward_ws_supported(s, true) && !pcfb::pconv_forward_umma2_supported(s, true);
}

extern "C" int pcfb_pconv_forward_supported(const pcfb_pconv_shape *s, int variant)
{
    if (!s) return 0;
    if (variant == 2 && mid1_path(s, variant)) return 1;
    if (variant == 4) return pcfb::pconv_forward_ws_supported(s, s->C_out > 0) ? 1 : 0;
    if (variant == 2 || variant == 3)
        return (pcfb::pconv_forward_umma_supported(s, s->C_out > 0) || pcfb::pconv_forward_umma2_supported(s, s->C_out > 0)) ? 1 : 0;
    return 1;
}

extern "C" size_t pcfb_pconv_forward_workspace(const pcfb_pconv_shape *s, int variant)
{
    if (!s) return 0;
    if (variant == 1) return 0;
    if (mid1_path(s, variant))
        return pcfb::align_up((size_t)s->n_out * (s->C_in + s->C_add) * sizeof(float), 256) + pcfb_gemm_nt_workspace(s->C_out, s->C_in + s->C_add);
    size_t ws = 0;                                      // auto may fall from one tcgen05 kernel to the next: size for all
    if (variant == 0 && compose_path(s))
        ws = pcfb::align_up((size_t)s->n_out * (s->C_in + s->C_add) * s->C_mid * sizeof(float), 256) +
             pcfb_gemm_nt_workspace(s->C_out, (s->C_in + s->C_add) * s->C_mid);
    if ((variant == 0 || variant == 4) && pcfb::pconv_forward_ws_supported(s, s->C_out > 0)) {
        const size_t w1 = pcfb::pconv_forward_ws_workspace(s);
        ws = w1 > ws ? w1 : ws;
    }
    if (variant == 4) return ws;
    if (variant != 3 && pcfb::pconv_forward_umma2_supported(s, s->C_out > 0)) {
        const size_t w2 = pcfb::pconv_forward_umma2_workspace(s);
        return w2 > ws ? w2 : ws;
    }
    if (ws) return ws;
    if (variant == 0 && compose_path(s))
        return pcfb::align_up((size_t)s->n_out * (s->C_in + s->C_add) * s->C_mid * sizeof(float), 256) +
               pcfb_gemm_nt_workspace(s->C_out, (s->C_in + s->C_add) * s->C_mid);
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
    PCFB_REQUIRE(variant >= 0 && variant <= 4, "pcfb_pconv_forward: unknown variant %d", variant);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // variants: 0 auto, 1 exact-fp32 SIMT, 2 tcgen05 (pipelined kernel when the tile fits, else the simple one),
    // 3 tcgen05 simple kernel only (kept for bisecting), 4 warp-specialised tcgen05 kernel only
    if (lin_w && mid1_path(s, variant)) {
        // C_mid == 1: P = weighted neighbour sum (streaming kernel), Y = P W^T + b on tcgen05
        const int C_cat = s->C_in + s->C_add;
        const size_t p_bytes = align_up((size_t)s->n_out * C_cat * sizeof(float), 256);
        const size_t nt_bytes = pcfb_gemm_nt_workspace(s->C_out, C_cat);
        PCFB_REQUIRE(workspace && workspace_bytes >= p_bytes + nt_bytes, "pcfb_pconv_forward: workspace too small");
        float *P = out_p ? out_p : static_cast<float *>(workspace);
        int rc;
        if ((rc = pconv_mid1_forward_p(s, feats, nei, weights, additional, P, st))) return rc;
        return pcfb_gemm_nt(P, C_cat, lin_w, C_cat, 0, lin_b, out_y, s->C_out, s->n_out, s->C_out, C_cat, 0,
                            static_cast<char *>(workspace) + p_bytes, nt_bytes, stream);
    }
    const bool small_compose = variant == 0 && lin_w && compose_path(s) && point_path(s);
    if (!small_compose && (variant == 0 || variant == 4)) {
        const bool aligned = ((uintptr_t)feats % 16 == 0) && (s->C_add == 0 || (uintptr_t)additional % 16 == 0) &&
                             (s->H == 0 || (uintptr_t)guidance % 16 == 0);
        if (lin_w && (variant == 4 || aligned) && pconv_forward_ws_supported(s, true))
            return pconv_forward_ws(s, feats, nei, weights, additional, guidance, lin_w, lin_b, out_y, out_p,
                                    workspace, workspace_bytes, st);
        if (variant == 4) {
            set_error("pcfb_pconv_forward: the warp-specialised variant does not support this shape");
            return PCFB_ERR_UNSUPPORTED;
        }
    }
    if (variant == 0 && lin_w && compose_path(s)) {
        const int KK = (s->C_in + s->C_add) * s->C_mid;
        const size_t p_bytes = align_up((size_t)s->n_out * KK * sizeof(float), 256);
        const size_t nt_bytes = pcfb_gemm_nt_workspace(s->C_out, KK);
        PCFB_REQUIRE(workspace && workspace_bytes >= p_bytes + nt_bytes, "pcfb_pconv_forward: workspace too small");
        float *P = out_p ? out_p : static_cast<float *>(workspace);
        int rc;
        if (point_path(s)) rc = pconv_point_fwd_p(s, feats, nei, weights, additional, guidance, P, st);
        else if (s->n_out <= 16384 && pconv_p_small_supported(s, weights, P)) rc = pconv_p_small(s, feats, nei, weights, additional, guidance, P, st);
        else if (pconv_midn_supported(s) && ((uintptr_t)feats % 16 == 0) && ((uintptr_t)P % 16 == 0) &&
                 (s->C_add == 0 || (uintptr_t)additional % 16 == 0))
            rc = pconv_midn_forward_p(s, feats, nei, weights, additional, guidance, P, st);      // C_mid 2..4: streaming weighted sums
        else rc = pconv_forward_simt(s, feats, nei, weights, additional, guidance, nullptr, nullptr, nullptr, P, st);
        if (rc) return rc;
        return pcfb_gemm_nt(P, KK, lin_w, KK, 0, lin_b, out_y, s->C_out, s->n_out, s->C_out, KK, 0,
                            static_cast<char *>(workspace) + p_bytes, nt_bytes, stream);
    }
    // contraction only (the unfused layers, PCONV_OPT: False), C_mid 2..4, more points than the small-level kernel takes
    if (variant == 0 && !lin_w && out_p && s->n_out > 16384 && pconv_midn_supported(s) && ((uintptr_t)feats % 16 == 0) &&
        ((uintptr_t)out_p % 16 == 0) && (s->C_add == 0 || (uintptr_t)additional % 16 == 0))
        return pconv_midn_forward_p(s, feats, nei, weights, additional, guidance, out_p, st);
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

// ---- backward: variant 0/2 = tensor-core composition, variant 1 = single fused SIMT kernel ----------------
// With a Linear, the two dense products run on tcgen05 (3xTF32): dP = dY W (pcfb_gemm_nt, column blocks of 128)
// and dW = dY^T P, db = column sums of dY (pcfb_gemm_tn, deterministic split over points); the per-point part
// (dw, dadd, dguidance, per-edge dE) then runs as the SIMT contraction-backward kernel on dP, and dx is the CSR
// segment sum over the inverse map.  Without a Linear the incoming gradient already is dP.
namespace pcfb {
struct BwdCompose { float *dP; void *nt_ws; size_t nt_bytes; void *tn_ws; size_t tn_bytes; float *p_re; void *simt_ws; size_t simt_bytes, bytes; };
static BwdCompose carve_compose(void *ws, const pcfb_pconv_shape &s, bool need_w, bool need_p) {
    Carver c(ws);
    BwdCompose w{};
    const int KK = (s.C_in + s.C_add) * s.C_mid;
    w.dP = c.take<float>((size_t)s.n_out * KK);
    w.nt_bytes = pcfb_gemm_nt_workspace((s.C_in + s.C_add) * s.C_mid, s.C_out);
    w.nt_ws = c.take<char>(w.nt_bytes);
    if (need_w) {
        w.tn_bytes = pcfb_gemm_tn_workspace(s.n_out, s.C_out, KK, 1);
        w.tn_ws = c.take<char>(w.tn_bytes);
        if (need_p) w.p_re = c.take<float>((size_t)s.n_out * KK);
    }
    pcfb_pconv_shape nolin = s; nolin.C_out = 0;
    w.simt_bytes = pconv_backward_simt_workspace(&nolin);
    w.simt_ws = c.take<char>(w.simt_bytes);
    w.bytes = align_up(c.off, 256);
    return w;
}
}  // namespace pcfb

extern "C" size_t pcfb_pconv_backward_workspace(const pcfb_pconv_shape *s, int variant)
{
    if (!s) return 0;
    const size_t fused = pcfb::pconv_backward_simt_workspace(s);
    if (variant == 1 || s->C_out <= 0 || s->C_out > 256) return fused;
    const size_t comp = pcfb::carve_compose(nullptr, *s, true, true).bytes;
    return comp > fused ? comp : fused;
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
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (variant != 1 && !lin_w && grad_p && s->n_out > 0 && pconv_bwd2_supported(s)) {
        pcfb_pconv_shape nolin0 = *s;
        nolin0.C_out = 0;
        const size_t need = pconv_backward_simt_workspace(&nolin0);
        PCFB_REQUIRE(!grad_feats || (workspace && workspace_bytes >= need), "pcfb_pconv_backward: workspace too small");
        PCFB_REQUIRE(!grad_feats || (inv_neighbors && inv_k && inv_idx), "pcfb_pconv_backward: grad_feats needs the inverse map");
        float *grad_edge = grad_feats ? static_cast<float *>(workspace) : nullptr;
        int rc0 = contraction_backward(s, grad_p, feats, nei, weights, additional, guidance, grad_weights, grad_additional, grad_guidance,
                                       grad_edge, st);
        if (rc0 || !grad_feats) return rc0;
        return pcfb_gather_backward(grad_edge, inv_neighbors, inv_k, inv_idx, s->n_in, s->n_out, s->K, s->C_in, grad_feats, stream);
    }
    if (variant == 1 || !lin_w || s->C_out > 256 || s->n_out == 0)
        return pconv_backward_simt(s, grad_y, grad_p, feats, nei, inv_neighbors, inv_k, inv_idx, weights, additional,
                                   guidance, lin_w, pconv_out, grad_feats, grad_weights, grad_additional, grad_guidance,
                                   grad_lin_w, grad_lin_b, workspace, workspace_bytes, st);
    PCFB_REQUIRE(grad_y != nullptr, "pcfb_pconv_backward: missing incoming gradient");
    const bool need_w = grad_lin_w || grad_lin_b;
    const bool need_p = need_w && !pconv_out;
    BwdCompose w = carve_compose(workspace, *s, need_w, need_p);
    if (!workspace || workspace_bytes < w.bytes) {
        set_error("pcfb_pconv_backward: workspace %zu < %zu", workspace_bytes, w.bytes);
        return PCFB_ERR_WORKSPACE;
    }
    const int KK = (s->C_in + s->C_add) * s->C_mid;
    int rc;
    // dP = dY * W  (one launch; column blocks of the KK outputs ride on blockIdx.y)
    if ((rc = pcfb_gemm_nt(grad_y, s->C_out, lin_w, KK, 1, nullptr, w.dP, KK, s->n_out, KK, s->C_out, 0,
                           w.nt_ws, w.nt_bytes, stream))) return rc;
    if (need_w) {
        const float *P = pconv_out;
        if (need_p) {
            if (pconv_mid1_supported(s)) rc = pconv_mid1_forward_p(s, feats, nei, weights, additional, w.p_re, st);
            else rc = pconv_forward_simt(s, feats, nei, weights, additional, guidance, nullptr, nullptr, nullptr, w.p_re, st);
            if (rc) return rc;
            P = w.p_re;
        }
        if ((rc = pcfb_gemm_tn(grad_y, s->C_out, P, KK, grad_lin_w, KK, grad_lin_b, s->n_out, s->C_out, KK, w.tn_ws, w.tn_bytes, stream))) return rc;
    }
    if (pconv_mid1_supported(s))
        return pconv_mid1_backward(s, w.dP, feats, nei, inv_neighbors, inv_k, inv_idx, weights, additional, grad_feats,
                                   grad_weights, grad_additional, st);
    if (pconv_bwd2_supported(s)) {
        float *grad_edge = grad_feats ? static_cast<float *>(w.simt_ws) : nullptr;     // [n_out, K, C_in] scratch
        PCFB_REQUIRE(!grad_feats || (inv_neighbors && inv_k && inv_idx), "pcfb_pconv_backward: grad_feats needs the inverse map");
        if ((rc = contraction_backward(s, w.dP, feats, nei, weights, additional, guidance, grad_weights, grad_additional, grad_guidance,
                                       grad_edge, st))) return rc;
        if (grad_feats)
            return pcfb_gather_backward(grad_edge, inv_neighbors, inv_k, inv_idx, s->n_in, s->n_out, s->K, s->C_in, grad_feats, stream);
        return PCFB_OK;
    }
    pcfb_pconv_shape nolin = *s;
    nolin.C_out = 0;
    return pconv_backward_simt(&nolin, nullptr, w.dP, feats, nei, inv_neighbors, inv_k, inv_idx, weights, additional, guidance,
                               nullptr, nullptr, grad_feats, grad_weights, grad_additional, grad_guidance, nullptr, nullptr,
                               w.simt_ws, w.simt_bytes, st);
}
