/* pcf_b200.h -- C ABI of libpcf_b200.so: the B200 (sm_100a) hot path of PointConvFormer.
 *
 * Drop-in boundary for the reference's native extension `pcf_cuda`
 * (/root/reference/cpp_wrappers/cpp_pcf_kernel/pcf_cuda.cpp:9-19, signatures include/pcf.h:38-250) and
 * for the third-party kNN it calls (/root/reference/knn_post_dataloader_utils.py:22-41) and its CPU
 * grid subsampling (/root/reference/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:9-110).
 *
 * Conventions (all entry points):
 *   - plain C, no torch types; every pointer is a DEVICE pointer unless named h_*;
 *   - the caller owns all memory (outputs and workspaces are caller-allocated); nothing is allocated,
 *     freed or synchronised inside; all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - tensors are dense row-major ("contiguous"), batch dimension folded by the caller (B == 1 packed
 *     representation, layers.py:216); indices are int64 like the reference's (datasetCommon.py:178-180);
 *   - return value 0 = success, otherwise an error code; pcfb_last_error() gives the message of the
 *     last failure on the calling thread.  The reference raises RuntimeError through TORCH_CHECK
 *     (include/pcf.h:14-24); the Python host mirror re-raises RuntimeError from these codes;
 *   - neighbour entries outside [0, n_in) (e.g. the -1 padding preserved by listToBatch,
 *     knn_post_dataloader_utils.py:131-148) contribute zero / are skipped, as the reference's opt
 *     kernels bounds-check (pconv_ops.cu:453,494; knn.cu:38,75).
 *   - pconv_out / "P" channel layout is c*C_mid + j (layers.py:713-716, pconv_ops.cu:89-90); gradients
 *     follow autograd through that layout (NOT the reference CUDA backward's mid*(C)+c indexing, which
 *     is inconsistent with its own forward -- SURVEY.md trap T1).
 */
#ifndef PCF_B200_H
#define PCF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    PCFB_OK = 0,
    PCFB_ERR_ARG = 1,        /* bad shape / null pointer / unsupported size */
    PCFB_ERR_CUDA = 2,       /* a CUDA runtime call failed (launch error etc.) */
    PCFB_ERR_WORKSPACE = 3,  /* workspace too small */
    PCFB_ERR_UNSUPPORTED = 4 /* shape not supported by the requested kernel variant */
};

const char *pcfb_last_error(void);
/* ABI / build info: "pcf_b200 <abi> sm_100a" */
const char *pcfb_version(void);
/* Number of kernel launches enqueued by this library since load (for bench.py's gpu_launches). */
uint64_t pcfb_launch_count(void);
/* Programmatic dependent launch for every kernel of the library (default on; env PCFB_PDL=0): each kernel starts with
 * griddepcontrol.wait and is launched with programmaticStreamSerializationAllowed, hiding the launch latency between the
 * ~1800 dependent kernels of a training step.  Returns the previous setting. */
int pcfb_set_pdl(int on);

/* ---------------------------------------------------------------------------------------------
 * kNN on packed scenes.  Replaces knn_keops (knn_post_dataloader_utils.py:22-41) + the per-scene loop
 * of compute_knn_packed (171-223) + the offsetting of listToBatch/prepare (113-167) for ONE edge set.
 *
 * ref_xyz [n_ref,3], qry_xyz [n_qry,3] fp32 packed over `n_seg` scenes; ref_off / qry_off are
 * int32[n_seg+1] exclusive prefix sums of the per-scene counts.  For every query q of scene s:
 *   out[q, 0..K) = ref_off[s] + the K references r of scene s with the smallest (d, r), ascending,
 *   d = ((qx-rx)^2 + (qy-ry)^2) + (qz-rz)^2 evaluated in fp32 with no FMA contraction.
 * If a scene has n < K references the n found neighbours are repeated cyclically (the reference draws
 * random indices there, knn_post_dataloader_utils.py:58-66 -- not reproducible).
 * 1 <= K <= 255 (inv_k is uint8 downstream, knn.cu:114).  Brute force, exact.
 * ------------------------------------------------------------------------------------------- */
int pcfb_knn_packed(const float *ref_xyz, const int32_t *ref_off, const float *qry_xyz,
                    const int32_t *qry_off, int n_seg, int n_ref, int n_qry, int K,
                    int64_t *out_idx, void *stream);

/* Same results, bit for bit, through a uniform grid over the reference cloud (exact
 * grid-accelerated kNN).  pcfb_knn_grid_build bins the references of every scene (cell edge = cell_hint,
 * enlarged on the device until each scene's dense grid has <= 2*n+64 cells; cell_hint <= 0 = automatic) into
 * the workspace; pcfb_knn_grid_query answers any number of query sets against it (1 <= K <= 64).  The grid
 * of one level serves its self-, forward- and propagate- edge sets.
 * qry_order (optional): a permutation of the queries in which neighbouring entries are neighbours in space -- the lanes
 * of a warp then walk the same cells.  NULL = natural order, except self queries (qry_xyz == ref_xyz), which use the
 * grid's own cell order.  pcfb_knn_grid_order returns that order of a built grid (a device pointer INTO its workspace,
 * n_ref entries: the reference indices sorted by cell) so that the grid of level l can order level l's points when they
 * are the queries against another level's grid.  The result does not depend on the order. */
size_t pcfb_knn_grid_workspace(int n_seg, int n_ref);
int pcfb_knn_grid_build(const float *ref_xyz, const int32_t *ref_off, int n_seg, int n_ref, float cell_hint,
                        void *workspace, size_t workspace_bytes, void *stream);
const int32_t *pcfb_knn_grid_order(int n_seg, int n_ref, const void *workspace);
int pcfb_knn_grid_query(const float *ref_xyz, int n_seg, int n_ref, const float *qry_xyz, const int32_t *qry_off,
                        int n_qry, int K, const int32_t *qry_order, int64_t *out_idx, const void *workspace,
                        size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * kNN inverse map (CSR transpose).  Replaces pcf_cuda.compute_knn_inverse
 * (include/pcf.h:183-186 -> src/knn.cu:104-168).
 * nei [n_out,K] int64; outputs inv_neighbors int32[n_out*K], inv_k uint8[n_out*K], inv_idx
 * int32[total+1].  Segment p = [inv_idx[p], inv_idx[p+1]) lists the (n, k) with nei[n,k] == p in
 * ascending (n, k) order (deterministic; the reference's order inside a segment is atomic-claim
 * order).  Unused tail entries are zero.  Workspace: pcfb_knn_inverse_workspace(n_out, K, total).
 * ------------------------------------------------------------------------------------------- */
size_t pcfb_knn_inverse_workspace(int n_out, int K, int total);
int pcfb_knn_inverse(const int64_t *nei, int n_out, int K, int total, int32_t *inv_neighbors,
                     uint8_t *inv_k, int32_t *inv_idx, void *workspace, size_t workspace_bytes,
                     void *stream);

/* ---------------------------------------------------------------------------------------------
 * Neighbour gather family (index_points, layer_utils.py:13-30, and the strided max-pool shortcut
 * layers.py:403-408,728-733).
 * ------------------------------------------------------------------------------------------- */
/* out[m,k,:] = feats[nei[m,k],:]   feats [n_in,C] -> out [n_out,K,C] */
int pcfb_gather(const float *feats, const int64_t *nei, int n_in, int n_out, int K, int C, float *out,
                void *stream);
/* grad_feats[p,:] = sum over (n,k) in inverse segment p of grad_out[n,k,:]   (no atomics) */
int pcfb_gather_backward(const float *grad_out, const int32_t *inv_neighbors, const uint8_t *inv_k,
                         const int32_t *inv_idx, int n_in, int n_out, int K, int C, float *grad_feats,
                         void *stream);
/* out[m,c] = max_k feats[nei[m,k],c]; arg[m,c] = the k attaining it (first on ties) */
int pcfb_gather_max(const float *feats, const int64_t *nei, int n_in, int n_out, int K, int C, float *out,
                    uint8_t *arg, void *stream);
int pcfb_gather_max_backward(const float *grad_out, const uint8_t *arg, const int32_t *inv_neighbors,
                             const uint8_t *inv_k, const int32_t *inv_idx, int n_in, int n_out, int K,
                             int C, float *grad_feats, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Edge geometry: localized xyz and the 12-d viewpoint-invariant features
 * (VI_coordinate_transform, layer_utils.py:176-231; prologue layers.py:337-353,660-682,846-863,1024-1037).
 * xyz_in/nrm_in [n_in,3] are the gathered cloud, xyz_out/nrm_out [n_out,3] the centres.
 * out_r [n_out,K,3] (may be NULL), out_vi [n_out,K,12] (may be NULL; needs the normals).
 * ------------------------------------------------------------------------------------------- */
int pcfb_edge_geometry(const float *xyz_in, const float *nrm_in, const float *xyz_out, const float *nrm_out,
                       const int64_t *nei, int n_in, int n_out, int K, float *out_r, float *out_vi,
                       void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused PointConv / PointConvFormer contraction (+ Linear).
 *   G[m,k,c]  = c <  C_in : feats[nei[m,k],c] * (guidance ? guidance[m,k,c % H] : 1)
 *               c >= C_in : additional[m,k,c-C_in]
 *   P[m,c*C_mid+j] = sum_k G[m,k,c] * weights[m,k,j]                       (C_cat = C_in + C_add)
 *   Y[m,o]   = sum_kk P[m,kk] * lin_w[o,kk] + lin_b[o]
 * Replaces pconv_linear_cutlass_forward (pcf.h:243-250 -> pconv_ops.cu:969-1269), pconv_linear_forward
 * (pcf.h:131-138), pconv_forward (pcf.h:81-86; lin_w == NULL -> only P) and pcf_forward (pcf.h:38-43;
 * guidance != NULL).  out_y [n_out,C_out] (NULL iff lin_w NULL), out_p [n_out,C_cat*C_mid] (may be NULL
 * when lin_w given).  `variant`: 0 = auto, 1 = exact fp32 SIMT, 2 = tcgen05 (3xTF32 split Linear; pipelined
 * kernel when the tile fits in shared memory, else the simple one), 3 = tcgen05 simple kernel (bisecting),
 * 4 = warp-specialised tcgen05 kernel only (K = 16, C_mid = 16, C_out % 16 == 0, C_out <= 128; what auto prefers).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int n_in, n_out, K, C_in, C_add, C_mid, C_out, H; /* H = guidance heads (0 if none) */
} pcfb_pconv_shape;

/* 1 if `variant` can run this shape (the tcgen05 variant needs C_mid in {1,4,8,16}, C_out % 8 == 0, a Linear,
 * C_cat*C_mid % 4 == 0 and a tile that fits in shared memory), else 0. */
int pcfb_pconv_forward_supported(const pcfb_pconv_shape *s, int variant);
size_t pcfb_pconv_forward_workspace(const pcfb_pconv_shape *s, int variant);
int pcfb_pconv_forward(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei,
                       const float *weights, const float *additional, const float *guidance,
                       const float *lin_w, const float *lin_b, float *out_y, float *out_p,
                       void *workspace, size_t workspace_bytes, int variant, void *stream);

/* Backward of the above (autograd semantics).  Replaces pconv_linear_opt_backward (pcf.h:213-224 ->
 * pconv_ops.cu:863-948), pconv_linear_backward, pconv_backward, pcf_backward (pcf.h:60-66).
 * grad_y [n_out,C_out] when lin_w != NULL, else grad_p [n_out,C_cat*C_mid] is the incoming gradient.
 * pconv_out: the P saved by the forward (the reference saves it, layer_utils.py:52-54); may be NULL,
 * then P is recomputed into the workspace when grad_lin_w / grad_lin_b are requested.
 * Outputs (any may be NULL to skip): grad_feats [n_in,C_in] (needs the inverse map; summed per input
 * point over its inverse segment, no atomics), grad_weights [n_out,K,C_mid], grad_additional
 * [n_out,K,C_add], grad_guidance [n_out,K,H], grad_lin_w [C_out,C_cat*C_mid], grad_lin_b [C_out].
 * variant 0/2: dP = dY W and dW = dY^T P run as 3xTF32 tcgen05 GEMMs, the per-point part as a CUDA-core kernel on
 * dP; variant 1: one fused exact-fp32 CUDA-core kernel (bisecting reference).
 * ------------------------------------------------------------------------------------------- */
size_t pcfb_pconv_backward_workspace(const pcfb_pconv_shape *s, int variant);
int pcfb_pconv_backward(const pcfb_pconv_shape *s, const float *grad_y, const float *grad_p,
                        const float *feats, const int64_t *nei, const int32_t *inv_neighbors,
                        const uint8_t *inv_k, const int32_t *inv_idx, const float *weights,
                        const float *additional, const float *guidance, const float *lin_w,
                        const float *pconv_out, float *grad_feats, float *grad_weights, float *grad_additional,
                        float *grad_guidance, float *grad_lin_w, float *grad_lin_b, void *workspace,
                        size_t workspace_bytes, int variant, void *stream);

/* ---------------------------------------------------------------------------------------------
 * fp32-accurate dense products on tcgen05 (3xTF32 operand split, TMEM accumulators) for the Linear layers
 * around the contraction (Linear_BN / UnaryBlock, layer_utils.py:241-319; torch runs them as SIMT sgemm) and
 * for the two dense products of the fused backward (dP = dY W, dW = dY^T P; pconv_ops.cu:434-440,516-533).
 *   pcfb_gemm_nt : C[M,N] = act(A[M,K] * Wt + bias);  W is [N,K] (w_is_kn = 0, y = x W^T) or [K,N]
 *                  (w_is_kn = 1, dx = dy W); any N >= 1 (wide outputs run as column blocks of one launch);
 *                  act 0 none / 1 ReLU / 2 LeakyReLU(0.1).
 *   pcfb_gemm_tn : C[N1,N2] = A[M,N1]^T * B[M,N2], optionally rowsum[N1] = sum_m A[m,:] (bias gradient);
 *                  1 <= N1 <= 256; reduction over M split across CTAs, partials summed in fixed order.
 * ------------------------------------------------------------------------------------------- */
size_t pcfb_gemm_nt_workspace(int N, int K);
/* pcfb_gemm_nt = pcfb_gemm_nt_prepare (weights -> split tf32 operands in tensor-core order, into the workspace; depends on
 * the weights and on M only through the column-block width) + pcfb_gemm_nt_prepared (the product).  The two may run on
 * different streams: a Linear's weights are ready long before its input. */
int pcfb_gemm_nt_prepare(const float *W, int ldw, int w_is_kn, int M, int N, int K, void *workspace, size_t workspace_bytes,
                         void *stream);
int pcfb_gemm_nt_prepared(const float *A, int lda, const float *bias, float *C, int ldc, int M, int N, int K, int act,
                          const void *workspace, size_t workspace_bytes, void *stream);
int pcfb_gemm_nt(const float *A, int lda, const float *W, int ldw, int w_is_kn, const float *bias, float *C, int ldc,
                 int M, int N, int K, int act, void *workspace, size_t workspace_bytes, void *stream);
size_t pcfb_gemm_tn_workspace(int M, int N1, int N2, int with_rowsum);
int pcfb_gemm_tn(const float *A, int lda, const float *B, int ldb, float *C, int ldc, float *rowsum,
                 int M, int N1, int N2, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused small MLP layers with train-mode BatchNorm for the per-edge / per-point chains: WeightNet
 * (layers.py:127-191), positional-encoding WeightNet (layers.py:575-577), mlp_conv (layers.py:240-243), the
 * guidance MLP (layers.py:38-68) and UnaryBlock (layer_utils.py:281-319).  One layer = one streaming pass:
 *   y[E,cout] = act_in(x*in_scale + in_shift) W^T + b     (the BatchNorm of the layer below is folded into the load)
 * while sum(y-b), sum((y-b)^2) are accumulated per block (stat_partial [blocks][2][cout]); pcfb_bn_finalize
 * turns them (reduced in double, fixed order) into the layer's own (scale, shift) = (gamma*invstd,
 * beta - mean*gamma*invstd), updates the running statistics (momentum, unbiased variance) and saves mean / invstd.
 * act codes: 0 none, 1 ReLU, 2 LeakyReLU(0.1), 3 sigmoid.  Sizes: cin, cout <= 64 (pcfb_mlp_supported).
 * Backward per layer (pcfb_mlp_backward): dz = dA*act'(z), dy = scale*(dz - S1/E - xhat*S2/E) with sums = [S1|S2]
 * from pcfb_mlp_backward_stats (or produced on the fly by the layer above through prev_sums); outputs dA_prev,
 * dW, db; dgamma = S2, dbeta = S1.  All reductions are block partials summed in fixed order: deterministic.
 * d_count (device double, may be NULL) overrides the row count with a global one (SyncBatchNorm across ranks).
 * ------------------------------------------------------------------------------------------- */
int pcfb_mlp_supported(int cin, int cout);
size_t pcfb_mlp_workspace(int64_t E, int cin, int cout);
int pcfb_mlp_forward(const float *x, int ldx, int64_t E, int cin, int cout, const float *W, const float *b,
                     const float *in_scale, const float *in_shift, int in_act, float *y, int ldy,
                     float *stat_partial, int *h_nblocks, void *stream);
/* Inference-mode chain of three Linear (+ BatchNorm with running statistics, given as per-channel scale / shift; NULL =
 * folded or absent) + activation layers as ONE kernel: the WeightNet of a PointConv / PointConvFormer layer
 * (layers.py:127-171) under model.eval().  W, b, scale, shift: arrays of three device pointers (W[l] row-major
 * [c_{l+1}][c_l]); act: three activation codes.  Supported: c0 in {1..4, 12}, c1 = c2 = 8, c3 = 16. */
/* eval-mode BatchNorm as an affine map: scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale
 * (gamma / beta / invstd optional); one launch instead of torch's five per BatchNorm and forward. */
int pcfb_bn_eval_affine(const float *gamma, const float *beta, const float *running_mean, const float *running_var,
                        float eps, int C, float *scale, float *shift, float *invstd, void *stream);
int pcfb_mlp_chain_eval_supported(int c0, int c1, int c2, int c3);
int pcfb_mlp_chain_eval(const float *x, int ldx, int64_t E, int c0, int c1, int c2, int c3,
                        const float *const *W, const float *const *b, const float *const *scale,
                        const float *const *shift, const int *act, float *out, int ldo, void *stream);
int pcfb_bn_act(const float *y, int64_t rows, int C, const float *scale, const float *shift, int act, float *out,
                const float *residual /* optional [rows, C] */, int residual_after_act, void *stream);
int pcfb_mlp_backward_stats(const float *dA, int ldd, const float *y, int ldy, int64_t E, int C, const float *scale,
                            const float *shift, const float *mean, const float *invstd, int act,
                            float *sums /* NULL: leave the block partials [*h_nblocks][2][C] in the workspace for pcfb_bn_reduce_sums */,
                            int *h_nblocks, void *workspace, size_t workspace_bytes, void *stream);
int pcfb_mlp_backward(const float *dA, int ldd, const float *y, int ldy, int64_t E, int cin, int cout,
                      const float *W, const float *scale, const float *shift, const float *mean,
                      const float *invstd, const float *sums, int act,
                      const float *x_prev, int ldx, const float *in_scale, const float *in_shift, int in_act,
                      const float *prev_mean, const float *prev_invstd,
                      float *dA_prev, int ldp, float *prev_sums, float *dW, float *db, const double *d_count,
                      void *workspace, size_t workspace_bytes, void *stream);
int pcfb_sum_partials(const float *partial, int nblocks, int n, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * BatchNorm statistics: block partials -> per-channel sums -> [SyncBatchNorm exchange] -> scale / shift / running
 * statistics in ONE kernel (csrc/peer_reduce.cu).  Replaces torch.nn.BatchNorm's finalize and, across ranks,
 * torch.nn.SyncBatchNorm's all_gather / all_reduce of per-channel statistics (convert_sync_batchnorm,
 * train_ScanNet_DDP_WarmUP.py:190-195; sync_bn: True in configs/configPCF_Opt_10cm.yaml): ~540 latency-bound messages
 * per step.  partial: [nblocks][2][C] (sum, sum of squares -- pivoted by `pivot` -- or sum dz, sum dz*xhat).
 * peer_bases (NULL = single process): device array of `world` pointers, entry r = rank r's symmetric buffer of
 * pcfb_syncbn_buffer_bytes(world) bytes as mapped into THIS process (all-zero before the first call).  One CTA per 8
 * channels pushes its sums + the local row count into its slot of every rank's buffer, publishes an epoch flag
 * (st.release.sys), waits for the peers' flags (ld.acquire.sys) and adds the slots in rank order (bit-identical on all
 * ranks).  Every rank must issue the same sequence of calls per `channel` (< pcfb_syncbn_channels()); calls of one channel
 * must be stream ordered.  timeout_s > 0: a CTA that waited that long sets the buffer's error word (byte 0), writes NaN
 * and returns (no trap); 0 = wait forever.
 *   pcfb_bn_finalize: (scale, shift) = (gamma*invstd, beta - mean*gamma*invstd), mean / invstd saved, running statistics
 *     updated with the GLOBAL count (momentum < 0: cumulative average 1/num_batches_tracked, counter not touched; else
 *     batches_tracked += 1), count_out (optional) = global row count (device double, consumed by the backward kernels).
 *     d_count (optional) = global count already on the device (NCCL fallback: partial then holds the all-reduced sums).
 *   pcfb_bn_reduce_sums: sums_local[2][C] (dgamma / dbeta stay local, like torch.nn.SyncBatchNorm) and sums_global[2][C].
 * ------------------------------------------------------------------------------------------- */
size_t pcfb_syncbn_buffer_bytes(int world);
int pcfb_syncbn_channels(void);
int pcfb_bn_finalize(const float *partial, int nblocks, int C, int64_t count, const double *d_count, const float *pivot,
                     const float *gamma, const float *beta, float eps, float momentum, float *running_mean,
                     float *running_var, float *scale, float *shift, float *mean, float *invstd,
                     int64_t *batches_tracked, double *count_out, const void *peer_bases, int rank, int world, int channel,
                     double timeout_s, void *stream);
int pcfb_bn_reduce_sums(const float *partial, int nblocks, int C, float *sums_local, float *sums_global,
                        const void *peer_bases, int rank, int world, int channel, double timeout_s, void *stream);

/* BatchNorm (+ residual + activation) of a SMALL [rows, C] tensor (rows <= pcfb_bn_small_max_rows(), C % 4 == 0) as ONE
 * kernel per direction: a CTA owns 8 channels for all rows, so statistics, the SyncBatchNorm exchange (same protocol and
 * arguments as pcfb_bn_finalize / pcfb_bn_reduce_sums), the finalize and the apply pass need no second launch -- replaces
 * pcfb_bn_stats -> pcfb_bn_finalize -> pcfb_bn_act and pcfb_bn_backward_stats -> pcfb_bn_reduce_sums -> pcfb_bn_backward
 * on the coarse levels of the pyramid, where every launch is a bubble on the step's critical path.  Same formulas:
 *   forward : out = act(x*scale + shift (+ residual)) (+ residual after the activation), scale / shift / mean / invstd /
 *             running statistics / batches_tracked / count_out as pcfb_bn_finalize;
 *   backward: dz = dA * act'(.), sums_local[2][C] = this rank's (sum dz | sum dz*xhat) = (dbeta | dgamma),
 *             dX = scale * (dz - S1/E - xhat*S2/E) with the GLOBAL sums and count (d_count, NULL = rows), d_residual = dz. */
int pcfb_bn_small_max_rows(void);
int pcfb_bn_small_forward(const float *x, int64_t rows, int C, const float *pivot, const float *gamma, const float *beta,
                          float eps, float momentum, float *running_mean, float *running_var, int64_t *batches_tracked,
                          int act, const float *residual, int residual_after_act, float *out, float *scale, float *shift,
                          float *mean, float *invstd, double *count_out, const void *peer_bases, int rank, int world,
                          int channel, double timeout_s, void *stream);
int pcfb_bn_small_backward(const float *dA, const float *x, int64_t rows, int C, const float *scale, const float *shift,
                           const float *mean, const float *invstd, int act, const float *residual, const double *d_count,
                           float *sums_local, float *dX, float *d_residual, const void *peer_bases, int rank, int world,
                           int channel, double timeout_s, void *stream);

/* ---------------------------------------------------------------------------------------------
 * BatchNorm (+ activation) over a contiguous [rows, C] tensor, C % 4 == 0, C <= 1024: the BatchNorm + ReLU that
 * follows the fused contraction (layers.py:708-709, 721, 893-898, 1086-1092; `self.bn` / `linear.bn`) and the
 * Linear_BN (+ LeakyReLU) of the wide per-point blocks (layer_utils.py:241-319), replacing torch's
 * batch_norm + activation kernels (5 passes forward, 8 backward) by 3 and 5.
 *   forward (train): pcfb_bn_stats -> partial[blocks][2][C] = sum(x - pivot), sum((x - pivot)^2)  (pivot: any
 *     per-channel offset near the mean, e.g. the Linear bias; may be NULL) -> pcfb_bn_finalize (scale, shift,
 *     running statistics, mean, invstd) -> pcfb_bn_act (out = act(x*scale + shift)).  Eval: pcfb_bn_act only.
 *   backward: pcfb_bn_backward_stats -> sums[2][C] = (sum dz, sum dz*xhat) = (dbeta, dgamma), dz = dA*act'(z);
 *     pcfb_bn_backward: dX = scale*(dz - S1/E - xhat*S2/E) (sums == NULL: eval mode, dX = scale*dz).
 *     d_count (device double, may be NULL) = global row count for SyncBatchNorm.
 *   residual (optional, [rows, C]): pcfb_bn_act computes act(x*scale + shift + r) (residual_after_act == 0: the tail of a
 *     PointConvFormer / PointConvStridePE block, leaky_relu(unary2(h) + shortcut), layers.py:413-415, 737-739) or
 *     act(x*scale + shift) + r (the decoder's skip connection, layers.py:1096-1097) in the same pass; for the first form the
 *     backward kernels take r to evaluate act' and pcfb_bn_backward also writes d_residual = dz.
 * Every reduction is block partials summed in fixed order (deterministic).
 * ------------------------------------------------------------------------------------------- */
int pcfb_bn_supported(int C);
size_t pcfb_bn_workspace(int64_t rows, int C);
int pcfb_bn_stats(const float *x, int64_t rows, int C, const float *pivot, float *partial, size_t partial_bytes,
                  int *h_nblocks, void *stream);
int pcfb_bn_backward_stats(const float *dA, const float *y, int64_t rows, int C, const float *scale, const float *shift,
                           const float *mean, const float *invstd, int act, const float *residual,
                           float *sums /* NULL: partials stay in the workspace */, int *h_nblocks, void *workspace,
                           size_t workspace_bytes, void *stream);
int pcfb_bn_backward(const float *dA, const float *y, int64_t rows, int C, const float *scale, const float *shift,
                     const float *mean, const float *invstd, const float *sums, int act, const double *d_count,
                     const float *residual, float *dX, float *d_residual, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Grid (voxel) subsampling with barycentres on packed scenes.  Replaces grid_subsampling()
 * (grid_subsampling.cpp:9-110) as called per level by subsample() (datasetCommon.py:384-420).
 * Three phases because the grid extent and the output size are data dependent (the host reads two
 * small arrays in between):
 *   pcfb_gridsub_bounds : per scene, origin = floor(min*(1/dl))*dl (float[n_seg*3]) and grid dims
 *                         N = floor((max-origin)/dl)+1 (int32[n_seg*3]), the reference's fp32
 *                         arithmetic (grid_subsampling.cpp:28-36).  workspace >= n_seg*24 bytes.
 *   pcfb_gridsub_count  : bins every point into the dense cell cell_off[s] + iX + NX*iY + NX*NY*iZ
 *                         (cell_off int32[n_seg+1] = host prefix sum of NX*NY*NZ, total_cells its
 *                         last entry), builds the cell->points CSR and writes the number of occupied
 *                         cells per scene to out_counts int32[n_seg].
 *   pcfb_gridsub_emit   : writes barycentres (sum * float(1.0/count)) and feature means (sum/count)
 *                         in ascending (scene, voxel key) order; sums accumulate sequentially in
 *                         input order, so results are bit-identical to the reference's running
 *                         sums (its own output ORDER is unordered_map iteration order, which is not
 *                         reproducible; ascending key is the canonical order here).
 * count and emit share the same workspace (pcfb_gridsub_workspace) which must stay untouched between.
 * ------------------------------------------------------------------------------------------- */
size_t pcfb_gridsub_workspace(int n_seg, int n_pts, int64_t total_cells);
int pcfb_gridsub_bounds(const float *xyz, const int32_t *seg_off, int n_seg, int n_pts, float dl,
                        float *out_origin, int32_t *out_dims, void *workspace, size_t workspace_bytes,
                        void *stream);
int pcfb_gridsub_count(const float *xyz, const int32_t *seg_off, int n_seg, int n_pts, float dl,
                       const float *origin, const int32_t *dims, const int32_t *cell_off,
                       int64_t total_cells, int32_t *out_counts, void *workspace, size_t workspace_bytes,
                       void *stream);
int pcfb_gridsub_emit(const float *xyz, const float *feats, int n_seg, int n_pts, int F,
                      int64_t total_cells, float *out_xyz, float *out_feats, void *workspace,
                      size_t workspace_bytes, void *stream);

/* Output-point count up to which pcfb_pconv_forward / pcfb_pconv_backward use the per-point CTA kernels of the coarse
 * pyramid levels (csrc/pconv_point.cu; K = 16, C_mid = 16) instead of the tiled ones; default 4000 (env
 * PCFB_POINT_KERNEL_MAX).  Returns the previous value.  Same results either way (exact fp32 contraction). */
int pcfb_set_point_kernel_max(int n_points);

/* ---------------------------------------------------------------------------------------------
 * Device-resident pyramid construction: raw scene -> voxelised level 0 -> grid-subsampled levels without a
 * host read per level.
 *   pcfb_voxelize: voxelize(coord, voxel_size, hash_type='ravel', mode='deterministic') of util/voxelize.py:44-70 on packed
 *     scenes: discrete = floor(coord / voxel) (float64, as numpy promotes it), key = Fortran-style ravel of discrete - per-scene min
 *     (ravel_hash_vec, voxelize.py:26-41); one point per occupied voxel -- the SMALLEST input index (the reference takes
 *     the first of an unstable argsort: implementation defined) -- emitted in ascending (scene, key) order, which is the
 *     reference's idx_sort order.  out_idx [<= n_pts] int32 packed point indices, out_seg_off [n_seg+1] device prefix of the
 *     per-scene output counts.
 *   pcfb_pyramid_level: one grid_subsampling() level (grid_subsampling.cpp:9-110, bit-identical sums, see pcfb_gridsub_*)
 *     whose INPUT size is only known on the device: seg_off [n_seg+1] is the device prefix array of the previous level
 *     (its last entry the point count, <= n_pts_max, the host's upper bound = the size of the buffers); out_seg_off feeds
 *     the next level.  The host reads all counts once, after the last level (datasetCommon.py:384-420 reads per level).
 * cells_max: host upper bound of the dense voxel table (from the level-0 bounding box); *status |= 1 if the data needs
 * more (nothing is emitted for that call).
 * ------------------------------------------------------------------------------------------- */
size_t pcfb_voxelize_workspace(int n_seg, int n_pts, int64_t cells_max);
int pcfb_voxelize(const float *xyz, const int32_t *seg_off, int n_seg, int n_pts, double voxel, int64_t cells_max,
                  int32_t *out_idx, int32_t *out_seg_off, int32_t *status, void *workspace, size_t workspace_bytes,
                  void *stream);
size_t pcfb_pyramid_level_workspace(int n_seg, int n_pts_max, int64_t cells_max);
int pcfb_pyramid_level(const float *xyz, const float *feats, int F, const int32_t *seg_off, int n_seg, int n_pts_max,
                       float dl, int64_t cells_max, float *out_xyz, float *out_feats, int32_t *out_seg_off,
                       int32_t *status, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Layer glue the reference leaves to chains of torch elementwise ops.
 *
 * Guidance input of the PointConvFormer layer (layers.py:372-382):
 *   q = cat(index_points(guidance_x, nei), feat_pe);  key = q[:, :, :1] if M == N else max_k q;  out = q - key
 * gx [n_in, G], pe [n_out, K, P], nei [n_out, K] -> out [n_out, K, G+P]; arg [n_out, G+P] (use_max only: the k
 * attaining the maximum, first on ties).  G, P multiples of 4.  -1 / out-of-range neighbours gather zeros.
 * Backward: d_gq [n_out, K, G] is the per-edge gradient of the gathered half (sum it per input point with
 * pcfb_gather_backward: the index_put_(accumulate) of the reference's autograd, without atomics), d_pe [n_out, K, P].
 * ------------------------------------------------------------------------------------------- */
int pcfb_guidance_input(const float *gx, const float *pe, const int64_t *nei, int n_in, int n_out, int K, int G, int P,
                        int use_max, float *out, uint8_t *arg, void *stream);
int pcfb_guidance_input_backward(const float *ds, const uint8_t *arg, int n_out, int K, int G, int P, int use_max,
                                 float *d_gq, float *d_pe, void *stream);

/* clip_grad_norm_(max_norm) + AdamW step on ONE flat fp32 buffer (train_ScanNet_DDP_WarmUP.py:421-424; torch.optim.AdamW
 * semantics: step += 1; p *= 1 - lr*wd; m, v moments; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)).  lr and step are DEVICE
 * scalars (float) so a captured step can be replayed under a learning-rate schedule; max_norm <= 0 disables clipping;
 * norm_out (optional) receives the gradient norm before clipping.  Two launches, fixed-order reductions. */
size_t pcfb_adamw_workspace(void);
int pcfb_adamw_clip_step(float *p, const float *g, float *m, float *v, int64_t n, const float *lr, float *step,
                         float beta1, float beta2, float eps, float weight_decay, float max_norm, float *norm_out,
                         void *workspace, size_t workspace_bytes, void *stream);

/* Cross entropy over logits [N, C] (C <= 64) with torch.nn.CrossEntropyLoss semantics (train_ScanNet_DDP_WarmUP.py:417:
 * class weights, ignore_index, label_smoothing, mean reduction = sum / sum of the target-class weights of the valid rows).
 * forward -> loss[1], den[1]; backward recomputes the softmax from the logits: dlogits = grad_scale[0] * dL/dlogits. */
size_t pcfb_ce_workspace(int64_t N);
int pcfb_ce_forward(const float *logits, const int64_t *labels, const float *weight, int64_t N, int C, int64_t ignore_index,
                    float smoothing, float *loss, float *den, void *workspace, size_t workspace_bytes, void *stream);
int pcfb_ce_backward(const float *logits, const int64_t *labels, const float *weight, int64_t N, int C, int64_t ignore_index,
                     float smoothing, const float *den, const float *grad_scale, float *dlogits, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Hardware self-test of the tcgen05 (UMMA) operand / accumulator conventions used by the fused
 * forward: D = A[M x K] * B[N x K]^T on one CTA (M in {64,128}, N % 8 == 0, K % 8 == 0, K <= 64).
 * h_params is a HOST array of 12 uint32: fill (lbo_a, sbo_a, lbo_b, sbo_b) = where element (r,k) is
 * written, byte offset (k/4)*lbo + (r/8)*sbo + (r%8)*16 + (k%4)*4 for a K-major operand and
 * (r%4)*4 + (r/4)*sbo + (k%8)*16 + (k/8)*lbo for an MN-major one; desc (lbo_a, sbo_a, lbo_b, sbo_b) =
 * what the shared-memory descriptors claim; (kstep_a, kstep_b) = start-address advance per K=8 step;
 * idesc; mn_flags (bit 0: A is MN-major, bit 1: B is MN-major).  raw receives the accumulator as stored in TMEM: [128 lanes][N columns].  status (device
 * int) is set to 1 if the MMA never completed.  No reference counterpart (test infrastructure of
 * the product kernel, exercised by tests/test_umma_selftest.py).
 * ------------------------------------------------------------------------------------------- */
int pcfb_selftest_umma(const float *A, const float *B, float *raw, int M, int N, int K,
                       const uint32_t *h_params, uint64_t desc_or, int split, int *status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PCF_B200_H */
