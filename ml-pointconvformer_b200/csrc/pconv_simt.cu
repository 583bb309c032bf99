// Fused PointConv / PointConvFormer contraction (+ Linear), exact-fp32 SIMT variant (sm_100a).
//
// This is the "variant 1" kernel family of pcfb_pconv_forward / pcfb_pconv_backward: every product
// and sum is an fp32 FMA on CUDA cores, so it is the bit-stable bisecting reference for the tcgen05
// variant (pconv_umma.cu) and the fallback for shapes that one does not cover.
//
// Math (reference: torch path /root/reference/layers.py:386-390 [PCF], 713-719 [StridePE], 890-898
// [PointConv], 1086-1092 [Transpose]; fused CUDA path pconv_ops.cu:128-225, 390-536, 969-1269):
//   G[m,k,c] = c < C_in ? x[nei[m,k],c] * g[m,k,c % H] : add[m,k,c-C_in]
//   P[m,c*C_mid+j] = sum_k G[m,k,c] w[m,k,j]            Y[m,o] = sum_kk P[m,kk] W[o,kk] + b[o]
// Backward (autograd of the above; NOT the reference CUDA backward, whose channel layout is
// inconsistent with its forward -- SURVEY.md trap T1):
//   dP = dY W ; dw[m,k,j] = sum_c dP[m,c,j] G[m,k,c] ; dG[m,k,c] = sum_j dP[m,c,j] w[m,k,j]
//   dadd = dG[c>=C_in] ; dg[m,k,h] = sum_{c%H==h} x dG ; per-edge dE[m,k,c] = dG g  (c<C_in)
//   dx[p,:] = sum over the inverse-map segment of p of dE  (pcfb_gather_backward: no atomics)
//   dW = dY^T P (+ db = column sums of dY) by a split-M GEMM with a fixed-order reduction.
//
// Tiling: one CTA = 32 output points x 256 threads; thread t -> point p = t % 32 (the lane), group
// g = t / 32 (the warp).  All shared-memory rows are padded to an odd number of floats so that the
// 32 lanes (32 different points) of a warp hit 32 different banks, while everything indexed by the
// warp-uniform group (W rows, channel offsets) is a broadcast.  Channels are processed in chunks of
// CC (32/16/8, chosen so the tile fits in 220 KB of shared memory).
#include "common.cuh"

namespace pcfb {

constexpr int TP = 32;          // points per CTA
constexpr int NT = 256;         // threads per CTA
constexpr int NG = NT / TP;     // groups (= warps) per CTA
constexpr int KS = 64;          // kk sub-chunk of the Linear staged in smem
constexpr int MAXO = 32;        // max outputs per thread -> C_out <= 256
constexpr size_t SMEM_BUDGET = 220 * 1024;

struct PconvArgs {
    pcfb_pconv_shape s;
    const float *feats, *weights, *additional, *guidance, *lin_w, *lin_b;
    const int64_t *nei;
    // forward outputs
    float *out_y, *out_p;
    // backward
    const float *grad_y, *grad_p;
    float *grad_weights, *grad_additional, *grad_guidance, *grad_edge;   // grad_edge: [n_out,K,C_in] scratch
    int CC;                      // channel chunk
};

__host__ __device__ inline int odd(int x) { return x | 1; }

struct SmemPlan {
    int w_stride, g_stride, p_stride, W_stride, dy_stride, gd_stride, dgd_stride;
    size_t off_w, off_g, off_p, off_W, off_dy, off_gd, off_dgd, off_nei, total;
};

__host__ __device__ inline SmemPlan plan_smem(const pcfb_pconv_shape &s, int CC, bool backward) {
    SmemPlan pl;
    const int cmid = s.C_mid;
    pl.w_stride = odd(s.K * cmid);
    pl.g_stride = odd(s.K * CC);
    pl.p_stride = odd(CC * cmid);
    pl.W_stride = KS + 1;
    pl.dy_stride = odd(s.C_out > 0 ? s.C_out : 1);
    pl.gd_stride = odd(s.K * (s.H > 0 ? s.H : 1));
    size_t o = 0;
    pl.off_w = o;  o += (size_t)TP * pl.w_stride;
    pl.off_g = o;  o += (size_t)TP * pl.g_stride;
    // p_s doubles as the staging buffer for Y (forward)
    size_t p_sz = (size_t)TP * pl.p_stride;
    if (!backward && (size_t)TP * pl.dy_stride > p_sz) p_sz = (size_t)TP * pl.dy_stride;
    pl.off_p = o;  o += p_sz;
    pl.off_W = o;  o += (s.C_out > 0) ? (size_t)s.C_out * pl.W_stride : 0;
    pl.off_dy = o; o += (backward && s.C_out > 0) ? (size_t)TP * pl.dy_stride : 0;
    pl.off_gd = o; o += (s.H > 0) ? (size_t)TP * pl.gd_stride : 0;
    pl.dgd_stride = odd(s.K * NG);
    pl.off_dgd = o; o += (backward && s.H > 0) ? (size_t)TP * pl.dgd_stride : 0;
    pl.off_nei = o; o += (size_t)TP * s.K;      // int32
    pl.total = o * sizeof(float);
    return pl;
}

static int choose_cc(const pcfb_pconv_shape &s, bool backward) {
    for (int cc = 32; cc >= 1; cc >>= 1)
        if (plan_smem(s, cc, backward).total <= SMEM_BUDGET) return cc;
    return 0;
}

// ---- staging helpers -----------------------------------------------------------------------
__device__ __forceinline__ void stage_nei(const PconvArgs &a, int m0, int *nei_s) {
    const int K = a.s.K;
    for (int i = threadIdx.x; i < TP * K; i += NT) {
        const int p = i / K;
        const int m = m0 + p;
        int v = -1;
        if (m < a.s.n_out) {
            const int64_t q = a.nei[(size_t)m * K + (i - p * K)];
            if (q >= 0 && q < a.s.n_in) v = (int)q;
        }
        nei_s[i] = v;
    }
}

__device__ __forceinline__ void stage_rows(const float *__restrict__ src, int row_len, int m0, int n_out,
                                           float *dst, int dst_stride) {
    // src rows [m0, m0+TP) are contiguous in global memory: linear coalesced copy into padded rows
    const int total = TP * row_len;
    const size_t base = (size_t)m0 * row_len;
    for (int i = threadIdx.x; i < total; i += NT) {
        const int p = i / row_len;
        dst[p * dst_stride + (i - p * row_len)] = (m0 + p < n_out) ? src[base + i] : 0.f;
    }
}

// G chunk: raw features (optionally already multiplied by guidance) and additional channels
template <bool APPLY_GUIDANCE>
__device__ __forceinline__ void stage_g(const PconvArgs &a, int m0, int c0, int CCv, const int *nei_s,
                                        const float *gd_s, int gd_stride, float *g_s, int g_stride) {
    const int K = a.s.K, CC = a.CC, C_in = a.s.C_in, C_add = a.s.C_add, H = a.s.H;
    const int total = TP * K * CC;
    for (int i = threadIdx.x; i < total; i += NT) {
        const int cl = i % CC;
        const int pk = i / CC;
        const int p = pk / K, k = pk - p * K;
        const int c = c0 + cl;
        float v = 0.f;
        const int m = m0 + p;
        if (cl < CCv && m < a.s.n_out) {
            if (c < C_in) {
                const int q = nei_s[pk];
                if (q >= 0) {
                    v = __ldg(a.feats + (size_t)q * C_in + c);
                    if (APPLY_GUIDANCE && H > 0) v *= gd_s[p * gd_stride + k * H + (c % H)];
                }
            } else {
                v = __ldg(a.additional + ((size_t)m * K + k) * C_add + (c - C_in));
            }
        }
        g_s[p * g_stride + k * CC + cl] = v;
    }
}

// ---- forward ---------------------------------------------------------------------------------
template <int CMP>
__global__ void __launch_bounds__(NT, 1) pconv_fwd_simt_kernel(PconvArgs a)
{
    pdl_wait();
    extern __shared__ float smem[];
    const pcfb_pconv_shape &s = a.s;
    const SmemPlan pl = plan_smem(s, a.CC, false);
    float *w_s = smem + pl.off_w, *g_s = smem + pl.off_g, *p_s = smem + pl.off_p, *W_s = smem + pl.off_W;
    float *gd_s = smem + pl.off_gd;
    int *nei_s = reinterpret_cast<int *>(smem + pl.off_nei);
    const int K = s.K, cmid = s.C_mid, CC = a.CC, C_cat = s.C_in + s.C_add, KK = C_cat * cmid, C_out = s.C_out;
    const int p = threadIdx.x % TP, g = threadIdx.x / TP;
    const bool has_lin = a.lin_w != nullptr;

    for (int m0 = blockIdx.x * TP; m0 < s.n_out; m0 += gridDim.x * TP) {
        __syncthreads();
        stage_nei(a, m0, nei_s);
        stage_rows(a.weights, K * cmid, m0, s.n_out, w_s, pl.w_stride);
        if (s.H > 0) stage_rows(a.guidance, K * s.H, m0, s.n_out, gd_s, pl.gd_stride);
        float yacc[MAXO];
#pragma unroll
        for (int i = 0; i < MAXO; ++i) yacc[i] = 0.f;

        for (int c0 = 0; c0 < C_cat; c0 += CC) {
            const int CCv = min(CC, C_cat - c0);
            __syncthreads();                       // previous chunk fully consumed (g_s, p_s, W_s)
            stage_g<true>(a, m0, c0, CCv, nei_s, gd_s, pl.gd_stride, g_s, pl.g_stride);
            __syncthreads();
            // contraction 1: thread (p, g) owns channels cl = g + ci*NG of this chunk
            {
                float acc[4][CMP];
#pragma unroll
                for (int ci = 0; ci < 4; ++ci)
#pragma unroll
                    for (int j = 0; j < CMP; ++j) acc[ci][j] = 0.f;
                const float *gp = g_s + p * pl.g_stride;
                const float *wp = w_s + p * pl.w_stride;
                for (int k = 0; k < K; ++k) {
                    float wv[CMP];
#pragma unroll
                    for (int j = 0; j < CMP; ++j) wv[j] = (j < cmid) ? wp[k * cmid + j] : 0.f;
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        const int cl = g + ci * NG;
                        if (cl < CC) {
                            const float gv = gp[k * CC + cl];
#pragma unroll
                            for (int j = 0; j < CMP; ++j) acc[ci][j] = fmaf(gv, wv[j], acc[ci][j]);
                        }
                    }
                }
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const int cl = g + ci * NG;
                    if (cl < CC) {
#pragma unroll
                        for (int j = 0; j < CMP; ++j)
                            if (j < cmid) p_s[p * pl.p_stride + cl * cmid + j] = acc[ci][j];
                    }
                }
            }
            __syncthreads();
            const int seg = CCv * cmid;             // valid kk in this chunk, global kk = c0*cmid + kl
            if (a.out_p) {
                for (int i = threadIdx.x; i < TP * seg; i += NT) {
                    const int pp = i / seg, kl = i - pp * seg;
                    if (m0 + pp < s.n_out) a.out_p[(size_t)(m0 + pp) * KK + c0 * cmid + kl] = p_s[pp * pl.p_stride + kl];
                }
            }
            if (has_lin) {
                for (int ks0 = 0; ks0 < seg; ks0 += KS) {
                    const int ksv = min(KS, seg - ks0);
                    __syncthreads();
                    for (int i = threadIdx.x; i < C_out * KS; i += NT) {
                        const int o = i / KS, kl = i - o * KS;
                        W_s[o * pl.W_stride + kl] = (kl < ksv) ? __ldg(a.lin_w + (size_t)o * KK + c0 * cmid + ks0 + kl) : 0.f;
                    }
                    __syncthreads();
                    const float *pp = p_s + p * pl.p_stride + ks0;
                    for (int kl = 0; kl < ksv; ++kl) {
                        const float pv = pp[kl];
#pragma unroll
                        for (int oi = 0; oi < MAXO; ++oi) {
                            const int o = g + oi * NG;
                            if (o < C_out) yacc[oi] = fmaf(pv, W_s[o * pl.W_stride + kl], yacc[oi]);
                        }
                    }
                }
            }
        }
        if (has_lin) {
            __syncthreads();
            float *y_s = p_s;                        // [TP][dy_stride]
#pragma unroll
            for (int oi = 0; oi < MAXO; ++oi) {
                const int o = g + oi * NG;
                if (o < C_out) y_s[p * pl.dy_stride + o] = yacc[oi] + (a.lin_b ? __ldg(a.lin_b + o) : 0.f);
            }
            __syncthreads();
            for (int i = threadIdx.x; i < TP * C_out; i += NT) {
                const int pp = i / C_out, o = i - pp * C_out;
                if (m0 + pp < s.n_out) a.out_y[(size_t)(m0 + pp) * C_out + o] = y_s[pp * pl.dy_stride + o];
            }
        }
    }
}

// ---- backward main kernel --------------------------------------------------------------------
// KPT = ceil(K / NG): neighbours k = g + ki*NG owned by thread (p, g) for the dw accumulators.
template <int CMP, int KPT>
__global__ void __launch_bounds__(NT, 1) pconv_bwd_simt_kernel(PconvArgs a)
{
    pdl_wait();
    extern __shared__ float smem[];
    const pcfb_pconv_shape &s = a.s;
    const SmemPlan pl = plan_smem(s, a.CC, true);
    float *w_s = smem + pl.off_w, *g_s = smem + pl.off_g, *p_s = smem + pl.off_p, *W_s = smem + pl.off_W;
    float *dy_s = smem + pl.off_dy, *gd_s = smem + pl.off_gd, *dgd_s = smem + pl.off_dgd;
    int *nei_s = reinterpret_cast<int *>(smem + pl.off_nei);
    const int K = s.K, cmid = s.C_mid, CC = a.CC, C_in = s.C_in, C_add = s.C_add, C_cat = C_in + C_add;
    const int KK = C_cat * cmid, C_out = s.C_out, H = s.H;
    const int p = threadIdx.x % TP, g = threadIdx.x / TP;
    const bool has_lin = a.lin_w != nullptr;

    for (int m0 = blockIdx.x * TP; m0 < s.n_out; m0 += gridDim.x * TP) {
        __syncthreads();
        stage_nei(a, m0, nei_s);
        stage_rows(a.weights, K * cmid, m0, s.n_out, w_s, pl.w_stride);
        if (H > 0) stage_rows(a.guidance, K * H, m0, s.n_out, gd_s, pl.gd_stride);
        if (has_lin) stage_rows(a.grad_y, C_out, m0, s.n_out, dy_s, pl.dy_stride);

        if (H > 0)
            for (int i = threadIdx.x; i < TP * pl.dgd_stride; i += NT) dgd_s[i] = 0.f;
        float dw[KPT][CMP];                      // grad wrt weights[m, k = g+ki*NG, j]
#pragma unroll
        for (int ki = 0; ki < KPT; ++ki)
#pragma unroll
            for (int j = 0; j < CMP; ++j) dw[ki][j] = 0.f;

        for (int c0 = 0; c0 < C_cat; c0 += CC) {
            const int CCv = min(CC, C_cat - c0);
            const int seg = CCv * cmid;
            __syncthreads();
            stage_g<false>(a, m0, c0, CCv, nei_s, gd_s, pl.gd_stride, g_s, pl.g_stride);   // RAW x / additional
            // dP chunk -> p_s[p][kl]
            if (has_lin) {
                for (int ks0 = 0; ks0 < seg; ks0 += KS) {
                    const int ksv = min(KS, seg - ks0);
                    __syncthreads();
                    for (int i = threadIdx.x; i < C_out * KS; i += NT) {
                        const int o = i / KS, kl = i - o * KS;
                        W_s[o * pl.W_stride + kl] = (kl < ksv) ? __ldg(a.lin_w + (size_t)o * KK + c0 * cmid + ks0 + kl) : 0.f;
                    }
                    __syncthreads();
                    float acc[KS / NG];
#pragma unroll
                    for (int i = 0; i < KS / NG; ++i) acc[i] = 0.f;
                    const float *dyp = dy_s + p * pl.dy_stride;
                    for (int o = 0; o < C_out; ++o) {
                        const float dv = dyp[o];
                        const float *wr = W_s + o * pl.W_stride + g;
#pragma unroll
                        for (int i = 0; i < KS / NG; ++i) acc[i] = fmaf(dv, wr[i * NG], acc[i]);
                    }
#pragma unroll
                    for (int i = 0; i < KS / NG; ++i) {
                        const int kl = g + i * NG;
                        if (kl < ksv) p_s[p * pl.p_stride + ks0 + kl] = acc[i];
                    }
                }
            } else {
                for (int i = threadIdx.x; i < TP * seg; i += NT) {
                    const int pp = i / seg, kl = i - pp * seg;
                    p_s[pp * pl.p_stride + kl] = (m0 + pp < s.n_out) ? a.grad_p[(size_t)(m0 + pp) * KK + c0 * cmid + kl] : 0.f;
                }
            }
            __syncthreads();
            const float *gp = g_s + p * pl.g_stride;
            const float *dpp = p_s + p * pl.p_stride;
            const float *wp = w_s + p * pl.w_stride;
            const float *gdp = gd_s + p * pl.gd_stride;
            // dw[k,j] += sum_c dP[c,j] * G[k,c]   (G = x*guidance for c < C_in)
            for (int cl = 0; cl < CCv; ++cl) {
                float dpv[CMP];
#pragma unroll
                for (int j = 0; j < CMP; ++j) dpv[j] = (j < cmid) ? dpp[cl * cmid + j] : 0.f;
                const int c = c0 + cl;
#pragma unroll
                for (int ki = 0; ki < KPT; ++ki) {
                    const int k = g + ki * NG;
                    if (k < K) {
                        float gv = gp[k * CC + cl];
                        if (H > 0 && c < C_in) gv *= gdp[k * H + (c % H)];
#pragma unroll
                        for (int j = 0; j < CMP; ++j) dw[ki][j] = fmaf(dpv[j], gv, dw[ki][j]);
                    }
                }
            }
            __syncthreads();                       // every thread is done reading g_s for dw
            // dG[k,c] = sum_j dP[c,j] w[k,j] for the channels cl = g + ci*NG owned by this thread;
            // written back IN PLACE over g_s (the same thread reads then writes each element).  For
            // guided channels the slot first yields the raw feature x: x*dG goes to this thread's
            // private column of the guidance-gradient accumulator, and the edge gradient is dG*g.
            for (int ci = 0; ci < 4; ++ci) {
                const int cl = g + ci * NG;
                if (cl >= CCv) break;
                const int c = c0 + cl;
                const bool guided = (H > 0 && c < C_in);
                float dpv[CMP];
#pragma unroll
                for (int j = 0; j < CMP; ++j) dpv[j] = (j < cmid) ? dpp[cl * cmid + j] : 0.f;
                for (int k = 0; k < K; ++k) {
                    float dg = 0.f;
#pragma unroll
                    for (int j = 0; j < CMP; ++j) if (j < cmid) dg = fmaf(dpv[j], wp[k * cmid + j], dg);
                    float *slot = g_s + p * pl.g_stride + k * CC + cl;
                    if (guided) {
                        dgd_s[p * pl.dgd_stride + k * NG + g] += (*slot) * dg;
                        dg *= gdp[k * H + (c % H)];
                    }
                    *slot = dg;
                }
            }
            __syncthreads();
            // coalesced write-out of the per-edge gradient rows: c < C_in -> grad_edge, else grad_additional
            for (int i = threadIdx.x; i < TP * K * CCv; i += NT) {
                const int cl = i % CCv;
                const int pk = i / CCv;
                const int pp = pk / K, k = pk - pp * K;
                const int m = m0 + pp;
                if (m >= s.n_out) continue;
                const int c = c0 + cl;
                const float v = g_s[pp * pl.g_stride + k * CC + cl];
                if (c < C_in) { if (a.grad_edge) a.grad_edge[((size_t)m * K + k) * C_in + c] = v; }
                else if (a.grad_additional) a.grad_additional[((size_t)m * K + k) * C_add + (c - C_in)] = v;
            }
        }
        // write dw: stage through w_s (no longer needed) for a coalesced store
        __syncthreads();
        if (a.grad_weights) {
#pragma unroll
            for (int ki = 0; ki < KPT; ++ki) {
                const int k = g + ki * NG;
                if (k < K) {
#pragma unroll
                    for (int j = 0; j < CMP; ++j) if (j < cmid) w_s[p * pl.w_stride + k * cmid + j] = dw[ki][j];
                }
            }
            __syncthreads();
            const int row = K * cmid;
            for (int i = threadIdx.x; i < TP * row; i += NT) {
                const int pp = i / row;
                if (m0 + pp < s.n_out) a.grad_weights[(size_t)m0 * row + i] = w_s[pp * pl.w_stride + (i - pp * row)];
            }
        }
        // guidance gradient: reduce the NG private columns that share a head (column g <-> head g % H)
        if (H > 0 && a.grad_guidance) {
            const int row = K * H;
            for (int i = threadIdx.x; i < TP * row; i += NT) {
                const int pp = i / row, r = i - pp * row;
                const int k = r / H, h = r - k * H;
                float sum = 0.f;
                for (int gg = h; gg < NG; gg += H) sum += dgd_s[pp * pl.dgd_stride + k * NG + gg];
                if (m0 + pp < s.n_out) a.grad_guidance[(size_t)m0 * row + i] = sum;
            }
        }
    }
}


// ---- dW = dY^T P (+ db) : split-M SIMT GEMM with fixed-order reduction -------------------------
// C[o][kk] = sum_m dY[m][o] * P[m][kk] for kk < KK, and the virtual column kk == KK carries
// sum_m dY[m][o] (= grad of the bias).  grid = (ceil((KK+1)/64), ceil(C_out/32), S).
constexpr int GW_TO = 32, GW_TK = 64, GW_MB = 16;
__global__ void __launch_bounds__(256)
gradw_partial_kernel(const float *__restrict__ dY, const float *__restrict__ P, int M, int C_out, int KK,
                     int slice, float *__restrict__ partial)
{
    pdl_wait();
    __shared__ float A_s[GW_MB][GW_TO + 1];
    __shared__ __align__(16) float B_s[GW_MB][GW_TK];
    const int kk0 = blockIdx.x * GW_TK, o0 = blockIdx.y * GW_TO;
    const int mbeg = blockIdx.z * slice, mend = min(M, mbeg + slice);
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int m0 = mbeg; m0 < mend; m0 += GW_MB) {
        __syncthreads();
        for (int i = threadIdx.x; i < GW_MB * GW_TO; i += 256) {
            const int mm = i / GW_TO, o = i % GW_TO;
            A_s[mm][o] = (m0 + mm < mend && o0 + o < C_out) ? dY[(size_t)(m0 + mm) * C_out + o0 + o] : 0.f;
        }
        for (int i = threadIdx.x; i < GW_MB * GW_TK; i += 256) {
            const int mm = i / GW_TK, kl = i % GW_TK;
            const int kk = kk0 + kl;
            float v = 0.f;
            if (m0 + mm < mend) v = (kk < KK) ? P[(size_t)(m0 + mm) * KK + kk] : (kk == KK ? 1.f : 0.f);
            B_s[mm][kl] = v;
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < GW_MB; ++mm) {
            const float a0 = A_s[mm][ty * 2], a1 = A_s[mm][ty * 2 + 1];
            const float4 b = *reinterpret_cast<const float4 *>(&B_s[mm][tx * 4]);
            acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
            acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
            acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
            acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
        }
    }
    const int ld = KK + 1;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int o = o0 + ty * 2 + r;
        if (o >= C_out) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int kk = kk0 + tx * 4 + q;
            if (kk <= KK) partial[((size_t)blockIdx.z * C_out + o) * ld + kk] = acc[r][q];
        }
    }
}

__global__ void gradw_reduce_kernel(const float *__restrict__ partial, int S, int C_out, int KK,
                                    float *__restrict__ grad_w, float *__restrict__ grad_b)
{
    pdl_wait();
    const int ld = KK + 1;
    const int total = C_out * ld;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float sum = 0.f;
        for (int z = 0; z < S; ++z) sum += partial[(size_t)z * total + i];     // fixed order: deterministic
        const int o = i / ld, kk = i - o * ld;
        if (kk < KK) { if (grad_w) grad_w[(size_t)o * KK + kk] = sum; }
        else if (grad_b) grad_b[o] = sum;
    }
}

// ---- host side ---------------------------------------------------------------------------------
static int validate_shape(const pcfb_pconv_shape &s, const char *who, bool has_lin) {
    PCFB_REQUIRE(s.n_in >= 1 && s.n_out >= 0 && s.K >= 1 && s.C_in >= 1 && s.C_add >= 0 && s.C_mid >= 1,
                 "%s: bad shape", who);
    PCFB_REQUIRE(s.C_mid <= 16, "%s: C_mid=%d > 16 unsupported", who, s.C_mid);
    PCFB_REQUIRE(s.K <= 64, "%s: K=%d > 64 unsupported by the fused contraction", who, s.K);
    PCFB_REQUIRE(!has_lin || (s.C_out >= 1 && s.C_out <= NG * MAXO), "%s: C_out=%d outside [1,%d]", who, s.C_out, NG * MAXO);
    PCFB_REQUIRE(s.H == 0 || ((s.H == 1 || s.H == 2 || s.H == 4 || s.H == 8) && s.C_in % s.H == 0),
                 "%s: guidance heads H=%d must be 1,2,4 or 8 and divide C_in=%d", who, s.H, s.C_in);
    return PCFB_OK;
}

template <typename Kern>
static int set_smem(Kern kern, size_t bytes) {
    PCFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUDGET));
    (void)bytes;
    return PCFB_OK;
}

static int launch_fwd_simt(const PconvArgs &a, cudaStream_t st) {
    const SmemPlan pl = plan_smem(a.s, a.CC, false);
    const int grid = max(1, min(ceil_div(a.s.n_out, TP), kNumSMs * 8));
    int rc;
#define LAUNCH_FWD(CMP)                                                                     \
    do {                                                                                    \
        if ((rc = set_smem(pconv_fwd_simt_kernel<CMP>, pl.total))) return rc;               \
        launch_k(pconv_fwd_simt_kernel<CMP>, grid, NT, pl.total, st, a);                          \
    } while (0)
    if (a.s.C_mid == 1) LAUNCH_FWD(1);
    else if (a.s.C_mid <= 4) LAUNCH_FWD(4);
    else if (a.s.C_mid <= 8) LAUNCH_FWD(8);
    else LAUNCH_FWD(16);
#undef LAUNCH_FWD
    return check_launch("pconv_fwd_simt_kernel");
}

static int launch_bwd_simt(const PconvArgs &a, cudaStream_t st) {
    const SmemPlan pl = plan_smem(a.s, a.CC, true);
    const int grid = max(1, min(ceil_div(a.s.n_out, TP), kNumSMs * 8));
    const int kpt = ceil_div(a.s.K, NG);
    int rc;
#define LAUNCH_BWD(CMP, KPT)                                                                \
    do {                                                                                    \
        if ((rc = set_smem(pconv_bwd_simt_kernel<CMP, KPT>, pl.total))) return rc;          \
        launch_k(pconv_bwd_simt_kernel<CMP, KPT>, grid, NT, pl.total, st, a);                     \
    } while (0)
#define LAUNCH_BWD_K(CMP)                                                                   \
    do {                                                                                    \
        if (kpt <= 2) LAUNCH_BWD(CMP, 2); else if (kpt <= 4) LAUNCH_BWD(CMP, 4); else LAUNCH_BWD(CMP, 8); \
    } while (0)
    if (a.s.C_mid == 1) LAUNCH_BWD_K(1);
    else if (a.s.C_mid <= 4) LAUNCH_BWD_K(4);
    else if (a.s.C_mid <= 8) LAUNCH_BWD_K(8);
    else LAUNCH_BWD_K(16);
#undef LAUNCH_BWD_K
#undef LAUNCH_BWD
    return check_launch("pconv_bwd_simt_kernel");
}

int pconv_forward_simt(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                       const float *additional, const float *guidance, const float *lin_w, const float *lin_b,
                       float *out_y, float *out_p, cudaStream_t st)
{
    int rc;
    if ((rc = validate_shape(*s, "pcfb_pconv_forward", lin_w != nullptr))) return rc;
    if (s->n_out == 0) return PCFB_OK;
    PconvArgs a{};
    a.s = *s;
    if (!lin_w) a.s.C_out = 0;
    a.feats = feats; a.nei = nei; a.weights = weights; a.additional = additional; a.guidance = guidance;
    a.lin_w = lin_w; a.lin_b = lin_b; a.out_y = out_y; a.out_p = out_p;
    a.CC = choose_cc(a.s, false);
    PCFB_REQUIRE(a.CC > 0, "pcfb_pconv_forward: tile does not fit in shared memory (K=%d C_mid=%d C_out=%d)", s->K, s->C_mid, s->C_out);
    return launch_fwd_simt(a, st);
}

struct BwdWorkspace {
    float *grad_edge;   // [n_out*K*C_in]
    float *p_recompute; // [n_out*KK]
    float *partial;     // [S*C_out*(KK+1)]
    int S, slice;
    size_t bytes;
};

static BwdWorkspace carve_bwd(void *ws, const pcfb_pconv_shape &s, bool need_edge, bool need_p, bool need_w) {
    Carver c(ws);
    BwdWorkspace w{};
    const int KK = (s.C_in + s.C_add) * s.C_mid;
    if (need_edge) w.grad_edge = c.take<float>((size_t)s.n_out * s.K * s.C_in);
    if (need_p) w.p_recompute = c.take<float>((size_t)s.n_out * KK);
    if (need_w) {
        const int tiles = ceil_div(KK + 1, GW_TK) * ceil_div(s.C_out, GW_TO);
        int S = ceil_div(4 * kNumSMs, tiles);
        const int maxS = max(1, ceil_div(s.n_out, 4 * GW_MB));
        if (S > maxS) S = maxS;
        if (S < 1) S = 1;
        w.S = S;
        w.slice = ceil_div(ceil_div(max(s.n_out, 1), S), GW_MB) * GW_MB;
        w.partial = c.take<float>((size_t)S * s.C_out * (KK + 1));
    }
    w.bytes = align_up(c.off, 256);
    return w;
}

size_t pconv_backward_simt_workspace(const pcfb_pconv_shape *s) {
    // worst case: everything requested, P not supplied
    return carve_bwd(nullptr, *s, true, true, s->C_out > 0).bytes;
}

int pconv_backward_simt(const pcfb_pconv_shape *s, const float *grad_y, const float *grad_p, const float *feats,
                        const int64_t *nei, const int32_t *inv_n, const uint8_t *inv_k, const int32_t *inv_idx,
                        const float *weights, const float *additional, const float *guidance, const float *lin_w,
                        const float *pconv_out, float *grad_feats, float *grad_weights, float *grad_additional,
                        float *grad_guidance, float *grad_lin_w, float *grad_lin_b, void *workspace,
                        size_t workspace_bytes, cudaStream_t st)
{
    int rc;
    const bool has_lin = lin_w != nullptr;
    if ((rc = validate_shape(*s, "pcfb_pconv_backward", has_lin))) return rc;
    PCFB_REQUIRE(has_lin ? grad_y != nullptr : grad_p != nullptr, "pcfb_pconv_backward: missing incoming gradient");
    PCFB_REQUIRE(!grad_feats || (inv_n && inv_k && inv_idx), "pcfb_pconv_backward: grad_feats needs the inverse map");
    const bool need_w = has_lin && (grad_lin_w || grad_lin_b);
    const bool need_p = need_w && !pconv_out;
    BwdWorkspace w = carve_bwd(workspace, *s, grad_feats != nullptr, need_p, need_w);
    if (w.bytes > 0 && (!workspace || workspace_bytes < w.bytes)) {
        set_error("pcfb_pconv_backward: workspace %zu < %zu", workspace_bytes, w.bytes);
        return PCFB_ERR_WORKSPACE;
    }
    if (s->n_out > 0) {
        PconvArgs a{};
        a.s = *s;
        if (!has_lin) a.s.C_out = 0;
        a.feats = feats; a.nei = nei; a.weights = weights; a.additional = additional; a.guidance = guidance;
        a.lin_w = lin_w; a.grad_y = grad_y; a.grad_p = grad_p;
        a.grad_weights = grad_weights; a.grad_additional = grad_additional; a.grad_guidance = grad_guidance;
        a.grad_edge = w.grad_edge;
        a.CC = choose_cc(a.s, true);
        PCFB_REQUIRE(a.CC > 0, "pcfb_pconv_backward: tile does not fit in shared memory (K=%d C_mid=%d C_out=%d)", s->K, s->C_mid, s->C_out);
        if ((rc = launch_bwd_simt(a, st))) return rc;
    }
    if (grad_feats) {
        if ((rc = pcfb_gather_backward(w.grad_edge, inv_n, inv_k, inv_idx, s->n_in, s->n_out, s->K, s->C_in, grad_feats, st))) return rc;
    }
    if (need_w) {
        const int KK = (s->C_in + s->C_add) * s->C_mid;
        const float *P = pconv_out;
        if (need_p) {
            if ((rc = pconv_forward_simt(s, feats, nei, weights, additional, guidance, nullptr, nullptr, nullptr, w.p_recompute, st))) return rc;
            P = w.p_recompute;
        }
        dim3 grid(ceil_div(KK + 1, GW_TK), ceil_div(s->C_out, GW_TO), w.S);
        launch_k(gradw_partial_kernel, grid, 256, 0, st, grad_y, P, s->n_out, s->C_out, KK, w.slice, w.partial);
        if ((rc = check_launch("gradw_partial_kernel"))) return rc;
        const int total = s->C_out * (KK + 1);
        launch_k(gradw_reduce_kernel, min(ceil_div(total, 256), kNumSMs * 8), 256, 0, st, w.partial, w.S, s->C_out, KK, grad_lin_w, grad_lin_b);
        if ((rc = check_launch("gradw_reduce_kernel"))) return rc;
    }
    return PCFB_OK;
}

}  // namespace pcfb
