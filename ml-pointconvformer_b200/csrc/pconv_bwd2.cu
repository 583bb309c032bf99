// Contraction backward of the fused PointConv / PointConvFormer layer, pipelined CUDA-core kernel (sm_100a).
//
// Input: dP [M, C_cat*C_mid] (= dY W, produced on tcgen05 by pcfb_gemm_nt), the gathered features G, the weightnet
// output w and the guidance g.  Output (autograd of P[m,c*C_mid+j] = sum_k G[m,k,c] w[m,k,j]):
//   dw[m,k,j]  = sum_c dP[m,c,j] * G[m,k,c]                (G = x*g for c < C_in, = additional otherwise)
//   dG[m,k,c]  = sum_j dP[m,c,j] * w[m,k,j]  ->  per-edge gradient dE = dG*g (c < C_in, summed per input point by the
//                inverse-map CSR kernel: no atomics), dadd = dG (c >= C_in), dg[m,k,h] = sum_{c%H==h} x*dG.
// Replaces the per-point part of pconv_linear_fused_cuda_backward_kernel_opt / pcf_cuda_backward_kernel
// (/root/reference/cpp_wrappers/cpp_pcf_kernel/src/pconv_ops.cu:390-536, pcf_ops.cu:86-141; 1.6 G atomicAdds there).
//
// Tile = 64 points, thread = (point p, quarter kg of the 16 neighbours).  A thread owns dw and dg of its 4 neighbours
// for the whole tile and the matching 4 x C_mid weightnet values (all in registers); per channel it needs the
// C_mid-vector dP[p,c,:] (shared by the 4 threads of the point, read as LDS.128) and 4 gathered values:
// 2*4*C_mid FMAs per 4 + C_mid/4 shared-memory loads.  dP and G stream through a 3-stage cp.async ring.
#include "common.cuh"
#include "umma.cuh"

namespace pcfb {

constexpr int BT = 64;         // points per tile (level-0/1 sized launches)
constexpr int BT_SMALL = 16;   // points per tile on the coarse levels: the tile's work is a serial FMA stream per CTA, so a 184-point
                               // level that ran as 3 CTAs for 115 us runs as 12 (and a 1 k-point level as 65 instead of 17)
constexpr int BK = 16;         // neighbours
constexpr int BCC = 4;         // channels per chunk

struct Bwd2Args {
    pcfb_pconv_shape s;
    const float *dP, *feats, *weights, *additional, *guidance;
    const int64_t *nei;
    float *grad_weights, *grad_additional, *grad_guidance, *grad_edge;
    int n_chunks, vec_ok;
};

__host__ __device__ inline int b2_pad4odd(int x) {
    int y = (x + 3) & ~3;
    if (((y >> 2) & 1) == 0) y += 4;
    return y;
}

struct Bwd2Plan { int GS, DS, GDS; size_t off_g, off_d, off_gd, off_nei, total; };

__host__ __device__ inline Bwd2Plan b2_plan(const pcfb_pconv_shape &s, int BT = 64) {
    Bwd2Plan pl;
    pl.GS = b2_pad4odd(BK * BCC);
    pl.DS = b2_pad4odd(BCC * s.C_mid);
    pl.GDS = b2_pad4odd(BK * (s.H > 0 ? s.H : 1));
    size_t o = 0;
    pl.off_g = o;   o += 3 * (size_t)BT * pl.GS * 4;
    pl.off_d = o;   o += 3 * (size_t)BT * pl.DS * 4;
    pl.off_gd = o;  o += (s.H > 0) ? (size_t)BT * pl.GDS * 4 : 0;
    o = align_up(o, 16);
    pl.off_nei = o; o += 2 * (size_t)BT * BK * 8;
    pl.total = align_up(o, 16);
    return pl;
}

__device__ __forceinline__ void b2_cp16(void *dst, const void *src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(umma::smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void b2_cp16_ca(void *dst, const void *src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" :: "r"(umma::smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void b2_cp4(void *dst, const void *src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" :: "r"(umma::smem_u32(dst)), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void b2_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void b2_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

template <int CMID, bool GUIDE, int BT>
__global__ void __launch_bounds__(4 * BT, 1) pconv_bwd2_kernel(Bwd2Args a)
{
    pdl_wait();
    constexpr int BNT = 4 * BT;                                 // thread = (point, quarter of the 16 neighbours)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const pcfb_pconv_shape &s = a.s;
    const Bwd2Plan pl = b2_plan(s, BT);
    float *g_s = reinterpret_cast<float *>(smem_raw + pl.off_g);
    float *d_s = reinterpret_cast<float *>(smem_raw + pl.off_d);
    float *gd_s = reinterpret_cast<float *>(smem_raw + pl.off_gd);
    long long *nei_s = reinterpret_cast<long long *>(smem_raw + pl.off_nei);
    constexpr int K = BK, CC = BCC, KQ = 4;                     // KQ neighbours per thread
    const int C_in = s.C_in, C_add = s.C_add, C_cat = C_in + C_add, KK = C_cat * CMID, H = s.H;
    const int n_in = s.n_in, n_out = s.n_out, n_chunks = a.n_chunks;
    const int tid = threadIdx.x;
    const int p = tid & (BT - 1), kg = tid / BT;
    const int GS = pl.GS, DS = pl.DS;

    auto issue_nei = [&](int m0, long long *dst) {
        const size_t base = (size_t)m0 * K;
        for (int i = tid * 2; i < BT * K; i += BNT * 2) b2_cp16(dst + i, a.nei + base + i, m0 + i / K < n_out);
    };
    auto issue_chunk = [&](int m0, int chunk, const long long *nei_cur) {
        float *gdst = g_s + (size_t)(chunk % 3) * BT * GS;
        float *ddst = d_s + (size_t)(chunk % 3) * BT * DS;
        const int c0 = chunk * CC;
        if (a.vec_ok) {
            for (int pk = tid; pk < BT * K; pk += BNT) {
                const int pp = pk / K, k = pk - pp * K;
                const int m = m0 + pp;
                const float *src = a.feats;
                bool valid = false;
                if (c0 < C_cat && m < n_out) {
                    if (c0 < C_in) {
                        const long long q = nei_cur[pk];
                        if (q >= 0 && q < n_in) { src = a.feats + (size_t)q * C_in + c0; valid = true; }
                    } else {
                        src = a.additional + ((size_t)m * K + k) * C_add + (c0 - C_in); valid = true;
                    }
                }
                b2_cp16_ca(gdst + pp * GS + k * CC, src, valid);
            }
        } else {
            for (int i = tid; i < BT * K * CC; i += BNT) {
                const int cl = i % CC, pk = i / CC;
                const int pp = pk / K, k = pk - pp * K;
                const int c = c0 + cl, m = m0 + pp;
                const float *src = a.feats;
                bool valid = false;
                if (c < C_cat && m < n_out) {
                    if (c < C_in) {
                        const long long q = nei_cur[pk];
                        if (q >= 0 && q < n_in) { src = a.feats + (size_t)q * C_in + c; valid = true; }
                    } else {
                        src = a.additional + ((size_t)m * K + k) * C_add + (c - C_in); valid = true;
                    }
                }
                b2_cp4(gdst + pp * GS + k * CC + cl, src, valid);
            }
        }
        // dP chunk: CC*CMID contiguous floats per point
        constexpr int UPP = CC * CMID / 4;                       // 16-byte units per point
        for (int i = tid; i < BT * UPP; i += BNT) {
            const int pp = i / UPP, u = i - pp * UPP;
            const int m = m0 + pp, kk = c0 * CMID + u * 4;
            b2_cp16(ddst + pp * DS + u * 4, a.dP + (size_t)(m < n_out ? m : 0) * KK + (kk < KK ? kk : 0), m < n_out && kk < KK);
        }
    };

    const int first = blockIdx.x * BT;
    if (first < n_out) issue_nei(first, nei_s);
    b2_commit();
    int tile_it = 0;
    for (int m0 = first; m0 < n_out; m0 += gridDim.x * BT, ++tile_it) {
        const long long *nei_cur = nei_s + (size_t)(tile_it & 1) * BT * K;
        long long *nei_nxt = nei_s + (size_t)((tile_it + 1) & 1) * BT * K;
        b2_wait<0>();
        __syncthreads();
        if (GUIDE) {
            const int grow = K * H;
            const size_t gbase = (size_t)m0 * grow;
            for (int i = tid * 4; i < BT * grow; i += BNT * 4) {
                const int pp = i / grow;
                b2_cp16(gd_s + pp * pl.GDS + (i - pp * grow), a.guidance + gbase + i, m0 + pp < n_out);
            }
        }
        issue_chunk(m0, 0, nei_cur);
        b2_commit();
        if (n_chunks > 1) issue_chunk(m0, 1, nei_cur);
        b2_commit();
        if (n_chunks > 2) issue_chunk(m0, 2, nei_cur);
        {
            const int m_next = m0 + gridDim.x * BT;
            if (m_next < n_out) issue_nei(m_next, nei_nxt);
        }
        b2_commit();

        // registers for the whole tile: weightnet rows of my 4 neighbours, their dw accumulators, guidance + its gradient
        const int m = m0 + p;
        const bool live = m < n_out;
        float wreg[KQ][CMID], dw[KQ][CMID];
        {
            const float *wsrc = a.weights + ((size_t)(live ? m : 0) * K + kg * KQ) * CMID;
#pragma unroll
            for (int kk = 0; kk < KQ; ++kk)
#pragma unroll
                for (int j = 0; j < CMID; j += 4) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live) v = __ldg(reinterpret_cast<const float4 *>(wsrc + kk * CMID + j));
                    wreg[kk][j] = v.x; wreg[kk][j + 1] = v.y; wreg[kk][j + 2] = v.z; wreg[kk][j + 3] = v.w;
                    dw[kk][j] = dw[kk][j + 1] = dw[kk][j + 2] = dw[kk][j + 3] = 0.f;
                }
        }
        float dgd[GUIDE ? KQ : 1][GUIDE ? 8 : 1];
        if (GUIDE) {
#pragma unroll
            for (int kk = 0; kk < KQ; ++kk)
#pragma unroll
                for (int h = 0; h < 8; ++h) dgd[kk][h] = 0.f;
        }

        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            b2_wait<2>();
            __syncthreads();
            const float *gp = g_s + (size_t)(chunk % 3) * BT * GS + p * GS + kg * KQ * CC;
            const float *dp = d_s + (size_t)(chunk % 3) * BT * DS + p * DS;
            const int c0 = chunk * CC;
            float gx[KQ][CC];                                     // raw gathered values of my 4 neighbours x 4 channels
#pragma unroll
            for (int kk = 0; kk < KQ; ++kk) {
                const float4 v = *reinterpret_cast<const float4 *>(gp + kk * CC);
                gx[kk][0] = v.x; gx[kk][1] = v.y; gx[kk][2] = v.z; gx[kk][3] = v.w;
            }
            float dG[KQ][CC];
#pragma unroll
            for (int cl = 0; cl < CC; ++cl) {
                const int c = c0 + cl;
                float dpv[CMID];
#pragma unroll
                for (int j = 0; j < CMID; j += 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(dp + cl * CMID + j);
                    dpv[j] = v.x; dpv[j + 1] = v.y; dpv[j + 2] = v.z; dpv[j + 3] = v.w;
                }
                const int h = GUIDE ? (c % H) : 0;
#pragma unroll
                for (int kk = 0; kk < KQ; ++kk) {
                    float gval = gx[kk][cl];
                    float gfac = 1.f;
                    if (GUIDE && c < C_in) gfac = gd_s[p * pl.GDS + (kg * KQ + kk) * H + h];
                    const float geff = gval * gfac;
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < CMID; ++j) {
                        dw[kk][j] = fmaf(dpv[j], geff, dw[kk][j]);
                        acc = fmaf(dpv[j], wreg[kk][j], acc);
                    }
                    if (GUIDE && c < C_in) {
                        // head index is a runtime value: accumulate through a predicated unrolled select
#pragma unroll
                        for (int hh = 0; hh < 8; ++hh) if (hh == h) dgd[kk][hh] = fmaf(gval, acc, dgd[kk][hh]);
                        acc *= gfac;
                    }
                    dG[kk][cl] = acc;
                }
            }
            // per-edge gradient rows of my 4 neighbours, 4 channels each
            if (live && c0 < C_cat) {
#pragma unroll
                for (int kk = 0; kk < KQ; ++kk) {
                    const size_t edge = (size_t)m * K + kg * KQ + kk;
                    if (a.vec_ok) {
                        const float4 v = make_float4(dG[kk][0], dG[kk][1], dG[kk][2], dG[kk][3]);
                        if (c0 < C_in) { if (a.grad_edge) *reinterpret_cast<float4 *>(a.grad_edge + edge * C_in + c0) = v; }
                        else if (a.grad_additional) *reinterpret_cast<float4 *>(a.grad_additional + edge * C_add + (c0 - C_in)) = v;
                    } else {
#pragma unroll
                        for (int cl = 0; cl < CC; ++cl) {
                            const int c = c0 + cl;
                            if (c < C_in) { if (a.grad_edge) a.grad_edge[edge * C_in + c] = dG[kk][cl]; }
                            else if (c < C_cat && a.grad_additional) a.grad_additional[edge * C_add + (c - C_in)] = dG[kk][cl];
                        }
                    }
                }
            }
            __syncthreads();                                     // ring slot chunk%3 consumed by everyone
            if (chunk + 3 < n_chunks) issue_chunk(m0, chunk + 3, nei_cur);
            b2_commit();
        }
        if (live) {
            if (a.grad_weights) {
                float *dst = a.grad_weights + ((size_t)m * K + kg * KQ) * CMID;
#pragma unroll
                for (int kk = 0; kk < KQ; ++kk)
#pragma unroll
                    for (int j = 0; j < CMID; j += 4)
                        *reinterpret_cast<float4 *>(dst + kk * CMID + j) = make_float4(dw[kk][j], dw[kk][j + 1], dw[kk][j + 2], dw[kk][j + 3]);
            }
            if (GUIDE && a.grad_guidance) {
                float *dst = a.grad_guidance + ((size_t)m * K + kg * KQ) * H;
#pragma unroll
                for (int kk = 0; kk < KQ; ++kk)
#pragma unroll
                    for (int h = 0; h < 8; ++h) if (h < H) dst[kk * H + h] = dgd[kk][h];
            }
        }
    }
    b2_wait<0>();
}

bool pconv_bwd2_supported(const pcfb_pconv_shape *s) {
    if (s->K != BK) return false;
    if (!(s->C_mid == 4 || s->C_mid == 8 || s->C_mid == 16)) return false;
    if (s->H != 0 && !((s->H == 1 || s->H == 2 || s->H == 4 || s->H == 8) && s->C_in % s->H == 0)) return false;
    return b2_plan(*s).total <= 200 * 1024;
}

template <int CMID, bool GUIDE>
static int launch_bwd2(const Bwd2Args &a, const Bwd2Plan &, cudaStream_t st) {
    // fewer 64-point tiles than ~1.5 per SM: the launch is bounded by one CTA's serial work, use 16-point tiles
    if (ceil_div(a.s.n_out, BT) < kNumSMs * 3 / 2) {
        const Bwd2Plan pl = b2_plan(a.s, BT_SMALL);
        PCFB_CUDA(cudaFuncSetAttribute(pconv_bwd2_kernel<CMID, GUIDE, BT_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const int grid = max(1, min(ceil_div(a.s.n_out, BT_SMALL), kNumSMs * 6));
        launch_k(pconv_bwd2_kernel<CMID, GUIDE, BT_SMALL>, grid, 4 * BT_SMALL, pl.total, st, a);
        return check_launch("pconv_bwd2_kernel");
    }
    const Bwd2Plan pl = b2_plan(a.s, BT);
    PCFB_CUDA(cudaFuncSetAttribute(pconv_bwd2_kernel<CMID, GUIDE, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int per_sm = (pl.total + 1024 <= 113 * 1024) ? 2 : 1;   // register use decides the real residency
    const int grid = max(1, min(ceil_div(a.s.n_out, BT), kNumSMs * per_sm));
    launch_k(pconv_bwd2_kernel<CMID, GUIDE, BT>, grid, 4 * BT, pl.total, st, a);
    return check_launch("pconv_bwd2_kernel");
}

// dP given: writes grad_weights / grad_additional / grad_guidance / grad_edge (per-edge rows for the CSR sum)
int pconv_bwd2(const pcfb_pconv_shape *s, const float *dP, const float *feats, const int64_t *nei, const float *weights,
               const float *additional, const float *guidance, float *grad_weights, float *grad_additional,
               float *grad_guidance, float *grad_edge, cudaStream_t st)
{
    PCFB_REQUIRE(pconv_bwd2_supported(s), "pconv_bwd2: unsupported shape");
    PCFB_REQUIRE(((uintptr_t)dP % 16 == 0) && ((uintptr_t)weights % 16 == 0) && ((uintptr_t)nei % 16 == 0) &&
                 (!grad_weights || (uintptr_t)grad_weights % 16 == 0), "pconv_bwd2: unaligned buffers");
    if (s->n_out == 0) return PCFB_OK;
    Bwd2Args a{};
    a.s = *s;
    a.dP = dP; a.feats = feats; a.nei = nei; a.weights = weights; a.additional = additional; a.guidance = guidance;
    a.grad_weights = grad_weights; a.grad_additional = grad_additional; a.grad_guidance = grad_guidance; a.grad_edge = grad_edge;
    a.n_chunks = ceil_div(s->C_in + s->C_add, BCC);
    a.vec_ok = (s->C_in % 4 == 0) && (s->C_add % 4 == 0) && ((uintptr_t)feats % 16 == 0) &&
               (s->C_add == 0 || (uintptr_t)additional % 16 == 0) &&
               (!grad_edge || (uintptr_t)grad_edge % 16 == 0) && (!grad_additional || (uintptr_t)grad_additional % 16 == 0) &&
               (s->H == 0 || (uintptr_t)guidance % 16 == 0);
    PCFB_REQUIRE(s->H == 0 || ((s->K * s->H) % 4 == 0 && (uintptr_t)guidance % 16 == 0), "pconv_bwd2: guidance rows must be 16-byte aligned");
    const Bwd2Plan pl = b2_plan(*s);
    const bool guide = s->H > 0;
    switch (s->C_mid) {
        case 4:  return guide ? launch_bwd2<4, true>(a, pl, st) : launch_bwd2<4, false>(a, pl, st);
        case 8:  return guide ? launch_bwd2<8, true>(a, pl, st) : launch_bwd2<8, false>(a, pl, st);
        default: return guide ? launch_bwd2<16, true>(a, pl, st) : launch_bwd2<16, false>(a, pl, st);
    }
}

}  // namespace pcfb
