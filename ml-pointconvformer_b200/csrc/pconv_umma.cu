// Fused PointConv / PointConvFormer forward on the Blackwell tensor cores (sm_100a, tcgen05 + TMEM).
//
// Replaces pconv_linear_cutlass_forward (/root/reference/cpp_wrappers/cpp_pcf_kernel/src/pconv_ops.cu:
// 969-1269: gather kernel -> SIMT CUTLASS batched GEMM -> SIMT CUTLASS GEMM -> bias, >= 5 HBM round trips
// of [N,512] tensors) with ONE kernel that touches HBM once per operand.
//
// Per CTA: a tile of 64 output points.  The K2 = C_cat*C_mid reduction of the Linear is walked in
// chunks of CK = 4*KPT (64/32/16) columns:
//   SIMT  : G chunk staged from the (L2-resident) feature rows / additional features, contraction 1
//           P[p, kk] = sum_k G[p,k,c] w[p,k,j] in fp32 registers (thread = (point, quarter of the chunk)),
//   split : P -> (hi, lo) tf32 pair written as the A operand (K-major, no-swizzle 8x16B core matrices),
//           the W chunk likewise as the B operand,
//   UMMA  : D[64 x C_out] += A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T  (tcgen05.mma kind::tf32, M=64,
//           N=C_out, accumulators in TMEM) -- the 3xTF32 split keeps ~2^-21 relative accuracy, which the
//           1e-4 parity bar needs (single-pass tf32 measures 4e-4),
//   the A/B chunk buffers are double buffered; an mbarrier armed by tcgen05.commit frees a buffer, so
//   the tensor pipe works on chunk i while the CUDA cores produce chunk i+1.
// Epilogue: tcgen05.ld D -> + bias -> Y.  P can optionally be streamed out (the reference saves it for
// the backward, layer_utils.py:52-54).
#include "common.cuh"
#include "umma.cuh"

namespace pcfb {

constexpr int UT = 64;        // points per tile = UMMA M
constexpr int UNT = 256;      // threads
constexpr size_t U_SMEM_BUDGET = 225 * 1024;

struct UmmaArgs {
    pcfb_pconv_shape s;
    const float *feats, *weights, *additional, *guidance, *lin_w, *lin_b;
    const int64_t *nei;
    float *out_y, *out_p;
    int tmem_cols;
};

struct UPlan {
    int CK, CC;                    // kk per chunk, channels per chunk
    int w_stride, g_stride, gd_stride;
    uint32_t a_bytes, b_bytes;     // one (hi or lo) operand buffer
    size_t off_A, off_B, off_w, off_g, off_gd, off_nei, off_bar, total;   // byte offsets
};

__host__ __device__ inline UPlan u_plan(const pcfb_pconv_shape &s, int KPT) {
    UPlan pl;
    pl.CK = 4 * KPT;
    pl.CC = (pl.CK + s.C_mid - 1) / s.C_mid;
    pl.w_stride = s.K * s.C_mid + 4;                       // multiple of 4 floats (LDS.128), /4 odd-ish
    if (((pl.w_stride / 4) & 1) == 0) pl.w_stride += 4;
    pl.g_stride = (s.K * pl.CC) | 1;
    pl.gd_stride = (s.K * (s.H > 0 ? s.H : 1)) | 1;
    pl.a_bytes = UT * pl.CK * 4;
    pl.b_bytes = s.C_out * pl.CK * 4;
    size_t o = 0;
    pl.off_A = o;  o += 4 * (size_t)pl.a_bytes;            // [buf][hi/lo]
    pl.off_B = o;  o += 4 * (size_t)pl.b_bytes;
    o = align_up(o, 16);
    pl.off_w = o;  o += (size_t)UT * pl.w_stride * 4;
    pl.off_g = o;  o += (size_t)UT * pl.g_stride * 4;
    pl.off_gd = o; o += (s.H > 0) ? (size_t)UT * pl.gd_stride * 4 : 0;
    pl.off_nei = o; o += (size_t)UT * s.K * 4;
    o = align_up(o, 16);
    pl.off_bar = o; o += 64;
    pl.total = o;
    return pl;
}

static int u_choose_kpt(const pcfb_pconv_shape &s) {
    for (int kpt = 16; kpt >= 4; kpt >>= 1) {
        if ((4 * kpt) % s.C_mid != 0 && s.C_mid % (4 * kpt) != 0) continue;
        if (u_plan(s, kpt).total <= U_SMEM_BUDGET) return kpt;
    }
    return 0;
}

template <int CMID, int KPT>
__global__ void __launch_bounds__(UNT, 1) pconv_fwd_umma_kernel(UmmaArgs a)
{
    pdl_wait();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const pcfb_pconv_shape &s = a.s;
    const UPlan pl = u_plan(s, KPT);
    constexpr int CK = 4 * KPT;
    const int CC = pl.CC;
    unsigned char *A_base = smem_raw + pl.off_A;
    unsigned char *B_base = smem_raw + pl.off_B;
    float *w_s = reinterpret_cast<float *>(smem_raw + pl.off_w);
    float *g_s = reinterpret_cast<float *>(smem_raw + pl.off_g);
    float *gd_s = reinterpret_cast<float *>(smem_raw + pl.off_gd);
    int *nei_s = reinterpret_cast<int *>(smem_raw + pl.off_nei);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + pl.off_bar);     // [0],[1]: chunk buffers
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + pl.off_bar + 32);

    const int K = s.K, C_in = s.C_in, C_add = s.C_add, C_cat = C_in + C_add, KK = C_cat * CMID, C_out = s.C_out, H = s.H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = tid & (UT - 1), g = tid >> 6;             // point within tile, quarter of the chunk

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "r"(umma::smem_u32(tmem_slot)), "r"(a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        umma::mbar_init(&bars[0], 1);
        umma::mbar_init(&bars[1], 1);
        umma::fence_mbar_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = *tmem_slot;

    const uint32_t idesc = umma::make_idesc_tf32(UT, C_out);
    const uint32_t lbo_a = UT * 16, lbo_b = (uint32_t)C_out * 16, sbo = 128;
    uint32_t uses0 = 0, uses1 = 0;                            // completed-or-pending commits per buffer

    for (int m0 = blockIdx.x * UT; m0 < s.n_out; m0 += gridDim.x * UT) {
        // ---- per-tile staging: neighbour ids, weightnet output, guidance ----
        __syncthreads();
        for (int i = tid; i < UT * K; i += UNT) {
            const int pp = i / K, m = m0 + pp;
            int v = -1;
            if (m < s.n_out) {
                const int64_t q = a.nei[(size_t)m * K + (i - pp * K)];
                if (q >= 0 && q < s.n_in) v = (int)q;
            }
            nei_s[i] = v;
        }
        {
            const int row = K * CMID, total = UT * row;
            const size_t base = (size_t)m0 * row;
            for (int i = tid; i < total; i += UNT) {
                const int pp = i / row;
                w_s[pp * pl.w_stride + (i - pp * row)] = (m0 + pp < s.n_out) ? a.weights[base + i] : 0.f;
            }
        }
        if (H > 0) {
            const int row = K * H, total = UT * row;
            const size_t base = (size_t)m0 * row;
            for (int i = tid; i < total; i += UNT) {
                const int pp = i / row;
                gd_s[pp * pl.gd_stride + (i - pp * row)] = (m0 + pp < s.n_out) ? a.guidance[base + i] : 0.f;
            }
        }

        int chunk = 0;
        for (int kk0 = 0; kk0 < KK; kk0 += CK, ++chunk) {
            const int buf = chunk & 1;
            const int c0 = kk0 / CMID;                         // first channel of this chunk
            __syncthreads();                                   // g_s free (previous contraction done), nei/w/gd visible
            // ---- stage the G chunk: [UT][K][CC] ----
            for (int i = tid; i < UT * K * CC; i += UNT) {
                const int cl = i % CC, pk = i / CC;
                const int pp = pk / K, k = pk - pp * K;
                const int c = c0 + cl, m = m0 + pp;
                float v = 0.f;
                if (c < C_cat && m < s.n_out) {
                    if (c < C_in) {
                        const int q = nei_s[pk];
                        if (q >= 0) {
                            v = __ldg(a.feats + (size_t)q * C_in + c);
                            if (H > 0) v *= gd_s[pp * pl.gd_stride + k * H + (c % H)];
                        }
                    } else {
                        v = __ldg(a.additional + ((size_t)m * K + k) * C_add + (c - C_in));
                    }
                }
                g_s[pp * pl.g_stride + k * CC + cl] = v;
            }
            // ---- wait until the tensor pipe has drained this buffer (MMAs of chunk-2) ----
            {
                const uint32_t u = buf ? uses1 : uses0;
                if (u > 0 && !umma::mbar_wait(&bars[buf], (u - 1) & 1)) __trap();
            }
            // ---- W chunk -> B operand (hi, lo), K-major core-matrix layout ----
            {
                unsigned char *Bh = B_base + (size_t)(buf * 2 + 0) * pl.b_bytes;
                unsigned char *Bl = B_base + (size_t)(buf * 2 + 1) * pl.b_bytes;
                constexpr int QN = CK / 4;
                for (int i = tid; i < C_out * QN; i += UNT) {
                    const int o = i / QN, q = i - o * QN;
                    const int kk = kk0 + 4 * q;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (kk < KK) v = __ldg(reinterpret_cast<const float4 *>(a.lin_w + (size_t)o * KK + kk));
                    float4 hi, lo;
                    umma::split_tf32(v.x, hi.x, lo.x); umma::split_tf32(v.y, hi.y, lo.y);
                    umma::split_tf32(v.z, hi.z, lo.z); umma::split_tf32(v.w, hi.w, lo.w);
                    *reinterpret_cast<float4 *>(Bh + (size_t)q * lbo_b + o * 16) = hi;
                    *reinterpret_cast<float4 *>(Bl + (size_t)q * lbo_b + o * 16) = lo;
                }
            }
            __syncthreads();                                   // g_s complete
            // ---- contraction 1 for kk_local in [g*KPT, (g+1)*KPT) of point p ----
            float acc[KPT];
#pragma unroll
            for (int t = 0; t < KPT; ++t) acc[t] = 0.f;
            {
                const float *gp = g_s + p * pl.g_stride;
                const float *wp = w_s + p * pl.w_stride;
                if (CMID >= KPT) {
                    // one channel (or part of one) per thread: cl fixed, j = j0 + t
                    const int cl = (g * KPT) / CMID, j0 = (g * KPT) % CMID;
                    for (int k = 0; k < K; ++k) {
                        const float gv = gp[k * CC + cl];
                        const float *wr = wp + k * CMID + j0;
                        if (KPT % 4 == 0 && CMID % 4 == 0) {
#pragma unroll
                            for (int t = 0; t < KPT; t += 4) {
                                const float4 wv = *reinterpret_cast<const float4 *>(wr + t);
                                acc[t] = fmaf(gv, wv.x, acc[t]); acc[t + 1] = fmaf(gv, wv.y, acc[t + 1]);
                                acc[t + 2] = fmaf(gv, wv.z, acc[t + 2]); acc[t + 3] = fmaf(gv, wv.w, acc[t + 3]);
                            }
                        } else {
#pragma unroll
                            for (int t = 0; t < KPT; ++t) acc[t] = fmaf(gv, wr[t], acc[t]);
                        }
                    }
                } else {
                    // several channels per thread: kk_local = g*KPT + t -> (cl, j)
                    for (int k = 0; k < K; ++k) {
                        float wv[CMID];
#pragma unroll
                        for (int j = 0; j < CMID; ++j) wv[j] = wp[k * CMID + j];
#pragma unroll
                        for (int t = 0; t < KPT; ++t) {
                            const int kl = g * KPT + t;
                            acc[t] = fmaf(gp[k * CC + kl / CMID], wv[kl % CMID], acc[t]);
                        }
                    }
                }
            }
            // ---- P -> out_p (optional) and -> A operand (hi, lo) ----
            {
                const int m = m0 + p;
                if (a.out_p && m < s.n_out) {
                    float *dst = a.out_p + (size_t)m * KK + kk0 + g * KPT;
#pragma unroll
                    for (int t = 0; t < KPT; t += 4)
                        if (kk0 + g * KPT + t < KK)
                            *reinterpret_cast<float4 *>(dst + t) = make_float4(acc[t], acc[t + 1], acc[t + 2], acc[t + 3]);
                }
                unsigned char *Ah = A_base + (size_t)(buf * 2 + 0) * pl.a_bytes;
                unsigned char *Al = A_base + (size_t)(buf * 2 + 1) * pl.a_bytes;
#pragma unroll
                for (int t = 0; t < KPT; t += 4) {
                    float4 hi, lo;
                    umma::split_tf32(acc[t], hi.x, lo.x); umma::split_tf32(acc[t + 1], hi.y, lo.y);
                    umma::split_tf32(acc[t + 2], hi.z, lo.z); umma::split_tf32(acc[t + 3], hi.w, lo.w);
                    const int q = (g * KPT + t) / 4;
                    *reinterpret_cast<float4 *>(Ah + (size_t)q * lbo_a + p * 16) = hi;
                    *reinterpret_cast<float4 *>(Al + (size_t)q * lbo_a + p * 16) = lo;
                }
            }
            umma::fence_proxy_async();                         // operand writes -> visible to the tensor pipe
            umma::fence_before_sync();
            __syncthreads();
            // ---- one thread feeds the tensor core ----
            if (tid == 0) {
                umma::fence_after_sync();
                const uint32_t ah = umma::smem_u32(A_base + (size_t)(buf * 2 + 0) * pl.a_bytes);
                const uint32_t al = umma::smem_u32(A_base + (size_t)(buf * 2 + 1) * pl.a_bytes);
                const uint32_t bh = umma::smem_u32(B_base + (size_t)(buf * 2 + 0) * pl.b_bytes);
                const uint32_t bl = umma::smem_u32(B_base + (size_t)(buf * 2 + 1) * pl.b_bytes);
#pragma unroll
                for (int ks = 0; ks < CK / 8; ++ks) {
                    const uint32_t ao = ks * 2 * lbo_a, bo = ks * 2 * lbo_b;
                    const uint64_t dah = umma::make_smem_desc(ah + ao, lbo_a, sbo);
                    const uint64_t dal = umma::make_smem_desc(al + ao, lbo_a, sbo);
                    const uint64_t dbh = umma::make_smem_desc(bh + bo, lbo_b, sbo);
                    const uint64_t dbl = umma::make_smem_desc(bl + bo, lbo_b, sbo);
                    umma::mma_tf32_ss(tmem_d, dal, dbh, idesc, (chunk > 0 || ks > 0) ? 1u : 0u);
                    umma::mma_tf32_ss(tmem_d, dah, dbl, idesc, 1u);
                    umma::mma_tf32_ss(tmem_d, dah, dbh, idesc, 1u);
                }
                umma::commit(&bars[buf]);
            }
            if (buf) ++uses1; else ++uses0;
        }
        // ---- epilogue: wait for the last commit (covers every MMA of the tile), D -> Y ----
        {
            const int lastbuf = (chunk - 1) & 1;
            const uint32_t u = lastbuf ? uses1 : uses0;
            if (!umma::mbar_wait(&bars[lastbuf], (u - 1) & 1)) __trap();
        }
        umma::fence_after_sync();
        if (warp < 4) {
            // M = 64 accumulator layout: row r lives in TMEM lane (r % 16) + 32 * (r / 16)
            const int row = warp * 16 + lane;
            const int m = m0 + row;
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
            for (int o0 = 0; o0 < C_out; o0 += 8) {
                float v[8];
                umma::tmem_ld8(taddr + o0, v);
                if (lane < 16 && m < s.n_out) {
                    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                    if (a.lin_b) {
                        b0 = __ldg(reinterpret_cast<const float4 *>(a.lin_b + o0));
                        b1 = __ldg(reinterpret_cast<const float4 *>(a.lin_b + o0 + 4));
                    }
                    float4 *dst = reinterpret_cast<float4 *>(a.out_y + (size_t)m * C_out + o0);
                    dst[0] = make_float4(v[0] + b0.x, v[1] + b0.y, v[2] + b0.z, v[3] + b0.w);
                    dst[1] = make_float4(v[4] + b1.x, v[5] + b1.y, v[6] + b1.z, v[7] + b1.w);
                }
            }
        }
        umma::fence_before_sync();
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(a.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// Descriptor self-test: D = A[M x K] * B[N x K]^T on one CTA, operands filled with a caller-chosen
// core-matrix layout, descriptors built from caller-chosen fields; dumps the RAW accumulator
// (128 TMEM lanes x N columns) so the host can verify the operand/accumulator conventions on hardware.
// ---------------------------------------------------------------------------------------------------
struct SelftestArgs {
    const float *A, *B;
    float *raw;                 // [128][N]
    int M, N, K;
    uint32_t fill_lbo_a, fill_sbo_a, fill_lbo_b, fill_sbo_b;     // where element (r,k) is written
    uint32_t desc_lbo_a, desc_sbo_a, desc_lbo_b, desc_sbo_b;     // what the descriptors claim
    uint32_t kstep_a, kstep_b;                                    // start-address advance per K=8 step
    uint32_t idesc;
    uint64_t desc_or;                                             // extra descriptor bits
    int split;                                                    // 1: 3xTF32
    uint32_t mn_flags;                                            // bit0: A is MN-major, bit1: B is MN-major
    int *status;
};

__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(SelftestArgs t)
{
    pdl_wait();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t a_bytes = 64 * 1024 / 2;      // generous fixed carve: A_hi, A_lo, B_hi, B_lo of 32 KB each
    unsigned char *Ah = smem_raw, *Al = smem_raw + a_bytes, *Bh = smem_raw + 2 * a_bytes, *Bl = smem_raw + 3 * a_bytes;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + 4 * a_bytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + 4 * a_bytes + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < 4 * a_bytes / 4; i += 128) reinterpret_cast<float *>(smem_raw)[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < t.M * t.K; i += 128) {
        const int r = i / t.K, k = i - r * t.K;
        float hi, lo;
        const float x = t.A[i];
        if (t.split) umma::split_tf32(x, hi, lo); else { hi = x; lo = 0.f; }
        const uint32_t off = (t.mn_flags & 1)
            ? (r % 4) * 4 + (r / 4) * t.fill_sbo_a + (k % 8) * 16 + (k / 8) * t.fill_lbo_a       // MN-major: 4 rows contiguous
            : (k / 4) * t.fill_lbo_a + (r / 8) * t.fill_sbo_a + (r % 8) * 16 + (k % 4) * 4;
        *reinterpret_cast<float *>(Ah + off) = hi;
        *reinterpret_cast<float *>(Al + off) = lo;
    }
    for (int i = tid; i < t.N * t.K; i += 128) {
        const int r = i / t.K, k = i - r * t.K;
        float hi, lo;
        const float x = t.B[i];
        if (t.split) umma::split_tf32(x, hi, lo); else { hi = x; lo = 0.f; }
        const uint32_t off = (t.mn_flags & 2)
            ? (r % 4) * 4 + (r / 4) * t.fill_sbo_b + (k % 8) * 16 + (k / 8) * t.fill_lbo_b
            : (k / 4) * t.fill_lbo_b + (r / 8) * t.fill_sbo_b + (r % 8) * 16 + (k % 4) * 4;
        *reinterpret_cast<float *>(Bh + off) = hi;
        *reinterpret_cast<float *>(Bl + off) = lo;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "r"(umma::smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) { umma::mbar_init(bar, 1); umma::fence_mbar_init(); }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = *tmem_slot;
    if (tid == 0) {
        for (int ks = 0; ks < t.K / 8; ++ks) {
            const uint64_t dah = umma::make_smem_desc(umma::smem_u32(Ah) + ks * t.kstep_a, t.desc_lbo_a, t.desc_sbo_a) | t.desc_or;
            const uint64_t dal = umma::make_smem_desc(umma::smem_u32(Al) + ks * t.kstep_a, t.desc_lbo_a, t.desc_sbo_a) | t.desc_or;
            const uint64_t dbh = umma::make_smem_desc(umma::smem_u32(Bh) + ks * t.kstep_b, t.desc_lbo_b, t.desc_sbo_b) | t.desc_or;
            const uint64_t dbl = umma::make_smem_desc(umma::smem_u32(Bl) + ks * t.kstep_b, t.desc_lbo_b, t.desc_sbo_b) | t.desc_or;
            if (t.split) {
                umma::mma_tf32_ss(tmem_d, dal, dbh, t.idesc, ks > 0 ? 1u : 0u);
                umma::mma_tf32_ss(tmem_d, dah, dbl, t.idesc, 1u);
                umma::mma_tf32_ss(tmem_d, dah, dbh, t.idesc, 1u);
            } else {
                umma::mma_tf32_ss(tmem_d, dah, dbh, t.idesc, ks > 0 ? 1u : 0u);
            }
        }
        umma::commit(bar);
    }
    const bool ok = umma::mbar_wait(bar, 0);
    if (!ok && tid == 0) *t.status = 1;
    umma::fence_after_sync();
    if (ok) {
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < t.N; c0 += 8) {
            float v[8];
            umma::tmem_ld8(taddr + c0, v);
            for (int j = 0; j < 8; ++j) t.raw[(size_t)tid * t.N + c0 + j] = v[j];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(256) : "memory");
}

// ---- host side ---------------------------------------------------------------------------------
bool pconv_forward_umma_supported(const pcfb_pconv_shape *s, bool has_lin) {
    if (!has_lin) return false;
    if (!(s->C_mid == 1 || s->C_mid == 4 || s->C_mid == 8 || s->C_mid == 16)) return false;
    if (s->C_out < 8 || s->C_out > 256 || s->C_out % 8 != 0) return false;
    const int KK = (s->C_in + s->C_add) * s->C_mid;
    if (KK % 4 != 0) return false;
    if (s->K < 1 || s->K > 64) return false;
    if (s->H != 0 && !((s->H == 1 || s->H == 2 || s->H == 4 || s->H == 8) && s->C_in % s->H == 0)) return false;
    return u_choose_kpt(*s) != 0;
}

size_t pconv_forward_umma_workspace(const pcfb_pconv_shape *) { return 0; }

template <int CMID, int KPT>
static int launch_umma(const UmmaArgs &a, const UPlan &pl, int grid, cudaStream_t st) {
    PCFB_CUDA(cudaFuncSetAttribute(pconv_fwd_umma_kernel<CMID, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)U_SMEM_BUDGET));
    launch_k(pconv_fwd_umma_kernel<CMID, KPT>, grid, UNT, pl.total, st, a);
    return check_launch("pconv_fwd_umma_kernel");
}

int pconv_forward_umma(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                       const float *additional, const float *guidance, const float *lin_w, const float *lin_b,
                       float *out_y, float *out_p, void *, size_t, cudaStream_t st)
{
    PCFB_REQUIRE(pconv_forward_umma_supported(s, lin_w != nullptr), "pcfb_pconv_forward: shape unsupported by the tcgen05 variant");
    PCFB_REQUIRE(((uintptr_t)lin_w % 16 == 0) && ((uintptr_t)out_y % 16 == 0) && (!out_p || (uintptr_t)out_p % 16 == 0) &&
                 (!lin_b || (uintptr_t)lin_b % 16 == 0), "pcfb_pconv_forward: tcgen05 variant needs 16-byte aligned lin_w/lin_b/out_y/out_p");
    if (s->n_out == 0) return PCFB_OK;
    UmmaArgs a{};
    a.s = *s;
    a.feats = feats; a.nei = nei; a.weights = weights; a.additional = additional; a.guidance = guidance;
    a.lin_w = lin_w; a.lin_b = lin_b; a.out_y = out_y; a.out_p = out_p;
    int cols = 32;
    while (cols < s->C_out) cols <<= 1;
    a.tmem_cols = cols;
    const int kpt = u_choose_kpt(*s);
    const UPlan pl = u_plan(*s, kpt);
    const int grid = max(1, min(ceil_div(s->n_out, UT), kNumSMs));
#define U_CASE(CMID)                                                                 \
    case CMID:                                                                       \
        if (kpt == 16) return launch_umma<CMID, 16>(a, pl, grid, st);                \
        if (kpt == 8) return launch_umma<CMID, 8>(a, pl, grid, st);                  \
        return launch_umma<CMID, 4>(a, pl, grid, st);
    switch (s->C_mid) {
        U_CASE(1)
        U_CASE(4)
        U_CASE(8)
        U_CASE(16)
    }
#undef U_CASE
    set_error("pcfb_pconv_forward: unreachable C_mid");
    return PCFB_ERR_UNSUPPORTED;
}

}  // namespace pcfb

// Hardware self-test of the UMMA descriptor conventions (used by tests/test_umma_selftest.py).
extern "C" int pcfb_selftest_umma(const float *A, const float *B, float *raw, int M, int N, int K,
                                  const uint32_t *h_params /* 12 host uint32 */, uint64_t desc_or, int split,
                                  int *status, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(A && B && raw && h_params && status, "pcfb_selftest_umma: null pointer");
    PCFB_REQUIRE((M == 64 || M == 128) && N >= 8 && N <= 256 && N % 8 == 0 && K >= 8 && K % 8 == 0 && K <= 64,
                 "pcfb_selftest_umma: bad sizes");
    SelftestArgs t{};
    t.A = A; t.B = B; t.raw = raw; t.M = M; t.N = N; t.K = K;
    t.fill_lbo_a = h_params[0]; t.fill_sbo_a = h_params[1]; t.fill_lbo_b = h_params[2]; t.fill_sbo_b = h_params[3];
    t.desc_lbo_a = h_params[4]; t.desc_sbo_a = h_params[5]; t.desc_lbo_b = h_params[6]; t.desc_sbo_b = h_params[7];
    t.kstep_a = h_params[8]; t.kstep_b = h_params[9]; t.idesc = h_params[10]; t.mn_flags = h_params[11];
    t.desc_or = desc_or; t.split = split; t.status = status;
    const size_t smem = 4 * 32 * 1024 + 64;
    PCFB_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_k(umma_selftest_kernel, 1, 128, smem, static_cast<cudaStream_t>(stream), t);
    return check_launch("umma_selftest_kernel");
}
