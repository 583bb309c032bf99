// Fused PointConv / PointConvFormer forward, tcgen05 variant, pipelined (sm_100a).
//
// Same maths and operand conventions as pconv_umma.cu (3xTF32 Linear on tcgen05 with TMEM accumulators, 64-point
// tiles = UMMA M), rebuilt around memory-level parallelism -- the first version staged every chunk with
// dependent load->store loops and ran at 4% of the HBM roofline:
//   * all global->shared traffic is cp.async (LDGSTS): the weightnet tile, the gathered feature rows and the
//     additional features land directly in shared memory, one commit group per chunk, prefetch distance 1;
//   * the Linear weights are split into (hi, lo) tf32 and laid out in the UMMA core-matrix order ONCE per call by
//     a prep kernel, so the B operand of every chunk is a contiguous cp.async copy (triple buffered);
//   * contraction 1 is register tiled TC channels x TJ weights per thread (2 LDS.128 per 16 FMA instead of 17
//     scalar loads) and issued as packed FFMA2 (fma.rn.f32x2);
//   * guidance is applied inside the contraction (the staged tile holds raw features);
//   * the next tile's neighbour table is prefetched while the current tile computes.
// Thread t: point p = t % 64 (lanes = consecutive points, conflict-free LDS.128 with rows padded to 4*odd floats),
// g = t / 64 selects the (channel group, weight group) of the chunk.
#include "common.cuh"
#include "umma.cuh"

namespace pcfb {

constexpr int VT = 64;         // points per tile = UMMA M
constexpr int VNT = 256;
constexpr int VK = 16;         // neighbours (every shipped config uses K = 16; other K -> simple/SIMT kernels)
constexpr int VKPT = 8;        // outputs per thread per chunk
constexpr int VCK = 4 * VKPT;  // kk columns per chunk
constexpr size_t V_SMEM_BUDGET = 225 * 1024;

struct U2Args {
    pcfb_pconv_shape s;
    const float *feats, *weights, *additional, *guidance, *lin_b;
    const float *w_prep;       // [chunks][hi,lo][CK/4][C_out][4]
    const int64_t *nei;
    float *out_y, *out_p;
    int tmem_cols, n_chunks, vec_ok;   // vec_ok: multi-float cp.async path usable for the G rows
};

__host__ __device__ inline int pad4odd(int x) {      // smallest y >= x with y % 4 == 0 and (y / 4) odd
    int y = (x + 3) & ~3;
    if (((y >> 2) & 1) == 0) y += 4;
    return y;
}

struct V2Plan {
    int CC, GS, GDS;
    uint32_t a_bytes, b_bytes;
    size_t off_A, off_B, off_g, off_gd, off_nei, off_bar, total;
};

__host__ __device__ inline V2Plan v2_plan(const pcfb_pconv_shape &s) {
    V2Plan pl;
    pl.CC = VCK / s.C_mid;
    pl.GS = (pl.CC >= 4) ? pad4odd(VK * pl.CC) : (VK * pl.CC + 2);
    pl.GDS = pad4odd(VK * (s.H > 0 ? s.H : 1));
    pl.a_bytes = VT * VCK * 4;
    pl.b_bytes = s.C_out * VCK * 4;
    size_t o = 0;
    pl.off_A = o;   o += 4 * (size_t)pl.a_bytes;            // [2 bufs][hi|lo]
    pl.off_B = o;   o += 6 * (size_t)pl.b_bytes;            // [3 bufs][hi|lo]
    pl.off_g = o;   o += 3 * (size_t)VT * pl.GS * 4;        // 3-stage ring
    o = align_up(o, 16);
    pl.off_gd = o;  o += (s.H > 0) ? (size_t)VT * pl.GDS * 4 : 0;
    pl.off_nei = o; o += 2 * (size_t)VT * VK * 8;           // raw int64, double buffered across tiles
    o = align_up(o, 16);
    pl.off_bar = o; o += 64;
    pl.total = o;
    return pl;
}

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(umma::smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(void *dst, const void *src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" :: "r"(umma::smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async8_ca(void *dst, const void *src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" :: "r"(umma::smem_u32(dst)), "l"(src), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" :: "r"(umma::smem_u32(dst)), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

__device__ __forceinline__ void ffma2(float &a0, float &a1, float x0, float x1, float y0, float y1) {
    // (a0, a1) += (x0*y0, x1*y1) as one packed FFMA2
    unsigned long long acc, xs, ys;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(xs) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ys) : "f"(y0), "f"(y1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xs), "l"(ys));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc));
}

__device__ __forceinline__ float4 ldg_nc_f4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// W [C_out][KK] -> (hi, lo) tf32 pairs in UMMA K-major core-matrix order, one contiguous block per chunk
__global__ void prep_w_kernel(const float *__restrict__ W, int C_out, int KK, int CK, int n_chunks, float *__restrict__ out)
{
    pdl_wait();
    const int units = CK / 4;
    const int64_t total = (int64_t)n_chunks * units * C_out;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(i % C_out);
        const int q = (int)((i / C_out) % units);
        const int ch = (int)(i / ((int64_t)C_out * units));
        const int kk = ch * CK + q * 4;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (kk + e < KK) ? W[(size_t)o * KK + kk + e] : 0.f;
        float4 hi, lo;
        umma::split_tf32(v[0], hi.x, lo.x); umma::split_tf32(v[1], hi.y, lo.y);
        umma::split_tf32(v[2], hi.z, lo.z); umma::split_tf32(v[3], hi.w, lo.w);
        const size_t blk = (size_t)C_out * CK;                    // floats per (chunk, hi|lo)
        const size_t off = (size_t)q * C_out * 4 + (size_t)o * 4;
        *reinterpret_cast<float4 *>(out + ((size_t)ch * 2 + 0) * blk + off) = hi;
        *reinterpret_cast<float4 *>(out + ((size_t)ch * 2 + 1) * blk + off) = lo;
    }
}

void prep_w_launch(const float *lin_w, int C_out, int KK, int CK, int n_chunks, float *out, cudaStream_t st) {
    const int64_t total = (int64_t)n_chunks * (CK / 4) * C_out;
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    launch_k(prep_w_kernel, blocks < 1 ? 1 : blocks, 256, 0, st, lin_w, C_out, KK, CK, n_chunks, out);
}

// Tiling of one chunk (CK = 32 kk columns = CC channels x CMID weights) over the 4 thread groups of a point:
//   TJ = min(CMID, 4) weights x TC = 8 / TJ channels per thread; NJ = CMID / TJ weight groups, NCG = 4 / NJ
//   channel groups.  The TJ x 16 weightnet values a thread needs are the same for every chunk of the tile, so
//   they live in REGISTERS (loaded once per tile straight from global memory) -- no shared-memory copy of the
//   weightnet tile, and the inner loop reads only TC gathered features per neighbour from shared memory.
template <int CMID, bool GUIDE>
__global__ void __launch_bounds__(VNT, 2) pconv_fwd_umma2_kernel(U2Args a)
{
    pdl_wait();
    constexpr int CK = VCK, KPT = VKPT, K = VK;
    constexpr int CC = CK / CMID;
    constexpr int TJ = CMID < 4 ? CMID : 4;
    constexpr int TC = KPT / TJ;
    constexpr int NJ = CMID / TJ;
    constexpr int NCG = 4 / NJ;
    static_assert(NCG * TC == CC, "tiling does not cover the chunk");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const pcfb_pconv_shape &s = a.s;
    const V2Plan pl = v2_plan(s);
    unsigned char *A_base = smem_raw + pl.off_A;
    unsigned char *B_base = smem_raw + pl.off_B;
    float *g_s = reinterpret_cast<float *>(smem_raw + pl.off_g);
    float *gd_s = reinterpret_cast<float *>(smem_raw + pl.off_gd);
    long long *nei_s = reinterpret_cast<long long *>(smem_raw + pl.off_nei);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + pl.off_bar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + pl.off_bar + 32);

    const int C_in = s.C_in, C_add = s.C_add, C_cat = C_in + C_add, KK = C_cat * CMID, C_out = s.C_out, H = s.H;
    const int n_in = s.n_in, n_out = s.n_out;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p = tid & (VT - 1), g = tid >> 6;
    const int jg = g % NJ, cg = g / NJ;
    const int n_chunks = a.n_chunks;
    const int GS = pl.GS;

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
    const uint32_t idesc = umma::make_idesc_tf32(VT, C_out);
    const uint32_t lbo_a = VT * 16, lbo_b = (uint32_t)C_out * 16, sbo = 128;
    uint32_t uses0 = 0, uses1 = 0;

    auto issue_nei = [&](int m0, long long *dst) {                 // raw int64 rows of the tile, 16 B per cp.async
        const size_t base = (size_t)m0 * K;
        for (int i = tid * 2; i < VT * K; i += VNT * 2)
            cp_async16(dst + i, a.nei + base + i, m0 + i / K < n_out);
    };
    auto issue_g = [&](int m0, int chunk, const long long *nei_cur) {
        float *gdst = g_s + (size_t)(chunk % 3) * VT * GS;
        const int c0 = chunk * CC;
        constexpr int PIECE = CC >= 4 ? 4 : CC;
        if (a.vec_ok) {
            constexpr int PPR = CC / PIECE;
            for (int i = tid; i < VT * K * PPR; i += VNT) {
                const int piece = i % PPR, pk = i / PPR;
                const int pp = pk / K, k = pk - pp * K;
                const int c = c0 + piece * PIECE, m = m0 + pp;
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
                float *dst = gdst + pp * GS + k * CC + piece * PIECE;
                if (PIECE == 4) cp_async16_ca(dst, src, valid);
                else if (PIECE == 2) cp_async8_ca(dst, src, valid);
                else cp_async4(dst, src, valid);
            }
        } else {
            for (int i = tid; i < VT * K * CC; i += VNT) {
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
                cp_async4(gdst + pp * GS + k * CC + cl, src, valid);
            }
        }
    };
    auto issue_b = [&](int chunk) {
        unsigned char *bdst = B_base + (size_t)(chunk % 3) * 2 * pl.b_bytes;
        const unsigned char *bsrc = reinterpret_cast<const unsigned char *>(a.w_prep) + (size_t)chunk * 2 * pl.b_bytes;
        for (uint32_t i = tid * 16; i < 2 * pl.b_bytes; i += VNT * 16) cp_async16(bdst + i, bsrc + i, true);
    };

    const int first = blockIdx.x * VT;
    if (first < n_out) issue_nei(first, nei_s);
    cp_async_commit();
    int tile_it = 0;
    for (int m0 = first; m0 < n_out; m0 += gridDim.x * VT, ++tile_it) {
        const long long *nei_cur = nei_s + (size_t)(tile_it & 1) * VT * K;
        long long *nei_nxt = nei_s + (size_t)((tile_it + 1) & 1) * VT * K;
        cp_async_wait<0>();
        __syncthreads();                                   // nei_cur landed; previous tile fully consumed
        // ---- tile prologue: [guidance tile, G0, B0, B1] [G1] [G2 + next tile's neighbour table] ----
        if (GUIDE) {
            const int grow = K * H;
            const size_t gbase = (size_t)m0 * grow;
            if ((grow & 3) == 0 && ((uintptr_t)a.guidance % 16 == 0)) {
                for (int i = tid * 4; i < VT * grow; i += VNT * 4) {
                    const int pp = i / grow;
                    cp_async16(gd_s + pp * pl.GDS + (i - pp * grow), a.guidance + gbase + i, m0 + pp < n_out);
                }
            } else {
                for (int i = tid; i < VT * grow; i += VNT) {
                    const int pp = i / grow;
                    cp_async4(gd_s + pp * pl.GDS + (i - pp * grow), a.guidance + gbase + i, m0 + pp < n_out);
                }
            }
        }
        issue_g(m0, 0, nei_cur);
        issue_b(0);
        if (n_chunks > 1) issue_b(1);
        cp_async_commit();
        if (n_chunks > 1) issue_g(m0, 1, nei_cur);
        cp_async_commit();
        if (n_chunks > 2) issue_g(m0, 2, nei_cur);
        {
            const int m_next = m0 + gridDim.x * VT;
            if (m_next < n_out) issue_nei(m_next, nei_nxt);
        }
        cp_async_commit();
        // ---- this thread's weightnet values: w[m][k][jg*TJ .. +TJ) for all k, kept in registers for the tile ----
        float wreg[K][TJ];
        {
            const int m = m0 + p;
            const float *wsrc = a.weights + (size_t)(m < n_out ? m : 0) * K * CMID + jg * TJ;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (TJ == 4) {
                    const float4 v = ldg_nc_f4(wsrc + k * CMID);
                    wreg[k][0] = v.x; wreg[k][1] = v.y; wreg[k][2] = v.z; wreg[k][3] = v.w;
                } else {
#pragma unroll
                    for (int t = 0; t < TJ; ++t) wreg[k][t] = __ldg(wsrc + k * CMID + t);
                }
            }
            if (m >= n_out) {
#pragma unroll
                for (int k = 0; k < K; ++k)
#pragma unroll
                    for (int t = 0; t < TJ; ++t) wreg[k][t] = 0.f;
            }
        }

        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int buf = chunk & 1;
            cp_async_wait<2>();                                // G(chunk) has landed (two younger groups may be in flight)
            __syncthreads();
            // ---- contraction 1: acc[tc*TJ + tj] = sum_k G[k][cg*TC+tc] * w[k][jg*TJ+tj] ----
            float acc[KPT];
#pragma unroll
            for (int t = 0; t < KPT; ++t) acc[t] = 0.f;
            {
                const float *gp = g_s + (size_t)(chunk % 3) * VT * GS + p * GS + cg * TC;
                int hofs[TC];
                if (GUIDE) {
#pragma unroll
                    for (int t = 0; t < TC; ++t) {
                        const int c = chunk * CC + cg * TC + t;
                        hofs[t] = (c < C_in) ? (c % H) : -1;
                    }
                }
                const float *gdp = gd_s + p * pl.GDS;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float gv[TC];
                    if (TC == 2) {
                        const float2 v = *reinterpret_cast<const float2 *>(gp + k * CC);
                        gv[0] = v.x; gv[1] = v.y;
                    } else {
#pragma unroll
                        for (int t = 0; t < TC; t += 4) {
                            const float4 v = *reinterpret_cast<const float4 *>(gp + k * CC + t);
                            gv[t] = v.x; gv[t + 1] = v.y; gv[t + 2] = v.z; gv[t + 3] = v.w;
                        }
                    }
                    if (GUIDE) {
#pragma unroll
                        for (int t = 0; t < TC; ++t)
                            if (hofs[t] >= 0) gv[t] *= gdp[k * H + hofs[t]];
                    }
                    if (TJ >= 2) {
#pragma unroll
                        for (int tc = 0; tc < TC; ++tc)
#pragma unroll
                            for (int tj = 0; tj < TJ; tj += 2)
                                ffma2(acc[tc * TJ + tj], acc[tc * TJ + tj + 1], gv[tc], gv[tc], wreg[k][tj], wreg[k][tj + 1]);
                    } else {
#pragma unroll
                        for (int tc = 0; tc < TC; tc += 2)
                            ffma2(acc[tc], acc[tc + 1], gv[tc], gv[tc + 1], wreg[k][0], wreg[k][0]);
                    }
                }
            }
            // ---- A operand buffer must be free: MMA of chunk-2 (same buffer) complete ----
            {
                const uint32_t u = buf ? uses1 : uses0;
                if (u > 0 && !umma::mbar_wait(&bars[buf], (u - 1) & 1)) __trap();
            }
            {
                unsigned char *Ah = A_base + (size_t)(buf * 2 + 0) * pl.a_bytes;
                unsigned char *Al = A_base + (size_t)(buf * 2 + 1) * pl.a_bytes;
                const int m = m0 + p;
#pragma unroll
                for (int u = 0; u < KPT / 4; ++u) {
                    const int q = (TJ == 4) ? ((cg * TC + u) * (CMID / 4) + jg) : ((cg * TC) / 4 + u);
                    float4 hi, lo;
                    umma::split_tf32(acc[4 * u], hi.x, lo.x); umma::split_tf32(acc[4 * u + 1], hi.y, lo.y);
                    umma::split_tf32(acc[4 * u + 2], hi.z, lo.z); umma::split_tf32(acc[4 * u + 3], hi.w, lo.w);
                    *reinterpret_cast<float4 *>(Ah + (size_t)q * lbo_a + p * 16) = hi;
                    *reinterpret_cast<float4 *>(Al + (size_t)q * lbo_a + p * 16) = lo;
                    if (a.out_p && m < n_out) {
                        const int kk = chunk * CK + q * 4;
                        if (kk < KK)
                            *reinterpret_cast<float4 *>(a.out_p + (size_t)m * KK + kk) =
                                make_float4(acc[4 * u], acc[4 * u + 1], acc[4 * u + 2], acc[4 * u + 3]);
                    }
                }
            }
            cp_async_wait<1>();                                // B(chunk) landed (it travels one group behind G(chunk+1))
            umma::fence_proxy_async();
            umma::fence_before_sync();
            __syncthreads();                                   // A written, B visible, g_s[chunk%3] consumed by everyone
            if (tid == 0) {
                umma::fence_after_sync();
                const uint32_t ah = umma::smem_u32(A_base + (size_t)(buf * 2 + 0) * pl.a_bytes);
                const uint32_t al = umma::smem_u32(A_base + (size_t)(buf * 2 + 1) * pl.a_bytes);
                const uint32_t bh = umma::smem_u32(B_base + (size_t)((chunk % 3) * 2 + 0) * pl.b_bytes);
                const uint32_t bl = umma::smem_u32(B_base + (size_t)((chunk % 3) * 2 + 1) * pl.b_bytes);
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
            // ---- prefetch: G(chunk+3) into the ring slot just consumed, B(chunk+2) into the slot MMA(chunk-1) read ----
            if (chunk + 3 < n_chunks) issue_g(m0, chunk + 3, nei_cur);
            if (chunk + 2 < n_chunks) {
                if (chunk >= 1) {
                    const int pb = (chunk - 1) & 1;
                    const uint32_t u = pb ? uses1 : uses0;
                    if (!umma::mbar_wait(&bars[pb], (u - 1) & 1)) __trap();
                }
                issue_b(chunk + 2);
            }
            cp_async_commit();                                 // exactly one group per iteration keeps the counts uniform
        }
        // ---- epilogue ----
        {
            const int lastbuf = (n_chunks - 1) & 1;
            const uint32_t u = lastbuf ? uses1 : uses0;
            if (!umma::mbar_wait(&bars[lastbuf], (u - 1) & 1)) __trap();
        }
        umma::fence_after_sync();
        if (warp < 4) {
            const int row = warp * 16 + lane;                  // M = 64: row r <-> TMEM lane (r % 16) + 32 * (r / 16)
            const int m = m0 + row;
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
            for (int o0 = 0; o0 < C_out; o0 += 8) {
                float v[8];
                umma::tmem_ld8(taddr + o0, v);
                if (lane < 16 && m < n_out) {
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
    cp_async_wait<0>();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(a.tmem_cols) : "memory");
}

// ---- host side ---------------------------------------------------------------------------------
bool pconv_forward_umma2_supported(const pcfb_pconv_shape *s, bool has_lin) {
    if (!has_lin) return false;
    if (s->K != VK) return false;
    if (!(s->C_mid == 1 || s->C_mid == 4 || s->C_mid == 8 || s->C_mid == 16)) return false;
    if (s->C_out < 8 || s->C_out > 256 || s->C_out % 8 != 0) return false;
    if (((s->C_in + s->C_add) * s->C_mid) % 4 != 0) return false;
    if (s->H != 0 && !((s->H == 1 || s->H == 2 || s->H == 4 || s->H == 8) && s->C_in % s->H == 0)) return false;
    return v2_plan(*s).total <= V_SMEM_BUDGET;
}

size_t pconv_forward_umma2_workspace(const pcfb_pconv_shape *s) {
    const int KK = (s->C_in + s->C_add) * s->C_mid;
    const int n_chunks = ceil_div(KK, VCK);
    return align_up((size_t)n_chunks * 2 * s->C_out * VCK * sizeof(float), 256);
}

template <int CMID, bool GUIDE>
static int launch_v2(const U2Args &a, const V2Plan &pl, cudaStream_t st) {
    PCFB_CUDA(cudaFuncSetAttribute(pconv_fwd_umma2_kernel<CMID, GUIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM_BUDGET));
    const int per_sm = (pl.total + 1024 <= 113 * 1024) ? 2 : 1;
    const int grid = max(1, min(ceil_div(a.s.n_out, VT), kNumSMs * per_sm));
    launch_k(pconv_fwd_umma2_kernel<CMID, GUIDE>, grid, VNT, pl.total, st, a);
    return check_launch("pconv_fwd_umma2_kernel");
}

int pconv_forward_umma2(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                        const float *additional, const float *guidance, const float *lin_w, const float *lin_b,
                        float *out_y, float *out_p, void *workspace, size_t workspace_bytes, cudaStream_t st)
{
    PCFB_REQUIRE(pconv_forward_umma2_supported(s, lin_w != nullptr), "pcfb_pconv_forward: shape unsupported by the tcgen05 variant");
    PCFB_REQUIRE(((uintptr_t)lin_w % 16 == 0) && ((uintptr_t)out_y % 16 == 0) && (!out_p || (uintptr_t)out_p % 16 == 0) &&
                 (!lin_b || (uintptr_t)lin_b % 16 == 0) && ((uintptr_t)weights % 16 == 0) && ((uintptr_t)nei % 16 == 0),
                 "pcfb_pconv_forward: tcgen05 variant needs 16-byte aligned weights/nei/lin_w/lin_b/out_y/out_p");
    const size_t need = pconv_forward_umma2_workspace(s);
    if (!workspace || workspace_bytes < need) { set_error("pcfb_pconv_forward: workspace %zu < %zu", workspace_bytes, need); return PCFB_ERR_WORKSPACE; }
    if (s->n_out == 0) return PCFB_OK;
    const V2Plan pl = v2_plan(*s);
    const int KK = (s->C_in + s->C_add) * s->C_mid;
    U2Args a{};
    a.s = *s;
    a.feats = feats; a.nei = nei; a.weights = weights; a.additional = additional; a.guidance = guidance;
    a.lin_b = lin_b; a.out_y = out_y; a.out_p = out_p;
    a.w_prep = static_cast<const float *>(workspace);
    a.n_chunks = ceil_div(KK, VCK);
    int cols = 32;
    while (cols < s->C_out) cols <<= 1;
    a.tmem_cols = cols;
    // multi-float gathers need every piece of a (point, k) row aligned and not straddling the feats/additional boundary
    const int piece = pl.CC >= 4 ? 4 : pl.CC;
    a.vec_ok = (s->C_in % piece == 0) && (s->C_add % piece == 0) && ((uintptr_t)feats % 16 == 0) &&
               (s->C_add == 0 || (uintptr_t)additional % 16 == 0);
    int rc;
    {
        const int64_t total = (int64_t)a.n_chunks * (VCK / 4) * s->C_out;
        const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
        launch_k(prep_w_kernel, blocks < 1 ? 1 : blocks, 256, 0, st, lin_w, s->C_out, KK, VCK, a.n_chunks, static_cast<float *>(workspace));
        if ((rc = check_launch("prep_w_kernel"))) return rc;
    }
    const bool guide = s->H > 0;
#define V2_CASE(CMID)                                                                \
    case CMID:                                                                       \
        return guide ? launch_v2<CMID, true>(a, pl, st) : launch_v2<CMID, false>(a, pl, st);
    switch (s->C_mid) {
        V2_CASE(1)
        V2_CASE(4)
        V2_CASE(8)
        V2_CASE(16)
    }
#undef V2_CASE
    set_error("pcfb_pconv_forward: unreachable C_mid");
    return PCFB_ERR_UNSUPPORTED;
}

}  // namespace pcfb
