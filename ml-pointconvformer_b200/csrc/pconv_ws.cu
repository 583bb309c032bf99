// Fused PointConv / PointConvFormer forward, warp-specialised tcgen05 variant (sm_100a).
//
//   P[m, c*16 + j] = sum_k G[m,k,c] * w[m,k,j]          (CUDA cores, fp32 FFMA2)
//   Y[m, :]        = P[m, :] . W^T + b                  (tcgen05.mma kind::tf32, 3xTF32, TMEM accumulator)
//
// What changed against pconv_umma2.cu (whose ncu profile showed 13% FFMA2 among 934 warp instructions per point,
// the rest address arithmetic of 8-byte gather pieces, block barriers and the tf32 split):
//   * no block barrier in the main loop.  11 compute warps own 8 points each and run a PRIVATE cp.async ring for the
//     gathered rows (warp-level sync only); one extra warp issues every tcgen05.mma and streams the prepared Linear
//     weights with cp.async.bulk; hand-offs are mbarriers (A full/empty, B full/empty, D full/empty);
//   * gathers are issued from per-lane row pointers held in registers (4 neighbour rows per lane), 16 bytes per
//     cp.async, no index arithmetic, no shared-memory copy of the neighbour table;
//   * the accumulator in TMEM is double buffered, the epilogue of tile t runs inside tile t+1;
//   * tile = 88 points (rows 88..127 of the M=128 UMMA read stale shared memory and are never stored).
// Lane l of a compute warp: point pl = l / 4 of the warp's 8, weight quarter jq = l % 4 (weights 4jq .. 4jq+3).  Its
// 16x4 weightnet values stay in registers for the whole tile.
#include "common.cuh"
#include "umma.cuh"
#include <cstdlib>

// Ablation switches used while profiling (scripts/ws_ablate.sh): compile with -DPCFB_WS_ABLATE=1 and set PCFB_WS_DEBUG
// to a bit mask (1 no gather, 2 no MMA hand-off, 4 no FMA loop, 8 no MMAs, 32 no B refill, 64 no proxy fence, 128 no A
// stores).  Results are wrong with any bit set; the shipped library compiles them out.
#ifndef PCFB_WS_ABLATE
#define PCFB_WS_ABLATE 0
#endif
#define WS_DBG(bit) (PCFB_WS_ABLATE && (a.dbg & (bit)))

namespace pcfb {

constexpr int WS_NW = 11;                        // compute warps (12 warps -> 168 registers per thread: no spills, no rematerialised indices)
constexpr int WS_PT = WS_NW * 8;                 // points per tile
constexpr int WS_EPI = (WS_PT + 31) / 32;        // warps that drain the accumulator
constexpr int WS_NT = (WS_NW + 1) * 32;          // threads
constexpr int WS_K = 16;
constexpr int WS_CG = 8;                         // channels per ring stage (x 8 neighbours)
constexpr int WS_PSTRIDE = 8 * WS_CG + 4;        // floats per point per stage; +4 shifts each point by 4 banks
constexpr int WS_STAGE = 8 * WS_PSTRIDE;         // floats per warp per stage
constexpr int WS_CK = 32;                        // kk columns per MMA chunk (= 2 channels x 16 weights)
// bytes between K-units of the A operand: rows x 16 B, +32 so that the four K-units a quarter-warp stores to (one per
// weight quarter jq) start 8 banks apart -- with the unpadded stride every STS.128 of the tile was a 4-way bank conflict
// (ncu: 16 wavefronts per instruction instead of 4, 47% of all shared-memory wavefronts of the kernel)
constexpr uint32_t WS_LBO_A = WS_PT * 16 + 32;
constexpr uint32_t WS_A_HALF = (WS_CK / 4) * WS_LBO_A;   // bytes of one (hi | lo) A chunk
constexpr size_t WS_SMEM_MAX = 227 * 1024;

struct WsArgs {
    pcfb_pconv_shape s;
    const float *feats, *weights, *additional, *guidance, *lin_b;
    const float *w_prep;       // [chunks][hi,lo][CK/4][C_out][4]   (prep_w_kernel of pconv_umma2.cu)
    const int64_t *nei;
    float *out_y, *out_p;
    int tmem_cols;             // columns of ONE accumulator buffer (power of two >= 32)
    int n_groups, n_chunks, n_tiles, sb, na;   // sb: B ring slots, na: A buffers
    int dbg;                   // ablation switches for profiling (PCFB_WS_DEBUG): 1 no gather, 2 no MMA hand-off, 4 no FMA loop
};

struct WsPlan {
    uint32_t b_bytes;          // one B chunk (hi + lo)
    size_t off_ring, off_A, off_B, off_bar, total;
};

__host__ __device__ inline WsPlan ws_plan(int C_out, int stages, int sb, int na) {
    WsPlan pl;
    pl.b_bytes = (uint32_t)C_out * WS_CK * 4 * 2;
    size_t o = 0;
    pl.off_ring = o; o += (size_t)WS_NW * stages * WS_STAGE * 4;
    o = align_up(o, 128);
    pl.off_A = o;    o += (size_t)na * 2 * WS_A_HALF;
    pl.off_B = o;    o += (size_t)sb * pl.b_bytes;          // also the overrun area of the last A unit (rows 88..127)
    o = align_up(o, 16);
    pl.off_bar = o;  o += 256;
    pl.total = o;
    return pl;
}

namespace ws {
__device__ __forceinline__ void cp16(uint32_t dst, const void *src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void commit_group() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void wait_group() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

__device__ __forceinline__ void ffma2(float &a0, float &a1, float x, float y0, float y1) {
    unsigned long long acc, xs, ys;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(xs) : "f"(x));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ys) : "f"(y0), "f"(y1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xs), "l"(ys));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc));
}
// x = hi + lo with hi on the tf32 grid (round to nearest, ties away): 3 integer/fp instructions instead of the
// multi-instruction cvt.rna.tf32 sequence
__device__ __forceinline__ void split(float x, float &hi, float &lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xffffe000u);      // rounded, not truncated by the tensor core (umma.cuh)
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {          // true in exactly one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
// lane-predicated forms (pred != 0 executes): the predicate lives inside the asm so the caller's control flow stays warp-uniform
__device__ __forceinline__ void mbar_expect_tx_if(uint64_t *bar, uint32_t bytes, uint32_t pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}\n"
                 :: "r"(umma::smem_u32(bar)), "r"(bytes), "r"(pred) : "memory");
}
__device__ __forceinline__ void bulk_g2s_if(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar, uint32_t pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                 "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}\n"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(umma::smem_u32(bar)), "r"(pred) : "memory");
}
__device__ __forceinline__ void mma_if(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate, uint32_t pred) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(pred) : "memory");
}
__device__ __forceinline__ void commit_if(uint64_t *bar, uint32_t pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
                 :: "r"(umma::smem_u32(bar)), "r"(pred) : "memory");
}
__device__ __forceinline__ void wait_or_trap(uint64_t *bar, uint32_t parity) {
    if (!umma::mbar_wait(bar, parity)) __trap();
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];\n" :: "l"(p)); }
}  // namespace ws

// barrier slots
enum { WS_A_FULL = 0, WS_A_EMPTY = 4, WS_D_FULL = 8, WS_D_EMPTY = 10, WS_B_FULL = 12, WS_B_EMPTY = 20, WS_NBAR = 28 };   // <= 4 A buffers, <= 8 B slots

// GQ: 0 = no guidance, 1 = H in {1,2,4} (one quad of head values per neighbour), 2 = H == 8 (two quads)
template <int STAGES, int GQ>
__global__ void __launch_bounds__(WS_NT, 1) pconv_fwd_ws_kernel(WsArgs a)
{
    pdl_wait();
    constexpr int K = WS_K, S = STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const pcfb_pconv_shape &s = a.s;
    const int C_in = s.C_in, C_add = s.C_add, C_cat = C_in + C_add, KK = C_cat * 16, C_out = s.C_out, H = s.H;
    const int n_in = s.n_in, n_out = s.n_out;
    const int n_chunks = a.n_chunks, n_tiles = a.n_tiles, SB = a.sb, NA = a.na;
    const WsPlan pl_ = ws_plan(C_out, S, SB, NA);
    float *ring_all = reinterpret_cast<float *>(smem_raw + pl_.off_ring);
    unsigned char *A_base = smem_raw + pl_.off_A;
    unsigned char *B_base = smem_raw + pl_.off_B;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + pl_.off_bar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + pl_.off_bar + WS_NBAR * 8);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // broadcast: the compiler can treat the role branch as warp-uniform

    if (warp == WS_NW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "r"(umma::smem_u32(tmem_slot)), "r"(2 * a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            umma::mbar_init(&bars[WS_A_FULL + i], WS_NW);
            umma::mbar_init(&bars[WS_A_EMPTY + i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&bars[WS_D_FULL + i], 1);
            umma::mbar_init(&bars[WS_D_EMPTY + i], WS_EPI);
        }
        for (int i = 0; i < 8; ++i) {
            umma::mbar_init(&bars[WS_B_FULL + i], 1);
            umma::mbar_init(&bars[WS_B_EMPTY + i], 1);
        }
        umma::fence_mbar_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles of this CTA

    if (warp == WS_NW) {
        // ==================== MMA issuer + Linear-weight producer ====================
        // The whole warp walks the loop in uniform control flow (the role branch is on a shuffled, hence provably
        // uniform, warp index; waits, counters and descriptors stay in uniform registers); one elected lane issues the
        // tcgen05 / bulk-copy instructions.  With a plain `if (lane == 0)` loop every UTCHMMA was wrapped in
        // R2UR x6 + ELECT + BRA.U.ANY and this single thread limited the kernel (~1000 cycles per 12-MMA chunk).
        const uint32_t idesc = umma::make_idesc_tf32(128, C_out);
        const uint32_t lbo_b = (uint32_t)C_out * 16, sbo = 128;
        const uint32_t bbytes = pl_.b_bytes;
        const unsigned char *wsrc = reinterpret_cast<const unsigned char *>(a.w_prep);
        const int total = my_tiles * n_chunks;
        const uint32_t a_u32 = umma::smem_u32(A_base), b_u32 = umma::smem_u32(B_base);
        // descriptors without the start address; the address field is the low 14 bits (bytes >> 4) and never carries
        // out of it for shared-memory addresses
        const uint64_t da0 = umma::make_smem_desc(0, WS_LBO_A, sbo), db0 = umma::make_smem_desc(0, lbo_b, sbo);
        if (ws::elect_one()) {
            for (int j = 0; j < SB && j < total; ++j) {
                ws::mbar_expect_tx(&bars[WS_B_FULL + j], bbytes);
                ws::bulk_g2s(b_u32 + (uint32_t)j * bbytes, wsrc + (size_t)(j % n_chunks) * bbytes, bbytes, &bars[WS_B_FULL + j]);
            }
        }
        __syncwarp();
        const int lag = (SB >= 3) ? 2 : 1;
        int i = 0, wch = SB % n_chunks;                               // wch: W chunk of the next refill (chunk i-lag+SB)
        int bs = 0, bs_use = 0;                                       // slot of chunk i and how often it was used before
        int ab = 0, ab_use = 0;                                       // A buffer of chunk i, ditto
        for (int t = 0; t < (WS_DBG(2) ? 0 : my_tiles); ++t) {
            const int db = t & 1;
            if (t >= 2) ws::wait_or_trap(&bars[WS_D_EMPTY + db], ((t >> 1) - 1) & 1);
            umma::fence_after_sync();
            const uint32_t dcol = tmem_d + (uint32_t)(db * a.tmem_cols);
            for (int ch = 0; ch < n_chunks; ++ch, ++i) {
                if (!WS_DBG(32) || bs_use == 0) ws::wait_or_trap(&bars[WS_B_FULL + bs], bs_use & 1);
                ws::wait_or_trap(&bars[WS_A_FULL + ab], ab_use & 1);
                if (!WS_DBG(64)) umma::fence_proxy_async();         // A tile written with st.shared by the compute warps
                umma::fence_after_sync();
                const uint64_t dah = da0 + ((a_u32 + (uint32_t)ab * 2 * WS_A_HALF) >> 4);
                const uint64_t dal = dah + (WS_A_HALF >> 4);
                const uint64_t dbh = db0 + ((b_u32 + (uint32_t)bs * bbytes) >> 4);
                const uint64_t dbl = dbh + (bbytes >> 5);
                if (ws::elect_one()) {
                    if (!WS_DBG(8)) {
#pragma unroll
                        for (int ks = 0; ks < WS_CK / 8; ++ks) {
                            const uint32_t ao = (ks * 2 * WS_LBO_A) >> 4, bo = (ks * 2 * lbo_b) >> 4;
                            umma::mma_tf32_ss(dcol, dal + ao, dbh + bo, idesc, (ch > 0 || ks > 0) ? 1u : 0u);
                            umma::mma_tf32_ss(dcol, dah + ao, dbl + bo, idesc, 1u);
                            umma::mma_tf32_ss(dcol, dah + ao, dbh + bo, idesc, 1u);
                        }
                        umma::commit(&bars[WS_A_EMPTY + ab]);
                        umma::commit(&bars[WS_B_EMPTY + bs]);
                        if (ch == n_chunks - 1) umma::commit(&bars[WS_D_FULL + db]);
                    } else {                                          // ablation 8: no MMAs, plain arrives instead of commits
                        ws::mbar_arrive(&bars[WS_A_EMPTY + ab]);
                        ws::mbar_arrive(&bars[WS_B_EMPTY + bs]);
                        if (ch == n_chunks - 1) ws::mbar_arrive(&bars[WS_D_FULL + db]);
                    }
                }
                __syncwarp();
                // Refill the slot that chunk i-LAG used with chunk i-LAG+SB.  LAG = 2 when the ring has >= 3 slots: those
                // MMAs were committed two iterations ago, so the wait below does not drain the tensor pipe.
                if (i >= lag && i - lag + SB < total && !WS_DBG(32)) {
                    int ps = bs - lag, ps_use = bs_use;
                    if (ps < 0) { ps += SB; --ps_use; }
                    ws::wait_or_trap(&bars[WS_B_EMPTY + ps], ps_use & 1);
                    if (ws::elect_one()) {
                        ws::mbar_expect_tx(&bars[WS_B_FULL + ps], bbytes);
                        ws::bulk_g2s(b_u32 + (uint32_t)ps * bbytes, wsrc + (size_t)wch * bbytes, bbytes, &bars[WS_B_FULL + ps]);
                    }
                    __syncwarp();
                    if (++wch == n_chunks) wch = 0;
                }
                if (++bs == SB) { bs = 0; ++bs_use; }
                if (++ab == NA) { ab = 0; ++ab_use; }
            }
        }
    } else {
        // ================================================ compute warps ================================================
        // Ring stage = [8 points][8 neighbours][8 channels] (32 B per gathered row piece = one L2 sector; the two 16 B
        // halves are requested by adjacent lanes of the same cp.async so they land as one shared-memory wavefront --
        // with 16 B pieces of 32 different rows every piece was its own wavefront, ncu: 28 per LDGSTS).  A channel octet
        // therefore takes two stages (neighbour halves kh = 0, 1); its 8x4 accumulators live across both.
        const int pl = lane >> 2, jq = lane & 3;
        const int hh = jq & 1, ks = jq >> 1;                            // gather role: 16 B half of the piece, neighbour block
        const int row = warp * 8 + pl;                                  // row of the tile
        float *ring = ring_all + (size_t)warp * S * WS_STAGE;
        const uint32_t ring_u32 = umma::smem_u32(ring);
        const uint32_t my_dst = (uint32_t)(pl * WS_PSTRIDE + ks * 4 * 8 + hh * 4) * 4;   // this lane's first piece in a stage (bytes)
        const int NO = (C_cat + 7) >> 3;                                // channel octets per tile

        // ---- issue cursor: (tile, octet, neighbour half) of the next stage to request; the lane's 8 row offsets ----
        uint32_t q_off[8];                                              // rows k = 8 ks + kk, kk = 0..7
        uint32_t vmask = 0, add_off = 0;                                // bits 0..7 neighbour valid, bit 8 point valid
        auto load_rows = [&](int tile) {
            vmask = 0; add_off = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) q_off[i] = 0;
            const int m = tile * WS_PT + row;
            if (tile < n_tiles && m < n_out) {
                const longlong2 *src = reinterpret_cast<const longlong2 *>(a.nei + (size_t)m * K + ks * 8);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const longlong2 v = __ldg(src + i);
                    if (v.x >= 0 && v.x < n_in) { q_off[2 * i] = (uint32_t)v.x * (uint32_t)C_in; vmask |= 1u << (2 * i); }
                    if (v.y >= 0 && v.y < n_in) { q_off[2 * i + 1] = (uint32_t)v.y * (uint32_t)C_in; vmask |= 2u << (2 * i); }
                }
                vmask |= 256u;
                add_off = ((uint32_t)m * K + ks * 8) * (uint32_t)C_add;
            }
        };
        int i_tile = blockIdx.x, i_o = 0, i_kh = 0, i_slot = 0;
        load_rows(i_tile);
        auto issue_next = [&]() {
            if (i_tile < n_tiles && !WS_DBG(1)) {
                const int c = i_o * 8 + hh * 4;                         // this lane's 4 channels of the octet
                const uint32_t dst = ring_u32 + (uint32_t)i_slot * (WS_STAGE * 4) + my_dst;
                if (c < C_in) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int kk = i_kh * 4 + i;
                        ws::cp16(dst + i * 32, a.feats + (i_kh ? q_off[4 + i] : q_off[i]) + c, (vmask >> kk) & 1u);
                    }
                } else {
                    const bool ok = c < C_cat && (vmask & 256u);
                    const float *src = ok ? a.additional + add_off + (size_t)(i_kh * 4) * C_add + (c - C_in) : a.feats;
                    const int step = ok ? C_add : 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) ws::cp16(dst + i * 32, src + i * step, ok);
                }
            }
            if (i_tile < n_tiles) {
                if (++i_kh == 2) {
                    i_kh = 0;
                    if (++i_o == NO) { i_o = 0; i_tile += gridDim.x; load_rows(i_tile); }
                }
            }
            if (++i_slot == S) i_slot = 0;
            ws::commit_group();
        };
#pragma unroll 1
        for (int p = 0; p < S - 1; ++p) issue_next();

        int ab = 0, ab_use = 0;                                           // A buffer of the next chunk and its use count
        int c_slot = 0;
        int pend_tile = -1, pend_it = 0;                                  // epilogue owed by warps 0..WS_EPI-1
        auto epilogue = [&](int tile, int t_it) {
            const int db = t_it & 1;
            ws::wait_or_trap(&bars[WS_D_FULL + db], (uint32_t)(t_it >> 1) & 1);
            umma::fence_after_sync();
            const int r = warp * 32 + lane;
            const int m = (r < WS_PT) ? tile * WS_PT + r : n_out;
            const uint32_t taddr = tmem_d + (uint32_t)(db * a.tmem_cols) + ((uint32_t)(warp * 32) << 16);
            for (int o0 = 0; o0 < C_out; o0 += 16) {
                float v[16];
                ws::tmem_ld16(taddr + o0, v);
                if (m < n_out) {
                    float4 *dst = reinterpret_cast<float4 *>(a.out_y + (size_t)m * C_out + o0);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (a.lin_b) b = __ldg(reinterpret_cast<const float4 *>(a.lin_b + o0) + u);
                        dst[u] = make_float4(v[4 * u] + b.x, v[4 * u + 1] + b.y, v[4 * u + 2] + b.z, v[4 * u + 3] + b.w);
                    }
                }
            }
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) ws::mbar_arrive(&bars[WS_D_EMPTY + db]);
        };

        // this lane's weightnet values w[m][k][4jq .. 4jq+3], all 16 k, in registers for the tile
        float wreg[K][4];
        auto load_w = [&](int tile) {
            const int m = tile * WS_PT + row;
            const bool ok = tile < n_tiles && m < n_out;
            const float *wsrc = a.weights + (size_t)(ok ? m : 0) * K * 16 + jq * 4;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                // rows past the end read row 0 (finite values); their results are never stored.  No select on the loaded
                // values: the FSELs made the warp wait for all 16 loads right here (ncu: 5.7% of the samples) instead of
                // at the first FFMA2 of the next tile
                const float4 v = __ldg(reinterpret_cast<const float4 *>(wsrc + k * 16));
                wreg[k][0] = v.x; wreg[k][1] = v.y; wreg[k][2] = v.z; wreg[k][3] = v.w;
            }
        };
        load_w(blockIdx.x);

        int t_it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t_it) {
            const int m = tile * WS_PT + row;
            const bool pvalid = m < n_out;
            {   // pull the tile after next towards L2 (weightnet rows 8 x 1 KB per warp, neighbour table, guidance)
                const int mn = (tile + 2 * (int)gridDim.x) * WS_PT + warp * 8;
                if (mn + 8 <= n_out) {
                    const char *pn = reinterpret_cast<const char *>(a.weights + (size_t)mn * K * 16);
                    ws::prefetch_l2(pn + lane * 256);
                    ws::prefetch_l2(pn + lane * 256 + 128);
                    if (lane < 8) ws::prefetch_l2(reinterpret_cast<const char *>(a.nei + (size_t)mn * K) + lane * 128);
                    if (GQ > 0) {
                        const char *gn = reinterpret_cast<const char *>(a.guidance + (size_t)mn * K * H);
                        if (lane * 128 < 8 * K * H * 4) ws::prefetch_l2(gn + lane * 128);
                    }
                }
            }
            const float *gsrc = (GQ > 0) ? a.guidance + ((size_t)(pvalid ? m : 0) * K + ks * 8) * H : nullptr;

            for (int o = 0; o < NO; ++o) {
                float acc[8][4];
#pragma unroll
                for (int c = 0; c < 8; ++c)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;
#pragma unroll
                for (int kh = 0; kh < 2; ++kh) {
                    float4 gq[4];                                         // guidance of the lane's pieces (rows 8ks + 4kh + i)
                    const bool guided = GQ > 0 && (o * 8 + hh * 4) < C_in;
                    if (guided) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float *gp_ = gsrc + (kh * 4 + i) * H;
                            if (GQ == 2) {
                                gq[i] = __ldg(reinterpret_cast<const float4 *>(gp_) + hh);
                            } else if (H == 4) {
                                gq[i] = __ldg(reinterpret_cast<const float4 *>(gp_));
                            } else if (H == 2) {
                                const float2 v = __ldg(reinterpret_cast<const float2 *>(gp_));
                                gq[i] = make_float4(v.x, v.y, v.x, v.y);
                            } else {
                                const float v = __ldg(gp_);
                                gq[i] = make_float4(v, v, v, v);
                            }
                        }
                    }
                    ws::wait_group<S - 2>();
                    __syncwarp();
                    issue_next();
                    float *slot = ring + (size_t)c_slot * WS_STAGE;
                    if (++c_slot == S) c_slot = 0;
                    if (GQ > 0) {                                         // guidance multiply, each lane on the pieces it fetched
                        if (guided) {
                            unsigned char *rp = reinterpret_cast<unsigned char *>(slot) + my_dst;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float4 v = *reinterpret_cast<float4 *>(rp + i * 32);
                                v.x *= gq[i].x; v.y *= gq[i].y; v.z *= gq[i].z; v.w *= gq[i].w;
                                *reinterpret_cast<float4 *>(rp + i * 32) = v;
                            }
                        }
                        __syncwarp();
                    }
                    // ---- contraction 1 over this stage's 8 neighbours: acc[c][j] += G[k][c] w[k][j] ----
                    if (!WS_DBG(4)) {
                        const float4 *gp = reinterpret_cast<const float4 *>(slot + pl * WS_PSTRIDE);
#pragma unroll
                        for (int kl = 0; kl < 8; ++kl) {
                            const int k = (kl >> 2) * 8 + kh * 4 + (kl & 3);
                            const float4 v0 = gp[2 * kl], v1 = gp[2 * kl + 1];
                            const float g8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                ws::ffma2(acc[c][0], acc[c][1], g8[c], wreg[k][0], wreg[k][1]);
                                ws::ffma2(acc[c][2], acc[c][3], g8[c], wreg[k][2], wreg[k][3]);
                            }
                        }
                    }
                }
                if (o == NO - 1) load_w(tile + (int)gridDim.x);           // next tile's weightnet values fly during the hand-off
                // ---- four A chunks (2 channels each): split to (hi, lo), hand to the MMA warp ----
#pragma unroll
                for (int cp = 0; cp < 4; ++cp) {
                    if (o * 8 + 2 * cp >= C_cat) break;                   // octet padded past the last channel
                    if WS_DBG(2) {                              // ablation: keep the values alive, skip the hand-off
                        if (acc[2 * cp][0] + acc[2 * cp + 1][3] == 123.456f) a.out_y[0] = 1.f;
                        continue;
                    }
                    if (ab_use > 0 && !WS_DBG(16)) ws::wait_or_trap(&bars[WS_A_EMPTY + ab], (uint32_t)(ab_use - 1) & 1);
                    unsigned char *Ah = A_base + (size_t)ab * 2 * WS_A_HALF;
                    unsigned char *Al = Ah + WS_A_HALF;
#pragma unroll
                    for (int cl = 0; cl < (WS_DBG(128) ? 0 : 2); ++cl) {
                        const int c = 2 * cp + cl;
                        float4 hi, lo;
                        ws::split(acc[c][0], hi.x, lo.x); ws::split(acc[c][1], hi.y, lo.y);
                        ws::split(acc[c][2], hi.z, lo.z); ws::split(acc[c][3], hi.w, lo.w);
                        const uint32_t off = (uint32_t)(cl * 4 + jq) * WS_LBO_A + (uint32_t)row * 16;
                        *reinterpret_cast<float4 *>(Ah + off) = hi;
                        *reinterpret_cast<float4 *>(Al + off) = lo;
                        if (a.out_p && pvalid)
                            *reinterpret_cast<float4 *>(a.out_p + (size_t)m * KK + (size_t)(o * 8 + c) * 16 + jq * 4) =
                                make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                    }
                    // hand-off: the warp's stores are ordered before lane 0's release-arrive by __syncwarp; the generic->async
                    // proxy fence is executed once by the consumer (MMA warp) after its acquire.  A writer-side
                    // fence.proxy.async here compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC and cost 20% of the kernel (ncu).
                    __syncwarp();
                    if (lane == 0) ws::mbar_arrive(&bars[WS_A_FULL + ab]);
                    if (++ab == NA) { ab = 0; ++ab_use; }
                }
                if (o == 0 && pend_tile >= 0 && warp < WS_EPI && !WS_DBG(2)) { epilogue(pend_tile, pend_it); }
                if (o == 0) pend_tile = -1;
            }
            pend_tile = tile; pend_it = t_it;
        }
        if (pend_tile >= 0 && warp < WS_EPI && !WS_DBG(2)) epilogue(pend_tile, pend_it);
        ws::wait_group<0>();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == WS_NW)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(2 * a.tmem_cols) : "memory");
}

// ---- host side ---------------------------------------------------------------------------------
void prep_w_launch(const float *lin_w, int C_out, int KK, int CK, int n_chunks, float *out, cudaStream_t st);   // pconv_umma2.cu

static bool ws_config(const pcfb_pconv_shape *s, int *stages, int *sb, int *na) {
    // (ring stages, B slots, A buffers) in order of preference; the first that fits wins
    const int cand[][3] = {{4, 4, 4}, {3, 4, 4}, {3, 3, 4}, {3, 3, 3}, {3, 2, 3}, {3, 2, 2}};
    for (auto &c : cand)
        if (ws_plan(s->C_out, c[0], c[1], c[2]).total <= WS_SMEM_MAX) { *stages = c[0]; *sb = c[1]; *na = c[2]; return true; }
    return false;
}

bool pconv_forward_ws_supported(const pcfb_pconv_shape *s, bool has_lin) {
    if (!has_lin) return false;
    if (s->K != WS_K || s->C_mid != 16) return false;
    if (s->C_out < 16 || s->C_out > 128 || s->C_out % 16 != 0) return false;
    if (s->C_in % 4 != 0 || s->C_add % 4 != 0 || s->C_in + s->C_add < 4) return false;
    if (s->H != 0 && !((s->H == 1 || s->H == 2 || s->H == 4 || s->H == 8) && s->C_in % s->H == 0)) return false;
    if ((uint64_t)s->n_in * (uint64_t)s->C_in >= (1ull << 32)) return false;
    if ((uint64_t)s->n_out * WS_K * (uint64_t)(s->C_add > 0 ? s->C_add : 1) >= (1ull << 32)) return false;
    int st, sb, na;
    return ws_config(s, &st, &sb, &na);
}

size_t pconv_forward_ws_workspace(const pcfb_pconv_shape *s) {
    const int KK = (s->C_in + s->C_add) * s->C_mid;
    return align_up((size_t)(KK / WS_CK) * 2 * s->C_out * WS_CK * sizeof(float), 256);
}

template <int STAGES, int GQ>
static int launch_ws(const WsArgs &a, size_t smem, cudaStream_t st) {
    PCFB_CUDA(cudaFuncSetAttribute(pconv_fwd_ws_kernel<STAGES, GQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_MAX));
    const int grid = max(1, min(a.n_tiles, kNumSMs));
    launch_k(pconv_fwd_ws_kernel<STAGES, GQ>, grid, WS_NT, smem, st, a);
    return check_launch("pconv_fwd_ws_kernel");
}

int pconv_forward_ws(const pcfb_pconv_shape *s, const float *feats, const int64_t *nei, const float *weights,
                     const float *additional, const float *guidance, const float *lin_w, const float *lin_b,
                     float *out_y, float *out_p, void *workspace, size_t workspace_bytes, cudaStream_t st)
{
    PCFB_REQUIRE(pconv_forward_ws_supported(s, lin_w != nullptr), "pcfb_pconv_forward: shape unsupported by the warp-specialised variant");
    PCFB_REQUIRE(((uintptr_t)lin_w % 16 == 0) && ((uintptr_t)out_y % 16 == 0) && (!out_p || (uintptr_t)out_p % 16 == 0) &&
                 (!lin_b || (uintptr_t)lin_b % 16 == 0) && ((uintptr_t)weights % 16 == 0) && ((uintptr_t)nei % 16 == 0) &&
                 ((uintptr_t)feats % 16 == 0) && (s->C_add == 0 || (uintptr_t)additional % 16 == 0) &&
                 (s->H == 0 || (uintptr_t)guidance % 16 == 0),
                 "pcfb_pconv_forward: the tcgen05 variants need 16-byte aligned tensors");
    const size_t need = pconv_forward_ws_workspace(s);
    if (!workspace || workspace_bytes < need) { set_error("pcfb_pconv_forward: workspace %zu < %zu", workspace_bytes, need); return PCFB_ERR_WORKSPACE; }
    if (s->n_out == 0) return PCFB_OK;
    int stages = 0, sb = 0, na = 0;
    ws_config(s, &stages, &sb, &na);
    const WsPlan pl = ws_plan(s->C_out, stages, sb, na);
    const int C_cat = s->C_in + s->C_add, KK = C_cat * 16;
    WsArgs a{};
    a.s = *s;
    a.feats = feats; a.nei = nei; a.weights = weights; a.additional = additional; a.guidance = guidance;
    a.lin_b = lin_b; a.out_y = out_y; a.out_p = out_p;
    a.w_prep = static_cast<const float *>(workspace);
    a.n_groups = (C_cat + WS_CG - 1) / WS_CG;
    a.n_chunks = KK / WS_CK;
    a.n_tiles = ceil_div(s->n_out, WS_PT);
    a.sb = sb; a.na = na;
#if PCFB_WS_ABLATE
    { const char *e = getenv("PCFB_WS_DEBUG"); a.dbg = e ? atoi(e) : 0; }
#endif
    int cols = 32;
    while (cols < s->C_out) cols <<= 1;
    a.tmem_cols = cols;
    prep_w_launch(lin_w, s->C_out, KK, WS_CK, a.n_chunks, static_cast<float *>(workspace), st);
    int rc;
    if ((rc = check_launch("prep_w_kernel"))) return rc;
    const int gq = s->H == 0 ? 0 : (s->H == 8 ? 2 : 1);
#define WS_CASE(ST, GQ) if (stages == ST && gq == GQ) return launch_ws<ST, GQ>(a, pl.total, st);
    WS_CASE(4, 0) WS_CASE(4, 1) WS_CASE(4, 2)
    WS_CASE(3, 0) WS_CASE(3, 1) WS_CASE(3, 2)
#undef WS_CASE
    set_error("pcfb_pconv_forward: unreachable configuration");
    return PCFB_ERR_UNSUPPORTED;
}

}  // namespace pcfb
