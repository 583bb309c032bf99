// fp32-accurate GEMMs on the Blackwell tensor cores (sm_100a): 3xTF32 operand splitting on tcgen05.mma with
// TMEM accumulators.  These serve the dense per-point / per-edge Linear layers around the fused contraction
// (Linear_BN / UnaryBlock, /root/reference/layer_utils.py:241-319) and the two dense products
// of the fused backward (row a12: dP = dY W and dW = dY^T P, /root/reference/.../pconv_ops.cu:434-440,516-533).
// torch's fp32 path for the same products is a SIMT sgemm (cutlass_80_simt_sgemm); single-pass TF32 misses the
// 1e-4 parity bar, hence the split: x = hi + lo, D += A_lo B_hi + A_hi B_lo + A_hi B_hi.
//
//   gemm_nt : C[M x N] = A[M x K] * Bt (+ bias), A row-major (lda), Bt given as the weight matrix either
//             [N x K] (y = x W^T) or [K x N] (dx = dy W); the small B is split and laid out in the UMMA
//             K-major core-matrix order once by a prep kernel and streamed with cp.async.  Tile = 128 rows
//             (UMMA M = 128: TMEM lane == row), N <= 256 per launch column block.
//   gemm_tn : C[N1 x N2] = A[M x N1]^T * B[M x N2] (+ optional column of row sums = bias gradient), the
//             reduction runs over the M rows: split over CTAs, fp32 partials reduced in fixed order
//             (deterministic).  Operands are transposed in registers while staging (4x4 blocks), so both are
//             ordinary K-major descriptors.
#include "common.cuh"
#include <stdlib.h>
#include "umma.cuh"

namespace pcfb {

constexpr int GT_M = 128;      // rows per tile (gemm_nt)
constexpr int GT_KC = 32;      // K floats per chunk
constexpr int G_NT = 256;      // threads

__device__ __forceinline__ void g_cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(umma::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void g_cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ bool g_elect_one() {          // true in exactly one lane of the converged warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
template <int N>
__device__ __forceinline__ void g_cp_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }

// ---- B preparation: weight matrix -> (hi, lo) tf32, K-major core-matrix order, zero padded ----
// out[chunk][hi|lo][q = k/4 within chunk][n][4];  src element (n, k) = trans ? W[k*ldw + n] : W[n*ldw + k]
// Column blocks (blockIdx.y): block y covers weight columns n in [y*Nblk, min(N_total, (y+1)*Nblk)) and writes its own
// prepared image of block_floats floats.
__global__ void gemm_prep_b_kernel(const float *__restrict__ W, int ldw, int trans, int N_total, int Nblk, int Npad, int K, int n_chunks,
                                   size_t block_floats, float *__restrict__ out)
{
    pdl_wait();
    const int n_col0 = blockIdx.y * Nblk;
    const int N = min(Nblk, N_total - n_col0);
    W += trans ? (size_t)n_col0 : (size_t)n_col0 * ldw;
    out += (size_t)blockIdx.y * block_floats;
    const int units = GT_KC / 4;
    const int64_t total = (int64_t)n_chunks * units * Npad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % Npad);
        const int q = (int)((i / Npad) % units);
        const int ch = (int)(i / ((int64_t)Npad * units));
        const int k0 = ch * GT_KC + q * 4;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = k0 + e;
            v[e] = (n < N && k < K) ? (trans ? W[(size_t)k * ldw + n] : W[(size_t)n * ldw + k]) : 0.f;
        }
        float4 hi, lo;
        umma::split_tf32(v[0], hi.x, lo.x); umma::split_tf32(v[1], hi.y, lo.y);
        umma::split_tf32(v[2], hi.z, lo.z); umma::split_tf32(v[3], hi.w, lo.w);
        const size_t blk = (size_t)Npad * GT_KC;
        const size_t off = (size_t)q * Npad * 4 + (size_t)n * 4;
        *reinterpret_cast<float4 *>(out + ((size_t)ch * 2 + 0) * blk + off) = hi;
        *reinterpret_cast<float4 *>(out + ((size_t)ch * 2 + 1) * blk + off) = lo;
    }
}

struct GemmNtArgs {
    const float *A;          // [M][lda]
    const float *b_prep;     // prepared B
    const float *bias;       // [N] or null
    float *C;                // [M][ldc]
    int M, N, Npad, K, lda, ldc, n_chunks, tmem_cols;
    int Nblk;                // columns per column block (blockIdx.y); N is the total
    size_t b_block_floats;   // prepared-B floats per column block
    int b_resident;          // whole prepared B lives in shared memory
    int vec_a;               // 16-byte loads of A rows allowed
    int act;                 // 0 none, 1 relu, 2 leaky relu 0.1
};

struct GemmNtPlan { uint32_t a_bytes, b_bytes; int b_slots; size_t off_A, off_B, off_bar, total; };

__host__ __device__ inline GemmNtPlan gemm_nt_plan(int Npad, int n_chunks, bool resident) {
    GemmNtPlan pl;
    pl.a_bytes = GT_M * GT_KC * 4;                       // 16 KB per (hi | lo)
    pl.b_bytes = (uint32_t)Npad * GT_KC * 4;
    pl.b_slots = resident ? n_chunks : 3;
    size_t o = 0;
    pl.off_A = o; o += 4 * (size_t)pl.a_bytes;           // [2 bufs][hi|lo]
    pl.off_B = o; o += (size_t)pl.b_slots * 2 * pl.b_bytes;
    o = align_up(o, 16);
    pl.off_bar = o; o += 64;
    pl.total = o;
    return pl;
}

__global__ void __launch_bounds__(G_NT, 1) gemm_nt_kernel(GemmNtArgs a)
{
    pdl_wait();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const GemmNtPlan pl = gemm_nt_plan(a.Npad, a.n_chunks, a.b_resident != 0);
    unsigned char *A_base = smem_raw + pl.off_A;
    unsigned char *B_base = smem_raw + pl.off_B;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + pl.off_bar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + pl.off_bar + 32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_col0 = blockIdx.y * a.Nblk;                    // this CTA's column block
    const int M = a.M, N = min(a.Nblk, a.N - n_col0), Npad = a.Npad, K = a.K, n_chunks = a.n_chunks;
    a.b_prep += (size_t)blockIdx.y * a.b_block_floats;
    a.C += n_col0;
    if (a.bias) a.bias += n_col0;

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
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);          // provably warp-uniform copy for the MMA-issue branch
    const uint32_t idesc = umma::make_idesc_tf32(GT_M, Npad);
    const uint32_t lbo_a = GT_M * 16, lbo_b = (uint32_t)Npad * 16, sbo = 128;
    const uint64_t da0 = umma::make_smem_desc(0, lbo_a, sbo), db0 = umma::make_smem_desc(0, lbo_b, sbo);
    const uint32_t a_u32 = umma::smem_u32(A_base), b_u32 = umma::smem_u32(B_base);
    uint32_t uses0 = 0, uses1 = 0;

    auto issue_b = [&](int chunk, int slot) {
        unsigned char *dst = B_base + (size_t)slot * 2 * pl.b_bytes;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(a.b_prep) + (size_t)chunk * 2 * pl.b_bytes;
        for (uint32_t i = tid * 16; i < 2 * pl.b_bytes; i += G_NT * 16) g_cp_async16(dst + i, src + i);
    };
    if (a.b_resident) {
        for (int c = 0; c < n_chunks; ++c) issue_b(c, c);
        g_cp_commit();
        g_cp_wait<0>();
        umma::fence_proxy_async();
        __syncthreads();
    }

    // this thread's share of an A chunk: 4 float4 = (row, q) pairs, rows consecutive across threads
    float4 areg[4];
    auto load_a = [&](int m0, int chunk) {
        const int k0 = chunk * GT_KC;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + G_NT * i;
            const int row = idx & (GT_M - 1), q = idx >> 7;
            const int m = m0 + row, k = k0 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M && k < K) {
                const float *src = a.A + (size_t)m * a.lda + k;
                if (a.vec_a && k + 3 < K) {
                    v = *reinterpret_cast<const float4 *>(src);
                } else {
                    v.x = src[0];
                    if (k + 1 < K) v.y = src[1];
                    if (k + 2 < K) v.z = src[2];
                    if (k + 3 < K) v.w = src[3];
                }
            }
            areg[i] = v;
        }
    };
    auto store_a = [&](int buf) {
        unsigned char *Ah = A_base + (size_t)(buf * 2 + 0) * pl.a_bytes;
        unsigned char *Al = A_base + (size_t)(buf * 2 + 1) * pl.a_bytes;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + G_NT * i;
            const int row = idx & (GT_M - 1), q = idx >> 7;
            float4 hi, lo;
            umma::split_tf32(areg[i].x, hi.x, lo.x); umma::split_tf32(areg[i].y, hi.y, lo.y);
            umma::split_tf32(areg[i].z, hi.z, lo.z); umma::split_tf32(areg[i].w, hi.w, lo.w);
            *reinterpret_cast<float4 *>(Ah + (size_t)q * lbo_a + row * 16) = hi;
            *reinterpret_cast<float4 *>(Al + (size_t)q * lbo_a + row * 16) = lo;
        }
    };

    int it = 0;                                   // global chunk counter (A double buffering across tiles)
    for (int m0 = blockIdx.x * GT_M; m0 < M; m0 += gridDim.x * GT_M) {
        if (!a.b_resident) {
            issue_b(0, 0); g_cp_commit();
            if (n_chunks > 1) issue_b(1, 1);
            g_cp_commit();
        }
        load_a(m0, 0);
        for (int chunk = 0; chunk < n_chunks; ++chunk, ++it) {
            const int buf = it & 1;
            {   // A buffer free? (MMA that read it two chunks ago)
                const uint32_t u = buf ? uses1 : uses0;
                if (u > 0 && !umma::mbar_wait(&bars[buf], (u - 1) & 1)) __trap();
            }
            store_a(buf);
            if (chunk + 1 < n_chunks) load_a(m0, chunk + 1);        // next chunk's global loads in flight during the MMA
            if (!a.b_resident) g_cp_wait<1>();
            umma::fence_proxy_async();
            umma::fence_before_sync();
            __syncthreads();
            if (warp_u == 0) {                                       // warp-uniform branch: the MMAs issue back to back from one elected lane
                umma::fence_after_sync();
                const int slot = a.b_resident ? chunk : (chunk % 3);
                const uint64_t dah0 = da0 + ((a_u32 + (uint32_t)(buf * 2 + 0) * pl.a_bytes) >> 4);
                const uint64_t dal0 = da0 + ((a_u32 + (uint32_t)(buf * 2 + 1) * pl.a_bytes) >> 4);
                const uint64_t dbh0 = db0 + ((b_u32 + (uint32_t)(slot * 2 + 0) * pl.b_bytes) >> 4);
                const uint64_t dbl0 = db0 + ((b_u32 + (uint32_t)(slot * 2 + 1) * pl.b_bytes) >> 4);
                if (g_elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < GT_KC / 8; ++ks) {
                        const uint32_t ao = (ks * 2 * lbo_a) >> 4, bo = (ks * 2 * lbo_b) >> 4;
                        umma::mma_tf32_ss(tmem_d, dal0 + ao, dbh0 + bo, idesc, (chunk > 0 || ks > 0) ? 1u : 0u);
                        umma::mma_tf32_ss(tmem_d, dah0 + ao, dbl0 + bo, idesc, 1u);
                        umma::mma_tf32_ss(tmem_d, dah0 + ao, dbh0 + bo, idesc, 1u);
                    }
                    umma::commit(&bars[buf]);
                }
                __syncwarp();
            }
            if (buf) ++uses1; else ++uses0;
            if (!a.b_resident) {
                if (chunk + 2 < n_chunks) {
                    if (chunk >= 1) {                    // slot (chunk+2)%3 was read by MMA(chunk-1)
                        const int pb = (it - 1) & 1;
                        const uint32_t u = pb ? uses1 : uses0;
                        if (!umma::mbar_wait(&bars[pb], (u - 1) & 1)) __trap();
                    }
                    issue_b(chunk + 2, (chunk + 2) % 3);
                }
                g_cp_commit();
            }
        }
        // ---- epilogue: all 8 warps; warp w reads TMEM lanes 32*(w%4).., columns split between w<4 and w>=4 ----
        {
            const int lastbuf = (it - 1) & 1;
            const uint32_t u = lastbuf ? uses1 : uses0;
            if (!umma::mbar_wait(&bars[lastbuf], (u - 1) & 1)) __trap();
        }
        umma::fence_after_sync();
        {
            const int row = (warp & 3) * 32 + lane;
            const int m = m0 + row;
            const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
            const int half = Npad / 2 >= 8 ? ((Npad / 2 + 7) & ~7) : Npad;     // columns handled by warps 0-3
            const int c_begin = (warp < 4) ? 0 : half, c_end = (warp < 4) ? half : Npad;
            for (int c0 = c_begin; c0 < c_end; c0 += 8) {
                float v[8];
                umma::tmem_ld8(taddr + c0, v);
                if (m < M) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = c0 + j;
                        if (c < N) {
                            float x = v[j] + (a.bias ? __ldg(a.bias + c) : 0.f);
                            if (a.act == 1) x = fmaxf(x, 0.f);
                            else if (a.act == 2) x = x > 0.f ? x : 0.1f * x;
                            v[j] = x;
                        }
                    }
                    float *dst = a.C + (size_t)m * a.ldc + c0;
                    if (c0 + 8 <= N && ((a.ldc & 3) == 0)) {
                        reinterpret_cast<float4 *>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
                        reinterpret_cast<float4 *>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) if (c0 + j < N) dst[j] = v[j];
                    }
                }
            }
        }
        umma::fence_before_sync();
        if (!a.b_resident) g_cp_wait<0>();
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(a.tmem_cols) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// gemm_nt, pipelined variant for SMALL grids (every CTA owns one 128-row tile: the coarse levels of the pyramid).
//
// gemm_nt_kernel above walks the K chunks in lock step: wait for the chunk's global loads, split, st.shared, fence,
// __syncthreads, one lane issues the MMAs -- ~2 300 cycles per 32-wide chunk when a CTA is alone on its SM (ncu,
// profiles/ncu_gemm_small_r02.txt: 12 % of the samples wait for the A loads, 14 % at the block barrier, the rest is the
// split arithmetic with 8 warps and nothing to overlap it with).  With 148+ tiles that latency hides behind other CTAs'
// work; a 1 k-row product is 9 tiles and its 32 chunks are a 40 us chain on the critical path of the step.
// Here the chain is cut into independent actors that only meet at mbarriers:
//   * warps 0-15 (producers): A chunk -> registers (three chunks in flight per thread) -> hi / lo tf32 -> stage s of an
//     S-deep ring; one mbarrier arrive per warp on full[s]; nobody waits for the other warps;
//   * warp 16: waits full[s] and bfull[s], issues the chunk's 12 MMAs, tcgen05.commit -> empty[s];
//   * the prepared weights of a chunk are ONE contiguous block: a single cp.async.bulk per chunk (issued by thread 0 as
//     soon as empty[s] frees the stage) that completes on bfull[s].
// Stage = [A hi 16 KB][A lo 16 KB][B hi][B lo]; S = 2..4 stages, whatever fits.
// ------------------------------------------------------------------------------------------------------------
namespace gp {
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
__device__ __forceinline__ void wait_or_trap(uint64_t *bar, uint32_t parity) {
    if (!umma::mbar_wait(bar, parity)) __trap();
}
}  // namespace gp

// a thread's i-th (row, q) piece of an A chunk (q = 16-byte column group of the 32-wide chunk): a warp instruction covers
// 8 consecutive rows x 4 consecutive q -- 64 contiguous bytes per row from global memory, and every 8-lane phase of the
// st.shared.v4 into the canonical layout (q * 2048 + row * 16) hits 8 different 16-byte banks.  16 producer warps: the
// chunk is paced by the producers' instruction stream (split + addresses), so twice the warps per scheduler of the
// 8-warp version (profiles/ncu_gemmpipe_r02.txt: 28 % issue utilisation with two warps per scheduler).
constexpr int GP_PROD = 512;                 // producer threads
constexpr int GP_NP = GT_M * (GT_KC / 4) / GP_PROD;   // float4 pieces per thread and chunk (2)
__device__ __forceinline__ int gp_row(int tid, int i) { return (tid >> 5) * 8 + (tid & 7); }
__device__ __forceinline__ int gp_q(int tid, int i) { return ((tid >> 3) & 3) + 4 * i; }
constexpr int GP_THREADS = GP_PROD + 32;     // 16 producer warps + the MMA warp
constexpr int GP_MAX_STAGES = 4;

struct GemmPipePlan { uint32_t a_bytes, b_bytes, stage_bytes; int stages; size_t off_bar, total; };

__host__ __device__ inline GemmPipePlan gemm_pipe_plan(int Npad, int n_chunks) {
    GemmPipePlan pl;
    pl.a_bytes = GT_M * GT_KC * 4;
    pl.b_bytes = (uint32_t)Npad * GT_KC * 4;
    pl.stage_bytes = 2 * pl.a_bytes + 2 * pl.b_bytes;
    int S = (int)((220u * 1024u) / pl.stage_bytes);
    if (S > GP_MAX_STAGES) S = GP_MAX_STAGES;
    if (S > n_chunks) S = n_chunks;
    pl.stages = S;
    pl.off_bar = (size_t)(S > 0 ? S : 1) * pl.stage_bytes;
    pl.total = pl.off_bar + 128;
    return pl;
}

__global__ void __launch_bounds__(GP_THREADS, 1) gemm_nt_pipe_kernel(GemmNtArgs a)
{
    pdl_wait();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const GemmPipePlan pl = gemm_pipe_plan(a.Npad, a.n_chunks);
    const int S = pl.stages;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + pl.off_bar);      // full[S] | empty[S] | bfull[S] | done
    uint64_t *full = bars, *empty = bars + GP_MAX_STAGES, *bfull = bars + 2 * GP_MAX_STAGES, *done = bars + 3 * GP_MAX_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * GP_MAX_STAGES + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_col0 = blockIdx.y * a.Nblk;
    const int M = a.M, N = min(a.Nblk, a.N - n_col0), Npad = a.Npad, K = a.K, n_chunks = a.n_chunks;
    const int m0 = blockIdx.x * GT_M;
    const unsigned char *b_src = reinterpret_cast<const unsigned char *>(a.b_prep + (size_t)blockIdx.y * a.b_block_floats);
    a.C += n_col0;
    if (a.bias) a.bias += n_col0;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "r"(umma::smem_u32(tmem_slot)), "r"(a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        for (int s = 0; s < S; ++s) {
            umma::mbar_init(&full[s], GP_PROD / 32);
            umma::mbar_init(&empty[s], 1);
            umma::mbar_init(&bfull[s], 1);
        }
        umma::mbar_init(done, 1);
        umma::fence_mbar_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const uint32_t lbo_a = GT_M * 16, lbo_b = (uint32_t)Npad * 16, sbo = 128;
    const uint32_t stage_u32 = umma::smem_u32(smem_raw);

    if (warp_u == GP_PROD / 32) {
        // ==================== MMA issuer ====================
        const uint32_t idesc = umma::make_idesc_tf32(GT_M, Npad);
        const uint64_t da0 = umma::make_smem_desc(0, lbo_a, sbo), db0 = umma::make_smem_desc(0, lbo_b, sbo);
        int s = 0;
        uint32_t use = 0;
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            gp::wait_or_trap(&bfull[s], use & 1);
            gp::wait_or_trap(&full[s], use & 1);
            umma::fence_after_sync();
            const uint32_t base = stage_u32 + (uint32_t)s * pl.stage_bytes;
            const uint64_t dah0 = da0 + (base >> 4), dal0 = da0 + ((base + pl.a_bytes) >> 4);
            const uint64_t dbh0 = db0 + ((base + 2 * pl.a_bytes) >> 4), dbl0 = db0 + ((base + 2 * pl.a_bytes + pl.b_bytes) >> 4);
            if (g_elect_one()) {
#pragma unroll
                for (int ks = 0; ks < GT_KC / 8; ++ks) {
                    const uint32_t ao = (ks * 2 * lbo_a) >> 4, bo = (ks * 2 * lbo_b) >> 4;
                    umma::mma_tf32_ss(tmem_d, dal0 + ao, dbh0 + bo, idesc, (chunk > 0 || ks > 0) ? 1u : 0u);
                    umma::mma_tf32_ss(tmem_d, dah0 + ao, dbl0 + bo, idesc, 1u);
                    umma::mma_tf32_ss(tmem_d, dah0 + ao, dbh0 + bo, idesc, 1u);
                }
                umma::commit(&empty[s]);
                if (chunk == n_chunks - 1) umma::commit(done);
            }
            __syncwarp();
            if (++s == S) { s = 0; ++use; }
        }
    } else {
        // ==================== producers: A chunk -> hi / lo -> stage; epilogue ====================
        // this thread's GP_NP pieces: row pointers and validity once per tile, a pointer bump per chunk
        const float *src0[GP_NP];
        bool row_ok[GP_NP];
#pragma unroll
        for (int i = 0; i < GP_NP; ++i) {
            const int m = m0 + gp_row(tid, i);
            row_ok[i] = m < M;
            src0[i] = a.A + (size_t)(row_ok[i] ? m : 0) * a.lda + 4 * gp_q(tid, i);
        }
        const int n_fast = a.vec_a ? K / GT_KC : 0;             // chunks that need no column checks
        auto load_a = [&](float4 (&r)[GP_NP], int chunk) {
            const int k0 = chunk * GT_KC;
#pragma unroll
            for (int i = 0; i < GP_NP; ++i) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row_ok[i]) {
                    const float *src = src0[i] + k0;
                    if (chunk < n_fast) {
                        v = *reinterpret_cast<const float4 *>(src);
                    } else {
                        const int k = k0 + 4 * gp_q(tid, i);
                        if (k < K) v.x = src[0];
                        if (k + 1 < K) v.y = src[1];
                        if (k + 2 < K) v.z = src[2];
                        if (k + 3 < K) v.w = src[3];
                    }
                }
                r[i] = v;
            }
        };
        auto b_copy = [&](int chunk, int s) {                    // one thread: the chunk's prepared weights, one bulk copy
            gp::mbar_expect_tx(&bfull[s], 2 * pl.b_bytes);
            gp::bulk_g2s(stage_u32 + (uint32_t)s * pl.stage_bytes + 2 * pl.a_bytes, b_src + (size_t)chunk * 2 * pl.b_bytes, 2 * pl.b_bytes, &bfull[s]);
        };
        int s = 0;
        uint32_t use = 0;
        auto produce = [&](float4 (&r)[GP_NP], int chunk) {
            if (use > 0) {                                        // stage free? (the MMAs that read it S chunks ago)
                gp::wait_or_trap(&empty[s], (use - 1) & 1);
                if (tid == 0) b_copy(chunk, s);
            }
            unsigned char *Ah = smem_raw + (size_t)s * pl.stage_bytes, *Al = Ah + pl.a_bytes;
#pragma unroll
            for (int i = 0; i < GP_NP; ++i) {
                const int row = gp_row(tid, i), q = gp_q(tid, i);
                float4 hi, lo;
                umma::split_tf32(r[i].x, hi.x, lo.x); umma::split_tf32(r[i].y, hi.y, lo.y);
                umma::split_tf32(r[i].z, hi.z, lo.z); umma::split_tf32(r[i].w, hi.w, lo.w);
                *reinterpret_cast<float4 *>(Ah + (size_t)q * lbo_a + row * 16) = hi;
                *reinterpret_cast<float4 *>(Al + (size_t)q * lbo_a + row * 16) = lo;
            }
            if (chunk + 3 < n_chunks) load_a(r, chunk + 3);      // this register set is free again: three chunks in flight
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) gp::mbar_arrive(&full[s]);
            if (++s == S) { s = 0; ++use; }
        };
        if (tid == 0)
            for (int c = 0; c < S; ++c) b_copy(c, c);
        float4 r0[GP_NP], r1[GP_NP], r2[GP_NP];
        load_a(r0, 0);
        if (n_chunks > 1) load_a(r1, 1);
        if (n_chunks > 2) load_a(r2, 2);
        for (int chunk = 0; chunk < n_chunks; chunk += 3) {
            produce(r0, chunk);
            if (chunk + 1 < n_chunks) produce(r1, chunk + 1);
            if (chunk + 2 < n_chunks) produce(r2, chunk + 2);
        }
        // ---- epilogue: warp w reads TMEM lanes 32*(w%4).., the columns split over the four warps of a lane quarter ----
        gp::wait_or_trap(done, 0);
        umma::fence_after_sync();
        const int row = (warp & 3) * 32 + lane;
        const int m = m0 + row;
        const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
        const int cq = ((Npad + 3) / 4 + 7) & ~7;                // columns per warp of the quarter (multiple of 8)
        const int c_begin = min((warp >> 2) * cq, Npad), c_end = min(c_begin + cq, Npad);
        for (int c0 = c_begin; c0 < c_end; c0 += 8) {
            float v[8];
            umma::tmem_ld8(taddr + c0, v);
            if (m < M) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + j;
                    if (c < N) {
                        float x = v[j] + (a.bias ? __ldg(a.bias + c) : 0.f);
                        if (a.act == 1) x = fmaxf(x, 0.f);
                        else if (a.act == 2) x = x > 0.f ? x : 0.1f * x;
                        v[j] = x;
                    }
                }
                float *dst = a.C + (size_t)m * a.ldc + c0;
                if (c0 + 8 <= N && ((a.ldc & 3) == 0)) {
                    reinterpret_cast<float4 *>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
                    reinterpret_cast<float4 *>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (c0 + j < N) dst[j] = v[j];
                }
            }
        }
        umma::fence_before_sync();
    }
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(a.tmem_cols) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// gemm_tn: C[N1 x N2] (+ column N2 = row sums of A^T, i.e. sum_m A[m][n1]) = A^T B, reduction over rows m
// ------------------------------------------------------------------------------------------------------------

struct GemmTnArgs {
    const float *A, *B;            // A [M][lda] (N1 used columns), B [M][ldb] (N2 used columns)
    float *partial;                // [slices][N1][ldp], ldp = N2 + (ones ? 1 : 0)
    int M, N1, N1pad, N2, lda, ldb, ldp, ones, slice_rows, n2_tile, tmem_cols, RC;   // RC: rows per chunk
    int N1blk;                     // A columns per blockIdx.z block (N1pad pads one block)
};

constexpr uint32_t TN_SBO = 144;
struct GemmTnPlan { uint32_t a_bytes, b_bytes; size_t off_A, off_B, off_bar, total; };

__host__ __device__ inline GemmTnPlan gemm_tn_plan(int N1pad, int n2_tile, int RC) {
    GemmTnPlan pl;
    // 8-column groups (the UMMA core matrices, 128 B) sit TN_SBO = 144 B apart: with the dense 128 B stride the staging
    // stores of a warp (16 B every 64 B) were 4-way bank conflicts (ncu: 14.8 wavefronts per STS.128)
    pl.a_bytes = (uint32_t)(N1pad / 8) * TN_SBO * (RC / 4);
    pl.b_bytes = (uint32_t)(n2_tile / 8) * TN_SBO * (RC / 4);
    size_t o = 0;
    pl.off_A = o; o += 4 * (size_t)pl.a_bytes;
    pl.off_B = o; o += 4 * (size_t)pl.b_bytes;
    o = align_up(o, 16);
    pl.off_bar = o; o += 64;
    pl.total = o;
    return pl;
}

__global__ void __launch_bounds__(G_NT, 2) gemm_tn_kernel(GemmTnArgs a)
{
    pdl_wait();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const GemmTnPlan pl = gemm_tn_plan(a.N1pad, a.n2_tile, a.RC);
    const int RC = a.RC;
    unsigned char *A_base = smem_raw + pl.off_A;
    unsigned char *B_base = smem_raw + pl.off_B;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + pl.off_bar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_raw + pl.off_bar + 32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

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
    // operand buffers start zeroed: padding columns are never written again
    for (uint32_t i = tid * 16; i < 4 * pl.a_bytes + 4 * pl.b_bytes; i += G_NT * 16)
        *reinterpret_cast<float4 *>(smem_raw + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const bool vec_a = ((a.lda & 3) == 0) && ((uintptr_t)a.A % 16 == 0) && ((a.N1blk & 3) == 0);
    const bool vec_b = ((a.ldb & 3) == 0) && ((uintptr_t)a.B % 16 == 0) && ((a.n2_tile & 3) == 0);

    const int n1_0 = blockIdx.z * a.N1blk;                         // first A column (= output row) of this CTA
    const int n1_cols = min(a.N1blk, a.N1 - n1_0);
    const int n2_0 = blockIdx.x * a.n2_tile;                       // first B column of this CTA
    const int n2_cols = min(a.n2_tile, a.ldp - n2_0);              // incl. the ones column in the last tile
    const int slice = blockIdx.y;
    const int m_begin = slice * a.slice_rows, m_end = min(a.M, m_begin + a.slice_rows);
    const int n_mblk = a.N1pad / 128 > 0 ? a.N1pad / 128 : 1;      // MMA row blocks of 128 (or one of 64)
    const int mma_m = a.N1pad >= 128 ? 128 : 64;
    const uint32_t idesc = umma::make_idesc_tf32(mma_m, a.n2_tile);
    const uint32_t lbo_a = (uint32_t)(a.N1pad / 8) * TN_SBO, lbo_b = (uint32_t)(a.n2_tile / 8) * TN_SBO, sbo = TN_SBO;
    const uint64_t da0 = umma::make_smem_desc(0, lbo_a, sbo), db0 = umma::make_smem_desc(0, lbo_b, sbo);
    const uint32_t a_u32 = umma::smem_u32(A_base), b_u32 = umma::smem_u32(B_base);
    const int ones_col = a.ones ? a.N2 : -1;
    uint32_t uses0 = 0, uses1 = 0;

    // ---- staging as work items of 4 rows x 4 columns (transposed in registers).  Items [0, items_a) belong to A, the rest
    // to B.  The global loads of the NEXT chunk are issued into registers (TN_PRE items per thread) right after this
    // chunk's registers went to shared memory, so they fly while the MMAs of this chunk run and the buffer of the next one
    // frees up; the first version loaded and stored item by item and ran at 0.9 TB/s. ----
    const int b_cols = max(0, min(a.n2_tile, a.N2 - n2_0));
    const int cg_a = (n1_cols + 3) / 4;
    const int cg_b = (b_cols + ((ones_col >= n2_0 && ones_col < n2_0 + a.n2_tile) ? 1 : 0) + 3) / 4;
    const int rg = RC / 4;
    const int items_a = cg_a * rg, items = items_a + cg_b * rg;
    // Per-thread item descriptors are fixed for the whole kernel (the divisions happen once): source pointer of the
    // item's first element at row 0 of a chunk, row offset inside the chunk, shared-memory offset, and whether the four
    // float4 loads are legal (else the generic element-wise path).
    struct Item { const float *p; uint32_t off; int r, ld, flags; };          // flags: 1 = A operand, 2 = vector loads
    auto make_item = [&](int item) {
        Item t;
        const bool is_a = item < items_a;
        const int it = is_a ? item : item - items_a;
        const int cgroups = is_a ? cg_a : cg_b;
        const int c = (it % cgroups) * 4;
        t.r = (it / cgroups) * 4;
        t.ld = is_a ? a.lda : a.ldb;
        t.p = (is_a ? a.A + n1_0 : a.B + n2_0) + c;
        const bool full = c + 3 < (is_a ? n1_cols : b_cols);
        t.flags = (is_a ? 1 : 0) | (((is_a ? vec_a : vec_b) && full) ? 2 : 0);
        // (hi) byte offset of column c of row group r/4: K-unit * LBO + (c/8) * SBO + (c%8) * 16 ; lo = hi + operand bytes
        t.off = (uint32_t)(is_a ? pl.off_A : pl.off_B) + (uint32_t)(t.r / 4) * (is_a ? lbo_a : lbo_b) + (uint32_t)(c / 8) * TN_SBO + (uint32_t)(c % 8) * 16;
        return t;
    };
    auto load_item = [&](const Item &t, int item, int m0, float (&v)[4][4]) {
        if (t.flags & 2) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int m = m0 + t.r + rr;
                float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < m_end) t4 = __ldg(reinterpret_cast<const float4 *>(t.p + (size_t)m * t.ld));
                v[rr][0] = t4.x; v[rr][1] = t4.y; v[rr][2] = t4.z; v[rr][3] = t4.w;
            }
        } else {                                                       // ragged / unaligned columns, the ones column
            const bool is_a = (t.flags & 1) != 0;
            const int it = is_a ? item : item - items_a;
            const int c = (it % (is_a ? cg_a : cg_b)) * 4;
            const int ncols = is_a ? n1_cols : b_cols, col0 = is_a ? n1_0 : n2_0, oc = is_a ? -1 : ones_col;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int m = m0 + t.r + rr;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int col = c + cc;
                    float x = 0.f;
                    if (m < m_end) {
                        if (col < ncols) x = __ldg(t.p + (size_t)m * t.ld + cc);
                        else if (col0 + col == oc) x = 1.f;
                    }
                    v[rr][cc] = x;
                }
            }
        }
    };
    auto store_item = [&](const Item &t, const float (&v)[4][4], int buf) {
        const uint32_t opb = (t.flags & 1) ? pl.a_bytes : pl.b_bytes;
        unsigned char *dst_hi = smem_raw + t.off + (size_t)(buf * 2) * opb;
        unsigned char *dst_lo = dst_hi + opb;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            float4 hi, lo;
            umma::split_tf32(v[0][cc], hi.x, lo.x); umma::split_tf32(v[1][cc], hi.y, lo.y);
            umma::split_tf32(v[2][cc], hi.z, lo.z); umma::split_tf32(v[3][cc], hi.w, lo.w);
            *reinterpret_cast<float4 *>(dst_hi + cc * 16) = hi;     // c % 8 is 0 or 4: the four columns stay inside one group
            *reinterpret_cast<float4 *>(dst_lo + cc * 16) = lo;
        }
    };
    constexpr int TN_PRE = 3;
    float pre[TN_PRE][4][4];
    Item its[TN_PRE];
#pragma unroll
    for (int i = 0; i < TN_PRE; ++i) {
        its[i] = make_item(min(tid + i * G_NT, items - 1));
        if (tid + i * G_NT < items && m_begin < m_end) load_item(its[i], tid + i * G_NT, m_begin, pre[i]);
    }

    int chunk = 0;
    for (int m0 = m_begin; m0 < m_end; m0 += RC, ++chunk) {
        const int buf = chunk & 1;
        {
            const uint32_t u = buf ? uses1 : uses0;
            if (u > 0 && !umma::mbar_wait(&bars[buf], (u - 1) & 1)) __trap();
        }
#pragma unroll
        for (int i = 0; i < TN_PRE; ++i)
            if (tid + i * G_NT < items) store_item(its[i], pre[i], buf);
        for (int item = tid + TN_PRE * G_NT; item < items; item += G_NT) {      // tiles wider than the register window
            float v[4][4];
            const Item t = make_item(item);
            load_item(t, item, m0, v);
            store_item(t, v, buf);
        }
        if (m0 + RC < m_end) {
#pragma unroll
            for (int i = 0; i < TN_PRE; ++i)
                if (tid + i * G_NT < items) load_item(its[i], tid + i * G_NT, m0 + RC, pre[i]);
        }
        umma::fence_proxy_async();
        umma::fence_before_sync();
        __syncthreads();
        if (warp_u == 0) {                                           // warp-uniform branch, one elected lane issues
            umma::fence_after_sync();
            const uint64_t dah0 = da0 + ((a_u32 + (uint32_t)(buf * 2 + 0) * pl.a_bytes) >> 4);
            const uint64_t dal0 = da0 + ((a_u32 + (uint32_t)(buf * 2 + 1) * pl.a_bytes) >> 4);
            const uint64_t dbh0 = db0 + ((b_u32 + (uint32_t)(buf * 2 + 0) * pl.b_bytes) >> 4);
            const uint64_t dbl0 = db0 + ((b_u32 + (uint32_t)(buf * 2 + 1) * pl.b_bytes) >> 4);
            if (g_elect_one()) {
                for (int mb = 0; mb < n_mblk; ++mb) {
                    const uint32_t d = tmem_d + mb * a.n2_tile;
                    const uint32_t arow = (uint32_t)(mb * 16) * TN_SBO;   // 128 rows = 16 column groups further inside each K-unit
                    for (int ks = 0; ks < RC / 8; ++ks) {
                        const uint32_t ao = (ks * 2 * lbo_a + arow) >> 4, bo = (ks * 2 * lbo_b) >> 4;
                        umma::mma_tf32_ss(d, dal0 + ao, dbh0 + bo, idesc, (chunk > 0 || ks > 0) ? 1u : 0u);
                        umma::mma_tf32_ss(d, dah0 + ao, dbl0 + bo, idesc, 1u);
                        umma::mma_tf32_ss(d, dah0 + ao, dbh0 + bo, idesc, 1u);
                    }
                }
                umma::commit(&bars[buf]);
            }
            __syncwarp();
        }
        if (buf) ++uses1; else ++uses0;
    }
    // ---- epilogue: partial[slice][n1][n2_0 + c] ----
    if (chunk > 0) {
        const int lastbuf = (chunk - 1) & 1;
        const uint32_t u = lastbuf ? uses1 : uses0;
        if (!umma::mbar_wait(&bars[lastbuf], (u - 1) & 1)) __trap();
    }
    umma::fence_after_sync();
    if (warp < 4) {
        for (int mb = 0; mb < n_mblk; ++mb) {
            int n1;
            bool lane_ok = true;
            if (mma_m == 128) n1 = mb * 128 + warp * 32 + lane;
            else { n1 = warp * 16 + lane; lane_ok = lane < 16; }   // M = 64: row r <-> lane (r%16) + 32*(r/16)
            const uint32_t taddr = tmem_d + mb * a.n2_tile + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < a.n2_tile; c0 += 8) {
                float v[8];
                if (chunk > 0) umma::tmem_ld8(taddr + c0, v);
                else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = 0.f;
                }
                if (lane_ok && n1 < n1_cols) {
                    float *dst = a.partial + ((size_t)slice * a.N1 + n1_0 + n1) * a.ldp + n2_0 + c0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (c0 + j < n2_cols) dst[j] = v[j];
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_d), "r"(a.tmem_cols) : "memory");
}

__global__ void gemm_tn_reduce_kernel(const float *__restrict__ partial, int S, int N1, int N2, int ldp,
                                      float *__restrict__ C, int ldc, float *__restrict__ rowsum, int transposed)
{
    pdl_wait();
    const int total = N1 * ldp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int z = 0; z < S; ++z) s += partial[(size_t)z * total + i];          // fixed order: deterministic
        const int n1 = i / ldp, c = i - n1 * ldp;
        if (c < N2) { if (C) C[transposed ? (size_t)c * ldc + n1 : (size_t)n1 * ldc + c] = s; }
        else if (rowsum) rowsum[n1] = s;
    }
}

// column sums of a row-major [M x N] matrix (N <= 256), two deterministic stages: partial[slice][n], then out[n]
__global__ void gemm_colsum_kernel(const float *__restrict__ A, int lda, int M, int N, int slice_rows, float *__restrict__ partial)
{
    pdl_wait();
    __shared__ float red[256];
    const int n = threadIdx.x % N, sub = threadIdx.x / N, nsub = blockDim.x / N;
    const int m_begin = blockIdx.x * slice_rows, m_end = min(M, m_begin + slice_rows);
    float s = 0.f;
    if (sub < nsub)
        for (int m = m_begin + sub; m < m_end; m += nsub) s += __ldg(A + (size_t)m * lda + n);
    red[threadIdx.x] = (sub < nsub) ? s : 0.f;
    __syncthreads();
    if (threadIdx.x < N) {
        float t = 0.f;
        for (int q = 0; q < nsub; ++q) t += red[q * N + threadIdx.x];
        partial[(size_t)blockIdx.x * N + threadIdx.x] = t;
    }
}
__global__ void gemm_colsum_reduce_kernel(const float *__restrict__ partial, int S, int N, float *__restrict__ out)
{
    pdl_wait();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int z = 0; z < S; ++z) s += partial[(size_t)z * N + n];
    out[n] = s;
}

// ---- host side ---------------------------------------------------------------------------------------------
static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct NtSetup { int Nblk, n_blocks, Npad, n_chunks, resident; size_t block_bytes, prep_bytes; GemmNtPlan plan; };

// One column block when the tile (A double buffer + B ring or resident B) fits for the whole width, else column
// blocks of 128 (64 for very long K) handled by blockIdx.y of the same launch.  Few rows (M <= 0: unknown / many): the
// 184- and 1 k-point levels gave 2 .. 8 row tiles, i.e. 2 .. 8 CTAs each streaming the whole weight matrix (31 us for a
// 184 x 1536 x 192 product, profiles/step_kernels_by_grid_r01.txt); narrower column blocks spread the weight stream over
// ~NT_TARGET_CTAS CTAs instead.
constexpr int NT_TARGET_CTAS = 48;          // measured on the replayed step: 8 -> 20.56, 24 -> 20.39, 48 -> 20.26, 96 -> 20.26 ms (scripts/ab_step.sh)
static int nt_target_ctas() {                       // PCFB_NT_TARGET: A/B override of NT_TARGET_CTAS
    static int v = -1;
    if (v < 0) { const char *e = getenv("PCFB_NT_TARGET"); v = e ? atoi(e) : NT_TARGET_CTAS; if (v < 1) v = 1; }
    return v;
}
static bool nt_pipe_enabled() {                     // PCFB_GEMM_PIPE=0: A/B switch for the pipelined small-grid variant
    static int on = -1;
    if (on < 0) { const char *e = getenv("PCFB_GEMM_PIPE"); on = (e && e[0] == '0') ? 0 : 1; }
    return on != 0;
}
static NtSetup nt_setup(int N, int K, int M = 0) {
    NtSetup s;
    s.n_chunks = ceil_div(K, GT_KC);
    const int tiles = M > 0 ? ceil_div(M, GT_M) : (1 << 20);
    // N > 256: equal blocks of <= 128 columns rounded up to 16 (288 = 3 x 96: three blocks of 128 would leave the last one
    // three-quarters empty and still read A three times)
    const int even = N > 256 ? round_up(ceil_div(N, ceil_div(N, 128)), 16) : 128;
    const int widths[] = {N, even, 64, 32, 16};
    for (int wi = 0; wi < 5; ++wi) {
        const int nb = widths[wi];
        if (wi > 0 && nb >= N) continue;
        if (nb > 256) continue;
        if (wi < 4 && nb > 16 && tiles * ceil_div(N, nb) < nt_target_ctas() && N > 16) continue;   // too few CTAs: try narrower blocks
        s.Nblk = nb;
        s.n_blocks = ceil_div(N, nb);
        s.Npad = round_up(nb < 16 ? 16 : nb, 16);
        s.block_bytes = align_up((size_t)s.n_chunks * 2 * s.Npad * GT_KC * sizeof(float), 256);
        s.prep_bytes = s.block_bytes * s.n_blocks;
        s.resident = (s.block_bytes <= 128 * 1024) ? 1 : 0;
        s.plan = gemm_nt_plan(s.Npad, s.n_chunks, s.resident != 0);
        if (s.plan.total > 225 * 1024) { s.resident = 0; s.plan = gemm_nt_plan(s.Npad, s.n_chunks, false); }
        if (s.plan.total <= 225 * 1024) break;
    }
    return s;
}

}  // namespace pcfb

extern "C" size_t pcfb_gemm_nt_workspace(int N, int K)
{
    // any M: the widest prepared-B footprint over all column-block widths nt_setup may pick
    using namespace pcfb;
    size_t best = nt_setup(N, K).prep_bytes;
    const int n_chunks = ceil_div(K, GT_KC);
    const int even = N > 256 ? round_up(ceil_div(N, ceil_div(N, 128)), 16) : 128;
    const int widths[] = {N, even, 64, 32, 16};
    for (int wi = 0; wi < 5; ++wi) {
        const int nb = widths[wi];
        if ((wi > 0 && nb >= N) || nb > 256) continue;
        const int npad = round_up(nb < 16 ? 16 : nb, 16);
        const size_t bytes = align_up((size_t)n_chunks * 2 * npad * GT_KC * sizeof(float), 256) * (size_t)ceil_div(N, nb);
        if (bytes > best) best = bytes;
    }
    return best;
}

// phase bit 1: prepare B (weights -> hi / lo tf32 in UMMA order) into the workspace; bit 2: run the product from a prepared
// workspace.  The weights of a Linear are known long before its input is: pcfb_gemm_nt_prepare lets the host run the
// preparation on another stream, off the critical path of the step (~160 of them per training step).
static int gemm_nt_phases(int phases, const float *A, int lda, const float *W, int ldw, int w_is_kn, const float *bias, float *C, int ldc,
                          int M, int N, int K, int act, void *workspace, size_t workspace_bytes, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(M >= 0 && N >= 1 && K >= 1, "pcfb_gemm_nt: need N >= 1, K >= 1 (N=%d K=%d)", N, K);
    PCFB_REQUIRE(workspace && (!(phases & 1) || W) && (!(phases & 2) || (A && C)), "pcfb_gemm_nt: null pointer");
    PCFB_REQUIRE(!(phases & 2) || (lda >= K && ldc >= N), "pcfb_gemm_nt: bad leading dimensions");
    NtSetup s = nt_setup(N, K, M);
    PCFB_REQUIRE(s.plan.total <= 225 * 1024, "pcfb_gemm_nt: tile does not fit in shared memory (N=%d K=%d)", N, K);
    PCFB_REQUIRE(s.n_blocks <= 65535, "pcfb_gemm_nt: too many column blocks");
    if (workspace_bytes < s.prep_bytes) { set_error("pcfb_gemm_nt: workspace %zu < %zu", workspace_bytes, s.prep_bytes); return PCFB_ERR_WORKSPACE; }
    if (M == 0) return PCFB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if (phases & 1) {
        const int64_t total = (int64_t)s.n_chunks * (GT_KC / 4) * s.Npad;
        const int blocks = (int)((total + 255) / 256 < 592 ? (total + 255) / 256 : 592);
        launch_k(gemm_prep_b_kernel, dim3(blocks < 1 ? 1 : blocks, s.n_blocks), 256, 0, st, W, ldw, w_is_kn, N, s.Nblk, s.Npad, K, s.n_chunks,
                                                                              s.block_bytes / sizeof(float), static_cast<float *>(workspace));
        if ((rc = check_launch("gemm_prep_b_kernel"))) return rc;
    }
    if (!(phases & 2)) return PCFB_OK;
    GemmNtArgs a{};
    a.A = A; a.b_prep = static_cast<const float *>(workspace); a.bias = bias; a.C = C;
    a.M = M; a.N = N; a.Npad = s.Npad; a.K = K; a.lda = lda; a.ldc = ldc; a.n_chunks = s.n_chunks;
    a.Nblk = s.Nblk; a.b_block_floats = s.block_bytes / sizeof(float);
    a.b_resident = s.resident; a.act = act;
    a.vec_a = ((lda & 3) == 0) && ((uintptr_t)A % 16 == 0);
    int cols = 32;
    while (cols < s.Npad) cols <<= 1;
    a.tmem_cols = cols;
    static bool attr = false;
    if (!attr) { PCFB_CUDA(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024)); attr = true; }
    const int tiles = ceil_div(M, GT_M);
    if (nt_pipe_enabled() && (int64_t)tiles * s.n_blocks <= kNumSMs) {
        // every CTA owns one row tile and is alone on its SM: the pipelined variant (see gemm_nt_pipe_kernel)
        const GemmPipePlan pp = gemm_pipe_plan(s.Npad, s.n_chunks);
        if (pp.stages >= 2 || s.n_chunks == 1) {
            static bool attr_p = false;
            if (!attr_p) { PCFB_CUDA(cudaFuncSetAttribute(gemm_nt_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024)); attr_p = true; }
            launch_k(gemm_nt_pipe_kernel, dim3(tiles, s.n_blocks), GP_THREADS, pp.total, st, a);
            return check_launch("gemm_nt_pipe_kernel");
        }
    }
    const int per_sm = (s.plan.total + 1024 <= 113 * 1024) ? 2 : 1;
    int gx = ceil_div(kNumSMs * per_sm, s.n_blocks);               // CTAs per column block: the whole grid is about one wave
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    launch_k(gemm_nt_kernel, dim3(gx, s.n_blocks), G_NT, s.plan.total, st, a);
    return check_launch("gemm_nt_kernel");
}

extern "C" int pcfb_gemm_nt(const float *A, int lda, const float *W, int ldw, int w_is_kn, const float *bias, float *C, int ldc,
                            int M, int N, int K, int act, void *workspace, size_t workspace_bytes, void *stream)
{
    PCFB_REQUIRE(A && W && C, "pcfb_gemm_nt: null pointer");
    return gemm_nt_phases(3, A, lda, W, ldw, w_is_kn, bias, C, ldc, M, N, K, act, workspace, workspace_bytes, stream);
}

extern "C" int pcfb_gemm_nt_prepare(const float *W, int ldw, int w_is_kn, int M, int N, int K, void *workspace, size_t workspace_bytes,
                                    void *stream)
{
    return gemm_nt_phases(1, nullptr, 0, W, ldw, w_is_kn, nullptr, nullptr, 0, M, N, K, 0, workspace, workspace_bytes, stream);
}

extern "C" int pcfb_gemm_nt_prepared(const float *A, int lda, const float *bias, float *C, int ldc, int M, int N, int K, int act,
                                     const void *workspace, size_t workspace_bytes, void *stream)
{
    return gemm_nt_phases(2, A, lda, nullptr, 0, 0, bias, C, ldc, M, N, K, act, const_cast<void *>(workspace), workspace_bytes, stream);
}

namespace pcfb {
struct TnSetup { int swap, n1, n2, N1blk, n1_blocks, N1pad, n2_tile, n2_tiles, S, slice_rows, ldp, RC, ones; size_t ws_bytes, colsum_off; GemmTnPlan plan; };
// The UMMA M side (TMEM lanes) holds up to 256 output rows per CTA, the N side up to 256 columns.  An MMA costs
// max(M,128)*N/256 cycles, so the LONG output dimension belongs on the M side: dW = dY^T P (32 x 512) runs as
// (P^T dY)^T -- four 128-row blocks of N = 32 instead of one half-empty 64-row block of N = 256 twice (2x the tensor time).
static TnSetup tn_setup(int M, int N1, int N2, int want_rowsum) {
    TnSetup s;
    s.swap = 0;   // (N2 > N1 && N2 >= 128): measured slower on B200 at every model shape (the staging, not the tensor pipe, bounds this kernel)
    s.n1 = s.swap ? N2 : N1;                                      // rows of the product the kernel computes
    s.n2 = s.swap ? N1 : N2;
    s.ones = (want_rowsum && !s.swap) ? 1 : 0;                    // swapped: the bias gradient is a separate column sum
    s.N1blk = s.n1 <= 256 ? s.n1 : 256;
    s.n1_blocks = ceil_div(s.n1, s.N1blk);
    s.N1pad = s.N1blk <= 64 ? 64 : round_up(s.N1blk, 128);
    s.ldp = s.n2 + s.ones;
    const int max_tile = s.N1pad > 128 ? 128 : 256;               // TMEM columns: n_mblk * n2_tile <= 512; smem
    s.n2_tile = round_up(s.ldp < max_tile ? s.ldp : max_tile, 16);
    if (s.n2_tile < 16) s.n2_tile = 16;
    s.n2_tiles = ceil_div(s.ldp, s.n2_tile);
    s.RC = 128;                                                   // rows per chunk: as deep as shared memory allows
    while (s.RC > 32 && gemm_tn_plan(s.N1pad, s.n2_tile, s.RC).total > 200 * 1024) s.RC >>= 1;
    // two co-resident CTAs per SM hide each other's stage -> barrier -> MMA bubbles: take a shallower chunk if that is
    // what it costs to get under half of the shared memory
    int per_sm = 1;
    if (gemm_tn_plan(s.N1pad, s.n2_tile, s.RC).total > 110 * 1024 && s.RC >= 32 &&
        gemm_tn_plan(s.N1pad, s.n2_tile, s.RC / 2).total <= 110 * 1024 && M >= 16384) { s.RC >>= 1; per_sm = 2; }
    else if (gemm_tn_plan(s.N1pad, s.n2_tile, s.RC).total <= 110 * 1024) per_sm = 2;
    int S = ceil_div(2 * per_sm * kNumSMs, s.n2_tiles * s.n1_blocks);
    const int maxS = ceil_div(M > 0 ? M : 1, 4 * s.RC);
    if (S > maxS) S = maxS;
    if (S < 1) S = 1;
    s.slice_rows = round_up(ceil_div(M > 0 ? M : 1, S), s.RC);
    s.S = ceil_div(M > 0 ? M : 1, s.slice_rows);
    s.colsum_off = align_up((size_t)s.S * s.n1 * s.ldp * sizeof(float), 256);
    s.ws_bytes = s.colsum_off + ((want_rowsum && s.swap) ? align_up((size_t)s.S * N1 * sizeof(float), 256) : 0);
    s.plan = gemm_tn_plan(s.N1pad, s.n2_tile, s.RC);
    return s;
}
}  // namespace pcfb

extern "C" size_t pcfb_gemm_tn_workspace(int M, int N1, int N2, int with_rowsum)
{
    return pcfb::tn_setup(M, N1, N2, with_rowsum).ws_bytes;
}

extern "C" int pcfb_gemm_tn(const float *A, int lda, const float *B, int ldb, float *C, int ldc, float *rowsum,
                            int M, int N1, int N2, void *workspace, size_t workspace_bytes, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(M >= 0 && N1 >= 1 && N1 <= 256 && N2 >= 1, "pcfb_gemm_tn: need 1 <= N1 <= 256, N2 >= 1 (N1=%d N2=%d)", N1, N2);
    PCFB_REQUIRE(A && B && workspace && (C || rowsum), "pcfb_gemm_tn: null pointer");
    PCFB_REQUIRE(lda >= N1 && ldb >= N2, "pcfb_gemm_tn: bad leading dimensions");
    TnSetup s = tn_setup(M, N1, N2, rowsum != nullptr);
    PCFB_REQUIRE(s.plan.total <= 225 * 1024, "pcfb_gemm_tn: tile does not fit in shared memory");
    if (workspace_bytes < s.ws_bytes) { set_error("pcfb_gemm_tn: workspace %zu < %zu", workspace_bytes, s.ws_bytes); return PCFB_ERR_WORKSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GemmTnArgs a{};
    a.A = s.swap ? B : A; a.B = s.swap ? A : B; a.partial = static_cast<float *>(workspace);
    a.M = M; a.N1 = s.n1; a.N1pad = s.N1pad; a.N1blk = s.N1blk; a.N2 = s.n2;
    a.lda = s.swap ? ldb : lda; a.ldb = s.swap ? lda : ldb; a.ldp = s.ldp; a.ones = s.ones;
    a.slice_rows = s.slice_rows; a.n2_tile = s.n2_tile; a.RC = s.RC;
    int cols = 32;
    const int need_cols = (s.N1pad >= 128 ? s.N1pad / 128 : 1) * s.n2_tile;
    while (cols < need_cols) cols <<= 1;
    a.tmem_cols = cols;
    static bool attr = false;
    if (!attr) { PCFB_CUDA(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024)); attr = true; }
    int rc;
    dim3 grid(s.n2_tiles, s.S, s.n1_blocks);
    launch_k(gemm_tn_kernel, grid, G_NT, s.plan.total, st, a);
    if ((rc = check_launch("gemm_tn_kernel"))) return rc;
    const int total = s.n1 * s.ldp;
    launch_k(gemm_tn_reduce_kernel, min(ceil_div(total, 256), kNumSMs * 4), 256, 0, st, a.partial, s.S, s.n1, s.n2, s.ldp, C, ldc,
                                                                            s.swap ? nullptr : rowsum, s.swap);
    if ((rc = check_launch("gemm_tn_reduce_kernel"))) return rc;
    if (rowsum && s.swap) {
        float *cpart = reinterpret_cast<float *>(static_cast<char *>(workspace) + s.colsum_off);
        launch_k(gemm_colsum_kernel, s.S, 256, 0, st, A, lda, M, N1, s.slice_rows, cpart);
        if ((rc = check_launch("gemm_colsum_kernel"))) return rc;
        launch_k(gemm_colsum_reduce_kernel, ceil_div(N1, 128), 128, 0, st, cpart, s.S, N1, rowsum);
        if ((rc = check_launch("gemm_colsum_reduce_kernel"))) return rc;
    }
    return PCFB_OK;
}
