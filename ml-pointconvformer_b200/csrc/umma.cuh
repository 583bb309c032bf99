// Thin inline-PTX layer over the Blackwell (sm_100a) tensor-core path: tcgen05.mma (kind::tf32) with
// TMEM accumulators, shared-memory matrix descriptors, mbarrier completion.  Bit layouts follow the
// CUTLASS headers vendored in this image (cute/arch/mma_sm100_desc.hpp: SmemDescriptor,
// InstrDescriptor) -- re-derived here, nothing is included from CUTLASS.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcfb {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- shared-memory matrix descriptor (K-major, no swizzle: "interleaved" 8x16B core matrices) ----
// element (row r, k) of a tf32 operand lives at byte
//     (k/4)*LBO + (r/8)*SBO + (r%8)*16 + (k%4)*4
// bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// ---- instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major ----
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4)                       // c_format = F32
         | (2u << 7) | (2u << 10)          // a_format = b_format = TF32
         | ((uint32_t)(N >> 3) << 17)      // n_dim
         | ((uint32_t)(M >> 4) << 24);     // m_dim
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// all previously issued tcgen05.mma of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                 :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- TMEM ----
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_in_smem) {      // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                 :: "r"(smem_u32(slot_in_smem)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {             // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(taddr), "n"(NCOLS) : "memory");
}

// 32 lanes x 8 consecutive columns: thread i of the warp gets TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: returns false on timeout (so a broken pipeline faults loudly instead of hanging the GPU)
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 24); ++spin)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}

// ---- 3xTF32 operand split: x ~ hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi) ----
// The tensor core reads only the upper 19 bits of an operand (it truncates): left as the exact fp32 remainder, lo would
// lose up to 2^-10 of itself = 2^-21 |x|; rounded to nearest here the split carries 22 mantissa bits (unit round-off
// 2^-22 |x|, fp32 itself: 2^-24) -- measured on the full-width model golden: logits error vs float64 1.5e-4 -> see DESIGN.md 4.
// cvt.rna.tf32.f32 (round to nearest, ties away) is emulated on sm_100a as add-half-ulp, an Inf / NaN test, a mask and a
// select (4 instructions); for finite values the test and the select are dead weight -- half an ulp added to the bit
// pattern and the low 13 bits cleared is the same number (NaN stays NaN, Inf stays Inf): 5 instructions per split value
// instead of 9, in kernels whose producers are bound by exactly this arithmetic (profiles/ncu_gemm_small_r02.txt).
__device__ __forceinline__ float rna_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = rna_tf32(x);
    lo = rna_tf32(x - hi);
}

}  // namespace umma
}  // namespace pcfb
