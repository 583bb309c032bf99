// Brute-force exact kNN on packed scenes (sm_100a).
//
// Replaces the reference's third-party kNN (pykeops argKmin, /root/reference/knn_post_dataloader_utils.py:
// 22-41), the scene loop of compute_knn_packed (171-223) and prepare()'s offsetting (113-167).
//
// Design: one thread owns one query and keeps its K best (distance, index) pairs sorted in REGISTERS
// (fully unrolled compare-swap chain, K_PAD in {16,32,64}); the CTA streams the scene's reference
// points through shared memory as SoA tiles (float4 broadcast loads: 4 references per 3 LDS.128).
// References are visited in ascending index and accepted with a strict '<', so ties resolve to the
// lowest index -- bit-identical to a stable sort of the distances (the oracle).  The distance is
// evaluated with explicit round-to-nearest intrinsics so that ptxas cannot contract it into FMAs:
//     d = fadd(fadd(fmul(dx,dx), fmul(dy,dy)), fmul(dz,dz)),  dx = fsub(qx, rx) ...
// The kernel is fp32-issue bound (9 lane-ops per pair), not HBM bound: 12 B per point in, 8K B out.
#include "common.cuh"

namespace pcfb {

constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 1024;     // reference points per smem tile (12 KB)

__device__ __forceinline__ float sqdist(float qx, float qy, float qz, float rx, float ry, float rz) {
    const float dx = __fsub_rn(qx, rx), dy = __fsub_rn(qy, ry), dz = __fsub_rn(qz, rz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// which scene does packed index i belong to (off[s] <= i < off[s+1]); n_seg small
__device__ __forceinline__ int find_segment(const int32_t *__restrict__ off, int n_seg, int i) {
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <int KP>
struct TopK {
    float d[KP];
    int id[KP];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < KP; ++i) { d[i] = __int_as_float(0x7f800000); id[i] = -1; }
    }
    // precondition: dist < d[KP-1]
    __device__ __forceinline__ void insert(float dist, int idx) {
        d[KP - 1] = dist; id[KP - 1] = idx;
#pragma unroll
        for (int i = KP - 1; i > 0; --i) {
            const bool sw = d[i] < d[i - 1];          // strict: equal distance keeps the earlier index first
            const float dl = sw ? d[i] : d[i - 1], dh = sw ? d[i - 1] : d[i];
            const int il = sw ? id[i] : id[i - 1], ih = sw ? id[i - 1] : id[i];
            d[i - 1] = dl; d[i] = dh; id[i - 1] = il; id[i] = ih;
        }
    }
};

template <int KP>
__global__ void __launch_bounds__(KNN_THREADS)
knn_packed_kernel(const float *__restrict__ ref, const int32_t *__restrict__ ref_off,
                  const float *__restrict__ qry, const int32_t *__restrict__ qry_off,
                  int n_seg, int n_qry, int K, int64_t *__restrict__ out)
{
    pdl_wait();
    __shared__ __align__(16) float sx[KNN_TILE], sy[KNN_TILE], sz[KNN_TILE];
    __shared__ int s_lo, s_hi;

    const int q = blockIdx.x * KNN_THREADS + threadIdx.x;
    const bool active = q < n_qry;
    int my_lo = 0, my_hi = 0;           // this query's reference range [my_lo, my_hi)
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const int s = find_segment(qry_off, n_seg, q);
        my_lo = ref_off[s]; my_hi = ref_off[s + 1];
        qx = qry[3 * q]; qy = qry[3 * q + 1]; qz = qry[3 * q + 2];
    }
    // CTA-wide union of reference ranges (a CTA can straddle a scene boundary)
    if (threadIdx.x == 0) {
        const int q0 = blockIdx.x * KNN_THREADS;
        const int q1 = min(q0 + KNN_THREADS, n_qry) - 1;
        s_lo = ref_off[find_segment(qry_off, n_seg, q0)];
        s_hi = ref_off[find_segment(qry_off, n_seg, q1) + 1];
    }
    __syncthreads();
    const int lo = s_lo, hi = s_hi;

    TopK<KP> best;
    best.init();

    for (int base = lo; base < hi; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, hi - base);
        __syncthreads();
        // coalesced AoS read -> SoA smem; pad the tail with +inf coordinates (never accepted)
        for (int i = threadIdx.x; i < 3 * KNN_TILE; i += KNN_THREADS) {
            const int p = i / 3, c = i - 3 * p;
            const float v = (p < cnt) ? ref[3 * (size_t)base + i] : __int_as_float(0x7f800000);
            (c == 0 ? sx : (c == 1 ? sy : sz))[p] = v;
        }
        __syncthreads();
        if (!active) continue;
        // clip to this query's own scene
        const int j0 = max(my_lo - base, 0), j1 = min(my_hi - base, cnt);
        if (j0 >= j1) continue;
        const int j0a = j0 & ~3, j1a = (j1 + 3) & ~3;
        for (int j = j0a; j < j1a; j += 4) {
            const float4 rx = *reinterpret_cast<const float4 *>(&sx[j]);
            const float4 ry = *reinterpret_cast<const float4 *>(&sy[j]);
            const float4 rz = *reinterpret_cast<const float4 *>(&sz[j]);
            float d0 = sqdist(qx, qy, qz, rx.x, ry.x, rz.x);
            float d1 = sqdist(qx, qy, qz, rx.y, ry.y, rz.y);
            float d2 = sqdist(qx, qy, qz, rx.z, ry.z, rz.z);
            float d3 = sqdist(qx, qy, qz, rx.w, ry.w, rz.w);
            // out-of-scene lanes of a straddling tile: reject
            if (j + 0 < j0 || j + 0 >= j1) d0 = __int_as_float(0x7f800000);
            if (j + 1 < j0 || j + 1 >= j1) d1 = __int_as_float(0x7f800000);
            if (j + 2 < j0 || j + 2 >= j1) d2 = __int_as_float(0x7f800000);
            if (j + 3 < j0 || j + 3 >= j1) d3 = __int_as_float(0x7f800000);
            const float worst = best.d[KP - 1];
            if (d0 < worst || d1 < worst || d2 < worst || d3 < worst) {
                if (d0 < best.d[KP - 1]) best.insert(d0, base + j + 0);
                if (d1 < best.d[KP - 1]) best.insert(d1, base + j + 1);
                if (d2 < best.d[KP - 1]) best.insert(d2, base + j + 2);
                if (d3 < best.d[KP - 1]) best.insert(d3, base + j + 3);
            }
        }
    }
    if (!active) return;
    // n_ref(scene) < K: repeat the found neighbours cyclically (deterministic stand-in for the
    // reference's random fallback, knn_post_dataloader_utils.py:58-66)
    const int found = min(my_hi - my_lo, K);
    int64_t *o = out + (size_t)q * K;
#pragma unroll
    for (int i = 0; i < KP; ++i) {
        if (i < K) {
            int v = best.id[i];
            if (i >= found && found > 0) {
                const int src = i % found;
                v = -1;
#pragma unroll
                for (int t = 0; t < KP; ++t) if (t == src) v = best.id[t];
            }
            o[i] = (int64_t)v;
        }
    }
}

// Generic K (<= 255): lists live in shared memory, one column per thread (conflict-free).
constexpr int KNN_GEN_THREADS = 64;
__global__ void __launch_bounds__(KNN_GEN_THREADS)
knn_packed_generic_kernel(const float *__restrict__ ref, const int32_t *__restrict__ ref_off,
                          const float *__restrict__ qry, const int32_t *__restrict__ qry_off,
                          int n_seg, int n_qry, int K, int64_t *__restrict__ out)
{
    pdl_wait();
    extern __shared__ float smem[];
    float *ld = smem;                                              // [K][T]
    int *li = reinterpret_cast<int *>(smem + (size_t)K * KNN_GEN_THREADS);  // [K][T]
    const int t = threadIdx.x;
    const int q = blockIdx.x * KNN_GEN_THREADS + t;
    if (q >= n_qry) return;
    const int s = find_segment(qry_off, n_seg, q);
    const int lo = ref_off[s], hi = ref_off[s + 1];
    const float qx = qry[3 * q], qy = qry[3 * q + 1], qz = qry[3 * q + 2];
    for (int i = 0; i < K; ++i) { ld[i * KNN_GEN_THREADS + t] = __int_as_float(0x7f800000); li[i * KNN_GEN_THREADS + t] = -1; }
    float worst = __int_as_float(0x7f800000);
    for (int r = lo; r < hi; ++r) {
        const float d = sqdist(qx, qy, qz, __ldg(ref + 3 * (size_t)r), __ldg(ref + 3 * (size_t)r + 1), __ldg(ref + 3 * (size_t)r + 2));
        if (d < worst) {
            int pos = K - 1;
            while (pos > 0 && d < ld[(pos - 1) * KNN_GEN_THREADS + t]) {
                ld[pos * KNN_GEN_THREADS + t] = ld[(pos - 1) * KNN_GEN_THREADS + t];
                li[pos * KNN_GEN_THREADS + t] = li[(pos - 1) * KNN_GEN_THREADS + t];
                --pos;
            }
            ld[pos * KNN_GEN_THREADS + t] = d; li[pos * KNN_GEN_THREADS + t] = r;
            worst = ld[(K - 1) * KNN_GEN_THREADS + t];
        }
    }
    const int found = min(hi - lo, K);
    for (int i = 0; i < K; ++i) {
        const int src = (i < found || found == 0) ? i : (i % found);
        out[(size_t)q * K + i] = (int64_t)li[src * KNN_GEN_THREADS + t];
    }
}

}  // namespace pcfb

extern "C" int pcfb_knn_packed(const float *ref_xyz, const int32_t *ref_off, const float *qry_xyz,
                               const int32_t *qry_off, int n_seg, int n_ref, int n_qry, int K,
                               int64_t *out_idx, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(K >= 1 && K <= 255, "pcfb_knn_packed: K=%d outside [1,255]", K);
    PCFB_REQUIRE(n_seg >= 1 && n_ref >= 0 && n_qry >= 0, "pcfb_knn_packed: bad sizes");
    if (n_qry == 0) return PCFB_OK;
    PCFB_REQUIRE(ref_xyz && ref_off && qry_xyz && qry_off && out_idx, "pcfb_knn_packed: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ceil_div(n_qry, KNN_THREADS);
    if (K <= 16)
        launch_k(knn_packed_kernel<16>, grid, KNN_THREADS, 0, st, ref_xyz, ref_off, qry_xyz, qry_off, n_seg, n_qry, K, out_idx);
    else if (K <= 32)
        launch_k(knn_packed_kernel<32>, grid, KNN_THREADS, 0, st, ref_xyz, ref_off, qry_xyz, qry_off, n_seg, n_qry, K, out_idx);
    else if (K <= 64)
        launch_k(knn_packed_kernel<64>, grid, KNN_THREADS, 0, st, ref_xyz, ref_off, qry_xyz, qry_off, n_seg, n_qry, K, out_idx);
    else {
        const size_t smem = (size_t)K * KNN_GEN_THREADS * 8;
        static bool attr_set = false;
        if (!attr_set) {
            PCFB_CUDA(cudaFuncSetAttribute(knn_packed_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 255 * KNN_GEN_THREADS * 8));
            attr_set = true;
        }
        launch_k(knn_packed_generic_kernel, ceil_div(n_qry, KNN_GEN_THREADS), KNN_GEN_THREADS, smem, st, ref_xyz, ref_off, qry_xyz, qry_off, n_seg, n_qry, K, out_idx);
    }
    return check_launch("pcfb_knn_packed");
}
