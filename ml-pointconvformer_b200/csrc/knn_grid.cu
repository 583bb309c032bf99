// Exact kNN on packed scenes through a uniform grid (sm_100a) -- same results, bit for bit, as the
// brute-force kernel in knn.cu (and therefore as the oracle), at O(N * candidates) instead of O(N^2).
//
// "a grid-hash-accelerated exact kNN (same results)".  Distances are evaluated with
// exactly the same fp32 arithmetic (no FMA contraction) and the K best are kept in (distance, index)
// lexicographic order, so the result does not depend on the order candidates are visited in.
//
// build : per scene bounding box -> cell edge h (caller hint, enlarged on the device until the dense grid
//         fits the caller's cell budget; no host round trip) -> dense cell id per reference + histogram -> exclusive scan
//         (cell -> first slot) -> slot-claim fill of the cell-sorted index list and float4 copy.  The order of the
//         references INSIDE a cell is whatever the atomics give: the query result does not depend on it.
// query : one thread per query walks Chebyshev shells R = 0,1,2,... of cells around its own cell.  The references are
//         kept as a CELL-SORTED float4 copy (x, y, z, index bits), and cells that are neighbours along x are neighbours
//         in that copy, so a row of the shell is ONE contiguous range: one independent 16-byte load per candidate
//         instead of the cell -> index -> coordinates chain of dependent loads.  Self queries (query cloud == reference
//         cloud) are processed in cell-sorted order: the lanes of a warp then sit in the same or adjacent cells, walk
//         the same ranges (no divergence in the trip counts, every load a broadcast) -- measured in profiles/README.md.
//         Every reference NOT yet visited after shell R differs from the query by more than R*h along some axis
//         (up to the fp32 rounding of the cell index, bounded explicitly in the kernel), so once the current
//         K-th best squared distance is below ((R - margin)*h)^2 no unvisited reference can enter: exact.
#include "common.cuh"
#include "scan.cuh"

namespace pcfb {

struct GridPlan {          // one per scene, lives in the workspace
    float ox, oy, oz, h;
    int nx, ny, nz, cell_off;
    int ref_lo, ref_hi;    // reference range of the scene
};

__device__ __forceinline__ int g_float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float g_ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ int g_find_seg(const int32_t *__restrict__ off, int n_seg, int i) {
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void kg_init_kernel(int *__restrict__ mm, int n_seg) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_seg * 6) mm[i] = (i % 6 < 3) ? 0x7fffffff : (int)0x80000000;
}

__global__ void kg_bounds_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ off, int n_seg, int n,
                                 int *__restrict__ mm)
{
    pdl_wait();
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {         // warp-uniform trip count
        const int i = base + threadIdx.x;
        const bool in = i < n;
        const size_t j = in ? (size_t)i : 0;
        warp_scene_minmax(in, in ? g_find_seg(off, n_seg, i) : 0, g_float_to_ordered(xyz[3 * j]), g_float_to_ordered(xyz[3 * j + 1]),
                          g_float_to_ordered(xyz[3 * j + 2]), mm);
    }
}

// single thread: per-scene grid geometry; h grows by 2^(1/3) until the scene's dense grid fits its budget
__global__ void kg_plan_kernel(const int *__restrict__ mm, const int32_t *__restrict__ off, int n_seg, float cell_hint,
                               GridPlan *__restrict__ plans)
{
    pdl_wait();
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int cell_off = 0;
    for (int s = 0; s < n_seg; ++s) {
        GridPlan p;
        p.ref_lo = off[s]; p.ref_hi = off[s + 1];
        const int n = p.ref_hi - p.ref_lo;
        const long long budget = 2ll * n + 64;
        if (n <= 0) {
            p.ox = p.oy = p.oz = 0.f; p.h = 1.f; p.nx = p.ny = p.nz = 1;
        } else {
            const float mnx = g_ordered_to_float(mm[s * 6 + 0]), mny = g_ordered_to_float(mm[s * 6 + 1]), mnz = g_ordered_to_float(mm[s * 6 + 2]);
            const float ex = g_ordered_to_float(mm[s * 6 + 3]) - mnx, ey = g_ordered_to_float(mm[s * 6 + 4]) - mny, ez = g_ordered_to_float(mm[s * 6 + 5]) - mnz;
            float h = cell_hint;
            if (!(h > 0.f)) h = fmaxf(fmaxf(ex, ey), ez) * (1.0f / 1024.0f);
            if (!(h > 1e-20f)) h = 1.f;                 // degenerate cloud (all points equal)
            while (true) {
                const long long nx = (long long)floorf(ex / h) + 1, ny = (long long)floorf(ey / h) + 1, nz = (long long)floorf(ez / h) + 1;
                if (nx * ny * nz <= budget) { p.nx = (int)nx; p.ny = (int)ny; p.nz = (int)nz; break; }
                h *= 1.2599211f;
            }
            p.ox = mnx; p.oy = mny; p.oz = mnz; p.h = h;
        }
        p.cell_off = cell_off;
        cell_off += p.nx * p.ny * p.nz;
        plans[s] = p;
    }
}

__device__ __forceinline__ int cell_coord(float x, float o, float h, int n) {
    const int c = (int)floorf(__fdiv_rn(__fsub_rn(x, o), h));
    return min(max(c, 0), n - 1);
}

__global__ void kg_cell_count_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ off, int n_seg, int n,
                                     const GridPlan *__restrict__ plans, int32_t *__restrict__ cell,
                                     int32_t *__restrict__ counts)
{
    pdl_wait();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const GridPlan p = plans[g_find_seg(off, n_seg, i)];
        const int cx = cell_coord(xyz[3 * (size_t)i], p.ox, p.h, p.nx);
        const int cy = cell_coord(xyz[3 * (size_t)i + 1], p.oy, p.h, p.ny);
        const int cz = cell_coord(xyz[3 * (size_t)i + 2], p.oz, p.h, p.nz);
        const int c = p.cell_off + cx + p.nx * (cy + p.ny * cz);
        cell[i] = c;
        atomicAdd(&counts[c], 1);
    }
}

// slot claim: reference i takes one of its cell's slots (counts[] still holds the histogram and is counted down);
// writes the cell-sorted index list and the cell-sorted float4 copy (x, y, z, index bits)
__global__ void kg_fill_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ cell, int n,
                               const int32_t *__restrict__ cell_ptr, int32_t *__restrict__ counts,
                               int32_t *__restrict__ cell_pts, float4 *__restrict__ sorted)
{
    pdl_wait();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = cell[i];
        const int pos = cell_ptr[c] + atomicSub(&counts[c], 1) - 1;
        cell_pts[pos] = i;
        sorted[pos] = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], __int_as_float(i));
    }
}

// K best (distance, index) pairs in ascending lexicographic order, in registers.  Distances are kept as their IEEE bit
// patterns (a sum of squares is >= +0, so unsigned integer order == float order): (distance, index) then compares as ONE
// 64-bit integer -- two ISETP instead of three float / integer compares plus predicate logic.
template <int KP>
struct TopKLex {
    unsigned d[KP];
    int id[KP];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < KP; ++i) { d[i] = 0x7f800000u; id[i] = 0x7fffffff; }
    }
    static __device__ __forceinline__ unsigned long long key(unsigned dist, int idx) {
        return ((unsigned long long)dist << 32) | (unsigned)idx;
    }
    __device__ __forceinline__ bool accepts(unsigned dist, int idx) const { return key(dist, idx) < key(d[KP - 1], id[KP - 1]); }
    // sorted insertion, branch free, INDEPENDENT comparisons (no bubble chain): position i takes its left neighbour if the
    // new element sorts before that neighbour, the new element if it sorts before the old occupant, else keeps the occupant
    // (2 ISETP + 4 SEL per position; the first version of this, written with nested ?:, compiled to 292 instructions of
    // branches per insertion -- profiles/ncu_knn_r02.txt)
    __device__ __forceinline__ void insert(unsigned dist, int idx) {
        const unsigned long long nk = key(dist, idx);
        bool lt_i = true;                                          // accepts() held: the new element sorts before d[KP-1]
#pragma unroll
        for (int i = KP - 1; i > 0; --i) {
            const bool lt_l = nk < key(d[i - 1], id[i - 1]);
            const unsigned td = lt_i ? dist : d[i];
            const int ti = lt_i ? idx : id[i];
            d[i] = lt_l ? d[i - 1] : td;
            id[i] = lt_l ? id[i - 1] : ti;
            lt_i = lt_l;
        }
        d[0] = lt_i ? dist : d[0];
        id[0] = lt_i ? idx : id[0];
    }
    __device__ __forceinline__ float kth(int K) const {         // d[K-1] without dynamic register indexing
        unsigned r = d[KP - 1];
#pragma unroll
        for (int i = 0; i < KP; ++i) if (i == K - 1) r = d[i];
        return __uint_as_float(r);
    }
};

constexpr int KG_THREADS = 128;
// (dz, dy) of the nine rows of the 3 x 3 x 3 cube, nearest first, as 2-bit fields (value + 1)
constexpr unsigned kg_pack9(int a0, int a1, int a2, int a3, int a4, int a5, int a6, int a7, int a8) {
    return (unsigned)(a0 + 1) | (unsigned)(a1 + 1) << 2 | (unsigned)(a2 + 1) << 4 | (unsigned)(a3 + 1) << 6 | (unsigned)(a4 + 1) << 8 |
           (unsigned)(a5 + 1) << 10 | (unsigned)(a6 + 1) << 12 | (unsigned)(a7 + 1) << 14 | (unsigned)(a8 + 1) << 16;
}
constexpr unsigned KG_CUBE_DZ = kg_pack9(0, 0, 0, -1, 1, -1, -1, 1, 1);
constexpr unsigned KG_CUBE_DY = kg_pack9(0, -1, 1, 0, 0, -1, 1, -1, 1);

template <int KP>
__global__ void __launch_bounds__(KG_THREADS)
knn_grid_query_kernel(const float4 *__restrict__ sorted, const GridPlan *__restrict__ plans,
                      const int32_t *__restrict__ cell_ptr, const int32_t *__restrict__ order,
                      const float *__restrict__ qry, const int32_t *__restrict__ qry_off, int n_seg, int n_qry, int K,
                      int64_t *__restrict__ out)
{
    pdl_wait();
    __shared__ int2 row_s[9][KG_THREADS];                         // the nine row ranges of every thread's cube
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_qry) return;
    const int q = order ? order[i] : i;
    const GridPlan p = plans[g_find_seg(qry_off, n_seg, q)];
    const float qx = qry[3 * (size_t)q], qy = qry[3 * (size_t)q + 1], qz = qry[3 * (size_t)q + 2];
    TopKLex<KP> best;
    best.init();
    const int n_ref = p.ref_hi - p.ref_lo;
    if (n_ref > 0) {
        const int cx = cell_coord(qx, p.ox, p.h, p.nx), cy = cell_coord(qy, p.oy, p.h, p.ny), cz = cell_coord(qz, p.oz, p.h, p.nz);
        auto candidate = [&](const float4 rr) {
            const float dx = __fsub_rn(qx, rr.x), dy2 = __fsub_rn(qy, rr.y), dz2 = __fsub_rn(qz, rr.z);
            const float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy2, dy2)), __fmul_rn(dz2, dz2));
            const int idx = __float_as_int(rr.w);
            if (best.accepts(__float_as_uint(dist), idx)) best.insert(__float_as_uint(dist), idx);
        };
        for (int R = 1;; ++R) {
            const int x0 = max(cx - R, 0), x1 = min(cx + R, p.nx - 1);
            if (R == 1) {
                // Shells 0 and 1 together (the whole 3 x 3 x 3 cube -- shell 0 alone can never terminate the search): nine
                // rows, each ONE contiguous range of the cell-sorted copy, nearest first (the query's own row, the four
                // rows sharing a face with it, the four diagonal ones: the list fills with near candidates early and most of
                // the later ones fail the K-th-distance test, which skips the insertion).  The ranges go to shared memory
                // and the candidates of all nine are walked in ONE flat loop: a lane whose row is empty or short moves on
                // to its next row instead of idling until the longest row of the warp is done.
#pragma unroll
                for (int r = 0; r < 9; ++r) {
                    const int z = cz + (int)((KG_CUBE_DZ >> (2 * r)) & 3u) - 1, y = cy + (int)((KG_CUBE_DY >> (2 * r)) & 3u) - 1;
                    const bool ok = z >= 0 && z < p.nz && y >= 0 && y < p.ny;
                    const int c = p.cell_off + p.nx * (y + p.ny * z);
                    row_s[r][threadIdx.x] = ok ? make_int2(__ldg(cell_ptr + c + x0), __ldg(cell_ptr + c + x1 + 1)) : make_int2(0, 0);
                }
                int r = 0;
                int2 ee = row_s[0][threadIdx.x];
                while (true) {
                    while (ee.x >= ee.y && r < 8) ee = row_s[++r][threadIdx.x];
                    if (ee.x >= ee.y) break;
                    candidate(__ldg(sorted + ee.x));
                    ++ee.x;
                }
            } else {
                const int side = 2 * R + 1, n_rows = side * side;
                for (int r = 0; r < n_rows; ++r) {
                    const int dz = r / side - R, dy = r - (dz + R) * side - R;
                    const int z = cz + dz, y = cy + dy;
                    if (z < 0 || z >= p.nz || y < 0 || y >= p.ny) continue;
                    const int c = p.cell_off + p.nx * (y + p.ny * z);
                    // a row on a z / y face of the shell: the whole x range; an inner row: its two end cells
                    const bool full_row = dz == -R || dz == R || dy == -R || dy == R;
                    for (int seg = 0; seg < (full_row ? 1 : 2); ++seg) {
                        const int xa = full_row ? x0 : (seg == 0 ? cx - R : cx + R);
                        const int xb = full_row ? x1 : xa;
                        if (xa < 0 || xb >= p.nx) continue;
                        const int e1 = __ldg(cell_ptr + c + xb + 1);
                        for (int e = __ldg(cell_ptr + c + xa); e < e1; ++e) candidate(__ldg(sorted + e));
                    }
                }
            }
            const bool all = (cx - R <= 0) && (cy - R <= 0) && (cz - R <= 0) && (cx + R >= p.nx - 1) && (cy + R >= p.ny - 1) && (cz + R >= p.nz - 1);
            if (all) break;
            // unvisited references differ from the query by more than (R - rounding of the cell index) * h
            const float errc = 3e-7f * (float)max(max(p.nx, p.ny), p.nz);
            const float reach = ((float)R - 2.f * errc - 1e-4f) * p.h;
            if (reach > 0.f && best.kth(K) < reach * reach) break;
        }
    }
    const int found = min(n_ref, K);
    int64_t *o = out + (size_t)q * K;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        if (k < K) {
            int v = best.id[k];
            if (k >= found) {
                v = -1;
                if (found > 0) {
                    const int src = k % found;
#pragma unroll
                    for (int t = 0; t < KP; ++t) if (t == src) v = best.id[t];
                }
            }
            o[k] = (int64_t)v;
        }
    }
}

struct KgWorkspace {
    int *mm;
    GridPlan *plans;
    int32_t *cell, *cell_ptr, *cell_pts;
    float4 *sorted;
    int32_t *counts;                 // zero-initialised region: counts, scan state, scan ticket (one memset)
    unsigned long long *state;
    unsigned int *ticket;
    size_t zero_bytes, bytes;
    int max_cells;
};

static KgWorkspace carve_kg(void *ws, int n_seg, int n_ref) {
    Carver c(ws);
    KgWorkspace w{};
    w.max_cells = 2 * n_ref + 64 * n_seg;
    w.mm = c.take<int>((size_t)n_seg * 6 + 1);
    w.plans = c.take<GridPlan>((size_t)n_seg);
    w.cell = c.take<int32_t>((size_t)n_ref + 1);
    w.cell_ptr = c.take<int32_t>((size_t)w.max_cells + 2);
    w.cell_pts = c.take<int32_t>((size_t)n_ref + 1);
    w.sorted = c.take<float4>((size_t)n_ref + 1);
    w.counts = c.take<int32_t>((size_t)w.max_cells + 2);
    w.state = c.take<unsigned long long>((size_t)ceil_div(w.max_cells + 1, SCAN_TILE) + 1);
    w.ticket = c.take<unsigned int>(4);
    w.zero_bytes = (size_t)((char *)(w.ticket + 4) - (char *)w.counts);
    w.bytes = align_up(c.off, 256);
    return w;
}

static inline int kg_blocks(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace pcfb

extern "C" size_t pcfb_knn_grid_workspace(int n_seg, int n_ref)
{
    return pcfb::carve_kg(nullptr, n_seg, n_ref).bytes;
}

extern "C" int pcfb_knn_grid_build(const float *ref_xyz, const int32_t *ref_off, int n_seg, int n_ref, float cell_hint,
                                   void *workspace, size_t workspace_bytes, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(n_seg >= 1 && n_ref >= 0, "pcfb_knn_grid_build: bad sizes");
    PCFB_REQUIRE((int64_t)2 * n_ref + 64ll * n_seg < (1ll << 30), "pcfb_knn_grid_build: cloud too large");
    PCFB_REQUIRE(ref_off && workspace && (n_ref == 0 || ref_xyz), "pcfb_knn_grid_build: null pointer");
    KgWorkspace w = carve_kg(workspace, n_seg, n_ref);
    if (workspace_bytes < w.bytes) { set_error("pcfb_knn_grid_build: workspace %zu < %zu", workspace_bytes, w.bytes); return PCFB_ERR_WORKSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    launch_k(kg_init_kernel, ceil_div(n_seg * 6, 256), 256, 0, st, w.mm, n_seg);
    if ((rc = check_launch("kg_init_kernel"))) return rc;
    if (n_ref > 0) {
        launch_k(kg_bounds_kernel, kg_blocks(n_ref), 256, 0, st, ref_xyz, ref_off, n_seg, n_ref, w.mm);
        if ((rc = check_launch("kg_bounds_kernel"))) return rc;
    }
    launch_k(kg_plan_kernel, 1, 32, 0, st, w.mm, ref_off, n_seg, cell_hint, w.plans);
    if ((rc = check_launch("kg_plan_kernel"))) return rc;
    PCFB_CUDA(cudaMemsetAsync(w.counts, 0, w.zero_bytes, st));
    if (n_ref > 0) {
        launch_k(kg_cell_count_kernel, kg_blocks(n_ref), 256, 0, st, ref_xyz, ref_off, n_seg, n_ref, (const GridPlan *)w.plans, w.cell, w.counts);
        if ((rc = check_launch("kg_cell_count_kernel"))) return rc;
    }
    // exclusive scan over max_cells + 1 entries: cell_ptr[c] = first slot of cell c, cell_ptr[max_cells] = n_ref
    launch_k(inv_scan_kernel, ceil_div(w.max_cells + 1, SCAN_TILE), SCAN_THREADS, 0, st, (const int32_t *)w.counts, w.max_cells, w.cell_ptr,
             w.state, w.ticket);
    if ((rc = check_launch("inv_scan_kernel"))) return rc;
    if (n_ref > 0) {
        launch_k(kg_fill_kernel, kg_blocks(n_ref), 256, 0, st, ref_xyz, (const int32_t *)w.cell, n_ref, (const int32_t *)w.cell_ptr, w.counts,
                 w.cell_pts, w.sorted);
        if ((rc = check_launch("kg_fill_kernel"))) return rc;
    }
    return PCFB_OK;
}

extern "C" const int32_t *pcfb_knn_grid_order(int n_seg, int n_ref, const void *workspace)
{
    if (!workspace || n_seg < 1 || n_ref < 0) return nullptr;
    return pcfb::carve_kg(const_cast<void *>(workspace), n_seg, n_ref).cell_pts;
}

extern "C" int pcfb_knn_grid_query(const float *ref_xyz, int n_seg, int n_ref, const float *qry_xyz, const int32_t *qry_off,
                                   int n_qry, int K, const int32_t *qry_order, int64_t *out_idx, const void *workspace,
                                   size_t workspace_bytes, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(K >= 1 && K <= 64, "pcfb_knn_grid_query: K=%d outside [1,64] (use pcfb_knn_packed for larger K)", K);
    PCFB_REQUIRE(n_seg >= 1 && n_ref >= 0 && n_qry >= 0, "pcfb_knn_grid_query: bad sizes");
    if (n_qry == 0) return PCFB_OK;
    PCFB_REQUIRE(qry_xyz && qry_off && out_idx && workspace && (n_ref == 0 || ref_xyz), "pcfb_knn_grid_query: null pointer");
    KgWorkspace w = carve_kg(const_cast<void *>(workspace), n_seg, n_ref);
    if (workspace_bytes < w.bytes) { set_error("pcfb_knn_grid_query: workspace %zu < %zu", workspace_bytes, w.bytes); return PCFB_ERR_WORKSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = ceil_div(n_qry, KG_THREADS);
    const float4 *sorted = w.sorted;
    const int32_t *cell_ptr = w.cell_ptr;
    const GridPlan *plans = w.plans;
    // spatially coherent order of the queries: the caller's permutation (e.g. another grid's cell order of the same
    // cloud), or -- self queries, the query cloud IS the reference cloud -- this grid's own cell order
    const int32_t *order = qry_order ? qry_order : ((qry_xyz == ref_xyz && n_qry == n_ref) ? w.cell_pts : nullptr);
    if (K <= 16)
        launch_k(knn_grid_query_kernel<16>, grid, KG_THREADS, 0, st, sorted, plans, cell_ptr, order, qry_xyz, qry_off, n_seg, n_qry, K, out_idx);
    else if (K <= 32)
        launch_k(knn_grid_query_kernel<32>, grid, KG_THREADS, 0, st, sorted, plans, cell_ptr, order, qry_xyz, qry_off, n_seg, n_qry, K, out_idx);
    else
        launch_k(knn_grid_query_kernel<64>, grid, KG_THREADS, 0, st, sorted, plans, cell_ptr, order, qry_xyz, qry_off, n_seg, n_qry, K, out_idx);
    return check_launch("knn_grid_query_kernel");
}
