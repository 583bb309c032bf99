// Grid (voxel) subsampling with barycentres on packed scenes (sm_100a).
//
// Replaces grid_subsampling() (/root/reference/cpp_wrappers/cpp_subsampling/grid_subsampling/
// grid_subsampling.cpp:9-110; SampledData grid_subsampling.h:13-83; min/max_point cloud.cpp:27-66),
// which the reference runs single-threaded on the CPU inside DataLoader workers
// (datasetCommon.py:384-420), accumulating into an unordered_map.
//
// Here: (1) per-scene bounding box by ordered-int atomics -> origin / grid dims with the reference's
// fp32 arithmetic; (2) every point gets a dense cell id = cell_off[scene] + iX + NX*iY + NX*NY*iZ;
// (3) the point->cell map is CSR-transposed with the same histogram/scan/fill/rank-sort kernels as
// the kNN inverse map (K = 1), which lists each cell's points in ascending input order; (4) one
// thread per occupied cell adds its points sequentially (__fadd_rn, input order) so that sums are
// BIT-IDENTICAL to the reference's running sums; output order is ascending (scene, voxel key).
// All arithmetic that decides a voxel index uses explicit _rn intrinsics (no FMA contraction).
#include "common.cuh"
#include "scan.cuh"

namespace pcfb {

__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) {
    return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

__device__ __forceinline__ int find_seg(const int32_t *__restrict__ off, int n_seg, int i) {
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void gs_init_bounds_kernel(int *__restrict__ mm, int n_seg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_seg * 6) mm[i] = (i % 6 < 3) ? 0x7fffffff : (int)0x80000000;
}

__global__ void gs_bounds_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ seg_off, int n_seg,
                                 int n_pts, int *__restrict__ mm)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pts; i += gridDim.x * blockDim.x) {
        const int s = find_seg(seg_off, n_seg, i);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int o = float_to_ordered(xyz[3 * (size_t)i + d]);
            atomicMin(&mm[s * 6 + d], o);
            atomicMax(&mm[s * 6 + 3 + d], o);
        }
    }
}

// origin = floor(min * (1/dl)) * dl ; N = floor((max - origin)/dl) + 1   (grid_subsampling.cpp:28-36)
__global__ void gs_finish_bounds_kernel(const int *__restrict__ mm, const int32_t *__restrict__ seg_off, int n_seg,
                                        float dl, float *__restrict__ origin, int32_t *__restrict__ dims)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const bool empty = seg_off[s + 1] <= seg_off[s];
    const float inv = __fdiv_rn(1.0f, dl);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (empty) { origin[3 * s + d] = 0.f; dims[3 * s + d] = 0; continue; }
        const float mn = ordered_to_float(mm[s * 6 + d]), mx = ordered_to_float(mm[s * 6 + 3 + d]);
        const float o = __fmul_rn(floorf(__fmul_rn(mn, inv)), dl);
        origin[3 * s + d] = o;
        dims[3 * s + d] = (int)floorf(__fdiv_rn(__fsub_rn(mx, o), dl)) + 1;
    }
}

__global__ void gs_cell_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ seg_off, int n_seg,
                               int n_pts, float dl, const float *__restrict__ origin, const int32_t *__restrict__ dims,
                               const int32_t *__restrict__ cell_off, int64_t *__restrict__ cell)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pts; i += gridDim.x * blockDim.x) {
        const int s = find_seg(seg_off, n_seg, i);
        const int ix = (int)floorf(__fdiv_rn(__fsub_rn(xyz[3 * (size_t)i], origin[3 * s]), dl));
        const int iy = (int)floorf(__fdiv_rn(__fsub_rn(xyz[3 * (size_t)i + 1], origin[3 * s + 1]), dl));
        const int iz = (int)floorf(__fdiv_rn(__fsub_rn(xyz[3 * (size_t)i + 2], origin[3 * s + 2]), dl));
        const int64_t nx = dims[3 * s], ny = dims[3 * s + 1];
        cell[i] = (int64_t)cell_off[s] + ix + nx * iy + nx * ny * iz;
    }
}

__global__ void gs_flag_kernel(const int32_t *__restrict__ cell_ptr, int total_cells, int32_t *__restrict__ flag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_cells; i += gridDim.x * blockDim.x)
        flag[i] = (cell_ptr[i + 1] > cell_ptr[i]) ? 1 : 0;
}

__global__ void gs_counts_kernel(const int32_t *__restrict__ rank, const int32_t *__restrict__ cell_off, int n_seg,
                                 int32_t *__restrict__ out_counts) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_seg) out_counts[s] = rank[cell_off[s + 1]] - rank[cell_off[s]];
}

__global__ void gs_emit_kernel(const float *__restrict__ xyz, const float *__restrict__ feats, int F,
                               const int32_t *__restrict__ cell_ptr, const int32_t *__restrict__ pts,
                               const int32_t *__restrict__ rank, int total_cells,
                               float *__restrict__ out_xyz, float *__restrict__ out_feats)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < total_cells; c += gridDim.x * blockDim.x) {
        const int beg = cell_ptr[c], end = cell_ptr[c + 1];
        if (end <= beg) continue;
        const int o = rank[c];
        const int cnt = end - beg;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int e = beg; e < end; ++e) {
            const size_t i = (size_t)pts[e];
            sx = __fadd_rn(sx, xyz[3 * i]); sy = __fadd_rn(sy, xyz[3 * i + 1]); sz = __fadd_rn(sz, xyz[3 * i + 2]);
        }
        const float inv = __double2float_rn(1.0 / (double)cnt);       // point * (1.0 / count), cpp:91
        out_xyz[3 * (size_t)o] = __fmul_rn(sx, inv);
        out_xyz[3 * (size_t)o + 1] = __fmul_rn(sy, inv);
        out_xyz[3 * (size_t)o + 2] = __fmul_rn(sz, inv);
        const float fc = (float)cnt;
        for (int f = 0; f < F; ++f) {
            float sf = 0.f;
            for (int e = beg; e < end; ++e) sf = __fadd_rn(sf, feats[(size_t)pts[e] * F + f]);
            out_feats[(size_t)o * F + f] = __fdiv_rn(sf, fc);        // f / count, cpp:94-98
        }
    }
}

struct GsWorkspace {
    int *mm;             // [n_seg*6] ordered-int min/max
    int64_t *cell;       // [n_pts]
    int32_t *cell_ptr;   // [cells+1]   CSR offsets (inv_idx)
    int32_t *pts;        // [n_pts]     point ids per cell, ascending
    uint8_t *zero_k;     // [n_pts]
    int32_t *flag;       // [cells+1]
    int32_t *rank;       // [cells+1]   exclusive scan of flag
    unsigned long long *state;
    unsigned int *ticket;
    void *inv_ws;
    size_t inv_ws_bytes, bytes;
};

static GsWorkspace carve_gs(void *ws, int n_seg, int n_pts, int64_t cells) {
    Carver c(ws);
    GsWorkspace w{};
    w.mm = c.take<int>((size_t)n_seg * 6 + 1);
    w.cell = c.take<int64_t>((size_t)n_pts + 1);
    w.cell_ptr = c.take<int32_t>((size_t)cells + 2);
    w.pts = c.take<int32_t>((size_t)n_pts + 1);
    w.zero_k = c.take<uint8_t>((size_t)n_pts + 1);
    w.flag = c.take<int32_t>((size_t)cells + 2);
    w.rank = c.take<int32_t>((size_t)cells + 2);
    w.state = c.take<unsigned long long>((size_t)(cells + 1) / SCAN_TILE + 2);
    w.ticket = c.take<unsigned int>(4);
    w.inv_ws_bytes = pcfb_knn_inverse_workspace(n_pts, 1, (int)cells);
    w.inv_ws = c.take<char>(w.inv_ws_bytes);
    w.bytes = align_up(c.off, 256);
    return w;
}

static inline int blocks_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace pcfb

extern "C" size_t pcfb_gridsub_workspace(int n_seg, int n_pts, int64_t total_cells)
{
    return pcfb::carve_gs(nullptr, n_seg, n_pts, total_cells).bytes;
}

extern "C" int pcfb_gridsub_bounds(const float *xyz, const int32_t *seg_off, int n_seg, int n_pts, float dl,
                                   float *out_origin, int32_t *out_dims, void *workspace, size_t workspace_bytes,
                                   void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(n_seg >= 1 && n_pts >= 0 && dl > 0.f, "pcfb_gridsub_bounds: bad arguments");
    PCFB_REQUIRE(xyz && seg_off && out_origin && out_dims && workspace, "pcfb_gridsub_bounds: null pointer");
    PCFB_REQUIRE(workspace_bytes >= (size_t)n_seg * 6 * sizeof(int), "pcfb_gridsub_bounds: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int *mm = static_cast<int *>(workspace);
    int rc;
    gs_init_bounds_kernel<<<ceil_div(n_seg * 6, 256), 256, 0, st>>>(mm, n_seg);
    if ((rc = check_launch("gs_init_bounds_kernel"))) return rc;
    if (n_pts > 0) {
        gs_bounds_kernel<<<blocks_for(n_pts), 256, 0, st>>>(xyz, seg_off, n_seg, n_pts, mm);
        if ((rc = check_launch("gs_bounds_kernel"))) return rc;
    }
    gs_finish_bounds_kernel<<<ceil_div(n_seg, 128), 128, 0, st>>>(mm, seg_off, n_seg, dl, out_origin, out_dims);
    return check_launch("gs_finish_bounds_kernel");
}

extern "C" int pcfb_gridsub_count(const float *xyz, const int32_t *seg_off, int n_seg, int n_pts, float dl,
                                  const float *origin, const int32_t *dims, const int32_t *cell_off,
                                  int64_t total_cells, int32_t *out_counts, void *workspace,
                                  size_t workspace_bytes, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(n_seg >= 1 && n_pts >= 0 && dl > 0.f, "pcfb_gridsub_count: bad arguments");
    PCFB_REQUIRE(total_cells >= 0 && total_cells < (1ll << 30), "pcfb_gridsub_count: %lld cells is too many for the dense binning", (long long)total_cells);
    PCFB_REQUIRE(xyz && seg_off && origin && dims && cell_off && out_counts && workspace, "pcfb_gridsub_count: null pointer");
    GsWorkspace w = carve_gs(workspace, n_seg, n_pts, total_cells);
    if (workspace_bytes < w.bytes) { set_error("pcfb_gridsub_count: workspace %zu < %zu", workspace_bytes, w.bytes); return PCFB_ERR_WORKSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if (n_pts > 0) {
        gs_cell_kernel<<<blocks_for(n_pts), 256, 0, st>>>(xyz, seg_off, n_seg, n_pts, dl, origin, dims, cell_off, w.cell);
        if ((rc = check_launch("gs_cell_kernel"))) return rc;
    }
    if ((rc = pcfb_knn_inverse(w.cell, n_pts, 1, (int)total_cells, w.pts, w.zero_k, w.cell_ptr, w.inv_ws, w.inv_ws_bytes, stream))) return rc;
    // occupied-cell rank = exclusive scan of (count > 0)
    PCFB_CUDA(cudaMemsetAsync(w.state, 0, (size_t)((char *)w.inv_ws - (char *)w.state), st));
    PCFB_CUDA(cudaMemsetAsync(w.flag + total_cells, 0, sizeof(int32_t), st));
    if (total_cells > 0) {
        gs_flag_kernel<<<blocks_for(total_cells), 256, 0, st>>>(w.cell_ptr, (int)total_cells, w.flag);
        if ((rc = check_launch("gs_flag_kernel"))) return rc;
    }
    inv_scan_kernel<<<ceil_div((int)total_cells + 1, SCAN_TILE), SCAN_THREADS, 0, st>>>(w.flag, (int)total_cells, w.rank, w.state, w.ticket);
    if ((rc = check_launch("inv_scan_kernel"))) return rc;
    gs_counts_kernel<<<ceil_div(n_seg, 128), 128, 0, st>>>(w.rank, cell_off, n_seg, out_counts);
    return check_launch("gs_counts_kernel");
}

extern "C" int pcfb_gridsub_emit(const float *xyz, const float *feats, int n_seg, int n_pts, int F,
                                 int64_t total_cells, float *out_xyz, float *out_feats, void *workspace,
                                 size_t workspace_bytes, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(F >= 0 && n_pts >= 0 && total_cells >= 0, "pcfb_gridsub_emit: bad arguments");
    PCFB_REQUIRE(xyz && out_xyz && workspace && (F == 0 || (feats && out_feats)), "pcfb_gridsub_emit: null pointer");
    GsWorkspace w = carve_gs(workspace, n_seg, n_pts, total_cells);
    if (workspace_bytes < w.bytes) { set_error("pcfb_gridsub_emit: workspace %zu < %zu", workspace_bytes, w.bytes); return PCFB_ERR_WORKSPACE; }
    if (total_cells == 0) return PCFB_OK;
    gs_emit_kernel<<<blocks_for(total_cells), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        xyz, feats, F, w.cell_ptr, w.pts, w.rank, (int)total_cells, out_xyz, out_feats);
    return check_launch("gs_emit_kernel");
}
