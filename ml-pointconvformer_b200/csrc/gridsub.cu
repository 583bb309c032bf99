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
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_seg * 6) mm[i] = (i % 6 < 3) ? 0x7fffffff : (int)0x80000000;
}

__global__ void gs_bounds_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ seg_off, int n_seg,
                                 int n_pts, int *__restrict__ mm)
{
    pdl_wait();
    for (int base = blockIdx.x * blockDim.x; base < n_pts; base += gridDim.x * blockDim.x) {     // warp-uniform trip count
        const int i = base + threadIdx.x;
        const bool in = i < n_pts;
        const size_t j = in ? (size_t)i : 0;
        warp_scene_minmax(in, in ? find_seg(seg_off, n_seg, i) : 0, float_to_ordered(xyz[3 * j]), float_to_ordered(xyz[3 * j + 1]),
                          float_to_ordered(xyz[3 * j + 2]), mm);
    }
}

// origin = floor(min * (1/dl)) * dl ; N = floor((max - origin)/dl) + 1   (grid_subsampling.cpp:28-36)
__global__ void gs_finish_bounds_kernel(const int *__restrict__ mm, const int32_t *__restrict__ seg_off, int n_seg,
                                        float dl, float *__restrict__ origin, int32_t *__restrict__ dims)
{
    pdl_wait();
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
    pdl_wait();
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
    pdl_wait();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_cells; i += gridDim.x * blockDim.x)
        flag[i] = (cell_ptr[i + 1] > cell_ptr[i]) ? 1 : 0;
}

__global__ void gs_counts_kernel(const int32_t *__restrict__ rank, const int32_t *__restrict__ cell_off, int n_seg,
                                 int32_t *__restrict__ out_counts) {
    pdl_wait();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_seg) out_counts[s] = rank[cell_off[s + 1]] - rank[cell_off[s]];
}

__global__ void gs_emit_kernel(const float *__restrict__ xyz, const float *__restrict__ feats, int F,
                               const int32_t *__restrict__ cell_ptr, const int32_t *__restrict__ pts,
                               const int32_t *__restrict__ rank, int total_cells,
                               float *__restrict__ out_xyz, float *__restrict__ out_feats)
{
    pdl_wait();
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


// ---- device-sized variants: the whole pyramid without per-level host reads ---------------------------------------------
// The input size of level l is only known on the device (it is the output count of level l-1): kernels run over the host's
// upper bound n_pts_max and guard with the device total seg_off[n_seg]; points past it get cell id -1, which the CSR
// transpose ignores.  The dense cell table is sized by a host upper bound (cells_max, from the level-0 bounding box).
__global__ void gs_bounds_dev_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ seg_off, int n_seg,
                                     int n_pts_max, int *__restrict__ mm)
{
    pdl_wait();
    const int n_pts = min(seg_off[n_seg], n_pts_max);
    for (int base = blockIdx.x * blockDim.x; base < n_pts; base += gridDim.x * blockDim.x) {     // warp-uniform trip count
        const int i = base + threadIdx.x;
        const bool in = i < n_pts;
        const size_t j = in ? (size_t)i : 0;
        warp_scene_minmax(in, in ? find_seg(seg_off, n_seg, i) : 0, float_to_ordered(xyz[3 * j]), float_to_ordered(xyz[3 * j + 1]),
                          float_to_ordered(xyz[3 * j + 2]), mm);
    }
}

// cell_off = exclusive prefix of the per-scene cell counts; status |= 1 (and an empty grid) if they exceed cells_max
__global__ void gs_cell_off_kernel(int32_t *__restrict__ dims, int n_seg, long long cells_max, int32_t *__restrict__ cell_off,
                                   int32_t *__restrict__ status)
{
    pdl_wait();
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    long long tot = 0;
    for (int s = 0; s < n_seg; ++s) {
        cell_off[s] = (int32_t)tot;
        tot += (long long)dims[3 * s] * dims[3 * s + 1] * dims[3 * s + 2];
        if (tot > cells_max) break;
    }
    if (tot > cells_max) {
        if (status) atomicOr(status, 1);
        for (int s = 0; s < n_seg; ++s) { cell_off[s] = 0; dims[3 * s] = dims[3 * s + 1] = dims[3 * s + 2] = 0; }
        tot = 0;
    }
    cell_off[n_seg] = (int32_t)tot;
}

__global__ void gs_cell_dev_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ seg_off, int n_seg,
                                   int n_pts_max, float dl, const float *__restrict__ origin, const int32_t *__restrict__ dims,
                                   const int32_t *__restrict__ cell_off, int64_t *__restrict__ cell)
{
    pdl_wait();
    const int n_pts = min(seg_off[n_seg], n_pts_max);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pts_max; i += gridDim.x * blockDim.x) {
        if (i >= n_pts) { cell[i] = -1; continue; }
        const int s = find_seg(seg_off, n_seg, i);
        const int64_t nx = dims[3 * s], ny = dims[3 * s + 1], nz = dims[3 * s + 2];
        if (nx * ny * nz == 0) { cell[i] = -1; continue; }
        const int ix = (int)floorf(__fdiv_rn(__fsub_rn(xyz[3 * (size_t)i], origin[3 * s]), dl));
        const int iy = (int)floorf(__fdiv_rn(__fsub_rn(xyz[3 * (size_t)i + 1], origin[3 * s + 1]), dl));
        const int iz = (int)floorf(__fdiv_rn(__fsub_rn(xyz[3 * (size_t)i + 2], origin[3 * s + 2]), dl));
        cell[i] = (int64_t)cell_off[s] + ix + nx * iy + nx * ny * iz;
    }
}

// per-scene output counts -> the next level's scene offsets
__global__ void gs_out_off_kernel(const int32_t *__restrict__ rank, const int32_t *__restrict__ cell_off, int n_seg,
                                  int32_t *__restrict__ out_seg_off)
{
    pdl_wait();
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int tot = 0;
    out_seg_off[0] = 0;
    for (int s = 0; s < n_seg; ++s) {
        tot += rank[cell_off[s + 1]] - rank[cell_off[s]];
        out_seg_off[s + 1] = tot;
    }
}

// ---- voxelisation: one point per occupied voxel (the smallest input index), util/voxelize.py:44-70 'deterministic' -----
// discrete = floor(coord / voxel) in float64 (numpy promotes float32 coordinates / np.array(voxel_size) to float64, NEP 50);
// key = ravel of (discrete - per-scene min) -- ravel_hash_vec, voxelize.py:26-41
__global__ void vx_cell_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ seg_off, int n_seg, int n_pts,
                               double voxel, const int *__restrict__ mm, const int32_t *__restrict__ dims,
                               const int32_t *__restrict__ cell_off, int32_t *__restrict__ first)
{
    pdl_wait();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pts; i += gridDim.x * blockDim.x) {
        const int s = find_seg(seg_off, n_seg, i);
        const int64_t nx = dims[3 * s], ny = dims[3 * s + 1], nz = dims[3 * s + 2];
        if (nx * ny * nz == 0) continue;
        int d[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double lo = floor((double)ordered_to_float(mm[s * 6 + a]) / voxel);
            d[a] = (int)(floor((double)xyz[3 * (size_t)i + a] / voxel) - lo);
        }
        // Fortran-style ravel of voxelize.py:36-40: ((x * ny + y) * nz + z): ascending key = the reference's sorted order
        const int64_t key = ((int64_t)d[0] * ny + d[1]) * nz + d[2];
        atomicMin(&first[cell_off[s] + key], i);                   // integer min: deterministic
    }
}

__global__ void vx_dims_kernel(const int *__restrict__ mm, const int32_t *__restrict__ seg_off, int n_seg, double voxel,
                               int32_t *__restrict__ dims)
{
    pdl_wait();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const bool empty = seg_off[s + 1] <= seg_off[s];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (empty) { dims[3 * s + a] = 0; continue; }
        const double lo = floor((double)ordered_to_float(mm[s * 6 + a]) / voxel);
        const double hi = floor((double)ordered_to_float(mm[s * 6 + 3 + a]) / voxel);
        dims[3 * s + a] = (int)(hi - lo) + 1;
    }
}

__global__ void vx_fill_kernel(int32_t *__restrict__ first, long long n) {
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) first[i] = 0x7fffffff;
}
__global__ void vx_flag_kernel(const int32_t *__restrict__ first, const int32_t *__restrict__ cell_off, int n_seg, int cells_max,
                               int32_t *__restrict__ flag) {
    pdl_wait();
    const int total = cell_off[n_seg];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= cells_max; i += gridDim.x * blockDim.x)
        flag[i] = (i < total && first[i] != 0x7fffffff) ? 1 : 0;
}
__global__ void vx_emit_kernel(const int32_t *__restrict__ first, const int32_t *__restrict__ rank, const int32_t *__restrict__ cell_off,
                               int n_seg, int32_t *__restrict__ out_idx) {
    pdl_wait();
    const int total = cell_off[n_seg];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        if (first[i] != 0x7fffffff) out_idx[rank[i]] = first[i];
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
    launch_k(gs_init_bounds_kernel, ceil_div(n_seg * 6, 256), 256, 0, st, mm, n_seg);
    if ((rc = check_launch("gs_init_bounds_kernel"))) return rc;
    if (n_pts > 0) {
        launch_k(gs_bounds_kernel, blocks_for(n_pts), 256, 0, st, xyz, seg_off, n_seg, n_pts, mm);
        if ((rc = check_launch("gs_bounds_kernel"))) return rc;
    }
    launch_k(gs_finish_bounds_kernel, ceil_div(n_seg, 128), 128, 0, st, mm, seg_off, n_seg, dl, out_origin, out_dims);
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
        launch_k(gs_cell_kernel, blocks_for(n_pts), 256, 0, st, xyz, seg_off, n_seg, n_pts, dl, origin, dims, cell_off, w.cell);
        if ((rc = check_launch("gs_cell_kernel"))) return rc;
    }
    if ((rc = pcfb_knn_inverse(w.cell, n_pts, 1, (int)total_cells, w.pts, w.zero_k, w.cell_ptr, w.inv_ws, w.inv_ws_bytes, stream))) return rc;
    // occupied-cell rank = exclusive scan of (count > 0)
    PCFB_CUDA(cudaMemsetAsync(w.state, 0, (size_t)((char *)w.inv_ws - (char *)w.state), st));
    PCFB_CUDA(cudaMemsetAsync(w.flag + total_cells, 0, sizeof(int32_t), st));
    if (total_cells > 0) {
        launch_k(gs_flag_kernel, blocks_for(total_cells), 256, 0, st, w.cell_ptr, (int)total_cells, w.flag);
        if ((rc = check_launch("gs_flag_kernel"))) return rc;
    }
    launch_k(inv_scan_kernel, ceil_div((int)total_cells + 1, SCAN_TILE), SCAN_THREADS, 0, st, w.flag, (int)total_cells, w.rank, w.state, w.ticket);
    if ((rc = check_launch("inv_scan_kernel"))) return rc;
    launch_k(gs_counts_kernel, ceil_div(n_seg, 128), 128, 0, st, w.rank, cell_off, n_seg, out_counts);
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
    launch_k(gs_emit_kernel, blocks_for(total_cells), 256, 0, static_cast<cudaStream_t>(stream), xyz, feats, F, w.cell_ptr, w.pts, w.rank, (int)total_cells, out_xyz, out_feats);
    return check_launch("gs_emit_kernel");
}


// ---- one pyramid level with device-side sizes (no host read) -----------------------------------------------------------
namespace pcfb {
struct GsDevExtra { float *origin; int32_t *dims, *cell_off; size_t bytes; };
static GsDevExtra carve_gs_dev(void *ws, size_t base_bytes, int n_seg) {
    Carver c(ws);
    c.off = base_bytes;
    GsDevExtra e{};
    e.origin = c.take<float>((size_t)n_seg * 3 + 1);
    e.dims = c.take<int32_t>((size_t)n_seg * 3 + 1);
    e.cell_off = c.take<int32_t>((size_t)n_seg + 2);
    e.bytes = align_up(c.off, 256);
    return e;
}
}  // namespace pcfb

extern "C" size_t pcfb_pyramid_level_workspace(int n_seg, int n_pts_max, int64_t cells_max)
{
    using namespace pcfb;
    return carve_gs_dev(nullptr, carve_gs(nullptr, n_seg, n_pts_max, cells_max).bytes, n_seg).bytes;
}

extern "C" int pcfb_pyramid_level(const float *xyz, const float *feats, int F, const int32_t *seg_off, int n_seg, int n_pts_max,
                                  float dl, int64_t cells_max, float *out_xyz, float *out_feats, int32_t *out_seg_off,
                                  int32_t *status, void *workspace, size_t workspace_bytes, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(n_seg >= 1 && n_pts_max >= 0 && dl > 0.f && F >= 0, "pcfb_pyramid_level: bad arguments");
    PCFB_REQUIRE(cells_max >= 1 && cells_max < (1ll << 30), "pcfb_pyramid_level: cells_max = %lld outside [1, 2^30)", (long long)cells_max);
    PCFB_REQUIRE(xyz && seg_off && out_xyz && out_seg_off && workspace && (F == 0 || (feats && out_feats)), "pcfb_pyramid_level: null pointer");
    GsWorkspace w = carve_gs(workspace, n_seg, n_pts_max, cells_max);
    GsDevExtra e = carve_gs_dev(workspace, w.bytes, n_seg);
    if (workspace_bytes < e.bytes) { set_error("pcfb_pyramid_level: workspace %zu < %zu", workspace_bytes, e.bytes); return PCFB_ERR_WORKSPACE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    launch_k(gs_init_bounds_kernel, ceil_div(n_seg * 6, 256), 256, 0, st, w.mm, n_seg);
    if ((rc = check_launch("gs_init_bounds_kernel"))) return rc;
    launch_k(gs_bounds_dev_kernel, blocks_for(n_pts_max), 256, 0, st, xyz, seg_off, n_seg, n_pts_max, w.mm);
    if ((rc = check_launch("gs_bounds_dev_kernel"))) return rc;
    launch_k(gs_finish_bounds_kernel, ceil_div(n_seg, 128), 128, 0, st, w.mm, seg_off, n_seg, dl, e.origin, e.dims);
    if ((rc = check_launch("gs_finish_bounds_kernel"))) return rc;
    launch_k(gs_cell_off_kernel, 1, 32, 0, st, e.dims, n_seg, (long long)cells_max, e.cell_off, status);
    if ((rc = check_launch("gs_cell_off_kernel"))) return rc;
    launch_k(gs_cell_dev_kernel, blocks_for(n_pts_max), 256, 0, st, xyz, seg_off, n_seg, n_pts_max, dl, e.origin, e.dims, e.cell_off, w.cell);
    if ((rc = check_launch("gs_cell_dev_kernel"))) return rc;
    if ((rc = pcfb_knn_inverse(w.cell, n_pts_max, 1, (int)cells_max, w.pts, w.zero_k, w.cell_ptr, w.inv_ws, w.inv_ws_bytes, stream))) return rc;
    PCFB_CUDA(cudaMemsetAsync(w.state, 0, (size_t)((char *)w.inv_ws - (char *)w.state), st));
    PCFB_CUDA(cudaMemsetAsync(w.flag + cells_max, 0, sizeof(int32_t), st));
    launch_k(gs_flag_kernel, blocks_for(cells_max), 256, 0, st, w.cell_ptr, (int)cells_max, w.flag);
    if ((rc = check_launch("gs_flag_kernel"))) return rc;
    launch_k(inv_scan_kernel, ceil_div((int)cells_max + 1, SCAN_TILE), SCAN_THREADS, 0, st, w.flag, (int)cells_max, w.rank, w.state, w.ticket);
    if ((rc = check_launch("inv_scan_kernel"))) return rc;
    launch_k(gs_out_off_kernel, 1, 32, 0, st, w.rank, e.cell_off, n_seg, out_seg_off);
    if ((rc = check_launch("gs_out_off_kernel"))) return rc;
    launch_k(gs_emit_kernel, blocks_for(cells_max), 256, 0, st, xyz, feats, F, w.cell_ptr, w.pts, w.rank, (int)cells_max, out_xyz, out_feats);
    return check_launch("gs_emit_kernel");
}

extern "C" size_t pcfb_voxelize_workspace(int n_seg, int n_pts, int64_t cells_max)
{
    using namespace pcfb;
    Carver c(nullptr);
    c.take<int>((size_t)n_seg * 6 + 1);
    c.take<int32_t>((size_t)n_seg * 3 + 1);
    c.take<int32_t>((size_t)n_seg + 2);
    c.take<int32_t>((size_t)cells_max + 2);
    c.take<int32_t>((size_t)cells_max + 2);
    c.take<int32_t>((size_t)cells_max + 2);
    c.take<unsigned long long>((size_t)(cells_max + 1) / SCAN_TILE + 2);
    c.take<unsigned int>(4);
    (void)n_pts;
    return align_up(c.off, 256);
}

extern "C" int pcfb_voxelize(const float *xyz, const int32_t *seg_off, int n_seg, int n_pts, double voxel, int64_t cells_max,
                             int32_t *out_idx, int32_t *out_seg_off, int32_t *status, void *workspace, size_t workspace_bytes,
                             void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(n_seg >= 1 && n_pts >= 0 && voxel > 0.0, "pcfb_voxelize: bad arguments");
    PCFB_REQUIRE(cells_max >= 1 && cells_max < (1ll << 30), "pcfb_voxelize: cells_max = %lld outside [1, 2^30)", (long long)cells_max);
    PCFB_REQUIRE(xyz && seg_off && out_idx && out_seg_off && workspace, "pcfb_voxelize: null pointer");
    if (workspace_bytes < pcfb_voxelize_workspace(n_seg, n_pts, cells_max)) { set_error("pcfb_voxelize: workspace too small"); return PCFB_ERR_WORKSPACE; }
    Carver c(workspace);
    int *mm = c.take<int>((size_t)n_seg * 6 + 1);
    int32_t *dims = c.take<int32_t>((size_t)n_seg * 3 + 1);
    int32_t *cell_off = c.take<int32_t>((size_t)n_seg + 2);
    int32_t *first = c.take<int32_t>((size_t)cells_max + 2);
    int32_t *flag = c.take<int32_t>((size_t)cells_max + 2);
    int32_t *rank = c.take<int32_t>((size_t)cells_max + 2);
    unsigned long long *state = c.take<unsigned long long>((size_t)(cells_max + 1) / SCAN_TILE + 2);
    unsigned int *ticket = c.take<unsigned int>(4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    launch_k(gs_init_bounds_kernel, ceil_div(n_seg * 6, 256), 256, 0, st, mm, n_seg);
    if ((rc = check_launch("gs_init_bounds_kernel"))) return rc;
    if (n_pts > 0) {
        launch_k(gs_bounds_kernel, blocks_for(n_pts), 256, 0, st, xyz, seg_off, n_seg, n_pts, mm);
        if ((rc = check_launch("gs_bounds_kernel"))) return rc;
    }
    launch_k(vx_dims_kernel, ceil_div(n_seg, 128), 128, 0, st, mm, seg_off, n_seg, voxel, dims);
    if ((rc = check_launch("vx_dims_kernel"))) return rc;
    launch_k(gs_cell_off_kernel, 1, 32, 0, st, dims, n_seg, (long long)cells_max, cell_off, status);
    if ((rc = check_launch("gs_cell_off_kernel"))) return rc;
    launch_k(vx_fill_kernel, blocks_for(cells_max + 1), 256, 0, st, first, (long long)cells_max + 1);
    if ((rc = check_launch("vx_fill_kernel"))) return rc;
    if (n_pts > 0) {
        launch_k(vx_cell_kernel, blocks_for(n_pts), 256, 0, st, xyz, seg_off, n_seg, n_pts, voxel, mm, dims, cell_off, first);
        if ((rc = check_launch("vx_cell_kernel"))) return rc;
    }
    PCFB_CUDA(cudaMemsetAsync(state, 0, (size_t)((char *)ticket + 4 * sizeof(unsigned int) - (char *)state), st));
    launch_k(vx_flag_kernel, blocks_for(cells_max + 1), 256, 0, st, first, cell_off, n_seg, (int)cells_max, flag);
    if ((rc = check_launch("vx_flag_kernel"))) return rc;
    launch_k(inv_scan_kernel, ceil_div((int)cells_max + 1, SCAN_TILE), SCAN_THREADS, 0, st, flag, (int)cells_max, rank, state, ticket);
    if ((rc = check_launch("inv_scan_kernel"))) return rc;
    launch_k(gs_out_off_kernel, 1, 32, 0, st, rank, cell_off, n_seg, out_seg_off);
    if ((rc = check_launch("gs_out_off_kernel"))) return rc;
    launch_k(vx_emit_kernel, blocks_for(cells_max), 256, 0, st, first, rank, cell_off, n_seg, out_idx);
    return check_launch("vx_emit_kernel");
}
