// Edge geometry (sm_100a): localized xyz r = xyz_in[nei] - centre and the 12-d viewpoint-invariant
// features of VI_coordinate_transform (/root/reference/layer_utils.py:176-231), one thread per edge.
// Inputs are 24 B per point (L2-resident), outputs 12 / 48 B per edge -> HBM-write bound.
// F.normalize semantics: x / max(||x||, 1e-12).
#include "common.cuh"

namespace pcfb {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 ld3(const float *p) { return V3{p[0], p[1], p[2]}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ V3 normalize3(V3 a) {
    const float n = sqrtf(dot3(a, a));
    const float inv = 1.0f / fmaxf(n, 1e-12f);
    return V3{a.x * inv, a.y * inv, a.z * inv};
}

__global__ void edge_geometry_kernel(const float *__restrict__ xyz_in, const float *__restrict__ nrm_in,
                                     const float *__restrict__ xyz_out, const float *__restrict__ nrm_out,
                                     const int64_t *__restrict__ nei, int n_in, int64_t n_edges, int K,
                                     float *__restrict__ out_r, float *__restrict__ out_vi)
{
    pdl_wait();
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = e / K;
        int64_t p = nei[e];
        const bool valid = p >= 0 && p < n_in;
        if (!valid) p = 0;
        const V3 c = ld3(xyz_out + 3 * m);
        const V3 g = ld3(xyz_in + 3 * p);
        V3 r{g.x - c.x, g.y - c.y, g.z - c.z};
        if (!valid) r = V3{0.f, 0.f, 0.f};
        if (out_r) { out_r[3 * e] = r.x; out_r[3 * e + 1] = r.y; out_r[3 * e + 2] = r.z; }
        if (out_vi) {
            V3 nj = ld3(nrm_in + 3 * p);
            if (!valid) nj = V3{0.f, 0.f, 0.f};
            const V3 ni = ld3(nrm_out + 3 * m);
            const V3 rh = normalize3(r);
            const float nr = dot3(ni, rh);
            const V3 v = normalize3(V3{ni.x - nr * rh.x, ni.y - nr * rh.y, ni.z - nr * rh.z});
            const V3 w = normalize3(cross3(rh, v));
            const float t3 = dot3(rh, nj);
            float4 a, b, d;
            a.x = dot3(nj, ni);            // theta1
            a.y = nr;                      // theta2 = r_hat . n_i
            a.z = t3;                      // theta3
            a.w = dot3(r, ni);             // theta4
            b.x = t3;                      // theta5
            b.y = dot3(nj, v);             // theta6
            b.z = dot3(nj, w);             // theta7
            b.w = dot3(r, cross3(nj, ni)); // theta8
            d.x = sqrtf(dot3(r, r));       // theta9
            d.y = r.x; d.z = r.y; d.w = r.z;
            float4 *o = reinterpret_cast<float4 *>(out_vi + 12 * e);
            o[0] = a; o[1] = b; o[2] = d;
        }
    }
}

}  // namespace pcfb

extern "C" int pcfb_edge_geometry(const float *xyz_in, const float *nrm_in, const float *xyz_out,
                                  const float *nrm_out, const int64_t *nei, int n_in, int n_out, int K,
                                  float *out_r, float *out_vi, void *stream)
{
    using namespace pcfb;
    PCFB_REQUIRE(n_in >= 1 && n_out >= 0 && K >= 1, "pcfb_edge_geometry: bad sizes");
    if (n_out == 0) return PCFB_OK;
    PCFB_REQUIRE(xyz_in && xyz_out && nei, "pcfb_edge_geometry: null pointer");
    PCFB_REQUIRE(!out_vi || (nrm_in && nrm_out), "pcfb_edge_geometry: VI features need normals");
    PCFB_REQUIRE(!out_vi || ((uintptr_t)out_vi % 16 == 0), "pcfb_edge_geometry: out_vi must be 16-byte aligned");
    const int64_t E = (int64_t)n_out * K;
    int64_t blocks = (E + 255) / 256;
    if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
    launch_k(edge_geometry_kernel, (int)blocks, 256, 0, static_cast<cudaStream_t>(stream), xyz_in, nrm_in, xyz_out, nrm_out, nei, n_in, E, K, out_r, out_vi);
    return check_launch("pcfb_edge_geometry");
}
