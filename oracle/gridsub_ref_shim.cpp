// extern "C" shim around the UNMODIFIED reference grid_subsampling()
// (/root/reference/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:9-110),
// compiled from the reference sources where they lie (oracle/Makefile `make ref`) into
// oracle/_ref/libgridsub_ref.so.  TEST INFRASTRUCTURE: validates oracle/grid_subsample.py and
// generates tests/golden/grid_subsample_*.npz.  The reference's CPython wrapper.cpp does not compile
// against numpy 2.x, hence this 30-line C ABI instead.
#include "grid_subsampling.h"
#include <cstring>

extern "C" int gridsub_ref(const float *pts, const float *feats, int n, int fdim, float dl,
                           float *out_pts, float *out_feats, int max_out)
{
    std::vector<PointXYZ> in_p(n), out_p;
    for (int i = 0; i < n; ++i) in_p[i] = PointXYZ(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    std::vector<float> in_f, out_f;
    if (fdim > 0) in_f.assign(feats, feats + (size_t)n * fdim);
    std::vector<int> in_c, out_c;
    grid_subsampling(in_p, out_p, in_f, out_f, in_c, out_c, dl, 0);
    int m = (int)out_p.size();
    if (m > max_out) return -m;
    for (int i = 0; i < m; ++i) { out_pts[3 * i] = out_p[i].x; out_pts[3 * i + 1] = out_p[i].y; out_pts[3 * i + 2] = out_p[i].z; }
    if (fdim > 0) std::memcpy(out_feats, out_f.data(), sizeof(float) * (size_t)m * fdim);
    return m;
}
