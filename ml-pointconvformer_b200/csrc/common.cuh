// Shared helpers for libpcf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "pcf_b200.h"

namespace pcfb {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return PCFB_ERR_CUDA;
    }
    return PCFB_OK;
}

#define PCFB_REQUIRE(cond, ...)                      \
    do {                                             \
        if (!(cond)) {                               \
            pcfb::set_error(__VA_ARGS__);            \
            return PCFB_ERR_ARG;                     \
        }                                            \
    } while (0)

#define PCFB_CUDA(call)                                                     \
    do {                                                                    \
        cudaError_t e_ = (call);                                            \
        if (e_ != cudaSuccess) {                                            \
            pcfb::set_error("%s: %s", #call, cudaGetErrorString(e_));       \
            return PCFB_ERR_CUDA;                                           \
        }                                                                   \
    } while (0)

constexpr int kNumSMs = 148;   // B200

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------------
// A training step is a chain of ~1800 DEPENDENT kernels, most of them sub-wave; between two of them the GPU idled ~1.7 us
// (profiles/step_timeline_r02.txt: 3.6 ms of gaps on the critical path).  Every kernel of this library starts with
// pdl_wait() (griddepcontrol.wait: returns once all prerequisite grids have completed and their memory is visible) and is
// launched with programmaticStreamSerializationAllowed, so that its CTAs are already resident, waiting, when its stream
// predecessor drains -- the launch latency is paid while the predecessor runs.  Nothing is read or written before the
// wait, so the semantics are exactly those of an ordinary stream-ordered launch.  PCFB_PDL=0 turns the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors surface through check_launch()
}

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// carve typed arrays out of a caller-provided workspace
struct Carver {
    char *base;
    size_t off;
    explicit Carver(void *p) : base(static_cast<char *>(p)), off(0) {}
    template <typename T>
    T *take(size_t n) {
        off = align_up(off, 256);
        T *r = reinterpret_cast<T *>(base + off);
        off += n * sizeof(T);
        return r;
    }
};

// Per-scene bounding boxes: lanes of a warp that belong to the same scene combine their three ordered-int coordinates with
// warp reductions and ONE lane issues the six atomics -- packed scenes are contiguous, so this is one atomic per warp and
// coordinate instead of 32 to the same address (16 packed rooms: 305 k points hammering 96 integers took 0.6 ms per level,
// scripts/profile_infer.py).  `in`: this lane has a point; all 32 lanes of the warp must call.
__device__ __forceinline__ void warp_scene_minmax(bool in, int s, int ox, int oy, int oz, int *__restrict__ mm)
{
    const unsigned act = __ballot_sync(0xffffffffu, in);
    if (!in) return;
    const unsigned grp = __match_any_sync(act, s);
    const bool leader = (int)(threadIdx.x & 31) == __ffs(grp) - 1;
    const int o[3] = {ox, oy, oz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int mn = __reduce_min_sync(grp, o[d]), mx = __reduce_max_sync(grp, o[d]);
        if (leader) { atomicMin(&mm[s * 6 + d], mn); atomicMax(&mm[s * 6 + 3 + d], mx); }
    }
}

}  // namespace pcfb
