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

}  // namespace pcfb
