// Activation codes of the C ABI (include/pcf_b200.h: act = 0 none, 1 ReLU, 2 LeakyReLU(0.1), 3 sigmoid) and their
// forward / derivative, shared by the chain kernels (mlp.cu) and the one-kernel BatchNorm of small tensors (peer_reduce.cu).
#pragma once

namespace pcfb {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_SIGMOID = 3 };

__device__ __forceinline__ float act_fwd(float z, int act) {
    if (act == ACT_RELU) return fmaxf(z, 0.f);
    if (act == ACT_LEAKY) return z > 0.f ? z : 0.1f * z;
    if (act == ACT_SIGMOID) return 1.f / (1.f + __expf(-z));
    return z;
}
// derivative given pre-activation z and activation value a
__device__ __forceinline__ float act_bwd(float z, float a, int act) {
    if (act == ACT_RELU) return z > 0.f ? 1.f : 0.f;
    if (act == ACT_LEAKY) return z > 0.f ? 1.f : 0.1f;
    if (act == ACT_SIGMOID) return a * (1.f - a);
    return 1.f;
}

}  // namespace pcfb
