// Inference-mode per-edge MLP chain as ONE kernel (sm_100a): out = act3(s3 * (W3 act2(s2 * (W2 act1(s1 * (W1 x + b1) + t1) + b2) + t2) + b3) + t3).
//
// The WeightNet of every PointConv / PointConvFormer layer (/root/reference/layers.py:127-171: Linear_BN + ReLU x 3 on the
// per-edge geometry features, 12 -> 8 -> 8 -> 16 with viewpoint-invariant features, 3 -> 8 -> 8 -> 16 without) is a pure
// function of its input row once its BatchNorms use running statistics (model.eval(); folded into the Linear by
// replace_batchnorm in the reference's test drivers, test_ScanNet_simple.py:139-141): no statistics pass separates the
// layers.  The training path (mlp.cu) runs one kernel per layer plus a final BatchNorm + activation pass:
// 80 + 64 + 96 + 128 = 368 B per edge and four launches; here a thread keeps its row in registers through the three
// layers: 48 + 64 = 112 B per edge, one launch.  This is the first slice of the module-level fusion:
// the chain no longer writes its intermediates; the step after it is evaluating it inside the contraction kernel.
// Weights / biases / affine maps sit in shared memory (every lane reads the same element: broadcast loads).
#include "common.cuh"
#include "act.cuh"

namespace pcfb {

constexpr int ME_THREADS = 256;

struct ChainEvalArgs {
    const float *W[3], *b[3], *scale[3], *shift[3];   // scale / shift null: identity (BatchNorm folded or absent)
    int act[3];
    int c0;                                           // real input width (<= C0; e.g. 3 of 4)
};

template <int C0, int C1, int C2, int C3>
__global__ void __launch_bounds__(ME_THREADS)
mlp_chain3_eval_kernel(const float *__restrict__ x, int ldx, int64_t E, ChainEvalArgs a, float *__restrict__ out, int ldo)
{
    pdl_wait();
    __shared__ float W1_s[C1 * C0], W2_s[C2 * C1], W3_s[C3 * C2];
    __shared__ float b_s[C1 + C2 + C3], sc_s[C1 + C2 + C3], sh_s[C1 + C2 + C3];
    const int t = threadIdx.x;
    for (int i = t; i < C1 * C0; i += ME_THREADS) { const int o = i / C0, k = i - o * C0; W1_s[i] = k < a.c0 ? a.W[0][o * a.c0 + k] : 0.f; }
    for (int i = t; i < C2 * C1; i += ME_THREADS) W2_s[i] = a.W[1][i];
    for (int i = t; i < C3 * C2; i += ME_THREADS) W3_s[i] = a.W[2][i];
    for (int i = t; i < C1 + C2 + C3; i += ME_THREADS) {
        const int l = i < C1 ? 0 : (i < C1 + C2 ? 1 : 2), o = i - (l == 0 ? 0 : (l == 1 ? C1 : C1 + C2));
        b_s[i] = a.b[l] ? a.b[l][o] : 0.f;
        sc_s[i] = a.scale[l] ? a.scale[l][o] : 1.f;
        sh_s[i] = a.shift[l] ? a.shift[l][o] : 0.f;
    }
    __syncthreads();
    const bool vec_in = (a.c0 == C0) && ((C0 & 3) == 0) && ((ldx & 3) == 0) && ((uintptr_t)x % 16 == 0);
    const bool vec_out = ((ldo & 3) == 0) && ((uintptr_t)out % 16 == 0);
    for (int64_t row = (int64_t)blockIdx.x * ME_THREADS + t; row < E; row += (int64_t)gridDim.x * ME_THREADS) {
        float v0[C0];
        const float *xr = x + row * ldx;
        if (vec_in) {
#pragma unroll
            for (int k = 0; k < C0; k += 4) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(xr + k));
                v0[k] = q.x; v0[k + 1] = q.y; v0[k + 2] = q.z; v0[k + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < C0; ++k) v0[k] = k < a.c0 ? __ldg(xr + k) : 0.f;
        }
        float v1[C1];
#pragma unroll
        for (int o = 0; o < C1; ++o) {
            float acc = b_s[o];
#pragma unroll
            for (int k = 0; k < C0; ++k) acc = fmaf(W1_s[o * C0 + k], v0[k], acc);
            v1[o] = act_fwd(fmaf(acc, sc_s[o], sh_s[o]), a.act[0]);
        }
        float v2[C2];
#pragma unroll
        for (int o = 0; o < C2; ++o) {
            float acc = b_s[C1 + o];
#pragma unroll
            for (int k = 0; k < C1; ++k) acc = fmaf(W2_s[o * C1 + k], v1[k], acc);
            v2[o] = act_fwd(fmaf(acc, sc_s[C1 + o], sh_s[C1 + o]), a.act[1]);
        }
        float v3[C3];
#pragma unroll
        for (int o = 0; o < C3; ++o) {
            float acc = b_s[C1 + C2 + o];
#pragma unroll
            for (int k = 0; k < C2; ++k) acc = fmaf(W3_s[o * C2 + k], v2[k], acc);
            v3[o] = act_fwd(fmaf(acc, sc_s[C1 + C2 + o], sh_s[C1 + C2 + o]), a.act[2]);
        }
        float *orow = out + row * ldo;
        if (vec_out) {
#pragma unroll
            for (int o = 0; o < C3; o += 4) *reinterpret_cast<float4 *>(orow + o) = make_float4(v3[o], v3[o + 1], v3[o + 2], v3[o + 3]);
        } else {
#pragma unroll
            for (int o = 0; o < C3; ++o) orow[o] = v3[o];
        }
    }
}

}  // namespace pcfb

using namespace pcfb;

// the chains with a fused inference kernel: (c0 <= 4 | c0 == 12) -> 8 -> 8 -> 16 (WeightNet without / with VI features)
extern "C" int pcfb_mlp_chain_eval_supported(int c0, int c1, int c2, int c3)
{
    return (c1 == 8 && c2 == 8 && c3 == 16 && ((c0 >= 1 && c0 <= 4) || c0 == 12)) ? 1 : 0;
}

extern "C" int pcfb_mlp_chain_eval(const float *x, int ldx, int64_t E, int c0, int c1, int c2, int c3,
                                   const float *const *W, const float *const *b, const float *const *scale,
                                   const float *const *shift, const int *act, float *out, int ldo, void *stream)
{
    PCFB_REQUIRE(pcfb_mlp_chain_eval_supported(c0, c1, c2, c3), "pcfb_mlp_chain_eval: unsupported chain %d -> %d -> %d -> %d", c0, c1, c2, c3);
    PCFB_REQUIRE(x && out && W && act && W[0] && W[1] && W[2] && E >= 0 && ldx >= c0 && ldo >= c3, "pcfb_mlp_chain_eval: null pointer or bad leading dimension");
    if (E == 0) return PCFB_OK;
    ChainEvalArgs a{};
    for (int l = 0; l < 3; ++l) {
        a.W[l] = W[l]; a.b[l] = b ? b[l] : nullptr;
        a.scale[l] = scale ? scale[l] : nullptr; a.shift[l] = shift ? shift[l] : nullptr;
        PCFB_REQUIRE((a.scale[l] == nullptr) == (a.shift[l] == nullptr), "pcfb_mlp_chain_eval: scale and shift go together");
        a.act[l] = act[l];
    }
    a.c0 = c0;
    int64_t blocks = (E + ME_THREADS - 1) / ME_THREADS;
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (c0 == 12) launch_k(mlp_chain3_eval_kernel<12, 8, 8, 16>, (int)blocks, ME_THREADS, 0, st, x, ldx, E, a, out, ldo);
    else launch_k(mlp_chain3_eval_kernel<4, 8, 8, 16>, (int)blocks, ME_THREADS, 0, st, x, ldx, E, a, out, ldo);
    return check_launch("mlp_chain3_eval_kernel");
}
