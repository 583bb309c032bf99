// Layer glue that the reference leaves to chains of torch elementwise ops (sm_100a, HBM-bound streaming kernels):
//
//  * guidance input of the PointConvFormer layer (/root/reference/layers.py:372-382):
//        q = cat(index_points(guidance_x, nei), feat_pe) ; key = q[:, :, :1] if M == N else max_k q ; s = q - key
//    (gather + cat + max + expand + sub = 5 HBM passes over a [M, K, 64] tensor in torch) as ONE pass, and its backward
//    (sub / max / cat / gather backward) as ONE pass that emits the per-edge gradient of the gathered half (summed per
//    input point by pcfb_gather_backward through the kNN inverse map, no atomics) and the gradient of feat_pe.
//  * gradient clipping + AdamW on the flat parameter buffer (/root/reference/train_ScanNet_DDP_WarmUP.py:421-424:
//    clip_grad_norm_(10) + optimizer.step()): two launches, fixed-order reductions.
//  * cross entropy with label smoothing / class weights / ignore_index and its gradient
//    (train_ScanNet_DDP_WarmUP.py:417, torch.nn.CrossEntropyLoss semantics).
#include "common.cuh"

namespace pcfb {

static inline int glue_grid(int64_t work, int threads) {
    int64_t b = (work + threads - 1) / threads;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

// thread = (output point m, float4 channel group c4 of the G + P channels); K neighbours walked twice (key, then q - key)
__global__ void __launch_bounds__(256)
guidance_input_kernel(const float4 *__restrict__ gx, const float4 *__restrict__ pe, const int64_t *__restrict__ nei, int n_in,
                      int64_t M, int K, int G4, int P4, int use_max, float4 *__restrict__ out, uchar4 *__restrict__ arg)
{
    pdl_wait();
    const int C4 = G4 + P4;
    const int64_t total = M * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / C4;
        const int c = (int)(i - m * C4);
        const bool from_gx = c < G4;
        auto load = [&](int k) -> float4 {
            if (from_gx) {
                const int64_t p = __ldg(nei + m * K + k);
                if (p < 0 || p >= n_in) return make_float4(0.f, 0.f, 0.f, 0.f);     // padding: index_points gives zeros
                return __ldg(gx + p * G4 + c);
            }
            return __ldg(pe + (m * K + k) * P4 + (c - G4));
        };
        float4 key = load(0);
        uchar4 ak = make_uchar4(0, 0, 0, 0);
        if (use_max) {
            for (int k = 1; k < K; ++k) {                                           // strict >: the first maximum wins
                const float4 v = load(k);
                if (v.x > key.x) { key.x = v.x; ak.x = (unsigned char)k; }
                if (v.y > key.y) { key.y = v.y; ak.y = (unsigned char)k; }
                if (v.z > key.z) { key.z = v.z; ak.z = (unsigned char)k; }
                if (v.w > key.w) { key.w = v.w; ak.w = (unsigned char)k; }
            }
            if (arg) arg[i] = ak;
        }
        for (int k = 0; k < K; ++k) out[(m * K + k) * C4 + c] = f4_sub(load(k), key);
    }
}

// dq[m,k,c] = ds[m,k,c] - [k == key(m,c)] * sum_k' ds[m,k',c] ; channels [0,G) -> d_gq (per edge), [G, G+P) -> d_pe
__global__ void __launch_bounds__(256)
guidance_input_bwd_kernel(const float4 *__restrict__ ds, const uchar4 *__restrict__ arg, int64_t M, int K, int G4, int P4,
                          int use_max, float4 *__restrict__ d_gq, float4 *__restrict__ d_pe)
{
    pdl_wait();
    const int C4 = G4 + P4;
    const int64_t total = M * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / C4;
        const int c = (int)(i - m * C4);
        float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < K; ++k) {
            const float4 v = __ldg(ds + (m * K + k) * C4 + c);
            tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
        }
        const uchar4 ak = use_max ? arg[i] : make_uchar4(0, 0, 0, 0);
        float4 *dst = c < G4 ? d_gq : d_pe;
        const int W4 = c < G4 ? G4 : P4, cc = c < G4 ? c : c - G4;
        if (!dst) continue;
        for (int k = 0; k < K; ++k) {
            float4 v = __ldg(ds + (m * K + k) * C4 + c);
            if (k == ak.x) v.x -= tot.x;
            if (k == ak.y) v.y -= tot.y;
            if (k == ak.z) v.z -= tot.z;
            if (k == ak.w) v.w -= tot.w;
            dst[(m * K + k) * W4 + cc] = v;
        }
    }
}

// ---- gradient clipping + AdamW on one flat buffer ---------------------------------------------------------------------
constexpr int OPT_THREADS = 256;
constexpr int OPT_MAX_BLOCKS = 2 * kNumSMs;

// partial[b] = sum of g^2 over block b's grid-stride slice (double); block 0 also advances the step counter
__global__ void __launch_bounds__(OPT_THREADS)
sqnorm_partials_kernel(const float *__restrict__ g, int64_t n, double *__restrict__ partial, float *__restrict__ step)
{
    pdl_wait();
    __shared__ double red[OPT_THREADS / 32];
    double s = 0.0;
    const int64_t n4 = n >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g4 + i);
        s += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; s += (double)v * v; }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < OPT_THREADS / 32; ++w) t += red[w];
        partial[blockIdx.x] = t;
        if (blockIdx.x == 0 && step) *step += 1.f;
    }
}

struct AdamArgs {
    float *p; const float *g; float *m, *v; int64_t n;
    const double *partial; int n_partial;
    const float *lr, *step;          // device scalars: the learning rate can change between replays of a captured step
    float beta1, beta2, eps, weight_decay, max_norm;
    float *norm_out;                 // optional: total gradient norm before clipping
};

__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, float coef, float decay, float b1, float b2,
                                         float step_size, float inv_sqrt_bc2, float eps)
{
    g *= coef;
    p *= decay;
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
}

__global__ void __launch_bounds__(OPT_THREADS)
adamw_clip_kernel(AdamArgs a)
{
    pdl_wait();
    __shared__ double tot_s;
    if (threadIdx.x < 32) {                     // every block sums the same partials in the same order: same coefficient
        double t = 0.0;
        for (int b = threadIdx.x; b < a.n_partial; b += 32) t += a.partial[b];
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) t += __shfl_xor_sync(0xffffffffu, t, sft);
        if (threadIdx.x == 0) tot_s = t;
    }
    __syncthreads();
    const float norm = (float)sqrt(tot_s);
    if (a.norm_out && blockIdx.x == 0 && threadIdx.x == 0) *a.norm_out = norm;
    float coef = 1.f;
    if (a.max_norm > 0.f) { coef = a.max_norm / (norm + 1e-6f); if (coef > 1.f) coef = 1.f; }
    const float lr = *a.lr, t = *a.step;
    const float bc1 = 1.f - powf(a.beta1, t), bc2 = 1.f - powf(a.beta2, t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * a.weight_decay;
    const int64_t n4 = a.n >> 2;
    float4 *p4 = reinterpret_cast<float4 *>(a.p), *m4 = reinterpret_cast<float4 *>(a.m), *v4 = reinterpret_cast<float4 *>(a.v);
    const float4 *g4 = reinterpret_cast<const float4 *>(a.g);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = p4[i], m = m4[i], v = v4[i];
        const float4 g = __ldg(g4 + i);
        adam_one(p.x, g.x, m.x, v.x, coef, decay, a.beta1, a.beta2, step_size, inv_sqrt_bc2, a.eps);
        adam_one(p.y, g.y, m.y, v.y, coef, decay, a.beta1, a.beta2, step_size, inv_sqrt_bc2, a.eps);
        adam_one(p.z, g.z, m.z, v.z, coef, decay, a.beta1, a.beta2, step_size, inv_sqrt_bc2, a.eps);
        adam_one(p.w, g.w, m.w, v.w, coef, decay, a.beta1, a.beta2, step_size, inv_sqrt_bc2, a.eps);
        p4[i] = p; m4[i] = m; v4[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(a.n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        adam_one(a.p[i], a.g[i], a.m[i], a.v[i], coef, decay, a.beta1, a.beta2, step_size, inv_sqrt_bc2, a.eps);
    }
}

// ---- cross entropy (label smoothing, class weights, ignore_index), thread = row ------------------------------------------
constexpr int CE_MAXC = 64;
constexpr int CE_THREADS = 128;

struct CeRow { float mx, lse; };

template <bool BWD>
__global__ void __launch_bounds__(CE_THREADS)
ce_kernel(const float *__restrict__ logits, const int64_t *__restrict__ labels, const float *__restrict__ weight, int64_t N, int C,
          int64_t ignore_index, float smoothing, double *__restrict__ partial /* fwd: [blocks][2] */,
          const float *__restrict__ den /* bwd: sum of weights of the valid rows */, const float *__restrict__ gscale,
          float *__restrict__ dlogits)
{
    pdl_wait();
    __shared__ float w_s[CE_MAXC];
    __shared__ float wsum_s;
    __shared__ double red[CE_THREADS / 32][2];
    if (threadIdx.x < C) w_s[threadIdx.x] = weight ? weight[threadIdx.x] : 1.f;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int c = 0; c < C; ++c) t += w_s[c]; wsum_s = t; }
    __syncthreads();
    const float wsum = wsum_s, eps_c = smoothing / (float)C;
    float gmul = 0.f;
    if (BWD) gmul = (gscale ? *gscale : 1.f) / *den;
    double loss = 0.0, wacc = 0.0;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < N; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t y = labels[r];
        const bool valid = y != ignore_index && y >= 0 && y < C;
        const float *row = logits + r * C;
        if (!valid) {
            if (BWD) for (int c = 0; c < C; ++c) dlogits[r * C + c] = 0.f;
            continue;
        }
        float v[CE_MAXC];
        float mx = -__int_as_float(0x7f800000);
#pragma unroll 4
        for (int c = 0; c < C; ++c) { v[c] = __ldg(row + c); mx = fmaxf(mx, v[c]); }
        float se = 0.f;
#pragma unroll 4
        for (int c = 0; c < C; ++c) se += __expf(v[c] - mx);
        const float lse = mx + __logf(se);
        const float wy = w_s[(int)y];
        if (!BWD) {
            float sm = 0.f;
            for (int c = 0; c < C; ++c) sm += w_s[c] * (lse - v[c]);
            loss += (double)((1.f - smoothing) * wy * (lse - v[(int)y]) + eps_c * sm);
            wacc += (double)wy;
        } else {
            for (int c = 0; c < C; ++c) {
                const float p = __expf(v[c] - lse);
                const float d = (1.f - smoothing) * wy * (p - (c == (int)y ? 1.f : 0.f)) + eps_c * (p * wsum - w_s[c]);
                dlogits[r * C + c] = gmul * d;
            }
        }
    }
    if (!BWD) {
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) { loss += __shfl_xor_sync(0xffffffffu, loss, sft); wacc += __shfl_xor_sync(0xffffffffu, wacc, sft); }
        if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = loss; red[threadIdx.x >> 5][1] = wacc; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < CE_THREADS / 32; ++w) { a += red[w][0]; b += red[w][1]; }
            partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b;
        }
    }
}

__global__ void ce_finalize_kernel(const double *__restrict__ partial, int nblocks, float *__restrict__ loss, float *__restrict__ den)
{
    pdl_wait();
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) { a += partial[2 * i]; b += partial[2 * i + 1]; }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, sft); b += __shfl_xor_sync(0xffffffffu, b, sft); }
    if (threadIdx.x == 0) { *loss = (float)(a / b); *den = (float)b; }
}

// eval-mode BatchNorm as a per-channel affine map: scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale
// (torch: rsqrt, two multiplies, a subtraction and a contiguous copy -- five launches per BatchNorm and forward)
__global__ void bn_eval_affine_kernel(const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ rm,
                                      const float *__restrict__ rv, float eps, int C, float *__restrict__ scale,
                                      float *__restrict__ shift, float *__restrict__ invstd)
{
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float is = rsqrtf(rv[c] + eps);
    const float sc = gamma ? gamma[c] * is : is;
    scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.f) - rm[c] * sc;
    if (invstd) invstd[c] = is;
}

}  // namespace pcfb

using namespace pcfb;

extern "C" int pcfb_guidance_input(const float *gx, const float *pe, const int64_t *nei, int n_in, int n_out, int K, int G, int P,
                                   int use_max, float *out, uint8_t *arg, void *stream)
{
    PCFB_REQUIRE(n_in >= 0 && n_out >= 0 && K >= 1 && K <= 255 && G >= 0 && P >= 0 && G + P >= 4 && (G & 3) == 0 && (P & 3) == 0,
                 "pcfb_guidance_input: bad sizes (G = %d, P = %d must be multiples of 4)", G, P);
    if (n_out == 0) return PCFB_OK;
    PCFB_REQUIRE((G == 0 || gx) && (P == 0 || pe) && nei && out && (!use_max || arg), "pcfb_guidance_input: null pointer");
    PCFB_REQUIRE(((uintptr_t)gx % 16 == 0) && ((uintptr_t)pe % 16 == 0) && ((uintptr_t)out % 16 == 0), "pcfb_guidance_input: misaligned pointer");
    const int64_t work = (int64_t)n_out * ((G + P) / 4);
    launch_k(guidance_input_kernel, glue_grid(work, 256), 256, 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const float4 *>(gx), reinterpret_cast<const float4 *>(pe), nei, n_in, n_out, K, G / 4, P / 4, use_max,
        reinterpret_cast<float4 *>(out), reinterpret_cast<uchar4 *>(arg));
    return check_launch("guidance_input_kernel");
}

extern "C" int pcfb_guidance_input_backward(const float *ds, const uint8_t *arg, int n_out, int K, int G, int P, int use_max,
                                            float *d_gq, float *d_pe, void *stream)
{
    PCFB_REQUIRE(n_out >= 0 && K >= 1 && K <= 255 && (G & 3) == 0 && (P & 3) == 0 && G + P >= 4, "pcfb_guidance_input_backward: bad sizes");
    if (n_out == 0) return PCFB_OK;
    PCFB_REQUIRE(ds && (!use_max || arg), "pcfb_guidance_input_backward: null pointer");
    PCFB_REQUIRE(((uintptr_t)ds % 16 == 0) && ((uintptr_t)d_gq % 16 == 0) && ((uintptr_t)d_pe % 16 == 0), "pcfb_guidance_input_backward: misaligned pointer");
    const int64_t work = (int64_t)n_out * ((G + P) / 4);
    launch_k(guidance_input_bwd_kernel, glue_grid(work, 256), 256, 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const float4 *>(ds), reinterpret_cast<const uchar4 *>(arg), n_out, K, G / 4, P / 4, use_max,
        reinterpret_cast<float4 *>(d_gq), reinterpret_cast<float4 *>(d_pe));
    return check_launch("guidance_input_bwd_kernel");
}

extern "C" size_t pcfb_adamw_workspace(void) { return (size_t)OPT_MAX_BLOCKS * sizeof(double); }

extern "C" int pcfb_adamw_clip_step(float *p, const float *g, float *m, float *v, int64_t n, const float *lr, float *step,
                                    float beta1, float beta2, float eps, float weight_decay, float max_norm, float *norm_out,
                                    void *workspace, size_t workspace_bytes, void *stream)
{
    PCFB_REQUIRE(p && g && m && v && lr && step && workspace && n >= 0, "pcfb_adamw_clip_step: null pointer");
    PCFB_REQUIRE(workspace_bytes >= pcfb_adamw_workspace(), "pcfb_adamw_clip_step: workspace too small");
    PCFB_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) && ((uintptr_t)v % 16 == 0) &&
                 ((uintptr_t)workspace % 8 == 0), "pcfb_adamw_clip_step: misaligned pointer");
    if (n == 0) return PCFB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int64_t b = ((n >> 2) + OPT_THREADS - 1) / OPT_THREADS;
    const int blocks = (int)(b < 1 ? 1 : (b > OPT_MAX_BLOCKS ? OPT_MAX_BLOCKS : b));
    double *partial = static_cast<double *>(workspace);
    launch_k(sqnorm_partials_kernel, blocks, OPT_THREADS, 0, st, g, n, partial, step);
    int rc = check_launch("sqnorm_partials_kernel");
    if (rc) return rc;
    AdamArgs a{p, g, m, v, n, partial, blocks, lr, step, beta1, beta2, eps, weight_decay, max_norm, norm_out};
    launch_k(adamw_clip_kernel, blocks, OPT_THREADS, 0, st, a);
    return check_launch("adamw_clip_kernel");
}

extern "C" size_t pcfb_ce_workspace(int64_t N)
{
    int64_t b = (N + CE_THREADS - 1) / CE_THREADS;
    if (b > (int64_t)kNumSMs * 8) b = (int64_t)kNumSMs * 8;
    return (size_t)(b < 1 ? 1 : b) * 2 * sizeof(double);
}

static int ce_blocks(int64_t N)
{
    int64_t b = (N + CE_THREADS - 1) / CE_THREADS;
    if (b > (int64_t)kNumSMs * 8) b = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : b);
}

extern "C" int pcfb_ce_forward(const float *logits, const int64_t *labels, const float *weight, int64_t N, int C, int64_t ignore_index,
                               float smoothing, float *loss, float *den, void *workspace, size_t workspace_bytes, void *stream)
{
    PCFB_REQUIRE(logits && labels && loss && den && workspace && N >= 0 && C >= 1 && C <= CE_MAXC, "pcfb_ce_forward: bad arguments (C <= %d)", CE_MAXC);
    PCFB_REQUIRE(workspace_bytes >= pcfb_ce_workspace(N), "pcfb_ce_forward: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = ce_blocks(N);
    double *partial = static_cast<double *>(workspace);
    launch_k(ce_kernel<false>, blocks, CE_THREADS, 0, st, logits, labels, weight, N, C, ignore_index, smoothing, partial, nullptr, nullptr, nullptr);
    int rc = check_launch("ce_kernel<fwd>");
    if (rc) return rc;
    launch_k(ce_finalize_kernel, 1, 32, 0, st, partial, blocks, loss, den);
    return check_launch("ce_finalize_kernel");
}

extern "C" int pcfb_ce_backward(const float *logits, const int64_t *labels, const float *weight, int64_t N, int C, int64_t ignore_index,
                                float smoothing, const float *den, const float *grad_scale, float *dlogits, void *stream)
{
    PCFB_REQUIRE(logits && labels && den && dlogits && N >= 0 && C >= 1 && C <= CE_MAXC, "pcfb_ce_backward: bad arguments");
    if (N == 0) return PCFB_OK;
    launch_k(ce_kernel<true>, ce_blocks(N), CE_THREADS, 0, static_cast<cudaStream_t>(stream), logits, labels, weight, N, C, ignore_index, smoothing,
                                                                                       nullptr, den, grad_scale, dlogits);
    return check_launch("ce_kernel<bwd>");
}

extern "C" int pcfb_bn_eval_affine(const float *gamma, const float *beta, const float *running_mean, const float *running_var,
                                   float eps, int C, float *scale, float *shift, float *invstd, void *stream)
{
    PCFB_REQUIRE(running_mean && running_var && scale && shift && C >= 1, "pcfb_bn_eval_affine: null pointer");
    launch_k(bn_eval_affine_kernel, ceil_div(C, 128), 128, 0, static_cast<cudaStream_t>(stream), gamma, beta, running_mean, running_var, eps, C,
             scale, shift, invstd);
    return check_launch("bn_eval_affine_kernel");
}
