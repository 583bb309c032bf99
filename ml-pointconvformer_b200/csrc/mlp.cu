// Fused per-edge / per-point MLP layers with train-mode BatchNorm (sm_100a).
//
// Covers SURVEY.md rows a8 / a9 / a13: WeightNet (Linear -> BN2d(batch stats) -> ReLU chains,
// /root/reference/layers.py:127-191), the positional-encoding WeightNet (layers.py:575-577), mlp_conv
// (layers.py:240-243) and the guidance MLP (layers.py:38-68), which the reference (and torch) run as ~5 full
// HBM passes per layer (Linear, BN statistics, BN apply, ReLU, ...) over [E, C] tensors with E = 1.6 M edges.
//
// Here one layer is ONE streaming pass:  y = act_in(x * scale_in + shift_in) W^T + b  with the BatchNorm of the
// *previous* layer folded into the load (scale/shift) and the batch statistics of y accumulated on the fly
// (pivoted by the bias, block partials reduced in double by a finalize kernel -> deterministic).  The BatchNorm of
// y itself is only a (scale, shift) pair that the next consumer applies.  Backward per layer: one pass producing
// the input gradient (with the BN-backward sums of the layer below accumulated on the fly) and one pass producing
// the weight / bias gradient (warps own output channels, rows live in lanes, fixed-order reductions, no atomics).
// One thread owns one row; weights sit in shared memory, zero padded to the template maxima.
#include "common.cuh"
#include "act.cuh"

namespace pcfb {

constexpr int ML_THREADS = 256;
constexpr int ML_WARPS = ML_THREADS / 32;


template <int CMAX>
__device__ __forceinline__ void load_row(const float *__restrict__ base, int ld, int64_t row, int c, bool vec, float *out) {
    const float *p = base + row * ld;
    if (vec) {
#pragma unroll
        for (int i = 0; i < CMAX; i += 4) {
            if (i < c) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(p + i));
                out[i] = v.x; out[i + 1] = v.y; out[i + 2] = v.z; out[i + 3] = v.w;
            } else { out[i] = out[i + 1] = out[i + 2] = out[i + 3] = 0.f; }
        }
    } else {
#pragma unroll
        for (int i = 0; i < CMAX; ++i) out[i] = (i < c) ? __ldg(p + i) : 0.f;
    }
}

template <int CMAX>
__device__ __forceinline__ void store_row(float *__restrict__ base, int ld, int64_t row, int c, bool vec, const float *v) {
    float *p = base + row * ld;
    if (vec) {
#pragma unroll
        for (int i = 0; i < CMAX; i += 4)
            if (i < c) *reinterpret_cast<float4 *>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < CMAX; ++i) if (i < c) p[i] = v[i];
    }
}

struct MlpLayer {            // everything one layer needs; pointers may be null where noted
    const float *W, *b;      // [cout][cin], [cout] (b may be null)
    int cin, cout;
    const float *in_scale, *in_shift;   // folded BN of the layer below applied on load (null = identity)
    int in_act;
};

// ------------------------------------------------------------------------------------------------------------
// forward: y = act_in(x*s+t) W^T + b ; stats partial[block][2][cout] = sum (y-b), sum (y-b)^2
// ------------------------------------------------------------------------------------------------------------
template <int CIN, int COUT>
__global__ void __launch_bounds__(ML_THREADS)
mlp_fwd_kernel(const float *__restrict__ x, int ldx, int64_t E, MlpLayer L, float *__restrict__ y, int ldy,
               float *__restrict__ stat_partial, int rows_per_thread)
{
    pdl_wait();
    __shared__ __align__(16) float W_s[COUT * CIN];
    __shared__ float b_s[COUT], s_s[CIN], t_s[CIN];
    __shared__ float red[ML_WARPS][2 * COUT];
    for (int i = threadIdx.x; i < COUT * CIN; i += ML_THREADS) {
        const int o = i / CIN, k = i - o * CIN;
        W_s[i] = (o < L.cout && k < L.cin) ? L.W[o * L.cin + k] : 0.f;
    }
    for (int i = threadIdx.x; i < COUT; i += ML_THREADS) b_s[i] = (L.b && i < L.cout) ? L.b[i] : 0.f;
    for (int i = threadIdx.x; i < CIN; i += ML_THREADS) {
        s_s[i] = (L.in_scale && i < L.cin) ? L.in_scale[i] : 1.f;
        t_s[i] = (L.in_shift && i < L.cin) ? L.in_shift[i] : 0.f;
    }
    __syncthreads();
    const bool vec_in = ((ldx & 3) == 0) && ((L.cin & 3) == 0) && ((uintptr_t)x % 16 == 0);
    const bool vec_out = ((ldy & 3) == 0) && ((L.cout & 3) == 0) && ((uintptr_t)y % 16 == 0);
    float s1[COUT], s2[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) { s1[o] = 0.f; s2[o] = 0.f; }
    const int64_t block_row0 = (int64_t)blockIdx.x * ML_THREADS * rows_per_thread;
    for (int r = 0; r < rows_per_thread; ++r) {
        const int64_t row = block_row0 + (int64_t)r * ML_THREADS + threadIdx.x;
        if (row >= E) break;
        float a[CIN];
        load_row<CIN>(x, ldx, row, L.cin, vec_in, a);
#pragma unroll
        for (int k = 0; k < CIN; ++k) a[k] = act_fwd(fmaf(a[k], s_s[k], t_s[k]), L.in_act);
        float acc[COUT];
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < CIN; k += 4) {
                const float4 w = *reinterpret_cast<const float4 *>(&W_s[o * CIN + k]);
                v = fmaf(a[k], w.x, v); v = fmaf(a[k + 1], w.y, v); v = fmaf(a[k + 2], w.z, v); v = fmaf(a[k + 3], w.w, v);
            }
            acc[o] = v;
            s1[o] += v; s2[o] = fmaf(v, v, s2[o]);
        }
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] += b_s[o];
        store_row<COUT>(y, ldy, row, L.cout, vec_out, acc);
    }
    if (stat_partial) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
            float a1 = s1[o], a2 = s2[o];
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, sft); a2 += __shfl_xor_sync(0xffffffffu, a2, sft); }
            if (lane == 0) { red[warp][o] = a1; red[warp][COUT + o] = a2; }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * COUT; i += ML_THREADS) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < ML_WARPS; ++w) v += red[w][i];
            const int which = i / COUT, o = i - which * COUT;
            if (o < L.cout) stat_partial[((size_t)blockIdx.x * 2 + which) * L.cout + o] = v;
        }
    }
}

// (BatchNorm finalize: bn_reduce_kernel in peer_reduce.cu -- partial sums -> [SyncBatchNorm exchange] -> scale / shift / running stats)

// a = act(y*scale+shift), elementwise (the chain's last BatchNorm + activation, materialised for its consumer)
// res (optional): a residual added before the activation (res_after == 0: act(bn(y) + r), the tail of a PointConvFormer
// block, layers.py:415) or after it (act(bn(y)) + r, the decoder's skip connection, layers.py:1096-1097)
__global__ void bn_act_kernel(const float *__restrict__ y, int64_t n, int C, const float *__restrict__ scale,
                              const float *__restrict__ shift, int act, float *__restrict__ out,
                              const float *__restrict__ res, int res_after)
{
    pdl_wait();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float z = scale ? fmaf(y[i], scale[c], shift[c]) : y[i];
        if (res && !res_after) z += res[i];
        float o = act_fwd(z, act);
        if (res && res_after) o += res[i];
        out[i] = o;
    }
}

__global__ void __launch_bounds__(256)
bn_act4_kernel(const float *__restrict__ y, int64_t rows, int C, const float *__restrict__ scale, const float *__restrict__ shift,
               int act, float *__restrict__ out, int c4, int ry, const float *__restrict__ res, int res_after);   // float4 form, defined below

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
struct MlpBnCtx {            // BatchNorm + activation of one layer's output y (train mode); no BN: scale = null
    const float *scale, *shift, *mean, *invstd;   // scale = gamma*invstd, shift = beta - mean*scale
    const float *sums;       // [2][C]: S1 = sum dz, S2 = sum dz * xhat   (null when no BN)
    int act;
    float inv_count;
    const double *d_count;   // optional device-side global row count (SyncBatchNorm): overrides inv_count
};

__device__ __forceinline__ float ctx_inv_count(const MlpBnCtx &B) { return B.d_count ? (float)(1.0 / *B.d_count) : B.inv_count; }

// slope of act'(z) for z <= 0 (ReLU 0, LeakyReLU 0.1, none 1): act(z) = z > 0 ? z : slope*z, act'(z) = z > 0 ? 1 : slope
__host__ __device__ inline float act_slope(int act) { return act == ACT_RELU ? 0.f : (act == ACT_LEAKY ? 0.1f : 1.f); }

// Per-channel constants of one layer's BatchNorm + activation backward, filled into shared memory once per CTA so that the
// row loops read broadcast LDS instead of six global loads per element:
//   z = y*sc + sh ;  xhat = (y - mu)*is ;  dy = sc*dz + c1*(y - mu) + c0  with c1 = -sc*S2*is/E, c0 = -sc*S1/E
// (no BatchNorm: sc = 1, the rest 0; padded channels: everything 0, so dy = 0 without a bounds test)
template <int C>
struct BnConst { float sc[C], sh[C], mu[C], is[C], c1[C], c0[C]; };

template <int C>
__device__ __forceinline__ void bn_const_fill(BnConst<C> &K, const MlpBnCtx &B, int c, float inv_count)
{
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        float sc = 0.f, sh = 0.f, mu = 0.f, is = 0.f, c1 = 0.f, c0 = 0.f;
        if (i < c) {
            sc = 1.f;
            if (B.scale) {
                sc = B.scale[i]; sh = B.shift[i];
                if (B.mean) { mu = B.mean[i]; is = B.invstd[i]; }
                if (B.sums) { c1 = -sc * B.sums[c + i] * inv_count * is; c0 = -sc * B.sums[i] * inv_count; }
            }
        }
        K.sc[i] = sc; K.sh[i] = sh; K.mu[i] = mu; K.is[i] = is; K.c1[i] = c1; K.c0[i] = c0;
    }
}

// dz = dA * act'(z) for the C channels of one row (SIG: sigmoid, else slope form); returns dz in d[]
template <int C, bool SIG>
__device__ __forceinline__ void bn_row_dz(float *d, const float *yv, const BnConst<C> &K, float slope)
{
#pragma unroll
    for (int o = 0; o < C; ++o) {
        const float z = fmaf(yv[o], K.sc[o], K.sh[o]);
        float gz;
        if (SIG) { const float av = 1.f / (1.f + __expf(-z)); gz = av * (1.f - av); }
        else gz = z > 0.f ? 1.f : slope;
        d[o] *= gz;
    }
}
// dy from dz (in place)
template <int C>
__device__ __forceinline__ void bn_row_dy(float *d, const float *yv, const BnConst<C> &K)
{
#pragma unroll
    for (int o = 0; o < C; ++o) d[o] = fmaf(K.sc[o], d[o], fmaf(K.c1[o], yv[o] - K.mu[o], K.c0[o]));
}

// stats of one layer's output gradient: partial[block][2][C] = sum dz, sum dz*xhat  (thread = row)
template <int COUT>
__global__ void __launch_bounds__(ML_THREADS)
mlp_bwd_stats_kernel(const float *__restrict__ dA, int ldd, const float *__restrict__ y, int ldy, int64_t E, int C,
                     MlpBnCtx B, float *__restrict__ partial, int rows_per_thread)
{
    pdl_wait();
    __shared__ float red[ML_WARPS][2 * COUT];
    __shared__ BnConst<COUT> K;
    bn_const_fill<COUT>(K, B, C, 0.f);
    __syncthreads();
    const bool vec_d = ((ldd & 3) == 0) && ((C & 3) == 0) && ((uintptr_t)dA % 16 == 0);
    const bool vec_y = ((ldy & 3) == 0) && ((C & 3) == 0) && ((uintptr_t)y % 16 == 0);
    const bool sig = B.act == ACT_SIGMOID;
    const float slope = act_slope(B.act);
    float s1[COUT], s2[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) { s1[o] = 0.f; s2[o] = 0.f; }
    const int64_t block_row0 = (int64_t)blockIdx.x * ML_THREADS * rows_per_thread;
    for (int r = 0; r < rows_per_thread; ++r) {
        const int64_t row = block_row0 + (int64_t)r * ML_THREADS + threadIdx.x;
        if (row >= E) break;
        float d[COUT], yv[COUT];
        load_row<COUT>(dA, ldd, row, C, vec_d, d);              // channels >= C load as 0
        load_row<COUT>(y, ldy, row, C, vec_y, yv);
        if (sig) bn_row_dz<COUT, true>(d, yv, K, slope); else bn_row_dz<COUT, false>(d, yv, K, slope);
#pragma unroll
        for (int o = 0; o < COUT; ++o) { s1[o] += d[o]; s2[o] = fmaf(d[o], (yv[o] - K.mu[o]) * K.is[o], s2[o]); }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
        float a1 = s1[o], a2 = s2[o];
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, sft); a2 += __shfl_xor_sync(0xffffffffu, a2, sft); }
        if (lane == 0) { red[warp][o] = a1; red[warp][COUT + o] = a2; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * COUT; i += ML_THREADS) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < ML_WARPS; ++w) v += red[w][i];
        const int which = i / COUT, o = i - which * COUT;
        if (o < C) partial[((size_t)blockIdx.x * 2 + which) * C + o] = v;
    }
}

// sums[2][C] = fixed-order (double) reduction of the block partials; also used for dgamma (= S2) / dbeta (= S1)
__global__ void sum_partials_kernel(const float *__restrict__ partial, int nblocks, int n, float *__restrict__ out)
{
    pdl_wait();
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    double s = 0.0;
    int b = lane;
    for (; b + 96 < nblocks; b += 128) {
        const float a0 = partial[(size_t)b * n + i], a1 = partial[(size_t)(b + 32) * n + i];
        const float a2 = partial[(size_t)(b + 64) * n + i], a3 = partial[(size_t)(b + 96) * n + i];
        s += (double)a0; s += (double)a1; s += (double)a2; s += (double)a3;
    }
    for (; b < nblocks; b += 32) s += (double)partial[(size_t)b * n + i];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
    if (lane == 0) out[i] = (float)s;
}

// input-gradient pass: dA_prev[row][k] = sum_o dy_o W[o][k]; optionally the BN-backward sums of the layer below
// (S1 = sum dz_prev, S2 = sum dz_prev * xhat_prev with dz_prev = dA_prev * act'_prev) are accumulated on the fly
template <int CIN, int COUT, bool FUSE_PREV>
__global__ void __launch_bounds__(ML_THREADS)
mlp_bwd_input_kernel(const float *__restrict__ dA, int ldd, const float *__restrict__ y, int ldy, int64_t E,
                     const float *__restrict__ W, int cin, int cout, MlpBnCtx B,
                     float *__restrict__ dA_prev, int ldp,
                     const float *__restrict__ y_prev, int ldyp, MlpBnCtx Bp, float *__restrict__ prev_partial,
                     int rows_per_thread)
{
    pdl_wait();
    __shared__ __align__(16) float Wt_s[CIN * COUT];          // transposed: [k][o]
    __shared__ float red[ML_WARPS][2 * CIN];
    __shared__ BnConst<COUT> K;                                 // this layer
    __shared__ BnConst<CIN> Kp;                                 // the layer below (only its sc / sh / mu / is are used)
    for (int i = threadIdx.x; i < CIN * COUT; i += ML_THREADS) {
        const int k = i / COUT, o = i - k * COUT;
        Wt_s[i] = (o < cout && k < cin) ? W[o * cin + k] : 0.f;
    }
    bn_const_fill<COUT>(K, B, cout, ctx_inv_count(B));
    if (FUSE_PREV && prev_partial) bn_const_fill<CIN>(Kp, Bp, cin, 0.f);
    __syncthreads();
    const bool vec_d = ((ldd & 3) == 0) && ((cout & 3) == 0) && ((uintptr_t)dA % 16 == 0);
    const bool vec_y = ((ldy & 3) == 0) && ((cout & 3) == 0) && ((uintptr_t)y % 16 == 0);
    const bool vec_p = ((ldp & 3) == 0) && ((cin & 3) == 0) && ((uintptr_t)dA_prev % 16 == 0);
    const bool vec_yp = y_prev && ((ldyp & 3) == 0) && ((cin & 3) == 0) && ((uintptr_t)y_prev % 16 == 0);
    const bool sig = B.act == ACT_SIGMOID;
    const float slope = act_slope(B.act), pslope = act_slope(Bp.act);
    constexpr int SN = FUSE_PREV ? CIN : 1;
    float s1[SN], s2[SN];
#pragma unroll
    for (int k = 0; k < SN; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    const int64_t block_row0 = (int64_t)blockIdx.x * ML_THREADS * rows_per_thread;
    for (int r = 0; r < rows_per_thread; ++r) {
        const int64_t row = block_row0 + (int64_t)r * ML_THREADS + threadIdx.x;
        if (row >= E) break;
        float d[COUT], yv[COUT];
        load_row<COUT>(dA, ldd, row, cout, vec_d, d);
        load_row<COUT>(y, ldy, row, cout, vec_y, yv);
        if (sig) bn_row_dz<COUT, true>(d, yv, K, slope); else bn_row_dz<COUT, false>(d, yv, K, slope);
        bn_row_dy<COUT>(d, yv, K);
        float g[CIN];
#pragma unroll
        for (int k = 0; k < CIN; ++k) {
            float v = 0.f;
#pragma unroll
            for (int o = 0; o < COUT; o += 4) {
                const float4 w = *reinterpret_cast<const float4 *>(&Wt_s[k * COUT + o]);
                v = fmaf(d[o], w.x, v); v = fmaf(d[o + 1], w.y, v); v = fmaf(d[o + 2], w.z, v); v = fmaf(d[o + 3], w.w, v);
            }
            g[k] = v;
        }
        store_row<CIN>(dA_prev, ldp, row, cin, vec_p, g);
        if (FUSE_PREV && prev_partial) {
            float yp[CIN];
            load_row<CIN>(y_prev, ldyp, row, cin, vec_yp, yp);
#pragma unroll
            for (int k = 0; k < CIN; ++k) {                     // padded channels: W column 0 -> g 0 -> no contribution
                const float z = fmaf(yp[k], Kp.sc[k], Kp.sh[k]);
                const float dz = g[k] * (z > 0.f ? 1.f : pslope);
                s1[k % SN] += dz; s2[k % SN] = fmaf(dz, (yp[k] - Kp.mu[k]) * Kp.is[k], s2[k % SN]);
            }
        }
    }
    if (FUSE_PREV && prev_partial) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < SN; ++k) {
            float a1 = s1[k], a2 = s2[k];
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, sft); a2 += __shfl_xor_sync(0xffffffffu, a2, sft); }
            if (lane == 0) { red[warp][k] = a1; red[warp][CIN + k] = a2; }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * CIN; i += ML_THREADS) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < ML_WARPS; ++w) v += red[w][i];
            const int which = i / CIN, k = i - which * CIN;
            if (k < cin) prev_partial[((size_t)blockIdx.x * 2 + which) * cin + k] = v;
        }
    }
}

// weight-gradient pass.  Tile = 128 rows: threads 0-127 turn (dA, y) of one row into dy[cout], threads 128-255 turn
// the stored input of one row into a_prev[cin]; both land in shared memory.  Then every thread owns a 4x4 block of
// the [cout x cin] gradient and a slice of the tile's rows (2 LDS.128 per 16 FMA), accumulating in registers over
// all tiles of the CTA; slices are combined once at the end in fixed order.
// partial[block][cout][cin+1]: dW[o][k] = sum dy_o * a_prev[k], last column = db[o] = sum dy_o
constexpr int MW_TILE = 128;
template <int CIN, int COUT>
__global__ void __launch_bounds__(ML_THREADS)
mlp_bwd_weight_kernel(const float *__restrict__ dA, int ldd, const float *__restrict__ y, int ldy, int64_t E, int cin, int cout,
                      MlpBnCtx B, const float *__restrict__ x_prev, int ldx, const float *__restrict__ in_scale,
                      const float *__restrict__ in_shift, int in_act, float *__restrict__ partial)
{
    pdl_wait();
    constexpr int DS = COUT + 4, AS = CIN + 4;                  // padded row strides (floats), 16-byte aligned rows
    constexpr int NPB = (COUT / 4) * (CIN / 4);                 // 4x4 blocks of the gradient
    constexpr int NS = ML_THREADS / NPB;                        // row slices
    static_assert(NPB <= ML_THREADS && ML_THREADS % NPB == 0, "bad tiling");
    __shared__ __align__(16) float dy_s[MW_TILE * DS];
    __shared__ __align__(16) float a_s[MW_TILE * AS];
    __shared__ BnConst<COUT> K;                                 // this layer's BatchNorm / activation backward constants
    __shared__ float ps_s[CIN], pt_s[CIN];                      // the layer below: a = act(x*ps + pt)
    const int t = threadIdx.x;
    bn_const_fill<COUT>(K, B, cout, ctx_inv_count(B));
    for (int i = t; i < CIN; i += ML_THREADS) {
        ps_s[i] = i < cin ? (in_scale ? in_scale[i] : 1.f) : 0.f;
        pt_s[i] = (i < cin && in_scale) ? in_shift[i] : 0.f;
    }
    const bool sig = B.act == ACT_SIGMOID, psig = in_act == ACT_SIGMOID;
    const float slope = act_slope(B.act), pslope = act_slope(in_act);
    const int pb = t % NPB, slice = t / NPB;
    const int o4 = (pb / (CIN / 4)) * 4, k4 = (pb % (CIN / 4)) * 4;
    const bool vec_d = ((ldd & 3) == 0) && ((cout & 3) == 0) && ((uintptr_t)dA % 16 == 0);
    const bool vec_y = ((ldy & 3) == 0) && ((cout & 3) == 0) && ((uintptr_t)y % 16 == 0);
    const bool vec_x = ((ldx & 3) == 0) && ((cin & 3) == 0) && ((uintptr_t)x_prev % 16 == 0);
    float acc[4][4], accb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { accb[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f; }
    const int64_t n_tiles = (E + MW_TILE - 1) / MW_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * MW_TILE + (t & (MW_TILE - 1));
        __syncthreads();                                        // previous tile consumed
        if (t < MW_TILE) {
            float d[COUT], yv[COUT];
            if (row < E) {
                load_row<COUT>(dA, ldd, row, cout, vec_d, d); load_row<COUT>(y, ldy, row, cout, vec_y, yv);
                if (sig) bn_row_dz<COUT, true>(d, yv, K, slope); else bn_row_dz<COUT, false>(d, yv, K, slope);
                bn_row_dy<COUT>(d, yv, K);
            } else {
#pragma unroll
                for (int o = 0; o < COUT; ++o) d[o] = 0.f;
            }
#pragma unroll
            for (int o = 0; o < COUT; o += 4)
                *reinterpret_cast<float4 *>(&dy_s[t * DS + o]) = make_float4(d[o], d[o + 1], d[o + 2], d[o + 3]);
        } else {
            const int r = t - MW_TILE;
            float a[CIN];
            if (row < E) {
                load_row<CIN>(x_prev, ldx, row, cin, vec_x, a);
#pragma unroll
                for (int k = 0; k < CIN; ++k) {                 // padded channels: ps = pt = 0 -> a = 0 (sigmoid: masked below)
                    const float v = fmaf(a[k], ps_s[k], pt_s[k]);
                    a[k] = psig ? (k < cin ? 1.f / (1.f + __expf(-v)) : 0.f) : (v > 0.f ? v : pslope * v);
                }
            } else {
#pragma unroll
                for (int k = 0; k < CIN; ++k) a[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < CIN; k += 4)
                *reinterpret_cast<float4 *>(&a_s[r * AS + k]) = make_float4(a[k], a[k + 1], a[k + 2], a[k + 3]);
        }
        __syncthreads();
#pragma unroll 4
        for (int r = slice; r < MW_TILE; r += NS) {
            const float4 dv = *reinterpret_cast<const float4 *>(&dy_s[r * DS + o4]);
            const float4 av = *reinterpret_cast<const float4 *>(&a_s[r * AS + k4]);
            const float dd[4] = {dv.x, dv.y, dv.z, dv.w}, aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                accb[i] += dd[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dd[i], aa[j], acc[i][j]);
            }
        }
    }
    // combine the NS row slices in fixed order (reuse dy_s / a_s as scratch: NS * NPB * 20 floats)
    __syncthreads();
    float *red = dy_s;                                          // [NS][NPB][20] <= MW_TILE*DS + MW_TILE*AS floats (contiguous arrays)
    constexpr int RED_FLOATS = NS * NPB * 20;
    static_assert(RED_FLOATS <= MW_TILE * DS + MW_TILE * AS, "reduction scratch does not fit");
    float *mine = red + ((size_t)slice * NPB + pb) * 20;
#pragma unroll
    for (int i = 0; i < 4; ++i) { mine[16 + i] = accb[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) mine[i * 4 + j] = acc[i][j]; }
    __syncthreads();
    for (int e = t; e < NPB * 20; e += ML_THREADS) {
        const int b2 = e / 20, idx = e - b2 * 20;
        float v = 0.f;
        for (int sl = 0; sl < NS; ++sl) v += red[((size_t)sl * NPB + b2) * 20 + idx];
        const int bo = (b2 / (CIN / 4)) * 4, bk = (b2 % (CIN / 4)) * 4;
        if (idx < 16) {
            const int o = bo + idx / 4, k = bk + idx % 4;
            if (o < cout && k < cin) partial[((size_t)blockIdx.x * cout + o) * (cin + 1) + k] = v;
        } else if (bk == 0) {
            const int o = bo + (idx - 16);
            if (o < cout) partial[((size_t)blockIdx.x * cout + o) * (cin + 1) + cin] = v;
        }
    }
}


// Fused backward of one layer: weight gradient + input gradient (+ the BatchNorm-backward sums of the layer below) in
// ONE sweep over the rows -- (dA, y, x_prev) are read once instead of twice.  Every WARP is autonomous: it owns 16 rows
// per iteration, stages them in its private slice of shared memory and only ever executes __syncwarp in the main loop
// (the first version staged 128-row tiles per CTA behind two block barriers and sat at the barrier 40% of the time, ncu).
//   lanes 0-15 : dy[row] from (dA, y) with this layer's BatchNorm/activation backward      -> dy_s
//   lanes 16-31: stored input row -> x_s (raw) and a_s = act(bn(x))
//   then every lane: CIN/2 input-gradient channels of one row (+ lower-BN sums), and its 4x4 block of dW over 8 or 16 rows.
// Block partials are combined once at the end in fixed order (no atomics).
constexpr int MF_ROWS = 16;
// SIG: this layer's activation is the sigmoid (last layer of the guidance MLP); the layer below never is (host checks).
// Per-channel constants live in shared memory (ncu on the first version: 96 warp instructions per row, 6x the FMA minimum,
// most of them per-element global loads of scale / shift / mean / invstd / sums and the run-time activation switch):
//   this layer:  z = y*sc + sh,  dy = sc*dz + c1*(y - mu) + c0   with c1 = -sc*S2*invstd/E, c0 = -sc*S1/E  (no BN: sc 1, rest 0)
//   layer below: a = act(x*psc + psh),  xhat = (x - pmu)*pis
template <int CIN, int COUT, bool SIG>
__global__ void __launch_bounds__(ML_THREADS, 3)
mlp_bwd_fused_kernel(const float *__restrict__ dA, int ldd, const float *__restrict__ y, int ldy, int64_t E,
                     const float *__restrict__ W, int cin, int cout, MlpBnCtx B,
                     const float *__restrict__ x_prev, int ldx, MlpBnCtx Bp,
                     float *__restrict__ dA_prev, int ldp, float *__restrict__ prev_partial, float *__restrict__ w_partial)
{
    pdl_wait();
    constexpr int DS = COUT + 4, AS = CIN + 4;
    constexpr int NPB = (COUT / 4) * (CIN / 4);                 // 4x4 blocks of dW: 16 or 32
    constexpr int RH = 32 / NPB;                                // row halves per block (2 when NPB == 16, else 1)
    constexpr int RPL = MF_ROWS / RH;                           // rows each lane accumulates per iteration
    constexpr int KH = CIN / 2, OH = COUT / 2;
    static_assert(NPB == 4 || NPB == 8 || NPB == 16 || NPB == 32, "warp tiling needs 4 .. 32 blocks of dW per warp");
    constexpr int SLICE = MF_ROWS * (DS + AS);                  // floats of one warp's staging slice
    constexpr int STAGE_FLOATS = ML_WARPS * SLICE, RED_FLOATS = ML_WARPS * NPB * 20;
    __shared__ __align__(16) float stage_s[STAGE_FLOATS > RED_FLOATS ? STAGE_FLOATS : RED_FLOATS];   // staging, then the dW reduction scratch
    __shared__ __align__(16) float Wt_s[CIN * COUT];            // transposed: [k][o]
    __shared__ __align__(16) float cst_s[5 * COUT + 4 * CIN];   // sc | sh | mu | c1 | c0 (this layer), psc | psh | pmu | pis (below)
    __shared__ float reds[ML_WARPS][2][2 * KH];
    float (*redw)[NPB * 20] = reinterpret_cast<float (*)[NPB * 20]>(stage_s);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float inv_count = ctx_inv_count(B);
    const bool prev_bn = prev_partial != nullptr;
    for (int i = t; i < CIN * COUT; i += ML_THREADS) {
        const int k = i / COUT, o = i - k * COUT;
        Wt_s[i] = (o < cout && k < cin) ? W[o * cin + k] : 0.f;
    }
    for (int i = t; i < COUT; i += ML_THREADS) {
        float sc = 0.f, sh = 0.f, mu = 0.f, c1 = 0.f, c0 = 0.f;
        if (i < cout) {
            sc = 1.f;
            if (B.scale) {
                sc = B.scale[i]; sh = B.shift[i]; mu = B.mean[i];
                c1 = -sc * B.sums[cout + i] * inv_count * B.invstd[i];
                c0 = -sc * B.sums[i] * inv_count;
            }
        }
        cst_s[i] = sc; cst_s[COUT + i] = sh; cst_s[2 * COUT + i] = mu; cst_s[3 * COUT + i] = c1; cst_s[4 * COUT + i] = c0;
    }
    for (int i = t; i < CIN; i += ML_THREADS) {
        float psc = 0.f, psh = 0.f, pmu = 0.f, pis = 0.f;
        if (i < cin) {
            psc = 1.f;
            if (Bp.scale) { psc = Bp.scale[i]; psh = Bp.shift[i]; }
            if (prev_bn) { pmu = Bp.mean[i]; pis = Bp.invstd[i]; }
        }
        float *pc = cst_s + 5 * COUT;
        pc[i] = psc; pc[CIN + i] = psh; pc[2 * CIN + i] = pmu; pc[3 * CIN + i] = pis;
    }
    __syncthreads();
    const float slope = act_slope(B.act), pslope = act_slope(Bp.act);
    const int pb = lane % NPB, rh = lane / NPB;
    const int o4 = (pb / (CIN / 4)) * 4, k4 = (pb % (CIN / 4)) * 4;
    const int r_in = lane & (MF_ROWS - 1), kh0 = (lane >> 4) * KH, oh0 = (lane >> 4) * OH;
    const bool vec_dy = ((ldd & 3) == 0) && ((ldy & 3) == 0) && cout == COUT && ((uintptr_t)dA % 16 == 0) && ((uintptr_t)y % 16 == 0);
    const bool vec_x = ((ldx & 3) == 0) && cin == CIN && ((uintptr_t)x_prev % 16 == 0);
    const bool vec_p = ((ldp & 3) == 0) && ((uintptr_t)dA_prev % 16 == 0) && (KH % 4 == 0) && cin == CIN;
    const float *csc = cst_s + oh0, *csh = cst_s + COUT + oh0, *cmu = cst_s + 2 * COUT + oh0, *cc1 = cst_s + 3 * COUT + oh0,
                *cc0 = cst_s + 4 * COUT + oh0;
    const float *psc = cst_s + 5 * COUT + kh0, *psh = psc + CIN, *pmu = psc + 2 * CIN, *pis = psc + 3 * CIN;
    float *dyw = stage_s + warp * SLICE, *aw = dyw + MF_ROWS * DS;
    float acc[4][4], accb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { accb[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f; }
    float s1[KH], s2[KH];
#pragma unroll
    for (int k = 0; k < KH; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    const int64_t n_groups = (E + MF_ROWS - 1) / MF_ROWS;
    for (int64_t g = (int64_t)blockIdx.x * ML_WARPS + warp; g < n_groups; g += (int64_t)gridDim.x * ML_WARPS) {
        const int64_t row = g * MF_ROWS + r_in;
        const bool live = row < E;
        const size_t rc = (size_t)(live ? row : E - 1);        // rows past the end re-read the last row and are masked to zero
        const float valid = live ? 1.f : 0.f;
        __syncwarp();                                           // the previous group's reads of the warp's slice are done
        float xr[KH];                                           // my half of the stored input row: staging AND the lower-BN sums
        {   // staging, uniform over the warp: lane (row, half) handles half of the row's output and input channels
            float d[OH], yv[OH];
            if (vec_dy) {
#pragma unroll
                for (int o = 0; o < OH; o += 4) {
                    const float4 dv = __ldg(reinterpret_cast<const float4 *>(dA + rc * ldd + oh0 + o));
                    const float4 yy = __ldg(reinterpret_cast<const float4 *>(y + rc * ldy + oh0 + o));
                    d[o] = dv.x; d[o + 1] = dv.y; d[o + 2] = dv.z; d[o + 3] = dv.w;
                    yv[o] = yy.x; yv[o + 1] = yy.y; yv[o + 2] = yy.z; yv[o + 3] = yy.w;
                }
            } else {
#pragma unroll
                for (int o = 0; o < OH; ++o) {
                    const bool ok = oh0 + o < cout;
                    d[o] = ok ? __ldg(dA + rc * ldd + oh0 + o) : 0.f;
                    yv[o] = ok ? __ldg(y + rc * ldy + oh0 + o) : 0.f;
                }
            }
            if (vec_x) {
#pragma unroll
                for (int k = 0; k < KH; k += 4) {
                    const float4 xv = __ldg(reinterpret_cast<const float4 *>(x_prev + rc * ldx + kh0 + k));
                    xr[k] = xv.x; xr[k + 1] = xv.y; xr[k + 2] = xv.z; xr[k + 3] = xv.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < KH; ++k) xr[k] = (kh0 + k < cin) ? __ldg(x_prev + rc * ldx + kh0 + k) : 0.f;
            }
#pragma unroll
            for (int o = 0; o < OH; ++o) {
                const float sc = csc[o];
                const float z = fmaf(yv[o], sc, csh[o]);
                float gz;
                if (SIG) { const float av = 1.f / (1.f + __expf(-z)); gz = av * (1.f - av); }
                else gz = z > 0.f ? 1.f : slope;
                const float dz = d[o] * gz;
                d[o] = valid * fmaf(sc, dz, fmaf(cc1[o], yv[o] - cmu[o], cc0[o]));      // padded channels: all constants 0
            }
#pragma unroll
            for (int o = 0; o < OH; o += 4)
                *reinterpret_cast<float4 *>(&dyw[r_in * DS + oh0 + o]) = make_float4(d[o], d[o + 1], d[o + 2], d[o + 3]);
            float a[KH];
#pragma unroll
            for (int k = 0; k < KH; ++k) {
                const float v = fmaf(xr[k], psc[k], psh[k]);
                a[k] = valid * (v > 0.f ? v : pslope * v);
            }
#pragma unroll
            for (int k = 0; k < KH; k += 4)
                *reinterpret_cast<float4 *>(&aw[r_in * AS + kh0 + k]) = make_float4(a[k], a[k + 1], a[k + 2], a[k + 3]);
        }
        __syncwarp();
        if (live && dA_prev) {                                  // input gradient: KH channels of one row per lane (skipped for a chain's first layer)
            float d[COUT];
#pragma unroll
            for (int o = 0; o < COUT; o += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(&dyw[r_in * DS + o]);
                d[o] = v.x; d[o + 1] = v.y; d[o + 2] = v.z; d[o + 3] = v.w;
            }
            float gk[KH];
#pragma unroll
            for (int k = 0; k < KH; ++k) {
                float v = 0.f;
#pragma unroll
                for (int o = 0; o < COUT; o += 4) {
                    const float4 w = *reinterpret_cast<const float4 *>(&Wt_s[(kh0 + k) * COUT + o]);
                    v = fmaf(d[o], w.x, v); v = fmaf(d[o + 1], w.y, v); v = fmaf(d[o + 2], w.z, v); v = fmaf(d[o + 3], w.w, v);
                }
                gk[k] = v;
            }
            if (vec_p) {
#pragma unroll
                for (int k = 0; k < KH; k += 4)
                    *reinterpret_cast<float4 *>(dA_prev + (size_t)row * ldp + kh0 + k) = make_float4(gk[k], gk[k + 1], gk[k + 2], gk[k + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < KH; ++k) if (kh0 + k < cin) dA_prev[(size_t)row * ldp + kh0 + k] = gk[k];
            }
            if (prev_bn) {
#pragma unroll
                for (int k = 0; k < KH; ++k) {                  // padded channels: W column 0 -> gk 0 -> no contribution
                    const float z = fmaf(xr[k], psc[k], psh[k]);
                    const float dz = gk[k] * (z > 0.f ? 1.f : pslope);
                    s1[k] += dz; s2[k] = fmaf(dz, (xr[k] - pmu[k]) * pis[k], s2[k]);
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < RPL; ++rr) {                      // dW: this lane's 4x4 block over its rows of the group
            const int r = rh * RPL + rr;
            const float4 dv = *reinterpret_cast<const float4 *>(&dyw[r * DS + o4]);
            const float4 av = *reinterpret_cast<const float4 *>(&aw[r * AS + k4]);
            const float dd[4] = {dv.x, dv.y, dv.z, dv.w}, aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                accb[i] += dd[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dd[i], aa[j], acc[i][j]);
            }
        }
    }
    // ---- block partials, fixed order: lanes -> warps ----
    __syncthreads();                                            // every warp is out of its staging slice: reuse it as scratch
#pragma unroll
    for (int off = NPB; off < 32; off <<= 1) {                  // the RH row groups of a block sit NPB lanes apart: fixed tree
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            accb[i] += __shfl_xor_sync(0xffffffffu, accb[i], off);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], off);
        }
    }
    if (lane < NPB) {
        float *mine = &redw[warp][pb * 20];
#pragma unroll
        for (int i = 0; i < 4; ++i) { mine[16 + i] = accb[i];
#pragma unroll
            for (int j = 0; j < 4; ++j) mine[i * 4 + j] = acc[i][j]; }
    }
    if (prev_bn) {                                              // lanes 0-15 hold channels [0, KH), lanes 16-31 [KH, CIN)
#pragma unroll
        for (int k = 0; k < KH; ++k) {
            float a1 = s1[k], a2 = s2[k];
#pragma unroll
            for (int sft = 8; sft > 0; sft >>= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, sft); a2 += __shfl_xor_sync(0xffffffffu, a2, sft); }
            if ((lane & 15) == 0) { reds[warp][lane >> 4][k] = a1; reds[warp][lane >> 4][KH + k] = a2; }
        }
    }
    __syncthreads();
    for (int e = t; e < NPB * 20; e += ML_THREADS) {
        const int b2 = e / 20, idx = e - b2 * 20;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < ML_WARPS; ++w) v += redw[w][e];
        const int bo = (b2 / (CIN / 4)) * 4, bk = (b2 % (CIN / 4)) * 4;
        if (idx < 16) {
            const int o = bo + idx / 4, k = bk + idx % 4;
            if (o < cout && k < cin) w_partial[((size_t)blockIdx.x * cout + o) * (cin + 1) + k] = v;
        } else if (bk == 0) {
            const int o = bo + (idx - 16);
            if (o < cout) w_partial[((size_t)blockIdx.x * cout + o) * (cin + 1) + cin] = v;
        }
    }
    if (prev_bn) {
        for (int i = t; i < 2 * CIN; i += ML_THREADS) {
            const int which = i / CIN, k = i - which * CIN;
            const int half = k / KH, kl = k - half * KH;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < ML_WARPS; ++w) v += reds[w][half][which * KH + kl];
            if (k < cin) prev_partial[((size_t)blockIdx.x * 2 + which) * cin + k] = v;
        }
    }
}

// dW[o][k], db[o] from block partials (fixed order, double)
__global__ void mlp_weight_finalize_kernel(const float *__restrict__ partial, int nblocks, int cout, int cin,
                                           float *__restrict__ dW, float *__restrict__ db)
{
    pdl_wait();
    const int n = cout * (cin + 1);
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    double s = 0.0;
    int b = lane;
    for (; b + 96 < nblocks; b += 128) {
        const float a0 = partial[(size_t)b * n + i], a1 = partial[(size_t)(b + 32) * n + i];
        const float a2 = partial[(size_t)(b + 64) * n + i], a3 = partial[(size_t)(b + 96) * n + i];
        s += (double)a0; s += (double)a1; s += (double)a2; s += (double)a3;
    }
    for (; b < nblocks; b += 32) s += (double)partial[(size_t)b * n + i];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
    if (lane != 0) return;
    const int o = i / (cin + 1), k = i - o * (cin + 1);
    if (k < cin) { if (dW) dW[o * cin + k] = (float)s; }
    else if (db) db[o] = (float)s;
}

// fused-backward finalize: dW / db from the weight partials AND the lower layer's BatchNorm-backward sums from theirs,
// one launch (warp per output element, fixed order, double accumulation)
__global__ void mlp_fused_finalize_kernel(const float *__restrict__ w_partial, const float *__restrict__ p_partial, int nblocks,
                                          int cout, int cin, float *__restrict__ dW, float *__restrict__ db,
                                          float *__restrict__ prev_sums)
{
    pdl_wait();
    const int n_w = cout * (cin + 1), n_p = p_partial ? 2 * cin : 0;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_w + n_p) return;
    const bool is_w = i < n_w;
    const float *src = is_w ? w_partial + i : p_partial + (i - n_w);
    const int stride = is_w ? n_w : n_p;
    double s = 0.0;
    int b = lane;
    for (; b + 96 < nblocks; b += 128) {                 // four independent loads in flight, same summation order
        const float a0 = src[(size_t)b * stride], a1 = src[(size_t)(b + 32) * stride];
        const float a2 = src[(size_t)(b + 64) * stride], a3 = src[(size_t)(b + 96) * stride];
        s += (double)a0; s += (double)a1; s += (double)a2; s += (double)a3;
    }
    for (; b < nblocks; b += 32) s += (double)src[(size_t)b * stride];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
    if (lane != 0) return;
    if (is_w) {
        const int o = i / (cin + 1), k = i - o * (cin + 1);
        if (k < cin) { if (dW) dW[o * cin + k] = (float)s; }
        else if (db) db[o] = (float)s;
    } else {
        prev_sums[i - n_w] = (float)s;
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
static inline int cmax_of(int c) { return c <= 16 ? 16 : (c <= 32 ? 32 : 64); }
// Template sizes of the forward / fused-backward kernels: the WeightNet layers (12 -> 8 -> 8 -> 16, 35 chains per step, the
// largest kernel group of the step) get exact 8-wide instances instead of 16-wide padding (ncu r02: 517 instructions per row
// in the forward for 96 useful FMAs, 1555 in the fused backward).
static inline void chain_dims(int cin, int cout, int *ci, int *co) {
    *ci = cmax_of(cin); *co = cmax_of(cout);
    if (cout <= 8 && *ci <= 16) *co = 8;
    if (cin <= 8 && *co <= 16) *ci = 8;
}
constexpr int ML_RPT = 8;          // most rows per thread in the streaming passes
// Rows per thread: 8 for the level-0 edge tensors (1.6 M rows), fewer as the tensor shrinks so that the coarse levels
// (3 k .. 80 k rows) still spread over the SMs -- with a fixed 8 a 16 k-row pass ran as 9 blocks whose threads walked
// 8 rows one after the other (ncu: 47-56 us for a pass that moves 4 MB; 14 of the model's 25 PointConvFormer layers
// live on those levels).
static inline int mlp_rpt(int64_t E) {
    const int64_t r = (E + (int64_t)ML_THREADS * kNumSMs * 4 - 1) / ((int64_t)ML_THREADS * kNumSMs * 4);
    return r < 1 ? 1 : (r > ML_RPT ? ML_RPT : (int)r);
}
static inline int mlp_blocks(int64_t E) {
    const int64_t per = (int64_t)ML_THREADS * mlp_rpt(E);
    return (int)((E + per - 1) / per);
}
static inline int mlp_wblocks(int64_t E, int *gpb) {
    const int64_t tiles = (E + 127) / 128;
    int64_t blocks = tiles < 3 * kNumSMs ? tiles : 3 * kNumSMs;
    if (blocks < 1) blocks = 1;
    *gpb = 0;
    return (int)blocks;
}

}  // namespace pcfb

using namespace pcfb;

extern "C" int pcfb_mlp_supported(int cin, int cout)
{
    if (cin < 1 || cout < 1 || cin > 64 || cout > 64) return 0;
    const int ci = cmax_of(cin), co = cmax_of(cout);
    // template instances: (16,16) (16,32) (16,64) (32,16) (32,32) (64,16)
    return (ci == 16) || (ci == 32 && co <= 32) || (ci == 64 && co == 16) ? 1 : 0;
}

// number of floats of scratch needed by the forward stats / backward partials for E rows
extern "C" size_t pcfb_mlp_workspace(int64_t E, int cin, int cout)
{
    int gpb;
    const size_t a = (size_t)mlp_blocks(E) * 2 * (size_t)(cout > cin ? cout : cin);
    const size_t b = (size_t)mlp_wblocks(E, &gpb) * ((size_t)cout * (cin + 1) + 2 * (size_t)cin);   // fused backward: dW partials + lower-BN sums
    return align_up((a > b ? a : b) * sizeof(float) + 1024, 256);
}

#define ML_DISPATCH_IO(CIN_V, COUT_V, CALL)                                         \
    do {                                                                            \
        if (CIN_V == 16 && COUT_V == 16) { constexpr int CI = 16, CO = 16; CALL; }   \
        else if (CIN_V == 16 && COUT_V == 32) { constexpr int CI = 16, CO = 32; CALL; } \
        else if (CIN_V == 16 && COUT_V == 64) { constexpr int CI = 16, CO = 64; CALL; } \
        else if (CIN_V == 32 && COUT_V == 16) { constexpr int CI = 32, CO = 16; CALL; } \
        else if (CIN_V == 32 && COUT_V == 32) { constexpr int CI = 32, CO = 32; CALL; } \
        else { constexpr int CI = 64, CO = 16; CALL; }                               \
    } while (0)

// y[E,cout] = act_in(x*in_scale+in_shift) W^T + b ; stat_partial (optional): [blocks][2][cout]; returns #blocks via *nblocks
extern "C" int pcfb_mlp_forward(const float *x, int ldx, int64_t E, int cin, int cout, const float *W, const float *b,
                                const float *in_scale, const float *in_shift, int in_act, float *y, int ldy,
                                float *stat_partial, int *nblocks, void *stream)
{
    PCFB_REQUIRE(pcfb_mlp_supported(cin, cout), "pcfb_mlp_forward: unsupported layer size %d -> %d", cin, cout);
    PCFB_REQUIRE(x && W && y && E >= 0, "pcfb_mlp_forward: null pointer");
    const int blocks = mlp_blocks(E);
    if (nblocks) *nblocks = blocks;
    if (E == 0) return PCFB_OK;
    MlpLayer L{W, b, cin, cout, in_scale, in_shift, in_act};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int ci, co;
    chain_dims(cin, cout, &ci, &co);
    if (ci == 8 && co == 8) launch_k(mlp_fwd_kernel<8, 8>, blocks, ML_THREADS, 0, st, x, ldx, E, L, y, ldy, stat_partial, mlp_rpt(E));
    else if (ci == 8 && co == 16) launch_k(mlp_fwd_kernel<8, 16>, blocks, ML_THREADS, 0, st, x, ldx, E, L, y, ldy, stat_partial, mlp_rpt(E));
    else if (ci == 16 && co == 8) launch_k(mlp_fwd_kernel<16, 8>, blocks, ML_THREADS, 0, st, x, ldx, E, L, y, ldy, stat_partial, mlp_rpt(E));
    else ML_DISPATCH_IO(ci, co, (launch_k(mlp_fwd_kernel<CI, CO>, blocks, ML_THREADS, 0, st, x, ldx, E, L, y, ldy, stat_partial, mlp_rpt(E))));
    return check_launch("mlp_fwd_kernel");
}

extern "C" int pcfb_bn_act(const float *y, int64_t rows, int C, const float *scale, const float *shift, int act, float *out,
                           const float *residual, int residual_after_act, void *stream)
{
    PCFB_REQUIRE(y && out, "pcfb_bn_act: null pointer");
    const int64_t n = rows * C;
    if (n == 0) return PCFB_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    if ((C & 3) == 0 && C <= 1024 && ((uintptr_t)y % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)residual % 16 == 0) &&
        (!scale || (((uintptr_t)scale % 16 == 0) && ((uintptr_t)shift % 16 == 0)))) {
        const int c4 = C >> 2, ry = 256 / c4;
        int64_t b4 = (rows + ry - 1) / ry;
        if (b4 > (int64_t)kNumSMs * 8) b4 = (int64_t)kNumSMs * 8;
        launch_k(bn_act4_kernel, (int)b4, 256, 0, static_cast<cudaStream_t>(stream), y, rows, C, scale, shift, act, out, c4, ry, residual,
                                                                               residual_after_act);
        return check_launch("bn_act4_kernel");
    }
    launch_k(bn_act_kernel, (int)blocks, 256, 0, static_cast<cudaStream_t>(stream), y, n, C, scale, shift, act, out, residual, residual_after_act);
    return check_launch("bn_act_kernel");
}

// BN-backward sums of one layer: sums[2][C] = (sum dz, sum dz*xhat) for dz = dA * act'(y*scale+shift)
extern "C" int pcfb_mlp_backward_stats(const float *dA, int ldd, const float *y, int ldy, int64_t E, int C, const float *scale,
                                       const float *shift, const float *mean, const float *invstd, int act, float *sums,
                                       int *nblocks, void *workspace, size_t workspace_bytes, void *stream)
{
    PCFB_REQUIRE(C >= 1 && C <= 64 && dA && y && scale && shift && mean && invstd && workspace, "pcfb_mlp_backward_stats: bad arguments");
    const int blocks = mlp_blocks(E);
    if (nblocks) *nblocks = E > 0 ? blocks : 0;
    PCFB_REQUIRE(workspace_bytes >= (size_t)blocks * 2 * C * sizeof(float), "pcfb_mlp_backward_stats: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(workspace);
    MlpBnCtx B{scale, shift, mean, invstd, nullptr, act, 0.f, nullptr};
    int rc;
    if (E > 0) {
        const int co = C <= 8 ? 8 : cmax_of(C);
        if (co == 8) launch_k(mlp_bwd_stats_kernel<8>, blocks, ML_THREADS, 0, st, dA, ldd, y, ldy, E, C, B, partial, mlp_rpt(E));
        else if (co == 16) launch_k(mlp_bwd_stats_kernel<16>, blocks, ML_THREADS, 0, st, dA, ldd, y, ldy, E, C, B, partial, mlp_rpt(E));
        else if (co == 32) launch_k(mlp_bwd_stats_kernel<32>, blocks, ML_THREADS, 0, st, dA, ldd, y, ldy, E, C, B, partial, mlp_rpt(E));
        else launch_k(mlp_bwd_stats_kernel<64>, blocks, ML_THREADS, 0, st, dA, ldd, y, ldy, E, C, B, partial, mlp_rpt(E));
        if ((rc = check_launch("mlp_bwd_stats_kernel"))) return rc;
    }
    if (!sums) return PCFB_OK;                                   // the caller reduces the partials itself (pcfb_bn_reduce_sums)
    launch_k(sum_partials_kernel, ceil_div(2 * C * 32, 128), 128, 0, st, partial, E > 0 ? blocks : 0, 2 * C, sums);
    return check_launch("sum_partials_kernel");
}

// One layer backward.  Inputs: dA (grad wrt this layer's activation output), y (its pre-BN output), BN context
// (scale/shift/mean/invstd/sums; scale == NULL -> no BN), W.  Outputs: dA_prev [E,cin] (may be NULL), dW [cout,cin],
// db [cout] (may be NULL).  x_prev (+ in_scale/in_shift/in_act) is the layer's input as stored (pre-BN output of the layer
// below, or the raw chain input).  If prev_sums != NULL the BN-backward sums of the layer below (whose BN context is
// prev_*) are produced on the fly: prev_sums[2][cin].
extern "C" int pcfb_mlp_backward(const float *dA, int ldd, const float *y, int ldy, int64_t E, int cin, int cout,
                                 const float *W, const float *scale, const float *shift, const float *mean,
                                 const float *invstd, const float *sums, int act,
                                 const float *x_prev, int ldx, const float *in_scale, const float *in_shift, int in_act,
                                 const float *prev_mean, const float *prev_invstd,
                                 float *dA_prev, int ldp, float *prev_sums, float *dW, float *db, const double *d_count,
                                 void *workspace, size_t workspace_bytes, void *stream)
{
    PCFB_REQUIRE(pcfb_mlp_supported(cin, cout), "pcfb_mlp_backward: unsupported layer size %d -> %d", cin, cout);
    PCFB_REQUIRE(dA && y && W && x_prev && workspace, "pcfb_mlp_backward: null pointer");
    PCFB_REQUIRE(!scale || (shift && mean && invstd && sums), "pcfb_mlp_backward: incomplete BatchNorm context");
    PCFB_REQUIRE(!prev_sums || (dA_prev && in_scale && in_shift && prev_mean && prev_invstd), "pcfb_mlp_backward: incomplete lower BatchNorm context");
    PCFB_REQUIRE(!prev_sums || cin <= 32, "pcfb_mlp_backward: fused lower-layer sums need cin <= 32 (use pcfb_mlp_backward_stats)");
    PCFB_REQUIRE(workspace_bytes >= pcfb_mlp_workspace(E, cin, cout), "pcfb_mlp_backward: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(workspace);
    const float inv_count = E > 0 ? (float)(1.0 / (double)E) : 0.f;
    MlpBnCtx B{scale, shift, mean, invstd, sums, act, inv_count, d_count};
    MlpBnCtx Bp{in_scale, in_shift, prev_mean, prev_invstd, nullptr, in_act, inv_count, d_count};
    int ci = cmax_of(cin), co = cmax_of(cout);
    int rc;
    if ((dW || db) && ci == 16 && co <= 32 && E > 0 && in_act != ACT_SIGMOID) {
        chain_dims(cin, cout, &ci, &co);   // one sweep (dA_prev == NULL: weight gradient only): input gradient + weight gradient (+ lower-BN sums)
        int gpb;
        const int blocks = mlp_wblocks(E, &gpb);
        (void)gpb;
        float *w_part = partial;
        float *p_part = partial + (size_t)blocks * cout * (cin + 1);
#define ML_FUSED_CASE(CI_, CO_)                                                                                   \
        if (ci == CI_ && co == CO_) {                                                                             \
            if (act == ACT_SIGMOID)                                                                               \
                launch_k(mlp_bwd_fused_kernel<CI_, CO_, true>, blocks, ML_THREADS, 0, st, dA, ldd, y, ldy, E, W, cin, cout, B, x_prev, ldx, Bp, \
                                                                                    dA_prev, ldp, prev_sums ? p_part : nullptr, w_part); \
            else                                                                                                  \
                launch_k(mlp_bwd_fused_kernel<CI_, CO_, false>, blocks, ML_THREADS, 0, st, dA, ldd, y, ldy, E, W, cin, cout, B, x_prev, ldx, Bp, \
                                                                                     dA_prev, ldp, prev_sums ? p_part : nullptr, w_part); \
        }
        ML_FUSED_CASE(8, 8) ML_FUSED_CASE(8, 16) ML_FUSED_CASE(16, 8) ML_FUSED_CASE(16, 16) ML_FUSED_CASE(16, 32)
#undef ML_FUSED_CASE
        ci = cmax_of(cin); co = cmax_of(cout);
        if ((rc = check_launch("mlp_bwd_fused_kernel"))) return rc;
        const int n_fin = cout * (cin + 1) + (prev_sums ? 2 * cin : 0);
        launch_k(mlp_fused_finalize_kernel, ceil_div(n_fin * 32, 256), 256, 0, st, w_part, prev_sums ? p_part : nullptr, blocks, cout, cin,
                                                                            dW, db, prev_sums);
        return check_launch("mlp_fused_finalize_kernel");
    }
    if (dA_prev) {
        const int blocks = mlp_blocks(E);
        if (E > 0) {
            ML_DISPATCH_IO(ci, co, (launch_k(mlp_bwd_input_kernel<CI, CO, (CI <= 32)>, blocks, ML_THREADS, 0, st, dA, ldd, y, ldy, E, W, cin, cout, B, dA_prev, ldp, x_prev, ldx, Bp, prev_sums ? partial : nullptr, mlp_rpt(E))));
            if ((rc = check_launch("mlp_bwd_input_kernel"))) return rc;
        }
        if (prev_sums) {
            launch_k(sum_partials_kernel, ceil_div(2 * cin * 32, 128), 128, 0, st, partial, E > 0 ? blocks : 0, 2 * cin, prev_sums);
            if ((rc = check_launch("sum_partials_kernel"))) return rc;
        }
    }
    if (dW || db) {
        int gpb;
        const int blocks = mlp_wblocks(E, &gpb);
        (void)gpb;
        if (E > 0) {
            ML_DISPATCH_IO(ci, co, (launch_k(mlp_bwd_weight_kernel<CI, CO>, blocks, ML_THREADS, 0, st, dA, ldd, y, ldy, E, cin, cout, B, x_prev, ldx, in_scale, in_shift, in_act, partial)));
            if ((rc = check_launch("mlp_bwd_weight_kernel"))) return rc;
        }
        launch_k(mlp_weight_finalize_kernel, ceil_div(cout * (cin + 1) * 32, 256), 256, 0, st, partial, E > 0 ? blocks : 0, cout, cin, dW, db);
        if ((rc = check_launch("mlp_weight_finalize_kernel"))) return rc;
    }
    return PCFB_OK;
}

extern "C" int pcfb_sum_partials(const float *partial, int nblocks, int n, float *out, void *stream)
{
    PCFB_REQUIRE(partial && out && n >= 1, "pcfb_sum_partials: bad arguments");
    launch_k(sum_partials_kernel, ceil_div(n * 32, 128), 128, 0, static_cast<cudaStream_t>(stream), partial, nblocks, n, out);
    return check_launch("sum_partials_kernel");
}

// ------------------------------------------------------------------------------------------------------------
// BatchNorm (+ activation) over a contiguous [rows, C] tensor for any C % 4 == 0, C <= 1024: the BatchNorm + ReLU
// behind the fused contraction (/root/reference/layers.py:708-709,721, 893-898, 1086-1092) and the Linear_BN of the
// wide per-point blocks (layer_utils.py:241-319) whose Linear runs on gemm_nt.  Train mode = 2 reads + 1 write in the
// forward (statistics, apply + activation), 2 + 2 reads + 1 write in the backward; all sums are block partials reduced
// in fixed order (deterministic, no atomics).
// Thread layout: tx = float4 column group (C/4 of them), ty = row lane (256 / (C/4) rows in flight per block).
// ------------------------------------------------------------------------------------------------------------
namespace pcfb {

constexpr int BA_THREADS = 256;

struct BaGeom { int c4, ry, rows_per_block, blocks; };

static BaGeom ba_geom(int64_t rows, int C) {
    BaGeom g;
    g.c4 = C >> 2;
    g.ry = BA_THREADS / g.c4;
    int64_t rpb = (rows + (int64_t)kNumSMs * 8 - 1) / ((int64_t)kNumSMs * 8);        // <= 8 blocks per SM
    if (rpb < (int64_t)g.ry * 4) rpb = (int64_t)g.ry * 4;
    rpb = (rpb + g.ry - 1) / g.ry * g.ry;
    g.rows_per_block = (int)rpb;
    g.blocks = rows > 0 ? (int)((rows + rpb - 1) / rpb) : 0;
    return g;
}

// fixed-order reduction of the per-row-lane accumulators of one block: partial[block][which][C]
__device__ __forceinline__ void ba_block_reduce(float4 s1, float4 s2, int tx, int ty, int c4, int ry, bool active,
                                                float4 *sm, float *__restrict__ partial, int C)
{
    if (active) { sm[ty * c4 + tx] = s1; sm[(ry + ty) * c4 + tx] = s2; }
    __syncthreads();
    if (active && ty == 0) {
        float4 a = sm[tx], b = sm[ry * c4 + tx];
        for (int r = 1; r < ry; ++r) {
            const float4 u = sm[r * c4 + tx], v = sm[(ry + r) * c4 + tx];
            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
            b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
        }
        *reinterpret_cast<float4 *>(partial + ((size_t)blockIdx.x * 2 + 0) * C + 4 * tx) = a;
        *reinterpret_cast<float4 *>(partial + ((size_t)blockIdx.x * 2 + 1) * C + 4 * tx) = b;
    }
}

// partial[block][2][C] = sum (x - pivot), sum (x - pivot)^2 over the block's rows
__global__ void __launch_bounds__(BA_THREADS)
bn_stats_kernel(const float *__restrict__ x, int64_t rows, int C, const float *__restrict__ pivot, float *__restrict__ partial,
                int c4, int ry, int rows_per_block)
{
    pdl_wait();
    extern __shared__ float4 ba_sm[];
    const int tx = threadIdx.x % c4, ty = threadIdx.x / c4;
    const bool active = ty < ry;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    if (active) {
        // scalar loads: the pivot is usually a bias parameter, which may be a 4-byte aligned view into a flat buffer
        const float4 pv = pivot ? make_float4(__ldg(pivot + 4 * tx), __ldg(pivot + 4 * tx + 1), __ldg(pivot + 4 * tx + 2), __ldg(pivot + 4 * tx + 3))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
        const int64_t r1 = min(rows, r0 + rows_per_block);
        for (int64_t r = r0 + ty; r < r1; r += ry) {
            float4 v = __ldg(reinterpret_cast<const float4 *>(x + r * C) + tx);
            v.x -= pv.x; v.y -= pv.y; v.z -= pv.z; v.w -= pv.w;
            s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
            s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
        }
    }
    ba_block_reduce(s1, s2, tx, ty, c4, ry, active, ba_sm, partial, C);
}

// out = act(y * scale + shift), float4 columns (the vector form of bn_act_kernel)
__global__ void __launch_bounds__(BA_THREADS)
bn_act4_kernel(const float *__restrict__ y, int64_t rows, int C, const float *__restrict__ scale, const float *__restrict__ shift,
               int act, float *__restrict__ out, int c4, int ry, const float *__restrict__ res, int res_after)
{
    pdl_wait();
    const int tx = threadIdx.x % c4, ty = threadIdx.x / c4;
    if (ty >= ry) return;
    const float4 sc = scale ? __ldg(reinterpret_cast<const float4 *>(scale) + tx) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 sh = scale ? __ldg(reinterpret_cast<const float4 *>(shift) + tx) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = (int64_t)blockIdx.x * ry + ty; r < rows; r += (int64_t)gridDim.x * ry) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(y + r * C) + tx);
        float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (res) rr = __ldg(reinterpret_cast<const float4 *>(res + r * C) + tx);
        const float4 pre = res_after ? make_float4(0.f, 0.f, 0.f, 0.f) : rr, post = res_after ? rr : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 o;
        o.x = act_fwd(fmaf(v.x, sc.x, sh.x) + pre.x, act) + post.x; o.y = act_fwd(fmaf(v.y, sc.y, sh.y) + pre.y, act) + post.y;
        o.z = act_fwd(fmaf(v.z, sc.z, sh.z) + pre.z, act) + post.z; o.w = act_fwd(fmaf(v.w, sc.w, sh.w) + pre.w, act) + post.w;
        *(reinterpret_cast<float4 *>(out + r * C) + tx) = o;
    }
}

__device__ __forceinline__ float ba_dz(float d, float yv, float sc, float sh, int act, float r = 0.f) {
    const float z = fmaf(yv, sc, sh) + r;                 // r: the residual added before the activation (0 otherwise)
    return d * act_bwd(z, act_fwd(z, act), act);
}

// partial[block][2][C] = sum dz, sum dz * xhat   with dz = dA * act'(y*scale+shift), xhat = (y - mean) * invstd
__global__ void __launch_bounds__(BA_THREADS)
bn_bwd_stats_kernel(const float *__restrict__ dA, const float *__restrict__ y, int64_t rows, int C, const float *__restrict__ scale,
                    const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ invstd, int act,
                    float *__restrict__ partial, int c4, int ry, int rows_per_block, const float *__restrict__ res)
{
    pdl_wait();
    extern __shared__ float4 ba_sm[];
    const int tx = threadIdx.x % c4, ty = threadIdx.x / c4;
    const bool active = ty < ry;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    if (active) {
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + tx), sh = __ldg(reinterpret_cast<const float4 *>(shift) + tx);
        const float4 mu = __ldg(reinterpret_cast<const float4 *>(mean) + tx), is = __ldg(reinterpret_cast<const float4 *>(invstd) + tx);
        const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
        const int64_t r1 = min(rows, r0 + rows_per_block);
        for (int64_t r = r0 + ty; r < r1; r += ry) {
            const float4 d = __ldg(reinterpret_cast<const float4 *>(dA + r * C) + tx);
            const float4 v = __ldg(reinterpret_cast<const float4 *>(y + r * C) + tx);
            float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
            if (res) rr = __ldg(reinterpret_cast<const float4 *>(res + r * C) + tx);
            const float dx_ = ba_dz(d.x, v.x, sc.x, sh.x, act, rr.x), dy_ = ba_dz(d.y, v.y, sc.y, sh.y, act, rr.y);
            const float dz_ = ba_dz(d.z, v.z, sc.z, sh.z, act, rr.z), dw_ = ba_dz(d.w, v.w, sc.w, sh.w, act, rr.w);
            s1.x += dx_; s1.y += dy_; s1.z += dz_; s1.w += dw_;
            s2.x = fmaf(dx_, (v.x - mu.x) * is.x, s2.x); s2.y = fmaf(dy_, (v.y - mu.y) * is.y, s2.y);
            s2.z = fmaf(dz_, (v.z - mu.z) * is.z, s2.z); s2.w = fmaf(dw_, (v.w - mu.w) * is.w, s2.w);
        }
    }
    ba_block_reduce(s1, s2, tx, ty, c4, ry, active, ba_sm, partial, C);
}

// train (sums != null): dX = scale * (dz - S1/E - xhat * S2/E);  eval: dX = scale * dz
__global__ void __launch_bounds__(BA_THREADS)
bn_bwd_kernel(const float *__restrict__ dA, const float *__restrict__ y, int64_t rows, int C, const float *__restrict__ scale,
              const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ invstd,
              const float *__restrict__ sums, int act, float inv_count, const double *__restrict__ d_count,
              float *__restrict__ dX, int c4, int ry, const float *__restrict__ res, float *__restrict__ dR)
{
    pdl_wait();
    const int tx = threadIdx.x % c4, ty = threadIdx.x / c4;
    if (ty >= ry) return;
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + tx), sh = __ldg(reinterpret_cast<const float4 *>(shift) + tx);
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = mu, m1 = mu, m2 = mu;
    if (sums) {
        const float ic = d_count ? (float)(1.0 / *d_count) : inv_count;
        mu = __ldg(reinterpret_cast<const float4 *>(mean) + tx); is = __ldg(reinterpret_cast<const float4 *>(invstd) + tx);
        m1 = __ldg(reinterpret_cast<const float4 *>(sums) + tx); m2 = __ldg(reinterpret_cast<const float4 *>(sums + C) + tx);
        m1.x *= ic; m1.y *= ic; m1.z *= ic; m1.w *= ic;
        m2.x *= ic; m2.y *= ic; m2.z *= ic; m2.w *= ic;
    }
    for (int64_t r = (int64_t)blockIdx.x * ry + ty; r < rows; r += (int64_t)gridDim.x * ry) {
        const float4 d = __ldg(reinterpret_cast<const float4 *>(dA + r * C) + tx);
        const float4 v = __ldg(reinterpret_cast<const float4 *>(y + r * C) + tx);
        float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (res) rr = __ldg(reinterpret_cast<const float4 *>(res + r * C) + tx);
        const float4 dz = make_float4(ba_dz(d.x, v.x, sc.x, sh.x, act, rr.x), ba_dz(d.y, v.y, sc.y, sh.y, act, rr.y),
                                      ba_dz(d.z, v.z, sc.z, sh.z, act, rr.z), ba_dz(d.w, v.w, sc.w, sh.w, act, rr.w));
        if (dR) *(reinterpret_cast<float4 *>(dR + r * C) + tx) = dz;        // gradient of the pre-activation residual
        float4 o;
        o.x = sc.x * (dz.x - m1.x - (v.x - mu.x) * is.x * m2.x);
        o.y = sc.y * (dz.y - m1.y - (v.y - mu.y) * is.y * m2.y);
        o.z = sc.z * (dz.z - m1.z - (v.z - mu.z) * is.z * m2.z);
        o.w = sc.w * (dz.w - m1.w - (v.w - mu.w) * is.w * m2.w);
        if (dX) *(reinterpret_cast<float4 *>(dX + r * C) + tx) = o;
    }
}

static int ba_stream_blocks(int64_t rows, int ry) {
    int64_t b = (rows + ry - 1) / ry;
    if (b > (int64_t)kNumSMs * 8) b = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace pcfb

using namespace pcfb;

extern "C" int pcfb_bn_supported(int C) { return C >= 4 && C <= 1024 && (C & 3) == 0; }

extern "C" size_t pcfb_bn_workspace(int64_t rows, int C)
{
    if (!pcfb_bn_supported(C)) return 0;
    return align_up((size_t)(ba_geom(rows, C).blocks + 1) * 2 * C * sizeof(float), 256);
}

extern "C" int pcfb_bn_stats(const float *x, int64_t rows, int C, const float *pivot, float *partial, size_t partial_bytes,
                             int *nblocks, void *stream)
{
    PCFB_REQUIRE(pcfb_bn_supported(C), "pcfb_bn_stats: C = %d unsupported (multiple of 4, <= 1024)", C);
    PCFB_REQUIRE(x && partial && rows >= 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)partial % 16 == 0),
                 "pcfb_bn_stats: null or misaligned pointer");
    const BaGeom g = ba_geom(rows, C);
    PCFB_REQUIRE(partial_bytes >= (size_t)g.blocks * 2 * C * sizeof(float), "pcfb_bn_stats: workspace too small");
    if (nblocks) *nblocks = g.blocks;
    if (g.blocks == 0) return PCFB_OK;
    launch_k(bn_stats_kernel, g.blocks, BA_THREADS, (size_t)2 * g.ry * g.c4 * sizeof(float4), static_cast<cudaStream_t>(stream), x, rows, C, pivot, partial, g.c4, g.ry, g.rows_per_block);
    return check_launch("bn_stats_kernel");
}

extern "C" int pcfb_bn_backward_stats(const float *dA, const float *y, int64_t rows, int C, const float *scale, const float *shift,
                                      const float *mean, const float *invstd, int act, const float *residual, float *sums, int *nblocks,
                                      void *workspace, size_t workspace_bytes, void *stream)
{
    PCFB_REQUIRE(pcfb_bn_supported(C), "pcfb_bn_backward_stats: C = %d unsupported (multiple of 4, <= 1024)", C);
    PCFB_REQUIRE(dA && y && scale && shift && mean && invstd && workspace && ((uintptr_t)dA % 16 == 0) &&
                 ((uintptr_t)y % 16 == 0), "pcfb_bn_backward_stats: null or misaligned pointer");
    const BaGeom g = ba_geom(rows, C);
    if (nblocks) *nblocks = g.blocks;
    PCFB_REQUIRE(workspace_bytes >= (size_t)g.blocks * 2 * C * sizeof(float), "pcfb_bn_backward_stats: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(workspace);
    int rc;
    if (g.blocks > 0) {
        launch_k(bn_bwd_stats_kernel, g.blocks, BA_THREADS, (size_t)2 * g.ry * g.c4 * sizeof(float4), st, dA, y, rows, C, scale, shift, mean, invstd, act, partial, g.c4, g.ry, g.rows_per_block, residual);
        if ((rc = check_launch("bn_bwd_stats_kernel"))) return rc;
    }
    if (!sums) return PCFB_OK;                                   // the caller reduces the partials itself (pcfb_bn_reduce_sums)
    launch_k(sum_partials_kernel, ceil_div(2 * C * 32, 128), 128, 0, st, partial, g.blocks, 2 * C, sums);
    return check_launch("sum_partials_kernel");
}

extern "C" int pcfb_bn_backward(const float *dA, const float *y, int64_t rows, int C, const float *scale, const float *shift,
                                const float *mean, const float *invstd, const float *sums, int act, const double *d_count,
                                const float *residual, float *dX, float *d_residual, void *stream)
{
    PCFB_REQUIRE(pcfb_bn_supported(C), "pcfb_bn_backward: C = %d unsupported (multiple of 4, <= 1024)", C);
    PCFB_REQUIRE(dA && y && scale && shift && (dX || d_residual) && (!sums || (mean && invstd)) && ((uintptr_t)dA % 16 == 0) &&
                 ((uintptr_t)y % 16 == 0) && ((uintptr_t)dX % 16 == 0) && ((uintptr_t)residual % 16 == 0) &&
                 ((uintptr_t)d_residual % 16 == 0), "pcfb_bn_backward: null or misaligned pointer");
    if (rows == 0) return PCFB_OK;
    const int c4 = C >> 2, ry = BA_THREADS / c4;
    launch_k(bn_bwd_kernel, ba_stream_blocks(rows, ry), BA_THREADS, 0, static_cast<cudaStream_t>(stream), dA, y, rows, C, scale, shift, mean, invstd, sums, act, (float)(1.0 / (double)rows), d_count, dX, c4, ry, residual, d_residual);
    return check_launch("bn_bwd_kernel");
}
