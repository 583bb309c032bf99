// BatchNorm statistics: block partials -> per-channel sums -> (SyncBatchNorm: exchange over NVLink / NVSwitch peer memory)
// -> scale / shift / running statistics, as ONE kernel (sm_100a).
//
// The reference trains with sync_bn: True (configs/configPCF_Opt_10cm.yaml:14; torch.nn.SyncBatchNorm,
// /root/reference/train_ScanNet_DDP_WarmUP.py:190-195): every one of PCF_Normal's 273 BatchNorms exchanges (sum, sum^2, count)
// in the forward and (sum dz, sum dz*xhat) in the backward -- ~540 messages of <= 3 KB per step, pure latency.  An NCCL
// all-reduce inside the captured step costs ~14 us each.  Round 1 ran three launches per exchange (sum the block partials,
// a single-CTA peer all-reduce, the BatchNorm finalize) plus an all-reduce of the row counts per step; this kernel does all
// of it in one launch and carries the row count inside the message, so nothing needs to know the pyramid's level sizes.
//
// Grid: one CTA per group of 8 channels (16 columns: sum and sum^2 / sum dz and sum dz*xhat of 8 channels).
//   1. local reduction of the block partials, fixed order, double (thread = (column, slice of the blocks));
//   2. world > 1: the CTA PUSHES its 16 sums + the local row count (17 doubles = 34 32-bit words) into its slot of every
//      rank's symmetric buffer as 8-byte {word, epoch} stores through the peer mapping -- every 8-byte store carries its own
//      validity tag (the NCCL "LL" idea), so there is no fence and no separate flag: the receiver polls each word until its
//      tag equals the call's epoch, then adds the ranks' values in rank order: the same order on every rank => bit-identical,
//      deterministic.  (Round-2 measurement: data + __threadfence_system + release flag + acquire spin cost ~10 us per
//      exchange at 2 GPUs)  CTA b of rank r talks only to CTA b of the other ranks.  No reset between
//      calls: epochs only grow, the slots are double buffered by epoch parity (a peer cannot reach use e+2 of a slot before
//      this rank has pushed use e+1, i.e. finished reading use e).  `channel` selects an independent set of slots / flags / epochs: exchanges enqueued on
//      different CUDA streams (streams.py: the branches of a layer run concurrently) use different channels, and every rank
//      issues the same sequence of exchanges per channel;
//   3. mode FINALIZE: mean / variance -> scale, shift, saved mean / invstd, running statistics (global count);
//      mode SUMS: the local sums (dgamma / dbeta stay local, as in torch.nn.SyncBatchNorm) and the global sums.
// A peer that never arrives: after `timeout_s` (host-configurable; the NCCL scale of minutes by default) the CTA sets the
// error word of its buffer, writes NaN results and RETURNS -- the context survives, the failure is loud.
//
// Symmetric buffer layout per rank (bytes): [0,4) error word | [1024, +NCH*MAXB*4) epochs |
// [65536, ...) slots[NCH][2][MAXB][world][34] of {uint32 word, uint32 epoch}.
#include "common.cuh"
#include <stdlib.h>
#include "act.cuh"

namespace pcfb {

constexpr int SB_NCH = 4;            // exchange channels (main stream + 3 side streams)
constexpr int SB_MAXB = 128;         // CTAs per exchange = channel groups of 8 -> C <= 1024
constexpr int SB_MAXW = 16;          // ranks
constexpr int SB_MSG = 17;           // 16 column sums + the row count (doubles)
constexpr int SB_WORDS = 2 * SB_MSG; // 32-bit words per message, each sent as an 8-byte {word, epoch} store
constexpr int SB_EPOCH_OFF = 1024;
constexpr int SB_DATA_OFF = 65536;
constexpr int SB_THREADS = 256;
constexpr int SB_MAXT = 1024;        // bn_reduce_kernel with many block partials

__device__ __forceinline__ void sb_st_ll(unsigned long long *p, unsigned long long v) {   // one 8-byte store: single-copy atomic
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;\n" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long sb_ld_ll(const unsigned long long *p) {     // never from L1: written by the peers
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long sb_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}

struct SbXchg {                      // the exchange side of a call: null bases / world <= 1 = no exchange
    const unsigned long long *bases;
    int rank, world, channel;
    unsigned long long timeout_ns;
    __device__ __forceinline__ bool on() const { return bases && world > 1; }
};

// thread 0 at kernel start: this use's epoch of (channel, group) -- the load overlaps the local reduction
__device__ __forceinline__ uint32_t sb_epoch_begin(const SbXchg &x, int group) {
    uint32_t *ctr = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(x.bases[x.rank]) + SB_EPOCH_OFF) +
                    x.channel * SB_MAXB + group;
    const uint32_t ep = *ctr + 1;
    *ctr = ep;
    return ep;
}

// Whole CTA (any block size >= 32).  msg_s[SB_MSG] (this rank's 16 sums + row count) must be visible to the block and
// *timed_out_s zero; on return glob_s[0..15] (and glob_s[16] when take_count) hold the sums over all ranks, added in rank
// order, visible to the block.
__device__ __forceinline__ void sb_exchange(const SbXchg &x, int group, uint32_t ep, const double *msg_s, double *glob_s,
                                            bool take_count, uint32_t (*recv_s)[SB_WORDS], int *timed_out_s)
{
    const int t = threadIdx.x;
    unsigned char *mine = reinterpret_cast<unsigned char *>(x.bases[x.rank]);
    const size_t par_off = SB_DATA_OFF + ((((size_t)x.channel * 2 + (ep & 1u)) * SB_MAXB + group) * x.world) * SB_WORDS * sizeof(unsigned long long);
    const uint32_t *msg_w = reinterpret_cast<const uint32_t *>(msg_s);
    for (int j = t; j < SB_WORDS * x.world; j += (int)blockDim.x) {    // push: my words (+ the epoch tag) into my slot of every rank's buffer
        const int q = j / SB_WORDS, i = j - q * SB_WORDS;
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(x.bases[q]) + par_off) +
                                  (size_t)x.rank * SB_WORDS + i;
        sb_st_ll(dst, ((unsigned long long)ep << 32) | msg_w[i]);
    }
    for (int j = t; j < SB_WORDS * x.world; j += (int)blockDim.x) {    // receive: poll every word of every rank until its tag is this epoch
        const int q = j / SB_WORDS, i = j - q * SB_WORDS;
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(mine + par_off) + (size_t)q * SB_WORDS + i;
        const unsigned long long t0 = sb_now_ns();
        unsigned long long v = sb_ld_ll(src);
        while ((uint32_t)(v >> 32) != ep) {
            if (x.timeout_ns && sb_now_ns() - t0 > x.timeout_ns) {
                *reinterpret_cast<volatile uint32_t *>(mine) = 1u;            // error word: read by the host (fused_mlp.peer_error)
                *timed_out_s = 1;
                if (i == 0)
                    printf("pcfb SyncBatchNorm exchange: rank %d gave up waiting for rank %d (channel %d, cta %d, epoch %u)\n",
                           x.rank, q, x.channel, group, ep);
                break;
            }
            v = sb_ld_ll(src);
        }
        recv_s[q][i] = (uint32_t)v;
    }
    __syncthreads();
    if (t < SB_MSG) {
        double v = 0.0;
        for (int q = 0; q < x.world; ++q) {                       // rank order: identical everywhere
            const unsigned long long bits = ((unsigned long long)recv_s[q][2 * t + 1] << 32) | recv_s[q][2 * t];
            v += __longlong_as_double((long long)bits);
        }
        if (*timed_out_s) v = __longlong_as_double(0x7ff8000000000000ll);
        if (t < 16 || take_count) glob_s[t] = v;
    }
    __syncthreads();
}

// scale / shift / saved statistics / running statistics of channel ch from the global sums of (y - pivot), (y - pivot)^2
struct SbFinalize {
    const float *pivot, *gamma, *beta;
    float eps, momentum;             // momentum < 0: cumulative moving average, 1 / num_batches_tracked (already incremented)
    float *running_mean, *running_var, *scale, *shift, *mean, *invstd;
    long long *batches_tracked;      // incremented by the kernel when momentum >= 0
};
__device__ __forceinline__ void sb_finalize_channel(const SbFinalize &f, int ch, double s1, double s2, double count, float &sc, float &sh)
{
    const double m_p = s1 / count;                                // mean of (y - pivot)
    double var = s2 / count - m_p * m_p;
    if (var < 0.0) var = 0.0;
    const double mean = m_p + (f.pivot ? (double)f.pivot[ch] : 0.0);
    const float invstd = (float)(1.0 / sqrt(var + (double)f.eps));
    const float g = f.gamma ? f.gamma[ch] : 1.f, bt = f.beta ? f.beta[ch] : 0.f;
    sc = g * invstd;
    sh = bt - (float)mean * g * invstd;
    f.scale[ch] = sc;
    f.shift[ch] = sh;
    if (f.mean) f.mean[ch] = (float)mean;
    if (f.invstd) f.invstd[ch] = invstd;
    if (f.running_mean) {
        float mom = f.momentum;
        if (mom < 0.f) mom = f.batches_tracked ? 1.f / (float)(*f.batches_tracked) : 1.f;     // cumulative moving average
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        f.running_mean[ch] = (1.f - mom) * f.running_mean[ch] + mom * (float)mean;
        f.running_var[ch] = (1.f - mom) * f.running_var[ch] + mom * (float)unbiased;
    }
}

struct SbArgs {
    const float *partial;            // [nblocks][2][C]
    int nblocks, C, mode;            // mode 0 = finalize, 1 = sums
    double count;                    // local row count
    const double *d_count;           // optional: global row count already known on the device (NCCL fallback path)
    SbXchg x;
    SbFinalize f;
    double *count_out;
    float *sums_local, *sums_global; // [2][C] (mode 1)
};

// blockDim.x = 256 (16 slices of the block partials per column) or 1024 (64 slices: level-0 tensors leave ~800 partials per
// column, and a 50-deep chain of dependent L2 loads per thread was most of this kernel's 5.6 us)
__global__ void __launch_bounds__(SB_MAXT)
bn_reduce_kernel(SbArgs a)
{
    pdl_wait();
    __shared__ double part_s[SB_MAXT / 16][17];   // [slice][column]
    __shared__ double msg_s[SB_MSG], glob_s[SB_MSG];
    __shared__ uint32_t recv_s[SB_MAXW][SB_WORDS];
    __shared__ uint32_t ep_s;
    __shared__ int timed_out_s;
    const int t = threadIdx.x, col = t & 15, slice = t >> 4, n_slices = (int)blockDim.x >> 4;
    const int c0 = blockIdx.x * 8;
    const int which = col >> 3, c = c0 + (col & 7);
    const bool exchange = a.x.on();
    if (exchange && t == 0) ep_s = sb_epoch_begin(a.x, blockIdx.x);
    // 1. local reduction, fixed order: slice s adds blocks s, s + n_slices, ... ; slices are then added 0 .. n_slices-1
    double s = 0.0;
    if (c < a.C) {
        const float *src = a.partial + (size_t)which * a.C + c;
        const size_t stride = (size_t)2 * a.C;
        int b = slice;
        for (; b + 3 * n_slices < a.nblocks; b += 4 * n_slices) {      // four independent loads in flight
            const float v0 = src[(size_t)b * stride], v1 = src[(size_t)(b + n_slices) * stride];
            const float v2 = src[(size_t)(b + 2 * n_slices) * stride], v3 = src[(size_t)(b + 3 * n_slices) * stride];
            s += (double)v0; s += (double)v1; s += (double)v2; s += (double)v3;
        }
        for (; b < a.nblocks; b += n_slices) s += (double)src[(size_t)b * stride];
    }
    part_s[slice][col] = s;
    if (t == 0) timed_out_s = 0;
    __syncthreads();
    if (t < 16) {
        double v = 0.0;
        for (int k = 0; k < n_slices; ++k) v += part_s[k][t];
        msg_s[t] = v;
        glob_s[t] = v;
    }
    if (t == 16) { msg_s[16] = a.count; glob_s[16] = a.d_count ? *a.d_count : a.count; }
    __syncthreads();
    // 2. exchange
    if (exchange) sb_exchange(a.x, blockIdx.x, ep_s, msg_s, glob_s, !a.d_count, recv_s, &timed_out_s);
    // 3. outputs
    if (a.mode == 1) {
        if (t < 16 && c < a.C) {
            if (a.sums_local) a.sums_local[(size_t)which * a.C + c] = (float)msg_s[t];
            if (a.sums_global) a.sums_global[(size_t)which * a.C + c] = (float)glob_s[t];
        }
        return;
    }
    const double count = glob_s[16];
    if (blockIdx.x == 0 && t == 0) {
        if (a.count_out) *a.count_out = count;
        if (a.f.batches_tracked && a.f.momentum >= 0.f) *a.f.batches_tracked += 1;
    }
    if (t < 8 && c0 + t < a.C) {
        float sc, sh;
        sb_finalize_channel(a.f, c0 + t, glob_s[t], glob_s[8 + t], count, sc, sh);
    }
}

// ------------------------------------------------------------------------------------------------------------
// BatchNorm (+ residual + activation) of a SMALL tensor as ONE kernel per direction.
//
// Statistics are per channel, so a CTA that owns a group of 8 channels for ALL rows needs nobody else: it sums its
// columns, (SyncBatchNorm: exchanges the 17 doubles with the same CTA of the other ranks,) finalizes and applies -- one
// launch instead of statistics -> bn_reduce -> apply (three launches and two dependencies per BatchNorm; 19 of
// PCF_Normal's 25 layers live on levels of <= 5 k points where each of those launches is a 4-6 us bubble on the critical
// path of the step: profiles/step_timeline_r02.txt).  The second pass re-reads the CTA's 32-byte column slice of every
// row from L1 / L2.  Thread = (row lane 0..127, half 0..1): a float4 of each row it visits.
// ------------------------------------------------------------------------------------------------------------
constexpr int SBS_MAXT = 1024;                    // threads per CTA: 256, or 1024 above 4 k rows (row lanes = threads / 2)

// 8 column sums (float4 s1 | float4 s2 of this thread's half) over the block's 128 row lanes -> msg_s[0..15] in double,
// fixed order: xor-shuffles inside the warp, then the 8 warps added 0..7
__device__ __forceinline__ void sbs_block_sums(float4 s1, float4 s2, double *msg_s, float (*red_s)[2][8])
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    float v[8] = {s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w};
#pragma unroll
    for (int off = 2; off < 32; off <<= 1)
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
    if (lane < 2)
#pragma unroll
        for (int k = 0; k < 8; ++k) red_s[warp][lane][k] = v[k];
    __syncthreads();
    if (t < 16) {                                 // column t: which = t >> 3 (s1 | s2), channel t & 7 = half * 4 + component
        const int which = t >> 3, ch = t & 7, n_warps = (int)blockDim.x >> 5;
        double acc = 0.0;
        for (int w = 0; w < n_warps; ++w) acc += (double)red_s[w][ch >> 2][which * 4 + (ch & 3)];
        msg_s[t] = acc;
    }
}

struct BnSmallFwd {
    const float *x, *res;
    float *out;
    long long rows;
    int C, act, res_after;
    SbFinalize f;
    double *count_out;
    SbXchg xc;
};

// T threads per CTA, SBS_U rows per load batch
template <int T, int SBS_U>
__global__ void __launch_bounds__(T)
bn_small_fwd_kernel(BnSmallFwd a)
{
    pdl_wait();
    __shared__ double msg_s[SB_MSG], glob_s[SB_MSG];
    __shared__ uint32_t recv_s[SB_MAXW][SB_WORDS];
    __shared__ float red_s[SBS_MAXT / 32][2][8];
    __shared__ float sc_s[8], sh_s[8];
    __shared__ uint32_t ep_s;
    __shared__ int timed_out_s;
    constexpr int SBS_RL = T / 2;                 // row lanes
    const int t = threadIdx.x, half = t & 1, rl = t >> 1;
    const int c0 = blockIdx.x * 8, c = c0 + 4 * half;
    const bool valid = c < a.C;                    // C is a multiple of 4: a float4 is inside or outside as a whole
    const bool exchange = a.xc.on();
    if (exchange && t == 0) ep_s = sb_epoch_begin(a.xc, blockIdx.x);
    if (t == 0) timed_out_s = 0;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    if (valid) {
        const float4 pv = a.f.pivot ? make_float4(__ldg(a.f.pivot + c), __ldg(a.f.pivot + c + 1), __ldg(a.f.pivot + c + 2), __ldg(a.f.pivot + c + 3))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        // SBS_U rows per batch: the loads of a batch are all in flight before the first is used (a CTA walks every row of
        // its channels, so the trip count is rows / 128 and an un-batched loop is a chain of L2 latencies)
        for (long long r0 = rl; r0 < a.rows; r0 += (long long)SBS_U * SBS_RL) {
            float4 vb[SBS_U];
#pragma unroll
            for (int u = 0; u < SBS_U; ++u) {
                const long long r = r0 + (long long)u * SBS_RL;
                vb[u] = r < a.rows ? __ldg(reinterpret_cast<const float4 *>(a.x + r * a.C + c)) : pv;    // pv - pv = 0: adds nothing
            }
#pragma unroll
            for (int u = 0; u < SBS_U; ++u) {
                float4 v = vb[u];
                v.x -= pv.x; v.y -= pv.y; v.z -= pv.z; v.w -= pv.w;
                s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
                s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
            }
        }
    }
    sbs_block_sums(s1, s2, msg_s, red_s);
    if (t == 16) msg_s[16] = (double)a.rows;
    __syncthreads();
    if (t < SB_MSG) glob_s[t] = msg_s[t];
    __syncthreads();
    if (exchange) sb_exchange(a.xc, blockIdx.x, ep_s, msg_s, glob_s, true, recv_s, &timed_out_s);
    const double count = glob_s[16];
    if (blockIdx.x == 0 && t == 0) {
        if (a.count_out) *a.count_out = count;
        if (a.f.batches_tracked && a.f.momentum >= 0.f) *a.f.batches_tracked += 1;
    }
    if (t < 8 && c0 + t < a.C) sb_finalize_channel(a.f, c0 + t, glob_s[t], glob_s[8 + t], count, sc_s[t], sh_s[t]);
    __syncthreads();
    if (!valid) return;
    const float4 sc = make_float4(sc_s[4 * half], sc_s[4 * half + 1], sc_s[4 * half + 2], sc_s[4 * half + 3]);
    const float4 sh = make_float4(sh_s[4 * half], sh_s[4 * half + 1], sh_s[4 * half + 2], sh_s[4 * half + 3]);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long r0 = rl; r0 < a.rows; r0 += (long long)SBS_U * SBS_RL) {
        float4 vb[SBS_U], rb[SBS_U];
#pragma unroll
        for (int u = 0; u < SBS_U; ++u) {
            const long long r = r0 + (long long)u * SBS_RL;
            vb[u] = r < a.rows ? __ldg(reinterpret_cast<const float4 *>(a.x + r * a.C + c)) : zero4;
            rb[u] = (a.res && r < a.rows) ? __ldg(reinterpret_cast<const float4 *>(a.res + r * a.C + c)) : zero4;
        }
#pragma unroll
        for (int u = 0; u < SBS_U; ++u) {
            const long long r = r0 + (long long)u * SBS_RL;
            if (r >= a.rows) break;
            const float4 v = vb[u], rr = rb[u];
            const float4 pre = a.res_after ? zero4 : rr, post = a.res_after ? rr : zero4;
            float4 o;
            o.x = act_fwd(fmaf(v.x, sc.x, sh.x) + pre.x, a.act) + post.x; o.y = act_fwd(fmaf(v.y, sc.y, sh.y) + pre.y, a.act) + post.y;
            o.z = act_fwd(fmaf(v.z, sc.z, sh.z) + pre.z, a.act) + post.z; o.w = act_fwd(fmaf(v.w, sc.w, sh.w) + pre.w, a.act) + post.w;
            *reinterpret_cast<float4 *>(a.out + r * a.C + c) = o;
        }
    }
}

struct BnSmallBwd {
    const float *dA, *x, *res, *scale, *shift, *mean, *invstd;
    const double *d_count;           // global row count (SyncBatchNorm), else null: rows
    float *sums_local, *dX, *dR;
    long long rows;
    int C, act;
    SbXchg xc;
};

__device__ __forceinline__ float sbs_dz(float d, float yv, float sc, float sh, int act, float r) {
    const float z = fmaf(yv, sc, sh) + r;         // r: the residual added before the activation (0 otherwise)
    return d * act_bwd(z, act_fwd(z, act), act);
}

// dz = dA * act'(x*scale+shift (+res)), xhat = (x - mean) * invstd;  S1 = sum dz, S2 = sum dz*xhat over the GLOBAL batch;
// dX = scale * (dz - S1/E - xhat * S2/E), dR = dz; sums_local = this rank's (S1 | S2) = (dbeta | dgamma)
template <int T, int SBS_UB>
__global__ void __launch_bounds__(T)
bn_small_bwd_kernel(BnSmallBwd a)
{
    pdl_wait();
    __shared__ double msg_s[SB_MSG], glob_s[SB_MSG];
    __shared__ uint32_t recv_s[SB_MAXW][SB_WORDS];
    __shared__ float red_s[SBS_MAXT / 32][2][8];
    __shared__ uint32_t ep_s;
    __shared__ int timed_out_s;
    constexpr int SBS_RL = T / 2;                 // row lanes
    const int t = threadIdx.x, half = t & 1, rl = t >> 1;
    const int c0 = blockIdx.x * 8, c = c0 + 4 * half;
    const bool valid = c < a.C;
    const bool exchange = a.xc.on();
    if (exchange && t == 0) ep_s = sb_epoch_begin(a.xc, blockIdx.x);
    if (t == 0) timed_out_s = 0;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    float4 sc = s1, sh = s1, mu = s1, is = s1;
    if (valid) {
        sc = __ldg(reinterpret_cast<const float4 *>(a.scale + c)); sh = __ldg(reinterpret_cast<const float4 *>(a.shift + c));
        mu = __ldg(reinterpret_cast<const float4 *>(a.mean + c)); is = __ldg(reinterpret_cast<const float4 *>(a.invstd + c));
        for (long long r0 = rl; r0 < a.rows; r0 += (long long)SBS_UB * SBS_RL) {
          float4 db[SBS_UB], vb[SBS_UB], rb[SBS_UB];
#pragma unroll
          for (int u = 0; u < SBS_UB; ++u) {
              const long long r = r0 + (long long)u * SBS_RL;
              const bool in = r < a.rows;                        // past the end: dA = 0 -> dz = 0 adds nothing
              db[u] = in ? __ldg(reinterpret_cast<const float4 *>(a.dA + r * a.C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
              vb[u] = in ? __ldg(reinterpret_cast<const float4 *>(a.x + r * a.C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
              rb[u] = (in && a.res) ? __ldg(reinterpret_cast<const float4 *>(a.res + r * a.C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < SBS_UB; ++u) {
            const float4 d = db[u], v = vb[u], rr = rb[u];
            const float dx_ = sbs_dz(d.x, v.x, sc.x, sh.x, a.act, rr.x), dy_ = sbs_dz(d.y, v.y, sc.y, sh.y, a.act, rr.y);
            const float dz_ = sbs_dz(d.z, v.z, sc.z, sh.z, a.act, rr.z), dw_ = sbs_dz(d.w, v.w, sc.w, sh.w, a.act, rr.w);
            s1.x += dx_; s1.y += dy_; s1.z += dz_; s1.w += dw_;
            s2.x = fmaf(dx_, (v.x - mu.x) * is.x, s2.x); s2.y = fmaf(dy_, (v.y - mu.y) * is.y, s2.y);
            s2.z = fmaf(dz_, (v.z - mu.z) * is.z, s2.z); s2.w = fmaf(dw_, (v.w - mu.w) * is.w, s2.w);
          }
        }
    }
    sbs_block_sums(s1, s2, msg_s, red_s);
    if (t == 16) msg_s[16] = 0.0;
    __syncthreads();
    if (t < SB_MSG) glob_s[t] = msg_s[t];
    __syncthreads();
    if (exchange) sb_exchange(a.xc, blockIdx.x, ep_s, msg_s, glob_s, false, recv_s, &timed_out_s);
    if (t < 16 && a.sums_local) {
        const int which = t >> 3, ch = c0 + (t & 7);
        if (ch < a.C) a.sums_local[(size_t)which * a.C + ch] = (float)msg_s[t];
    }
    if (!valid || !(a.dX || a.dR)) return;
    const float ic = a.d_count ? (float)(1.0 / *a.d_count) : (float)(1.0 / (double)a.rows);
    float4 m1, m2;
    m1.x = (float)glob_s[4 * half] * ic; m1.y = (float)glob_s[4 * half + 1] * ic; m1.z = (float)glob_s[4 * half + 2] * ic; m1.w = (float)glob_s[4 * half + 3] * ic;
    m2.x = (float)glob_s[8 + 4 * half] * ic; m2.y = (float)glob_s[8 + 4 * half + 1] * ic; m2.z = (float)glob_s[8 + 4 * half + 2] * ic; m2.w = (float)glob_s[8 + 4 * half + 3] * ic;
    for (long long r0 = rl; r0 < a.rows; r0 += (long long)SBS_UB * SBS_RL) {
      float4 db[SBS_UB], vb[SBS_UB], rb[SBS_UB];
#pragma unroll
      for (int u = 0; u < SBS_UB; ++u) {
          const long long r = r0 + (long long)u * SBS_RL;
          const bool in = r < a.rows;
          db[u] = in ? __ldg(reinterpret_cast<const float4 *>(a.dA + r * a.C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          vb[u] = in ? __ldg(reinterpret_cast<const float4 *>(a.x + r * a.C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          rb[u] = (in && a.res) ? __ldg(reinterpret_cast<const float4 *>(a.res + r * a.C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < SBS_UB; ++u) {
        const long long r = r0 + (long long)u * SBS_RL;
        if (r >= a.rows) break;
        const float4 d = db[u], v = vb[u], rr = rb[u];
        const float4 dz = make_float4(sbs_dz(d.x, v.x, sc.x, sh.x, a.act, rr.x), sbs_dz(d.y, v.y, sc.y, sh.y, a.act, rr.y),
                                      sbs_dz(d.z, v.z, sc.z, sh.z, a.act, rr.z), sbs_dz(d.w, v.w, sc.w, sh.w, a.act, rr.w));
        if (a.dR) *reinterpret_cast<float4 *>(a.dR + r * a.C + c) = dz;
        if (a.dX) {
            float4 o;
            o.x = sc.x * (dz.x - m1.x - (v.x - mu.x) * is.x * m2.x);
            o.y = sc.y * (dz.y - m1.y - (v.y - mu.y) * is.y * m2.y);
            o.z = sc.z * (dz.z - m1.z - (v.z - mu.z) * is.z * m2.z);
            o.w = sc.w * (dz.w - m1.w - (v.w - mu.w) * is.w * m2.w);
            *reinterpret_cast<float4 *>(a.dX + r * a.C + c) = o;
        }
      }
    }
}

}  // namespace pcfb

using namespace pcfb;

extern "C" size_t pcfb_syncbn_buffer_bytes(int world)
{
    if (world < 1 || world > SB_MAXW) return 0;
    return SB_DATA_OFF + (size_t)SB_NCH * 2 * SB_MAXB * world * SB_WORDS * sizeof(unsigned long long);
}

extern "C" int pcfb_syncbn_channels(void) { return SB_NCH; }

static int sb_check(int C, const void *peer_bases, int rank, int world, int channel, const char *what)
{
    PCFB_REQUIRE(C >= 1 && C <= 8 * SB_MAXB, "%s: C = %d outside [1, %d]", what, C, 8 * SB_MAXB);
    PCFB_REQUIRE(!peer_bases || (world >= 1 && world <= SB_MAXW && rank >= 0 && rank < world && channel >= 0 && channel < SB_NCH),
                 "%s: bad rank %d / world %d / channel %d", what, rank, world, channel);
    return PCFB_OK;
}

static int sb_wide_min() {                           // block partials above which bn_reduce runs with 64 slices (PCFB_BNR_WIDE_MIN)
    static int v = -1;
    if (v < 0) { const char *e = getenv("PCFB_BNR_WIDE_MIN"); v = e ? atoi(e) : 128; }
    return v;
}
static unsigned long long sb_timeout(double timeout_s) { return timeout_s > 0.0 ? (unsigned long long)(timeout_s * 1e9) : 0ull; }

static SbXchg sb_xchg(const void *peer_bases, int rank, int world, int channel, double timeout_s)
{
    SbXchg x;
    x.bases = static_cast<const unsigned long long *>(peer_bases); x.rank = rank; x.world = peer_bases ? world : 1; x.channel = channel;
    x.timeout_ns = sb_timeout(timeout_s);
    return x;
}

extern "C" int pcfb_bn_finalize(const float *partial, int nblocks, int C, int64_t count, const double *d_count, const float *pivot,
                                const float *gamma, const float *beta, float eps, float momentum, float *running_mean,
                                float *running_var, float *scale, float *shift, float *mean, float *invstd,
                                int64_t *batches_tracked, double *count_out, const void *peer_bases, int rank, int world,
                                int channel, double timeout_s, void *stream)
{
    PCFB_REQUIRE(partial && scale && shift && nblocks >= 0, "pcfb_bn_finalize: null pointer");
    int rc = sb_check(C, peer_bases, rank, world, channel, "pcfb_bn_finalize");
    if (rc) return rc;
    SbArgs a{};
    a.partial = partial; a.nblocks = nblocks; a.C = C; a.mode = 0; a.count = (double)count; a.d_count = d_count;
    a.x = sb_xchg(peer_bases, rank, world, channel, timeout_s);
    a.f.pivot = pivot; a.f.gamma = gamma; a.f.beta = beta; a.f.eps = eps; a.f.momentum = momentum;
    a.f.running_mean = running_mean; a.f.running_var = running_var; a.f.scale = scale; a.f.shift = shift; a.f.mean = mean; a.f.invstd = invstd;
    a.f.batches_tracked = reinterpret_cast<long long *>(batches_tracked); a.count_out = count_out;
    launch_k(bn_reduce_kernel, ceil_div(C, 8), nblocks > sb_wide_min() ? SB_MAXT : SB_THREADS, 0, static_cast<cudaStream_t>(stream), a);
    return check_launch("bn_reduce_kernel<finalize>");
}

extern "C" int pcfb_bn_reduce_sums(const float *partial, int nblocks, int C, float *sums_local, float *sums_global,
                                   const void *peer_bases, int rank, int world, int channel, double timeout_s, void *stream)
{
    PCFB_REQUIRE(partial && (sums_local || sums_global) && nblocks >= 0, "pcfb_bn_reduce_sums: null pointer");
    int rc = sb_check(C, peer_bases, rank, world, channel, "pcfb_bn_reduce_sums");
    if (rc) return rc;
    SbArgs a{};
    a.partial = partial; a.nblocks = nblocks; a.C = C; a.mode = 1; a.count = 0.0;
    a.x = sb_xchg(peer_bases, rank, world, channel, timeout_s);
    a.sums_local = sums_local; a.sums_global = sums_global;
    launch_k(bn_reduce_kernel, ceil_div(C, 8), nblocks > sb_wide_min() ? SB_MAXT : SB_THREADS, 0, static_cast<cudaStream_t>(stream), a);
    return check_launch("bn_reduce_kernel<sums>");
}

// rows up to which the one-kernel BatchNorm beats statistics -> reduce -> apply.  Each CTA walks ALL rows of its 8
// channels, i.e. C/8 SMs pull the tensor through their own L2 ports: measured on the 10 cm pyramid, the 184- and 1 k-point
// levels gain 0.65 ms per step, the 5 k-point level loses 0.4 ms (scripts/bn_small_sweep.sh, profiles/README.md).
extern "C" int pcfb_bn_small_max_rows(void) { return 2048; }

extern "C" int pcfb_bn_small_forward(const float *x, int64_t rows, int C, const float *pivot, const float *gamma, const float *beta,
                                     float eps, float momentum, float *running_mean, float *running_var, int64_t *batches_tracked,
                                     int act, const float *residual, int residual_after_act, float *out, float *scale, float *shift,
                                     float *mean, float *invstd, double *count_out, const void *peer_bases, int rank, int world,
                                     int channel, double timeout_s, void *stream)
{
    PCFB_REQUIRE(x && out && scale && shift && rows >= 1, "pcfb_bn_small_forward: null pointer or no rows");
    PCFB_REQUIRE(C >= 4 && (C & 3) == 0, "pcfb_bn_small_forward: C = %d must be a multiple of 4", C);
    PCFB_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)residual % 16 == 0),
                 "pcfb_bn_small_forward: misaligned pointer");
    int rc = sb_check(C, peer_bases, rank, world, channel, "pcfb_bn_small_forward");
    if (rc) return rc;
    BnSmallFwd a{};
    a.x = x; a.res = residual; a.out = out; a.rows = rows; a.C = C; a.act = act; a.res_after = residual_after_act;
    a.f.pivot = pivot; a.f.gamma = gamma; a.f.beta = beta; a.f.eps = eps; a.f.momentum = momentum;
    a.f.running_mean = running_mean; a.f.running_var = running_var; a.f.scale = scale; a.f.shift = shift; a.f.mean = mean; a.f.invstd = invstd;
    a.f.batches_tracked = reinterpret_cast<long long *>(batches_tracked); a.count_out = count_out;
    a.xc = sb_xchg(peer_bases, rank, world, channel, timeout_s);
    // 256 threads (128 row lanes, 8 rows per load batch); 1024 threads (64 registers: batches of 4) only for callers that
    // force the one-kernel path on larger tensors (PCFB_BN_SMALL_ROWS): at 1 k rows the big CTA costs 9 us against 5
    if (rows > 4096) launch_k(bn_small_fwd_kernel<SBS_MAXT, 4>, ceil_div(C, 8), SBS_MAXT, 0, static_cast<cudaStream_t>(stream), a);
    else launch_k(bn_small_fwd_kernel<SB_THREADS, 8>, ceil_div(C, 8), SB_THREADS, 0, static_cast<cudaStream_t>(stream), a);
    return check_launch("bn_small_fwd_kernel");
}

extern "C" int pcfb_bn_small_backward(const float *dA, const float *x, int64_t rows, int C, const float *scale, const float *shift,
                                      const float *mean, const float *invstd, int act, const float *residual, const double *d_count,
                                      float *sums_local, float *dX, float *d_residual, const void *peer_bases, int rank, int world,
                                      int channel, double timeout_s, void *stream)
{
    PCFB_REQUIRE(dA && x && scale && shift && mean && invstd && rows >= 1, "pcfb_bn_small_backward: null pointer or no rows");
    PCFB_REQUIRE(C >= 4 && (C & 3) == 0, "pcfb_bn_small_backward: C = %d must be a multiple of 4", C);
    PCFB_REQUIRE(((uintptr_t)dA % 16 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dX % 16 == 0) && ((uintptr_t)residual % 16 == 0) &&
                 ((uintptr_t)d_residual % 16 == 0) && ((uintptr_t)scale % 16 == 0) && ((uintptr_t)shift % 16 == 0) &&
                 ((uintptr_t)mean % 16 == 0) && ((uintptr_t)invstd % 16 == 0), "pcfb_bn_small_backward: misaligned pointer");
    int rc = sb_check(C, peer_bases, rank, world, channel, "pcfb_bn_small_backward");
    if (rc) return rc;
    BnSmallBwd a{};
    a.dA = dA; a.x = x; a.res = residual; a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd; a.d_count = d_count;
    a.sums_local = sums_local; a.dX = dX; a.dR = d_residual; a.rows = rows; a.C = C; a.act = act;
    a.xc = sb_xchg(peer_bases, rank, world, channel, timeout_s);
    if (rows > 4096) launch_k(bn_small_bwd_kernel<SBS_MAXT, 2>, ceil_div(C, 8), SBS_MAXT, 0, static_cast<cudaStream_t>(stream), a);
    else launch_k(bn_small_bwd_kernel<SB_THREADS, 4>, ceil_div(C, 8), SB_THREADS, 0, static_cast<cudaStream_t>(stream), a);
    return check_launch("bn_small_bwd_kernel");
}
