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

namespace pcfb {

constexpr int SB_NCH = 4;            // exchange channels (main stream + 3 side streams)
constexpr int SB_MAXB = 128;         // CTAs per exchange = channel groups of 8 -> C <= 1024
constexpr int SB_MAXW = 16;          // ranks
constexpr int SB_MSG = 17;           // 16 column sums + the row count (doubles)
constexpr int SB_WORDS = 2 * SB_MSG; // 32-bit words per message, each sent as an 8-byte {word, epoch} store
constexpr int SB_EPOCH_OFF = 1024;
constexpr int SB_DATA_OFF = 65536;
constexpr int SB_THREADS = 256;

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

struct SbArgs {
    const float *partial;            // [nblocks][2][C]
    int nblocks, C, mode;            // mode 0 = finalize, 1 = sums
    double count;                    // local row count
    const double *d_count;           // optional: global row count already known on the device (NCCL fallback path)
    // exchange
    const unsigned long long *bases; // null / world <= 1: no exchange
    int rank, world, channel;
    unsigned long long timeout_ns;
    // finalize
    const float *pivot, *gamma, *beta;
    float eps, momentum;             // momentum < 0: cumulative moving average, 1 / num_batches_tracked (already incremented)
    float *running_mean, *running_var, *scale, *shift, *mean, *invstd;
    long long *batches_tracked;      // incremented here when momentum >= 0
    double *count_out;
    // sums
    float *sums_local, *sums_global; // [2][C]
};

__global__ void __launch_bounds__(SB_THREADS)
bn_reduce_kernel(SbArgs a)
{
    pdl_wait();
    __shared__ double part_s[16][17];             // [slice][column]
    __shared__ double msg_s[SB_MSG], glob_s[SB_MSG];
    __shared__ uint32_t ep_s;
    __shared__ int timed_out_s;
    const int t = threadIdx.x, col = t & 15, slice = t >> 4;
    const int c0 = blockIdx.x * 8;
    const int which = col >> 3, c = c0 + (col & 7);
    const bool exchange = a.bases && a.world > 1;
    if (exchange && t == 0) {                      // this use's epoch: read early, the load overlaps the local reduction
        uint32_t *ctr = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(a.bases[a.rank]) + SB_EPOCH_OFF) +
                        a.channel * SB_MAXB + blockIdx.x;
        ep_s = *ctr + 1;
        *ctr = ep_s;
    }
    // 1. local reduction, fixed order: slice s adds blocks s, s+16, ... ; slices are then added 0..15
    double s = 0.0;
    if (c < a.C) {
        const float *src = a.partial + (size_t)which * a.C + c;
        const size_t stride = (size_t)2 * a.C;
        int b = slice;
        for (; b + 48 < a.nblocks; b += 64) {      // four independent loads in flight
            const float v0 = src[(size_t)b * stride], v1 = src[(size_t)(b + 16) * stride];
            const float v2 = src[(size_t)(b + 32) * stride], v3 = src[(size_t)(b + 48) * stride];
            s += (double)v0; s += (double)v1; s += (double)v2; s += (double)v3;
        }
        for (; b < a.nblocks; b += 16) s += (double)src[(size_t)b * stride];
    }
    part_s[slice][col] = s;
    if (t == 0) timed_out_s = 0;
    __syncthreads();
    if (t < 16) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) v += part_s[k][t];
        msg_s[t] = v;
        glob_s[t] = v;
    }
    if (t == 16) { msg_s[16] = a.count; glob_s[16] = a.d_count ? *a.d_count : a.count; }
    __syncthreads();
    // 2. exchange
    if (exchange) {
        __shared__ uint32_t recv_s[SB_MAXW][SB_WORDS];
        unsigned char *mine = reinterpret_cast<unsigned char *>(a.bases[a.rank]);
        const uint32_t ep = ep_s;                  // (written before the block barriers of step 1)
        const size_t par_off = SB_DATA_OFF + ((((size_t)a.channel * 2 + (ep & 1u)) * SB_MAXB + blockIdx.x) * a.world) * SB_WORDS * sizeof(unsigned long long);
        const uint32_t *msg_w = reinterpret_cast<const uint32_t *>(msg_s);
        for (int j = t; j < SB_WORDS * a.world; j += SB_THREADS) {    // push: my words (+ the epoch tag) into my slot of every rank's buffer
            const int q = j / SB_WORDS, i = j - q * SB_WORDS;
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(a.bases[q]) + par_off) +
                                      (size_t)a.rank * SB_WORDS + i;
            sb_st_ll(dst, ((unsigned long long)ep << 32) | msg_w[i]);
        }
        for (int j = t; j < SB_WORDS * a.world; j += SB_THREADS) {    // receive: poll every word of every rank until its tag is this epoch
            const int q = j / SB_WORDS, i = j - q * SB_WORDS;
            const unsigned long long *src = reinterpret_cast<const unsigned long long *>(mine + par_off) + (size_t)q * SB_WORDS + i;
            const unsigned long long t0 = sb_now_ns();
            unsigned long long v = sb_ld_ll(src);
            while ((uint32_t)(v >> 32) != ep) {
                if (a.timeout_ns && sb_now_ns() - t0 > a.timeout_ns) {
                    *reinterpret_cast<volatile uint32_t *>(mine) = 1u;            // error word: read by the host (fused_mlp.peer_error)
                    timed_out_s = 1;
                    if (i == 0)
                        printf("pcfb SyncBatchNorm exchange: rank %d gave up waiting for rank %d (channel %d, cta %d, epoch %u)\n",
                               a.rank, q, a.channel, (int)blockIdx.x, ep);
                    break;
                }
                v = sb_ld_ll(src);
            }
            recv_s[q][i] = (uint32_t)v;
        }
        __syncthreads();
        if (t < SB_MSG) {
            double v = 0.0;
            for (int q = 0; q < a.world; ++q) {                       // rank order: identical everywhere
                const unsigned long long bits = ((unsigned long long)recv_s[q][2 * t + 1] << 32) | recv_s[q][2 * t];
                v += __longlong_as_double((long long)bits);
            }
            if (timed_out_s) v = __longlong_as_double(0x7ff8000000000000ll);
            if (t < 16 || !a.d_count) glob_s[t] = v;
        }
        __syncthreads();
    }
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
        if (a.batches_tracked && a.momentum >= 0.f) *a.batches_tracked += 1;
    }
    if (t < 8 && c0 + t < a.C) {
        const int ch = c0 + t;
        const double m_p = glob_s[t] / count;                    // mean of (y - pivot)
        double var = glob_s[8 + t] / count - m_p * m_p;
        if (var < 0.0) var = 0.0;
        const double mean = m_p + (a.pivot ? (double)a.pivot[ch] : 0.0);
        const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
        const float g = a.gamma ? a.gamma[ch] : 1.f, bt = a.beta ? a.beta[ch] : 0.f;
        a.scale[ch] = g * invstd;
        a.shift[ch] = bt - (float)mean * g * invstd;
        if (a.mean) a.mean[ch] = (float)mean;
        if (a.invstd) a.invstd[ch] = invstd;
        if (a.running_mean) {
            float mom = a.momentum;
            if (mom < 0.f) mom = a.batches_tracked ? 1.f / (float)(*a.batches_tracked) : 1.f;     // cumulative moving average
            const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
            a.running_mean[ch] = (1.f - mom) * a.running_mean[ch] + mom * (float)mean;
            a.running_var[ch] = (1.f - mom) * a.running_var[ch] + mom * (float)unbiased;
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

static unsigned long long sb_timeout(double timeout_s) { return timeout_s > 0.0 ? (unsigned long long)(timeout_s * 1e9) : 0ull; }

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
    a.bases = static_cast<const unsigned long long *>(peer_bases); a.rank = rank; a.world = peer_bases ? world : 1; a.channel = channel;
    a.timeout_ns = sb_timeout(timeout_s);
    a.pivot = pivot; a.gamma = gamma; a.beta = beta; a.eps = eps; a.momentum = momentum;
    a.running_mean = running_mean; a.running_var = running_var; a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd;
    a.batches_tracked = reinterpret_cast<long long *>(batches_tracked); a.count_out = count_out;
    launch_k(bn_reduce_kernel, ceil_div(C, 8), SB_THREADS, 0, static_cast<cudaStream_t>(stream), a);
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
    a.bases = static_cast<const unsigned long long *>(peer_bases); a.rank = rank; a.world = peer_bases ? world : 1; a.channel = channel;
    a.timeout_ns = sb_timeout(timeout_s);
    a.sums_local = sums_local; a.sums_global = sums_global;
    launch_k(bn_reduce_kernel, ceil_div(C, 8), SB_THREADS, 0, static_cast<cudaStream_t>(stream), a);
    return check_launch("bn_reduce_kernel<sums>");
}
