// qs_exchange.cuh -- the peer-memory moment exchange as a device function, shared by xchg_merge_kernel (qs_exchange.cu) and the
// kernel that finishes the env step's observation moments (qs_vecnorm.cu: moments_final_xchg_kernel, one launch for both).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct qs_xchg {
    int device, rank, world, d, len;
    void* local;                 // this rank's buffer
    size_t bytes;
    void** peer_host;            // [world] mapped base pointers (own = local)
    void** peer_dev;             // device copy of the above
    unsigned long long* seq;     // device: step counter
    int* failed;                 // device: sticky timeout flag
    bool connected;
};

namespace qs {

// A double travels as two 8-byte words (32 data bits | 32-bit tag of the step): an aligned 8-byte store is one NVLink transaction, so a
// word is either old or complete -- the receiver polls the words themselves and no fence, flag or second round trip is needed
// (the low-latency protocol of NCCL's LL collectives).  The tag is never 0 (fresh buffers are zeroed) and differs between the two
// steps that share a parity slot.
__device__ __forceinline__ void ll_store(unsigned long long* p, double v, uint32_t tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | t), "l"((b >> 32) | t) : "memory");
}
__device__ __forceinline__ bool ll_load(const unsigned long long* p, uint32_t tag, double& v) {
    unsigned long long a, b;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    v = __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
    return (uint32_t)(a >> 32) == tag && (uint32_t)(b >> 32) == tag;
}
// bounded wait for one double of this step (~10 s at 1.9 GHz: ranks may be seconds apart at start-up)
__device__ __forceinline__ bool ll_wait(const unsigned long long* p, uint32_t tag, double& v) {
    if (ll_load(p, tag, v)) return true;
    const long long t0 = clock64();
    while (!ll_load(p, tag, v)) {
        if (clock64() - t0 > 20000000000ll) return false;
        __nanosleep(32);
    }
    return true;
}

// The whole exchange step for one CTA of >= 1 + 2d threads (every thread of the CTA must call it: it contains CTA barriers).
//     buffer of a rank:  ll[2 (parity of the step)][world (source rank)][1 + 2d][2] words
//   1. thread c < 1 + 2d stores local[c], tagged with the step, into slot [parity][rank][c] of EVERY rank's buffer (its own included);
//   2. thread c < d waits until the count and its own column's mean and M2 of every source rank carry this step's tag in the local
//      buffer (bounded; a timeout is agreed on per source rank by the whole CTA, so every column merges the same set of ranks);
//   3. Chan merge in rank order into the running statistics.
// Two parities suffice: a rank can only reach step s + 2 after every rank has published step s + 1, which each rank does after it
// finished reading step s.
__device__ __forceinline__ void xchg_merge_body(void* const* __restrict__ peers, int rank, int world, int d, unsigned long long* seq, int* failed,
                                                double* stats, const double* local) {
    const int len = 1 + 2 * d, c = threadIdx.x;
    __shared__ unsigned long long s_seq;
    __shared__ int s_arrived[64];
    if (c == 0) s_seq = ++(*seq);
    if (c < 64) s_arrived[c] = 1;
    __syncthreads();
    const unsigned long long s = s_seq;
    const int par = (int)(s & 1ull);
    const uint32_t tag = (uint32_t)(s & 0x7fffffffull) | 0x80000000u;
    if (c < len) {
        const double v = local[c];
        for (int q0 = 0; q0 < world; ++q0) {
            const int q = (rank + 1 + q0) % world;                      // the peers first, the local copy last
            ll_store(reinterpret_cast<unsigned long long*>(peers[q]) + (((size_t)par * world + rank) * len + c) * 2, v, tag);
        }
    }
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers[rank]) + (size_t)par * world * len * 2;
    if (c < d) {
        for (int q = 0; q < world; ++q) {
            const unsigned long long* m = mine + (size_t)q * len * 2;
            double t;
            const bool ok = ll_wait(m, tag, t) && ll_wait(m + (1 + c) * 2, tag, t) && ll_wait(m + (1 + d + c) * 2, tag, t);
            if (!ok) { s_arrived[q] = 0; *failed = 1; }
        }
    }
    __syncthreads();
    // Chan merge in rank order (RunningMeanStd.update_from_moments, k batches) -- same arithmetic as vecnorm_merge_kernel
    double count = 0.0, mean = 0.0, var = 0.0;
    if (c < d) {
        count = stats[0]; mean = stats[1 + c]; var = stats[1 + d + c];
        for (int q = 0; q < world; ++q) {
            if (!s_arrived[q]) continue;                                // timed out (for some column): skip what never arrived
            const unsigned long long* m = mine + (size_t)q * len * 2;
            double bn, mq, m2q;
            ll_load(m, tag, bn);
            ll_load(m + (1 + c) * 2, tag, mq);
            ll_load(m + (1 + d + c) * 2, tag, m2q);
            if (bn <= 0.0) continue;
            const double delta = mq - mean;
            const double tot = count + bn;
            mean = mean + delta * bn / tot;
            const double M2 = var * count + m2q + delta * delta * count * bn / tot;
            var = M2 / tot;
            count = tot;
        }
    }
    __syncthreads();                                                     // every column has read stats[0] before it is rewritten
    if (c < d) {
        stats[1 + c] = mean;
        stats[1 + d + c] = var;
        if (c == 0) stats[0] = count;
    }
}

}  // namespace qs
