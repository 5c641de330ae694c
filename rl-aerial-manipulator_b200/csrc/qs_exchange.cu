// qs_exchange.cu -- the one exchange step of the sharded env path (VecNormalize moments), fused with its consumer:
// all-gather over NVLink peer memory + Chan merge in ONE kernel (include/quadsim.h: qs_xchg_*).
//
// Per rank one cudaMalloc'ed buffer, shared with the peers through CUDA IPC:
//     slots[2][world][len]   doubles   len = 1 + 2d   (parity of the step, source rank)
//     flags[2][world]        uint64    sequence number of the step whose triplet the slot holds
// Step s (1, 2, ...; the counter lives in device memory so a captured graph replays correctly), parity p = s & 1:
//   1. thread c < len stores local[c] into slots[p][rank][c] of EVERY rank's buffer (its own included), then fences system-wide;
//   2. one thread publishes flags[p][rank] = s in every buffer (after the CTA barrier, so all data stores are fenced);
//   3. thread q acquire-spins until its own buffer's flags[p][q] >= s (bounded: ~10 s, then a sticky error), CTA barrier;
//   4. Chan merge of slots[p][0..world) in rank order into the running statistics.
// Two parities suffice: a rank can only reach step s+2 after every rank has published step s+1, which each rank does after it
// finished reading step s.
#include "../../include/quadsim.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <new>

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

thread_local char g_xchg_error[256] = "";

static int xfail(const char* what, cudaError_t err) {
    snprintf(g_xchg_error, sizeof(g_xchg_error), "%s: %s", what, cudaGetErrorString(err));
    return QS_ECUDA;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(128) xchg_merge_kernel(void* const* __restrict__ peers, int rank, int world, int d,
                                                         unsigned long long* seq, int* failed, double* __restrict__ stats,
                                                         const double* __restrict__ local) {
    const int len = 1 + 2 * d, c = threadIdx.x;
    __shared__ unsigned long long s_seq;
    if (c == 0) s_seq = ++(*seq);
    __syncthreads();
    const unsigned long long s = s_seq;
    const int par = (int)(s & 1ull);
    const size_t slot_doubles = (size_t)2 * world * len;
    // 1. publish the data
    if (c < len) {
        const double v = local[c];
        for (int q = 0; q < world; ++q) {
            double* slots = reinterpret_cast<double*>(peers[q]);
            slots[((size_t)par * world + rank) * len + c] = v;
        }
        __threadfence_system();
    }
    __syncthreads();
    // 2. publish the flags
    if (c < world) {
        unsigned long long* flags = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(peers[c]) + slot_doubles);
        st_release_sys(flags + (size_t)par * world + rank, s);
    }
    // 3. wait for everybody's triplet of this step in the local buffer
    const double* my_slots = reinterpret_cast<const double*>(peers[rank]);
    const unsigned long long* my_flags = reinterpret_cast<const unsigned long long*>(my_slots + slot_doubles) + (size_t)par * world;
    // one decision per source rank (thread q waits for rank q), shared by all columns: after a timeout every column merges the
    // same set of ranks, so the statistics stay self-consistent and the sticky flag says they are incomplete
    __shared__ int s_arrived[64];
    if (c < world) {
        bool ok = true;
        const long long t0 = clock64();
        while (ld_acquire_sys(my_flags + c) < s) {
            if (clock64() - t0 > 20000000000ll) { ok = false; break; }  // ~10 s at 1.9 GHz: ranks may be seconds apart at start-up
            __nanosleep(64);
        }
        s_arrived[c] = ok ? 1 : 0;
        if (!ok) *failed = 1;
    }
    __syncthreads();
    // 4. Chan merge in rank order (RunningMeanStd.update_from_moments, k batches) -- same arithmetic as vecnorm_merge_kernel
    double count = 0.0, mean = 0.0, var = 0.0;
    if (c < d) {
        count = stats[0]; mean = stats[1 + c]; var = stats[1 + d + c];
        for (int q = 0; q < world; ++q) {
            const volatile double* m = my_slots + ((size_t)par * world + q) * len;   // written by a peer: never from a stale L1 line
            if (!s_arrived[q]) continue;                                // timed out: skip what never arrived
            (void)ld_acquire_sys(my_flags + q);                         // acquire in THIS thread before reading the peer's data
            const double bn = m[0];
            if (bn <= 0.0) continue;
            const double delta = m[1 + c] - mean;
            const double tot = count + bn;
            mean = mean + delta * bn / tot;
            const double M2 = var * count + m[1 + d + c] + delta * delta * count * bn / tot;
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

using namespace qs;

extern "C" {

const char* qs_xchg_last_error(void) { return g_xchg_error; }

int qs_xchg_create(int device, int rank, int world, int d, qs_xchg** out, unsigned char* ipc_handle_out) {
    if (!out || !ipc_handle_out || world < 1 || world > 64 || rank < 0 || rank >= world || d < 1 || d > 32) {
        snprintf(g_xchg_error, sizeof(g_xchg_error), "qs_xchg_create: bad argument (1 <= world <= 64, 1 <= d <= 32)");
        return QS_EINVAL;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess) return xfail("qs_xchg_create", err);
    qs_xchg* x = new (std::nothrow) qs_xchg();
    if (!x) { snprintf(g_xchg_error, sizeof(g_xchg_error), "qs_xchg_create: out of host memory"); return QS_EINVAL; }
    x->device = device; x->rank = rank; x->world = world; x->d = d; x->len = 1 + 2 * d; x->connected = false;
    x->bytes = (size_t)2 * world * x->len * sizeof(double) + (size_t)2 * world * sizeof(unsigned long long);
    x->peer_host = new void*[world]();
    x->local = nullptr; x->peer_dev = nullptr; x->seq = nullptr; x->failed = nullptr;
    if ((err = cudaMalloc(&x->local, x->bytes)) != cudaSuccess || (err = cudaMemset(x->local, 0, x->bytes)) != cudaSuccess ||
        (err = cudaMalloc(&x->peer_dev, sizeof(void*) * world)) != cudaSuccess ||
        (err = cudaMalloc(&x->seq, sizeof(unsigned long long))) != cudaSuccess || (err = cudaMemset(x->seq, 0, sizeof(unsigned long long))) != cudaSuccess ||
        (err = cudaMalloc(&x->failed, sizeof(int))) != cudaSuccess || (err = cudaMemset(x->failed, 0, sizeof(int))) != cudaSuccess) {
        qs_xchg_destroy(x);
        return xfail("qs_xchg_create", err);
    }
    cudaIpcMemHandle_t h;
    if ((err = cudaIpcGetMemHandle(&h, x->local)) != cudaSuccess) {
        qs_xchg_destroy(x);
        return xfail("qs_xchg_create: cudaIpcGetMemHandle", err);
    }
    memcpy(ipc_handle_out, &h, 64);
    if ((err = cudaDeviceSynchronize()) != cudaSuccess) { qs_xchg_destroy(x); return xfail("qs_xchg_create", err); }
    *out = x;
    return QS_OK;
}

int qs_xchg_connect(qs_xchg* x, const unsigned char* all_handles) {
    if (!x || !all_handles) { snprintf(g_xchg_error, sizeof(g_xchg_error), "qs_xchg_connect: null argument"); return QS_EINVAL; }
    cudaError_t err = cudaSetDevice(x->device);
    if (err != cudaSuccess) return xfail("qs_xchg_connect", err);
    for (int q = 0; q < x->world; ++q) {
        if (q == x->rank) { x->peer_host[q] = x->local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + (size_t)q * 64, 64);
        void* ptr = nullptr;
        if ((err = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) {
            snprintf(g_xchg_error, sizeof(g_xchg_error), "qs_xchg_connect: cudaIpcOpenMemHandle(rank %d): %s", q, cudaGetErrorString(err));
            return QS_ECUDA;
        }
        x->peer_host[q] = ptr;
    }
    if ((err = cudaMemcpy(x->peer_dev, x->peer_host, sizeof(void*) * x->world, cudaMemcpyHostToDevice)) != cudaSuccess)
        return xfail("qs_xchg_connect", err);
    x->connected = true;
    return QS_OK;
}

int qs_xchg_merge(qs_xchg* x, double* stats, const double* local_moments, void* stream) {
    if (!x || !stats || !local_moments) { snprintf(g_xchg_error, sizeof(g_xchg_error), "qs_xchg_merge: null argument"); return QS_EINVAL; }
    if (!x->connected) { snprintf(g_xchg_error, sizeof(g_xchg_error), "qs_xchg_merge: call qs_xchg_connect first"); return QS_EINVAL; }
    xchg_merge_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(x->peer_dev, x->rank, x->world, x->d, x->seq, x->failed, stats, local_moments);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return xfail("qs_xchg_merge", err);
    return QS_OK;
}

int qs_xchg_failed(qs_xchg* x) {
    if (!x) return 1;
    int f = 1;
    cudaSetDevice(x->device);
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    if (cudaMemcpy(&f, x->failed, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    return f;
}

int qs_xchg_destroy(qs_xchg* x) {
    if (!x) return QS_OK;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    if (x->peer_host) {
        for (int q = 0; q < x->world; ++q)
            if (q != x->rank && x->peer_host[q]) cudaIpcCloseMemHandle(x->peer_host[q]);
        delete[] x->peer_host;
    }
    if (x->local) cudaFree(x->local);
    if (x->peer_dev) cudaFree(x->peer_dev);
    if (x->seq) cudaFree(x->seq);
    if (x->failed) cudaFree(x->failed);
    delete x;
    return QS_OK;
}

}  // extern "C"
