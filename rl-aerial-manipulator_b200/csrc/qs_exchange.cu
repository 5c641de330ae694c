// qs_exchange.cu -- the one exchange step of the sharded env path (VecNormalize moments), fused with its consumer:
// all-gather over NVLink peer memory + Chan merge in ONE kernel (include/quadsim.h: qs_xchg_*).
//
// Per rank one cudaMalloc'ed buffer, shared with the peers through CUDA IPC; protocol and layout: qs_exchange.cuh (tagged 8-byte
// words, no fences or flags: one NVLink write latency per exchange).
#include "../../include/quadsim.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <new>

#include "qs_exchange.cuh"

namespace qs {

thread_local char g_xchg_error[256] = "";

static int xfail(const char* what, cudaError_t err) {
    snprintf(g_xchg_error, sizeof(g_xchg_error), "%s: %s", what, cudaGetErrorString(err));
    return QS_ECUDA;
}

__global__ void __launch_bounds__(128) xchg_merge_kernel(void* const* __restrict__ peers, int rank, int world, int d,
                                                         unsigned long long* seq, int* failed, double* __restrict__ stats,
                                                         const double* __restrict__ local) {
    xchg_merge_body(peers, rank, world, d, seq, failed, stats, local);
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
    x->bytes = (size_t)2 * world * x->len * 2 * sizeof(unsigned long long);
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
