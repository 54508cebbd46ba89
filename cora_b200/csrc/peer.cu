// Peer (NVLink / NVSwitch) memory for the one-process-per-GPU layout: buffers allocated here are
// exported as CUDA IPC handles, every other rank of the box maps them, and the producing kernels
// (C_l fill, apply) store straight into the consumer GPU's buffer -- the exchange of
// cora/core/skysim.py:128 (alm_array.redistribute) happens inside the kernel epilogue instead of
// in a separate collective.  A flag barrier over the same peer memory orders the steps.
#include "common.cuh"
#include "cora_b200.h"

namespace cb {

// One CTA; thread i < size signals rank i and waits for rank i's signal.
//   flags[r]  : rank r's flag array (u64[size]) as mapped in THIS process (flags[rank] is local)
//   epoch     : strictly increasing barrier number (1, 2, ...)
// All stores of earlier kernels on this stream are complete at kernel entry (stream order); the
// system fence + release store publishes them to the peer before it observes the flag.
// A peer that does not arrive within timeout_s: *status = 1 + its rank (sticky), and -- unless fatal == 0 --
// the kernel traps: the kernels queued behind the barrier would otherwise read buffers the slow peer
// has not finished writing and hand back wrong maps without any error.  After the trap every later CUDA
// call of the process fails loudly.
__global__ void peer_barrier_kernel(unsigned long long* const* __restrict__ flags, int rank, int size,
                                    unsigned long long epoch, double timeout_s, int* __restrict__ status, int fatal) {
    const int i = threadIdx.x;
    if (i >= size) return;
    __threadfence_system();
    unsigned long long* remote = flags[i] + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(remote), "l"(epoch) : "memory");
    const unsigned long long* mine = flags[rank] + i;
    unsigned long long t0, t1, v;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t0));
    const unsigned long long limit = (unsigned long long)(timeout_s * 1e9);
    while (true) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t1));
        if (t1 - t0 > limit) {                                        // a peer never arrived: report, do not hang
            atomicExch(status, 1 + i);
            __threadfence_system();
            if (fatal) __trap();
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace cb

using namespace cb;

extern "C" int cora_b200_peer_alloc(long long bytes, void** ptr_out, unsigned char* handle64_out) {
    CB_REQUIRE(bytes > 0 && ptr_out && handle64_out, 1, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    CB_CUDA(cudaMalloc(&p, (size_t)bytes));
    CB_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        return 100 + (int)e;
    }
    memcpy(handle64_out, &h, 64);
    *ptr_out = p;
    return 0;
}

extern "C" int cora_b200_peer_free(void* ptr) {
    if (ptr) CB_CUDA(cudaFree(ptr));
    return 0;
}

extern "C" int cora_b200_peer_open(const unsigned char* handle64, void** ptr_out) {
    CB_REQUIRE(handle64 && ptr_out, 1, "peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = p;
    return 0;
}

extern "C" int cora_b200_peer_close(void* ptr) {
    if (ptr) CB_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int cora_b200_peer_barrier(const void* flags_ptrs, int rank, int size, unsigned long long epoch,
                                      double timeout_s, int* status, int fatal, void* stream) {
    CB_REQUIRE(flags_ptrs && status && size >= 1 && size <= 1024 && rank >= 0 && rank < size && epoch > 0, 1,
               "peer_barrier: bad arguments");
    peer_barrier_kernel<<<1, ((size + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(
        (unsigned long long* const*)flags_ptrs, rank, size, epoch, timeout_s > 0 ? timeout_s : 60.0, status, fatal);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}
