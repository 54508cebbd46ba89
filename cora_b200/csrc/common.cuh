// Shared helpers for the cora_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace cb {

// ---- error plumbing: every C-ABI entry returns int (0 = OK) and never throws -------------
void set_error(const char* fmt, ...);
const char* get_error();

#define CB_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t _e = (call);                                                        \
        if (_e != cudaSuccess) {                                                        \
            cb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                 \
                          cudaGetErrorString(_e));                                      \
            return 100 + (int)_e;                                                       \
        }                                                                               \
    } while (0)

#define CB_REQUIRE(cond, code, ...)                                                     \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            cb::set_error(__VA_ARGS__);                                                 \
            return (code);                                                              \
        }                                                                               \
    } while (0)

#define CB_LAUNCH_CHECK()                                                               \
    do {                                                                                \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) {                                                        \
            cb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,             \
                          cudaGetErrorString(_e));                                      \
            return 100 + (int)_e;                                                       \
        }                                                                               \
    } while (0)

// launch counter (bench.py reports it as gpu_launches)
extern long long g_launches;
inline void count_launch(int n = 1) { g_launches += n; }

// ---- per-kernel device timing (cora_b200_timing_*): CUDA events recorded on the launching
// stream around each kernel while enabled; bench.py derives roofline.achieved from these.
enum KernelId {
    K_CL_FILL = 0, K_ROOT_PREP, K_CHOLESKY, K_EIGH, K_DRAW, K_APPLY, K_LAYOUT, K_LEGENDRE, K_PHASE, K_TABLE, K_COUNT
};
extern bool g_timing_on;
void timing_begin(int id, cudaStream_t st);
void timing_end(cudaStream_t st);
struct KTimer {
    cudaStream_t st;
    bool on;
    KTimer(int id, cudaStream_t s) : st(s), on(g_timing_on) { if (on) timing_begin(id, st); }
    ~KTimer() { if (on) timing_end(st); }
};

// ---- device helpers -----------------------------------------------------------------------
// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// Fragment ownership (lane = 4*g + t): A[g][t], B[t][g], C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace cb
