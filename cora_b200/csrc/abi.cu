// Library-level C ABI: version, error string, launch counter, FP64 peak probe.
#include "common.cuh"
#include "cora_b200.h"

#include <cstdarg>
#include <utility>
#include <vector>

namespace cb {

static thread_local char g_err[1024] = "";
long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

// ---- per-kernel timing -----------------------------------------------------------------------
bool g_timing_on = false;
struct TimedSpan { int id; cudaEvent_t e0, e1; };
static std::vector<TimedSpan> g_spans;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_free_events;
static int g_open = -1;

void timing_begin(int id, cudaStream_t st) {
    TimedSpan s;
    s.id = id;
    if (!g_free_events.empty()) {
        s.e0 = g_free_events.back().first; s.e1 = g_free_events.back().second;
        g_free_events.pop_back();
    } else {
        cudaEventCreate(&s.e0);
        cudaEventCreate(&s.e1);
    }
    cudaEventRecord(s.e0, st);
    g_spans.push_back(s);
    g_open = (int)g_spans.size() - 1;
}
void timing_end(cudaStream_t st) {
    if (g_open >= 0) cudaEventRecord(g_spans[g_open].e1, st);
    g_open = -1;
}

static const char* const g_kernel_names[K_COUNT] = {
    "cl_fill", "root_prepare", "cholesky", "eigh_jacobi", "draw", "apply", "alm_layout", "sht_legendre", "sht_phase", "ps_table"};

// 8 independent DMMA chains per warp; 256 FMA per DMMA per warp.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c0[i] = threadIdx.x; c1[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace cb

using namespace cb;

extern "C" int cora_b200_version(void) { return 100; }
extern "C" const char* cora_b200_last_error(void) { return get_error(); }
extern "C" long long cora_b200_launch_count(void) { return g_launches; }

extern "C" int cora_b200_timing_enable(int on) {
    // drop whatever was recorded and (re)arm
    for (auto& s : g_spans) g_free_events.push_back({s.e0, s.e1});
    g_spans.clear();
    g_open = -1;
    g_timing_on = on != 0;
    return 0;
}
extern "C" int cora_b200_timing_kinds(void) { return K_COUNT; }
extern "C" const char* cora_b200_timing_name(int id) { return (id >= 0 && id < K_COUNT) ? g_kernel_names[id] : ""; }
extern "C" int cora_b200_timing_read(double* ms_out, long long* launches_out, int n) {
    CB_REQUIRE(ms_out && launches_out && n >= K_COUNT, 1, "timing_read: need arrays of >= %d entries", (int)K_COUNT);
    for (int i = 0; i < n; i++) { ms_out[i] = 0.0; launches_out[i] = 0; }
    for (auto& s : g_spans) {
        CB_CUDA(cudaEventSynchronize(s.e1));
        float ms = 0.f;
        CB_CUDA(cudaEventElapsedTime(&ms, s.e0, s.e1));
        ms_out[s.id] += ms;
        launches_out[s.id] += 1;
        g_free_events.push_back({s.e0, s.e1});
    }
    g_spans.clear();
    return 0;
}

extern "C" int cora_b200_timing_trace(int* id_out, double* start_ms_out, double* dur_ms_out, int cap) {
    // one entry per recorded span, in launch order: start relative to the first span's start.  Does not clear.
    CB_REQUIRE(id_out && start_ms_out && dur_ms_out && cap >= 0, 1, "timing_trace: null output");
    int n = 0;
    for (auto& s : g_spans) {
        if (n >= cap) break;
        CB_CUDA(cudaEventSynchronize(s.e1));
        float t0 = 0.f, dt = 0.f;
        CB_CUDA(cudaEventElapsedTime(&t0, g_spans[0].e0, s.e0));
        CB_CUDA(cudaEventElapsedTime(&dt, s.e0, s.e1));
        id_out[n] = s.id; start_ms_out[n] = t0; dur_ms_out[n] = dt;
        n++;
    }
    return -n;   // (negative count: non-negative returns are error codes everywhere else in this ABI)
}

extern "C" int cora_b200_fp64_peak(double ms_budget, double* tflops_out, void* stream) {
    CB_REQUIRE(tflops_out, 1, "fp64_peak: null output");
    cudaStream_t st = (cudaStream_t)stream;
    int dev, nsm;
    CB_CUDA(cudaGetDevice(&dev));
    CB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const int grid = nsm * 4, iters = 4096;
    double* out;
    CB_CUDA(cudaMalloc(&out, sizeof(double) * grid * 256));
    cudaEvent_t e0, e1;
    CB_CUDA(cudaEventCreate(&e0));
    CB_CUDA(cudaEventCreate(&e1));
    fp64_peak_kernel<<<grid, 256, 0, st>>>(out, iters, 1.0000001, 1e-9);   // warm-up
    count_launch();
    double flop_per = 2.0 * 256 * 8 * (double)iters * 8.0 * grid;
    int reps = 1;
    double best = 0;
    // one launch is ~2.1 ms on B200; repeat until the budget is used, keep the best
    double spent = 0;
    while (spent < ms_budget && reps < 1000) {
        CB_CUDA(cudaEventRecord(e0, st));
        fp64_peak_kernel<<<grid, 256, 0, st>>>(out, iters, 1.0000001, 1e-9);
        count_launch();
        CB_CUDA(cudaEventRecord(e1, st));
        CB_CUDA(cudaEventSynchronize(e1));
        float ms;
        CB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        spent += ms;
        double tf = flop_per / ms * 1e-9;
        if (tf > best) best = tf;
        reps++;
    }
    CB_LAUNCH_CHECK();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops_out = best;
    return 0;
}
