// FP64 peak micro-benchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) throughput.
// MEASURED_PEAKS.json has no FP64 entry; the SHT / apply kernels are FP64-compute-bound,
// so their roofline denominator is measured here.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs.  256 FMA per warp instruction.
template <int NT>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double a, double b) {
    double c0[NT], c1[NT];
#pragma unroll
    for (int i = 0; i < NT; i++) { c0[i] = threadIdx.x; c1[i] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NT; i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k8: A 4 regs, B 2 regs, C 4 regs.  1024 FMA per warp instruction.
template <int NT>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters, double a, double b) {
    double c[NT][4];
#pragma unroll
    for (int i = 0; i < NT; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NT; i++) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs. 2048 FMA per warp instruction.
template <int NT>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters, double a, double b) {
    double c[NT][4];
#pragma unroll
    for (int i = 0; i < NT; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NT; i++) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, nsm, p.clockRate);
    double* out; CHECK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
    const int iters = 4096;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = nsm * bps;
        {
            float ms = time_it([&] { dfma_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * iters * 256.0 * grid;
            printf("DFMA  acc16 blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dmma884_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 256 * 8 * iters * 8.0 * grid;
            printf("DMMA m8n8k4  x8 blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dmma1688_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 1024 * 8 * iters * 8.0 * grid;
            printf("DMMA m16n8k8 x8 blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dmma16816_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 2048 * 8 * iters * 8.0 * grid;
            printf("DMMA m16n8k16 x8 blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
    }
    // sustained: 2 seconds of DFMA to see the clock settle
    {
        int grid = nsm * 4;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int n = 0;
        for (; n < 400; n++) dfma_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 16 * iters * 256.0 * grid * n;
        printf("DFMA sustained (%d launches, %.1f ms): %.2f TFLOP/s\n", n, ms, fl / ms * 1e-9);
        cudaEventRecord(e0);
        for (n = 0; n < 400; n++) dmma884_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 2.0 * 256 * 8 * iters * 8.0 * grid * n;
        printf("DMMA m8n8k4 sustained (%d launches, %.1f ms): %.2f TFLOP/s\n", n, ms, fl / ms * 1e-9);
    }
    return 0;
}
