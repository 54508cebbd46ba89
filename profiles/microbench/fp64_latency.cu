// FP64 dependent-issue latency on B200: one warp per CTA, chains of dependent DFMA / the recurrence step.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_chain(double* out, long long* cyc, double a, double b, int n) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x = fma(x, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// the Legendre step: pn = fma(c1 * x, p, -c2 * pp)
__global__ void recur_chain(double* out, long long* cyc, double c1, double c2, int n) {
    double x = 0.3 + 1e-3 * threadIdx.x, p = 1.0, pp = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            double pn = fma(c1 * x, p, -c2 * pp);
            pp = p; p = pn;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = p + pp;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    long long h;
    for (int rep = 0; rep < 2; rep++) {
        dfma_chain<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 1000);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dependent DFMA: %.2f cycles/op\n", (double)h / 16000.0);
        recur_chain<<<1, 32>>>(out, cyc, 1.9999, 0.9999, 1000);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("recurrence step (2 DMUL + DFMA): %.2f cycles/step\n", (double)h / 16000.0);
    }
    return 0;
}
