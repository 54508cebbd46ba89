// alm layout conversions between the library's PANEL layout and cora's dense [chan][l][m].
#include "common.cuh"
#include "cora_b200.h"

#include <algorithm>

namespace cb {

// one thread per (l, m<=l... full square) x channel tile; channels contiguous in the panel,
// m contiguous in the dense array -> stage through shared memory for coalescing on both sides.
__global__ void panel_to_dense_kernel(const double2* __restrict__ panel, long long stride, int chan0, int nchan,
                                      int lmax, double2* __restrict__ dense) {
    __shared__ double2 tile[32][33];
    const int L = lmax + 1;
    const int l = blockIdx.z;
    const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int m = m0 + r, c = c0 + threadIdx.x;
        double2 v = make_double2(0.0, 0.0);
        if (m <= l && c < nchan) {
            long long idx = (long long)m * (2 * lmax + 1 - m) / 2 + l;
            v = panel[idx * stride + chan0 + c];
        }
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, m = m0 + threadIdx.x;
        if (c < nchan && m < L) dense[((long long)c * L + l) * L + m] = tile[threadIdx.x][r];
    }
}

__global__ void dense_to_panel_kernel(const double2* __restrict__ dense, int nchan, int lmax,
                                      double2* __restrict__ panel, long long stride, int chan0) {
    __shared__ double2 tile[32][33];
    const int L = lmax + 1;
    const int l = blockIdx.z;
    const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, m = m0 + threadIdx.x;
        double2 v = make_double2(0.0, 0.0);
        if (c < nchan && m <= l) v = dense[((long long)c * L + l) * L + m];
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int m = m0 + r, c = c0 + threadIdx.x;
        if (m <= l && c < nchan) {
            long long idx = (long long)m * (2 * lmax + 1 - m) / 2 + l;
            panel[idx * stride + chan0 + c] = tile[threadIdx.x][r];
        }
    }
}

// received all-to-all slabs -> PANEL.  The rows (l, m = 0..l) of one l are consecutive in the
// receive buffer, `cb` channels each, starting at complex offset l_off[l].
__global__ void slabs_to_panel_kernel(const double2* __restrict__ recv, const long long* __restrict__ l_off, int lmax,
                                      int cb, double2* __restrict__ panel, long long stride, int chan0) {
    const int l = blockIdx.y;
    const long long n = (long long)(l + 1) * cb;
    const double2* src = recv + l_off[l];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(e / cb), c = (int)(e - (long long)m * cb);
        const long long idx = (long long)m * (2 * lmax + 1 - m) / 2 + l;
        panel[idx * stride + chan0 + c] = src[e];
    }
}

// panel[idx(l, m)][chan0 + c] *= scale[l, c]: a per-(l, channel) factor on every a_lm (Gaussian beam of
// healpy.smoothing, one width per channel).  grid (channel blocks, l), threads over (m, c).
__global__ void alm_scale_l_kernel(double2* __restrict__ panel, long long stride, int chan0, int nchan, int lmax,
                                   const double* __restrict__ scale) {
    const int l = blockIdx.y;
    const long long n = (long long)(l + 1) * nchan;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(e / nchan), c = (int)(e - (long long)m * nchan);
        const long long idx = (long long)m * (2 * lmax + 1 - m) / 2 + l;
        const double f = scale[(long long)l * nchan + c];
        double2 v = panel[idx * stride + chan0 + c];
        v.x *= f; v.y *= f;
        panel[idx * stride + chan0 + c] = v;
    }
}

}  // namespace cb

using namespace cb;

extern "C" int cora_b200_alm_scale_l(void* alm_panel, long long panel_stride, int chan0, int nchan, int lmax, const double* scale,
                                     void* stream) {
    CB_REQUIRE(alm_panel && scale && nchan >= 1 && lmax >= 0 && lmax + 1 <= 65535, 1, "alm_scale_l: bad arguments");
    KTimer kt(K_LAYOUT, (cudaStream_t)stream);
    const long long per_l = (long long)(lmax + 1) * nchan;
    dim3 grid((unsigned)std::min<long long>(64, (per_l + 255) / 256), lmax + 1);
    alm_scale_l_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((double2*)alm_panel, panel_stride, chan0, nchan, lmax, scale);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_alm_slabs_to_panel(const void* recv, const long long* l_off, int lmax, int nchan, void* alm_panel,
                                            long long panel_stride, int chan0, void* stream) {
    CB_REQUIRE(recv && l_off && alm_panel && nchan >= 1 && lmax >= 0 && lmax + 1 <= 65535, 1, "alm_slabs_to_panel: bad arguments");
    KTimer kt(K_LAYOUT, (cudaStream_t)stream);
    const long long per_l = (long long)(lmax + 1) * nchan;
    dim3 grid((unsigned)std::min<long long>(64, (per_l + 255) / 256), lmax + 1);
    slabs_to_panel_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const double2*)recv, l_off, lmax, nchan, (double2*)alm_panel,
                                                                  panel_stride, chan0);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_alm_panel_to_dense(const void* alm_panel, long long panel_stride, int chan0, int nchan, int lmax,
                                            void* dense, void* stream) {
    CB_REQUIRE(alm_panel && dense && nchan >= 1 && lmax >= 0, 1, "alm_panel_to_dense: bad arguments");
    CB_REQUIRE(lmax + 1 <= 65535, 1, "alm_panel_to_dense: lmax too large");
    dim3 grid(ceil_div(lmax + 1, 32), ceil_div(nchan, 32), lmax + 1);
    KTimer kt(K_LAYOUT, (cudaStream_t)stream);
    panel_to_dense_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const double2*)alm_panel, panel_stride, chan0, nchan,
                                                                          lmax, (double2*)dense);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_alm_dense_to_panel(const void* dense, int nchan, int lmax, void* alm_panel, long long panel_stride,
                                            int chan0, void* stream) {
    CB_REQUIRE(alm_panel && dense && nchan >= 1 && lmax >= 0, 1, "alm_dense_to_panel: bad arguments");
    CB_REQUIRE(lmax + 1 <= 65535, 1, "alm_dense_to_panel: lmax too large");
    dim3 grid(ceil_div(lmax + 1, 32), ceil_div(nchan, 32), lmax + 1);
    KTimer kt(K_LAYOUT, (cudaStream_t)stream);
    dense_to_panel_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((const double2*)dense, nchan, lmax,
                                                                          (double2*)alm_panel, panel_stride, chan0);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}
