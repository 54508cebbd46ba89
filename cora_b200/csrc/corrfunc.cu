// xi(r) -> C_l(chi, chi') front end (sm_100a): the Legendre transform of cora/signal/corrfunc.py:290-400
// (`corr_to_clarray`).  The reference evaluates the correlation function at the Gauss-Legendre nodes in
// mu = cos(theta), integrates over the radial bins, and contracts with the weighted Legendre polynomials:
//
//     C_l(x, x') = sum_i  [ P_l(mu_i) w_i 4 pi / sum(w) ]  xi_bar(mu_i; x, x')          (corrfunc.py:384-396)
//
// a (L x M) . (M x nx^2) FP64 matrix product.  Here: legendre_table_kernel builds the weighted P_l(mu_i) table
// by the three-term recurrence (scipy.special.lpn, corrfunc.py:281-287), dgemm_kernel is a DMMA.8x8x4
// tensor-core GEMM (128 x 64 tiles, 3-stage cp.async) that accumulates chunk after chunk of mu, so the
// (M x nx^2) integrand never has to exist in memory at once.
#include "common.cuh"
#include "cora_b200.h"

namespace cb {

// one thread per node i: out[l][i] = P_l(mu_i) * scale_i,  l = 0 .. lmax
__global__ void legendre_table_kernel(const double* __restrict__ mu, const double* __restrict__ scale, int n, int lmax,
                                      double* __restrict__ out, long long ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = mu[i], s = scale ? scale[i] : 1.0;
    double p0 = 1.0, p1 = x;
    out[i] = p0 * s;
    if (lmax >= 1) out[ld + i] = p1 * s;
    for (int l = 1; l < lmax; l++) {
        // (l + 1) P_{l+1} = (2l + 1) x P_l - l P_{l-1}
        const double p2 = ((2.0 * l + 1.0) * x * p1 - (double)l * p0) / (double)(l + 1);
        p0 = p1; p1 = p2;
        out[(long long)(l + 1) * ld + i] = p2 * s;
    }
}

// Fused integrand for tabulated correlation functions (cora_b200.corrfunc.TabulatedCorrelation): one thread per
// output (node i, bin a, bin b): cosine rule -> piecewise-linear xi(r) (numpy.interp semantics: clamped at both
// ends) -> the two-sided radial quadrature, all in registers.  Replaces, for such functions, the host evaluation
// `corr(cosine_rule(mu, xa, xa))` and the two matmuls of corrfunc.py:369-379.
//   out[i][a][b] = sum_{p,q} w_p w_q xi( sqrt((x_ap - x_bq)^2 + 2 x_ap x_bq (1 - mu_i)) )
// logx: the table abscissa is ln r (interpolation linear in ln r; r = 0 clamps to the first knot).
__global__ void __launch_bounds__(256) corr_bins_kernel(const double* __restrict__ mu, int nmu, const double* __restrict__ xa,
                                                        const double* __restrict__ xw, int nx, int xint,
                                                        const double* __restrict__ tr, const double* __restrict__ tv, int nt,
                                                        int logx, double* __restrict__ out) {
    extern __shared__ double cb_smem[];   // [2 nt] knots, values (when they fit: nt <= 4096)
    const bool tab_smem = nt <= 4096;
    double* sr = cb_smem;
    double* sv = sr + (tab_smem ? nt : 0);
    if (tab_smem)
        for (int k = threadIdx.x; k < nt; k += blockDim.x) { sr[k] = tr[k]; sv[k] = tv[k]; }
    __syncthreads();
    const double* R = tab_smem ? sr : tr;
    const double* V = tab_smem ? sv : tv;
    const long long n2 = (long long)nx * nx;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (e >= n2) return;
    const int a = (int)(e / nx), b = (int)(e - (long long)a * nx);
    const double om = 1.0 - mu[i];
    double acc = 0.0;
    for (int p = 0; p < xint; p++) {
        const double xp = xa[a * xint + p];
        double accq = 0.0;
        for (int q = 0; q < xint; q++) {
            const double xq = xa[b * xint + q];
            const double d = xp - xq;
            double r = sqrt(d * d + 2.0 * xp * xq * om);
            if (logx) r = log(r);          // log(0) = -inf clamps to the first knot below
            double v;
            if (!(r > R[0])) v = V[0];
            else if (r >= R[nt - 1]) v = V[nt - 1];
            else {
                int lo = 0, hi = nt - 1;   // R[lo] < r < R[hi]... invariant R[lo] <= r < R[hi]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (R[mid] <= r) lo = mid; else hi = mid;
                }
                const double slope = (V[lo + 1] - V[lo]) / (R[lo + 1] - R[lo]);
                v = slope * (r - R[lo]) + V[lo];
            }
            accq += xw[q] * v;
        }
        acc += xw[p] * accq;
    }
    out[(long long)i * n2 + e] = acc;
}

constexpr int GM_TM = 128, GM_TN = 64, GM_KC = 16, GM_ALD = 20, GM_BLD = 68, GM_STAGES = 3;
constexpr int GM_STAGE_DOUBLES = GM_TM * GM_ALD + GM_KC * GM_BLD;

struct GemmParams {
    const double *A, *B;
    double* C;
    long long lda, ldb, ldc;
    int m, n, k, accumulate;
};

// C[m x n] (+)= A[m x k] B[k x n], row-major.  FAST: every 16-byte piece of the A and B rows is aligned
// (even lda / ldb / k / n, 16-byte aligned bases) -> cp.async pipeline; otherwise plain loads.
template <bool FAST>
__global__ void __launch_bounds__(256, 2) dgemm_kernel(GemmParams P) {
    extern __shared__ __align__(16) double gm_smem[];
    const int r0 = blockIdx.y * GM_TM, c0 = blockIdx.x * GM_TN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int nit = (P.k + GM_KC - 1) / GM_KC;
    auto stage_load = [&](int it, int slot) {
        double* As = gm_smem + slot * GM_STAGE_DOUBLES;
        double* Bs = As + GM_TM * GM_ALD;
        const int k0 = it * GM_KC;
        if (FAST) {
#pragma unroll
            for (int q = 0; q < 4; q++) {          // A tile 128 x 16: 1024 pieces of 2 doubles
                const int e = tid + q * 256;
                const int rr = e >> 3, kk = (e & 7) * 2;
                const int r = r0 + rr, k = k0 + kk;
                const bool ok = (r < P.m) && (k < P.k) && (it < nit);
                cp_async16(As + rr * GM_ALD + kk, P.A + (ok ? ((long long)r * P.lda + k) : 0), ok);
            }
#pragma unroll
            for (int q = 0; q < 2; q++) {          // B tile 16 x 64: 512 pieces of 2 doubles
                const int e = tid + q * 256;
                const int kk = e >> 5, cc = (e & 31) * 2;
                const int k = k0 + kk, col = c0 + cc;
                const bool ok = (k < P.k) && (col < P.n) && (it < nit);
                cp_async16(Bs + kk * GM_BLD + cc, P.B + (ok ? ((long long)k * P.ldb + col) : 0), ok);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int e = tid + q * 256;
                const int rr = e >> 4, kk = e & 15;
                const int r = r0 + rr, k = k0 + kk;
                As[rr * GM_ALD + kk] = (r < P.m && k < P.k && it < nit) ? P.A[(long long)r * P.lda + k] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int e = tid + q * 256;
                const int kk = e >> 6, cc = e & 63;
                const int k = k0 + kk, col = c0 + cc;
                Bs[kk * GM_BLD + cc] = (k < P.k && col < P.n && it < nit) ? P.B[(long long)k * P.ldb + col] : 0.0;
            }
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; s++) stage_load(s, s);
    for (int it = 0; it < nit; it++) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();
        stage_load(it + GM_STAGES - 1, (it + GM_STAGES - 1) % GM_STAGES);
        const double* As = gm_smem + (it % GM_STAGES) * GM_STAGE_DOUBLES;
        const double* Bs = As + GM_TM * GM_ALD;
#pragma unroll
        for (int k4 = 0; k4 < GM_KC / 4; k4++) {
            double af[4], bf[4];
#pragma unroll
            for (int mb = 0; mb < 4; mb++) af[mb] = As[(wm * 32 + 8 * mb + g) * GM_ALD + k4 * 4 + t];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) bf[nb] = Bs[(k4 * 4 + t) * GM_BLD + wn * 32 + 8 * nb + g];
#pragma unroll
            for (int mb = 0; mb < 4; mb++)
#pragma unroll
                for (int nb = 0; nb < 4; nb++) dmma884(acc[mb][nb][0], acc[mb][nb][1], af[mb], bf[nb]);
        }
    }
    cp_async_wait<0>();
    // epilogue: this lane holds C[row g][cols 2t, 2t+1] of every 8 x 8 block
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
        const int r = r0 + wm * 32 + 8 * mb + g;
        if (r >= P.m) continue;
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            const int c = c0 + wn * 32 + 8 * nb + 2 * t;
            double* o = P.C + (long long)r * P.ldc + c;
            if (c < P.n) o[0] = (P.accumulate ? o[0] : 0.0) + acc[mb][nb][0];
            if (c + 1 < P.n) o[1] = (P.accumulate ? o[1] : 0.0) + acc[mb][nb][1];
        }
    }
}

}  // namespace cb

using namespace cb;

extern "C" int cora_b200_legendre_table(const double* mu, const double* scale, int n, int lmax, double* out, long long ld,
                                        void* stream) {
    CB_REQUIRE(mu && out && n >= 1 && lmax >= 0 && ld >= n, 1, "legendre_table: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    KTimer kt(K_TABLE, st);
    legendre_table_kernel<<<ceil_div(n, 128), 128, 0, st>>>(mu, scale, n, lmax, out, ld);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_corr_bins(const double* mu, int nmu, const double* xa, const double* xw, int nx, int xint,
                                   const double* tab_r, const double* tab_v, int nt, int logx, double* out, void* stream) {
    CB_REQUIRE(mu && xa && xw && tab_r && tab_v && out && nmu >= 1 && nx >= 1 && xint >= 1 && nt >= 2, 1, "corr_bins: bad arguments");
    CB_REQUIRE(nmu <= 65535, 3, "corr_bins: more than 65535 nodes per call");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (nt <= 4096) ? sizeof(double) * 2 * (size_t)nt : 0;
    CB_CUDA(cudaFuncSetAttribute(corr_bins_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    KTimer kt(K_CL_FILL, st);
    corr_bins_kernel<<<dim3(ceil_div((long long)nx * nx, 256), nmu), 256, smem, st>>>(mu, nmu, xa, xw, nx, xint, tab_r, tab_v, nt, logx, out);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_dgemm(const double* A, const double* B, double* C, int m, int n, int k, long long lda, long long ldb,
                               long long ldc, int accumulate, void* stream) {
    CB_REQUIRE(A && B && C && m >= 1 && n >= 1 && k >= 1 && lda >= k && ldb >= n && ldc >= n, 1, "dgemm: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    GemmParams P;
    P.A = A; P.B = B; P.C = C; P.lda = lda; P.ldb = ldb; P.ldc = ldc; P.m = m; P.n = n; P.k = k; P.accumulate = accumulate;
    const bool fast = (lda % 2 == 0) && (ldb % 2 == 0) && (k % 2 == 0) && (n % 2 == 0) && ((uintptr_t)A % 16 == 0) &&
                      ((uintptr_t)B % 16 == 0);
    dim3 grid(ceil_div(n, GM_TN), ceil_div(m, GM_TM));
    CB_REQUIRE(grid.y <= 65535, 3, "dgemm: too many row tiles (%u)", grid.y);
    const size_t smem = sizeof(double) * GM_STAGES * GM_STAGE_DOUBLES;
    KTimer kt(K_APPLY, st);
    if (fast) {
        CB_CUDA(cudaFuncSetAttribute(dgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dgemm_kernel<true><<<grid, 256, smem, st>>>(P);
    } else {
        CB_CUDA(cudaFuncSetAttribute(dgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dgemm_kernel<false><<<grid, 256, smem, st>>>(P);
    }
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}
