// Batched per-l matrix root with the reference's semantics (sm_100a).
//
//   corrm = C_l + 1e-14 max(diag C_l) I                     (cora/core/skysim.py:116-117)
//   try Cholesky (lower); on a non-positive pivot fall back to a symmetric eigen-
//   decomposition, zero eigenvalues < 1e-16 max, root = evecs sqrt(evals), all columns
//   kept, ascending eigenvalue order                        (cora/util/nputil.py:51-101)
//
// Cholesky: one CTA per matrix, left-looking blocked (32-wide panels) on the matrix resident
// in global/L2.  Fallback: one CTA per failed matrix, one-sided (Hestenes) Jacobi on the rows
// of the symmetric matrix with accumulated rotations; rows end up as lambda_i v_i^T.
#include "common.cuh"
#include "cora_b200.h"

#include <algorithm>

namespace cb {

constexpr int CH_NB = 32;      // panel width
constexpr int CH_THREADS = 256;

// jitter + copy lower triangle (upper zeroed); also per-matrix max(diag)
__global__ void root_prepare_kernel(const double* __restrict__ cl, int nz, double jitter_rel, double* __restrict__ root,
                                    double* __restrict__ dmax_out, const double* __restrict__ dmax_in) {
    __shared__ double red[256];
    const long long base = (long long)blockIdx.x * nz * nz;
    double mx = -1.0e308;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) mx = fmax(mx, cl[base + (long long)i * nz + i]);
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    // dmax_in: the diagonal maximum of a larger (block-diagonal) matrix this one is a block of
    const double dm = dmax_in ? dmax_in[blockIdx.x] : red[0];
    const double cmax = dm * jitter_rel;
    if (threadIdx.x == 0 && blockIdx.y == 0) dmax_out[blockIdx.x] = dm;
    // the copy is split over gridDim.y CTAs per matrix (each recomputes the cheap diagonal maximum)
    const long long n2 = (long long)nz * nz;
    const long long e0 = n2 * blockIdx.y / gridDim.y, e1 = n2 * (blockIdx.y + 1) / gridDim.y;
    for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int r = (int)(e / nz), c = (int)(e % nz);
        double v = 0.0;
        if (c <= r) v = cl[base + e] + (r == c ? cmax : 0.0);
        root[base + e] = v;
    }
}

// In-place lower Cholesky of root[l] (row-major, lower triangle holds the matrix), one CTA per
// matrix, left-looking over 32-wide block columns:
//   (1) panel update  A[kb:, kb:kb+32] -= L[kb:, :kb] L[kb:kb+32, :kb]^T  on the FP64 tensor cores
//       (DMMA.8x8x4; both operands are row slices of L, staged [row][k] through a cp.async double
//       buffer; 8 warps x 32 rows per pass),
//   (2) the 32 x 32 diagonal block factored by one warp in shared memory,
//   (3) the rows below solved against it by substitution, one row per thread.
// fail[l] = 1 if a pivot is <= 0 or NaN (LAPACK dpotrf info > 0).
constexpr int CH_KC = 16;                  // k per staged chunk
constexpr int CH_LD = 20;                  // [row][k] tiles, row pitch 20 doubles: conflict-free fragments, 16-byte rows
constexpr int CH_ROWS = 256;               // rows per pass (8 warps x 32)
constexpr int CH_STAGE = (CH_ROWS + CH_NB) * CH_LD;   // doubles per stage: row tile + the block column's own rows

template <bool FAST>
__global__ void __launch_bounds__(CH_THREADS, 2) cholesky_kernel(double* __restrict__ root, int nz, int* __restrict__ fail) {
    extern __shared__ __align__(16) double ch_smem[];
    double* stage = ch_smem;                              // [2][CH_STAGE]
    double (*Ld)[CH_NB + 1] = (double (*)[CH_NB + 1])(ch_smem + 2 * CH_STAGE);   // diagonal block
    __shared__ int s_fail;
    __shared__ double Linv[CH_NB];                        // reciprocals of the diagonal block's pivots
    double* A = root + (long long)blockIdx.x * nz * nz;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    if (tid == 0) s_fail = 0;
    __syncthreads();

    for (int kb = 0; kb < nz; kb += CH_NB) {
        const int nbk = min(CH_NB, nz - kb);
        // ---- (1) panel update with DMMA
        if (kb > 0) {
            const int nchunk = (kb + CH_KC - 1) / CH_KC;
            for (int r0 = kb; r0 < nz; r0 += CH_ROWS) {
                double acc[4][4][2];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
                auto load = [&](int ck, int slot) {
                    double* S = stage + slot * CH_STAGE;
                    const int p0 = ck * CH_KC;
                    // rows 0..255: the row tile; rows 256..287: the block column's own rows kb..kb+31
                    for (int e = tid; e < (CH_ROWS + CH_NB) * (CH_KC / 2); e += CH_THREADS) {
                        const int rr = e / (CH_KC / 2), kk = (e % (CH_KC / 2)) * 2;
                        const int r = (rr < CH_ROWS) ? r0 + rr : kb + (rr - CH_ROWS);
                        const int k = p0 + kk;
                        double* dst = S + rr * CH_LD + kk;
                        if (FAST) {
                            const bool ok = (r < nz) && (k < kb);   // kb is a multiple of 32: k, k+1 < kb together
                            cp_async16(dst, A + (ok ? ((long long)r * nz + k) : 0), ok);
                        } else {
                            dst[0] = (r < nz && k < kb) ? A[(long long)r * nz + k] : 0.0;
                            dst[1] = (r < nz && k + 1 < kb) ? A[(long long)r * nz + k + 1] : 0.0;
                        }
                    }
                    cp_async_commit();
                };
                load(0, 0);
                for (int ck = 0; ck < nchunk; ck++) {
                    cp_async_wait<0>();
                    __syncthreads();
                    if (ck + 1 < nchunk) load(ck + 1, (ck + 1) & 1);
                    const double* S = stage + (ck & 1) * CH_STAGE;
                    const double* Sa = S + warp * 32 * CH_LD;
                    const double* Sb = S + CH_ROWS * CH_LD;
                    if (r0 + warp * 32 < nz) {
#pragma unroll
                        for (int k4 = 0; k4 < CH_KC / 4; k4++) {
                            double af[4], bf[4];
#pragma unroll
                            for (int mb = 0; mb < 4; mb++) af[mb] = Sa[(8 * mb + g) * CH_LD + k4 * 4 + t];
#pragma unroll
                            for (int nb = 0; nb < 4; nb++) bf[nb] = Sb[(8 * nb + g) * CH_LD + k4 * 4 + t];
#pragma unroll
                            for (int mb = 0; mb < 4; mb++)
#pragma unroll
                                for (int nb = 0; nb < 4; nb++) dmma884(acc[mb][nb][0], acc[mb][nb][1], af[mb], bf[nb]);
                        }
                    }
                }
                __syncthreads();   // all fragment reads done before the next pass refills stage 0
                // C[g][2t], C[g][2t+1]: row 8 mb + g, columns 8 nb + 2t, + 1
#pragma unroll
                for (int mb = 0; mb < 4; mb++) {
                    const int r = r0 + warp * 32 + 8 * mb + g;
                    if (r >= nz) continue;
#pragma unroll
                    for (int nb = 0; nb < 4; nb++) {
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const int c = 8 * nb + 2 * t + q;
                            if (c < nbk && kb + c <= r) A[(long long)r * nz + kb + c] -= acc[mb][nb][q];
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- (2) diagonal block: one warp, lane = row
        for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
            const int rr = e >> 5, cc = e & 31;
            Ld[rr][cc] = (rr < nbk && cc <= rr) ? A[(long long)(kb + rr) * nz + kb + cc] : 0.0;
        }
        __syncthreads();
        if (warp == 0) {
            // lane = row, the row lives in registers; column jj of L is broadcast by shuffles
            // (no shared-memory read-modify-write chain: that loop was 60 % of a 256^2 factorisation)
            double r[CH_NB];
#pragma unroll
            for (int c = 0; c < CH_NB; c++) r[c] = Ld[lane][c];
            bool bad = false;
            double myinv = 0.0;
#pragma unroll
            for (int jj = 0; jj < CH_NB; jj++) {
                const double d = __shfl_sync(0xffffffffu, r[jj], jj);
                if (jj < nbk && !bad) {                       // warp-uniform
                    if (!(d > 0.0)) {
                        bad = true;
                    } else {
                        // one reciprocal square root per pivot (the divide and the square root are each a
                        // ~250-cycle software sequence on the critical path of all 32 pivots)
                        const double dinv = rsqrt(d);
                        const double dj = d * dinv;
                        double lij = r[jj] * dinv;            // rows above jj hold 0 here and stay 0
                        if (lane == jj) { lij = dj; myinv = dinv; }
                        r[jj] = lij;
#pragma unroll
                        for (int c = jj + 1; c < CH_NB; c++) {
                            const double lc = __shfl_sync(0xffffffffu, lij, c);   // L[c][jj]
                            r[c] -= lij * lc;                 // meaningful for lane >= c (lower triangle)
                        }
                    }
                }
            }
            if (bad && lane == 0) s_fail = 1;
#pragma unroll
            for (int c = 0; c < CH_NB; c++)
                if (c <= lane) Ld[lane][c] = r[c];
            Linv[lane] = myinv;
        }
        __syncthreads();
        if (s_fail) break;
        for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
            const int rr = e >> 5, cc = e & 31;
            if (rr < nbk && cc <= rr) A[(long long)(kb + rr) * nz + kb + cc] = Ld[rr][cc];
        }
        // ---- (3) triangular solve for the rows below: X L_kk^T = B, one row per thread, right-looking
        // (every finished x_p updates all later columns at once: independent FMAs instead of one
        // dependent chain per column; the pivots enter through their reciprocals)
        for (int r = kb + nbk + tid; r < nz; r += CH_THREADS) {
            double* Ar = A + (long long)r * nz + kb;
            double v[CH_NB];
#pragma unroll
            for (int cc = 0; cc < CH_NB; cc++) v[cc] = (cc < nbk) ? Ar[cc] : 0.0;
#pragma unroll
            for (int pp = 0; pp < CH_NB; pp++) {
                if (pp < nbk) {
                    const double xp = v[pp] * Linv[pp];
                    v[pp] = xp;
#pragma unroll
                    for (int cc = pp + 1; cc < CH_NB; cc++) v[cc] -= xp * Ld[cc][pp];   // Ld rows >= nbk are zero
                }
            }
#pragma unroll
            for (int cc = 0; cc < CH_NB; cc++)
                if (cc < nbk) Ar[cc] = v[cc];
        }
        __syncthreads();
    }
    if (tid == 0) fail[blockIdx.x] = s_fail;
}

// ----------------------------------------------------------------- Jacobi fallback
// G (rows) starts as the symmetrised jittered matrix; V starts as identity.  Rotating rows
// p,q of both by the same plane rotation until all rows of G are mutually orthogonal gives
// G = diag(lambda) V, rows of V the eigenvectors.  Round-robin ordering, one warp per pair.
__global__ void jacobi_init_kernel(const double* __restrict__ cl, const int* __restrict__ fail_list, int nz,
                                   double jitter_rel, const double* __restrict__ dmax, double* __restrict__ G,
                                   double* __restrict__ V, const int* __restrict__ nfail_ptr) {
    if (nfail_ptr && (int)blockIdx.x >= *nfail_ptr) return;   // device-side count: no host round trip
    const int l = fail_list[blockIdx.x];
    const long long src = (long long)l * nz * nz, dst = (long long)blockIdx.x * nz * nz;
    const double cmax = dmax[l] * jitter_rel;
    for (long long e = threadIdx.x; e < (long long)nz * nz; e += blockDim.x) {
        const int r = (int)(e / nz), c = (int)(e % nz);
        // LAPACK eigh(lower=True) reads the lower triangle only
        const double v = (c <= r) ? cl[src + e] : cl[src + (long long)c * nz + r];
        G[dst + e] = v + (r == c ? cmax : 0.0);
        V[dst + e] = (r == c) ? 1.0 : 0.0;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(1024) jacobi_kernel(double* __restrict__ Gall, double* __restrict__ Vall, int nz,
                                                      int max_sweeps, int* __restrict__ sweeps_out,
                                                      const int* __restrict__ nfail_ptr) {
    __shared__ int s_rot;
    if (nfail_ptr && (int)blockIdx.x >= *nfail_ptr) return;
    double* G = Gall + (long long)blockIdx.x * nz * nz;
    double* V = Vall + (long long)blockIdx.x * nz * nz;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int npl = nz + (nz & 1);          // players in the round-robin (pad to even)
    const int nsteps = npl - 1, npairs = npl / 2;
    {   // an all-zero matrix (C_0 of the l^-beta foreground spectra) is already diagonal: skip the sweep
        int nonzero = 0;
        for (long long e = threadIdx.x; e < (long long)nz * nz; e += blockDim.x) nonzero |= (G[e] != 0.0);
        if (!__syncthreads_or(nonzero)) {
            if (threadIdx.x == 0) sweeps_out[blockIdx.x] = 0;
            return;
        }
    }
    int sweep = 0;
    for (; sweep < max_sweeps; sweep++) {
        if (threadIdx.x == 0) s_rot = 0;
        __syncthreads();
        for (int step = 0; step < nsteps; step++) {
            for (int pi = warp; pi < npairs; pi += nwarp) {
                // circle method: player npl-1 fixed, others rotate
                int a = (pi == 0) ? npl - 1 : (step + pi) % (npl - 1);
                int b = (step + npl - 1 - pi) % (npl - 1);
                int p = min(a, b), q = max(a, b);
                if (q >= nz) continue;   // bye
                double* gp = G + (long long)p * nz;
                double* gq = G + (long long)q * nz;
                double app = 0, aqq = 0, apq = 0;
                for (int c = lane; c < nz; c += 32) {
                    const double x = gp[c], y = gq[c];
                    app = fma(x, x, app); aqq = fma(y, y, aqq); apq = fma(x, y, apq);
                }
                app = warp_sum(app); aqq = warp_sum(aqq); apq = warp_sum(apq);
                if (fabs(apq) <= 1e-15 * sqrt(app * aqq) || apq == 0.0) continue;
                const double zeta = (aqq - app) / (2.0 * apq);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int c = lane; c < nz; c += 32) {
                    const double x = gp[c], y = gq[c];
                    gp[c] = cs * x - sn * y;
                    gq[c] = sn * x + cs * y;
                }
                double* vp = V + (long long)p * nz;
                double* vq = V + (long long)q * nz;
                for (int c = lane; c < nz; c += 32) {
                    const double x = vp[c], y = vq[c];
                    vp[c] = cs * x - sn * y;
                    vq[c] = sn * x + cs * y;
                }
                if (lane == 0) s_rot = 1;
            }
            __syncthreads();
        }
        const int any = s_rot;
        __syncthreads();
        if (!any) break;
    }
    if (threadIdx.x == 0) sweeps_out[blockIdx.x] = sweep;
}

// eigenvalues lambda_i = v_i . g_i, clip, sort ascending, write root[:, rank] = v_i sqrt(lambda_i)
__global__ void __launch_bounds__(256) jacobi_finish_kernel(const double* __restrict__ Gall, const double* __restrict__ Vall,
                                                            const int* __restrict__ fail_list, int nz, double clip_rel,
                                                            double* __restrict__ root, int* __restrict__ num_pos,
                                                            double* __restrict__ evals_ws, int* __restrict__ rank_ws,
                                                            const int* __restrict__ nfail_ptr,
                                                            double* __restrict__ evals_sorted = nullptr,
                                                            int* __restrict__ used = nullptr) {
    // evals_sorted != nullptr: plain eigen-decomposition (scipy.linalg.eigh layout): column k of `root` is the
    // unit eigenvector of the k-th smallest eigenvalue, evals_sorted[l][k] that eigenvalue; no clipping
    __shared__ double s_max;
    __shared__ int s_npos;
    if (nfail_ptr && (int)blockIdx.x >= *nfail_ptr) return;
    const int l = fail_list[blockIdx.x];
    const double* G = Gall + (long long)blockIdx.x * nz * nz;
    const double* V = Vall + (long long)blockIdx.x * nz * nz;
    double* ev = evals_ws + (long long)blockIdx.x * nz;
    int* rank = rank_ws + (long long)blockIdx.x * nz;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = warp; i < nz; i += nwarp) {
        double s = 0.0;
        for (int c = lane; c < nz; c += 32) s = fma(V[(long long)i * nz + c], G[(long long)i * nz + c], s);
        s = warp_sum(s);
        if (lane == 0) ev[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double mx = -1.0e308;
        for (int i = 0; i < nz; i++) mx = fmax(mx, ev[i]);
        s_max = mx;
        s_npos = 0;
    }
    __syncthreads();
    const double thr = s_max * clip_rel;
    // rank of each eigenvalue in ascending order (ties broken by index)
    for (int i = threadIdx.x; i < nz; i += blockDim.x) {
        const double vi = ev[i];
        int rk = 0;
        for (int k = 0; k < nz; k++) {
            const double vk = ev[k];
            rk += (vk < vi) || (vk == vi && k < i);
        }
        rank[i] = rk;
        if (!(vi < thr) && vi != 0.0) atomicAdd(&s_npos, 1);
    }
    __syncthreads();
    double* R = root + (long long)l * nz * nz;
    for (long long e = threadIdx.x; e < (long long)nz * nz; e += blockDim.x) {
        const int i = (int)(e / nz), r = (int)(e % nz);     // eigenpair i, row r of the root
        const double lam = ev[i];
        const double sc = evals_sorted ? 1.0 : ((lam < thr) ? 0.0 : sqrt(fmax(lam, 0.0)));
        R[(long long)r * nz + rank[i]] = V[(long long)i * nz + r] * sc;
    }
    if (evals_sorted)
        for (int i = threadIdx.x; i < nz; i += blockDim.x) evals_sorted[(long long)l * nz + rank[i]] = ev[i];
    if (threadIdx.x == 0 && num_pos) num_pos[l] = s_npos;
    if (threadIdx.x == 0 && used) used[l] = 1 + (nz - s_npos);     // 1 + number of leading zero columns
}


// ----------------------------------------------------------------- pivoted-Cholesky fallback
// For large matrices the Jacobi eigen fallback is hopeless (11 s for one 1024^2 matrix: every
// rotation streams whole rows through L2), while the matrices that need a fallback are the
// numerically rank-deficient foreground covariances (rank ~50 of 1024).  A diagonally pivoted
// Cholesky with the reference's relative clip gives a root with the same defining property,
// M M^T = C + jitter to within clip_rel * trace element-wise, in O(nz^2 rank):
//   p = argmax d;  stop when d[p] <= clip_rel * trace;  L[:,k] = (A[:,p] - L[:, :k] L[p, :k]^T) / sqrt(d[p]);
//   d -= L[:,k]^2.
// Columns are stored like the eigen branch stores its own: discarded (zero) columns first, the
// retained ones last with the strongest in the last column; num_pos = rank.  One CTA per matrix.
// Lws: [slot][k][i] (column k of L contiguous in i).
__global__ void __launch_bounds__(1024) pchol_kernel(const double* __restrict__ cl, const int* __restrict__ fail_list, int nz,
                                                     double jitter_rel, const double* __restrict__ dmax, double clip_rel,
                                                     double* __restrict__ Lws_all, double* __restrict__ root,
                                                     int* __restrict__ num_pos, const int* __restrict__ nfail_ptr,
                                                     int* __restrict__ used) {
    extern __shared__ __align__(16) double pc_smem[];
    double* d = pc_smem;              // [nz] residual diagonal (-1 once a row has been a pivot)
    double* Lp = d + nz;              // [nz] row p of L
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ double s_piv, s_tau;
    __shared__ int s_p;
    if (nfail_ptr && (int)blockIdx.x >= *nfail_ptr) return;
    const int l = fail_list[blockIdx.x];
    const double* A = cl + (long long)l * nz * nz;
    double* Lws = Lws_all + (long long)blockIdx.x * nz * nz;
    const double cmax = dmax[l] * jitter_rel;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;

    double tr = 0.0;
    for (int i = tid; i < nz; i += blockDim.x) {
        const double v = A[(long long)i * nz + i] + cmax;
        d[i] = v;
        tr += v;
    }
    tr = warp_sum(tr);
    if (lane == 0) s_val[warp] = tr;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < nwarp; w++) t += s_val[w];
        s_tau = clip_rel * fmax(t, 0.0);
    }
    __syncthreads();
    int k = 0;
    for (; k < nz; k++) {
        // ---- pivot: largest residual diagonal entry, lowest index on ties
        double bv = -1.0e308;
        int bi = 0x7fffffff;
        for (int i = tid; i < nz; i += blockDim.x) {
            const double v = d[i];
            if (v > bv) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            double v = s_val[0];
            int ix = s_idx[0];
            for (int w = 1; w < nwarp; w++)
                if (s_val[w] > v || (s_val[w] == v && s_idx[w] < ix)) { v = s_val[w]; ix = s_idx[w]; }
            s_p = ix;
            s_piv = v;
        }
        __syncthreads();
        const int p = s_p;
        const double dp = s_piv;
        if (!(dp > s_tau) || !(dp > 0.0)) break;      // uniform
        const double piv = sqrt(dp);
        for (int j = tid; j < k; j += blockDim.x) Lp[j] = Lws[(long long)j * nz + p];
        __syncthreads();
        for (int i = tid; i < nz; i += blockDim.x) {
            double lik = 0.0;
            if (i == p) {
                lik = piv;
            } else if (d[i] >= 0.0) {                 // rows that were pivots before stay exactly zero
                // LAPACK-style: only the lower triangle of the input is read
                double v = (i < p) ? A[(long long)p * nz + i] : A[(long long)i * nz + p];
                for (int j = 0; j < k; j++) v = fma(-Lws[(long long)j * nz + i], Lp[j], v);
                lik = v / piv;
            }
            Lws[(long long)k * nz + i] = lik;
            if (i == p) d[i] = -1.0;
            else if (d[i] >= 0.0) d[i] = fmax(d[i] - lik * lik, 0.0);
        }
        __syncthreads();
    }
    // ---- root[i][nz - 1 - kk] = L[i][kk]; zero columns first
    double* R = root + (long long)l * nz * nz;
    for (long long e = tid; e < (long long)nz * nz; e += blockDim.x) {
        const int i = (int)(e / nz), c = (int)(e % nz);
        const int kk = nz - 1 - c;
        R[e] = (kk < k) ? Lws[(long long)kk * nz + i] : 0.0;
    }
    if (tid == 0) {
        num_pos[l] = k;
        if (used) used[l] = 1 + (nz - k);                          // 1 + number of leading zero columns
    }
}

__global__ void root_flags_kernel(const int* __restrict__ fail, int nl, int nz, int* __restrict__ used_eigh,
                                  int* __restrict__ num_pos, int* __restrict__ fail_list, int* __restrict__ nfail) {
    // single thread block: compact the failed indices (order preserved)
    if (threadIdx.x == 0) {
        int n = 0;
        for (int l = 0; l < nl; l++) {
            used_eigh[l] = fail[l];
            num_pos[l] = nz;
            if (fail[l]) fail_list[n++] = l;
        }
        *nfail = n;
    }
}

}  // namespace cb

using namespace cb;

// workspace: dmax[nl] | fail[nl] | fail_list[nl] | nfail | evals[nl*nz] | rank[nl*nz] | sweeps[nl] | G | V (eigh slots)
static long long root_fixed_bytes(int nl, int nz) {
    return 8LL * nl + 4LL * nl * 3 + 64 + 12LL * nl * nz + 10 * 256;
}

// Matrices up to this size take the Jacobi eigen fallback (the reference's exact eigen semantics);
// larger ones the pivoted Cholesky.  CORA_B200_JACOBI_MAX_NZ overrides (e.g. a huge value forces Jacobi).
static int g_jacobi_max_nz = [] { const char* e = getenv("CORA_B200_JACOBI_MAX_NZ"); return e ? atoi(e) : 128; }();

extern "C" long long cora_b200_root_workspace_bytes(int nl, int nz) {
    // room for every matrix to take the eigh path (2 nz^2 doubles each); a smaller workspace is
    // accepted and processed in waves
    return root_fixed_bytes(nl, nz) + 16LL * nz * nz * (long long)nl;
}

// max(diag) per matrix, optionally merged with the maxima already in dmax (merge != 0)
__global__ void diag_max_kernel(const double* __restrict__ cl, int nz, double* __restrict__ dmax, int merge) {
    __shared__ double red[256];
    const long long base = (long long)blockIdx.x * nz * nz;
    double mx = -1.0e308;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) mx = fmax(mx, cl[base + (long long)i * nz + i]);
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) dmax[blockIdx.x] = merge ? fmax(dmax[blockIdx.x], red[0]) : red[0];
}

extern "C" int cora_b200_diag_max(const double* cl, int nl, int nz, double* dmax, int merge, void* stream) {
    CB_REQUIRE(cl && dmax && nl >= 1 && nz >= 1, 1, "diag_max: bad arguments");
    diag_max_kernel<<<nl, 256, 0, (cudaStream_t)stream>>>(cl, nz, dmax, merge);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_root_batched(const double* cl, int nl, int nz, double jitter_rel, double clip_rel, double* root,
                                      int* used_eigh, int* num_pos, void* workspace, long long ws_bytes, void* stream) {
    return cora_b200_root_batched_block(cl, nl, nz, jitter_rel, clip_rel, nullptr, root, used_eigh, num_pos, workspace, ws_bytes, stream);
}

extern "C" int cora_b200_root_batched_block(const double* cl, int nl, int nz, double jitter_rel, double clip_rel,
                                            const double* diag_max, double* root, int* used_eigh, int* num_pos,
                                            void* workspace, long long ws_bytes, void* stream) {
    CB_REQUIRE(cl && root && used_eigh && num_pos && workspace, 1, "root_batched: null argument");
    CB_REQUIRE(nl >= 1 && nz >= 1, 1, "root_batched: bad sizes nl=%d nz=%d", nl, nz);
    CB_REQUIRE(ws_bytes >= root_fixed_bytes(nl, nz) + 16LL * nz * nz, 4,
               "root_batched: workspace too small (%lld B, need >= %lld B)", ws_bytes, root_fixed_bytes(nl, nz) + 16LL * nz * nz);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    auto take = [&](long long bytes) { char* p = ws; ws = (char*)(((uintptr_t)(ws + bytes) + 255) & ~(uintptr_t)255); return p; };
    double* dmax = (double*)take(8LL * nl);
    int* fail = (int*)take(4LL * nl);
    int* fail_list = (int*)take(4LL * nl);
    int* sweeps = (int*)take(4LL * nl);
    int* nfail_d = (int*)take(64);
    double* evals = (double*)take(8LL * nl * nz);
    int* rank = (int*)take(4LL * nl * nz);
    double* GV = (double*)ws;
    long long slots = ((char*)workspace + ws_bytes - ws) / (16LL * nz * nz);

    { KTimer kt(K_ROOT_PREP, st); root_prepare_kernel<<<dim3(nl, nl >= 1024 ? 2 : 8), 256, 0, st>>>(cl, nz, jitter_rel, root, dmax, diag_max); }
    count_launch();
    CB_LAUNCH_CHECK();
    {
        KTimer kt(K_CHOLESKY, st);
        const size_t smem = sizeof(double) * (2 * CH_STAGE + CH_NB * (CH_NB + 1));
        if (nz % 2 == 0) {
            CB_CUDA(cudaFuncSetAttribute(cholesky_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cholesky_kernel<true><<<nl, CH_THREADS, smem, st>>>(root, nz, fail);
        } else {
            CB_CUDA(cudaFuncSetAttribute(cholesky_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cholesky_kernel<false><<<nl, CH_THREADS, smem, st>>>(root, nz, fail);
        }
    }
    count_launch();
    CB_LAUNCH_CHECK();
    root_flags_kernel<<<1, 32, 0, st>>>(fail, nl, nz, used_eigh, num_pos, fail_list, nfail_d);
    count_launch();
    CB_LAUNCH_CHECK();
    const int threads = nz >= 512 ? 1024 : (nz >= 128 ? 512 : 256);
    auto wave = [&](int f0, int nb, const int* guard) -> int {
        double* G = GV;
        double* V = GV + (long long)nb * nz * nz;
        KTimer kt(K_EIGH, st);
        if (nz > g_jacobi_max_nz) {
            const size_t smem = sizeof(double) * 2 * (size_t)nz;
            CB_CUDA(cudaFuncSetAttribute(pchol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
            pchol_kernel<<<nb, nz >= 512 ? 1024 : 512, smem, st>>>(cl, fail_list + f0, nz, jitter_rel, dmax, clip_rel, G, root,
                                                                   num_pos, guard, used_eigh);
            count_launch();
            CB_LAUNCH_CHECK();
            return 0;
        }
        jacobi_init_kernel<<<nb, 256, 0, st>>>(cl, fail_list + f0, nz, jitter_rel, dmax, G, V, guard);
        count_launch();
        CB_LAUNCH_CHECK();
        jacobi_kernel<<<nb, threads, 0, st>>>(G, V, nz, 60, sweeps + f0, guard);
        count_launch();
        CB_LAUNCH_CHECK();
        jacobi_finish_kernel<<<nb, 256, 0, st>>>(G, V, fail_list + f0, nz, clip_rel, root, num_pos, evals, rank, guard, nullptr,
                                                 used_eigh);
        count_launch();
        CB_LAUNCH_CHECK();
        return 0;
    };
    if (slots >= nl) {
        // room for every matrix: launch the eigen fallback for all nl slots, CTAs beyond the
        // device-side failure count exit at once -- the host never waits for the count
        return wave(0, nl, nfail_d);
    }
    int nfail = 0;
    CB_CUDA(cudaMemcpyAsync(&nfail, nfail_d, sizeof(int), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    if (nfail == 0) return 0;
    CB_REQUIRE(slots >= 1, 4, "root_batched: no workspace for the eigen fallback");
    for (int f0 = 0; f0 < nfail; f0 += (int)slots) {
        const int nb = (int)std::min<long long>(slots, nfail - f0);
        if (int rc = wave(f0, nb, nullptr)) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------- plain eigh
__global__ void iota_kernel(int* p, int n, int base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = base + i;
}

extern "C" long long cora_b200_eigh_workspace_bytes(int nl, int nz) {
    return 4LL * nl + 8LL * nl + 12LL * nl * nz + 4LL * nl + 8 * 256 + 16LL * nz * nz * (long long)nl;
}

extern "C" int cora_b200_eigh_batched(const double* a, int nl, int nz, double* evecs, double* evals, void* workspace,
                                      long long ws_bytes, void* stream) {
    CB_REQUIRE(a && evecs && evals && workspace && nl >= 1 && nz >= 1, 1, "eigh_batched: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    auto take = [&](long long bytes) { char* p = ws; ws = (char*)(((uintptr_t)(ws + bytes) + 255) & ~(uintptr_t)255); return p; };
    int* list = (int*)take(4LL * nl);
    double* dzero = (double*)take(8LL * nl);
    double* ev_ws = (double*)take(8LL * nl * nz);
    int* rank = (int*)take(4LL * nl * nz);
    int* sweeps = (int*)take(4LL * nl);
    double* GV = (double*)ws;
    const long long slots = ((char*)workspace + ws_bytes - ws) / (16LL * nz * nz);
    CB_REQUIRE(slots >= 1, 4, "eigh_batched: workspace too small (%lld B)", ws_bytes);
    iota_kernel<<<ceil_div(nl, 256), 256, 0, st>>>(list, nl, 0);
    CB_CUDA(cudaMemsetAsync(dzero, 0, 8LL * nl, st));
    count_launch();
    CB_LAUNCH_CHECK();
    const int threads = nz >= 512 ? 1024 : (nz >= 128 ? 512 : 256);
    KTimer kt(K_EIGH, st);
    for (int f0 = 0; f0 < nl; f0 += (int)slots) {
        const int nb = (int)std::min<long long>(slots, nl - f0);
        double* G = GV;
        double* V = GV + (long long)nb * nz * nz;
        jacobi_init_kernel<<<nb, 256, 0, st>>>(a, list + f0, nz, 0.0, dzero, G, V, nullptr);
        jacobi_kernel<<<nb, threads, 0, st>>>(G, V, nz, 60, sweeps + f0, nullptr);
        jacobi_finish_kernel<<<nb, 256, 0, st>>>(G, V, list + f0, nz, 0.0, evecs, nullptr, ev_ws, rank, nullptr, evals);
        count_launch(3);
        CB_LAUNCH_CHECK();
    }
    return 0;
}
