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

// ----------------------------------------------------------------- Cholesky, fused tile passes
// Same left-looking algorithm and the same arithmetic per element as cholesky_kernel, restructured around what the
// ncu source page of the 1024-channel case showed (profiles/r02/ncu_chol_c3_pass1.txt): 14 % of the stall samples sat
// on the scattered global read-modify-write that ended the panel update, another ~8 % on re-reading the panel from
// global memory for the diagonal block and the row solve.  Here every 256-row tile of a block column makes ONE pass:
//   DMMA update into registers -> the tile (A - update) assembled in shared memory (coalesced 256-byte row reads) ->
//   [first tile: diagonal block factored by warp 0] -> row solve from shared memory, one row per thread ->
//   coalesced store of the finished L rows.
// The panel tile lives in the operand stage buffers, which are idle between two K loops.
constexpr int CH_PLD = 33;     // panel tile pitch: one row per thread reads conflict-free

// doubles of the operand ring of cholesky2_kernel (at least the 256 x 33 panel tile that aliases it)
__host__ __device__ constexpr int ch2_stage_doubles(int KC, int LD, int NST) {
    return (NST * (CH_ROWS + CH_NB) * LD > CH_ROWS * CH_PLD) ? NST * (CH_ROWS + CH_NB) * LD : CH_ROWS * CH_PLD;
}

// KC: k per staged chunk, LD: row pitch of the [row][k] operand tiles, NST: stages of the cp.async ring.
// (16, 20, 2): 64 DMMAs per warp between two CTA barriers, one chunk of prefetch; (8, 12, 3): 32 DMMAs, two chunks.
template <bool FAST, int KC, int LD, int NST>
__global__ void __launch_bounds__(CH_THREADS, 2) cholesky2_kernel(const double* __restrict__ cl, double* __restrict__ root, int nz,
                                                                  double jitter_rel, const double* __restrict__ dmax,
                                                                  int* __restrict__ fail) {
    // Reads the covariance (lower triangle of cl, + jitter_rel * dmax[l] on the diagonal: cora/core/skysim.py:116-117)
    // the one time each entry is needed and writes the finished root -- L below, exact zeros above the diagonal --
    // into `root`: the separate jitter-and-copy pass (root_prepare_kernel: 4 MB read + 8 MB written per 1024^2 matrix)
    // is gone.
    constexpr int STAGE = (CH_ROWS + CH_NB) * LD;
    extern __shared__ __align__(16) double ch_smem[];
    double* stage = ch_smem;                              // [NST][STAGE]; aliased by the panel tile P[256][CH_PLD]
    double* Pt = ch_smem;
    double (*Ld)[CH_NB + 1] = (double (*)[CH_NB + 1])(ch_smem + ch2_stage_doubles(KC, LD, NST));   // factored diagonal block
    __shared__ int s_fail;
    __shared__ double Linv[CH_NB];                        // reciprocals of the diagonal block's pivots
    double* A = root + (long long)blockIdx.x * nz * nz;
    const double* C = cl + (long long)blockIdx.x * nz * nz;
    const double cjit = dmax[blockIdx.x] * jitter_rel;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    if (tid == 0) s_fail = 0;
    __syncthreads();

    for (int kb = 0; kb < nz; kb += CH_NB) {
        const int nbk = min(CH_NB, nz - kb);
        const int nchunk = (kb + KC - 1) / KC;
        // zeros above the diagonal block of this block column (rows 0 .. kb-1), 256-byte runs
        for (int e = tid; e < kb * CH_NB; e += CH_THREADS) {
            const int rr = e >> 5, c = e & 31;
            if (c < nbk) A[(long long)rr * nz + kb + c] = 0.0;
        }
        for (int r0 = kb; r0 < nz; r0 += CH_ROWS) {
            // ---- (1) update of this tile's 256 x 32 panel entries on the tensor cores
            double acc[4][4][2];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
            if (kb > 0) {
                auto load = [&](int ck, int slot) {
                    double* S = stage + slot * STAGE;
                    const int p0 = ck * KC;
                    // rows 0..255: the row tile; rows 256..287: the block column's own rows kb..kb+31
                    for (int e = tid; e < (CH_ROWS + CH_NB) * (KC / 2); e += CH_THREADS) {
                        const int rr = e / (KC / 2), kk = (e % (KC / 2)) * 2;
                        const int r = (rr < CH_ROWS) ? r0 + rr : kb + (rr - CH_ROWS);
                        const int k = p0 + kk;
                        double* dst = S + rr * LD + kk;
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
#pragma unroll
                for (int st = 0; st < NST - 1; st++) {
                    if (st < nchunk) load(st, st); else cp_async_commit();
                }
                for (int ck = 0; ck < nchunk; ck++) {
                    cp_async_wait<NST - 2>();
                    __syncthreads();   // chunk ck landed for everyone; the slot refilled below was consumed one iteration ago
                    if (ck + NST - 1 < nchunk) load(ck + NST - 1, (ck + NST - 1) % NST); else cp_async_commit();
                    const double* S = stage + (ck % NST) * STAGE;
                    const double* Sa = S + warp * 32 * LD;
                    const double* Sb = S + CH_ROWS * LD;
                    if (r0 + warp * 32 < nz) {
#pragma unroll
                        for (int k4 = 0; k4 < KC / 4; k4++) {
                            double af[4], bf[4];
#pragma unroll
                            for (int mb = 0; mb < 4; mb++) af[mb] = Sa[(8 * mb + g) * LD + k4 * 4 + t];
#pragma unroll
                            for (int nb = 0; nb < 4; nb++) bf[nb] = Sb[(8 * nb + g) * LD + k4 * 4 + t];
#pragma unroll
                            for (int mb = 0; mb < 4; mb++)
#pragma unroll
                                for (int nb = 0; nb < 4; nb++) dmma884(acc[mb][nb][0], acc[mb][nb][1], af[mb], bf[nb]);
                        }
                    }
                }
                cp_async_wait<0>();
                __syncthreads();   // all fragment reads done: the stage buffers become the panel tile
            }
            // ---- (2) P = A[tile, kb:kb+32] - update.  The update is scattered from the fragments (C[g][2t], C[g][2t+1]:
            // row 8 mb + g, columns 8 nb + 2t, + 1), the matrix entries are added by coalesced row reads.
#pragma unroll
            for (int mb = 0; mb < 4; mb++) {
                double* pr = Pt + (warp * 32 + 8 * mb + g) * CH_PLD;
#pragma unroll
                for (int nb = 0; nb < 4; nb++) {
                    pr[8 * nb + 2 * t] = -acc[mb][nb][0];
                    pr[8 * nb + 2 * t + 1] = -acc[mb][nb][1];
                }
            }
            __syncthreads();
            for (int e = tid; e < CH_ROWS * CH_NB; e += CH_THREADS) {
                const int rr = e >> 5, c = e & 31;
                const int r = r0 + rr;
                double v = 0.0;
                if (r < nz && c < nbk && kb + c <= r) v = Pt[rr * CH_PLD + c] + (C[(long long)r * nz + kb + c] + (kb + c == r ? cjit : 0.0));
                Pt[rr * CH_PLD + c] = v;
            }
            __syncthreads();
            // ---- (3) first tile: the 32 x 32 diagonal block, one warp, lane = row (rows in registers, columns
            // broadcast by shuffles, one reciprocal square root per pivot)
            if (r0 == kb) {
                if (warp == 0) {
                    double r[CH_NB];
#pragma unroll
                    for (int c = 0; c < CH_NB; c++) r[c] = Pt[lane * CH_PLD + c];
                    bool bad = false;
                    double myinv = 0.0;
#pragma unroll
                    for (int jj = 0; jj < CH_NB; jj++) {
                        const double d = __shfl_sync(0xffffffffu, r[jj], jj);
                        if (jj < nbk && !bad) {                       // warp-uniform
                            if (!(d > 0.0)) {
                                bad = true;
                            } else {
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
                    for (int c = 0; c < CH_NB; c++) {
                        const double v = (c <= lane && lane < nbk) ? r[c] : 0.0;
                        Ld[lane][c] = v;
                        Pt[lane * CH_PLD + c] = v;
                    }
                    Linv[lane] = myinv;
                }
                __syncthreads();
                if (s_fail) break;
            }
            // ---- (4) rows below the diagonal block: X L_kk^T = B by substitution, one row per thread, right-looking
            {
                const int r = r0 + tid;
                if (r >= kb + nbk && r < nz) {
                    double* pr = Pt + tid * CH_PLD;
                    double v[CH_NB];
#pragma unroll
                    for (int cc = 0; cc < CH_NB; cc++) v[cc] = pr[cc];
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
                    for (int cc = 0; cc < CH_NB; cc++) pr[cc] = v[cc];
                }
            }
            __syncthreads();
            // ---- (5) finished rows of L back to global memory, 256-byte runs
            for (int e = tid; e < CH_ROWS * CH_NB; e += CH_THREADS) {
                const int rr = e >> 5, c = e & 31;
                const int r = r0 + rr;
                // (rows of the diagonal block: the zeros right of the diagonal are written too)
                if (r < nz && c < nbk && (kb + c <= r || r < kb + CH_NB)) A[(long long)r * nz + kb + c] = Pt[rr * CH_PLD + c];
            }
            __syncthreads();   // the next K loop refills the stage buffers
        }
        if (s_fail) break;
    }
    if (tid == 0) fail[blockIdx.x] = s_fail;
}

// ================================================================= eigen fallback
// Reference semantics (cora/util/nputil.py:86-96): eigh of the jittered matrix, eigenvalues below
// clip_rel * (largest eigenvalue) set to zero, root = evecs sqrt(evals), every column kept, ascending
// eigenvalue order (so the zeroed columns come first).  For the block-diagonal polarised covariance of
// cora/scripts/makesky.py:368-382 the jitter, the Cholesky-or-eigh decision and the clip threshold are
// those of the WHOLE matrix: blocks of one l share a slot group, `eig_max[l]` is the maximum over blocks.
//
// Two implementations of the same definition:
//  * exact: one-sided (Hestenes) Jacobi on the full nz x nz matrix (below) -- always for nz <= g_jacobi_max_nz,
//    and for any matrix the low-rank route cannot certify;
//  * low-rank (nz > g_jacobi_max_nz): the matrices that fail Cholesky are numerically rank-deficient covariances
//    (the smooth foreground spectra keep ~10-60 of 1024 modes above the clip).  A = L L^T by diagonally pivoted
//    Cholesky of the UN-jittered matrix down to the jitter level, the residual A - L L^T is verified to be at
//    that level element by element, the r columns of L are orthogonalised by one-sided Jacobi (L W = U,
//    eigenvalues lambda_k = |u_k|^2 of A to high relative accuracy), and the eigenpairs of the jittered matrix
//    are (lambda_k + c, u_k / |u_k|); the remaining nz - r eigenvalues equal c (+ the certified residual) and
//    are clipped because the route is only taken when c < clip threshold.  O(nz^2 r) instead of O(nz^3 sweeps).
constexpr int MAXB = 8;
struct RootBlocks {
    const double* cl[MAXB];
    double* root[MAXB];
    int nb;
};

struct WaveCtx {
    const int* fail_list;    // already offset by f0
    const int* nfail_ptr;    // device count of failed l (NULL: every slot group of the wave is live)
    int f0;
    int nl;                  // per-block stride of used / num_pos
    int nz;
    double jitter_rel, clip_rel;
    const double* dmax;      // [nl] global diagonal maximum
    double* eig_max;         // [nl] global maximum eigenvalue of the jittered matrix (atomic max)
    double* slots;           // per slot: G / L^T (nz^2 doubles) then V (nz^2 doubles)
    double* evals;           // [slot][nz]
    int* rank;               // [slot][nz]
    int* lr_rank;            // [slot]
    int* status;             // [slot] 0 = low-rank route holds the decomposition, 1 = exact Jacobi
    int* sweeps;             // [slot]
};

// slot -> (l, block); false if the slot is beyond the device-side failure count
__device__ __forceinline__ bool wave_slot(const WaveCtx& w, int nb, int slot, int& l, int& b) {
    const int wi = slot / nb;
    b = slot - wi * nb;
    if (w.nfail_ptr && w.f0 + wi >= *w.nfail_ptr) return false;
    l = w.fail_list[wi];
    return true;
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    // for non-negative doubles the IEEE bit pattern orders like the unsigned integer
    atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(fmax(v, 0.0)));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ----------------------------------------------------------------- exact: full Jacobi
// G (rows) starts as the symmetrised jittered matrix; V starts as identity.  Rotating rows
// p,q of both by the same plane rotation until all rows of G are mutually orthogonal gives
// G = diag(lambda) V, rows of V the eigenvectors.  Round-robin ordering, one warp per pair.
__global__ void jacobi_init_kernel(RootBlocks B, WaveCtx w) {
    int l, b;
    if (!wave_slot(w, B.nb, blockIdx.x, l, b) || w.status[blockIdx.x] != 1) return;
    const int nz = w.nz;
    const double* cl = B.cl[b];
    const long long src = (long long)l * nz * nz;
    double* G = w.slots + (long long)blockIdx.x * 2 * nz * nz;
    double* V = G + (long long)nz * nz;
    const double cmax = w.dmax[l] * w.jitter_rel;
    for (long long e = threadIdx.x; e < (long long)nz * nz; e += blockDim.x) {
        const int r = (int)(e / nz), c = (int)(e % nz);
        // LAPACK eigh(lower=True) reads the lower triangle only
        const double v = (c <= r) ? cl[src + e] : cl[src + (long long)c * nz + r];
        G[e] = v + (r == c ? cmax : 0.0);
        V[e] = (r == c) ? 1.0 : 0.0;
    }
}

// One-sided Jacobi on `nrow` rows of length `len` (G), optionally carrying V (same shape) along.
// Returns the number of sweeps.  All threads of the CTA call it.
__device__ int hestenes_sweeps(double* __restrict__ G, double* __restrict__ V, int nrow, int len, int max_sweeps, int* s_rot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int npl = nrow + (nrow & 1);          // players in the round-robin (pad to even)
    const int nsteps = npl - 1, npairs = npl / 2;
    int sweep = 0;
    for (; sweep < max_sweeps; sweep++) {
        if (threadIdx.x == 0) *s_rot = 0;
        __syncthreads();
        for (int step = 0; step < nsteps; step++) {
            for (int pi = warp; pi < npairs; pi += nwarp) {
                // circle method: player npl-1 fixed, others rotate
                int a = (pi == 0) ? npl - 1 : (step + pi) % (npl - 1);
                int b = (step + npl - 1 - pi) % (npl - 1);
                int p = min(a, b), q = max(a, b);
                if (q >= nrow) continue;   // bye
                double* gp = G + (long long)p * len;
                double* gq = G + (long long)q * len;
                double app = 0, aqq = 0, apq = 0;
                for (int c = lane; c < len; c += 32) {
                    const double x = gp[c], y = gq[c];
                    app = fma(x, x, app); aqq = fma(y, y, aqq); apq = fma(x, y, apq);
                }
                app = warp_sum(app); aqq = warp_sum(aqq); apq = warp_sum(apq);
                if (fabs(apq) <= 1e-15 * sqrt(app * aqq) || apq == 0.0) continue;
                const double zeta = (aqq - app) / (2.0 * apq);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int c = lane; c < len; c += 32) {
                    const double x = gp[c], y = gq[c];
                    gp[c] = cs * x - sn * y;
                    gq[c] = sn * x + cs * y;
                }
                if (V) {
                    double* vp = V + (long long)p * len;
                    double* vq = V + (long long)q * len;
                    for (int c = lane; c < len; c += 32) {
                        const double x = vp[c], y = vq[c];
                        vp[c] = cs * x - sn * y;
                        vq[c] = sn * x + cs * y;
                    }
                }
                if (lane == 0) *s_rot = 1;
            }
            __syncthreads();
        }
        const int any = *s_rot;
        __syncthreads();
        if (!any) break;
    }
    return sweep;
}

__global__ void __launch_bounds__(1024) jacobi_kernel(RootBlocks B, WaveCtx w, int max_sweeps) {
    __shared__ int s_rot;
    int l, b;
    if (!wave_slot(w, B.nb, blockIdx.x, l, b) || w.status[blockIdx.x] != 1) return;
    const int nz = w.nz;
    double* G = w.slots + (long long)blockIdx.x * 2 * nz * nz;
    double* V = G + (long long)nz * nz;
    {   // a diagonal matrix (all-zero C_0 of the l^-beta foreground spectra, jitter only) needs no sweep
        int offdiag = 0;
        for (long long e = threadIdx.x; e < (long long)nz * nz; e += blockDim.x)
            offdiag |= (G[e] != 0.0) && (e / nz != e % nz);
        if (!__syncthreads_or(offdiag)) {
            if (threadIdx.x == 0) w.sweeps[blockIdx.x] = 0;
            return;
        }
    }
    const int sweep = hestenes_sweeps(G, V, nz, nz, max_sweeps, &s_rot);
    if (threadIdx.x == 0) w.sweeps[blockIdx.x] = sweep;
}

// eigenvalues lambda_i = v_i . g_i of an exact-route slot; the largest goes into eig_max[l]
__global__ void __launch_bounds__(256) jacobi_evals_kernel(RootBlocks B, WaveCtx w) {
    __shared__ double s_red[8];
    int l, b;
    if (!wave_slot(w, B.nb, blockIdx.x, l, b) || w.status[blockIdx.x] != 1) return;
    const int nz = w.nz;
    const double* G = w.slots + (long long)blockIdx.x * 2 * nz * nz;
    const double* V = G + (long long)nz * nz;
    double* ev = w.evals + (long long)blockIdx.x * nz;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double mx = 0.0;
    for (int i = warp; i < nz; i += nwarp) {
        double s = 0.0;
        for (int c = lane; c < nz; c += 32) s = fma(V[(long long)i * nz + c], G[(long long)i * nz + c], s);
        s = warp_sum(s);
        if (lane == 0) ev[i] = s;
        mx = fmax(mx, s);
    }
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < nwarp; k++) mx = fmax(mx, s_red[k]);
        atomic_max_nonneg(w.eig_max + l, mx);
    }
}

// clip against the global maximum, sort ascending, write root[:, rank] = v_i sqrt(lambda_i)
__global__ void __launch_bounds__(256) jacobi_finish_kernel(RootBlocks B, WaveCtx w, int* __restrict__ num_pos,
                                                            int* __restrict__ used,
                                                            double* __restrict__ evals_sorted = nullptr) {
    // evals_sorted != nullptr: plain eigen-decomposition (scipy.linalg.eigh layout): column k of `root` is the
    // unit eigenvector of the k-th smallest eigenvalue, evals_sorted[l][k] that eigenvalue; no clipping
    __shared__ int s_npos;
    int l, b;
    if (!wave_slot(w, B.nb, blockIdx.x, l, b) || w.status[blockIdx.x] != 1) return;
    const int nz = w.nz;
    const double* G = w.slots + (long long)blockIdx.x * 2 * nz * nz;
    const double* V = G + (long long)nz * nz;
    (void)G;
    const double* ev = w.evals + (long long)blockIdx.x * nz;
    int* rank = w.rank + (long long)blockIdx.x * nz;
    if (threadIdx.x == 0) s_npos = 0;
    __syncthreads();
    const double thr = w.eig_max[l] * w.clip_rel;
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
    double* R = B.root[b] + (long long)l * nz * nz;
    for (long long e = threadIdx.x; e < (long long)nz * nz; e += blockDim.x) {
        const int i = (int)(e / nz), r = (int)(e % nz);     // eigenpair i, row r of the root
        const double lam = ev[i];
        const double sc = evals_sorted ? 1.0 : ((lam < thr) ? 0.0 : sqrt(fmax(lam, 0.0)));
        R[(long long)r * nz + rank[i]] = V[(long long)i * nz + r] * sc;
    }
    if (evals_sorted)
        for (int i = threadIdx.x; i < nz; i += blockDim.x) evals_sorted[(long long)l * nz + rank[i]] = ev[i];
    if (threadIdx.x == 0 && num_pos) num_pos[(long long)b * w.nl + l] = s_npos;
    if (threadIdx.x == 0 && used) used[(long long)b * w.nl + l] = 1 + (nz - s_npos);     // 1 + number of leading zero columns
}

// ----------------------------------------------------------------- low-rank route, step 1: pivoted Cholesky
//   p = argmax d;  stop when d[p] <= tau;  L[:,k] = (A[:,p] - L[:, :k] L[p, :k]^T) / sqrt(d[p]);  d -= L[:,k]^2
// tau = max(jitter / 4, 32 eps max(diag)): the jitter level when there is one (everything below it is clipped
// together with the jitter itself), else the round-off level of the matrix entries.  Then the certificate:
// max |A - L L^T| over the lower triangle (what LAPACK reads) must be <= 4 tau + 64 eps max(diag) -- true for
// a positive semi-definite residual with diagonal <= tau; a matrix with a significant negative eigenvalue
// fails it and is handed to the exact route (status 1).
// L^T is stored in the slot's first nz^2 doubles as [k][i] (column k of L contiguous in i).
__global__ void __launch_bounds__(1024) lr_pchol_kernel(RootBlocks B, WaveCtx w) {
    extern __shared__ __align__(16) double pc_smem[];
    const int nz = w.nz;
    double* d = pc_smem;              // [nz] residual diagonal (-1 once a row has been a pivot)
    double* Lp = d + nz;              // [nz] row p of L
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ double s_piv, s_tau, s_dmax;
    __shared__ int s_p;
    int l, b;
    if (!wave_slot(w, B.nb, blockIdx.x, l, b)) return;
    const double* A = B.cl[b] + (long long)l * nz * nz;
    double* Lws = w.slots + (long long)blockIdx.x * 2 * nz * nz;
    const double cmax = w.dmax[l] * w.jitter_rel;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;

    double mx = 0.0;
    for (int i = tid; i < nz; i += blockDim.x) {
        const double v = A[(long long)i * nz + i];
        d[i] = v;
        mx = fmax(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_val[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int k = 0; k < nwarp; k++) t = fmax(t, s_val[k]);
        s_dmax = t;
        s_tau = fmax(0.25 * cmax, 32.0 * 2.220446049250313e-16 * t);
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
            for (int q = 1; q < nwarp; q++)
                if (s_val[q] > v || (s_val[q] == v && s_idx[q] < ix)) { v = s_val[q]; ix = s_idx[q]; }
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
    // ---- certificate: |A - L L^T| over the lower triangle, in 4 x 4 register blocks (bi >= bj): per k a thread loads 4
    // values of column block bi (the same for its neighbours: broadcast) and 4 of bj (consecutive across threads) for 16
    // FMAs.  (One thread per column walking the rows read two operands per FMA from L2 -- 500 MB per 1024^2 matrix,
    // 3.3 ms per CTA, the largest part of the fallback.)
    const double bound = 4.0 * s_tau + 64.0 * 2.220446049250313e-16 * s_dmax;
    int bad = 0;
    const int nblk = (nz + 3) >> 2;
    const long long ntile = (long long)nblk * (nblk + 1) / 2;
    for (long long tix = tid; tix < ntile; tix += blockDim.x) {
        // tile index -> (bi, bj <= bi), row-major over the lower triangle of blocks
        int bi = (int)((sqrt(8.0 * (double)tix + 1.0) - 1.0) * 0.5);
        while ((long long)(bi + 1) * (bi + 2) / 2 <= tix) bi++;
        while ((long long)bi * (bi + 1) / 2 > tix) bi--;
        const int bj = (int)(tix - (long long)bi * (bi + 1) / 2);
        const int i0 = 4 * bi, j0 = 4 * bj;
        double acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] = 0.0;
        for (int kk = 0; kk < k; kk++) {
            const double* row = Lws + (long long)kk * nz;
            double li[4], lj[4];
#pragma unroll
            for (int r = 0; r < 4; r++) { li[r] = (i0 + r < nz) ? row[i0 + r] : 0.0; lj[r] = (j0 + r < nz) ? row[j0 + r] : 0.0; }
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = fma(li[r], lj[c], acc[r][c]);
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int i = i0 + r, j = j0 + c;
                if (i < nz && j <= i) bad |= !(fabs(A[(long long)i * nz + j] - acc[r][c]) <= bound);
            }
    }
    bad = __syncthreads_or(bad);
    if (tid == 0) {
        w.lr_rank[blockIdx.x] = k;
        w.status[blockIdx.x] = bad ? 1 : 0;
    }
}

// step 2: orthogonalise the r columns of L (rows of L^T) -> u_k, lambda_k = |u_k|^2; largest + jitter -> eig_max
constexpr int LJ_RMAX = 64;     // largest factor rank that is Gram-preconditioned (G and V in shared memory)
constexpr int LJ_XC = 128;      // columns of L^T transformed per chunk
__global__ void __launch_bounds__(1024) lr_jacobi_kernel(RootBlocks B, WaveCtx w, int max_sweeps, int precondition) {
    __shared__ int s_rot;
    __shared__ double s_red[32];
    int l, b;
    if (!wave_slot(w, B.nb, blockIdx.x, l, b) || w.status[blockIdx.x] != 0) return;
    const int nz = w.nz;
    const int r = w.lr_rank[blockIdx.x];
    double* Lt = w.slots + (long long)blockIdx.x * 2 * nz * nz;
    double* lam = w.evals + (long long)blockIdx.x * nz;
    const double cmax = w.dmax[l] * w.jitter_rel;
    int sweep = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (r > 1 && r <= LJ_RMAX && precondition) {
        // Gram preconditioning.  The one-sided Jacobi below streams the r rows (r x nz doubles, 480 KB for rank 60 at
        // 1024 channels: L2, not shared memory) twice per pair and needs ~8-10 sweeps from the pivoted-Cholesky factor.
        // The r x r Gram matrix G = L^T L fits shared memory: its eigenvectors W (one-sided Jacobi on G with accumulated
        // rotations, all in shared memory) make the rows of W^T L^T orthogonal up to the accuracy G allows (it squares
        // the condition number), after which the streaming Jacobi -- which restores the high relative accuracy --
        // converges in one or two sweeps.
        extern __shared__ __align__(16) double lj_smem[];
        double* Gs = lj_smem;                   // [r][r]
        double* Vs = Gs + LJ_RMAX * LJ_RMAX;    // [r][r]
        double* Xs = Vs + LJ_RMAX * LJ_RMAX;    // [r][LJ_XC] column chunk of L^T
        // (1) G = L^T L in 4 x 4 blocks of (p, q <= p): a warp streams 8 rows once for 16 dot products
        const int nb4 = (r + 3) >> 2;
        for (int tile = warp; tile < nb4 * (nb4 + 1) / 2; tile += nwarp) {
            int bp = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
            while ((bp + 1) * (bp + 2) / 2 <= tile) bp++;
            while (bp * (bp + 1) / 2 > tile) bp--;
            const int bq = tile - bp * (bp + 1) / 2;
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b2 = 0; b2 < 4; b2++) acc[a][b2] = 0.0;
            for (int c = lane; c < nz; c += 32) {
                double xp[4], xq[4];
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    xp[a] = (4 * bp + a < r) ? Lt[(long long)(4 * bp + a) * nz + c] : 0.0;
                    xq[a] = (4 * bq + a < r) ? Lt[(long long)(4 * bq + a) * nz + c] : 0.0;
                }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b2 = 0; b2 < 4; b2++) acc[a][b2] = fma(xp[a], xq[b2], acc[a][b2]);
            }
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b2 = 0; b2 < 4; b2++) {
                    const double v = warp_sum(acc[a][b2]);
                    const int pp = 4 * bp + a, qq = 4 * bq + b2;
                    if (lane == 0 && pp < r && qq < r) { Gs[pp * r + qq] = v; Gs[qq * r + pp] = v; }
                }
        }
        for (int e = threadIdx.x; e < r * r; e += blockDim.x) Vs[e] = (e / r == e % r) ? 1.0 : 0.0;
        __syncthreads();
        // (2) rows of Vs <- eigenvectors of G
        hestenes_sweeps(Gs, Vs, r, r, max_sweeps, &s_rot);
        __syncthreads();
        // (3) L^T <- V L^T, column chunk by column chunk (the chunk's old values of all rows are staged first)
        for (int c0 = 0; c0 < nz; c0 += LJ_XC) {
            const int nc = min(LJ_XC, nz - c0);
            for (int e = threadIdx.x; e < r * LJ_XC; e += blockDim.x) {
                const int pp = e / LJ_XC, cc = e - pp * LJ_XC;
                Xs[e] = (cc < nc) ? Lt[(long long)pp * nz + c0 + cc] : 0.0;
            }
            __syncthreads();
            for (int e = threadIdx.x; e < r * LJ_XC; e += blockDim.x) {
                const int kk = e / LJ_XC, cc = e - kk * LJ_XC;
                if (cc >= nc) continue;
                double v = 0.0;
                for (int pp = 0; pp < r; pp++) v = fma(Vs[kk * r + pp], Xs[pp * LJ_XC + cc], v);
                Lt[(long long)kk * nz + c0 + cc] = v;
            }
            __syncthreads();
        }
    }
    if (r > 1) sweep = hestenes_sweeps(Lt, nullptr, r, nz, max_sweeps, &s_rot);
    double mx = 0.0;
    for (int k = warp; k < r; k += nwarp) {
        double s = 0.0;
        for (int c = lane; c < nz; c += 32) { const double x = Lt[(long long)k * nz + c]; s = fma(x, x, s); }
        s = warp_sum(s);
        if (lane == 0) lam[k] = s;
        mx = fmax(mx, s);
    }
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < nwarp; k++) mx = fmax(mx, s_red[k]);
        atomic_max_nonneg(w.eig_max + l, mx + cmax);
        w.sweeps[blockIdx.x] = sweep;
    }
}

// step 3: the low-rank route is only valid if the nz - r eigenvalues it does not hold (all equal to the jitter c)
// are clipped, i.e. c < clip_rel * eig_max; otherwise the reference keeps the null space (scaled sqrt(c)) and the
// exact route has to produce it.  One thread per slot.
__global__ void lr_decide_kernel(RootBlocks B, WaveCtx w, int nslots) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nslots) return;
    int l, b;
    if (!wave_slot(w, B.nb, slot, l, b) || w.status[slot] != 0) return;
    const double c = w.dmax[l] * w.jitter_rel;
    if (w.lr_rank[slot] < w.nz && c > 0.0 && !(c < w.clip_rel * w.eig_max[l])) w.status[slot] = 1;
}

// step 4: root[:, nz - r + rank_k] = u_k sqrt((lambda_k + c) / lambda_k), zero where lambda_k + c is clipped
__global__ void __launch_bounds__(256) lr_finish_kernel(RootBlocks B, WaveCtx w, int* __restrict__ num_pos,
                                                        int* __restrict__ used) {
    extern __shared__ __align__(16) double lf_smem[];
    __shared__ int s_npos;
    int l, b;
    if (!wave_slot(w, B.nb, blockIdx.x, l, b) || w.status[blockIdx.x] != 0) return;
    const int nz = w.nz;
    const int r = w.lr_rank[blockIdx.x];
    double* scale = lf_smem;                 // [nz] by column position
    int* src = (int*)(lf_smem + nz);         // [nz] column position -> k
    const double* Lt = w.slots + (long long)blockIdx.x * 2 * nz * nz;
    const double* lam = w.evals + (long long)blockIdx.x * nz;
    const double c = w.dmax[l] * w.jitter_rel;
    const double thr = w.clip_rel * w.eig_max[l];
    if (threadIdx.x == 0) s_npos = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < r; k += blockDim.x) {
        const double vk = lam[k];
        int rk = 0;
        for (int q = 0; q < r; q++) {
            const double vq = lam[q];
            rk += (vq < vk) || (vq == vk && q < k);
        }
        const double mu = vk + c;
        const bool keep = !(mu < thr) && vk > 0.0;
        src[rk] = k;
        scale[rk] = keep ? sqrt(mu / vk) : 0.0;
        if (keep) atomicAdd(&s_npos, 1);
    }
    __syncthreads();
    double* R = B.root[b] + (long long)l * nz * nz;
    const int c0 = nz - r;
    for (long long e = threadIdx.x; e < (long long)nz * nz; e += blockDim.x) {
        const int i = (int)(e / nz), col = (int)(e % nz);
        double v = 0.0;
        if (col >= c0) {
            const double sc = scale[col - c0];
            if (sc != 0.0) v = Lt[(long long)src[col - c0] * nz + i] * sc;
        }
        R[e] = v;
    }
    if (threadIdx.x == 0) {
        num_pos[(long long)b * w.nl + l] = s_npos;
        used[(long long)b * w.nl + l] = 1 + (nz - s_npos);
    }
}

__global__ void fill_int_kernel(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// fail_any[l] = OR over blocks; used / num_pos defaults; failed l compacted (order preserved);
// zero_scale[l] = sqrt(jitter): root of an implicit all-zero block on the Cholesky branch
__global__ void root_flags_kernel(const int* __restrict__ fail, int nb, int nl, int nz, double jitter_rel,
                                  const double* __restrict__ dmax, int* __restrict__ used_eigh, int* __restrict__ num_pos,
                                  int* __restrict__ fail_list, int* __restrict__ nfail, double* __restrict__ zero_scale) {
    if (threadIdx.x == 0) {
        int n = 0;
        for (int l = 0; l < nl; l++) {
            int any = 0;
            for (int b = 0; b < nb; b++) any |= fail[b * nl + l];
            for (int b = 0; b < nb; b++) {
                used_eigh[b * nl + l] = any;
                num_pos[b * nl + l] = nz;
            }
            if (any) fail_list[n++] = l;
            if (zero_scale) zero_scale[l] = sqrt(fmax(jitter_rel * dmax[l], 0.0));
        }
        *nfail = n;
    }
}

// implicit all-zero blocks (Stokes V) on the eigen branch: every eigenvalue equals the jitter c, kept iff c >= thr
__global__ void zero_scale_kernel(WaveCtx w, int nwave, double* __restrict__ zero_scale) {
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwave) return;
    if (w.nfail_ptr && w.f0 + wi >= *w.nfail_ptr) return;
    const int l = w.fail_list[wi];
    const double c = w.dmax[l] * w.jitter_rel;
    zero_scale[l] = (c > 0.0 && !(c < w.clip_rel * fmax(w.eig_max[l], c))) ? sqrt(c) : 0.0;
}

}  // namespace cb

using namespace cb;

// workspace (nb blocks): dmax[nl] dmax2[nl] eig_max[nl] | fail[nb nl] fail_list[nl] nfail | per slot (nb nl):
// sweeps, lr_rank, status, evals[nz], rank[nz] | slots: 2 nz^2 doubles each
static long long root_fixed_bytes(int nb, int nl, int nz) {
    return 24LL * nl + 4LL * nb * nl + 4LL * nl + 64 + 12LL * nb * nl + 12LL * nb * nl * nz + 12 * 256;
}

// Matrices up to this size always take the exact Jacobi eigen fallback; larger ones the low-rank route first.
// CORA_B200_JACOBI_MAX_NZ overrides (e.g. a huge value forces the exact route everywhere).
// 1: the first Cholesky kernel (separate read-modify-write / diagonal / row-solve phases), for A/B runs
static const bool g_chol_v1 = [] { const char* e = getenv("CORA_B200_CHOL_V1"); return e && e[0] == '1'; }();
// low-rank eigen route: Gram-precondition the one-sided Jacobi on the factor (CORA_B200_LR_PRECOND=0 turns it off)
static const bool g_lr_precond = [] { const char* e = getenv("CORA_B200_LR_PRECOND"); return !(e && e[0] == '0'); }();
// operand ring of the fused-tile Cholesky: 0 = 2 stages of 16 k, 1 = 3 stages of 8 k (deeper prefetch, more barriers)
static const bool g_chol_deep = [] { const char* e = getenv("CORA_B200_CHOL_DEEP"); return e && e[0] == '1'; }();
static int g_jacobi_max_nz = [] { const char* e = getenv("CORA_B200_JACOBI_MAX_NZ"); return e ? atoi(e) : 128; }();

extern "C" long long cora_b200_root_workspace_bytes(int nl, int nz) {
    // room for every matrix to take the eigen path (2 nz^2 doubles each); a smaller workspace is
    // accepted and processed in waves
    return root_fixed_bytes(1, nl, nz) + 16LL * nz * nz * (long long)nl;
}

extern "C" long long cora_b200_root_multi_workspace_bytes(int nblocks, int nl, int nz) {
    return root_fixed_bytes(nblocks, nl, nz) + 16LL * nz * nz * (long long)nl * nblocks;
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
    CB_REQUIRE(cl && root, 1, "root_batched: null argument");
    const double* cls[1] = {cl};
    double* roots[1] = {root};
    return cora_b200_root_batched_multi(cls, 1, nl, nz, jitter_rel, clip_rel, roots, used_eigh, num_pos, nullptr, workspace,
                                        ws_bytes, stream);
}

extern "C" int cora_b200_root_batched_multi(const double* const* cl_blocks, int nblocks, int nl, int nz, double jitter_rel,
                                            double clip_rel, double* const* root_blocks, int* used_eigh, int* num_pos,
                                            double* zero_block_scale, void* workspace, long long ws_bytes, void* stream) {
    CB_REQUIRE(cl_blocks && root_blocks && used_eigh && num_pos && workspace, 1, "root_batched: null argument");
    CB_REQUIRE(nl >= 1 && nz >= 1 && nblocks >= 1 && nblocks <= MAXB, 1, "root_batched: bad sizes nl=%d nz=%d nblocks=%d", nl, nz,
               nblocks);
    const int nb = nblocks;
    const long long slot_bytes = 16LL * nz * nz;
    CB_REQUIRE(ws_bytes >= root_fixed_bytes(nb, nl, nz) + slot_bytes * nb, 4,
               "root_batched: workspace too small (%lld B, need >= %lld B)", ws_bytes, root_fixed_bytes(nb, nl, nz) + slot_bytes * nb);
    RootBlocks B;
    B.nb = nb;
    for (int b = 0; b < nb; b++) {
        CB_REQUIRE(cl_blocks[b] && root_blocks[b], 1, "root_batched: null block pointer");
        B.cl[b] = cl_blocks[b];
        B.root[b] = root_blocks[b];
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    auto take = [&](long long bytes) { char* p = ws; ws = (char*)(((uintptr_t)(ws + bytes) + 255) & ~(uintptr_t)255); return p; };
    double* dmax = (double*)take(8LL * nl);
    double* dmax2 = (double*)take(8LL * nl);
    double* eig_max = (double*)take(8LL * nl);
    int* fail = (int*)take(4LL * nb * nl);
    int* fail_list = (int*)take(4LL * nl);
    int* nfail_d = (int*)take(64);
    int* sweeps = (int*)take(4LL * nb * nl);
    int* lr_rank = (int*)take(4LL * nb * nl);
    int* status = (int*)take(4LL * nb * nl);
    double* evals = (double*)take(8LL * nb * nl * nz);
    int* rank = (int*)take(4LL * nb * nl * nz);
    double* slot_mem = (double*)ws;
    const long long groups = ((char*)workspace + ws_bytes - ws) / (slot_bytes * nb);   // l's whose blocks fit at once

    for (int b = 0; b < nb; b++) {
        diag_max_kernel<<<nl, 256, 0, st>>>(B.cl[b], nz, dmax, b > 0);
        count_launch();
    }
    CB_LAUNCH_CHECK();
    for (int b = 0; b < nb; b++) {
        if (g_chol_v1) {
            { KTimer kt(K_ROOT_PREP, st); root_prepare_kernel<<<dim3(nl, nl >= 1024 ? 2 : 8), 256, 0, st>>>(B.cl[b], nz, jitter_rel, B.root[b], dmax2, dmax); }
            count_launch();
            CB_LAUNCH_CHECK();
        }
        KTimer kt(K_CHOLESKY, st);
        const size_t smem = sizeof(double) * (2 * CH_STAGE + CH_NB * (CH_NB + 1));
        if (!g_chol_v1) {
            auto launch2 = [&](auto kern, size_t sm) -> int {
                CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                kern<<<nl, CH_THREADS, sm, st>>>(B.cl[b], B.root[b], nz, jitter_rel, dmax, fail + (long long)b * nl);
                return 0;
            };
            const size_t sm_a = sizeof(double) * (ch2_stage_doubles(16, 20, 2) + CH_NB * (CH_NB + 1));
            const size_t sm_b = sizeof(double) * (ch2_stage_doubles(8, 12, 3) + CH_NB * (CH_NB + 1));
            int rc2;
            if (nz % 2 == 0) rc2 = g_chol_deep ? launch2(cholesky2_kernel<true, 8, 12, 3>, sm_b) : launch2(cholesky2_kernel<true, 16, 20, 2>, sm_a);
            else rc2 = g_chol_deep ? launch2(cholesky2_kernel<false, 8, 12, 3>, sm_b) : launch2(cholesky2_kernel<false, 16, 20, 2>, sm_a);
            if (rc2) return rc2;
        } else if (nz % 2 == 0) {
            CB_CUDA(cudaFuncSetAttribute(cholesky_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cholesky_kernel<true><<<nl, CH_THREADS, smem, st>>>(B.root[b], nz, fail + (long long)b * nl);
        } else {
            CB_CUDA(cudaFuncSetAttribute(cholesky_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cholesky_kernel<false><<<nl, CH_THREADS, smem, st>>>(B.root[b], nz, fail + (long long)b * nl);
        }
        count_launch();
        CB_LAUNCH_CHECK();
    }
    root_flags_kernel<<<1, 32, 0, st>>>(fail, nb, nl, nz, jitter_rel, dmax, used_eigh, num_pos, fail_list, nfail_d, zero_block_scale);
    count_launch();
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMemsetAsync(eig_max, 0, 8LL * nl, st));
    const int jthreads = nz >= 512 ? 1024 : (nz >= 128 ? 512 : 256);
    const bool lowrank = nz > g_jacobi_max_nz;
    auto wave = [&](int f0, int nw, const int* guard) -> int {
        WaveCtx w;
        w.fail_list = fail_list + f0;
        w.nfail_ptr = guard;
        w.f0 = f0;
        w.nl = nl;
        w.nz = nz;
        w.jitter_rel = jitter_rel;
        w.clip_rel = clip_rel;
        w.dmax = dmax;
        w.eig_max = eig_max;
        w.slots = slot_mem;
        w.evals = evals;
        w.rank = rank;
        w.lr_rank = lr_rank;
        w.status = status;
        w.sweeps = sweeps;
        const int nslots = nw * nb;
        KTimer kt(K_EIGH, st);
        if (lowrank) {
            const size_t smem = sizeof(double) * 2 * (size_t)nz;
            CB_CUDA(cudaFuncSetAttribute(lr_pchol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
            lr_pchol_kernel<<<nslots, nz >= 512 ? 1024 : 512, smem, st>>>(B, w);
            {
                const size_t sm_j = sizeof(double) * (2 * LJ_RMAX * LJ_RMAX + LJ_RMAX * LJ_XC);
                CB_CUDA(cudaFuncSetAttribute(lr_jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_j));
                lr_jacobi_kernel<<<nslots, 1024, sm_j, st>>>(B, w, 60, g_lr_precond ? 1 : 0);
            }
            lr_decide_kernel<<<ceil_div(nslots, 128), 128, 0, st>>>(B, w, nslots);
            count_launch(3);
        } else {
            fill_int_kernel<<<ceil_div(nslots, 256), 256, 0, st>>>(status, nslots, 1);
            count_launch();
        }
        CB_LAUNCH_CHECK();
        jacobi_init_kernel<<<nslots, 256, 0, st>>>(B, w);
        jacobi_kernel<<<nslots, jthreads, 0, st>>>(B, w, 60);
        jacobi_evals_kernel<<<nslots, 256, 0, st>>>(B, w);
        count_launch(3);
        CB_LAUNCH_CHECK();
        if (lowrank) {
            const size_t smem = 12 * (size_t)nz + 16;
            CB_CUDA(cudaFuncSetAttribute(lr_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
            lr_finish_kernel<<<nslots, 256, smem, st>>>(B, w, num_pos, used_eigh);
            count_launch();
        }
        jacobi_finish_kernel<<<nslots, 256, 0, st>>>(B, w, num_pos, used_eigh, nullptr);
        count_launch();
        if (zero_block_scale) {
            zero_scale_kernel<<<ceil_div(nw, 128), 128, 0, st>>>(w, nw, zero_block_scale);
            count_launch();
        }
        CB_LAUNCH_CHECK();
        return 0;
    };
    // The eigen fallback is launched for every l in waves of `groups` slot groups; the CTAs of a wave beyond the
    // device-side failure count exit at once, so the host never waits for the count (no synchronisation in a step),
    // whatever the number of failures.  With room for every matrix (groups >= nl) that is a single wave.
    CB_REQUIRE(groups >= 1, 4, "root_batched: no workspace for the eigen fallback");
    for (int f0 = 0; f0 < nl; f0 += (int)std::min<long long>(groups, nl)) {
        const int nw = (int)std::min<long long>(groups, nl - f0);
        if (int rc = wave(f0, nw, nfail_d)) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------- plain eigh
__global__ void iota_kernel(int* p, int n, int base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = base + i;
}

extern "C" long long cora_b200_eigh_workspace_bytes(int nl, int nz) {
    return 4LL * nl + 16LL * nl + 12LL * nl * nz + 8LL * nl + 10 * 256 + 16LL * nz * nz * (long long)nl;
}

extern "C" int cora_b200_eigh_batched(const double* a, int nl, int nz, double* evecs, double* evals, void* workspace,
                                      long long ws_bytes, void* stream) {
    CB_REQUIRE(a && evecs && evals && workspace && nl >= 1 && nz >= 1, 1, "eigh_batched: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    auto take = [&](long long bytes) { char* p = ws; ws = (char*)(((uintptr_t)(ws + bytes) + 255) & ~(uintptr_t)255); return p; };
    int* list = (int*)take(4LL * nl);
    double* dzero = (double*)take(8LL * nl);
    double* eig_max = (double*)take(8LL * nl);
    double* ev_ws = (double*)take(8LL * nl * nz);
    int* rank = (int*)take(4LL * nl * nz);
    int* sweeps = (int*)take(4LL * nl);
    int* status = (int*)take(4LL * nl);
    double* GV = (double*)ws;
    const long long slots = ((char*)workspace + ws_bytes - ws) / (16LL * nz * nz);
    CB_REQUIRE(slots >= 1, 4, "eigh_batched: workspace too small (%lld B)", ws_bytes);
    iota_kernel<<<ceil_div(nl, 256), 256, 0, st>>>(list, nl, 0);
    fill_int_kernel<<<ceil_div(nl, 256), 256, 0, st>>>(status, nl, 1);
    CB_CUDA(cudaMemsetAsync(dzero, 0, 8LL * nl, st));
    CB_CUDA(cudaMemsetAsync(eig_max, 0, 8LL * nl, st));
    count_launch(2);
    CB_LAUNCH_CHECK();
    const int threads = nz >= 512 ? 1024 : (nz >= 128 ? 512 : 256);
    RootBlocks B;
    B.nb = 1;
    B.cl[0] = a;
    B.root[0] = evecs;
    KTimer kt(K_EIGH, st);
    for (int f0 = 0; f0 < nl; f0 += (int)slots) {
        const int nw = (int)std::min<long long>(slots, nl - f0);
        WaveCtx w;
        w.fail_list = list + f0;
        w.nfail_ptr = nullptr;
        w.f0 = f0;
        w.nl = nl;
        w.nz = nz;
        w.jitter_rel = 0.0;
        w.clip_rel = 0.0;
        w.dmax = dzero;
        w.eig_max = eig_max;
        w.slots = GV;
        w.evals = ev_ws;
        w.rank = rank;
        w.lr_rank = nullptr;
        w.status = status;
        w.sweeps = sweeps;
        jacobi_init_kernel<<<nw, 256, 0, st>>>(B, w);
        jacobi_kernel<<<nw, threads, 0, st>>>(B, w, 60);
        jacobi_evals_kernel<<<nw, 256, 0, st>>>(B, w);
        jacobi_finish_kernel<<<nw, 256, 0, st>>>(B, w, nullptr, nullptr, evals);
        count_launch(4);
        CB_LAUNCH_CHECK();
    }
    return 0;
}
