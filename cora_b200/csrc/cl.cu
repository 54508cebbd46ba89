// C_l(nu, nu') table fill for the two spectra on the hot path (sm_100a).
//
//   SCK foreground : closed form, separable in l and (nu, nu')      (gaussianfg.py:107-130)
//   21cm           : flat-sky DCT tables + bilinear lookup           (corr.py:891-982)
// both Romberg-averaged over each frequency channel as skysim.clarray does
// (cora/core/skysim.py:41-67): Cbar_l[i,j] = sum_ab w_a w_b C_l(s_ia, s_jb), sum_a w_a = 1.
#include "common.cuh"
#include "cora_b200.h"

#include <vector>

namespace cb {

constexpr int MAXZINT = 33;   // 2^5 + 1

// ------------------------------------------------------------------------------ SCK
__global__ void sck_bbar_kernel(double alpha, double nu_ref, double zeta, const double* __restrict__ nu,
                                const double* __restrict__ w, int nz, int zint, double* __restrict__ bbar) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= nz) return;
    double acc = 0.0;
    for (int a = 0; a < zint; a++) {
        const double n1 = nu[i * zint + a];
        const double v1 = pow(n1 / nu_ref, -2.0 * alpha);
        double inner = 0.0;
        for (int b = 0; b < zint; b++) {
            const double n2 = nu[j * zint + b];
            const double v2 = pow(n2 / nu_ref, -2.0 * alpha);
            const double lr = log(n1 / n2) / zeta;
            inner += w[b] * (sqrt(v1 * v2) * exp(-0.5 * lr * lr));
        }
        acc += w[a] * inner;
    }
    bbar[(long long)i * nz + j] = acc;
}

__global__ void sck_scale_kernel(double A, double beta, double l_ref, const double* __restrict__ bbar, int l0, int l_step,
                                 int nl, long long nz2, double* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int li = blockIdx.y;
    if (e >= nz2) return;
    const int l = l0 + li * l_step;
    const double al = (l == 0) ? 0.0 : A * pow((double)l / l_ref, -beta);
    out[(long long)li * nz2 + e] = al * bbar[e];
}

// ---------------------------------------------------------------------- 21cm tables
constexpr int NKPERP = 500;
constexpr int NKPAR = 32768;
constexpr double KPERP_MIN = 1e-4, KPERP_MAX = 40.0, KPAR_MAX = 20.0;

// device table layout: planar, tab[t][y][x] (t = dd, dv, vv; y = r_par index; x = k_perp index):
// a fixed-y row of one table is 500 contiguous doubles, which is what the fill kernel streams.
__device__ __forceinline__ long long tab_idx(int t, int y, int x) { return ((long long)t * NKPAR + y) * NKPERP + x; }

// natural cubic spline in (ln k, ln P), linear extrapolation outside (cubicspline.pyx:124-175)
__device__ double spline_eval(const double* __restrict__ xs, const double* __restrict__ ys,
                              const double* __restrict__ y2, int n, double x) {
    if (x < xs[0]) {
        const double h = xs[1] - xs[0];
        const double a = (ys[1] - ys[0]) / h;
        return (a - h * y2[1] / 6) * (x - xs[0]) + ys[0];
    }
    if (x >= xs[n - 1]) {
        const int kh = n - 1;
        const double h = xs[kh] - xs[kh - 1];
        const double a = (ys[kh] - ys[kh - 1]) / h;
        return (a + h * y2[kh - 1] / 6) * (x - xs[kh]) + ys[kh];
    }
    int kl = 0, kh = n;
    while (kh - kl > 1) {
        const int kn = (kh + kl) / 2;
        if (xs[kn] > x) kh = kn; else kl = kn;
    }
    const double h = xs[kh] - xs[kl];
    const double a = (xs[kh] - x) / h;
    const double b = (x - xs[kl]) / h;
    const double c = (a * a * a - a) * h * h / 6;
    const double d = (b * b * b - b) * h * h / 6;
    return a * ys[kl] + b * ys[kh] + c * y2[kl] + d * y2[kh];
}

// P(k_perp, k_par) {1, mu^2, mu^4}: pk[(x*3 + t) * NKPAR + n]   (corr.py:915-936, corr21cm.py:25-29)
__global__ void ps21_pk_kernel(const double* __restrict__ lnk, const double* __restrict__ lnp,
                               const double* __restrict__ y2, int nknot, double kstar, double* __restrict__ pk) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = blockIdx.y;
    if (n >= NKPAR) return;
    const double lstart = log10(KPERP_MIN), lstop = log10(KPERP_MAX);
    const double lstep = (lstop - lstart) / (double)(NKPERP - 1);
    const double lk = (x == NKPERP - 1) ? lstop : (double)x * lstep + lstart;
    const double kperp = pow(10.0, lk);
    const double kstep = KPAR_MAX / (double)(NKPAR - 1);
    const double kpar = (n == NKPAR - 1) ? KPAR_MAX : (double)n * kstep;
    const double k = sqrt(kpar * kpar + kperp * kperp);
    const double mu2 = kpar * kpar / (k * k);
    const double dd = exp(-0.5 * k * k / (kstar * kstar)) * exp(spline_eval(lnk, lnp, y2, nknot, log(k)));
    pk[((long long)x * 3 + 0) * NKPAR + n] = dd;
    pk[((long long)x * 3 + 1) * NKPAR + n] = dd * mu2;
    pk[((long long)x * 3 + 2) * NKPAR + n] = dd * mu2 * mu2;
}

__global__ void ps21_costab_kernel(double* ctab) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < 2 * (NKPAR - 1)) ctab[j] = cospi((double)j / (double)(NKPAR - 1));
}

// DCT-I along k_par by direct summation with an exact-argument cosine table:
//   X_k = x_0 + (-1)^k x_{N-1} + 2 sum_{n=1}^{N-2} x_n cos(pi n k/(N-1)),   times KPAR_MAX/(2N)
// (scipy.fftpack.dct type 1, corr.py:938-942).  One thread per output k, DCT_ROWS rows
// (x, table) at a time from shared memory; two-level summation keeps the rounding error
// at the 1e-15 level.  Output is planar: tab[t][k][x].
constexpr int DCT_ROWS = 12;   // 4 k_perp values x 3 tables
constexpr int DCT_CHUNK = 256;
__global__ void __launch_bounds__(256) ps21_dct_kernel(const double* __restrict__ pk, const double* __restrict__ ctab,
                                                       double* __restrict__ tab) {
    __shared__ double xs[DCT_CHUNK][DCT_ROWS];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;   // output index (y)
    const int row0 = blockIdx.y * DCT_ROWS;                // first (x*3+t) row
    const int PER = 2 * (NKPAR - 1);
    double tot[DCT_ROWS];
#pragma unroll
    for (int r = 0; r < DCT_ROWS; r++) tot[r] = 0.0;
    for (int n0 = 0; n0 < NKPAR; n0 += DCT_CHUNK) {
        __syncthreads();
        for (int e = threadIdx.x; e < DCT_CHUNK * DCT_ROWS; e += blockDim.x) {
            const int r = e / DCT_CHUNK, nn = e % DCT_CHUNK;
            xs[nn][r] = pk[(long long)(row0 + r) * NKPAR + n0 + nn];
        }
        __syncthreads();
        double part[DCT_ROWS];
#pragma unroll
        for (int r = 0; r < DCT_ROWS; r++) part[r] = 0.0;
        int idx = (int)(((long long)n0 * k) % PER);
        const int step = k % PER;
        for (int nn = 0; nn < DCT_CHUNK; nn++) {
            const int n = n0 + nn;
            const double c = ctab[idx];
            const double wgt = (n == 0 || n == NKPAR - 1) ? c : 2.0 * c;
#pragma unroll
            for (int r = 0; r < DCT_ROWS; r++) part[r] = fma(xs[nn][r], wgt, part[r]);
            idx += step;
            if (idx >= PER) idx -= PER;
        }
#pragma unroll
        for (int r = 0; r < DCT_ROWS; r++) tot[r] += part[r];
    }
    const double norm = KPAR_MAX / (2.0 * (double)NKPAR);
#pragma unroll
    for (int r = 0; r < DCT_ROWS; r++) {
        const int row = row0 + r;
        const int x = row / 3, t = row % 3;
        tab[tab_idx(t, k, x)] = tot[r] * norm;
    }
}

// ---------------------------------------------------------------------- 21cm fill
// One CTA per channel pair (i >= j, nu <-> nu' symmetry).  For a sample pair e = (a, b) the
// table row index y and every l-independent factor are fixed, and only the k_perp coordinate
// x(l, e) = (log10 l - shift_e) * xscale moves with l.  So per batch of 9 sample pairs the CTA
// first builds, for the x band the l block can reach, the y-interpolated and table-combined row
//     R_e[x] = sum_t c_t(e) [ (1 - wy_e) T_t[y0_e][x] + wy_e T_t[y1_e][x] ]
// in shared memory from coalesced row reads, then every thread evaluates its l's with two
// shared-memory reads and one interpolation in x.  Algebraically identical to the reference's
// 4-corner bilinear form (bilinearmap.pyx:49-57); it cuts the bytes moved per evaluation from
// 96 (12 table reads) to ~35, which is what bounds this kernel (L1 bandwidth).
struct PairPre {
    double shift;      // log10(xc * KPERP_MIN)
    double wy;         // fractional part of y
    double cdd, cdv, cvv;   // w_a w_b * prefactor * (b1 b2, f1 b2 + f2 b1, f1 f2)
    int y0, y1;
};

constexpr int FILL_LPT = 4;          // l's per thread per l block
constexpr int FILL_WMAX = 384;       // widest x band held in shared memory
constexpr int FILL_EB = 9;           // sample pairs per batch

__device__ __forceinline__ double fill_direct(const double* __restrict__ tab, const PairPre& p, double lx, double xscale) {
    double x = (lx - p.shift) * xscale;
    x = fmin(fmax(x, 0.0), (double)NKPERP - 1e-5);
    const unsigned x0 = (unsigned)x;
    const unsigned x1 = min(x0 + 1u, (unsigned)(NKPERP - 1));
    const double wx = x - (double)x0;
    // reference weights (bilinearmap.pyx:49-57): wa=(x1-x)(y1-y) [x0,y0], wb=(x1-x)(y-y0) [x0,y1],
    // wc=(x-x0)(y1-y) [x1,y0], wd=(x-x0)(y-y0) [x1,y1]
    const double wa = (1.0 - wx) * (1.0 - p.wy), wb = (1.0 - wx) * p.wy;
    const double wc = wx * (1.0 - p.wy), wd = wx * p.wy;
    double v[3];
#pragma unroll
    for (int t = 0; t < 3; t++)
        v[t] = wa * tab[tab_idx(t, p.y0, x0)] + wb * tab[tab_idx(t, p.y1, x0)] + wc * tab[tab_idx(t, p.y0, x1)] +
               wd * tab[tab_idx(t, p.y1, x1)];
    return p.cdd * v[0] + p.cdv * v[1] + p.cvv * v[2];
}

__global__ void __launch_bounds__(256) cl21_fill_kernel(const double* __restrict__ tab, const double* __restrict__ chi,
                                                        const double* __restrict__ bb, const double* __restrict__ ff,
                                                        const double* __restrict__ pf, const double* __restrict__ DD,
                                                        const double* __restrict__ w, int l0, int l_step, int nl, int nz,
                                                        int zint, double* __restrict__ out, long long pair0, int tile_step,
                                                        double* const* __restrict__ out_ptrs,
                                                        const int* __restrict__ l_owner, const int* __restrict__ l_row,
                                                        const long long* __restrict__ tile_start, int lower_only) {
    extern __shared__ __align__(16) unsigned char smraw[];
    PairPre* pre = (PairPre*)smraw;                                     // [zint*zint]
    double* R = (double*)(smraw + sizeof(PairPre) * zint * zint);        // [FILL_EB][FILL_WMAX]
    __shared__ double s_min, s_max;
    __shared__ double s_red[256];
    // decode the channel pair.  Pairs are enumerated diagonal by diagonal (d = i - j, then j): CTAs that
    // run together then share almost the same band of table rows (y ~ |chi_i - chi_j|), which keeps
    // the band in L2 instead of re-reading it from HBM.  first(d) = d nz - d (d - 1) / 2.
    int i, j;
    if (tile_start) {
        // tile enumeration of the row-weight kernel (4 CTAs per tile (i; j0 .. j0+3), pair0 = first tile)
        const long long tidx = pair0 + (long long)tile_step * (blockIdx.x >> 2);
        int lo = 0, hi = nz - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tile_start[mid] <= tidx) lo = mid; else hi = mid - 1;
        }
        j = 4 * (int)(tidx - tile_start[lo]) + (blockIdx.x & 3);
        i = 4 * (int)(tidx - tile_start[lo]) + lo;
        if (j > i) return;
    } else {
        const long long pidx = pair0 + blockIdx.x;
        int d = (int)(((2.0 * nz + 1.0) - sqrt((2.0 * nz + 1.0) * (2.0 * nz + 1.0) - 8.0 * (double)pidx)) * 0.5);
        d = max(0, min(nz - 1, d));
        while (d + 1 < nz && (long long)(d + 1) * nz - (long long)(d + 1) * d / 2 <= pidx) d++;
        while (d > 0 && (long long)d * nz - (long long)d * (d - 1) / 2 > pidx) d--;
        j = (int)(pidx - ((long long)d * nz - (long long)d * (d - 1) / 2));
        i = j + d;
    }
    const int npair = zint * zint;
    const int tid = threadIdx.x;
    const double PI = 3.14159265358979323846;
    for (int e = tid; e < npair; e += blockDim.x) {
        const int a = e / zint, b = e % zint;
        const int s1 = i * zint + a, s2 = j * zint + b;
        const double x1 = chi[s1], x2 = chi[s2];
        const double xc = 0.5 * (x1 + x2);
        const double rpar = fabs(x2 - x1);
        double y = rpar / (PI / KPAR_MAX);
        y = fmin(fmax(y, 0.0), (double)NKPAR - 1e-5);
        const unsigned y0 = (unsigned)y;
        PairPre p;
        p.y0 = (int)y0;
        p.y1 = (int)min(y0 + 1u, (unsigned)(NKPAR - 1));
        p.wy = y - (double)y0;
        p.shift = log10(xc * KPERP_MIN);
        const double pref = w[a] * w[b] * (DD[s1] * DD[s2] * pf[s1] * pf[s2] / (xc * xc * PI));
        p.cdd = pref * (bb[s1] * bb[s2]);
        p.cdv = pref * (ff[s1] * bb[s2] + ff[s2] * bb[s1]);
        p.cvv = pref * (ff[s1] * ff[s2]);
        pre[e] = p;
    }
    __syncthreads();
    if (tid == 0) {
        double mn = pre[0].shift, mx = pre[0].shift;
        for (int e = 1; e < npair; e++) { mn = fmin(mn, pre[e].shift); mx = fmax(mx, pre[e].shift); }
        s_min = mn; s_max = mx;
    }
    __syncthreads();
    const double smin = s_min, smax = s_max;
    const double xscale = (double)(NKPERP - 1) / log10(KPERP_MAX / KPERP_MIN);
    const double xtopclip = (double)NKPERP - 1e-5;
    const long long nz2 = (long long)nz * nz;
    const long long oij = (long long)i * nz + j, oji = (long long)j * nz + i;

    for (int lb = 0; lb < nl; lb += 256 * FILL_LPT) {
        const int lend = min(nl, lb + 256 * FILL_LPT);
        // x band reachable by l >= 1 of this block (one entry of margin either side)
        const int lfirst = max(1, l0 + lb * l_step), llast = max(1, l0 + (lend - 1) * l_step);
        const double xa = fmin(fmax((log10((double)lfirst) - smax) * xscale, 0.0), xtopclip);
        const double xb = fmin(fmax((log10((double)llast) - smin) * xscale, 0.0), xtopclip);
        const int xbase = max(0, (int)xa - 1);
        const int xtop = min(NKPERP - 1, (int)xb + 2);
        const int W = xtop - xbase + 1;
        const bool banded = (W <= FILL_WMAX);
        // no l >= 1 of this block is clipped, and none reaches the table's last column
        const double xa_raw = (log10((double)lfirst) - smax) * xscale, xb_raw = (log10((double)llast) - smin) * xscale;
        const bool noclip = (xa_raw >= 0.0) && (xb_raw < (double)(NKPERP - 1) - 1e-3);

        double lx[FILL_LPT], acc[FILL_LPT];
        int lv[FILL_LPT];
#pragma unroll
        for (int q = 0; q < FILL_LPT; q++) {
            const int li = lb + tid + q * 256;
            lv[q] = (li < lend) ? l0 + li * l_step : -1;
            lx[q] = log10(lv[q] <= 0 ? 1e-10 : (double)lv[q]);
            acc[q] = 0.0;
        }
        if (banded) {
            for (int e0 = 0; e0 < npair; e0 += FILL_EB) {
                const int ne = min(FILL_EB, npair - e0);
                __syncthreads();   // previous batch fully consumed
                // band rows R[el][x] = y-interpolated, coefficient-weighted table rows (L2 latency bound);
                // items (sample pair, x) dealt evenly over the CTA: W is rarely a multiple of 256, and a
                // per-x deal would leave most threads idle in its last pass
                const int nitems = W * ne;
#pragma unroll 4
                for (int w = tid; w < nitems; w += 256) {
                    const int el = w / W, xi = w - el * W;
                    const int x = xbase + xi;
                    const PairPre p = pre[e0 + el];
                    const double u0 = 1.0 - p.wy, u1 = p.wy;
                    const double dd = u0 * __ldg(tab + tab_idx(0, p.y0, x)) + u1 * __ldg(tab + tab_idx(0, p.y1, x));
                    const double dv = u0 * __ldg(tab + tab_idx(1, p.y0, x)) + u1 * __ldg(tab + tab_idx(1, p.y1, x));
                    const double vv = u0 * __ldg(tab + tab_idx(2, p.y0, x)) + u1 * __ldg(tab + tab_idx(2, p.y1, x));
                    R[el * FILL_WMAX + xi] = p.cdd * dd + p.cdv * dv + p.cvv * vv;
                }
                __syncthreads();
                // x-interpolation of the band rows.  An FP64 fmin/fmax pair costs ~16 instructions (NaN-aware
                // select sequences) -- a third of this loop -- so the clip of bilinearmap.pyx:44-45 is only
                // executed when some l of this block can actually reach it (CTA-uniform test on the band ends;
                // same arithmetic and bit-identical results for in-range x).
                if (noclip) {
#pragma unroll
                    for (int q = 0; q < FILL_LPT; q++) {
                        if (lv[q] < 1) continue;
                        double a = acc[q];
                        for (int el = 0; el < ne; el++) {
                            const double x = (lx[q] - pre[e0 + el].shift) * xscale;
                            const int x0 = (int)x;
                            const double wx = x - (double)x0;
                            const double* Rr = R + el * FILL_WMAX + (x0 - xbase);
                            a += (1.0 - wx) * Rr[0] + wx * Rr[1];
                        }
                        acc[q] = a;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < FILL_LPT; q++) {
                        if (lv[q] < 1) continue;
                        double a = acc[q];
                        for (int el = 0; el < ne; el++) {
                            double x = (lx[q] - pre[e0 + el].shift) * xscale;
                            x = fmin(fmax(x, 0.0), xtopclip);
                            const int x0 = (int)x;
                            const double wx = x - (double)x0;
                            const double* Rr = R + el * FILL_WMAX + (x0 - xbase);
                            const int up = (x0 < NKPERP - 1) ? 1 : 0;
                            a += (1.0 - wx) * Rr[0] + wx * Rr[up];
                        }
                        acc[q] = a;
                    }
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < FILL_LPT; q++) {
                if (lv[q] < 1) continue;
                for (int e = 0; e < npair; e++) acc[q] += fill_direct(tab, pre[e], lx[q], xscale);
            }
        }
        // l = 0 (replaced by 1e-10, corr.py:957): x clips to 0, outside the band -> direct, spread over threads
        if (l0 == 0 && lb == 0) {
            double v = 0.0;
            for (int e = tid; e < npair; e += 256) v += fill_direct(tab, pre[e], log10(1e-10), xscale);
            s_red[tid] = v;
            __syncthreads();
            if (tid == 0) {
                double tsum = 0.0;
                for (int k = 0; k < min(npair, 256); k++) tsum += s_red[k];
                acc[0] = tsum;     // thread 0, q = 0 holds li = 0 <-> l = 0
            }
        }
#pragma unroll
        for (int q = 0; q < FILL_LPT; q++) {
            const int li = lb + tid + q * 256;
            if (li >= lend) continue;
            double* o = out + (long long)li * nz2;
            if (out_ptrs) {   // pair-sharded fill: row l lives on the GPU that owns l (peer store over NVLink)
                const int l = l0 + li * l_step;
                o = out_ptrs[l_owner[l]] + (long long)l_row[l] * nz2;
            }
            o[oij] = acc[q];
            if (i != j && !lower_only) o[oji] = acc[q];
        }
    }
}


// ---------------------------------------------------------------------- 21cm fill, row-weight form
// The sum over the zint^2 sample pairs e of a channel pair is linear in the table entries:
//     R_e[x] = sum_t c_t(e) [ (1 - wy_e) T_t[y0_e][x] + wy_e T_t[y1_e][x] ]        (y-interpolated, combined row)
//     Cbar_l = sum_e interp_x(R_e; x_e(l)),   x_e(l) = (log10 l - shift_e) xscale = X_g(l) - delta_e,
// where the samples are split into G groups of neighbouring shift, X_g is taken at the mid shift of group g and
// |delta_e| <= Delta_g/2 << 1 (Delta ~ 0.12 / G at 256 channels, 0.03 / G at 1024: the sample points of one channel
// pair see almost the same k_perp).  Whenever every x_e(l) of a group lies in the cell k = floor(X_g) of X_g itself,
//     sum_{e in g} = S_g[k] + (X_g - k)(S_g[k+1] - S_g[k]) - (V_g[k+1] - V_g[k]),
//     S_g[x] = sum_{e in g} R_e[x],   V_g[x] = sum_{e in g} delta_e R_e[x],
// and S_g, V_g are weighted sums over the DISTINCT table rows of the pair's y window,
//     S_g[x] = sum_t sum_y A_gt[y] T_t[y][x],   A_gt[y] = sum_{e in g} c_t(e) [(1 - wy_e) 1(y = y0_e) + wy_e 1(y = y1_e)]
// (B_gt, V_g likewise with delta_e): ~30 rows x 3 tables per x at 1024 channels instead of 81 x 6 table reads, and G
// x-interpolations per l instead of 81.  The l whose X_g sits within Delta_g/2 of a cell boundary ("slow", a
// fraction Delta_g of them) have some x_e in the neighbouring cell; for those e the piecewise-linear interpolant
// differs from the linear extension of cell k by |x_e - b| D2_e[b] (b the boundary crossed, D2_e the second
// difference of R_e at b), which a flat pass over the (slow l, sample) candidates adds exactly.  Everything is
// summed in a fixed order: the table is deterministic.  It agrees with the per-sample-pair evaluation (the
// reference's order, `cl21_fill_kernel` above) to a few ulp of the row's largest entry (tests/test_gpu_spectra.py).
//
// One CTA walks a 1 x 4 tile of channel pairs (i; j0 .. j0+3): its four 8-byte stores per l land in one 32-byte
// sector within microseconds and merge in L2 (the per-pair kernel's lone 8-byte stores cost a DRAM
// read-modify-write each: 100 GB of DRAM traffic for a 12.9 GB table at 1024 channels).  With lower_only the
// mirror element (j, i) is not written: the root stage reads the lower triangle only (LAPACK-style).
constexpr int F3_THREADS = 320;
constexpr int F3_TILE = 4;        // j-adjacent channel pairs per CTA
constexpr int F3_NSLOTS = 1024;   // scratch rows for the parked pair results (>> resident CTAs: 3 per SM)
constexpr int F3_GMAX = 4;        // shift groups per channel pair: 4 for short y windows, 2 for tall ones
constexpr int F3_YSPLIT = 64;     // y windows taller than this use 2 groups (the band pass costs 2 G FMAs per table entry)
constexpr int F3_WMAX = 352;      // widest x band
constexpr int F3_YWMAX = 192;     // tallest y window of one pair (256 channels over 400-800 MHz: up to ~140 rows)
constexpr int F3_NPMAX = 81;      // largest zint^2 on the banded path (zromb <= 3; larger: direct evaluation)
constexpr int F3_LISTMAX = 768;   // (slow l, group) entries listed per pair (more: evaluated directly)
constexpr int F3_WDOUBLES = F3_GMAX * 3 * 96 * 2;   // doubles in the weight array (G x 3 x YW x (A, B); G YW <= 4 x 96 = 2 x 192)

__device__ __forceinline__ double fill_direct_all(const double* __restrict__ tab, const PairPre* pre, int npair, double lx,
                                                  double xscale) {
    double a = 0.0;
    for (int e = 0; e < npair; e++) a += fill_direct(tab, pre[e], lx, xscale);
    return a;
}

__global__ void log10_table_kernel(int l0, int l_step, int nl, double* __restrict__ lx) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nl) return;
    const int l = l0 + li * l_step;
    lx[li] = log10(l <= 0 ? 1e-10 : (double)l);
}

struct F3Smem {
    double W[F3_WDOUBLES];                  // (A, B) row weights [g][t][r]; after the band pass: the slow candidates' corrections
    double2 SV[F3_GMAX][F3_WMAX + 2];       // (S, V) band sums
    PairPre pre[F3_NPMAX];
    double delta[F3_NPMAX];
    double gsref[F3_GMAX], gdpos[F3_GMAX], gdneg[F3_GMAX];
    int slow_li[F3_LISTMAX];
    int queue[F3_WDOUBLES];                 // crossing candidates of the current chunk
    int rank[F3_NPMAX];
    short member[F3_GMAX][F3_NPMAX];        // sample index of the eg-th member (ascending shift) of group g (-1: none)
    unsigned char grp[F3_NPMAX];
    unsigned char slow_g[F3_LISTMAX], slow_n[F3_LISTMAX];   // group of the entry; number of entries of its l (first entry only, else 0)
    double smin, smax;
    int ymin, ymax, nslow, nqueue;
};

struct F3Ctx {
    const double* tab;
    const double* lxtab;
    double* out;
    double* const* out_ptrs;
    const int* l_owner;
    const int* l_row;
    long long nz2;
    double xscale;
    int l0, l_step, nl, li_first, lfirst, llast, npair;
    // The tile's pairs (i, j0 .. j0 + pp_last) are evaluated one after the other, but their results leave the CTA
    // together: pairs before the last park their value of every l in a per-CTA scratch row (global memory, L2
    // resident), the last pair's store picks them up and writes ONE run per l -- a single 32-byte vector store for a
    // full tile (SASS STG.256).  Over NVLink (tile-sharded fill: row l lives on the GPU that owns l) that is one
    // 32-byte write where four 8-byte writes used to go out, each with its own packet header.
    double* scratch;          // [F3_TILE - 1][nl]
    int pp, pp_last, full, ic, j0c, nzc;
    long long oi0;            // offset of (i, j0) inside a matrix
    __device__ __forceinline__ void store(int li, double v) const {
        if (pp < pp_last) { scratch[(long long)pp * nl + li] = v; return; }
        double* o = out + (long long)li * nz2;
        if (out_ptrs) {
            const int l = l0 + li * l_step;
            o = out_ptrs[l_owner[l]] + (long long)l_row[l] * nz2;
        }
        double h[F3_TILE];
#pragma unroll
        for (int q = 0; q < F3_TILE - 1; q++) h[q] = (q < pp_last) ? __ldcg(scratch + (long long)q * nl + li) : 0.0;
        double* run = o + oi0;
        if (pp_last == F3_TILE - 1 && (((uintptr_t)run) & 31) == 0) {
            // streaming 256-bit store: the table is written once and read by a later kernel
            asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(run), "d"(h[0]), "d"(h[1]), "d"(h[2]), "d"(v) : "memory");
        } else {
#pragma unroll
            for (int q = 0; q < F3_TILE - 1; q++)
                if (q < pp_last) __stcs(run + q, h[q]);
            __stcs(run + pp_last, v);
        }
        if (full) {             // symmetric table requested: the (j, i) entries, one column of the rows j0 .. j0 + pp_last
#pragma unroll
            for (int q = 0; q < F3_TILE; q++) {
                if (q > pp_last || j0c + q == ic) continue;
                __stcs(o + (long long)(j0c + q) * nzc + ic, q == pp_last ? v : h[q]);
            }
        }
    }
};

// The banded evaluation of one channel pair with G shift groups.  Returns false (uniformly) if the geometry does not
// fit (x clipped inside the band, band or window too large, very wide channels): the caller evaluates directly.
template <int G>
__device__ bool fill3_banded(F3Smem& sm, const F3Ctx& c) {
    const int tid = threadIdx.x;
    const int npair = c.npair;
    const double xscale = c.xscale;
    const int egn = (npair + G - 1) / G;
    // ---- groups of neighbouring shift: group = rank G / npair, members listed in ascending shift
    for (int e = tid; e < G * F3_NPMAX; e += F3_THREADS) sm.member[e / F3_NPMAX][e % F3_NPMAX] = -1;
    __syncthreads();
    for (int e = tid; e < npair; e += F3_THREADS) {
        const int rk = sm.rank[e];
        const int g = rk * G / npair;
        const int first = (g * npair + G - 1) / G;      // smallest rank with rank G / npair == g
        sm.grp[e] = (unsigned char)g;
        sm.member[g][rk - first] = (short)e;
    }
    __syncthreads();
    if (tid < G) {
        // the group's members are sorted: its extreme shifts are those of the first and the last member
        const int first = (tid * npair + G - 1) / G, next = ((tid + 1) * npair + G - 1) / G;
        double mn, mx;
        if (next > first) { mn = sm.pre[sm.member[tid][0]].shift; mx = sm.pre[sm.member[tid][next - first - 1]].shift; }
        else { mn = mx = sm.smin; }            // empty group (npair < G)
        const double sref = 0.5 * (mn + mx);
        sm.gsref[tid] = sref;
        sm.gdpos[tid] = (mx - sref) * xscale;
        sm.gdneg[tid] = (sref - mn) * xscale;
    }
    __syncthreads();
    const int ymin = sm.ymin, YW = sm.ymax - sm.ymin + 1;
    // x band reachable by the l >= 1 of this call (one cell of margin either side + the k + 1 entry); no clip of x
    // anywhere in it (bilinearmap.pyx:44-45 is then the identity), band and window fit, and the cells of one group's
    // samples differ by at most one
    const double xa_raw = (log10((double)c.lfirst) - sm.smax) * xscale, xb_raw = (log10((double)c.llast) - sm.smin) * xscale;
    const int xbase = (int)xa_raw - 1;
    const int W = (int)xb_raw + 3 - xbase;
    double dmaxg = 0.0;
#pragma unroll
    for (int g = 0; g < G; g++) dmaxg = fmax(dmaxg, sm.gdpos[g] + sm.gdneg[g]);
    if (!((xa_raw >= 2.0) && (xb_raw < (double)(NKPERP - 3)) && (W <= F3_WMAX) && (YW * G * 6 <= F3_WDOUBLES) && (dmaxg < 0.45)))
        return false;
    for (int e = tid; e < npair; e += F3_THREADS) sm.delta[e] = (sm.pre[e].shift - sm.gsref[sm.grp[e]]) * xscale;
    __syncthreads();
    // ---- row weights of the y window per group (fixed summation order: the group's members in ascending shift)
    double2* W2 = (double2*)sm.W;                  // [g][t][r], r < YW
    for (int it = tid; it < YW * G; it += F3_THREADS) {
        const int r = it / G, g = it - r * G;
        const int y = ymin + r;
        double a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
        for (int eg = 0; eg < egn; eg++) {
            const int e = sm.member[g][eg];
            if (e < 0) break;
            const int ey0 = sm.pre[e].y0, ey1 = sm.pre[e].y1;
            if (ey0 != y && ey1 != y) continue;
            const PairPre p = sm.pre[e];
            double u = 0.0;
            if (ey0 == y) u += 1.0 - p.wy;
            if (ey1 == y) u += p.wy;
            const double de = sm.delta[e];
            const double cdd = p.cdd * u, cdv = p.cdv * u, cvv = p.cvv * u;
            a0 += cdd; a1 += cdv; a2 += cvv;
            b0 = fma(de, cdd, b0); b1 = fma(de, cdv, b1); b2 = fma(de, cvv, b2);
        }
        W2[(g * 3 + 0) * YW + r] = make_double2(a0, b0);
        W2[(g * 3 + 1) * YW + r] = make_double2(a1, b1);
        W2[(g * 3 + 2) * YW + r] = make_double2(a2, b2);
    }
    __syncthreads();
    // ---- (S_g, V_g)[x] over the band: thread = x, coalesced row reads, 2 G FMAs per table entry
    for (int xi = tid; xi < W; xi += F3_THREADS) {
        double sv[G], vv[G];
#pragma unroll
        for (int g = 0; g < G; g++) sv[g] = vv[g] = 0.0;
        const double* col = c.tab + (long long)ymin * NKPERP + (xbase + xi);
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const double* ct = col + (long long)t * NKPAR * NKPERP;
#pragma unroll 16
            for (int r = 0; r < YW; r++) {
                const double v = __ldg(ct + (long long)r * NKPERP);
#pragma unroll
                for (int g = 0; g < G; g++) {
                    const double2 wg = W2[(g * 3 + t) * YW + r];
                    sv[g] = fma(wg.x, v, sv[g]);
                    vv[g] = fma(wg.y, v, vv[g]);
                }
            }
        }
#pragma unroll
        for (int g = 0; g < G; g++) sm.SV[g][xi] = make_double2(sv[g], vv[g]);
    }
    __syncthreads();
    // ---- G interpolations per l; an l with a group near a cell boundary books one entry per such group
    double rsref[G], rdpos[G], rdneg[G];
#pragma unroll
    for (int g = 0; g < G; g++) { rsref[g] = sm.gsref[g]; rdpos[g] = sm.gdpos[g] + 1e-9; rdneg[g] = 1.0 - sm.gdneg[g] - 1e-9; }
    auto fast_sum = [&](double lx) {
        double acc = 0.0;
#pragma unroll
        for (int g = 0; g < G; g++) {
            const double X = (lx - rsref[g]) * xscale;
            const int k = (int)X;
            const double f = X - (double)k;
            const double2 c0 = sm.SV[g][k - xbase], c1 = sm.SV[g][k - xbase + 1];
            acc += (c0.x + f * (c1.x - c0.x)) - (c1.y - c0.y);
        }
        return acc;
    };
    for (int li = c.li_first + tid; li < c.nl; li += F3_THREADS) {
        const double lx = c.lxtab[li];
        int mask = 0, cnt = 0;
        double acc = 0.0;
#pragma unroll
        for (int g = 0; g < G; g++) {
            const double X = (lx - rsref[g]) * xscale;
            const int k = (int)X;
            const double f = X - (double)k;
            if (f < rdpos[g] || f > rdneg[g]) { mask |= 1 << g; cnt++; }
            const double2 c0 = sm.SV[g][k - xbase], c1 = sm.SV[g][k - xbase + 1];
            acc += (c0.x + f * (c1.x - c0.x)) - (c1.y - c0.y);
        }
        if (mask == 0) { c.store(li, acc); continue; }
        const int slot = atomicAdd(&sm.nslow, cnt);
        if (slot + cnt > F3_LISTMAX) {        // no room (pathological geometry): evaluate directly; void straddled slots
            for (int q = slot; q < F3_LISTMAX; q++) { sm.slow_li[q] = li; sm.slow_g[q] = 0; sm.slow_n[q] = 0; }
            c.store(li, fill_direct_all(c.tab, sm.pre, npair, lx, xscale));
            continue;
        }
        int q = 0;
#pragma unroll
        for (int g = 0; g < G; g++)
            if (mask & (1 << g)) {
                sm.slow_li[slot + q] = li;
                sm.slow_g[slot + q] = (unsigned char)g;
                sm.slow_n[slot + q] = (unsigned char)(q == 0 ? cnt : 0);
                q++;
            }
    }
    __syncthreads();
    // ---- slow candidates (entry, member of its group): the exact correction of a sample in the neighbouring cell.
    // Pass 1, one thread per candidate: is the sample in another cell than its group's X?  (no table access; its
    // correction slot is zeroed, the crossing ones are queued.)  Pass 2, one thread per queued candidate: the second
    // difference from 18 table reads.  Pass 3, one thread per l: its candidates summed in a fixed order.  In chunks
    // whose corrections fit the weight array (idle now).
    const int nslow = min(sm.nslow, F3_LISTMAX);
    double* contrib = sm.W;
    const int chunk = max(1, F3_WDOUBLES / egn);   // entries per chunk
    for (int s0 = 0; s0 < nslow;) {
        int s1 = min(s0 + chunk, nslow);
        while (s1 < nslow && s1 > s0 + 1 && sm.slow_n[s1] == 0) s1--;       // do not split the entries of one l (uniform)
        if (tid == 0) sm.nqueue = 0;
        __syncthreads();
        for (int cc = tid; cc < (s1 - s0) * egn; cc += F3_THREADS) {
            const int sl = cc / egn, eg = cc - sl * egn;
            const int slot = s0 + sl;
            const int g = sm.slow_g[slot];
            const int e = sm.member[g][eg];
            contrib[cc] = 0.0;
            if (e >= 0) {
                const double lx = c.lxtab[sm.slow_li[slot]];
                const int k = (int)((lx - sm.gsref[g]) * xscale);
                const int ke = (int)((lx - sm.pre[e].shift) * xscale);
                if (ke != k) sm.queue[atomicAdd(&sm.nqueue, 1)] = cc;
            }
        }
        __syncthreads();
        const int nq = sm.nqueue;
        for (int qi = tid; qi < nq; qi += F3_THREADS) {
            const int cc = sm.queue[qi];
            const int sl = cc / egn, eg = cc - sl * egn;
            const int slot = s0 + sl;
            const int g = sm.slow_g[slot];
            const PairPre p = sm.pre[sm.member[g][eg]];
            const double lx = c.lxtab[sm.slow_li[slot]];
            const int k = (int)((lx - sm.gsref[g]) * xscale);
            const double xe = (lx - p.shift) * xscale;
            const int b = ((int)xe < k) ? k : k + 1;            // the boundary between cell k and the sample's cell
            const double* c0 = c.tab + (long long)p.y0 * NKPERP + b;
            const double* c1 = c.tab + (long long)p.y1 * NKPERP + b;
            double d2 = 0.0;
#pragma unroll
            for (int t = 0; t < 3; t++) {
                const long long off = (long long)t * NKPAR * NKPERP;
                const double r0 = __ldg(c0 + off + 1) - 2.0 * __ldg(c0 + off) + __ldg(c0 + off - 1);
                const double r1 = __ldg(c1 + off + 1) - 2.0 * __ldg(c1 + off) + __ldg(c1 + off - 1);
                const double ct = (t == 0) ? p.cdd : ((t == 1) ? p.cdv : p.cvv);
                d2 = fma(ct, (1.0 - p.wy) * r0 + p.wy * r1, d2);
            }
            contrib[cc] = fabs(xe - (double)b) * d2;
        }
        __syncthreads();
        for (int slot = s0 + tid; slot < s1; slot += F3_THREADS) {
            const int n = sm.slow_n[slot];
            if (n == 0) continue;                                // not the first entry of its l
            double corr = 0.0;
            const double* cp = contrib + (slot - s0) * egn;
            const int nn = min(n, s1 - slot) * egn;
            for (int q = 0; q < nn; q++) corr += cp[q];
            const int li = sm.slow_li[slot];
            c.store(li, fast_sum(c.lxtab[li]) + corr);
        }
        __syncthreads();
        s0 = s1;
    }
    return true;
}

__global__ void __launch_bounds__(F3_THREADS, 3) cl21_fill3_kernel(const double* __restrict__ tab, const double* __restrict__ chi,
                                                                   const double* __restrict__ bb, const double* __restrict__ ff,
                                                                   const double* __restrict__ pf, const double* __restrict__ DD,
                                                                   const double* __restrict__ w, const double* __restrict__ lxtab,
                                                                   const long long* __restrict__ tile_start,
                                                                   int l0, int l_step, int nl, int nz, int zint,
                                                                   double* __restrict__ out, long long tile0, int tile_step, int lower_only,
                                                                   double* const* __restrict__ out_ptrs,
                                                                   const int* __restrict__ l_owner, const int* __restrict__ l_row,
                                                                   double* __restrict__ scratch_all, int* __restrict__ slot_busy) {
    extern __shared__ __align__(16) unsigned char smraw[];
    F3Smem& sm = *(F3Smem*)smraw;
    const int npair = zint * zint;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tile (i; j0 .. j0+3): tiles are enumerated by dt = i - j0 first (tiles that run together share the band of
    // table rows, y ~ |chi_i - chi_j|), then by j0 / 4.  tile_start[dt] = index of the first tile of offset dt.
    const long long tidx = tile0 + (long long)tile_step * blockIdx.x;
    int lo = 0, hi = nz - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tile_start[mid] <= tidx) lo = mid; else hi = mid - 1;
    }
    const int dt = lo;
    const int j0 = F3_TILE * (int)(tidx - tile_start[dt]);
    const int i = j0 + dt;
    const double PI = 3.14159265358979323846;
    F3Ctx c;
    c.tab = tab; c.lxtab = lxtab; c.out = out; c.out_ptrs = out_ptrs; c.l_owner = l_owner; c.l_row = l_row;
    c.nz2 = (long long)nz * nz;
    c.xscale = (double)(NKPERP - 1) / log10(KPERP_MAX / KPERP_MIN);
    c.l0 = l0; c.l_step = l_step; c.nl = nl; c.npair = npair;
    c.li_first = (l0 >= 1) ? 0 : 1;   // first li with l = l0 + li l_step >= 1 (l0 >= 0, l_step >= 1)
    c.lfirst = l0 + c.li_first * l_step;
    c.llast = l0 + (nl - 1) * l_step;
    const bool have_l = (c.llast >= 1) && (c.li_first < nl);
    const double xscale = c.xscale;

    // sample-pair constants of (i, j); all of them fit shared memory up to zint = 9, else they are recomputed on the fly
    auto make_pre = [&](int j, int e) {
        const int a = e / zint, b = e % zint;
        const int s1 = i * zint + a, s2 = j * zint + b;
        const double x1 = chi[s1], x2 = chi[s2];
        const double xc = 0.5 * (x1 + x2);
        const double rpar = fabs(x2 - x1);
        double y = rpar / (PI / KPAR_MAX);
        y = fmin(fmax(y, 0.0), (double)NKPAR - 1e-5);
        const unsigned y0 = (unsigned)y;
        PairPre p;
        p.y0 = (int)y0;
        p.y1 = (int)min(y0 + 1u, (unsigned)(NKPAR - 1));
        p.wy = y - (double)y0;
        p.shift = log10(xc * KPERP_MIN);
        const double pref = w[a] * w[b] * (DD[s1] * DD[s2] * pf[s1] * pf[s2] / (xc * xc * PI));
        p.cdd = pref * (bb[s1] * bb[s2]);
        p.cdv = pref * (ff[s1] * bb[s2] + ff[s2] * bb[s1]);
        p.cvv = pref * (ff[s1] * ff[s2]);
        return p;
    };

    // scratch row of this CTA: claim one of the F3_NSLOTS slots (at most 3 x #SM CTAs are resident, far fewer than slots)
    __shared__ int s_slot;
    if (tid == 0) {
        int sl = (int)(blockIdx.x % F3_NSLOTS);
        while (atomicCAS(slot_busy + sl, 0, 1) != 0) sl = (sl + 1) % F3_NSLOTS;
        s_slot = sl;
    }
    __syncthreads();
    c.scratch = scratch_all + (long long)s_slot * (F3_TILE - 1) * nl;
    c.pp_last = min(F3_TILE - 1, i - j0);
    c.full = lower_only ? 0 : 1;
    c.ic = i; c.j0c = j0; c.nzc = nz;
    c.oi0 = (long long)i * nz + j0;

    for (int pp = 0; pp < F3_TILE; pp++) {
        const int j = j0 + pp;
        if (j > i) break;                      // uniform: partial tile on the diagonal
        c.pp = pp;
        __syncthreads();                       // the previous pair's shared state is no longer read (and its parked values are visible)
        const bool small = (npair <= F3_NPMAX);
        if (small) {
            for (int e = tid; e < npair; e += F3_THREADS) { sm.pre[e] = make_pre(j, e); sm.rank[e] = 0; }
        }
        if (tid == 0) sm.nslow = 0;
        __syncthreads();
        bool done = false;
        if (small && have_l) {
            // rank of every sample by shift (ties by index; three threads count a third of the samples each)
            for (int it = tid; it < 3 * npair; it += F3_THREADS) {
                const int e = it / 3, part = it - 3 * e;
                const int q0 = part * npair / 3, q1 = (part + 1) * npair / 3;
                const double se = sm.pre[e].shift;
                int rk = 0;
                for (int q = q0; q < q1; q++) {
                    const double sq = sm.pre[q].shift;
                    rk += (sq < se) || (sq == se && q < e);
                }
                atomicAdd(&sm.rank[e], rk);
            }
            if (warp == 9) {
                double mn = 1.0e308, mx = -1.0e308;
                int ymn = 0x7fffffff, ymx = -1;
                for (int e = lane; e < npair; e += 32) {
                    mn = fmin(mn, sm.pre[e].shift); mx = fmax(mx, sm.pre[e].shift);
                    ymn = min(ymn, sm.pre[e].y0); ymx = max(ymx, sm.pre[e].y1);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    ymn = min(ymn, __shfl_xor_sync(0xffffffffu, ymn, o)); ymx = max(ymx, __shfl_xor_sync(0xffffffffu, ymx, o));
                }
                if (lane == 0) { sm.smin = mn; sm.smax = mx; sm.ymin = ymn; sm.ymax = ymx; }
            }
            __syncthreads();
            const int YW = sm.ymax - sm.ymin + 1;
            done = (YW > F3_YSPLIT || npair < 4) ? fill3_banded<2>(sm, c) : fill3_banded<4>(sm, c);
        }
        if (!done && have_l) {
            // exotic geometry: the plain per-sample-pair evaluation
            if (small) {
                for (int li = c.li_first + tid; li < nl; li += F3_THREADS) c.store(li, fill_direct_all(tab, sm.pre, npair, lxtab[li], xscale));
            } else {
                for (int li = c.li_first + tid; li < nl; li += F3_THREADS) {
                    double a = 0.0;
                    for (int e = 0; e < npair; e++) a += fill_direct(tab, make_pre(j, e), lxtab[li], xscale);
                    c.store(li, a);
                }
            }
        }
        // l <= 0 (l = 0 replaced by 1e-10, corr.py:957): x clips to 0 -> direct, spread over the threads
        if (c.li_first > 0) {
            double v = 0.0;
            for (int e = tid; e < npair; e += F3_THREADS) v += fill_direct(tab, small ? sm.pre[e] : make_pre(j, e), log10(1e-10), xscale);
            __syncthreads();
            sm.W[tid] = v;
            __syncthreads();
            if (tid == 0) {
                double tsum = 0.0;
                for (int k = 0; k < F3_THREADS; k++) tsum += sm.W[k];
                for (int li = 0; li < min(c.li_first, nl); li++) c.store(li, tsum);
            }
        }
    }
    __syncthreads();
    if (tid == 0) atomicExch(slot_busy + s_slot, 0);
}

// upper triangle <- lower triangle (32 x 32 tiles through shared memory, full-sector writes)
__global__ void cl_symmetrize_kernel(double* __restrict__ cl, int nz) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    double* M = cl + (long long)blockIdx.z * nz * nz;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 256 threads: 8 rows per pass
    for (int r = ty; r < 32; r += 8) {
        const int i = bi * 32 + r, j = bj * 32 + tx;
        tile[r][tx] = (i < nz && j < nz) ? M[(long long)i * nz + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = bj * 32 + r, i = bi * 32 + tx;           // write element (j, i) = tile[i - bi 32][j - bj 32]
        if (i < nz && j < nz && j < i) M[(long long)j * nz + i] = tile[tx][r];
    }
}


// ------------------------------------------------------------- generic helpers
// Romberg average of an already-evaluated block: in[nl][nz][zint][nz][zint] -> out[nl][nz][nz]
// (the two si.romb calls + normalisation of cora/core/skysim.py:62-67 for a generic callable).
__global__ void romberg_reduce_kernel(const double* __restrict__ in, const double* __restrict__ w, int nz, int zint,
                                      double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y, l = blockIdx.z;
    if (j >= nz) return;
    const long long base = (((long long)l * nz + i) * zint) * nz * zint;
    double acc = 0.0;
    for (int a = 0; a < zint; a++) {
        double inner = 0.0;
        for (int b = 0; b < zint; b++) inner += w[b] * in[base + ((long long)a * nz + j) * zint + b];
        acc += w[a] * inner;
    }
    out[((long long)l * nz + i) * nz + j] = acc;
}

// point-wise spectra for arbitrary broadcast shapes (host broadcasts, device evaluates)
__global__ void sck_points_kernel(double A, double beta, double l_ref, double alpha, double nu_ref, double zeta,
                                  const double* __restrict__ l, const double* __restrict__ n1,
                                  const double* __restrict__ n2, long long n, double* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double ll = l[e];
    const double al = (ll == 0.0) ? 0.0 : A * pow(ll / l_ref, -beta);
    const double v1 = pow(n1[e] / nu_ref, -2.0 * alpha), v2 = pow(n2[e] / nu_ref, -2.0 * alpha);
    const double lr = log(n1[e] / n2[e]) / zeta;
    out[e] = al * (sqrt(v1 * v2) * exp(-0.5 * lr * lr));
}

// vec1/vec2: [5][n] rows chi, b, f, pf, D for the two redshift arguments
__global__ void cl21_points_kernel(const double* __restrict__ tab, const double* __restrict__ l,
                                   const double* __restrict__ v1, const double* __restrict__ v2, long long n,
                                   double* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double PI = 3.14159265358979323846;
    const double x1 = v1[e], x2 = v2[e];
    const double b1 = v1[n + e], b2 = v2[n + e], f1 = v1[2 * n + e], f2 = v2[2 * n + e];
    const double pf1 = v1[3 * n + e], pf2 = v2[3 * n + e], D1 = v1[4 * n + e], D2 = v2[4 * n + e];
    const double xc = 0.5 * (x1 + x2), rpar = fabs(x2 - x1);
    double ll = l[e];
    if (ll == 0.0) ll = 1e-10;
    double x = (log10(ll) - log10(xc * KPERP_MIN)) / log10(KPERP_MAX / KPERP_MIN) * (double)(NKPERP - 1);
    double y = rpar / (PI / KPAR_MAX);
    x = fmin(fmax(x, 0.0), (double)NKPERP - 1e-5);
    y = fmin(fmax(y, 0.0), (double)NKPAR - 1e-5);
    const unsigned x0 = (unsigned)x, y0 = (unsigned)y;
    const unsigned xn = min(x0 + 1u, (unsigned)(NKPERP - 1)), yn = min(y0 + 1u, (unsigned)(NKPAR - 1));
    const double wx = x - (double)x0, wy = y - (double)y0;
    const double wa = (1.0 - wx) * (1.0 - wy), wb = (1.0 - wx) * wy, wc = wx * (1.0 - wy), wd = wx * wy;
    double v[3];
#pragma unroll
    for (int t = 0; t < 3; t++)
        v[t] = wa * tab[tab_idx(t, y0, x0)] + wb * tab[tab_idx(t, yn, x0)] + wc * tab[tab_idx(t, y0, xn)] +
               wd * tab[tab_idx(t, yn, xn)];
    const double dd = v[0], dv = v[1], vv = v[2];
    out[e] = (D1 * D2 * pf1 * pf2 / (xc * xc * PI)) * ((b1 * b2) * dd + (f1 * b2 + f2 * b1) * dv + (f1 * f2) * vv);
}

// read back table entries in the reference's [x][y] indexing (tests / cache export)
__global__ void ps21_gather_kernel(const double* __restrict__ tab, const int* __restrict__ xs, const int* __restrict__ ys,
                                   int n, double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    for (int t = 0; t < 3; t++) out[3 * e + t] = tab[tab_idx(t, ys[e], xs[e])];
}

}  // namespace cb

using namespace cb;

static int dev_scratch(int id, size_t bytes, void** out);   // library-owned per-device scratch (defined below)

extern "C" int cora_b200_cl_fill_sck(double A, double beta, double l_ref, double alpha, double nu_ref, double zeta,
                                     const double* nu_samples, const double* w, int l0, int l_step, int nl, int nz,
                                     int zint, double* out_cl, void* stream) {
    CB_REQUIRE(nu_samples && w && out_cl, 1, "cl_fill_sck: null argument");
    CB_REQUIRE(nl >= 1 && nz >= 1 && zint >= 1 && zint <= MAXZINT && l0 >= 0 && l_step >= 1, 1, "cl_fill_sck: bad sizes nl=%d nz=%d zint=%d", nl, nz, zint);
    cudaStream_t st = (cudaStream_t)stream;
    KTimer kt(K_CL_FILL, st);
    double* bbar = nullptr;     // (one fill at a time per device: calls on different streams must not overlap)
    if (int rc = dev_scratch(2, sizeof(double) * (size_t)nz * nz, (void**)&bbar)) return rc;
    sck_bbar_kernel<<<dim3(ceil_div(nz, 128), nz), 128, 0, st>>>(alpha, nu_ref, zeta, nu_samples, w, nz, zint, bbar);
    count_launch();
    CB_LAUNCH_CHECK();
    const long long nz2 = (long long)nz * nz;
    sck_scale_kernel<<<dim3(ceil_div(nz2, 256), nl), 256, 0, st>>>(A, beta, l_ref, bbar, l0, l_step, nl, nz2, out_cl);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" long long cora_b200_ps_table_21cm_bytes(void) { return (long long)NKPERP * NKPAR * 3 * 8; }
extern "C" long long cora_b200_ps_table_21cm_workspace_bytes(void) {
    return (long long)NKPERP * NKPAR * 3 * 8 + (long long)2 * NKPAR * 8 + 3 * 8 * 4096 + 1024;
}

extern "C" int cora_b200_ps_table_21cm(const double* lnk_h, const double* lnp_h, const double* y2_h, int nknot, double kstar,
                                       double* tab, void* workspace, long long ws_bytes, void* stream) {
    CB_REQUIRE(lnk_h && lnp_h && y2_h && tab && workspace, 1, "ps_table_21cm: null argument");
    CB_REQUIRE(nknot >= 4 && nknot <= 4096, 1, "ps_table_21cm: need 4 <= nknot <= 4096 (got %d)", nknot);
    CB_REQUIRE(ws_bytes >= cora_b200_ps_table_21cm_workspace_bytes(), 4, "ps_table_21cm: workspace too small");
    CB_REQUIRE((NKPERP * 3) % DCT_ROWS == 0 && NKPAR % 256 == 0, 9, "ps_table_21cm: internal tiling");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    KTimer kt(K_TABLE, st);
    double* pk = (double*)ws;
    double* ctab = pk + (long long)NKPERP * NKPAR * 3;
    double* knots = ctab + 2 * NKPAR;
    CB_CUDA(cudaMemcpyAsync(knots, lnk_h, sizeof(double) * nknot, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(knots + 4096, lnp_h, sizeof(double) * nknot, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(knots + 8192, y2_h, sizeof(double) * nknot, cudaMemcpyHostToDevice, st));
    ps21_pk_kernel<<<dim3(NKPAR / 256, NKPERP), 256, 0, st>>>(knots, knots + 4096, knots + 8192, nknot, kstar, pk);
    count_launch();
    CB_LAUNCH_CHECK();
    ps21_costab_kernel<<<ceil_div(2 * (NKPAR - 1), 256), 256, 0, st>>>(ctab);
    count_launch();
    CB_LAUNCH_CHECK();
    ps21_dct_kernel<<<dim3(NKPAR / 256, NKPERP * 3 / DCT_ROWS), 256, 0, st>>>(pk, ctab, tab);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

// CORA_B200_FILL_V1=1: the per-sample-pair kernel of round 1 (A/B reference for the row-weight kernel)
static const bool g_fill_v1 = [] { const char* e = getenv("CORA_B200_FILL_V1"); return e && atoi(e) != 0; }();

// tiles (i; j0 = 4 jt .. j0 + 3), j0 <= i, enumerated by dt = i - j0 then jt: tile_start[dt], dt = 0 .. nz
static long long fill_ntiles(int nz) {
    long long n = 0;
    for (int dt = 0; dt < nz; dt++) n += (nz - 1 - dt) / F3_TILE + 1;
    return n;
}
struct TileTab { int dev, nz; long long* d; };
static std::vector<TileTab> g_tile_tabs;
static int fill_tile_table(int nz, const long long** out) {
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    for (auto& t : g_tile_tabs)
        if (t.dev == dev && t.nz == nz) { *out = t.d; return 0; }
    std::vector<long long> h(nz + 1);
    long long n = 0;
    for (int dt = 0; dt <= nz; dt++) {
        h[dt] = n;
        if (dt < nz) n += (nz - 1 - dt) / F3_TILE + 1;
    }
    TileTab t;
    t.dev = dev; t.nz = nz; t.d = nullptr;
    CB_CUDA(cudaMalloc(&t.d, sizeof(long long) * (nz + 1)));
    CB_CUDA(cudaMemcpy(t.d, h.data(), sizeof(long long) * (nz + 1), cudaMemcpyHostToDevice));
    g_tile_tabs.push_back(t);
    *out = t.d;
    return 0;
}

extern "C" long long cora_b200_cl_fill_21cm_ntiles(int nz) { return nz >= 1 ? fill_ntiles(nz) : 0; }

// Small per-device scratch buffers owned by the library (grow-only, keyed by purpose).  NOT cudaMallocAsync: the
// stream-ordered pool hands its memory back to the driver at every synchronisation (release threshold 0), so each step
// paid a driver allocation -- measured as a 50 .. 850 ms stall of the fill launch one call in three when the device
// memory is nearly full (profiles/e2e_diag.py), the jitter of the end-to-end numbers.
struct DevScratch { int dev, id; size_t bytes; void* p; };
static std::vector<DevScratch> g_dev_scratch;
static int dev_scratch(int id, size_t bytes, void** out) {
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    for (auto& f : g_dev_scratch)
        if (f.dev == dev && f.id == id) {
            if (f.bytes < bytes) {
                CB_CUDA(cudaDeviceSynchronize());
                cudaFree(f.p);
                f.p = nullptr; f.bytes = 0;
                CB_CUDA(cudaMalloc(&f.p, bytes));
                f.bytes = bytes;
            }
            *out = f.p;
            return 0;
        }
    DevScratch f;
    f.dev = dev; f.id = id; f.bytes = bytes; f.p = nullptr;
    CB_CUDA(cudaMalloc(&f.p, bytes));
    g_dev_scratch.push_back(f);
    *out = f.p;
    return 0;
}

// per-device scratch of the fill kernel (parked pair results, F3_NSLOTS rows of (F3_TILE - 1) nl doubles) and the slot
// flags; grows with nl, never shrinks
struct FillScratch { int dev; int nl; double* d; int* busy; };
static std::vector<FillScratch> g_fill_scratch;
static int fill_scratch(int nl, double** d, int** busy) {
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    for (auto& f : g_fill_scratch)
        if (f.dev == dev) {
            if (f.nl < nl) {
                CB_CUDA(cudaDeviceSynchronize());
                cudaFree(f.d);
                CB_CUDA(cudaMalloc(&f.d, sizeof(double) * (size_t)F3_NSLOTS * (F3_TILE - 1) * nl));
                f.nl = nl;
            }
            *d = f.d; *busy = f.busy;
            return 0;
        }
    FillScratch f;
    f.dev = dev; f.nl = nl; f.d = nullptr; f.busy = nullptr;
    CB_CUDA(cudaMalloc(&f.d, sizeof(double) * (size_t)F3_NSLOTS * (F3_TILE - 1) * nl));
    CB_CUDA(cudaMalloc(&f.busy, sizeof(int) * F3_NSLOTS));
    CB_CUDA(cudaMemset(f.busy, 0, sizeof(int) * F3_NSLOTS));
    g_fill_scratch.push_back(f);
    *d = f.d; *busy = f.busy;
    return 0;
}

static int fill21_launch(const double* tab, const double* chi, const double* b, const double* f, const double* pf,
                         const double* D, const double* w, int l0, int l_step, int nl, int nz, int zint, double* out_cl,
                         int lower_only, int variant, long long tile0, long long ntiles, int tile_step, double* const* out_ptrs,
                         const int* l_owner, const int* l_row, cudaStream_t st) {
    KTimer kt(K_CL_FILL, st);
    const long long* tstart = nullptr;
    if (int rc = fill_tile_table(nz, &tstart)) return rc;
    if (g_fill_v1 || variant == 1) {
        size_t smem = sizeof(PairPre) * (size_t)zint * zint + sizeof(double) * FILL_EB * FILL_WMAX;
        CB_CUDA(cudaFuncSetAttribute(cl21_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cl21_fill_kernel<<<(unsigned)(ntiles * F3_TILE), 256, smem, st>>>(tab, chi, b, f, pf, D, w, l0, l_step, nl, nz, zint, out_cl, tile0, tile_step,
                                                                         out_ptrs, l_owner, l_row, tstart, lower_only);
        count_launch();
        CB_LAUNCH_CHECK();
        return 0;
    }
    double* scratch = nullptr;
    int* busy = nullptr;
    if (int rc = fill_scratch(nl, &scratch, &busy)) return rc;
    double* lx = nullptr;
    if (int rc = dev_scratch(1, sizeof(double) * (size_t)nl, (void**)&lx)) return rc;
    log10_table_kernel<<<ceil_div(nl, 256), 256, 0, st>>>(l0, l_step, nl, lx);
    count_launch();
    const size_t smem = sizeof(F3Smem);
    CB_CUDA(cudaFuncSetAttribute(cl21_fill3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cl21_fill3_kernel<<<(unsigned)ntiles, F3_THREADS, smem, st>>>(tab, chi, b, f, pf, D, w, lx, tstart, l0, l_step, nl, nz, zint, out_cl,
                                                                 tile0, tile_step, lower_only, out_ptrs, l_owner, l_row, scratch, busy);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_cl_fill_21cm(const double* tab, const double* chi, const double* b, const double* f, const double* pf,
                                      const double* D, const double* w, int l0, int l_step, int nl, int nz, int zint,
                                      double* out_cl, int lower_only, int variant, void* stream) {
    CB_REQUIRE(tab && chi && b && f && pf && D && w && out_cl, 1, "cl_fill_21cm: null argument");
    CB_REQUIRE(nl >= 1 && nz >= 1 && zint >= 1 && zint <= MAXZINT && l0 >= 0 && l_step >= 1, 1, "cl_fill_21cm: bad sizes nl=%d nz=%d zint=%d", nl, nz, zint);
    const long long ntiles = fill_ntiles(nz);
    CB_REQUIRE((long long)nz * (nz + 1) / 2 < 2147483647LL, 3, "cl_fill_21cm: too many channel pairs");
    return fill21_launch(tab, chi, b, f, pf, D, w, l0, l_step, nl, nz, zint, out_cl, lower_only, variant, 0, ntiles, 1, nullptr, nullptr,
                         nullptr, (cudaStream_t)stream);
}

extern "C" int cora_b200_cl_fill_21cm_tiles(const double* tab, const double* chi, const double* b, const double* f,
                                            const double* pf, const double* D, const double* w, int nl, int nz, int zint,
                                            long long tile0, long long ntiles, int tile_step, int variant, const void* out_ptrs,
                                            const int* l_owner, const int* l_row, void* stream) {
    CB_REQUIRE(tab && chi && b && f && pf && D && w && out_ptrs && l_owner && l_row, 1, "cl_fill_21cm_tiles: null argument");
    CB_REQUIRE(nl >= 1 && nz >= 1 && zint >= 1 && zint <= MAXZINT, 1, "cl_fill_21cm_tiles: bad sizes nl=%d nz=%d zint=%d", nl, nz, zint);
    const long long all = fill_ntiles(nz);
    CB_REQUIRE(tile0 >= 0 && ntiles >= 0 && tile_step >= 1 && ntiles < 2147483647LL &&
                   (ntiles == 0 || tile0 + (ntiles - 1) * (long long)tile_step < all), 1,
               "cl_fill_21cm_tiles: tiles %lld + %d k, k < %lld, outside [0, %lld)", tile0, tile_step, ntiles, all);
    if (ntiles == 0) return 0;
    return fill21_launch(tab, chi, b, f, pf, D, w, 0, 1, nl, nz, zint, nullptr, 1, variant, tile0, ntiles, tile_step,
                         (double* const*)out_ptrs, l_owner, l_row, (cudaStream_t)stream);
}

extern "C" int cora_b200_cl_symmetrize(double* cl, int nl, int nz, void* stream) {
    CB_REQUIRE(cl && nl >= 1 && nz >= 1 && nl <= 65535, 1, "cl_symmetrize: bad arguments");
    const int nb = ceil_div(nz, 32);
    cl_symmetrize_kernel<<<dim3(nb, nb, nl), 256, 0, (cudaStream_t)stream>>>(cl, nz);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_cl_romberg_reduce(const double* in, const double* w, int nl, int nz, int zint, double* out,
                                           void* stream) {
    CB_REQUIRE(in && w && out && nl >= 1 && nz >= 1 && zint >= 1, 1, "cl_romberg_reduce: bad arguments");
    CB_REQUIRE(nl <= 65535 && nz <= 65535, 1, "cl_romberg_reduce: nl, nz must be <= 65535 per call");
    romberg_reduce_kernel<<<dim3(ceil_div(nz, 128), nz, nl), 128, 0, (cudaStream_t)stream>>>(in, w, nz, zint, out);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_aps_sck_points(double A, double beta, double l_ref, double alpha, double nu_ref, double zeta,
                                        const double* l, const double* nu1, const double* nu2, long long n, double* out,
                                        void* stream) {
    CB_REQUIRE(l && nu1 && nu2 && out && n >= 1, 1, "aps_sck_points: bad arguments");
    sck_points_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(A, beta, l_ref, alpha, nu_ref, zeta, l, nu1, nu2, n, out);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_aps_21cm_points(const double* tab, const double* l, const double* vec1, const double* vec2,
                                         long long n, double* out, void* stream) {
    CB_REQUIRE(tab && l && vec1 && vec2 && out && n >= 1, 1, "aps_21cm_points: bad arguments");
    cl21_points_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(tab, l, vec1, vec2, n, out);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_ps_table_21cm_gather(const double* tab, const int* x, const int* y, int n, double* out, void* stream) {
    CB_REQUIRE(tab && x && y && out && n >= 1, 1, "ps_table_21cm_gather: bad arguments");
    ps21_gather_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(tab, x, y, n, out);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}
