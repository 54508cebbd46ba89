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
                                                        int zint, double* __restrict__ out, long long pair0,
                                                        double* const* __restrict__ out_ptrs,
                                                        const int* __restrict__ l_owner, const int* __restrict__ l_row) {
    extern __shared__ __align__(16) unsigned char smraw[];
    PairPre* pre = (PairPre*)smraw;                                     // [zint*zint]
    double* R = (double*)(smraw + sizeof(PairPre) * zint * zint);        // [FILL_EB][FILL_WMAX]
    __shared__ double s_min, s_max;
    __shared__ double s_red[256];
    // decode the channel pair.  Pairs are enumerated diagonal by diagonal (d = i - j, then j): CTAs that
    // run together then share almost the same band of table rows (y ~ |chi_i - chi_j|), which keeps
    // the band in L2 instead of re-reading it from HBM.  first(d) = d nz - d (d - 1) / 2.
    const long long pidx = pair0 + blockIdx.x;
    int d = (int)(((2.0 * nz + 1.0) - sqrt((2.0 * nz + 1.0) * (2.0 * nz + 1.0) - 8.0 * (double)pidx)) * 0.5);
    d = max(0, min(nz - 1, d));
    while (d + 1 < nz && (long long)(d + 1) * nz - (long long)(d + 1) * d / 2 <= pidx) d++;
    while (d > 0 && (long long)d * nz - (long long)d * (d - 1) / 2 > pidx) d--;
    const int j = (int)(pidx - ((long long)d * nz - (long long)d * (d - 1) / 2));
    const int i = j + d;
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
            if (i != j) o[oji] = acc[q];
        }
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

extern "C" int cora_b200_cl_fill_sck(double A, double beta, double l_ref, double alpha, double nu_ref, double zeta,
                                     const double* nu_samples, const double* w, int l0, int l_step, int nl, int nz,
                                     int zint, double* out_cl, void* stream) {
    CB_REQUIRE(nu_samples && w && out_cl, 1, "cl_fill_sck: null argument");
    CB_REQUIRE(nl >= 1 && nz >= 1 && zint >= 1 && zint <= MAXZINT && l0 >= 0 && l_step >= 1, 1, "cl_fill_sck: bad sizes nl=%d nz=%d zint=%d", nl, nz, zint);
    cudaStream_t st = (cudaStream_t)stream;
    KTimer kt(K_CL_FILL, st);
    double* bbar;
    CB_CUDA(cudaMallocAsync(&bbar, sizeof(double) * (size_t)nz * nz, st));
    sck_bbar_kernel<<<dim3(ceil_div(nz, 128), nz), 128, 0, st>>>(alpha, nu_ref, zeta, nu_samples, w, nz, zint, bbar);
    count_launch();
    CB_LAUNCH_CHECK();
    const long long nz2 = (long long)nz * nz;
    sck_scale_kernel<<<dim3(ceil_div(nz2, 256), nl), 256, 0, st>>>(A, beta, l_ref, bbar, l0, l_step, nl, nz2, out_cl);
    count_launch();
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaFreeAsync(bbar, st));
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

extern "C" int cora_b200_cl_fill_21cm(const double* tab, const double* chi, const double* b, const double* f, const double* pf,
                                      const double* D, const double* w, int l0, int l_step, int nl, int nz, int zint,
                                      double* out_cl, void* stream) {
    CB_REQUIRE(tab && chi && b && f && pf && D && w && out_cl, 1, "cl_fill_21cm: null argument");
    CB_REQUIRE(nl >= 1 && nz >= 1 && zint >= 1 && zint <= MAXZINT && l0 >= 0 && l_step >= 1, 1, "cl_fill_21cm: bad sizes nl=%d nz=%d zint=%d", nl, nz, zint);
    const long long npairs = (long long)nz * (nz + 1) / 2;
    CB_REQUIRE(npairs < 2147483647LL, 3, "cl_fill_21cm: too many channel pairs");
    size_t smem = sizeof(PairPre) * (size_t)zint * zint + sizeof(double) * FILL_EB * FILL_WMAX;
    CB_CUDA(cudaFuncSetAttribute(cl21_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KTimer kt(K_CL_FILL, (cudaStream_t)stream);
    cl21_fill_kernel<<<(unsigned)npairs, 256, smem, (cudaStream_t)stream>>>(tab, chi, b, f, pf, D, w, l0, l_step, nl, nz, zint, out_cl,
                                                                            0, nullptr, nullptr, nullptr);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

extern "C" int cora_b200_cl_fill_21cm_pairs(const double* tab, const double* chi, const double* b, const double* f,
                                            const double* pf, const double* D, const double* w, int nl, int nz, int zint,
                                            long long pair0, long long npairs, const void* out_ptrs, const int* l_owner,
                                            const int* l_row, void* stream) {
    CB_REQUIRE(tab && chi && b && f && pf && D && w && out_ptrs && l_owner && l_row, 1, "cl_fill_21cm_pairs: null argument");
    CB_REQUIRE(nl >= 1 && nz >= 1 && zint >= 1 && zint <= MAXZINT, 1, "cl_fill_21cm_pairs: bad sizes nl=%d nz=%d zint=%d", nl, nz, zint);
    const long long all = (long long)nz * (nz + 1) / 2;
    CB_REQUIRE(pair0 >= 0 && npairs >= 0 && pair0 + npairs <= all && npairs < 2147483647LL, 1,
               "cl_fill_21cm_pairs: pair range [%lld, %lld) outside [0, %lld)", pair0, pair0 + npairs, all);
    if (npairs == 0) return 0;
    size_t smem = sizeof(PairPre) * (size_t)zint * zint + sizeof(double) * FILL_EB * FILL_WMAX;
    CB_CUDA(cudaFuncSetAttribute(cl21_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KTimer kt(K_CL_FILL, (cudaStream_t)stream);
    cl21_fill_kernel<<<(unsigned)npairs, 256, smem, (cudaStream_t)stream>>>(tab, chi, b, f, pf, D, w, 0, 1, nl, nz, zint, nullptr,
                                                                            pair0, (double* const*)out_ptrs, l_owner, l_row);
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
