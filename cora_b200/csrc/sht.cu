// Batched inverse spherical-harmonic transform (alm -> HEALPix RING maps) for sm_100a.
//
// Replaces the per-channel healpy.alm2map loop of cora/util/hputil.py:500-531 (call sites
// :388, :420, :427).  Two stages, batched over all frequency channels:
//
//   1. sht_legendre_kernel: for every m, F_m(ring, chan) = sum_l lambda_lm(theta_ring) a_lm(chan).
//      lambda_lm is produced on the fly by a scaled three-term recurrence in registers (one
//      ring per lane), handed to the FP64 tensor cores (DMMA.8x8x4) through a warp-private
//      shared-memory tile, and contracted against a cp.async double-buffered alm tile.  North /
//      south rings share lambda through the (l-m) parity split: two accumulator sets per warp.
//   2. sht_phase_kernel: per ring, fold m onto m mod nph (exact aliasing of the direct sum),
//      pair two channels into one complex transform, in-shared-memory power-of-two FFT
//      (rings with nph = 2^k) or Bluestein chirp-z on a power-of-two FFT (all other rings),
//      coalesced stores into map[chan][pix].
//
// Conventions follow SURVEY.md App. A.7-A.8 / oracle/sht.py.
#include "common.cuh"
#include "cora_b200.h"

#include <cuda.h>   // CUtensorMap (the encode entry point is fetched through the runtime, no -lcuda)

#include <cmath>
#include <cstdlib>
#include <vector>
#include <algorithm>

namespace cb {

// ------------------------------------------------------------------------------- plan
struct RingDesc {
    long long start;      // first pixel of the ring (RING order)
    int nph;              // pixels in ring
    int shifted;          // phi0 = pi/nph (1) or 0 (0)
    int cap;              // index into the per-size Bluestein tables (cap ring number i), or -1
    int logM;             // FFT size 2^logM: nph itself (power of two) or the Bluestein size >= 2 nph - 1
    int bluestein;
};

// rings launched together: all rings whose FFT fits Mmax points of shared memory per channel pair
struct PhaseClass {
    int Mmax, P, threads;           // smem FFT capacity, channel pairs per CTA, CTA size
    int blu;                        // 1: Bluestein rings only, 0: power-of-two rings only
    int nrings;
    int* d_rings;                   // ring indices of this class, longest ring first
};

struct ShtPlan {
    int nside, lmax, nrn;           // nrn = 2*nside north rings incl. equator
    long long npix, nalm;
    double *d_cth, *d_sth;          // [nrn]
    double2* d_rc;                  // [nalm] recurrence coefficients, row idx(l, m) holds those of step l -> l+1
    double* d_rc2;                  // [nalm][4] spin-2 coefficients of (l, m), built on first polarised use
    double* d_nm_mant;              // [lmax+1]   N_m = mant * 2^exp
    int* d_nm_exp;                  // [lmax+1]
    RingDesc* d_rings;              // [4*nside-1]
    std::vector<RingDesc> h_rings;
    double2* d_tw;                  // exp(2 pi i j / tw_n), j < tw_n/2
    int tw_n, log_tw;
    double2* d_chirp;               // Bluestein chirps  c_k = exp(i pi k^2 / nph), per cap ring
    double2* d_shph;                // ring phase e^{i pi k / nph}, k <= nph/2, per Bluestein ring (offsets = chirp's)
    double2* d_bhat;                // FFT_M(conj chirp)/M in bit-reversed order, per cap ring
    long long *d_chirp_off, *d_bhat_off;   // [nside] offsets by cap ring number
    std::vector<PhaseClass> classes;
};

// M = 4096 phase classes: 1 (default) = one channel pair per 256-thread CTA (74 KB, 3 CTAs per SM), 0 = two pairs per
// 512-thread CTA (147 KB, 1 CTA per SM).  Measured at nside 512 x 1024 channels: phase stage 48.2 -> 44.0 ms with 1.
static int g_phase_4096_split = [] { const char* e = getenv("CORA_B200_PHASE_4096_SPLIT"); return (e && e[0] == '0') ? 0 : 1; }();

static int ilog2(int x) { int l = 0; while ((1 << l) < x) l++; return l; }
static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// ------------------------------------------------------------------ in-smem batched FFT
// x: [P][M] complex.  In-place radix-2.  DIT: bit-reversed in -> natural out.
// DIF: natural in -> bit-reversed out.  sign = +1: exp(+2 pi i jk/M).
template <bool DIT>
__device__ __forceinline__ void fft_batch(double2* x, int M, int logM, int P, int sign,
                                          const double2* __restrict__ tw, int log_tw) {
    const int nb = M >> 1;
    const int total = P * nb;
    for (int st = 0; st < logM; st++) {
        const int s = DIT ? st : (logM - 1 - st);
        const int h = 1 << s;
        for (int w = threadIdx.x; w < total; w += blockDim.x) {
            const int p = w / nb;
            const int b = w - p * nb;
            const int q = b & (h - 1);
            const int i0 = ((b >> s) << (s + 1)) + q;
            const int i1 = i0 + h;
            double2 wv = tw[(size_t)q << (log_tw - 1 - s)];
            if (sign < 0) wv.y = -wv.y;
            double2* xp = x + (size_t)p * M;
            double2 u = xp[i0], v = xp[i1];
            if (DIT) {
                v = cmul(v, wv);
                xp[i0] = make_double2(u.x + v.x, u.y + v.y);
                xp[i1] = make_double2(u.x - v.x, u.y - v.y);
            } else {
                xp[i0] = make_double2(u.x + v.x, u.y + v.y);
                xp[i1] = cmul(make_double2(u.x - v.x, u.y - v.y), wv);
            }
        }
        __syncthreads();
    }
}

// lam_{l+1} = c1 x lam_l - c2 lam_{l-1}:  rc[idx(l, m)] = (c1, c2) of the step that produces l + 1
// (zeros once l + 1 > lmax, so the recurrence runs cleanly off the end of the table).
__global__ void recur_coef_kernel(double2* rc, int lmax) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (l < m || l > lmax) return;
    const double ln = (double)(l + 1), mm = (double)m;
    double a = 0.0, r = 0.0;
    if (l + 1 <= lmax) {
        a = sqrt((4.0 * ln * ln - 1.0) / (ln * ln - mm * mm));
        if (l + 1 >= m + 2) {
            const double lp = ln - 1.0;
            r = a / sqrt((4.0 * lp * lp - 1.0) / (lp * lp - mm * mm));
        }
    }
    rc[(long long)m * (2 * lmax + 1 - m) / 2 + l] = make_double2(a, r);
}

// spin-2 coefficients of (l, m) (the same expressions as sht_legendre_kernel<2>): with
// tn = 2 / sqrt((l+2)(l+1) l (l-1)),  tg = tn sqrt((2l+1)/(2l-1) (l^2 - m^2)):
//   X1 = -(c0 / sin^2 + c1) lam_l + c3 cos / sin^2 lam_{l-1},   X2 = -c2 cos / sin^2 lam_l + m c3 / sin^2 lam_{l-1}
// c = (tn (l - m^2), tn l (l-1) / 2, tn m (l-1), tg); zero for l < 2.
__global__ void spin2_coef_kernel(double* rc2, int lmax) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (l < m || l > lmax) return;
    const double ll = (double)l, mm = (double)m;
    double tn = 0.0, tg = 0.0;
    if (l >= 2) {
        tn = 2.0 / sqrt((ll + 2.0) * (ll + 1.0) * ll * (ll - 1.0));
        tg = tn * sqrt((2.0 * ll + 1.0) / (2.0 * ll - 1.0) * (ll * ll - mm * mm));
    }
    double* o = rc2 + 4 * ((long long)m * (2 * lmax + 1 - m) / 2 + l);
    o[0] = tn * (ll - mm * mm);
    o[1] = 0.5 * tn * ll * (ll - 1.0);
    o[2] = tn * mm * (ll - 1.0);
    o[3] = tg;
}

__global__ void twiddle_kernel(double2* tw, int n) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n / 2) {
        double s, c;
        sincospi(2.0 * (double)j / (double)n, &s, &c);
        tw[j] = make_double2(c, s);
    }
}

// chirp c_k = exp(+i pi k^2 / n) with k^2 reduced mod 2n exactly.
__device__ __forceinline__ double2 chirp_val(long long k, int n) {
    long long r = (k * k) % (2LL * n);
    double s, c;
    sincospi((double)r / (double)n, &s, &c);
    return make_double2(c, s);
}

// One CTA per cap ring size: chirp table and bhat = DIF-FFT_M(b)/M with b_d = conj(c_d), d = -(n-1)..(n-1).
__global__ void bluestein_setup_kernel(int nside, const long long* chirp_off, const long long* bhat_off,
                                       double2* chirp, double2* bhat, double2* shph, const double2* tw, int log_tw) {
    extern __shared__ __align__(16) double2 xs[];
    const int i = blockIdx.x + 1;  // cap ring number
    const int n = 4 * i;
    if ((n & (n - 1)) == 0) return;  // power of two: direct FFT, no tables
    int M = 1, logM = 0;
    while (M < 2 * n - 1) { M <<= 1; logM++; }
    for (int k = threadIdx.x; k < M; k += blockDim.x) xs[k] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double2 c = chirp_val(k, n);
        chirp[chirp_off[i] + k] = c;
        double sn, cs;
        sincospi((double)k / (double)n, &sn, &cs);
        shph[chirp_off[i] + k] = make_double2(cs, sn);
        double2 b = cconj(c);
        xs[k] = b;
        if (k > 0) xs[M - k] = b;
    }
    __syncthreads();
    fft_batch<false>(xs, M, logM, 1, -1, tw, log_tw);
    const double inv = 1.0 / (double)M;
    for (int k = threadIdx.x; k < M; k += blockDim.x)
        bhat[bhat_off[i] + k] = make_double2(xs[k].x * inv, xs[k].y * inv);
}

// --------------------------------------------------------------------- layout transpose
// healpy-packed channel-major alm[chan][idx] -> panel layout almT[idx][chan_batch].
__global__ void alm_transpose_kernel(const double2* __restrict__ in, long long in_stride, int nchan,
                                     long long nalm, double2* __restrict__ out, int out_stride) {
    __shared__ double2 tile[32][33];
    const long long i0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r;
        long long i = i0 + threadIdx.x;
        if (c < nchan && i < nalm) tile[r][threadIdx.x] = in[(long long)c * in_stride + i];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        long long i = i0 + r;
        int c = c0 + threadIdx.x;
        if (c < nchan && i < nalm) out[i * out_stride + c] = tile[threadIdx.x][r];
    }
}

// ------------------------------------------------------------------------ Legendre stage
constexpr int LEG_THREADS = 256;
constexpr int LEG_WARPS = 8;
constexpr int LEG_RT = 256;   // north rings per CTA (one per thread)
constexpr int LEG_NCH = 16;   // complex channels per CTA  (32 real columns = 4 n8 blocks)
constexpr int LEG_KC = 64;    // l's per alm chunk
constexpr int LEG_GPC = LEG_KC / 8;   // 8-l groups per chunk
constexpr int LEG_ALD = 36;   // A tile [8][36]: (par*4+t)*36 + ring -> conflict-free LDS.64 per half-warp
constexpr int LEG_BLD = 34;   // Bs[KC][34]: (2t+par)*34 + col

struct LegParams {
    const double2* almT;      // [nalm][alm_stride] panel layout (spin 2: the E panel)
    const double2* almB;      // spin 2 only: the B panel (same strides)
    long long alm_stride;     // in double2
    int chan0;                // first channel of this batch inside almT rows
    int nb;                   // channels in this batch ( = width of F rows)
    double2* F;               // [4*nside-1][lmax+1][nb]   (spin 2: Q)
    double2* F2;              // spin 2 only: U
    const double *cth, *sth;  // [nrn]
    const double* nm_mant;
    const int* nm_exp;
    const double2* rc;        // [nalm] recurrence coefficients (a_{l+1}, a_{l+1}/a_l) at idx(l, m)  (ws kernel)
    const double* rc2;        // [nalm][4] spin-2 coefficients of (l, m): tn (l - m^2), tn l (l-1) / 2, tn m (l-1), tg  (ws kernel)
    int nside, lmax, nrn, nrb, ncb, Lpad, ncg;   // ncg = ceil(nb / 4) channel groups in F
};

__device__ __forceinline__ void norm_frexp(double& m, long long& e) {
    int ee;
    m = frexp(m, &ee);
    e += ee;
}

// warp-uniformly predicated DMMA: a dead 8-ring block costs an issue slot, not pipe time
__device__ __forceinline__ void dmma884_p(double& c0, double& c1, double a, double b, int pred) {
    asm("{\n .reg .pred p;\n setp.ne.s32 p, %4, 0;\n"
        " @p mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n}\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b), "r"(pred));
}

// SPIN = 0: scalar synthesis, 16 complex channels per CTA (32 real GEMM columns).
// SPIN = 2: (E,B) -> (Q,U) with the HEALPix X1/X2 functions, 8 channels per
//   CTA; GEMM columns per channel are (Q_re, Q_im, U_re, U_im) and the K dimension is the
//   concatenation [X1 | X2] against B1 = (aE_re, aE_im, aB_re, aB_im), B2 = (-aB_im, aB_re,
//   aE_im, -aE_re), so Q = -(X1 aE + i X2 aB), U = -(X1 aB - i X2 aE) accumulate in place.
//   X2 has the opposite theta-parity of X1, so it feeds the other parity accumulator.
//
// Schedule.  The l range of one m is walked in groups of 8.  Each warp owns 32 rings (four
// octets dealt round-robin over the warps so that the polar, late-starting rings are spread
// evenly) and is software-pipelined: while the DMMAs of group g run from A-tile buffer g&1,
// the lane's recurrence for group g+1 fills buffer (g+1)&1 in the same basic block, so the
// dependent FP64 chain hides in the DMMA issue gaps.  The scale exponent is examined once per
// group; a lane is "live" for a whole group, and 8-ring blocks with no live lane are predicated
// off (whole warps: branched over).
template <int SPIN>
__global__ void __launch_bounds__(LEG_THREADS, 1) sht_legendre_kernel(LegParams P) {
    extern __shared__ __align__(16) double smem[];
    constexpr int NCH = (SPIN == 0) ? LEG_NCH : LEG_NCH / 2;   // channels per CTA
    constexpr int AROWS = (SPIN == 0) ? 8 : 16;
    double2* cc = (double2*)smem;                              // [Lpad]  (a_l, a_l / a_{l-1})
    double2* cs = cc + P.Lpad;                                 // [Lpad]  spin 2: (2 n_l, 2 n_l g_lm)
    double* Bs = (double*)(cs + (SPIN ? P.Lpad : 0));          // [2][KC][BLD]
    double* As = Bs + 2 * LEG_KC * LEG_BLD;                    // [WARPS][2][AROWS][ALD]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    int bid = blockIdx.x;
    const int cb = bid % P.ncb; bid /= P.ncb;
    const int rb = bid % P.nrb;
    const int m = bid / P.nrb;
    const int lmax = P.lmax;
    const int nk = lmax - m + 1;               // number of l's

    // recurrence coefficients for this m:  lam_k = c1[k] x lam_{k-1} - c2[k] lam_{k-2},  k = l - m
    for (int k = tid; k < P.Lpad; k += LEG_THREADS) {
        double a = 0.0, r = 0.0;
        const double l = (double)(m + k), mm = (double)m;
        if (k >= 1 && k < nk) {
            a = sqrt((4.0 * l * l - 1.0) / (l * l - mm * mm));
            if (k >= 2) {
                double lp = l - 1.0;
                double ap = sqrt((4.0 * lp * lp - 1.0) / (lp * lp - mm * mm));
                r = a / ap;
            }
        }
        cc[k] = make_double2(a, r);
        if (SPIN) {
            double tn = 0.0, tg = 0.0;
            if (k < nk && m + k >= 2) {
                tn = 2.0 / sqrt((l + 2.0) * (l + 1.0) * l * (l - 1.0));
                tg = tn * sqrt((2.0 * l + 1.0) / (2.0 * l - 1.0) * (l * l - mm * mm));
            }
            cs[k] = make_double2(tn, tg);
        }
    }

    // ring owned by this lane: octet (lane/8)*8 + warp of the CTA's 32 octets
    const int rn = rb * LEG_RT + (((lane >> 3) * 8 + warp) << 3) + (lane & 7);   // north ring index (ring number rn+1)
    const bool ring_ok = rn < P.nrn;
    double x = 0.0, p_cur = 0.0, p_prev = 0.0;
    double is2 = 0.0, cs2 = 0.0;   // spin 2: 1/sin^2, cos/sin^2
    int e = -(1 << 20);  // scale exponent (multiple of 256, <= 0); true value = p * 2^e
    if (ring_ok) {
        x = P.cth[rn];
        // seed lambda_mm = (-1)^m N_m sin^m(theta) as (mantissa, exponent)
        double bm = P.sth[rn];
        if (SPIN) { is2 = 1.0 / (bm * bm); cs2 = x * is2; }
        long long be = 0;
        norm_frexp(bm, be);
        double rm = 1.0;
        long long re = 0;
        int n = m;
        while (n) {
            if (n & 1) { rm *= bm; re += be; norm_frexp(rm, re); }
            bm *= bm; be *= 2; norm_frexp(bm, be);
            n >>= 1;
        }
        rm *= P.nm_mant[m];
        re += P.nm_exp[m];
        norm_frexp(rm, re);
        if (m & 1) rm = -rm;
        // e = 256*ceil(re/256) clipped to <= 0
        long long q = (re >= 0) ? 0 : -((-re) / 256);
        e = (int)(q * 256);
        long long sh = re - (long long)e;   // in (-256, 0] when e < 0, = re when e == 0
        p_cur = ldexp(rm, (int)sh);
    }

    double acc[2][4][4][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[a][b][c][0] = acc[a][b][c][1] = 0.0;

    // alm chunk loader: rows l = m + kc + j (j < KC); 16 x 16-byte pieces per row
    const long long row0 = (long long)m * (2 * lmax + 1 - m) / 2 + m;   // idx(l=m, m)
    const int nchunk = (nk + LEG_KC - 1) / LEG_KC;
    const int ngroups = (nk + 7) / 8;
    auto load_chunk = [&](int ck, int buf) {
#pragma unroll
        for (int qd = 0; qd < LEG_KC * 16 / LEG_THREADS; qd++) {
            int el = tid + qd * LEG_THREADS;
            int j = el >> 4, c = el & 15;
            int k = ck * LEG_KC + j;
            int ch;
            const double2* base;
            if (SPIN == 0) { ch = cb * NCH + c; base = P.almT; }
            else { ch = cb * NCH + (c >> 1); base = (c & 1) ? P.almB : P.almT; }
            bool ok = (k < nk) && (ch < P.nb);
            const double2* src = base + (ok ? ((row0 + k) * P.alm_stride + P.chan0 + ch) : 0);
            cp_async16(Bs + (size_t)buf * LEG_KC * LEG_BLD + j * LEG_BLD + 2 * c, src, ok);
        }
        cp_async_commit();
    };

    // 8 recurrence steps starting at k = kbase, emitted into A-tile buffer Ab (rows split by l parity).
    // Returns the ballot of lanes that are live (unscaled) for this group.
    auto recur8 = [&](double* Ab, int kbase) -> unsigned {
        const bool live = (e == 0);
        double2 c[8];
#pragma unroll
        for (int j = 0; j < 8; j++) c[j] = cc[kbase + 1 + j];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int arow = (j & 1) * 4 + (j >> 1);
            if (SPIN == 0) {
                Ab[arow * LEG_ALD + lane] = live ? p_cur : 0.0;
            } else {
                const double2 sc = cs[kbase + j];
                const double l = (double)(m + kbase + j), mm = (double)m;
                const double tn = sc.x, tg = sc.y;
                const double al = tn * (l - mm * mm), be = 0.5 * tn * l * (l - 1.0), de = tn * mm * (l - 1.0);
                const double x1 = -(al * is2 + be) * p_cur + tg * cs2 * p_prev;
                const double x2 = -de * cs2 * p_cur + mm * tg * is2 * p_prev;
                Ab[arow * LEG_ALD + lane] = live ? x1 : 0.0;
                Ab[(8 + arow) * LEG_ALD + lane] = live ? x2 : 0.0;
            }
            const double pn = fma(c[j].x * x, p_cur, -c[j].y * p_prev);
            p_prev = p_cur;
            p_cur = pn;
        }
        if (e < 0 && ((__double2hiint(p_cur) >> 20) & 0x7ff) > 1023 + 128) {
            p_cur *= 0x1p-256;
            p_prev *= 0x1p-256;
            e += 256;
        }
        return __ballot_sync(0xffffffffu, live);
    };

    double* Aw = As + warp * 2 * AROWS * LEG_ALD;
    load_chunk(0, 0);
    __syncthreads();   // coefficient arrays visible
    unsigned bal = recur8(Aw, 0);

    // spin 2: B2 fragment = +-B1 at a permuted column inside each group of four
    const int gperm = (g & 4) + 3 - (g & 3);
    const double gsign = ((g & 3) == 0 || (g & 3) == 3) ? -1.0 : 1.0;

    int gi = 0;   // group being contracted
    for (int ck = 0; ck < nchunk; ck++) {
        cp_async_wait<0>();
        __syncthreads();
        if (ck + 1 < nchunk) load_chunk(ck + 1, (ck + 1) & 1);
        const double* Bb = Bs + (size_t)(ck & 1) * LEG_KC * LEG_BLD;
        const int gend = min(ngroups, (ck + 1) * LEG_GPC);
#pragma unroll 1
        for (; gi < gend; gi++) {
            __syncwarp();   // buffer gi&1 (written one iteration ago by every lane) is visible
            const double* Ac = Aw + (gi & 1) * AROWS * LEG_ALD;
            double* An = Aw + ((gi + 1) & 1) * AROWS * LEG_ALD;
            const int grp = gi - ck * LEG_GPC;
            unsigned bal_next;
            if (bal) {
                int pm[4];
#pragma unroll
                for (int mb = 0; mb < 4; mb++) pm[mb] = (int)((bal >> (8 * mb)) & 0xffu);
                double af[2][4], bf[2][4];
#pragma unroll
                for (int par = 0; par < 2; par++) {
#pragma unroll
                    for (int mb = 0; mb < 4; mb++) af[par][mb] = Ac[(par * 4 + t) * LEG_ALD + 8 * mb + g];
#pragma unroll
                    for (int nb = 0; nb < 4; nb++) bf[par][nb] = Bb[(grp * 8 + 2 * t + par) * LEG_BLD + 8 * nb + g];
                }
                bal_next = recur8(An, (gi + 1) * 8);
#pragma unroll
                for (int par = 0; par < 2; par++)
#pragma unroll
                    for (int mb = 0; mb < 4; mb++)
#pragma unroll
                        for (int nb = 0; nb < 4; nb++)
                            dmma884_p(acc[par][mb][nb][0], acc[par][mb][nb][1], af[par][mb], bf[par][nb], pm[mb]);
                if (SPIN) {
                    // X2 rows of l-parity `par` have theta-parity 1-par
#pragma unroll
                    for (int par = 0; par < 2; par++) {
#pragma unroll
                        for (int mb = 0; mb < 4; mb++) af[par][mb] = Ac[(8 + par * 4 + t) * LEG_ALD + 8 * mb + g];
#pragma unroll
                        for (int nb = 0; nb < 4; nb++)
                            bf[par][nb] = gsign * Bb[(grp * 8 + 2 * t + par) * LEG_BLD + 8 * nb + gperm];
                    }
#pragma unroll
                    for (int par = 0; par < 2; par++)
#pragma unroll
                        for (int mb = 0; mb < 4; mb++)
#pragma unroll
                            for (int nb = 0; nb < 4; nb++)
                                dmma884_p(acc[par ^ 1][mb][nb][0], acc[par ^ 1][mb][nb][1], af[par][mb], bf[par][nb], pm[mb]);
                }
            } else {
                bal_next = recur8(An, (gi + 1) * 8);
            }
            bal = bal_next;
        }
    }

    // ---- epilogue: north = even + odd, south = even - odd;  F[ring][m][chan]
    const int L = lmax + 1;
    const int nring_tot = 4 * P.nside - 1;
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
        const int rr = rb * LEG_RT + ((mb * 8 + warp) << 3) + g;   // north ring index
        if (rr >= P.nrn) continue;
        const int r_n = rr;                       // ring array index of the north ring
        const int r_s = nring_tot - 1 - rr;       // mirror ring
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            double er = acc[0][mb][nb][0], ei = acc[0][mb][nb][1];
            double orr = acc[1][mb][nb][0], oi = acc[1][mb][nb][1];
            if (SPIN == 0) {
                const int ch = cb * NCH + 4 * nb + t;
                if (ch >= P.nb) continue;
                const int cg = ch >> 2, cc = ch & 3;
                P.F[(((long long)r_n * P.ncg + cg) * L + m) * 4 + cc] = make_double2(er + orr, ei + oi);
                if (r_s != r_n) P.F[(((long long)r_s * P.ncg + cg) * L + m) * 4 + cc] = make_double2(er - orr, ei - oi);
            } else {
                const int ch = cb * NCH + 2 * nb + (t >> 1);
                if (ch >= P.nb) continue;
                double2* Fo = (t & 1) ? P.F2 : P.F;   // even t: Q, odd t: U
                const int cg = ch >> 2, cc = ch & 3;
                Fo[(((long long)r_n * P.ncg + cg) * L + m) * 4 + cc] = make_double2(-(er + orr), -(ei + oi));
                if (r_s != r_n) Fo[(((long long)r_s * P.ncg + cg) * L + m) * 4 + cc] = make_double2(-(er - orr), -(ei - oi));
            }
        }
    }
}

// ------------------------------------------------- Legendre stage, warp-specialised (scalar)
// Same arithmetic as sht_legendre_kernel<0>, restructured for the Blackwell SM.  4 producer warps
// run the lambda_lm recurrence (one ring per lane) into a ring of shared-memory A tiles; 8 consumer
// warps do nothing but fragment loads and DMMAs (two consumers, 32 columns each, share a
// producer's 32 rings); the alm operand arrives through the TMA engine as 2-D tiled loads
// (cp.async.bulk.tensor, 128-byte swizzle; SASS UTMALDG) together with the chunk's recurrence
// coefficients (precomputed once per plan) into a 4-deep ring of B tiles.  All hand-offs are
// mbarrier full/empty pairs at chunk (32 l) granularity -- one probe costs 100-150 cycles.
//
// Measured behaviour that shapes this (profiles/r01): scalar FP64 and DMMA share one pipe and a
// warp's DFMA is starved while two other warps keep that pipe full of DMMAs, so a warp that
// alternates recurrence and DMMA (sht_legendre_kernel) serialises them; here the recurrence lives
// in its own warps, runs ahead by up to three chunks and only costs pipe time when it executes.
namespace lws {
constexpr int NCONS = 8, NPROD = 4, RPL = 1;   // consumer warps; producer warps; rings per producer lane
constexpr int NVP = NPROD * RPL;                // ring groups of 32 ("virtual producers")
constexpr int THREADS = 32 * (NCONS + NPROD);   // 12 warps: 2 consumers + 1 producer per SM sub-partition
constexpr int RT = 32 * NVP;     // rings per work item
constexpr int NCOL = 64;         // real columns per work item (32 channels)
constexpr int GPC = 4;           // 8-l groups per B chunk
constexpr int KC = 8 * GPC;      // l's per B chunk
constexpr int NSB = 4;           // B ring depth in chunks of GPC groups
// A ring depth: 3 chunks for the scalar transform; 2 for spin 2, whose A tiles hold X1 and X2 (16 rows per
// 8-l group, 147 KB for two stages)
template <int SPIN> struct Cfg { static constexpr int NSA = SPIN ? 2 : 3; static constexpr int AROWS = SPIN ? 16 : 8; };
constexpr int ALD = 36;          // A tile [AROWS][36]
// B tile of one chunk: NCOL/16 column blocks of [KC rows][16 doubles = 128 B], each written by one
// 2-D TMA box with the 128-byte swizzle (16-byte unit u of row r lands at unit u ^ (r & 7)), which
// makes the DMMA fragment reads below conflict-free without padding.
constexpr int BBLK = KC * 16;    // doubles per column block
constexpr int BSTAGE = (NCOL / 16) * BBLK;
// (no setmaxnreg: 12 warps x 168 registers fill the register file and both roles fit in 168)
}  // namespace lws

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    // test_wait, not try_wait: a poll must not suspend the thread until the hardware time limit
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n .reg .pred p;\n"
        "MBAR_WAIT:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra MBAR_DONE;\n"
        " bra MBAR_WAIT;\n"
        "MBAR_DONE:\n}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Persistent: one CTA per SM walks the work items (m, ring block, channel block) it is dealt
// round-robin -- items are ordered by m, i.e. by decreasing cost, so the deal is balanced, and the
// CTAs that run concurrently share the same few m (their alm rows hit in L2).  Roles run ahead of
// each other across item boundaries: while the consumers drain item i and write its F rows, the
// producers already compute the recurrence table and seeds of item i+1 and the TMA warp fetches
// its first alm chunks.
// 2-D tiled TMA load (SASS: UTMALDG): box of the tensor map at (c0 = column in doubles, c1 = row)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

// SPIN = 0: scalar synthesis, a work item = (m, 128 rings, 32 channels).
// SPIN = 2: (E, B) -> (Q, U), a work item = (m, 128 rings, 16 channels).  The producers emit the HEALPix X1 / X2
//   functions from lambda_l, lambda_{l-1} and four per-(l, m) coefficients that arrive with the
//   chunk; the 64 GEMM columns of an item are, per consumer half h (8 channels): [Q of 4 channels | Q of the next 4 |
//   U of 4 | U of the next 4], each channel a (re, im) pair.  Q = -(X1 aE + i X2 aB), U = -(X1 aB - i X2 aE): the X1
//   product reads the E tile for the Q blocks and the B tile for the U blocks; the X2 product reads the OTHER tile with
//   re <-> im swapped inside the channel's 16-byte unit and a sign, and -- X2 having the opposite theta parity --
//   feeds the other parity accumulator.  Both alm panels arrive through their own tensor map (blocks 0,1 = E tile,
//   2,3 = B tile of the B stage).
template <int SPIN>
__global__ void __launch_bounds__(lws::THREADS, 1) sht_legendre_ws_kernel(LegParams P, int nitems,
                                                                           const __grid_constant__ CUtensorMap tmapB,
                                                                           const __grid_constant__ CUtensorMap tmapB2) {
    using namespace lws;
    constexpr int NSA = Cfg<SPIN>::NSA, AROWS = Cfg<SPIN>::AROWS;
    extern __shared__ __align__(16) double smem[];
    // the 128-byte swizzle pattern repeats every 1024 bytes of shared-memory address: align the B ring explicitly
    double* Bs = (double*)((char*)smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u));   // [NSB][NCOL/16][KC][16]
    double2* Cs = (double2*)(Bs + NSB * BSTAGE);                    // [NSB][KC] recurrence coefficients of the chunk
    double* Cs2 = (double*)(Cs + NSB * KC);                         // [NSB][KC][4] spin-2 coefficients of the chunk
    double* As = Cs2 + (SPIN ? NSB * KC * 4 : 0);                   // [NVP][NSA][GPC][AROWS][ALD]
    unsigned* balA = (unsigned*)(As + NVP * NSA * GPC * AROWS * ALD);   // [NVP][NSA][GPC] live-lane ballots
    unsigned long long* bars = (unsigned long long*)(balA + NVP * NSA * GPC);
    unsigned long long* fullA = bars;                               // [NPROD][NSA]
    unsigned long long* emptyA = fullA + NPROD * NSA;               // [NPROD][NSA]
    unsigned long long* fullB = emptyA + NPROD * NSA;               // [NSB]
    unsigned long long* emptyB = fullB + NSB;                       // [NSB]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lmax = P.lmax;

    if (tid == 0) {
        for (int i = 0; i < NPROD * NSA; i++) { mbar_init(fullA + i, 1); mbar_init(emptyA + i, 2 * RPL); }
        for (int i = 0; i < NSB; i++) { mbar_init(fullB + i, 32); mbar_init(emptyB + i, NCONS + NPROD); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= NCONS) {
        // ================================================================= producer warp: lambda_lm
        const int p = warp - NCONS;
        unsigned cglob = 0;                       // chunks produced so far
        // alm chunks: producer p also feeds the B ring with the chunks whose running number is
        // p mod NPROD (one 512-byte row per lane through the TMA engine), polled from its wait loops
        // so that it never blocks the recurrence.  (b_it, b_c) = item and chunk of the next one.
        int b_it = blockIdx.x, b_c = 0;
        unsigned b_glob = 0;
        auto b_nchunk = [&](int it) { return (lmax - it / (P.ncb * P.nrb) + 1 + KC - 1) / KC; };
        auto b_advance = [&](int n) {
            for (int i = 0; i < n && b_it < nitems; i++) {
                b_glob++;
                if (++b_c == b_nchunk(b_it)) { b_c = 0; b_it += gridDim.x; }
            }
        };
        b_advance(p);
        auto service_B = [&]() {
            if (b_it >= nitems) return;
            const int sb = b_glob % NSB;
            unsigned ready = 0;
            if (lane == 0) ready = mbar_test(emptyB + sb, ((b_glob / NSB) & 1) ^ 1) ? 1u : 0u;
            if (!__shfl_sync(0xffffffffu, ready, 0)) return;
            const int bm = b_it / (P.ncb * P.nrb), bcb = b_it % P.ncb;
            const int bnk = lmax - bm + 1;
            const long long brow0 = (long long)bm * (2 * lmax + 1 - bm) / 2 + bm;   // idx(l = m, m)
            const int nrows = min(KC, bnk - b_c * KC);
            // coefficient rows past lmax: zeros (the alm rows there are finite -- they belong to the next m or
            // are zero-filled by the TMA unit beyond the tensor -- and meet lambda = 0)
            if (lane >= nrows) {
                Cs[sb * KC + lane] = make_double2(0.0, 0.0);
                if (SPIN) {
#pragma unroll
                    for (int q = 0; q < 4; q++) Cs2[(sb * KC + lane) * 4 + q] = 0.0;
                }
            }
            if (lane == 0) {
                double* dst = Bs + (size_t)sb * BSTAGE;
                const int row = (int)(brow0 + (long long)b_c * KC);
                if (SPIN == 0) {
#pragma unroll
                    for (int jb = 0; jb < NCOL / 16; jb++) tma_load_2d(dst + jb * BBLK, &tmapB, bcb * NCOL + 16 * jb, row, fullB + sb);
                } else {
                    // 16 channels = 32 doubles of each panel row: two boxes per panel
#pragma unroll
                    for (int jb = 0; jb < 2; jb++) {
                        tma_load_2d(dst + jb * BBLK, &tmapB, bcb * (NCOL / 2) + 16 * jb, row, fullB + sb);
                        tma_load_2d(dst + (2 + jb) * BBLK, &tmapB2, bcb * (NCOL / 2) + 16 * jb, row, fullB + sb);
                    }
                }
                bulk_g2s(Cs + sb * KC, P.rc + brow0 + (long long)b_c * KC, (unsigned)nrows * 16u, fullB + sb);
                if (SPIN) bulk_g2s(Cs2 + sb * KC * 4, P.rc2 + 4 * (brow0 + (long long)b_c * KC), (unsigned)nrows * 32u, fullB + sb);
                mbar_arrive_expect_tx(fullB + sb, (unsigned)(BSTAGE * 8) + (unsigned)nrows * (SPIN ? 48u : 16u));
            } else {
                mbar_arrive(fullB + sb);
            }
            b_advance(NPROD);
        };
        for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
            const int rb = (it / P.ncb) % P.nrb;
            const int m = it / (P.ncb * P.nrb);
            const int nk = lmax - m + 1;
            const int ngroups = (nk + 7) / 8;
            // RPL rings per lane (independent recurrence chains -> the FP64 latency overlaps): ring group
            // v = p * RPL + u; inside a group the lane owns octet (lane/8)*NVP + v of the item's octets
            double x[RPL], p_cur[RPL], p_prev[RPL];
            double is2[RPL], cs2[RPL];   // spin 2: 1/sin^2, cos/sin^2
            int e[RPL];
#pragma unroll
            for (int u = 0; u < RPL; u++) {
                const int v = p * RPL + u;
                const int rn = rb * RT + (((lane >> 3) * NVP + v) << 3) + (lane & 7);
                x[u] = 0.0; p_cur[u] = 0.0; p_prev[u] = 0.0; is2[u] = 0.0; cs2[u] = 0.0;
                e[u] = -(1 << 20);
                if (rn < P.nrn) {
                    x[u] = P.cth[rn];
                    double bm = P.sth[rn];
                    if (SPIN) { is2[u] = 1.0 / (bm * bm); cs2[u] = x[u] * is2[u]; }
                    long long be = 0;
                    norm_frexp(bm, be);
                    double rm = 1.0;
                    long long re = 0;
                    int n = m;
                    while (n) {
                        if (n & 1) { rm *= bm; re += be; norm_frexp(rm, re); }
                        bm *= bm; be *= 2; norm_frexp(bm, be);
                        n >>= 1;
                    }
                    rm *= P.nm_mant[m];
                    re += P.nm_exp[m];
                    norm_frexp(rm, re);
                    if (m & 1) rm = -rm;
                    long long q = (re >= 0) ? 0 : -((-re) / 256);
                    e[u] = (int)(q * 256);
                    p_cur[u] = ldexp(rm, (int)(re - (long long)e[u]));
                }
            }
            const double mm = (double)m;
            const int nchunk = (ngroups + GPC - 1) / GPC;
            for (int c = 0; c < nchunk; c++, cglob++) {
                const int sa = cglob % NSA, sbc = cglob % NSB;
                const int ng = min(GPC, ngroups - c * GPC);
                service_B();
                // hand-offs are per chunk (an mbarrier probe costs ~100-150 cycles): the chunk's
                // coefficients have landed, and the consumers are done with this A stage
                for (;;) {
                    unsigned ready = 0;
                    if (lane == 0)
                        ready = (mbar_test(fullB + sbc, (cglob / NSB) & 1) && mbar_test(emptyA + p * NSA + sa, ((cglob / NSA) & 1) ^ 1)) ? 1u : 0u;
                    if (__shfl_sync(0xffffffffu, ready, 0)) break;
                    service_B();
                }
                for (int grp = 0; grp < ng; grp++) {
                    bool live[RPL];
                    double* Ab[RPL];
#pragma unroll
                    for (int u = 0; u < RPL; u++) {
                        live[u] = (e[u] == 0);
                        Ab[u] = As + ((size_t)((p * RPL + u) * NSA + sa) * GPC + grp) * AROWS * ALD;
                    }
                    double2 cf[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) cf[j] = Cs[sbc * KC + grp * 8 + j];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int arow = (j & 1) * 4 + (j >> 1);
                        double sa1 = 0.0, sb1 = 0.0, sd1 = 0.0, sg1 = 0.0;
                        if (SPIN) {
                            const double* q4 = Cs2 + (size_t)(sbc * KC + grp * 8 + j) * 4;
                            sa1 = q4[0]; sb1 = q4[1]; sd1 = q4[2]; sg1 = q4[3];
                        }
#pragma unroll
                        for (int u = 0; u < RPL; u++) {
                            if (SPIN == 0) {
                                Ab[u][arow * ALD + lane] = live[u] ? p_cur[u] : 0.0;
                            } else {
                                const double x1 = -(sa1 * is2[u] + sb1) * p_cur[u] + sg1 * cs2[u] * p_prev[u];
                                const double x2 = -sd1 * cs2[u] * p_cur[u] + mm * sg1 * is2[u] * p_prev[u];
                                Ab[u][arow * ALD + lane] = live[u] ? x1 : 0.0;
                                Ab[u][(8 + arow) * ALD + lane] = live[u] ? x2 : 0.0;
                            }
                            const double pn = fma(cf[j].x * x[u], p_cur[u], -cf[j].y * p_prev[u]);
                            p_prev[u] = p_cur[u];
                            p_cur[u] = pn;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < RPL; u++) {
                        if (e[u] < 0 && ((__double2hiint(p_cur[u]) >> 20) & 0x7ff) > 1023 + 128) {
                            p_cur[u] *= 0x1p-256;
                            p_prev[u] *= 0x1p-256;
                            e[u] += 256;
                        }
                        const unsigned bal = __ballot_sync(0xffffffffu, live[u]);
                        if (lane == 0) balA[((p * RPL + u) * NSA + sa) * GPC + grp] = bal;
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(fullA + p * NSA + sa);
                    mbar_arrive(emptyB + sbc);
                }
            }
        }
        while (b_it < nitems) service_B();   // (nothing left in practice: a chunk is consumed before its groups end)
    } else {
        // ================================================================= consumer warp: DMMA
        const int v = warp >> 1, h = warp & 1;     // ring group, column half
        const int p = v / RPL;                       // the producer warp that feeds this ring group
        const int g = lane >> 2, t = lane & 3;
        const double* Ap = As + (size_t)v * NSA * GPC * AROWS * ALD;
        // lane constants of the fragment addresses (see the group loop)
        const int a_lane = t * ALD + g;
        int b_lane[2][2];
#pragma unroll
        for (int par = 0; par < 2; par++)
#pragma unroll
            for (int q = 0; q < 2; q++)
                b_lane[par][q] = (2 * t + par) * 16 + (((4 * q + (g >> 1)) ^ ((2 * t + par) & 7)) << 1) + (g & 1);
        const int L = lmax + 1;
        const int nring_tot = 4 * P.nside - 1;
        unsigned cglob = 0;
        for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
            // (channel block fastest.  Ring block fastest -- so that the CTAs running side by side share one (m, channel
            // block) and its alm rows reach DRAM once instead of 2.3 times -- was measured SLOWER, 149 -> 164 ms at nside
            // 512 x 1024 channels: DRAM is at 8 % of its peak here, and neighbouring CTAs then finish at different times.)
            const int cb = it % P.ncb;
            const int rb = (it / P.ncb) % P.nrb;
            const int m = it / (P.ncb * P.nrb);
            const int nk = lmax - m + 1;
            const int ngroups = (nk + 7) / 8;
            double acc[2][4][4][2];
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 4; b++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[a][b][c][0] = acc[a][b][c][1] = 0.0;
            const int nchunk = (ngroups + GPC - 1) / GPC;
#pragma unroll 1
            for (int c = 0; c < nchunk; c++, cglob++) {
                const int sb = cglob % NSB, sa = cglob % NSA;
                const int ng = min(GPC, ngroups - c * GPC);
                mbar_wait(fullB + sb, (cglob / NSB) & 1);
                mbar_wait(fullA + p * NSA + sa, (cglob / NSA) & 1);
                // every ring of a group live (the common case away from the poles): no predicates
                auto group_full = [&](const double* Ac, const double* Bc) {
#pragma unroll
                    for (int par = 0; par < 2; par++) {
                        double af[4], bf[4];
#pragma unroll
                        for (int mb = 0; mb < 4; mb++) af[mb] = Ac[par * 4 * ALD + 8 * mb];
#pragma unroll
                        for (int nb = 0; nb < 4; nb++)
                            bf[nb] = Bc[(SPIN ? (2 * (nb >> 1) + h) : (2 * h + (nb >> 1))) * BBLK + b_lane[par][nb & 1]];
#pragma unroll
                        for (int mb = 0; mb < 4; mb++)
#pragma unroll
                            for (int nb = 0; nb < 4; nb++) dmma884(acc[par][mb][nb][0], acc[par][mb][nb][1], af[mb], bf[nb]);
                    }
                    if (SPIN) {
                        // X2 rows of l-parity `par` have theta-parity 1 - par.  Q columns: (-aB_im, +aB_re) from the B tile,
                        // U columns: (+aE_im, -aE_re) from the E tile -- the channel's other component, signed
#pragma unroll
                        for (int par = 0; par < 2; par++) {
                            double af[4], bf[4];
#pragma unroll
                            for (int mb = 0; mb < 4; mb++) af[mb] = Ac[(8 + par * 4) * ALD + 8 * mb];
#pragma unroll
                            for (int nb = 0; nb < 4; nb++) {
                                const double vsw = Bc[(2 * (1 - (nb >> 1)) + h) * BBLK + (b_lane[par][nb & 1] ^ 1)];
                                bf[nb] = ((nb < 2) == ((g & 1) == 0)) ? -vsw : vsw;
                            }
#pragma unroll
                            for (int mb = 0; mb < 4; mb++)
#pragma unroll
                                for (int nb = 0; nb < 4; nb++)
                                    dmma884(acc[par ^ 1][mb][nb][0], acc[par ^ 1][mb][nb][1], af[mb], bf[nb]);
                        }
                    }
                };
                // fragment addresses = lane constants (hoisted above the item loop) + group / stage offsets:
                //   A: row (par * 4 + t) of the group's tile, ring 8 mb + g
                //   B: row r = 8 grp + 2 t + par, 16-byte unit (4 (nb & 1) + g / 2) ^ (r & 7) of column block blk
                const double* Ac0 = Ap + sa * GPC * AROWS * ALD + a_lane;
                const double* Bc0 = Bs + (size_t)sb * BSTAGE;
                {
                    // whole chunk live: the four groups run unrolled, their offsets are immediates
                    const unsigned* bp = balA + (v * NSA + sa) * GPC;
                    if (SPIN == 0 && ng == GPC && (bp[0] & bp[1] & bp[2] & bp[3]) == 0xffffffffu) {   // (spin 2: the unrolled chunk spills)
#pragma unroll
                        for (int grp = 0; grp < GPC; grp++) group_full(Ac0 + grp * AROWS * ALD, Bc0 + grp * 128);
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(emptyA + p * NSA + sa);
                            mbar_arrive(emptyB + sb);
                        }
                        continue;
                    }
                }
#pragma unroll 1
                for (int grp = 0; grp < ng; grp++) {
                    const unsigned bal = balA[(v * NSA + sa) * GPC + grp];
                    if (!bal) continue;
                    const double* Ac = Ac0 + grp * AROWS * ALD;
                    const double* Bc = Bc0 + grp * 128;
                    if (bal == 0xffffffffu) { group_full(Ac, Bc); continue; }
                    int pm[4];
#pragma unroll
                    for (int mb = 0; mb < 4; mb++) pm[mb] = (int)((bal >> (8 * mb)) & 0xffu);
#pragma unroll
                    for (int par = 0; par < 2; par++) {
                        double af[4], bf[4];
#pragma unroll
                        for (int mb = 0; mb < 4; mb++) af[mb] = Ac[par * 4 * ALD + 8 * mb];
#pragma unroll
                        for (int nb = 0; nb < 4; nb++)
                            bf[nb] = Bc[(SPIN ? (2 * (nb >> 1) + h) : (2 * h + (nb >> 1))) * BBLK + b_lane[par][nb & 1]];
#pragma unroll
                        for (int mb = 0; mb < 4; mb++)
#pragma unroll
                            for (int nb = 0; nb < 4; nb++)
                                dmma884_p(acc[par][mb][nb][0], acc[par][mb][nb][1], af[mb], bf[nb], pm[mb]);
                    }
                    if (SPIN) {
#pragma unroll
                        for (int par = 0; par < 2; par++) {
                            double af[4], bf[4];
#pragma unroll
                            for (int mb = 0; mb < 4; mb++) af[mb] = Ac[(8 + par * 4) * ALD + 8 * mb];
#pragma unroll
                            for (int nb = 0; nb < 4; nb++) {
                                const double vsw = Bc[(2 * (1 - (nb >> 1)) + h) * BBLK + (b_lane[par][nb & 1] ^ 1)];
                                bf[nb] = ((nb < 2) == ((g & 1) == 0)) ? -vsw : vsw;
                            }
#pragma unroll
                            for (int mb = 0; mb < 4; mb++)
#pragma unroll
                                for (int nb = 0; nb < 4; nb++)
                                    dmma884_p(acc[par ^ 1][mb][nb][0], acc[par ^ 1][mb][nb][1], af[mb], bf[nb], pm[mb]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(emptyA + p * NSA + sa);
                    mbar_arrive(emptyB + sb);
                }
            }
            // ---- epilogue: north = even + odd, south = even - odd;  F[ring][cg][m][4]
#pragma unroll
            for (int mb = 0; mb < 4; mb++) {
                const int rr = rb * RT + ((mb * NVP + v) << 3) + g;   // north ring index
                if (rr >= P.nrn) continue;
                const int r_s = nring_tot - 1 - rr;                      // mirror ring
#pragma unroll
                for (int nb = 0; nb < 4; nb++) {
                    const double er = acc[0][mb][nb][0], ei = acc[0][mb][nb][1];
                    const double orr = acc[1][mb][nb][0], oi = acc[1][mb][nb][1];
                    if (SPIN == 0) {
                        const int ch = cb * (NCOL / 2) + 16 * h + 4 * nb + t;
                        if (ch >= P.nb) continue;
                        const int cg = ch >> 2, cq = ch & 3;
                        P.F[(((long long)rr * P.ncg + cg) * L + m) * 4 + cq] = make_double2(er + orr, ei + oi);
                        if (r_s != rr) P.F[(((long long)r_s * P.ncg + cg) * L + m) * 4 + cq] = make_double2(er - orr, ei - oi);
                    } else {
                        const int ch = cb * (NCOL / 4) + 8 * h + 4 * (nb & 1) + t;
                        if (ch >= P.nb) continue;
                        double2* Fo = (nb >> 1) ? P.F2 : P.F;             // column blocks 0,1: Q; 2,3: U
                        const int cg = ch >> 2, cq = ch & 3;
                        Fo[(((long long)rr * P.ncg + cg) * L + m) * 4 + cq] = make_double2(-(er + orr), -(ei + oi));
                        if (r_s != rr) Fo[(((long long)r_s * P.ncg + cg) * L + m) * 4 + cq] = make_double2(-(er - orr), -(ei - oi));
                    }
                }
            }
        }
    }
}

// --------------------------------------------------------------------------- phase stage
// F layout: [ring][channel group of 4][m][4] -- the 64 bytes the Legendre epilogue writes per
// (ring, m) are the 64 bytes one phase CTA reads, and a CTA's whole input is one contiguous block.
struct PhaseParams {
    const double2* F;        // [nring][ncg][L][4]
    double* map;             // [nchan][npix] (pointer already offset to the batch's first channel)
    long long npix;
    const RingDesc* rings;
    const int* ring_list;
    const double2* tw;
    const double2* chirp;
    const double2* bhat;
    const double2* shph;      // e^{i pi k / nph} per Bluestein ring (same offsets as chirp)
    const long long *chirp_off, *bhat_off;
    int lmax, nb, ncg, P, Mmax, log_tw;
};

__device__ __forceinline__ int bitrev(int v, int bits) { return (int)(__brev((unsigned)v) >> (32 - bits)); }

// shared-memory FFT buffers are padded by one element every 8: every fused pass below then reads
// and writes conflict-free per quarter-warp, whatever its stride
__device__ __forceinline__ int pidx(int i) { return i + (i >> 3); }

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
template <int SIGN> __device__ __forceinline__ double2 ld_tw(const double2* __restrict__ tw, int j) {
    double2 w = tw[j];
    if (SIGN < 0) w.y = -w.y;
    return w;
}
// multiply by (SIGN i)
template <int SIGN> __device__ __forceinline__ double2 mul_i(double2 a) {
    return SIGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}
// multiply by exp(SIGN 2 pi i q / 8), q = 0..3
template <int SIGN, int Q> __device__ __forceinline__ double2 mul_c8(double2 a) {
    const double r = 0.70710678118654752440;
    if (Q == 0) return a;
    if (Q == 2) return mul_i<SIGN>(a);
    if (Q == 1) return SIGN > 0 ? make_double2(r * (a.x - a.y), r * (a.x + a.y)) : make_double2(r * (a.x + a.y), r * (a.y - a.x));
    return SIGN > 0 ? make_double2(-r * (a.x + a.y), r * (a.x - a.y)) : make_double2(r * (a.y - a.x), -r * (a.x + a.y));
}

// Radix-2 decimation-in-frequency stages s, s-1, .. fused in registers (natural in -> bit-reversed
// out, in place; identical data flow to one radix-2 stage after the other).  UNIT: all twiddles are 1
// (the pass that covers stages 2, 1, 0 / 1, 0 / 0), so the multiplications by them are dropped.
template <int SIGN, bool UNIT>
__device__ __forceinline__ void dif8_core(double2 v[8], double2 w1, double2 w2, double2 w4) {
    {   // stage s: pairs (j, j+4), twiddle w1 c8^j
        double2 d0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
        v[0] = cadd(v[0], v[4]); v[1] = cadd(v[1], v[5]); v[2] = cadd(v[2], v[6]); v[3] = cadd(v[3], v[7]);
        if (UNIT) {
            v[4] = d0; v[5] = mul_c8<SIGN, 1>(d1); v[6] = mul_c8<SIGN, 2>(d2); v[7] = mul_c8<SIGN, 3>(d3);
        } else {
            v[4] = cmul(d0, w1);
            v[5] = cmul(mul_c8<SIGN, 1>(d1), w1);
            v[6] = cmul(mul_c8<SIGN, 2>(d2), w1);
            v[7] = cmul(mul_c8<SIGN, 3>(d3), w1);
        }
    }
#pragma unroll
    for (int h = 0; h < 8; h += 4) {   // stage s-1: pairs (j, j+2), twiddle w2 (SIGN i)^j
        double2 d0 = csub(v[h], v[h + 2]), d1 = csub(v[h + 1], v[h + 3]);
        v[h] = cadd(v[h], v[h + 2]); v[h + 1] = cadd(v[h + 1], v[h + 3]);
        if (UNIT) { v[h + 2] = d0; v[h + 3] = mul_i<SIGN>(d1); }
        else { v[h + 2] = cmul(d0, w2); v[h + 3] = cmul(mul_i<SIGN>(d1), w2); }
    }
#pragma unroll
    for (int h = 0; h < 8; h += 2) {   // stage s-2: pairs (j, j+1), twiddle w4
        double2 d = csub(v[h], v[h + 1]);
        v[h] = cadd(v[h], v[h + 1]);
        v[h + 1] = UNIT ? d : cmul(d, w4);
    }
}
template <int SIGN>
__device__ __forceinline__ void dif8(double2* x, int b, int s, const double2* __restrict__ tw, int log_tw) {
    const int S = 1 << (s - 2);
    const int low = b & (S - 1);
    const int base = ((b >> (s - 2)) << (s + 1)) | low;
    double2 v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = x[pidx(base + j * S)];
    const double2 w1 = ld_tw<SIGN>(tw, low << (log_tw - 1 - s));
    const double2 w2 = ld_tw<SIGN>(tw, low << (log_tw - s));
    const double2 w4 = ld_tw<SIGN>(tw, low << (log_tw - s + 1));
    dif8_core<SIGN, false>(v, w1, w2, w4);
#pragma unroll
    for (int j = 0; j < 8; j++) x[pidx(base + j * S)] = v[j];
}
template <int SIGN, bool UNIT>
__device__ __forceinline__ void dif4_core(double2 v[4], double2 w1, double2 w2) {
    double2 d0 = csub(v[0], v[2]), d1 = csub(v[1], v[3]);
    v[0] = cadd(v[0], v[2]); v[1] = cadd(v[1], v[3]);
    if (UNIT) { v[2] = d0; v[3] = mul_i<SIGN>(d1); }
    else { v[2] = cmul(d0, w1); v[3] = cmul(mul_i<SIGN>(d1), w1); }
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
        double2 d = csub(v[h], v[h + 1]);
        v[h] = cadd(v[h], v[h + 1]);
        v[h + 1] = UNIT ? d : cmul(d, w2);
    }
}
template <int SIGN>
__device__ __forceinline__ void dif4(double2* x, int b, int s, const double2* __restrict__ tw, int log_tw) {
    const int S = 1 << (s - 1);
    const int low = b & (S - 1);
    const int base = ((b >> (s - 1)) << (s + 1)) | low;
    double2 v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = x[pidx(base + j * S)];
    const double2 w1 = ld_tw<SIGN>(tw, low << (log_tw - 1 - s));
    const double2 w2 = ld_tw<SIGN>(tw, low << (log_tw - s));
    dif4_core<SIGN, false>(v, w1, w2);
#pragma unroll
    for (int j = 0; j < 4; j++) x[pidx(base + j * S)] = v[j];
}
template <int SIGN>
__device__ __forceinline__ void dif2(double2* x, int b, int s, const double2* __restrict__ tw, int log_tw) {
    const int S = 1 << s;
    const int low = b & (S - 1);
    const int base = ((b >> s) << (s + 1)) | low;
    const double2 u = x[pidx(base)], v = x[pidx(base + S)];
    x[pidx(base)] = cadd(u, v);
    x[pidx(base + S)] = cmul(csub(u, v), ld_tw<SIGN>(tw, low << (log_tw - 1 - s)));
}

// Decimation-in-time stages s, s+1, .. fused (bit-reversed in -> natural out, in place).
template <int SIGN, bool UNIT>
__device__ __forceinline__ void dit8_core(double2 v[8], double2 w4, double2 w2, double2 w1) {
#pragma unroll
    for (int h = 0; h < 8; h += 2) {   // stage s
        const double2 t = UNIT ? v[h + 1] : cmul(v[h + 1], w4);
        v[h + 1] = csub(v[h], t);
        v[h] = cadd(v[h], t);
    }
#pragma unroll
    for (int h = 0; h < 8; h += 4) {   // stage s+1
        const double2 t0 = UNIT ? v[h + 2] : cmul(v[h + 2], w2);
        const double2 t1 = mul_i<SIGN>(UNIT ? v[h + 3] : cmul(v[h + 3], w2));
        v[h + 2] = csub(v[h], t0); v[h] = cadd(v[h], t0);
        v[h + 3] = csub(v[h + 1], t1); v[h + 1] = cadd(v[h + 1], t1);
    }
    {   // stage s+2
        const double2 t0 = UNIT ? v[4] : cmul(v[4], w1);
        const double2 t1 = mul_c8<SIGN, 1>(UNIT ? v[5] : cmul(v[5], w1));
        const double2 t2 = mul_c8<SIGN, 2>(UNIT ? v[6] : cmul(v[6], w1));
        const double2 t3 = mul_c8<SIGN, 3>(UNIT ? v[7] : cmul(v[7], w1));
        v[4] = csub(v[0], t0); v[0] = cadd(v[0], t0);
        v[5] = csub(v[1], t1); v[1] = cadd(v[1], t1);
        v[6] = csub(v[2], t2); v[2] = cadd(v[2], t2);
        v[7] = csub(v[3], t3); v[3] = cadd(v[3], t3);
    }
}
template <int SIGN>
__device__ __forceinline__ void dit8(double2* x, int b, int s, const double2* __restrict__ tw, int log_tw) {
    const int S = 1 << s;
    const int low = b & (S - 1);
    const int base = ((b >> s) << (s + 3)) | low;
    double2 v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = x[pidx(base + j * S)];
    const double2 w4 = ld_tw<SIGN>(tw, low << (log_tw - 1 - s));
    const double2 w2 = ld_tw<SIGN>(tw, low << (log_tw - 2 - s));
    const double2 w1 = ld_tw<SIGN>(tw, low << (log_tw - 3 - s));
    dit8_core<SIGN, false>(v, w4, w2, w1);
#pragma unroll
    for (int j = 0; j < 8; j++) x[pidx(base + j * S)] = v[j];
}
template <int SIGN, bool UNIT>
__device__ __forceinline__ void dit4_core(double2 v[4], double2 w2, double2 w1) {
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
        const double2 t = UNIT ? v[h + 1] : cmul(v[h + 1], w2);
        v[h + 1] = csub(v[h], t);
        v[h] = cadd(v[h], t);
    }
    const double2 t0 = UNIT ? v[2] : cmul(v[2], w1);
    const double2 t1 = mul_i<SIGN>(UNIT ? v[3] : cmul(v[3], w1));
    v[2] = csub(v[0], t0); v[0] = cadd(v[0], t0);
    v[3] = csub(v[1], t1); v[1] = cadd(v[1], t1);
}
template <int SIGN>
__device__ __forceinline__ void dit4(double2* x, int b, int s, const double2* __restrict__ tw, int log_tw) {
    const int S = 1 << s;
    const int low = b & (S - 1);
    const int base = ((b >> s) << (s + 2)) | low;
    double2 v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = x[pidx(base + j * S)];
    const double2 w2 = ld_tw<SIGN>(tw, low << (log_tw - 1 - s));
    const double2 w1 = ld_tw<SIGN>(tw, low << (log_tw - 2 - s));
    dit4_core<SIGN, false>(v, w2, w1);
#pragma unroll
    for (int j = 0; j < 4; j++) x[pidx(base + j * S)] = v[j];
}
template <int SIGN>
__device__ __forceinline__ void dit2(double2* x, int b, int s, const double2* __restrict__ tw, int log_tw) {
    const int S = 1 << s;
    const int low = b & (S - 1);
    const int base = ((b >> s) << (s + 1)) | low;
    const double2 u = x[pidx(base)];
    const double2 t = cmul(x[pidx(base + S)], ld_tw<SIGN>(tw, low << (log_tw - 1 - s)));
    x[pidx(base)] = cadd(u, t);
    x[pidx(base + S)] = csub(u, t);
}

// P sequences of M points at xs + p * seq_stride (padded layout); whole CTA participates.
// s_stop: the DIF stops before the pass that would start at stage s_stop (-1: run every pass).
template <int SIGN>
__device__ void fft_dif_padded(double2* xs, int seq_stride, int M, int logM, int P, const double2* __restrict__ tw, int log_tw,
                               bool skip_last = false) {
    int s = logM - 1;
    const int rem = logM % 3;
    for (; s >= 2; s -= 3) {
        if (skip_last && rem == 0 && s == 2) return;
        const int lg = logM - 3, nbf = 1 << lg;
        for (int w = threadIdx.x; w < P * nbf; w += blockDim.x) {
            const int p = w >> lg;
            dif8<SIGN>(xs + (size_t)p * seq_stride, w & (nbf - 1), s, tw, log_tw);
        }
        __syncthreads();
    }
    if (skip_last) return;
    if (s == 1) {
        const int lg = logM - 2, nbf = 1 << lg;
        for (int w = threadIdx.x; w < P * nbf; w += blockDim.x) {
            const int p = w >> lg;
            dif4<SIGN>(xs + (size_t)p * seq_stride, w & (nbf - 1), 1, tw, log_tw);
        }
        __syncthreads();
    } else if (s == 0) {
        const int lg = logM - 1, nbf = 1 << lg;
        for (int w = threadIdx.x; w < P * nbf; w += blockDim.x) {
            const int p = w >> lg;
            dif2<SIGN>(xs + (size_t)p * seq_stride, w & (nbf - 1), 0, tw, log_tw);
        }
        __syncthreads();
    }
}
template <int SIGN>
__device__ void fft_dit_padded(double2* xs, int seq_stride, int M, int logM, int P, const double2* __restrict__ tw, int log_tw,
                               bool skip_first = false) {
    int s = 0;
    const int rem = logM % 3;
    if (skip_first) {
        s = (rem == 0) ? 3 : rem;
    } else if (rem == 1) {
        const int lg = logM - 1, nbf = 1 << lg;
        for (int w = threadIdx.x; w < P * nbf; w += blockDim.x) {
            const int p = w >> lg;
            dit2<SIGN>(xs + (size_t)p * seq_stride, w & (nbf - 1), 0, tw, log_tw);
        }
        __syncthreads();
        s = 1;
    } else if (rem == 2) {
        const int lg = logM - 2, nbf = 1 << lg;
        for (int w = threadIdx.x; w < P * nbf; w += blockDim.x) {
            const int p = w >> lg;
            dit4<SIGN>(xs + (size_t)p * seq_stride, w & (nbf - 1), 0, tw, log_tw);
        }
        __syncthreads();
        s = 2;
    }
    for (; s + 2 < logM; s += 3) {
        const int lg = logM - 3, nbf = 1 << lg;
        for (int w = threadIdx.x; w < P * nbf; w += blockDim.x) {
            const int p = w >> lg;
            dit8<SIGN>(xs + (size_t)p * seq_stride, w & (nbf - 1), s, tw, log_tw);
        }
        __syncthreads();
    }
}

// Bluestein convolution  x <- IFFT_M( FFT_M(x) . bhat )  on P padded sequences.  The last DIF pass, the pointwise
// product and the first DIT pass touch the same R = 8 / 2 / 4 (logM mod 3 = 0 / 1 / 2) consecutive elements and all
// their twiddles are 1: they run as one register pass (two shared-memory round trips and two barriers fewer).
template <int R>
__device__ __forceinline__ void blu_mid(double2* x, int b, const double2* __restrict__ bh) {
    double2 v[R];
    const int base = b * R;
#pragma unroll
    for (int j = 0; j < R; j++) v[j] = x[pidx(base + j)];
    const double2 one = make_double2(1.0, 0.0);
    if (R == 8) dif8_core<-1, true>(v, one, one, one);
    else if (R == 4) dif4_core<-1, true>(v, one, one);
    else { const double2 u = v[0]; v[0] = cadd(u, v[1]); v[1] = csub(u, v[1]); }
#pragma unroll
    for (int j = 0; j < R; j++) v[j] = cmul(v[j], bh[base + j]);
    if (R == 8) dit8_core<+1, true>(v, one, one, one);
    else if (R == 4) dit4_core<+1, true>(v, one, one);
    else { const double2 u = v[0]; v[0] = cadd(u, v[1]); v[1] = csub(u, v[1]); }
#pragma unroll
    for (int j = 0; j < R; j++) x[pidx(base + j)] = v[j];
}
__device__ void bluestein_conv_padded(double2* xs, int seq_stride, int M, int logM, int P, const double2* __restrict__ tw, int log_tw,
                                      const double2* __restrict__ bh) {
    fft_dif_padded<-1>(xs, seq_stride, M, logM, P, tw, log_tw, true);
    const int rem = logM % 3;
    const int lgr = (rem == 0) ? 3 : rem;          // log2 R
    const int lg = logM - lgr, nbf = 1 << lg;
    for (int w = threadIdx.x; w < P * nbf; w += blockDim.x) {
        double2* x = xs + (size_t)(w >> lg) * seq_stride;
        const int b = w & (nbf - 1);
        if (rem == 0) blu_mid<8>(x, b, bh);
        else if (rem == 1) blu_mid<2>(x, b, bh);
        else blu_mid<4>(x, b, bh);
    }
    __syncthreads();
    fft_dit_padded<+1>(xs, seq_stride, M, logM, P, tw, log_tw, true);
}

constexpr int PH_MAXSLICE_ITEMS = 512;   // fold work items (bin, pair, slice) staged for the slice reduction

// One fold work item: the aliased copies q = q0, q0 + qstep, .. of bin k (m = k + q n) and of its twin n - k, for the
// channel pair (c, c + 1) of the group.  The bin and its twin (n - k >= k, so its copies are a subset of the bin's) are
// walked together: the four F loads of an iteration are independent and in flight at once.
__device__ __forceinline__ void fold_bin(const double2* __restrict__ Fr, int c, bool has2, int k, int kk, bool twin, int n, int lmax,
                                         bool shifted, int q0, int qstep, double2& a1, double2& a2, double2& b1, double2& b2) {
    for (int q = q0; k + q * n <= lmax; q += qstep) {
        const int m = k + q * n, m2 = kk + q * n;
        const bool v2 = twin && (m2 <= lmax);
        double2 f1 = Fr[m * 4 + c];
        double2 f2 = has2 ? Fr[m * 4 + c + 1] : make_double2(0, 0);
        const double2 g1 = v2 ? Fr[m2 * 4 + c] : make_double2(0, 0);
        const double2 g2 = (v2 && has2) ? Fr[m2 * 4 + c + 1] : make_double2(0, 0);
        if (m == 0) { f1.y = 0.0; f2.y = 0.0; }
        double sg = (m == 0) ? 1.0 : 2.0;
        if (shifted && (q & 1)) sg = -sg;   // (-1)^q of the aliased copy on shifted rings
        a1.x += sg * f1.x; a1.y += sg * f1.y;
        a2.x += sg * f2.x; a2.y += sg * f2.y;
        if (v2) {
            const double sg2 = (shifted && (q & 1)) ? -2.0 : 2.0;
            b1.x += sg2 * g1.x; b1.y += sg2 * g1.y;
            b2.x += sg2 * g2.x; b2.y += sg2 * g2.y;
        }
    }
}

// PP: channel pairs per CTA (1 or 2); BLU: the class holds Bluestein rings (else power-of-two rings).
template <int THREADS, int PP, bool BLU>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? (PP == 1 ? 3 : 4) : 1) sht_phase_kernel(PhaseParams Q) {
    extern __shared__ __align__(16) double2 xs[];   // [PP][pidx(Mmax)] + slice partials
    const int r = Q.ring_list[blockIdx.x];
    const RingDesc rd = Q.rings[r];
    const int n = rd.nph, logM = rd.logM, M = 1 << logM;
    const int seq = pidx(Q.Mmax) + 1;               // padded sequence stride
    double2* part = xs + (size_t)PP * seq;          // [PH_MAXSLICE_ITEMS][4]
    constexpr int gpc = 2 / PP;                     // CTAs per channel group
    const int cg = blockIdx.y / gpc;
    const int pair0 = (blockIdx.y % gpc) * PP;      // first pair (of the group's two) handled here
    const int L = Q.lmax + 1;
    const double2* Fr = Q.F + (((long long)r * Q.ncg + cg) * L) * 4;
    const int half = n >> 1;
    const double2* chirp = BLU ? Q.chirp + Q.chirp_off[rd.cap] : nullptr;
    const double2* shph = BLU ? Q.shph + Q.chirp_off[rd.cap] : nullptr;   // e^{i pi k / n}, k <= n/2
    const bool shifted = rd.shifted != 0;
    const bool tw_phase = !BLU && (logM + 1 <= Q.log_tw);                  // e^{i pi k/n} = tw[k * tw_n / (2n)]

    if (BLU) {
#pragma unroll
        for (int p = 0; p < PP; p++)
            for (int k = n + threadIdx.x; k < M; k += THREADS) xs[(size_t)p * seq + pidx(k)] = make_double2(0.0, 0.0);
    }

    // bin sums -> the two points (k, n - k) of the pair's complex sequence
    auto finish = [&](int k, int kk, bool twin, int pair, double2 a1, double2 a2, double2 b1, double2 b2) {
        if (twin) {
            if (shifted) {
                // e^{i pi k/n} on bin k;  e^{i pi (n-k)/n} = -conj(e^{i pi k/n}) on bin n-k
                double2 ph;
                if (BLU) ph = shph[k];
                else if (tw_phase) ph = Q.tw[k << (Q.log_tw - 1 - logM)];
                else { double sn, cs; sincospi((double)k / (double)n, &sn, &cs); ph = make_double2(cs, sn); }
                const double2 phb = make_double2(-ph.x, ph.y);
                a1 = cmul(a1, ph); a2 = cmul(a2, ph);
                b1 = cmul(b1, phb); b2 = cmul(b2, phb);
            }
        } else {
            if (shifted && k != 0) {   // k == n/2: phase e^{i pi/2} = i
                a1 = make_double2(-a1.y, a1.x);
                a2 = make_double2(-a2.y, a2.x);
            }
            b1 = a1; b2 = a2;
        }
        // H = (G_k + conj(G_{n-k}))/2
        const double2 h1 = make_double2(0.5 * (a1.x + b1.x), 0.5 * (a1.y - b1.y));
        const double2 h2 = make_double2(0.5 * (a2.x + b2.x), 0.5 * (a2.y - b2.y));
        const double2 zk = make_double2(h1.x - h2.y, h1.y + h2.x);
        const double2 zn = make_double2(h1.x + h2.y, -h1.y + h2.x);
        double2* xp = xs + (size_t)pair * seq;
        if (BLU) {
            xp[pidx(k)] = cmul(zk, chirp[k]);
            if (twin) xp[pidx(kk)] = cmul(zn, chirp[kk]);
        } else {
            xp[pidx(bitrev(k, logM))] = zk;
            if (twin) xp[pidx(bitrev(kk, logM))] = zn;
        }
    };

    // ---- fold m -> m mod nph.  Work item = (bin k <= n/2, pair, slice): a slice sums every
    // nsl-th aliased m of its bin; short rings get several slices per bin so the CTA stays busy.
    const int nbin = (half + 1) * PP;
    int nsl = 1;
    while (nsl * 2 * nbin <= min(THREADS, PH_MAXSLICE_ITEMS) && nsl * 2 * n <= L) nsl *= 2;
    if (nsl == 1) {
        for (int w = threadIdx.x; w < nbin; w += THREADS) {
            const int pair = w % PP, k = w / PP;                   // PP is a compile-time 1 or 2
            const int c = 2 * (pair0 + pair);                      // channel inside the group of 4
            const int ch = cg * 4 + c;
            if (ch >= Q.nb) continue;
            const int kk = n - k;
            const bool twin = (k != 0 && kk != k);
            double2 a1 = make_double2(0, 0), a2 = a1, b1 = a1, b2 = a1;
            fold_bin(Fr, c, ch + 1 < Q.nb, k, kk, twin, n, Q.lmax, shifted, 0, 1, a1, a2, b1, b2);
            finish(k, kk, twin, pair, a1, a2, b1, b2);
        }
    } else {
        for (int w0 = 0; w0 < nbin * nsl; w0 += THREADS) {
            const int w = w0 + threadIdx.x;
            const bool act = w < nbin * nsl;
            const int sl = act ? w / nbin : 0;
            const int wb = act ? w - sl * nbin : 0;
            const int pair = wb % PP, k = wb / PP;
            const int c = 2 * (pair0 + pair);
            const int ch = cg * 4 + c;
            const bool has1 = act && ch < Q.nb;
            const int kk = n - k;
            const bool twin = (k != 0 && kk != k);
            double2 a1 = make_double2(0, 0), a2 = a1, b1 = a1, b2 = a1;
            if (has1) fold_bin(Fr, c, ch + 1 < Q.nb, k, kk, twin, n, Q.lmax, shifted, sl, nsl, a1, a2, b1, b2);
            // fixed-order reduction over the slices (deterministic)
            if (act) {
                double2* pp = part + (size_t)w * 4;
                pp[0] = a1; pp[1] = a2; pp[2] = b1; pp[3] = b2;
            }
            __syncthreads();
            if (act && sl == 0) {
                for (int s2 = 1; s2 < nsl; s2++) {
                    const double2* pp = part + (size_t)(s2 * nbin + wb) * 4;
                    a1 = cadd(a1, pp[0]); a2 = cadd(a2, pp[1]); b1 = cadd(b1, pp[2]); b2 = cadd(b2, pp[3]);
                }
                finish(k, kk, twin, pair, a1, a2, b1, b2);
            }
            __syncthreads();   // partials consumed before the next round overwrites them
        }
    }
    __syncthreads();

    if (BLU) bluestein_conv_padded(xs, seq, M, logM, PP, Q.tw, Q.log_tw, Q.bhat + Q.bhat_off[rd.cap]);
    else fft_dit_padded<+1>(xs, seq, M, logM, PP, Q.tw, Q.log_tw);

    // ---- store: real part -> first channel of the pair, imaginary part -> second
#pragma unroll
    for (int pair = 0; pair < PP; pair++) {
        const int ch = cg * 4 + 2 * (pair0 + pair);
        if (ch >= Q.nb) continue;
        double* o1 = Q.map + (long long)ch * Q.npix + rd.start;
        double* o2 = (ch + 1 < Q.nb) ? o1 + Q.npix : nullptr;
        const double2* xp = xs + (size_t)pair * seq;
        for (int j = threadIdx.x; j < n; j += THREADS) {
            double2 v = xp[pidx(j)];
            if (BLU) v = cmul(v, chirp[j]);
            o1[j] = v.x;
            if (o2) o2[j] = v.y;
        }
    }
}


// =============================================================================== analysis
// Forward transform (map -> alm), the adjoint of the two synthesis stages above times the HEALPix
// quadrature weight 4 pi / npix:
//   1. sht_ring_analysis_kernel: per ring, two channels packed into one complex sequence, the
//      same shared-memory FFT / Bluestein core as the synthesis run on the conjugated input
//      (forward DFT = conj of the backward DFT of the conjugate), un-aliased onto every m <= lmax
//      with the ring's phase e^{-i m phi0} and weight.
//   2. sht_legendre_adj_kernel: a_lm(chan) = sum_rings lambda_lm(theta_ring) F_m(ring, chan) on
//      the FP64 tensor cores; lambda from the same scaled recurrence, transposed roles (rows = l,
//      contraction over rings), north/south folded by l - m parity before the contraction.
// replaces healpy.map2alm as called by cora/util/hputil.py:228-230 (sphtrans_real).
struct AnaParams {
    const double* map;       // [nchan][npix] (pointer already offset to the batch's first channel)
    long long npix;
    double2* F;              // [nring][ncg][L][4]
    const RingDesc* rings;
    const int* ring_list;
    const double2* tw;
    const double2* chirp;
    const double2* bhat;
    const long long *chirp_off, *bhat_off;
    const double* wring;     // [2 nside] absolute ring weights (north rings incl. equator) or null
    double wscale;           // 4 pi / npix
    int nside, lmax, nb, ncg, P, Mmax, log_tw;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 4 : 1) sht_ring_analysis_kernel(AnaParams Q) {
    extern __shared__ __align__(16) double2 xs[];   // [P][pidx(Mmax)]
    const int r = Q.ring_list[blockIdx.x];
    const RingDesc rd = Q.rings[r];
    const int n = rd.nph, logM = rd.logM, M = 1 << logM, Pp = Q.P;
    const bool blu = rd.bluestein != 0;
    const int seq = pidx(Q.Mmax) + 1;
    const int gpc = 2 / Pp;
    const int cg = blockIdx.y / gpc;
    const int pair0 = (blockIdx.y % gpc) * Pp;
    const int L = Q.lmax + 1;
    const double2* chirp = blu ? Q.chirp + Q.chirp_off[rd.cap] : nullptr;

    // ---- load conj(z), z = f1 + i f2 (two channels per complex sequence)
    for (int w = threadIdx.x; w < Pp * M; w += blockDim.x) {
        const int p = w >> logM, j = w & (M - 1);
        const int ch = cg * 4 + 2 * (pair0 + p);
        double2 v = make_double2(0.0, 0.0);
        if (j < n && ch < Q.nb) {
            v.x = Q.map[(long long)ch * Q.npix + rd.start + j];
            if (ch + 1 < Q.nb) v.y = -Q.map[(long long)(ch + 1) * Q.npix + rd.start + j];
        }
        double2* xp = xs + (size_t)p * seq;
        if (blu) xp[pidx(j)] = (j < n) ? cmul(v, chirp[j]) : v;
        else xp[pidx(bitrev(j, logM))] = v;
    }
    __syncthreads();
    if (blu) bluestein_conv_padded(xs, seq, M, logM, Pp, Q.tw, Q.log_tw, Q.bhat + Q.bhat_off[rd.cap]);
    else fft_dit_padded<+1>(xs, seq, M, logM, Pp, Q.tw, Q.log_tw);
    // xs[k] (times chirp[k] for Bluestein) = Y_k = conj(Z_k), Z = DFT_n(z).
    const int nring = 4 * Q.nside - 1;
    const int north = min(r, nring - 1 - r);
    const double wr = Q.wscale * (Q.wring ? Q.wring[north] : 1.0);
    double2* Fr = Q.F + (((long long)r * Q.ncg + cg) * L) * 4;
    for (int w = threadIdx.x; w < L * Pp; w += blockDim.x) {
        const int m = w / Pp, p = w - m * Pp;
        const int c = 2 * (pair0 + p);
        const int ch = cg * 4 + c;
        if (ch >= Q.nb) continue;
        const int q = m / n, k = m - q * n, kk = (k == 0) ? 0 : n - k;
        const double2* xp = xs + (size_t)p * seq;
        double2 yk = xp[pidx(k)], yn = xp[pidx(kk)];
        if (blu) { yk = cmul(yk, chirp[k]); yn = cmul(yn, chirp[kk]); }
        // F1_k = (Z_k + conj(Z_{n-k}))/2 = (conj(Y_k) + Y_{n-k})/2,  F2_k = (conj(Y_k) - Y_{n-k})/(2i)
        double2 g1 = make_double2(0.5 * (yk.x + yn.x), 0.5 * (yn.y - yk.y));
        const double2 d = make_double2(0.5 * (yk.x - yn.x), 0.5 * (-yk.y - yn.y));
        double2 g2 = make_double2(d.y, -d.x);
        double2 ph = make_double2(wr, 0.0);
        if (rd.shifted) {   // e^{-i m phi0}, phi0 = pi/n:  e^{-i pi k/n} (-1)^q
            double sn, cs;
            sincospi((double)k / (double)n, &sn, &cs);
            const double sg = (q & 1) ? -wr : wr;
            ph = make_double2(sg * cs, -sg * sn);
        }
        Fr[m * 4 + c] = cmul(g1, ph);
        if (ch + 1 < Q.nb) Fr[m * 4 + c + 1] = cmul(g2, ph);
    }
}

constexpr int ADJ_THREADS = 256;
constexpr int ADJ_RPL = 1;                   // ring sets (of 32) per warp
constexpr int ADJ_RT = 32 * 8 * ADJ_RPL;     // north rings per CTA
constexpr int ADJ_NCH = 8;                   // complex channels per CTA (16 real columns)
constexpr int ADJ_BLD = 16;                  // Bs[ring][16], column c stored at c ^ 4 (ring & 3): fragment reads conflict-free
constexpr int ADJ_ALD = 36;                  // As[row][36]:  (g * 36 + t)
// shared memory: B 64 KB + A 36.9 KB (the C staging of the cross-warp reduction aliases the A tiles)
// = 101 KB -> two CTAs per SM, so one CTA's recurrence / barrier phases overlap the other's DMMAs
__device__ __forceinline__ int adj_bcol(int ring, int col) { return col ^ ((ring & 3) << 2); }

struct LegAdjParams {
    const double2* F;         // [nring][ncg][L][4]
    double2* alm;             // PANEL [nalm][alm_stride]
    long long alm_stride;
    int chan0, nb;
    const double *cth, *sth;
    const double* nm_mant;
    const int* nm_exp;
    const double2* rc;
    int nside, lmax, nrn, nrb, ncb, ncg, accumulate;
};

__global__ void __launch_bounds__(ADJ_THREADS, 2) sht_legendre_adj_kernel(LegAdjParams P) {
    extern __shared__ __align__(16) double smem[];
    double* Bs = smem;                                  // [2 parities][RT][BLD]
    double* As = Bs + 2 * ADJ_RT * ADJ_BLD;             // [8 warps][2 parities][8 l][ALD]  (576 doubles per warp)
    double* Cs = As;                                    // [8 warps][256] staged in each warp's own A tile after its DMMAs
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    int bid = blockIdx.x;
    const int cb = bid % P.ncb; bid /= P.ncb;
    const int rb = bid % P.nrb;
    const int m = bid / P.nrb;
    const int lmax = P.lmax, L = lmax + 1;
    const int nk = lmax - m + 1;
    const int nring_tot = 4 * P.nside - 1;

    // ---- B tile: (north + south, north - south) of F_m for this CTA's rings and channels
    for (int el = tid; el < ADJ_RT * ADJ_NCH; el += ADJ_THREADS) {
        const int rr = el / ADJ_NCH, c = el % ADJ_NCH;
        const int rn = rb * ADJ_RT + rr;
        const int ch = cb * ADJ_NCH + c;
        double2 ev = make_double2(0.0, 0.0), od = ev;
        if (rn < P.nrn && ch < P.nb) {
            const int cg = ch >> 2, cc = ch & 3;
            const double2 fn = P.F[(((long long)rn * P.ncg + cg) * L + m) * 4 + cc];
            const int rs = nring_tot - 1 - rn;
            if (rs != rn) {
                const double2 fs = P.F[(((long long)rs * P.ncg + cg) * L + m) * 4 + cc];
                ev = make_double2(fn.x + fs.x, fn.y + fs.y);
                od = make_double2(fn.x - fs.x, fn.y - fs.y);
            } else {
                ev = fn;   // equator: lambda vanishes for odd l - m
            }
        }
        Bs[(size_t)rr * ADJ_BLD + adj_bcol(rr, 2 * c)] = ev.x;
        Bs[(size_t)rr * ADJ_BLD + adj_bcol(rr, 2 * c + 1)] = ev.y;
        Bs[(size_t)(ADJ_RT + rr) * ADJ_BLD + adj_bcol(rr, 2 * c)] = od.x;
        Bs[(size_t)(ADJ_RT + rr) * ADJ_BLD + adj_bcol(rr, 2 * c + 1)] = od.y;
    }

    // ---- recurrence seeds: lane owns ring (set s, lane) of the warp's ADJ_RPL sets
    double x[ADJ_RPL], p_cur[ADJ_RPL], p_prev[ADJ_RPL];
    int e[ADJ_RPL];
#pragma unroll
    for (int s = 0; s < ADJ_RPL; s++) {
        const int rn = rb * ADJ_RT + (s * 8 + warp) * 32 + lane;
        x[s] = 0.0; p_cur[s] = 0.0; p_prev[s] = 0.0; e[s] = -(1 << 20);
        if (rn < P.nrn) {
            x[s] = P.cth[rn];
            double bm = P.sth[rn];
            long long be = 0;
            norm_frexp(bm, be);
            double rm = 1.0;
            long long re = 0;
            int n = m;
            while (n) {
                if (n & 1) { rm *= bm; re += be; norm_frexp(rm, re); }
                bm *= bm; be *= 2; norm_frexp(bm, be);
                n >>= 1;
            }
            rm *= P.nm_mant[m];
            re += P.nm_exp[m];
            norm_frexp(rm, re);
            if (m & 1) rm = -rm;
            long long q = (re >= 0) ? 0 : -((-re) / 256);
            e[s] = (int)(q * 256);
            p_cur[s] = ldexp(rm, (int)(re - (long long)e[s]));
        }
    }
    __syncthreads();   // B tile complete

    const long long row0 = (long long)m * (2 * lmax + 1 - m) / 2 + m;   // idx(l = m, m)
    double* Aw = As + warp * (2 * 8 * ADJ_ALD);
    const int ngroups = (nk + 15) / 16;
    for (int gi = 0; gi < ngroups; gi++) {
        const int kbase = gi * 16;
        double acc[2][2][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
#pragma unroll
        for (int s = 0; s < ADJ_RPL; s++) {
            const bool live = (e[s] == 0);
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int k = kbase + j;
                Aw[((j & 1) * 8 + (j >> 1)) * ADJ_ALD + lane] = (live && k < nk) ? p_cur[s] : 0.0;
                const double2 c = (k < nk) ? P.rc[row0 + k] : make_double2(0.0, 0.0);
                const double pn = fma(c.x * x[s], p_cur[s], -c.y * p_prev[s]);
                p_prev[s] = p_cur[s];
                p_cur[s] = pn;
                if ((j & 7) == 7 && e[s] < 0 && ((__double2hiint(p_cur[s]) >> 20) & 0x7ff) > 1023 + 128) {
                    p_cur[s] *= 0x1p-256;
                    p_prev[s] *= 0x1p-256;
                    e[s] += 256;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, live);
            __syncwarp();
            if (bal) {
                const int ring0 = (s * 8 + warp) * 32;
#pragma unroll
                for (int par = 0; par < 2; par++) {
#pragma unroll
                    for (int jb = 0; jb < 8; jb++) {
                        if (!((bal >> (4 * jb)) & 0xfu)) continue;   // four dead rings: nothing to add
                        const double af = Aw[(par * 8 + g) * ADJ_ALD + 4 * jb + t];
                        const double* brow = Bs + (size_t)(par * ADJ_RT + ring0 + 4 * jb + t) * ADJ_BLD;
                        dmma884(acc[par][0][0], acc[par][0][1], af, brow[g ^ (t << 2)]);          // ring & 3 == t
                        dmma884(acc[par][1][0], acc[par][1][1], af, brow[(8 + g) ^ (t << 2)]);
                    }
                }
            }
            __syncwarp();   // fragments read before the next set overwrites the tile
        }
        // ---- cross-warp reduction of the 16 l x 16 column tile, fixed order
#pragma unroll
        for (int par = 0; par < 2; par++)
#pragma unroll
            for (int nb = 0; nb < 2; nb++) {
                Cs[warp * 576 + ((par * 2 + nb) * 8 + g) * 8 + 2 * t] = acc[par][nb][0];
                Cs[warp * 576 + ((par * 2 + nb) * 8 + g) * 8 + 2 * t + 1] = acc[par][nb][1];
            }
        __syncthreads();
        {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) v += Cs[w * 576 + tid];
            const int c8 = tid & 7, gg = (tid >> 3) & 7, nb = (tid >> 6) & 1, par = tid >> 7;
            const int k = kbase + 2 * gg + par;
            const int col = nb * 8 + c8;
            const int ch = cb * ADJ_NCH + (col >> 1);
            if (k < nk && ch < P.nb) {
                double* dst = (double*)(P.alm + (row0 + k) * P.alm_stride + P.chan0 + ch) + (col & 1);
                if (P.nrb > 1) atomicAdd(dst, v);          // ring blocks of one (m, channel block) meet here
                else if (P.accumulate) *dst += v;
                else *dst = v;
            }
        }
        __syncthreads();   // Cs consumed
    }
}


// Spin-2 analysis: (Q, U) ring spectra -> (aE, aB), the Hermitian adjoint of sht_legendre_kernel<2>:
//   aE = -sum_r (X1 Q_m + i X2 U_m),  aB = -sum_r (X1 U_m - i X2 Q_m)      (weights are in F)
// GEMM form, 4 real columns per channel: C = X1^T B1 + X2^T B2 with B1 = -(Q_re, Q_im, U_re, U_im),
// B2 = (U_im, -U_re, -Q_im, Q_re) = +-B1 at the mirrored column of the group of four; X2 has the
// opposite theta-parity of X1, so it contracts with the other north/south combination.
constexpr int ADJ2_RT = 256;      // north rings per CTA (one set of 32 per warp)
constexpr int ADJ2_NCH = 4;       // channels per CTA (16 real columns)
constexpr int ADJ2_BLD = 20;      // Bs[ring][20]: (t * 20 + g) conflict-free per half-warp

struct LegAdj2Params {
    const double2 *FQ, *FU;   // [nring][ncg][L][4]
    double2 *almE, *almB;     // PANEL
    long long alm_stride;
    int chan0, nb;
    const double *cth, *sth;
    const double* nm_mant;
    const int* nm_exp;
    const double2* rc;
    int nside, lmax, nrn, nrb, ncb, ncg, accumulate;
};

__global__ void __launch_bounds__(ADJ_THREADS, 1) sht_legendre_adj2_kernel(LegAdj2Params P) {
    extern __shared__ __align__(16) double smem[];
    double* Bs = smem;                                   // [2 parities][RT][BLD]   (B1)
    double* As = Bs + 2 * ADJ2_RT * ADJ2_BLD;             // [8 warps][X1, X2][2 parities][8 l][ALD]
    double* Cs = As + 8 * 2 * 2 * 8 * ADJ_ALD;           // [8 warps][256]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    int bid = blockIdx.x;
    const int cb = bid % P.ncb; bid /= P.ncb;
    const int rb = bid % P.nrb;
    const int m = bid / P.nrb;
    const int lmax = P.lmax, L = lmax + 1;
    const int nk = lmax - m + 1;
    const int nring_tot = 4 * P.nside - 1;

    // ---- B1 tile
    for (int el = tid; el < ADJ2_RT * ADJ2_NCH * 2; el += ADJ_THREADS) {
        const int rr = el / (ADJ2_NCH * 2), c2 = el % (ADJ2_NCH * 2);
        const int c = c2 >> 1, isU = c2 & 1;
        const int rn = rb * ADJ2_RT + rr;
        const int ch = cb * ADJ2_NCH + c;
        double2 ev = make_double2(0.0, 0.0), od = ev;
        if (rn < P.nrn && ch < P.nb) {
            const double2* F = isU ? P.FU : P.FQ;
            const int cg = ch >> 2, cc = ch & 3;
            const double2 fn = F[(((long long)rn * P.ncg + cg) * L + m) * 4 + cc];
            const int rs = nring_tot - 1 - rn;
            if (rs != rn) {
                const double2 fs = F[(((long long)rs * P.ncg + cg) * L + m) * 4 + cc];
                ev = make_double2(fn.x + fs.x, fn.y + fs.y);
                od = make_double2(fn.x - fs.x, fn.y - fs.y);
            } else {
                // equator: X1 vanishes for odd l - m (keep the even combination), X2 for even l - m
                // (it contracts with the "odd" slot): both slots hold the ring itself
                ev = fn; od = fn;
            }
        }
        const int col = 4 * c + 2 * isU;
        Bs[(size_t)rr * ADJ2_BLD + col] = -ev.x;
        Bs[(size_t)rr * ADJ2_BLD + col + 1] = -ev.y;
        Bs[(size_t)(ADJ2_RT + rr) * ADJ2_BLD + col] = -od.x;
        Bs[(size_t)(ADJ2_RT + rr) * ADJ2_BLD + col + 1] = -od.y;
    }

    // ---- recurrence seed (one ring per lane)
    double x = 0.0, p_cur = 0.0, p_prev = 0.0, is2 = 0.0, cs2 = 0.0;
    int e = -(1 << 20);
    {
        const int rn = rb * ADJ2_RT + warp * 32 + lane;
        if (rn < P.nrn) {
            x = P.cth[rn];
            double bm = P.sth[rn];
            is2 = 1.0 / (bm * bm); cs2 = x * is2;
            long long be = 0;
            norm_frexp(bm, be);
            double rm = 1.0;
            long long re = 0;
            int n = m;
            while (n) {
                if (n & 1) { rm *= bm; re += be; norm_frexp(rm, re); }
                bm *= bm; be *= 2; norm_frexp(bm, be);
                n >>= 1;
            }
            rm *= P.nm_mant[m];
            re += P.nm_exp[m];
            norm_frexp(rm, re);
            if (m & 1) rm = -rm;
            long long q = (re >= 0) ? 0 : -((-re) / 256);
            e = (int)(q * 256);
            p_cur = ldexp(rm, (int)(re - (long long)e));
        }
    }
    __syncthreads();

    const long long row0 = (long long)m * (2 * lmax + 1 - m) / 2 + m;
    double* Aw = As + warp * (2 * 2 * 8 * ADJ_ALD);
    const int gperm = (g & 4) + 3 - (g & 3);
    const double gsign = ((g & 3) == 0 || (g & 3) == 3) ? -1.0 : 1.0;
    const double mm = (double)m;
    const int ngroups = (nk + 15) / 16;
    for (int gi = 0; gi < ngroups; gi++) {
        const int kbase = gi * 16;
        double acc[2][2][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
        // spin factors of the group's 16 l's: lane j < 16 computes those of l = m + kbase + j
        double tn_l = 0.0, tg_l = 0.0;
        {
            const double l = (double)(m + kbase + (lane & 15));
            if (kbase + (lane & 15) < nk && l >= 2.0) {
                tn_l = 2.0 / sqrt((l + 2.0) * (l + 1.0) * l * (l - 1.0));
                tg_l = tn_l * sqrt((2.0 * l + 1.0) / (2.0 * l - 1.0) * (l * l - mm * mm));
            }
        }
        const bool live = (e == 0);
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int k = kbase + j;
            const double tn = __shfl_sync(0xffffffffu, tn_l, j), tg = __shfl_sync(0xffffffffu, tg_l, j);
            const double l = (double)(m + k);
            const double al = tn * (l - mm * mm), be = 0.5 * tn * l * (l - 1.0), de = tn * mm * (l - 1.0);
            const double x1 = -(al * is2 + be) * p_cur + tg * cs2 * p_prev;
            const double x2 = -de * cs2 * p_cur + mm * tg * is2 * p_prev;
            const bool ok = live && k < nk;
            const int arow = (j & 1) * 8 + (j >> 1);
            Aw[arow * ADJ_ALD + lane] = ok ? x1 : 0.0;
            Aw[(16 + arow) * ADJ_ALD + lane] = ok ? x2 : 0.0;
            const double2 c = (k < nk) ? P.rc[row0 + k] : make_double2(0.0, 0.0);
            const double pn = fma(c.x * x, p_cur, -c.y * p_prev);
            p_prev = p_cur;
            p_cur = pn;
            if ((j & 7) == 7 && e < 0 && ((__double2hiint(p_cur) >> 20) & 0x7ff) > 1023 + 128) {
                p_cur *= 0x1p-256;
                p_prev *= 0x1p-256;
                e += 256;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, live);
        __syncwarp();
        if (bal) {
            const int ring0 = warp * 32;
#pragma unroll
            for (int par = 0; par < 2; par++) {
#pragma unroll
                for (int jb = 0; jb < 8; jb++) {
                    if (!((bal >> (4 * jb)) & 0xfu)) continue;
                    const double a1 = Aw[(par * 8 + g) * ADJ_ALD + 4 * jb + t];
                    const double a2 = Aw[(16 + par * 8 + g) * ADJ_ALD + 4 * jb + t];
                    const double* b1row = Bs + (size_t)(par * ADJ2_RT + ring0 + 4 * jb + t) * ADJ2_BLD;          // X1: same parity
                    const double* b2row = Bs + (size_t)((par ^ 1) * ADJ2_RT + ring0 + 4 * jb + t) * ADJ2_BLD;    // X2: opposite
                    dmma884(acc[par][0][0], acc[par][0][1], a1, b1row[g]);
                    dmma884(acc[par][1][0], acc[par][1][1], a1, b1row[8 + g]);
                    dmma884(acc[par][0][0], acc[par][0][1], a2, gsign * b2row[gperm]);
                    dmma884(acc[par][1][0], acc[par][1][1], a2, gsign * b2row[8 + gperm]);
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int par = 0; par < 2; par++)
#pragma unroll
            for (int nb = 0; nb < 2; nb++) {
                Cs[warp * 256 + ((par * 2 + nb) * 8 + g) * 8 + 2 * t] = acc[par][nb][0];
                Cs[warp * 256 + ((par * 2 + nb) * 8 + g) * 8 + 2 * t + 1] = acc[par][nb][1];
            }
        __syncthreads();
        {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) v += Cs[w * 256 + tid];
            const int c8 = tid & 7, gg = (tid >> 3) & 7, nb = (tid >> 6) & 1, par = tid >> 7;
            const int k = kbase + 2 * gg + par;
            const int col = nb * 8 + c8;                  // 4 * channel + (aE_re, aE_im, aB_re, aB_im)
            const int ch = cb * ADJ2_NCH + (col >> 2);
            if (k < nk && ch < P.nb) {
                double2* base = (col & 2) ? P.almB : P.almE;
                double* dst = (double*)(base + (row0 + k) * P.alm_stride + P.chan0 + ch) + (col & 1);
                if (P.nrb > 1) atomicAdd(dst, v);
                else if (P.accumulate) *dst += v;
                else *dst = v;
            }
        }
        __syncthreads();
    }
}

// out = a - b  (residual map of the Jacobi refinement)
__global__ void map_sub_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = i; e < n; e += stride) out[e] = a[e] - b[e];
}

// zero the rows of a PANEL batch (before ring blocks accumulate into it)
__global__ void panel_zero_kernel(double2* __restrict__ alm, long long nalm, long long stride, int chan0, int nb) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nalm * nb) return;
    const long long row = e / nb;
    const int c = (int)(e - row * nb);
    alm[row * stride + chan0 + c] = make_double2(0.0, 0.0);
}

}  // namespace cb

using namespace cb;

// =============================================================================== C ABI
extern "C" int cora_b200_sht_plan_create(int nside, int lmax, void** plan_out) {
    CB_REQUIRE(nside >= 1 && lmax >= 0 && plan_out, 1, "sht_plan_create: bad arguments (nside=%d lmax=%d)", nside, lmax);
    ShtPlan* pl = new ShtPlan();
    pl->nside = nside; pl->lmax = lmax; pl->nrn = 2 * nside;
    pl->npix = 12LL * nside * nside;
    pl->nalm = (long long)(lmax + 1) * (lmax + 2) / 2;
    const int nring = 4 * nside - 1;

    // ring geometry
    std::vector<double> cth(pl->nrn), sth(pl->nrn);
    pl->h_rings.resize(nring);
    const double nf = (double)nside;
    for (int i = 1; i <= 2 * nside; i++) {
        double c, s;
        if (i < nside) {
            double om = (double)i * (double)i / (3.0 * nf * nf);   // 1 - cos(theta)
            c = 1.0 - om;
            s = std::sqrt(om * (1.0 + c));
        } else {
            c = 4.0 / 3.0 - 2.0 * (double)i / (3.0 * nf);
            s = std::sqrt((1.0 - c) * (1.0 + c));
        }
        cth[i - 1] = c; sth[i - 1] = s;
    }
    for (int i = 1; i <= nring; i++) {
        int north = std::min(i, 4 * nside - i);
        bool cap = north < nside;
        RingDesc rd;
        rd.nph = cap ? 4 * north : 4 * nside;
        rd.shifted = (cap || (((north - nside) & 1) == 0)) ? 1 : 0;
        rd.cap = cap ? north : -1;
        if (i <= 2 * nside || !cap) {
            rd.start = cap ? 2LL * north * (north - 1) : 2LL * nside * (nside - 1) + 4LL * nside * (i - nside);
        } else {
            rd.start = pl->npix - 2LL * north * (north + 1);
        }
        pl->h_rings[i - 1] = rd;
    }

    // N_m = sqrt((2m+1)/(4 pi) prod_{k<=m} (2k-1)/(2k)) as mantissa/exponent, long double product
    std::vector<double> nm_mant(lmax + 1);
    std::vector<int> nm_exp(lmax + 1);
    {
        long double pm = 1.0L; long long pe = 0;
        const long double four_pi = 4.0L * 3.14159265358979323846264338327950288L;
        for (int m = 0; m <= lmax; m++) {
            if (m > 0) {
                pm *= (long double)(2 * m - 1) / (long double)(2 * m);
                int ee; pm = frexpl(pm, &ee); pe += ee;
            }
            long double v = pm * (long double)(2 * m + 1) / four_pi;
            long long ve = pe; int ee;
            v = frexpl(v, &ee); ve += ee;
            if (ve & 1) { v *= 2.0L; ve -= 1; }
            long double sq = sqrtl(v);
            long long se = ve / 2;
            sq = frexpl(sq, &ee); se += ee;
            nm_mant[m] = (double)sq; nm_exp[m] = (int)se;
        }
    }

    CB_CUDA(cudaMalloc(&pl->d_rc, sizeof(double2) * pl->nalm));
    recur_coef_kernel<<<dim3(ceil_div(lmax + 1, 128), lmax + 1), 128>>>(pl->d_rc, lmax);
    count_launch();
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMalloc(&pl->d_cth, sizeof(double) * pl->nrn));
    CB_CUDA(cudaMalloc(&pl->d_sth, sizeof(double) * pl->nrn));
    CB_CUDA(cudaMalloc(&pl->d_nm_mant, sizeof(double) * (lmax + 1)));
    CB_CUDA(cudaMalloc(&pl->d_nm_exp, sizeof(int) * (lmax + 1)));
    CB_CUDA(cudaMalloc(&pl->d_rings, sizeof(RingDesc) * nring));
    CB_CUDA(cudaMemcpy(pl->d_cth, cth.data(), sizeof(double) * pl->nrn, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(pl->d_sth, sth.data(), sizeof(double) * pl->nrn, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(pl->d_nm_mant, nm_mant.data(), sizeof(double) * (lmax + 1), cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(pl->d_nm_exp, nm_exp.data(), sizeof(int) * (lmax + 1), cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(pl->d_rings, pl->h_rings.data(), sizeof(RingDesc) * nring, cudaMemcpyHostToDevice));

    // FFT classes
    int maxM = 4;
    std::vector<long long> chirp_off(nside + 1, 0), bhat_off(nside + 1, 0);
    long long nchirp = 0, nbhat = 0;
    for (int i = 1; i < nside; i++) {
        int n = 4 * i;
        if (is_pow2(n)) { maxM = std::max(maxM, n); continue; }
        int M = 1; while (M < 2 * n - 1) M <<= 1;
        chirp_off[i] = nchirp; nchirp += n;
        bhat_off[i] = nbhat; nbhat += M;
        maxM = std::max(maxM, M);
    }
    maxM = std::max(maxM, 4 * nside);
    int belt_M = 4 * nside;
    bool belt_pow2 = is_pow2(belt_M);
    // a non-power-of-two nside makes the belt rings Bluestein rings too: give them table slot `nside`
    if (!belt_pow2) {
        int n = belt_M; int M = 1; while (M < 2 * n - 1) M <<= 1;
        chirp_off[nside] = nchirp; nchirp += n;
        bhat_off[nside] = nbhat; nbhat += M;
        maxM = std::max(maxM, M);
        for (auto& rd : pl->h_rings) if (rd.cap < 0) rd.cap = nside;
        CB_CUDA(cudaMemcpy(pl->d_rings, pl->h_rings.data(), sizeof(RingDesc) * nring, cudaMemcpyHostToDevice));
    }
    CB_REQUIRE(maxM <= 8192, 2, "sht_plan_create: ring FFT size %d exceeds the 8192-point shared-memory FFT (nside too large)", maxM);
    pl->tw_n = maxM; pl->log_tw = ilog2(maxM);
    CB_CUDA(cudaMalloc(&pl->d_tw, sizeof(double2) * std::max(1, maxM / 2)));
    twiddle_kernel<<<ceil_div(maxM / 2, 256), 256>>>(pl->d_tw, maxM); count_launch();
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaMalloc(&pl->d_chirp, sizeof(double2) * std::max(1LL, nchirp)));
    CB_CUDA(cudaMalloc(&pl->d_shph, sizeof(double2) * std::max(1LL, nchirp)));
    CB_CUDA(cudaMalloc(&pl->d_bhat, sizeof(double2) * std::max(1LL, nbhat)));
    CB_CUDA(cudaMalloc(&pl->d_chirp_off, sizeof(long long) * (nside + 1)));
    CB_CUDA(cudaMalloc(&pl->d_bhat_off, sizeof(long long) * (nside + 1)));
    CB_CUDA(cudaMemcpy(pl->d_chirp_off, chirp_off.data(), sizeof(long long) * (nside + 1), cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(pl->d_bhat_off, bhat_off.data(), sizeof(long long) * (nside + 1), cudaMemcpyHostToDevice));
    if (nbhat > 0) {
        CB_CUDA(cudaFuncSetAttribute(bluestein_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
        bluestein_setup_kernel<<<nside, 256, (size_t)maxM * 16>>>(nside, pl->d_chirp_off, pl->d_bhat_off, pl->d_chirp,
                                                                  pl->d_bhat, pl->d_shph, pl->d_tw, pl->log_tw);
        count_launch();
        CB_LAUNCH_CHECK();
    }
    CB_CUDA(cudaFuncSetAttribute(sht_phase_kernel<256, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CB_CUDA(cudaFuncSetAttribute(sht_phase_kernel<256, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CB_CUDA(cudaFuncSetAttribute(sht_phase_kernel<512, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CB_CUDA(cudaFuncSetAttribute(sht_phase_kernel<512, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CB_CUDA(cudaFuncSetAttribute(sht_phase_kernel<512, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CB_CUDA(cudaFuncSetAttribute(sht_phase_kernel<512, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CB_CUDA(cudaFuncSetAttribute(sht_phase_kernel<256, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CB_CUDA(cudaFuncSetAttribute(sht_phase_kernel<256, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

    // per-ring FFT size; rings launched together by shared-memory footprint, longest first
    for (auto& rd : pl->h_rings) {
        const int n = rd.nph;
        rd.bluestein = is_pow2(n) ? 0 : 1;
        int M = n; if (rd.bluestein) { M = 1; while (M < 2 * n - 1) M <<= 1; }
        rd.logM = ilog2(M);
    }
    CB_CUDA(cudaMemcpy(pl->d_rings, pl->h_rings.data(), sizeof(RingDesc) * nring, cudaMemcpyHostToDevice));
    {
        // classes: (shared-memory FFT size, Bluestein or power-of-two ring) -- one kernel instantiation each
        std::vector<std::vector<int>> bycls(32);
        for (int r = 0; r < nring; r++) bycls[2 * std::max(9, pl->h_rings[r].logM) + pl->h_rings[r].bluestein].push_back(r);
        for (int ci = 31; ci >= 18; ci--) {
            auto& v = bycls[ci];
            if (v.empty()) continue;
            const int lg = ci / 2;
            std::stable_sort(v.begin(), v.end(), [&](int x, int y) { return pl->h_rings[x].nph > pl->h_rings[y].nph; });
            PhaseClass pc;
            pc.blu = ci & 1;
            pc.Mmax = 1 << lg;
            pc.P = (pc.Mmax <= 4096) ? 2 : 1;          // 2 pairs x 4096 points (padded) = 147 KB
            pc.threads = (pc.Mmax >= 4096) ? 512 : 256;
            if (pc.Mmax == 4096 && g_phase_4096_split) { pc.P = 1; pc.threads = 256; }   // 74 KB: 3 CTAs of 256 threads per SM
            pc.nrings = (int)v.size();
            CB_CUDA(cudaMalloc(&pc.d_rings, sizeof(int) * v.size()));
            CB_CUDA(cudaMemcpy(pc.d_rings, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice));
            pl->classes.push_back(pc);
        }
    }
    CB_CUDA(cudaDeviceSynchronize());
    *plan_out = pl;
    return 0;
}

extern "C" int cora_b200_sht_plan_destroy(void* plan) {
    if (!plan) return 0;
    ShtPlan* pl = (ShtPlan*)plan;
    cudaFree(pl->d_rc);
    if (pl->d_rc2) cudaFree(pl->d_rc2);
    cudaFree(pl->d_cth); cudaFree(pl->d_sth); cudaFree(pl->d_nm_mant); cudaFree(pl->d_nm_exp);
    cudaFree(pl->d_rings); cudaFree(pl->d_tw); cudaFree(pl->d_chirp); cudaFree(pl->d_shph); cudaFree(pl->d_bhat);
    cudaFree(pl->d_chirp_off); cudaFree(pl->d_bhat_off);
    for (auto& pc : pl->classes) cudaFree(pc.d_rings);
    delete pl;
    return 0;
}

// bytes of workspace needed per channel of a batch
static long long ws_per_chan(const ShtPlan* pl, int layout) {
    long long f = (long long)(4 * pl->nside - 1) * (pl->lmax + 1) * 16;
    long long a = (layout == CORA_B200_ALM_PACKED) ? pl->nalm * 16 : 0;
    return f + a;
}

static long long round4(long long n) { return (n + 3) & ~3LL; }

extern "C" long long cora_b200_alm2map_workspace_bytes(void* plan, int layout, int nchan_batch) {
    if (!plan) return -1;
    return ws_per_chan((ShtPlan*)plan, layout) * round4(nchan_batch) + 256;
}

// Bit 0: warp-specialised scalar Legendre kernel, bit 1: warp-specialised spin-2 kernel (default 3: both); cleared bits
// fall back to the single-role kernel.  Switchable (CORA_B200_LEGENDRE_WS=0|1|2|3) so the kernels can be compared on
// the device.
static int g_legendre_ws = [] { const char* e = getenv("CORA_B200_LEGENDRE_WS"); return (e && e[0] >= '0' && e[0] <= '3') ? (e[0] - '0') : 3; }();

extern "C" int cora_b200_set_legendre_ws(int mask) {
    const int old = g_legendre_ws;
    if (mask >= 0) g_legendre_ws = mask & 3;
    return old;
}

// tensor map over a batch's alm panel columns: [nalm rows][2 nb doubles], row pitch = panel stride; columns past
// the batch are out of bounds -> zero-filled by the TMA unit
static int make_alm_tmap(CUtensorMap* tmap, const double2* base, int nb, long long nalm, long long alm_stride) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        CB_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, 5, "alm2map: cuTensorMapEncodeTiled is not available");
        encode = (EncodeFn)fn;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)nb * 2, (cuuint64_t)nalm};
    const cuuint64_t gstr[1] = {(cuuint64_t)alm_stride * 16};
    const cuuint32_t box[2] = {16, (cuuint32_t)lws::KC};
    const cuuint32_t estr[2] = {1, 1};
    CB_REQUIRE(((uintptr_t)base % 16) == 0, 5, "alm2map: alm panel must be 16-byte aligned");
    CUresult cr = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CB_REQUIRE(cr == CUDA_SUCCESS, 5, "alm2map: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return 0;
}

template <int SPIN>
static int run_legendre(ShtPlan* pl, const double2* almT, const double2* almB, long long alm_stride, int chan0, int nb,
                        double2* F, double2* F2, cudaStream_t st) {
    LegParams P;
    P.almT = almT; P.almB = almB; P.alm_stride = alm_stride; P.chan0 = chan0; P.nb = nb; P.F = F; P.F2 = F2;
    P.cth = pl->d_cth; P.sth = pl->d_sth; P.nm_mant = pl->d_nm_mant; P.nm_exp = pl->d_nm_exp; P.rc = pl->d_rc;
    P.rc2 = nullptr;
    P.nside = pl->nside; P.lmax = pl->lmax; P.nrn = pl->nrn;
    P.nrb = ceil_div(pl->nrn, LEG_RT);
    P.ncb = ceil_div(nb, SPIN ? LEG_NCH / 2 : LEG_NCH);
    P.ncg = ceil_div(nb, 4);
    P.Lpad = ((pl->lmax + 1 + 7) / 8 + 2) * 8 + 8;   // recur8 runs one group past the end
    if (g_legendre_ws & (SPIN ? 2 : 1)) {
        using namespace lws;
        constexpr int NSA = Cfg<SPIN>::NSA, AROWS = Cfg<SPIN>::AROWS;
        P.nrb = ceil_div(pl->nrn, RT);
        P.ncb = ceil_div(nb, SPIN ? NCOL / 4 : NCOL / 2);
        if (SPIN && !pl->d_rc2) {
            CB_CUDA(cudaMalloc(&pl->d_rc2, sizeof(double) * 4 * pl->nalm));
            spin2_coef_kernel<<<dim3(ceil_div(pl->lmax + 1, 128), pl->lmax + 1), 128, 0, st>>>(pl->d_rc2, pl->lmax);
            count_launch();
            CB_LAUNCH_CHECK();
        }
        P.rc2 = pl->d_rc2;
        size_t smem = 16 * (size_t)NSB * KC + (SPIN ? 32 * (size_t)NSB * KC : 0) +
                      8 * ((size_t)NSB * BSTAGE + (size_t)NVP * NSA * GPC * AROWS * ALD) +
                      4 * NVP * NSA * GPC + 8 * (2 * NPROD * NSA + 2 * NSB) + 1024;
        CUtensorMap tmap, tmap2;
        if (int rc = make_alm_tmap(&tmap, almT + chan0, nb, pl->nalm, alm_stride)) return rc;
        if (int rc = make_alm_tmap(&tmap2, (SPIN ? almB : almT) + chan0, nb, pl->nalm, alm_stride)) return rc;
        CB_REQUIRE(smem <= 227 * 1024, 3, "alm2map: the Legendre stage needs %zu B of shared memory (> 227 KB)", smem);
        CB_CUDA(cudaFuncSetAttribute(sht_legendre_ws_kernel<SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        long long nitems = (long long)(pl->lmax + 1) * P.nrb * P.ncb;
        CB_REQUIRE(nitems < 2147483647LL, 3, "alm2map: too many Legendre work items (%lld)", nitems);
        int dev, nsm;
        CB_CUDA(cudaGetDevice(&dev));
        CB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        const unsigned grid = (unsigned)std::min<long long>(nitems, nsm);   // persistent: one CTA per SM
        { KTimer kt(K_LEGENDRE, st); sht_legendre_ws_kernel<SPIN><<<grid, THREADS, smem, st>>>(P, (int)nitems, tmap, tmap2); }
        count_launch();
        CB_LAUNCH_CHECK();
        return 0;
    }
    size_t smem = sizeof(double) * ((SPIN ? 4 : 2) * (size_t)P.Lpad + 2 * LEG_KC * LEG_BLD + LEG_WARPS * 2 * (SPIN ? 16 : 8) * LEG_ALD);
    CB_REQUIRE(smem <= 227 * 1024, 3, "alm2map: lmax %d needs %zu B of shared memory (> 227 KB)", pl->lmax, smem);
    CB_CUDA(cudaFuncSetAttribute(sht_legendre_kernel<SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (long long)(pl->lmax + 1) * P.nrb * P.ncb;
    CB_REQUIRE(grid < 2147483647LL, 3, "alm2map: Legendre grid too large (%lld)", grid);
    { KTimer kt(K_LEGENDRE, st); sht_legendre_kernel<SPIN><<<(unsigned)grid, LEG_THREADS, smem, st>>>(P); }
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

static size_t phase_smem(const PhaseClass& pc) {
    const size_t seq = (size_t)pc.Mmax + pc.Mmax / 8 + 1;
    return pc.P * seq * 16 + (pc.Mmax <= 512 ? (size_t)PH_MAXSLICE_ITEMS * 4 * 16 : 0);
}

static int run_phase(const ShtPlan* pl, const double2* F, int nb, double* map, cudaStream_t st, long long map_stride = 0) {
    KTimer kt(K_PHASE, st);
    for (const auto& pc : pl->classes) {
        PhaseParams Q;
        Q.F = F; Q.map = map; Q.npix = map_stride > 0 ? map_stride : pl->npix;   // PhaseParams::npix is the channel stride of the output
        Q.rings = pl->d_rings; Q.ring_list = pc.d_rings;
        Q.tw = pl->d_tw; Q.chirp = pl->d_chirp; Q.bhat = pl->d_bhat;
        Q.chirp_off = pl->d_chirp_off; Q.bhat_off = pl->d_bhat_off;
        Q.shph = pl->d_shph;
        Q.lmax = pl->lmax; Q.nb = nb; Q.ncg = ceil_div(nb, 4); Q.P = pc.P; Q.Mmax = pc.Mmax; Q.log_tw = pl->log_tw;
        dim3 grid(pc.nrings, Q.ncg * (2 / pc.P));
        const size_t sm = phase_smem(pc);
        if (pc.threads == 256 && pc.P == 2) {
            if (pc.blu) sht_phase_kernel<256, 2, true><<<grid, 256, sm, st>>>(Q);
            else sht_phase_kernel<256, 2, false><<<grid, 256, sm, st>>>(Q);
        } else if (pc.threads == 256) {
            if (pc.blu) sht_phase_kernel<256, 1, true><<<grid, 256, sm, st>>>(Q);
            else sht_phase_kernel<256, 1, false><<<grid, 256, sm, st>>>(Q);
        } else if (pc.P == 2) {
            if (pc.blu) sht_phase_kernel<512, 2, true><<<grid, 512, sm, st>>>(Q);
            else sht_phase_kernel<512, 2, false><<<grid, 512, sm, st>>>(Q);
        } else {
            if (pc.blu) sht_phase_kernel<512, 1, true><<<grid, 512, sm, st>>>(Q);
            else sht_phase_kernel<512, 1, false><<<grid, 512, sm, st>>>(Q);
        }
        count_launch();
        CB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int cora_b200_alm2map(void* plan, const void* alm, int layout, long long alm_stride, int nchan, double* map,
                                 void* workspace, long long ws_bytes, void* stream) {
    return cora_b200_alm2map_strided(plan, alm, layout, alm_stride, nchan, map, 0, workspace, ws_bytes, stream);
}

extern "C" int cora_b200_alm2map_strided(void* plan, const void* alm, int layout, long long alm_stride, int nchan, double* map,
                                         long long map_stride, void* workspace, long long ws_bytes, void* stream) {
    CB_REQUIRE(plan && alm && map && workspace, 1, "alm2map: null argument");
    CB_REQUIRE(map_stride == 0 || map_stride >= ((ShtPlan*)plan)->npix, 1, "alm2map: map_stride %lld < npix", map_stride);
    const long long mstride = map_stride > 0 ? map_stride : ((ShtPlan*)plan)->npix;
    CB_REQUIRE(layout == CORA_B200_ALM_PACKED || layout == CORA_B200_ALM_PANEL, 1, "alm2map: unknown alm layout %d", layout);
    CB_REQUIRE(nchan >= 1, 1, "alm2map: nchan must be >= 1");
    ShtPlan* pl = (ShtPlan*)plan;
    cudaStream_t st = (cudaStream_t)stream;
    long long per = ws_per_chan(pl, layout);
    long long cap = ((ws_bytes - 256) / per) & ~3LL;   // F rows are padded to groups of 4 channels
    CB_REQUIRE(cap >= 4, 4, "alm2map: workspace too small (%lld B; need %lld B per 4 channels)", ws_bytes, 4 * per);
    int nbmax = (int)std::min<long long>(cap, nchan);
    if (nbmax < nchan && nbmax >= 16) nbmax -= nbmax % 16;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    for (int c0 = 0; c0 < nchan; c0 += nbmax) {
        int nb = std::min(nbmax, nchan - c0);
        double2* F = (double2*)ws;
        const double2* almT; long long stride; int chan0;
        if (layout == CORA_B200_ALM_PACKED) {
            double2* T = (double2*)(ws + (long long)(4 * pl->nside - 1) * (pl->lmax + 1) * 16 * round4(nbmax));
            dim3 grid(ceil_div(pl->nalm, 32), ceil_div(nb, 32));
            { KTimer kt(K_LAYOUT, st);
              alm_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>((const double2*)alm + (long long)c0 * alm_stride, alm_stride, nb,
                                                                pl->nalm, T, nb); }
            count_launch();
            CB_LAUNCH_CHECK();
            almT = T; stride = nb; chan0 = 0;
        } else {
            almT = (const double2*)alm; stride = alm_stride; chan0 = c0;
        }
        int rc = run_legendre<0>(pl, almT, nullptr, stride, chan0, nb, F, nullptr, st);
        if (rc) return rc;
        rc = run_phase(pl, F, nb, map + (long long)c0 * mstride, st, mstride);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int cora_b200_alm2map_spin2(void* plan, const void* almE, const void* almB, int layout, long long alm_stride, int nchan,
                                       double* mapQ, double* mapU, void* workspace, long long ws_bytes, void* stream) {
    return cora_b200_alm2map_spin2_strided(plan, almE, almB, layout, alm_stride, nchan, mapQ, mapU, 0, workspace, ws_bytes, stream);
}

extern "C" int cora_b200_alm2map_spin2_strided(void* plan, const void* almE, const void* almB, int layout, long long alm_stride,
                                               int nchan, double* mapQ, double* mapU, long long map_stride, void* workspace,
                                               long long ws_bytes, void* stream) {
    CB_REQUIRE(plan && almE && almB && mapQ && mapU && workspace, 1, "alm2map_spin2: null argument");
    CB_REQUIRE(map_stride == 0 || map_stride >= ((ShtPlan*)plan)->npix, 1, "alm2map_spin2: map_stride %lld < npix", map_stride);
    const long long mstride = map_stride > 0 ? map_stride : ((ShtPlan*)plan)->npix;
    CB_REQUIRE(layout == CORA_B200_ALM_PACKED || layout == CORA_B200_ALM_PANEL, 1, "alm2map_spin2: unknown alm layout %d", layout);
    CB_REQUIRE(nchan >= 1, 1, "alm2map_spin2: nchan must be >= 1");
    ShtPlan* pl = (ShtPlan*)plan;
    cudaStream_t st = (cudaStream_t)stream;
    const long long fbytes = (long long)(4 * pl->nside - 1) * (pl->lmax + 1) * 16;
    const long long per = 2 * fbytes + ((layout == CORA_B200_ALM_PACKED) ? 2 * pl->nalm * 16 : 0);
    long long cap = ((ws_bytes - 256) / per) & ~3LL;
    CB_REQUIRE(cap >= 4, 4, "alm2map_spin2: workspace too small (%lld B; need %lld B per 4 channels)", ws_bytes, 4 * per);
    int nbmax = (int)std::min<long long>(cap, nchan);
    if (nbmax < nchan && nbmax >= 8) nbmax -= nbmax % 8;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    for (int c0 = 0; c0 < nchan; c0 += nbmax) {
        int nb = std::min(nbmax, nchan - c0);
        double2* FQ = (double2*)ws;
        double2* FU = (double2*)(ws + fbytes * round4(nbmax));
        const double2 *pE, *pB; long long stride; int chan0;
        if (layout == CORA_B200_ALM_PACKED) {
            double2* TE = (double2*)(ws + 2 * fbytes * round4(nbmax));
            double2* TB = TE + pl->nalm * nbmax;
            dim3 grid(ceil_div(pl->nalm, 32), ceil_div(nb, 32));
            { KTimer kt(K_LAYOUT, st);
              alm_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>((const double2*)almE + (long long)c0 * alm_stride, alm_stride, nb,
                                                                pl->nalm, TE, nb);
              alm_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>((const double2*)almB + (long long)c0 * alm_stride, alm_stride, nb,
                                                                pl->nalm, TB, nb); }
            count_launch(2);
            CB_LAUNCH_CHECK();
            pE = TE; pB = TB; stride = nb; chan0 = 0;
        } else {
            pE = (const double2*)almE; pB = (const double2*)almB; stride = alm_stride; chan0 = c0;
        }
        int rc = run_legendre<2>(pl, pE, pB, stride, chan0, nb, FQ, FU, st);
        if (rc) return rc;
        rc = run_phase(pl, FQ, nb, mapQ + (long long)c0 * mstride, st, mstride);
        if (rc) return rc;
        rc = run_phase(pl, FU, nb, mapU + (long long)c0 * mstride, st, mstride);
        if (rc) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------- map2alm
extern "C" long long cora_b200_map2alm_workspace_bytes(void* plan, int nchan_batch) {
    if (!plan) return -1;
    const ShtPlan* pl = (const ShtPlan*)plan;
    return (long long)(4 * pl->nside - 1) * (pl->lmax + 1) * 16 * round4(nchan_batch) + 256;
}

extern "C" int cora_b200_map2alm(void* plan, const double* map, int nchan, const double* ring_weights, int accumulate,
                                 void* alm_panel, long long panel_stride, int chan0, void* workspace, long long ws_bytes,
                                 void* stream) {
    CB_REQUIRE(plan && map && alm_panel && workspace, 1, "map2alm: null argument");
    CB_REQUIRE(nchan >= 1 && chan0 >= 0 && panel_stride >= chan0 + nchan, 1, "map2alm: bad sizes (nchan=%d chan0=%d stride=%lld)", nchan,
               chan0, panel_stride);
    ShtPlan* pl = (ShtPlan*)plan;
    cudaStream_t st = (cudaStream_t)stream;
    const long long per = (long long)(4 * pl->nside - 1) * (pl->lmax + 1) * 16;
    long long cap = ((ws_bytes - 256) / per) & ~3LL;
    CB_REQUIRE(cap >= 4, 4, "map2alm: workspace too small (%lld B; need %lld B per 4 channels)", ws_bytes, 4 * per);
    int nbmax = (int)std::min<long long>(cap, nchan);
    if (nbmax < nchan && nbmax >= 8) nbmax -= nbmax % 8;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const size_t adj_smem = sizeof(double) * (2 * (size_t)ADJ_RT * ADJ_BLD + 8 * 2 * 8 * ADJ_ALD);
    CB_CUDA(cudaFuncSetAttribute(sht_legendre_adj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)adj_smem));
    CB_CUDA(cudaFuncSetAttribute(sht_ring_analysis_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CB_CUDA(cudaFuncSetAttribute(sht_ring_analysis_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int c0 = 0; c0 < nchan; c0 += nbmax) {
        const int nb = std::min(nbmax, nchan - c0);
        double2* F = (double2*)ws;
        {
            KTimer kt(K_PHASE, st);
            for (const auto& pc : pl->classes) {
                AnaParams Q;
                Q.map = map + (long long)c0 * pl->npix; Q.npix = pl->npix; Q.F = F; Q.rings = pl->d_rings; Q.ring_list = pc.d_rings;
                Q.tw = pl->d_tw; Q.chirp = pl->d_chirp; Q.bhat = pl->d_bhat;
                Q.chirp_off = pl->d_chirp_off; Q.bhat_off = pl->d_bhat_off;
                Q.wring = ring_weights; Q.wscale = 4.0 * 3.14159265358979323846 / (double)pl->npix; Q.nside = pl->nside;
                Q.lmax = pl->lmax; Q.nb = nb; Q.ncg = ceil_div(nb, 4); Q.P = pc.P; Q.Mmax = pc.Mmax; Q.log_tw = pl->log_tw;
                dim3 grid(pc.nrings, Q.ncg * (2 / pc.P));
                if (pc.threads == 256) sht_ring_analysis_kernel<256><<<grid, 256, phase_smem(pc), st>>>(Q);
                else sht_ring_analysis_kernel<512><<<grid, 512, phase_smem(pc), st>>>(Q);
                count_launch();
                CB_LAUNCH_CHECK();
            }
        }
        LegAdjParams P;
        P.F = F; P.alm = (double2*)alm_panel; P.alm_stride = panel_stride; P.chan0 = chan0 + c0; P.nb = nb;
        P.cth = pl->d_cth; P.sth = pl->d_sth; P.nm_mant = pl->d_nm_mant; P.nm_exp = pl->d_nm_exp; P.rc = pl->d_rc;
        P.nside = pl->nside; P.lmax = pl->lmax; P.nrn = pl->nrn;
        P.nrb = ceil_div(pl->nrn, ADJ_RT); P.ncb = ceil_div(nb, ADJ_NCH); P.ncg = ceil_div(nb, 4);
        P.accumulate = accumulate;
        if (P.nrb > 1 && !accumulate) {
            panel_zero_kernel<<<ceil_div(pl->nalm * nb, 256), 256, 0, st>>>((double2*)alm_panel, pl->nalm, panel_stride, chan0 + c0, nb);
            count_launch();
            CB_LAUNCH_CHECK();
        }
        const long long grid = (long long)(pl->lmax + 1) * P.nrb * P.ncb;
        CB_REQUIRE(grid < 2147483647LL, 3, "map2alm: Legendre grid too large (%lld)", grid);
        { KTimer kt(K_LEGENDRE, st); sht_legendre_adj_kernel<<<(unsigned)grid, ADJ_THREADS, adj_smem, st>>>(P); }
        count_launch();
        CB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int cora_b200_map_sub(const double* a, const double* b, long long n, double* out, void* stream) {
    CB_REQUIRE(a && b && out && n >= 1, 1, "map_sub: bad arguments");
    map_sub_kernel<<<(unsigned)std::min<long long>(ceil_div(n, 256), 148 * 16), 256, 0, (cudaStream_t)stream>>>(a, b, n, out);
    count_launch();
    CB_LAUNCH_CHECK();
    return 0;
}

static int run_ring_analysis(const ShtPlan* pl, const double* map, int nb, const double* ring_weights, double2* F, cudaStream_t st) {
    KTimer kt(K_PHASE, st);
    for (const auto& pc : pl->classes) {
        AnaParams Q;
        Q.map = map; Q.npix = pl->npix; Q.F = F; Q.rings = pl->d_rings; Q.ring_list = pc.d_rings;
        Q.tw = pl->d_tw; Q.chirp = pl->d_chirp; Q.bhat = pl->d_bhat;
        Q.chirp_off = pl->d_chirp_off; Q.bhat_off = pl->d_bhat_off;
        Q.wring = ring_weights; Q.wscale = 4.0 * 3.14159265358979323846 / (double)pl->npix; Q.nside = pl->nside;
        Q.lmax = pl->lmax; Q.nb = nb; Q.ncg = ceil_div(nb, 4); Q.P = pc.P; Q.Mmax = pc.Mmax; Q.log_tw = pl->log_tw;
        dim3 grid(pc.nrings, Q.ncg * (2 / pc.P));
        if (pc.threads == 256) sht_ring_analysis_kernel<256><<<grid, 256, phase_smem(pc), st>>>(Q);
        else sht_ring_analysis_kernel<512><<<grid, 512, phase_smem(pc), st>>>(Q);
        count_launch();
        CB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" long long cora_b200_map2alm_spin2_workspace_bytes(void* plan, int nchan_batch) {
    if (!plan) return -1;
    const ShtPlan* pl = (const ShtPlan*)plan;
    return 2 * (long long)(4 * pl->nside - 1) * (pl->lmax + 1) * 16 * round4(nchan_batch) + 256;
}

extern "C" int cora_b200_map2alm_spin2(void* plan, const double* mapQ, const double* mapU, int nchan, const double* ring_weights,
                                       int accumulate, void* almE_panel, void* almB_panel, long long panel_stride, int chan0,
                                       void* workspace, long long ws_bytes, void* stream) {
    CB_REQUIRE(plan && mapQ && mapU && almE_panel && almB_panel && workspace, 1, "map2alm_spin2: null argument");
    CB_REQUIRE(nchan >= 1 && chan0 >= 0 && panel_stride >= chan0 + nchan, 1, "map2alm_spin2: bad sizes");
    ShtPlan* pl = (ShtPlan*)plan;
    cudaStream_t st = (cudaStream_t)stream;
    const long long fbytes = (long long)(4 * pl->nside - 1) * (pl->lmax + 1) * 16;
    long long cap = ((ws_bytes - 256) / (2 * fbytes)) & ~3LL;
    CB_REQUIRE(cap >= 4, 4, "map2alm_spin2: workspace too small (%lld B; need %lld B per 4 channels)", ws_bytes, 8 * fbytes);
    int nbmax = (int)std::min<long long>(cap, nchan);
    if (nbmax < nchan && nbmax >= 8) nbmax -= nbmax % 8;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const size_t smem = sizeof(double) * (2 * (size_t)ADJ2_RT * ADJ2_BLD + 8 * 2 * 2 * 8 * ADJ_ALD + 8 * 256);
    CB_CUDA(cudaFuncSetAttribute(sht_legendre_adj2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CB_CUDA(cudaFuncSetAttribute(sht_ring_analysis_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CB_CUDA(cudaFuncSetAttribute(sht_ring_analysis_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int c0 = 0; c0 < nchan; c0 += nbmax) {
        const int nb = std::min(nbmax, nchan - c0);
        double2* FQ = (double2*)ws;
        double2* FU = (double2*)(ws + fbytes * round4(nbmax));
        int rc = run_ring_analysis(pl, mapQ + (long long)c0 * pl->npix, nb, ring_weights, FQ, st);
        if (rc) return rc;
        rc = run_ring_analysis(pl, mapU + (long long)c0 * pl->npix, nb, ring_weights, FU, st);
        if (rc) return rc;
        LegAdj2Params P;
        P.FQ = FQ; P.FU = FU; P.almE = (double2*)almE_panel; P.almB = (double2*)almB_panel; P.alm_stride = panel_stride;
        P.chan0 = chan0 + c0; P.nb = nb;
        P.cth = pl->d_cth; P.sth = pl->d_sth; P.nm_mant = pl->d_nm_mant; P.nm_exp = pl->d_nm_exp; P.rc = pl->d_rc;
        P.nside = pl->nside; P.lmax = pl->lmax; P.nrn = pl->nrn;
        P.nrb = ceil_div(pl->nrn, ADJ2_RT); P.ncb = ceil_div(nb, ADJ2_NCH); P.ncg = ceil_div(nb, 4);
        P.accumulate = accumulate;
        if (P.nrb > 1 && !accumulate) {
            panel_zero_kernel<<<ceil_div(pl->nalm * nb, 256), 256, 0, st>>>((double2*)almE_panel, pl->nalm, panel_stride, chan0 + c0, nb);
            panel_zero_kernel<<<ceil_div(pl->nalm * nb, 256), 256, 0, st>>>((double2*)almB_panel, pl->nalm, panel_stride, chan0 + c0, nb);
            count_launch(2);
            CB_LAUNCH_CHECK();
        }
        const long long grid = (long long)(pl->lmax + 1) * P.nrb * P.ncb;
        CB_REQUIRE(grid < 2147483647LL, 3, "map2alm_spin2: Legendre grid too large (%lld)", grid);
        { KTimer kt(K_LEGENDRE, st); sht_legendre_adj2_kernel<<<(unsigned)grid, ADJ_THREADS, smem, st>>>(P); }
        count_launch();
        CB_LAUNCH_CHECK();
    }
    return 0;
}
