// Gaussian draws + root application:  a_lm(nu) = sum_nu' M_l[nu, nu'] g_l[nu', m]   (sm_100a)
//
// replaces complex_std_normal + np.dot + the strided scatter of cora/core/skysim.py:119-121
// (cora/util/nputil.py:104-125).  Draws: counter-based Philox4x32-10, one call per complex
// variate, counter = (l, m, nu', 0), key = seed; Box-Muller in FP64.  Apply: batched FP64
// tensor-core GEMM (DMMA.8x8x4) per l, triangular roots (Cholesky) skip the zero half of K,
// output written straight into the PANEL layout the Legendre stage reads.
#include "common.cuh"
#include "cora_b200.h"

#include <algorithm>
#include <vector>

namespace cb {

// --------------------------------------------------------------------------- Philox
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0,
                                              unsigned k1, unsigned out[4]) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const unsigned hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// complex standard normal (re, im each N(0, 1/2)) for (l, m, nu')
__device__ __forceinline__ double2 philox_cnormal(unsigned long long seed, unsigned l, unsigned m, unsigned nu) {
    unsigned r[4];
    philox4x32_10(l, m, nu, 0u, (unsigned)seed, (unsigned)(seed >> 32), r);
    const unsigned long long a = ((unsigned long long)r[1] << 32) | r[0];
    const unsigned long long b = ((unsigned long long)r[3] << 32) | r[2];
    const double u1 = ((double)(a >> 11) + 0.5) * 0x1p-53;   // (0, 1)
    const double u2 = ((double)(b >> 11) + 0.5) * 0x1p-53;
    const double rad = sqrt(-log(u1));                       // sqrt(-2 ln u)/sqrt(2)
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    return make_double2(rad * c, rad * s);
}

struct LDesc {
    long long goff;   // offset (complex elements) of this l's draws in the gauss buffer
    long long orow;   // slab output: row of (l, m = 0); rows of one l are consecutive in m
    int l;            // global multipole
    int ld;           // row length (complex) of the gauss buffer for this l
};

__global__ void draw_kernel(const LDesc* __restrict__ ld, int nz, unsigned long long seed, double2* __restrict__ G, unsigned nu_ctr0) {
    const LDesc d = ld[blockIdx.z];
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int nu = blockIdx.y;
    if (m > d.l) return;
    G[d.goff + (long long)nu * d.ld + m] = philox_cnormal(seed, (unsigned)d.l, (unsigned)m, (unsigned)nu + nu_ctr0);
}

// ---------------------------------------------------------------------------- apply
constexpr int AP_TM = 128;   // nu rows per CTA
constexpr int AP_TN = 64;    // real columns per CTA (32 m's)
constexpr int AP_KC = 16;
constexpr int AP_ALD = 20;   // As[128][20]  (g*20 + t: conflict-free per half-warp)
constexpr int AP_BLD = 68;   // Bs[16][68]   (t*68 + g)

struct ApplyParams {
    const double* root;       // [nl][nz][nz]
    const LDesc* ldesc;
    const int* dense;         // per-l flag: 1 = dense root (eigh), 0 = lower triangular; may be null (all dense)
    const double2* G;
    double2* panel;
    long long panel_stride;
    // slab output (multi-GPU send buffer): element (row, nu) at nu_base[nu] + row * nu_width[nu];
    // null -> PANEL output
    const long long* nu_base;
    const int* nu_width;
    // peer output (multi-GPU, fused exchange): element (l, m, nu) at nu_ptr[nu][idx(l, m) * nu_width[nu]],
    // nu_ptr[nu] = the PANEL buffer of the GPU that owns channel nu, advanced to that channel's column
    double2* const* nu_ptr;
    int nz, lmax, chan0, nu0, nnu;
};

constexpr int AP_STAGES = 3;
constexpr int AP_STAGE_DOUBLES = AP_TM * AP_ALD + AP_KC * AP_BLD;

// FAST = true: operands staged with a 3-deep cp.async pipeline (needs an even nz so that every
// 16-byte piece of a root row is aligned); FAST = false: plain loads, any nz.
template <bool FAST>
__global__ void __launch_bounds__(256, 2) apply_kernel(ApplyParams P) {
    extern __shared__ __align__(16) double ap_smem[];
    const int li = blockIdx.z;
    const LDesc d = P.ldesc[li];
    const int m0 = blockIdx.x * (AP_TN / 2);
    if (m0 > d.l) return;
    const int r0 = P.nu0 + blockIdx.y * AP_TM;           // first nu row of the tile
    const int nz = P.nz;
    const double* M = P.root + (long long)li * nz * nz;
    const double* Gd = (const double*)(P.G + d.goff);    // rows nu', 2*ld doubles each
    const int gld = 2 * d.ld;
    // root kind (cora_b200_root_batched's used_eigh): 0 = lower triangular (Cholesky); v >= 1 = dense fallback
    // root whose first v - 1 columns are zero (clipped modes / beyond the numerical rank): those are skipped
    const int dflag = P.dense ? P.dense[li] : 1;
    const bool tri = (dflag == 0);
    const int kbeg = (dflag > 1) ? (min(dflag - 1, nz) & ~(AP_KC - 1)) : 0;
    const int kend = tri ? min(nz, r0 + AP_TM) : nz;
    const int ncols = 2 * (d.l + 1);                      // valid real columns

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int nit = (kend - kbeg + AP_KC - 1) / AP_KC;
    auto stage_load = [&](int it, int slot) {
        double* As = ap_smem + slot * AP_STAGE_DOUBLES;
        double* Bs = As + AP_TM * AP_ALD;
        const int k0 = kbeg + it * AP_KC;
        if (FAST) {
            // A tile 128 x 16: 1024 pieces of 2 doubles
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int e = tid + q * 256;
                const int rr = e >> 3, kk = (e & 7) * 2;
                const int r = r0 + rr, k = k0 + kk;
                const bool ok = (r < nz) && (k < nz) && (it < nit);
                cp_async16(As + rr * AP_ALD + kk, M + (ok ? ((long long)r * nz + k) : 0), ok);
            }
            // B tile 16 x 64: 512 pieces of one complex
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int e = tid + q * 256;
                const int kk = e >> 5, cc = (e & 31) * 2;
                const int k = k0 + kk, col = 2 * m0 + cc;
                const bool ok = (k < nz) && (col < ncols) && (it < nit);
                cp_async16(Bs + kk * AP_BLD + cc, Gd + (ok ? ((long long)k * gld + col) : 0), ok);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int e = tid + q * 256;
                const int rr = e >> 4, kk = e & 15;
                const int r = r0 + rr, k = k0 + kk;
                As[rr * AP_ALD + kk] = (r < nz && k < nz && it < nit) ? M[(long long)r * nz + k] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int e = tid + q * 256;
                const int kk = e >> 6, cc = e & 63;
                const int k = k0 + kk, col = 2 * m0 + cc;
                Bs[kk * AP_BLD + cc] = (k < nz && col < ncols && it < nit) ? Gd[(long long)k * gld + col] : 0.0;
            }
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < AP_STAGES - 1; s++) stage_load(s, s);
    for (int it = 0; it < nit; it++) {
        cp_async_wait<AP_STAGES - 2>();
        __syncthreads();   // stage `it` landed for everyone; the slot refilled below was consumed last iteration
        stage_load(it + AP_STAGES - 1, (it + AP_STAGES - 1) % AP_STAGES);
        const double* As = ap_smem + (it % AP_STAGES) * AP_STAGE_DOUBLES;
        const double* Bs = As + AP_TM * AP_ALD;
        // Triangular (Cholesky) roots: an 8-row block needs only the columns k <= its last row; everything
        // to the right is exact zeros.  The DMMA pipe is what bounds this kernel (82 % busy in ncu), so the
        // zero blocks are skipped per 8 rows x 4 columns -- warp-uniform tests, bit-identical sums.
        // (Measured alternatives: interleaving the 8-row blocks over the warps to even out the work is slower,
        // 2.02 ms vs 1.75 ms; re-pairing row groups over SM sub-partitions changes nothing.)
        const int kk0 = kbeg + it * AP_KC;
        const int rw = r0 + wm * 32;                     // first row of this warp's 32
        if (tri && kk0 > rw + 31) continue;              // the whole chunk lies right of this warp's rows
#pragma unroll
        for (int k4 = 0; k4 < AP_KC / 4; k4++) {
            const int kk = kk0 + 4 * k4;
            double af[4], bf[4];
#pragma unroll
            for (int mb = 0; mb < 4; mb++) af[mb] = As[(wm * 32 + 8 * mb + g) * AP_ALD + k4 * 4 + t];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) bf[nb] = Bs[(k4 * 4 + t) * AP_BLD + wn * 32 + 8 * nb + g];
#pragma unroll
            for (int mb = 0; mb < 4; mb++) {
                if (tri && kk > rw + 8 * mb + 7) continue;
#pragma unroll
                for (int nb = 0; nb < 4; nb++) dmma884(acc[mb][nb][0], acc[mb][nb][1], af[mb], bf[nb]);
            }
        }
    }
    cp_async_wait<0>();
    // epilogue: C[g][2t], C[g][2t+1] = (re, im) of (nu = row g, m = col t)
    const int lmax = P.lmax;
#pragma unroll
    for (int nb = 0; nb < 4; nb++) {
        const int m = m0 + wn * 16 + 4 * nb + t;
        if (m > d.l) continue;
        if (P.nu_ptr) {
            const long long idx = (long long)m * (2 * lmax + 1 - m) / 2 + d.l;
#pragma unroll
            for (int mb = 0; mb < 4; mb++) {
                const int nu = r0 + wm * 32 + 8 * mb + g;
                if (nu < nz) P.nu_ptr[nu][idx * P.nu_width[nu]] = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
            }
            continue;
        }
        if (P.nu_base) {
            const long long orow = d.orow + m;
#pragma unroll
            for (int mb = 0; mb < 4; mb++) {
                const int nu = r0 + wm * 32 + 8 * mb + g;
                if (nu < nz) P.panel[P.nu_base[nu] + orow * P.nu_width[nu]] = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
            }
            continue;
        }
        const long long idx = (long long)m * (2 * lmax + 1 - m) / 2 + d.l;
        double2* row = P.panel + idx * P.panel_stride + P.chan0;
#pragma unroll
        for (int mb = 0; mb < 4; mb++) {
            const int nu = r0 + wm * 32 + 8 * mb + g;
            if (nu < P.nu0 + P.nnu && nu < nz) row[nu - P.nu0] = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
        }
    }
}

}  // namespace cb

using namespace cb;

extern "C" long long cora_b200_draw_apply_workspace_bytes(int nz, int lmax_in_batch, int nl_batch) {
    return 16LL * nz * (long long)(lmax_in_batch + 1) * nl_batch + 64LL * nl_batch + 1024;
}

// Descriptor tables are tiny and identical from step to step: keep them in library-owned device
// memory keyed by content, so a steady-state step uploads nothing and never blocks the host.
struct DescEntry { int dev; std::vector<LDesc> hd; LDesc* dd; };
static std::vector<DescEntry> g_desc_cache;

static int cached_descs(const std::vector<LDesc>& hd, LDesc** out) {
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    for (auto& e : g_desc_cache)
        if (e.dev == dev && e.hd.size() == hd.size() && memcmp(e.hd.data(), hd.data(), sizeof(LDesc) * hd.size()) == 0) {
            *out = e.dd;
            return 0;
        }
    if (g_desc_cache.size() >= 512) {   // evict the oldest (never one in flight: a sync precedes the free)
        CB_CUDA(cudaDeviceSynchronize());
        cudaFree(g_desc_cache.front().dd);
        g_desc_cache.erase(g_desc_cache.begin());
    }
    DescEntry e;
    e.dev = dev; e.hd = hd; e.dd = nullptr;
    CB_CUDA(cudaMalloc(&e.dd, sizeof(LDesc) * hd.size()));
    CB_CUDA(cudaMemcpy(e.dd, hd.data(), sizeof(LDesc) * hd.size(), cudaMemcpyHostToDevice));
    g_desc_cache.push_back(e);
    *out = e.dd;
    return 0;
}

static int draw_apply_impl(const double* root, const int* l_list_h, const int* dense_flag, int nl, int nz, int lmax,
                           unsigned long long seed, const void* gauss, long long gauss_ld, void* alm_panel,
                           long long panel_stride, int chan0, int nu0, int nnu, const long long* row0_h,
                           const long long* nu_base, const int* nu_width, void* workspace, long long ws_bytes,
                           void* stream, const void* nu_ptr = nullptr, unsigned nu_ctr0 = 0) {
    CB_REQUIRE(root && l_list_h && (alm_panel || nu_ptr) && workspace, 1, "draw_apply: null argument");
    CB_REQUIRE(nl >= 1 && nz >= 1 && lmax >= 0 && nnu >= 1 && nu0 >= 0 && nu0 + nnu <= nz, 1, "draw_apply: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    long long avail = (char*)workspace + ws_bytes - ws;
    for (int i = 0; i < nl; i++)
        CB_REQUIRE(l_list_h[i] >= 0 && l_list_h[i] <= lmax, 1, "draw_apply: l_list[%d]=%d outside [0,%d]", i, l_list_h[i], lmax);
    const bool packed = gauss && gauss_ld < 0;     // draws made by cora_b200_draw: block i holds [nz][l_i + 1], blocks back to back
    if (gauss && !packed) CB_REQUIRE(gauss_ld >= lmax + 1 || nl == 0, 1, "draw_apply: gauss_ld %lld < lmax+1", gauss_ld);
    long long packed_off = 0;

    int i0 = 0;
    while (i0 < nl) {
        // choose a batch of l's whose descriptors (+ generated draws) fit the workspace
        std::vector<LDesc> hd;
        long long gneed = 0;
        int i1 = i0;
        while (i1 < nl) {
            LDesc d;
            d.l = l_list_h[i1];
            d.orow = row0_h ? row0_h[i1] : 0;
            if (packed) { d.goff = packed_off; d.ld = d.l + 1; packed_off += (long long)nz * d.ld; }
            else if (gauss) { d.goff = (long long)i1 * nz * gauss_ld; d.ld = (int)gauss_ld; }
            else { d.goff = gneed; d.ld = d.l + 1; }
            long long add = gauss ? 0 : (long long)nz * d.ld;
            long long bytes = ((long long)(hd.size() + 1) * sizeof(LDesc) + 255) / 256 * 256 + 16 * (gneed + add) + 512;
            if (bytes > avail && i1 > i0) { if (packed) packed_off -= (long long)nz * d.ld; break; }
            CB_REQUIRE(bytes <= avail, 4, "draw_apply: workspace too small (%lld B) for a single l (needs %lld B)", avail, bytes);
            gneed += add;
            hd.push_back(d);
            i1++;
            if (hd.size() >= 65535) break;
        }
        const int nb = (int)hd.size();
        LDesc* dd = nullptr;
        if (int rc = cached_descs(hd, &dd)) return rc;
        double2* Gbuf = (double2*)ws;
        int lbig = 0;
        for (auto& d : hd) lbig = std::max(lbig, d.l);
        const double2* Gsrc = (const double2*)gauss;
        if (!gauss) {
            { KTimer kt(K_DRAW, st); draw_kernel<<<dim3(ceil_div(lbig + 1, 128), nz, nb), 128, 0, st>>>(dd, nz, seed, Gbuf, nu_ctr0); }
            count_launch();
            CB_LAUNCH_CHECK();
            Gsrc = Gbuf;
        }
        ApplyParams P;
        P.root = root + (long long)i0 * nz * nz;
        P.ldesc = dd;
        P.dense = dense_flag ? dense_flag + i0 : nullptr;
        P.G = Gsrc; P.panel = (double2*)alm_panel; P.panel_stride = panel_stride;
        P.nu_base = nu_base; P.nu_width = nu_width; P.nu_ptr = (double2* const*)nu_ptr;
        P.nz = nz; P.lmax = lmax; P.chan0 = chan0; P.nu0 = nu0; P.nnu = nnu;
        dim3 grid(ceil_div(lbig + 1, AP_TN / 2), ceil_div(nnu, AP_TM), nb);
        {
            KTimer kt(K_APPLY, st);
            const size_t smem = sizeof(double) * AP_STAGES * AP_STAGE_DOUBLES;
            if (nz % 2 == 0) {
                CB_CUDA(cudaFuncSetAttribute(apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                apply_kernel<true><<<grid, 256, smem, st>>>(P);
            } else {
                CB_CUDA(cudaFuncSetAttribute(apply_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                apply_kernel<false><<<grid, 256, smem, st>>>(P);
            }
        }
        count_launch();
        CB_LAUNCH_CHECK();
        // the next batch's draws overwrite this batch's in the same workspace: safe without any host
        // synchronisation, because the draw kernel is ordered behind this apply on the same stream and the
        // descriptor tables are immutable, content-keyed device copies
        i0 = i1;
    }
    return 0;
}

extern "C" long long cora_b200_draw_bytes(const int* l_list_h, int nl, int nz) {
    if (!l_list_h || nl < 0 || nz < 1) return -1;
    long long n = 0;
    for (int i = 0; i < nl; i++) n += (long long)nz * (l_list_h[i] + 1);
    return 16 * n;
}

extern "C" int cora_b200_draw(const int* l_list_h, int nl, int nz, unsigned long long seed, int draw_counter0, void* gauss_packed,
                              long long bytes, void* stream) {
    CB_REQUIRE(l_list_h && gauss_packed && nl >= 1 && nz >= 1 && draw_counter0 >= 0, 1, "draw: bad arguments");
    CB_REQUIRE(bytes >= cora_b200_draw_bytes(l_list_h, nl, nz), 4, "draw: buffer too small (%lld B, need %lld B)", bytes,
               cora_b200_draw_bytes(l_list_h, nl, nz));
    cudaStream_t st = (cudaStream_t)stream;
    for (int i0 = 0; i0 < nl; i0 += 65535) {
        const int nb = std::min(65535, nl - i0);
        std::vector<LDesc> hd(nb);
        long long off = 0;
        for (int i = 0; i < i0; i++) off += (long long)nz * (l_list_h[i] + 1);
        int lbig = 0;
        for (int i = 0; i < nb; i++) {
            CB_REQUIRE(l_list_h[i0 + i] >= 0, 1, "draw: negative l");
            hd[i].l = l_list_h[i0 + i];
            hd[i].ld = hd[i].l + 1;
            hd[i].goff = off;
            hd[i].orow = 0;
            off += (long long)nz * hd[i].ld;
            lbig = std::max(lbig, hd[i].l);
        }
        LDesc* dd = nullptr;
        if (int rc = cached_descs(hd, &dd)) return rc;
        { KTimer kt(K_DRAW, st); draw_kernel<<<dim3(ceil_div(lbig + 1, 128), nz, nb), 128, 0, st>>>(dd, nz, seed, (double2*)gauss_packed, (unsigned)draw_counter0); }
        count_launch();
        CB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int cora_b200_draw_apply(const double* root, const int* l_list_h, const int* dense_flag, int nl, int nz, int lmax,
                                    unsigned long long seed, const void* gauss, long long gauss_ld, void* alm_panel,
                                    long long panel_stride, int chan0, int nu0, int nnu, void* workspace,
                                    long long ws_bytes, void* stream) {
    return draw_apply_impl(root, l_list_h, dense_flag, nl, nz, lmax, seed, gauss, gauss_ld, alm_panel, panel_stride, chan0,
                           nu0, nnu, nullptr, nullptr, nullptr, workspace, ws_bytes, stream);
}

extern "C" int cora_b200_draw_apply_slabs(const double* root, const int* l_list_h, const int* dense_flag, int nl, int nz,
                                          int lmax, unsigned long long seed, const void* gauss, long long gauss_ld,
                                          const long long* row0_h, const long long* nu_base, const int* nu_width,
                                          void* send, void* workspace, long long ws_bytes, void* stream) {
    CB_REQUIRE(row0_h && nu_base && nu_width, 1, "draw_apply_slabs: null slab description");
    return draw_apply_impl(root, l_list_h, dense_flag, nl, nz, lmax, seed, gauss, gauss_ld, send, 0, 0, 0, nz, row0_h,
                           nu_base, nu_width, workspace, ws_bytes, stream);
}

extern "C" int cora_b200_draw_apply_peers(const double* root, const int* l_list_h, const int* dense_flag, int nl, int nz,
                                          int lmax, unsigned long long seed, int draw_counter0, const void* gauss,
                                          long long gauss_ld, const void* nu_ptr, const int* nu_width, void* workspace,
                                          long long ws_bytes, void* stream) {
    CB_REQUIRE(nu_ptr && nu_width && draw_counter0 >= 0, 1, "draw_apply_peers: bad peer description");
    return draw_apply_impl(root, l_list_h, dense_flag, nl, nz, lmax, seed, gauss, gauss_ld, nullptr, 0, 0, 0, nz, nullptr,
                           nullptr, nu_width, workspace, ws_bytes, stream, nu_ptr, (unsigned)draw_counter0);
}
