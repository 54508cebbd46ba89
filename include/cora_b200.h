/* cora_b200 -- C ABI of the B200-native full-sky Gaussian field generator.
 *
 * Drop-in boundary for the hot path of radiocosmology/cora:
 *   cora.core.skysim.clarray / mkfullsky  ->  nputil.matrix_root_manynull, complex_std_normal,
 *   np.dot  ->  hputil.sphtrans_inv_sky  ->  healpy.alm2map.
 * The reference has no FFI layer of its own (it is Python + Cython); these are the entry
 * points a maintainer binds with ctypes.  Each declaration cites the
 * reference interface it replaces (file:line under the reference checkout).
 *
 * Conventions
 *   - every function returns int: 0 = OK, non-zero = error (text via cora_b200_last_error);
 *     nothing throws, nothing allocates caller-visible memory except the opaque plans;
 *   - all data pointers are DEVICE pointers on the current CUDA device unless the name ends
 *     in _h (host); `stream` is a cudaStream_t passed as void*;
 *   - complex numbers are interleaved (re, im) float64 pairs ("complex128");
 *   - work is asynchronous on `stream`; the caller synchronises.
 */
#ifndef CORA_B200_H
#define CORA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ library ------ */
int cora_b200_version(void);
const char* cora_b200_last_error(void);
/* number of CUDA kernels this library has launched since load (bench.py: gpu_launches) */
long long cora_b200_launch_count(void);
/* FP64 tensor-core (DMMA) peak of the current device measured on the spot, TFLOP/s.
 * Used as the roofline denominator of the FP64-bound stages (MEASURED_PEAKS.json has no
 * FP64 entry).  `ms_budget`: approximate run time of the probe. */
int cora_b200_fp64_peak(double ms_budget, double* tflops_out, void* stream);

/* ------------------------------------------------------------------ alm layouts -- */
/* PACKED: healpy order per channel, alm[chan * stride + idx(l,m)],
 *         idx(l,m) = m (2 lmax + 1 - m)/2 + l          (cora/util/hputil.py:124-152)
 * PANEL : the library's working layout, alm[idx(l,m) * stride + chan] (channels
 *         contiguous; what draw_apply writes and the Legendre stage reads).            */
#define CORA_B200_ALM_PACKED 0
#define CORA_B200_ALM_PANEL 1

/* ------------------------------------------------------------------ inverse SHT -- */
/* Plan for healpy.alm2map on the HEALPix RING grid, lmax = mmax, no beam / pixel window.
 * replaces: healpy.alm2map as called at cora/util/hputil.py:388 (scalar), :420-423
 * (T,E,B -> T,Q,U), :426-430 (V).                                                      */
int cora_b200_sht_plan_create(int nside, int lmax, void** plan_out);
int cora_b200_sht_plan_destroy(void* plan);
long long cora_b200_alm2map_workspace_bytes(void* plan, int layout, int nchan_batch);

/* Scalar synthesis of `nchan` channels: map[chan * npix + pix], float64, RING order.
 * replaces: the per-frequency loop of hputil.sphtrans_inv_sky -> sphtrans_inv_real
 * (cora/util/hputil.py:500-531, :369-391).  `alm_stride` is the stride (in complex
 * elements) between channels (PACKED) or between idx rows (PANEL).  The workspace may be
 * smaller than cora_b200_alm2map_workspace_bytes(nchan): channels are then processed in
 * batches.                                                                              */
int cora_b200_alm2map(void* plan, const void* alm, int layout, long long alm_stride, int nchan,
                      double* map, void* workspace, long long ws_bytes, void* stream);

/* Spin-2 synthesis (E,B) -> (Q,U), HEALPix sign convention.
 * replaces: the polarised part of healpy.alm2map([T,E,B]) at cora/util/hputil.py:419-423. */
int cora_b200_alm2map_spin2(void* plan, const void* almE, const void* almB, int layout,
                            long long alm_stride, int nchan, double* mapQ, double* mapU,
                            void* workspace, long long ws_bytes, void* stream);

/* PANEL -> cora dense alm[chan][l][m] (complex128[nchan, L, L], zeros for m > l)
 * replaces: the layout mkfullsky(alms=True) returns (cora/core/skysim.py:108-125).     */
int cora_b200_alm_panel_to_dense(const void* alm_panel, long long panel_stride, int chan0, int nchan,
                                 int lmax, void* dense, void* stream);
/* cora dense alm[chan][l][m] -> PANEL (the pack_alm step, cora/util/hputil.py:124-152) */
int cora_b200_alm_dense_to_panel(const void* dense, int nchan, int lmax, void* alm_panel,
                                 long long panel_stride, int chan0, void* stream);

/* ------------------------------------------------------------------ C_l fill ----- */
/* SCK foreground spectrum, Romberg-averaged over each channel.
 * replaces: skysim.clarray(ForegroundSCK.angular_powerspectrum, ...)
 * (cora/core/skysim.py:10-69 + cora/foreground/gaussianfg.py:107-130).
 * nu_samples[nz * zint]: per-channel sample frequencies; w[zint]: Romberg weights already
 * divided so that sum_ab w_a w_b = 1; out_cl[nl][nz][nz] for l = l0 .. l0+nl-1.          */
int cora_b200_cl_fill_sck(double A, double beta, double l_ref, double alpha, double nu_ref, double zeta,
                          const double* nu_samples, const double* w, int l0, int nl, int nz, int zint,
                          double* out_cl, void* stream);

/* 21cm: one-off P(k_perp, k_par) table -> three DCT-I tables (dd, dv, vv), stored
 * y-major and interleaved: tab[(y * nkperp + x) * 3 + {dd,dv,vv}].
 * replaces: the cache build of RedshiftCorrelation.angular_powerspectrum_fft
 * (cora/signal/corr.py:915-942) incl. the log-log cubic spline of ps_z1.5.dat
 * (cora/util/cubicspline.pyx:124-175,274-288) and the exp(-k^2/2k*^2) cut
 * (cora/signal/corr21cm.py:25-29).  Knots are host arrays (ln k, ln P, y2).              */
int cora_b200_ps_table_21cm(const double* lnk_h, const double* lnp_h, const double* y2_h, int nknot,
                            double kstar, double* tab, void* workspace, long long ws_bytes, void* stream);
long long cora_b200_ps_table_21cm_bytes(void);
long long cora_b200_ps_table_21cm_workspace_bytes(void);

/* 21cm C_l fill with fused bilinear lookup + Romberg average.
 * replaces: skysim.clarray(Corr21cm.angular_powerspectrum, ...) i.e. cora/core/skysim.py:41-67
 * + cora/signal/corr.py:944-982 + cora/util/bilinearmap.pyx:14-59.
 * Per-sample vectors (length nz*zint, host-computed exactly as the reference does):
 * chi (comoving distance), b, f, pf, D.                                                  */
int cora_b200_cl_fill_21cm(const double* tab, const double* chi, const double* b, const double* f,
                           const double* pf, const double* D, const double* w, int l0, int nl, int nz,
                           int zint, double* out_cl, void* stream);

/* Romberg average of a block evaluated by a generic host callable:
 * in[nl][nz][zint][nz][zint] -> out[nl][nz][nz]  (the two scipy.integrate.romb calls and the
 * normalisation of cora/core/skysim.py:62-67).  w as above (sum_a w_a = 1).                */
int cora_b200_cl_romberg_reduce(const double* in, const double* w, int nl, int nz, int zint, double* out,
                                void* stream);

/* Point-wise spectra for arbitrary (already broadcast) argument arrays of length n.
 * replaces: ForegroundSCK.angular_powerspectrum (cora/foreground/gaussianfg.py:40-41,107-130)
 * and RedshiftCorrelation.angular_powerspectrum_fft (cora/signal/corr.py:944-982).
 * vec1/vec2: [5][n] rows chi, b, f, pf, D of the two redshift arguments.                    */
int cora_b200_aps_sck_points(double A, double beta, double l_ref, double alpha, double nu_ref, double zeta,
                             const double* l, const double* nu1, const double* nu2, long long n, double* out,
                             void* stream);
int cora_b200_aps_21cm_points(const double* tab, const double* l, const double* vec1, const double* vec2,
                              long long n, double* out, void* stream);
/* Read table entries back in the reference's [x = k_perp index][y = r_par index] indexing:
 * out[3 e + {dd,dv,vv}] (the save_fft_cache view, cora/signal/corr.py:870-878).            */
int cora_b200_ps_table_21cm_gather(const double* tab, const int* x, const int* y, int n, double* out, void* stream);

/* ------------------------------------------------------------------ root --------- */
/* Batched matrix root with the reference's semantics: add jitter_rel * max(diag) to the
 * diagonal (cora/core/skysim.py:116-117), try Cholesky (lower), and where a pivot is not
 * positive fall back to a symmetric eigen-decomposition with eigenvalues below
 * clip_rel * max set to zero, root = evecs * sqrt(evals) (cora/util/nputil.py:51-101,
 * truncate=False).  cl / root: [nl][nz][nz].  used_eigh / num_pos: int[nl].
 * workspace: cora_b200_root_workspace_bytes(nl, nz).                                      */
long long cora_b200_root_workspace_bytes(int nl, int nz);
int cora_b200_root_batched(const double* cl, int nl, int nz, double jitter_rel, double clip_rel,
                           double* root, int* used_eigh, int* num_pos, void* workspace,
                           long long ws_bytes, void* stream);

/* ------------------------------------------------------------------ draw + apply - */
/* alm[nu, l, m] = sum_nu' M_l[nu, nu'] g_l[nu', m], written in PANEL layout.
 * replaces: complex_std_normal + np.dot + scatter (cora/core/skysim.py:119-121,
 * cora/util/nputil.py:104-125).
 *   root[nl][nz][nz]   roots in the order of l_list_h
 *   l_list_h[nl]       HOST array: global l of each root (any subset -> l-sharding)
 *   dense_flag[nl]     device int per root: 0 = lower-triangular (Cholesky) so the zero half
 *                      of the contraction is skipped, 1 = dense (eigh); NULL = all dense
 *   gauss == NULL      draws come from Philox4x32-10 keyed by `seed`, counter (l, m, nu', 0),
 *                      Box-Muller, (N + iN)/sqrt(2)
 *   gauss != NULL      injected draws, complex128 gauss[i][nu'][gauss_ld] (i = position in
 *                      l_list_h, columns m <= l used): the identical-draw parity path
 * Output rows nu in [nu0, nu0+nnu) go to alm_panel[idx(l,m) * panel_stride + chan0 + (nu-nu0)].
 * The workspace holds the generated draws of a batch of l's; any size that fits one l works. */
long long cora_b200_draw_apply_workspace_bytes(int nz, int lmax_in_batch, int nl_batch);
int cora_b200_draw_apply(const double* root, const int* l_list_h, const int* dense_flag, int nl, int nz,
                         int lmax, unsigned long long seed, const void* gauss, long long gauss_ld,
                         void* alm_panel, long long panel_stride, int chan0, int nu0, int nnu,
                         void* workspace, long long ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CORA_B200_H */
