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
 *     nothing throws; the only allocations are the opaque SHT plans, the peer buffers handed out
 *     by cora_b200_peer_alloc and a small library-owned cache of descriptor tables;
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

/* Per-kernel device timing.  While enabled, every kernel family the library launches is
 * bracketed by CUDA events on its launching stream; timing_read synchronises, returns the
 * summed milliseconds and launch counts per family (arrays of >= cora_b200_timing_kinds()
 * entries, names from cora_b200_timing_name) and clears the record.  bench.py computes
 * roofline.achieved from these. */
int cora_b200_timing_enable(int on);
int cora_b200_timing_kinds(void);
const char* cora_b200_timing_name(int id);
int cora_b200_timing_read(double* ms_out_h, long long* launches_out_h, int n);
/* Timeline of the recorded spans in launch order (family id, start relative to the first span, duration; ms).
 * Returns MINUS the number of entries written (<= cap), or a positive error code.  Does not clear the record. */
int cora_b200_timing_trace(int* id_out_h, double* start_ms_out_h, double* dur_ms_out_h, int cap);

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

/* As cora_b200_alm2map with an explicit channel stride of the output (map[chan * map_stride + pix],
 * map_stride >= npix; 0 = npix): lets T, Q, U, V be written straight into the reference's
 * [freq][pol][pix] layout (cora/util/hputil.py:516-521) with map_stride = npol * npix.       */
int cora_b200_alm2map_strided(void* plan, const void* alm, int layout, long long alm_stride, int nchan,
                              double* map, long long map_stride, void* workspace, long long ws_bytes, void* stream);

/* Spin-2 synthesis (E,B) -> (Q,U), HEALPix sign convention.
 * replaces: the polarised part of healpy.alm2map([T,E,B]) at cora/util/hputil.py:419-423. */
int cora_b200_alm2map_spin2(void* plan, const void* almE, const void* almB, int layout,
                            long long alm_stride, int nchan, double* mapQ, double* mapU,
                            void* workspace, long long ws_bytes, void* stream);

int cora_b200_alm2map_spin2_strided(void* plan, const void* almE, const void* almB, int layout,
                                    long long alm_stride, int nchan, double* mapQ, double* mapU,
                                    long long map_stride, void* workspace, long long ws_bytes, void* stream);

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
 * divided so that sum_ab w_a w_b = 1; out_cl[nl][nz][nz] for l = l0 + i * l_step, i < nl
 * (l_step = 1: a contiguous block; l_step = G: the interleaved l-shard of one of G GPUs). */
int cora_b200_cl_fill_sck(double A, double beta, double l_ref, double alpha, double nu_ref, double zeta,
                          const double* nu_samples, const double* w, int l0, int l_step, int nl, int nz, int zint,
                          double* out_cl, void* stream);

/* 21cm: one-off P(k_perp, k_par) table -> three DCT-I tables (dd, dv, vv), stored
 * planar: tab[({dd,dv,vv} * nkpar + y) * nkperp + x]  (y = r_par index, x = k_perp index).
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
                           const double* pf, const double* D, const double* w, int l0, int l_step, int nl,
                           int nz, int zint, double* out_cl, int lower_only, int variant, void* stream);
/* variant: 0 = the row-weight kernel (band sums over the distinct table rows of a channel pair's y window + exact
 * cell-boundary corrections; 2-3x faster for narrow channels, >= 512 channels over the band), 1 = the
 * per-sample-pair kernel (faster when the y window is taller than ~64 rows, e.g. 256 channels over 400 MHz).
 * The two agree to a few ulp of each row's scale.  The host picks from the comoving width of a channel.
 * lower_only != 0: only the entries (i, j <= i) are written (what the root stage reads, LAPACK-style); the
 * fill's 8-byte stores then merge into full 32-byte sectors.  cora_b200_cl_symmetrize mirrors the lower
 * triangle of every matrix into its upper triangle (cl[nl][nz][nz], in place).                        */
int cora_b200_cl_symmetrize(double* cl, int nl, int nz, void* stream);

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
 * truncate=False).  cl / root: [nl][nz][nz].  used_eigh / num_pos: int[nl]: used_eigh[l] = 0 for a
 * Cholesky (lower-triangular) root, else 1 + the number of leading all-zero columns of the fallback root
 * (= 1 + nz - num_pos[l]; the apply kernel skips those columns).
 * workspace: cora_b200_root_workspace_bytes(nl, nz).                                      */
long long cora_b200_root_workspace_bytes(int nl, int nz);
int cora_b200_root_batched(const double* cl, int nl, int nz, double jitter_rel, double clip_rel,
                           double* root, int* used_eigh, int* num_pos, void* workspace,
                           long long ws_bytes, void* stream);

/* Root of a block-diagonal covariance blockdiag(cl_blocks[0][l], ..., cl_blocks[nblocks-1][l], 0, ...) kept as
 * separate blocks (makesky gaussianfg --pol full builds blockdiag(T, E, B, V) densely, (4 nfreq)^2 per l:
 * cora/scripts/makesky.py:368-382) with the semantics the reference applies to the WHOLE matrix:
 *   - jitter = jitter_rel * max over every block's diagonal                    (cora/core/skysim.py:116-117)
 *   - Cholesky of every block; if ANY block of an l fails, every block of that l takes the eigen branch
 *     (scipy's cholesky of the whole matrix would have failed)                 (cora/util/nputil.py:81-86)
 *   - eigenvalues below clip_rel * (largest eigenvalue over ALL blocks) are zeroed  (nputil.py:89-90)
 * cl_blocks / root_blocks: HOST arrays of nblocks device pointers, each [nl][nz][nz]; the same pointer may
 * not appear twice in root_blocks.  used_eigh / num_pos: device int[nblocks][nl] (meaning as above, per block).
 * zero_block_scale (device float64[nl] or NULL): s with root = s I for an implicit all-zero block (Stokes V):
 * sqrt(jitter) on the Cholesky branch and where the jitter survives the clip, else 0.
 * Fallback implementation: nz <= 128 one-sided Jacobi on the full matrix; larger matrices first try the
 * low-rank route (pivoted Cholesky down to the jitter level, certified residual, one-sided Jacobi on the
 * factor's columns: the same eigenpairs in O(nz^2 rank)) and fall back to the full Jacobi when the matrix
 * is not numerically low-rank positive semi-definite or when the jitter itself would survive the clip.
 * workspace: cora_b200_root_multi_workspace_bytes(nblocks, nl, nz) (smaller: the fallback runs in waves guarded by
 * the device-side failure count -- the host never synchronises).  cora_b200_diag_max: dmax[l] = max_i cl[l][i][i], merged (max) into dmax when merge != 0. */
int cora_b200_diag_max(const double* cl, int nl, int nz, double* dmax, int merge, void* stream);
long long cora_b200_root_multi_workspace_bytes(int nblocks, int nl, int nz);
int cora_b200_root_batched_multi(const double* const* cl_blocks, int nblocks, int nl, int nz, double jitter_rel,
                                 double clip_rel, double* const* root_blocks, int* used_eigh, int* num_pos,
                                 double* zero_block_scale, void* workspace, long long ws_bytes, void* stream);

/* Batched symmetric eigen-decomposition in scipy.linalg.eigh's layout (lower triangle read):
 * evals[l][k] ascending, column k of evecs[l] the unit eigenvector of evals[l][k].
 * replaces: la.eigh(corr[i]) in mkconstrained (cora/core/skysim.py:183-185).  One-sided Jacobi,
 * one CTA per matrix; a smaller workspace is processed in waves.                           */
long long cora_b200_eigh_workspace_bytes(int nl, int nz);
int cora_b200_eigh_batched(const double* a, int nl, int nz, double* evecs, double* evals, void* workspace,
                           long long ws_bytes, void* stream);

/* ------------------------------------------------------------------ draw + apply - */
/* alm[nu, l, m] = sum_nu' M_l[nu, nu'] g_l[nu', m], written in PANEL layout.
 * replaces: complex_std_normal + np.dot + scatter (cora/core/skysim.py:119-121,
 * cora/util/nputil.py:104-125).
 *   root[nl][nz][nz]   roots in the order of l_list_h
 *   l_list_h[nl]       HOST array: global l of each root (any subset -> l-sharding)
 *   dense_flag[nl]     device int per root (cora_b200_root_batched's used_eigh): 0 = lower-triangular
 *                      (Cholesky), the zero half of the contraction is skipped; 1 = dense; v > 1 = dense
 *                      with v - 1 leading zero columns, which are skipped; NULL = all dense
 *   gauss == NULL      draws come from Philox4x32-10 keyed by `seed`, counter (l, m, nu', 0),
 *                      Box-Muller, (N + iN)/sqrt(2)
 *   gauss != NULL      injected draws, complex128 gauss[i][nu'][gauss_ld] (i = position in
 *                      l_list_h, columns m <= l used): the identical-draw parity path
 * Output rows nu in [nu0, nu0+nnu) go to alm_panel[idx(l,m) * panel_stride + chan0 + (nu-nu0)].
 * The workspace holds the generated draws of a batch of l's; any size that fits one l works. */
long long cora_b200_draw_apply_workspace_bytes(int nz, int lmax_in_batch, int nl_batch);
/* The draws alone, for overlap with the C_l fill and the root on another stream (they do not depend on C_l):
 * complex128 variates for the l's of l_list_h, block i = [nz][l_i + 1] (row nu', m contiguous), blocks back to
 * back; cora_b200_draw_bytes gives the size.  Pass the buffer to cora_b200_draw_apply* as `gauss` with
 * gauss_ld = -1 (same l_list): the result is bit-identical to letting draw_apply draw by itself.           */
long long cora_b200_draw_bytes(const int* l_list_h, int nl, int nz);
int cora_b200_draw(const int* l_list_h, int nl, int nz, unsigned long long seed, int draw_counter0,
                   void* gauss_packed, long long bytes, void* stream);
int cora_b200_draw_apply(const double* root, const int* l_list_h, const int* dense_flag, int nl, int nz,
                         int lmax, unsigned long long seed, const void* gauss, long long gauss_ld,
                         void* alm_panel, long long panel_stride, int chan0, int nu0, int nnu,
                         void* workspace, long long ws_bytes, void* stream);

/* ------------------------------------------------------------------ multi-GPU --- */
/* The l-sharded half of mkfullsky on one GPU of G (cora/core/skysim.py:108-121 for the local
 * l's of alm_array.enumerate(axis=2)), writing straight into the send buffer of the l -> nu
 * all-to-all (alm_array.redistribute(axis=0), cora/core/skysim.py:128): element (l, m, nu)
 * goes to send[nu_base[nu] + (row0_h[i] + m) * nu_width[nu]] (complex elements), i = position
 * of l in l_list_h.  With nu_base[nu] = slab_offset(dest(nu)) + (nu - first_nu(dest)) and
 * nu_width[nu] = channels owned by dest(nu), the buffer is one contiguous [rows][channels]
 * slab per destination GPU.  nu_base / nu_width: device arrays [nz]; row0_h: host [nl].      */
int cora_b200_draw_apply_slabs(const double* root, const int* l_list_h, const int* dense_flag, int nl, int nz,
                               int lmax, unsigned long long seed, const void* gauss, long long gauss_ld,
                               const long long* row0_h, const long long* nu_base, const int* nu_width,
                               void* send, void* workspace, long long ws_bytes, void* stream);

/* After the all-to-all: receive buffer -> PANEL rows idx(l, m) of this GPU's `nchan` channels.
 * l_off[lmax+1] (device): complex offset of row (l, m = 0) in `recv`; the rows m = 0..l of one
 * l are consecutive, nchan channels each (the unpack half of redistribute + pack_alm,
 * cora/core/skysim.py:128 + cora/util/hputil.py:124-152).                                   */
int cora_b200_alm_slabs_to_panel(const void* recv, const long long* l_off, int lmax, int nchan, void* alm_panel,
                                 long long panel_stride, int chan0, void* stream);

/* ------------------------------------------------------------------ forward SHT --- */
/* Scalar analysis of `nchan` RING maps: one HEALPix quadrature pass
 *   a_lm = (4 pi / npix) sum_rings w_ring lambda_lm(theta_ring) sum_j map(ring, j) exp(-i m phi_j)
 * written to (accumulate = 0) or added to (accumulate = 1) the PANEL array
 * alm_panel[idx(l, m) * panel_stride + chan0 + c].  healpy.map2alm's `iter` Jacobi refinements
 * are a += A(map - S a): cora_b200_alm2map + cora_b200_map_sub + this with accumulate = 1.
 * replaces: healpy.map2alm at cora/util/hputil.py:228-230 (sphtrans_real), :310-321
 * (sphtrans_real_pol, T and V), used by sphtrans_sky (:460-497) and sph_ps (:607-619).
 * ring_weights: device float64[2 nside] absolute weights of the northern rings incl. the equator
 * (healpy's use_weights=True tables are data files of the healpy distribution), or NULL = 1.   */
long long cora_b200_map2alm_workspace_bytes(void* plan, int nchan_batch);
int cora_b200_map2alm(void* plan, const double* map, int nchan, const double* ring_weights, int accumulate,
                      void* alm_panel, long long panel_stride, int chan0, void* workspace, long long ws_bytes,
                      void* stream);
/* Polarised analysis (Q, U) -> (aE, aB), the adjoint of cora_b200_alm2map_spin2 times 4 pi / npix:
 *   aE = -sum_r w (X1 Q_m + i X2 U_m),  aB = -sum_r w (X1 U_m - i X2 Q_m).
 * replaces: the E/B part of healpy.map2alm([T, Q, U]) at cora/util/hputil.py:310-312
 * (sphtrans_real_pol); T and V go through cora_b200_map2alm.                                   */
long long cora_b200_map2alm_spin2_workspace_bytes(void* plan, int nchan_batch);
int cora_b200_map2alm_spin2(void* plan, const double* mapQ, const double* mapU, int nchan,
                            const double* ring_weights, int accumulate, void* almE_panel, void* almB_panel,
                            long long panel_stride, int chan0, void* workspace, long long ws_bytes, void* stream);
/* out[i] = a[i] - b[i], i < n (the residual map of the refinement) */
int cora_b200_map_sub(const double* a, const double* b, long long n, double* out, void* stream);

/* ---- fused exchange over peer memory (NVLink / NVSwitch, one process per GPU) ------------
 * The reference moves data between ranks with caput's MPIArray.redistribute
 * (cora/core/skysim.py:128) after the compute.  Here the producing kernels store straight into
 * the consumer GPU's buffer, so the exchange happens in the kernel epilogue:
 *   - cora_b200_cl_fill_21cm_tiles: the 21cm C_l fill sharded over channel PAIRS (its cost per
 *     pair does not shrink with the l range, so l-sharding it would not scale); row l of the
 *     result is written to the GPU that owns l for the root/apply stage.
 *   - cora_b200_draw_apply_peers: apply writes a_lm(nu) into the PANEL buffer of the GPU that
 *     owns channel nu for the SHT stage.
 * Buffers come from cora_b200_peer_alloc (cudaMalloc + CUDA IPC handle, zero-filled); every
 * other process maps them with cora_b200_peer_open.  cora_b200_peer_barrier is a flag barrier
 * over the same memory (stream-ordered; `epoch` strictly increasing).  A peer that does not arrive
 * within timeout_s (<= 0: 60 s) sets *status = 1 + its rank instead of hanging the GPU and, with
 * fatal != 0, traps: the kernels queued behind the barrier must not run on half-written buffers, so
 * every later CUDA call of the process fails.  With fatal == 0 the caller must read *status before
 * trusting anything computed after the barrier.                                                 */
int cora_b200_peer_alloc(long long bytes, void** ptr_out, unsigned char* handle64_out);
int cora_b200_peer_free(void* ptr);
int cora_b200_peer_open(const unsigned char* handle64, void** ptr_out);
int cora_b200_peer_close(void* ptr);
/* flags_ptrs: device array [size] of pointers, entry r = rank r's flag array (u64[size]) as mapped here */
int cora_b200_peer_barrier(const void* flags_ptrs, int rank, int size, unsigned long long epoch,
                           double timeout_s, int* status, int fatal, void* stream);

/* 21cm fill for the channel-pair tiles tile0, tile0 + tile_step, ... (ntiles of them) of the
 * cora_b200_cl_fill_21cm_ntiles(nz) tiles (a tile = channel i against 4 consecutive channels j0 .. j0+3 <= i; tiles
 * ordered by i - j0, then j0), all l = 0..nl-1, lower triangle only.  Ranks take tiles INTERLEAVED (tile0 = rank,
 * tile_step = size): a tile's cost grows with |chi_i - chi_j|, i.e. along the enumeration, so contiguous ranges are
 * unbalanced (measured: 29.3 ms on each of two GPUs against 42.6 ms on one).  out_ptrs: device array of per-GPU C_l
 * buffers; element (l, i, j) goes to out_ptrs[l_owner[l]][(l_row[l] * nz + i) * nz + j], one 32-byte run per (l, tile)
 * (over NVLink for remote owners).  l_owner / l_row: device int[nl]. */
long long cora_b200_cl_fill_21cm_ntiles(int nz);
int cora_b200_cl_fill_21cm_tiles(const double* tab, const double* chi, const double* b, const double* f,
                                 const double* pf, const double* D, const double* w, int nl, int nz, int zint,
                                 long long tile0, long long ntiles, int tile_step, int variant, const void* out_ptrs,
                                 const int* l_owner, const int* l_row, void* stream);

/* draw + apply for the local l's with the exchange fused into the epilogue: element (l, m, nu)
 * is stored at nu_ptr[nu][idx(l, m) * nu_width[nu]] (complex elements), where nu_ptr[nu] (device
 * array [nz] of pointers) is the PANEL buffer of the GPU owning channel nu advanced to that
 * channel's column and nu_width[nu] that GPU's channel count.  draw_counter0 offsets the nu'
 * word of the Philox counter (block b of a block-diagonal covariance draws with b * nz, so the
 * blocks see independent streams -- the same draws as the dense (npol nz)^2 formulation).  Other
 * arguments as cora_b200_draw_apply.                                                                     */
int cora_b200_draw_apply_peers(const double* root, const int* l_list_h, const int* dense_flag, int nl, int nz,
                               int lmax, unsigned long long seed, int draw_counter0, const void* gauss,
                               long long gauss_ld, const void* nu_ptr, const int* nu_width, void* workspace,
                               long long ws_bytes, void* stream);

/* alm_panel[idx(l, m)][chan0 + c] *= scale[l * nchan + c] (device float64[(lmax+1) * nchan]): a per-(l, channel) factor
 * on every a_lm -- the Gaussian beam exp(-l (l+1) sigma^2 / 2) of healpy.smoothing as ConstrainedGalaxy calls it
 * (cora/foreground/galaxy.py:165-185), between cora_b200_map2alm and cora_b200_alm2map.                          */
int cora_b200_alm_scale_l(void* alm_panel, long long panel_stride, int chan0, int nchan, int lmax, const double* scale,
                          void* stream);

/* Which Legendre kernels the inverse SHT launches: bit 0 = warp-specialised scalar kernel, bit 1 = warp-specialised
 * spin-2 kernel (default 3, or the CORA_B200_LEGENDRE_WS environment variable at load); cleared bits select the
 * single-role kernel.  Returns the previous mask; mask < 0 only queries.  (A/B measurements and tests.)          */
int cora_b200_set_legendre_ws(int mask);

/* ---- xi(r) -> C_l(chi, chi') front end --------------------------------------------------------
 * replaces: legendre_array + the weighted Legendre contraction of corr_to_clarray
 *           (cora/signal/corrfunc.py:265-287, :384-396).
 * legendre_table: out[l * ld + i] = P_l(mu[i]) * scale[i] (scale may be null), l = 0..lmax, i < n, by the
 * three-term recurrence scipy.special.lpn uses.  mu, scale, out: device pointers.
 * dgemm: C[m x n] (+)= A[m x k] B[k x n], row-major FP64 on the tensor cores (DMMA.8x8x4); accumulate != 0 adds to C
 * (corr_to_clarray feeds the Gauss-Legendre nodes chunk by chunk: np.dot(lm, corr_array), corrfunc.py:395).   */
/* corr_bins: the bin-averaged integrand for a TABULATED correlation function (piecewise linear in r, or in ln r when
 * logx != 0; clamped at both ends like numpy.interp), fused with the cosine rule and the two-sided radial quadrature:
 * out[i][a][b] = sum_{p,q} xw[p] xw[q] xi(|x_ap - x_bq| at angle mu_i), xa[nx * xint] the radial samples, i < nmu
 * (corrfunc.py:369-379 for callables of that form).  All pointers are device pointers.                          */
int cora_b200_corr_bins(const double* mu, int nmu, const double* xa, const double* xw, int nx, int xint,
                        const double* tab_r, const double* tab_v, int nt, int logx, double* out, void* stream);
int cora_b200_legendre_table(const double* mu, const double* scale, int n, int lmax, double* out, long long ld,
                             void* stream);
int cora_b200_dgemm(const double* A, const double* B, double* C, int m, int n, int k, long long lda, long long ldb,
                    long long ldc, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CORA_B200_H */
