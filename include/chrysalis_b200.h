/*
 * chrysalis_b200 - C ABI of the B200-native (sm_100a) hot path of rockdeme/chrysalis:
 *   detect_svgs (Moran's I)  ->  pca  ->  aa
 *
 * The reference (pure Python, /root/reference/chrysalis/core.py) has no FFI; its boundary
 * for this path is three Python callables and the third-party numeric calls inside them.
 * Each entry point below replaces one of those numeric call sites; the Python host
 * (chrysalis_b200/core.py) keeps the reference signatures and calls these through ctypes.
 *
 * Conventions
 *   - plain C types only; every buffer is a caller-owned DEVICE pointer unless a name says host
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), returns
 *     0 on success, non-zero on failure; chr_last_error() gives the message (thread-local)
 *   - no hidden global state, no allocation on the hot path: scratch is a caller-provided
 *     workspace whose size comes from the matching *_workspace_bytes() query
 *   - matrices are row-major; `ld` is the row stride in elements
 */
#ifndef CHRYSALIS_B200_H
#define CHRYSALIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* chr_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------ */
int chr_version(void);
const char* chr_last_error(void);
/* device time of the dominant kernel of the most recent call of each family that was made
 * with CHR_FLAG_TIMING (two cudaEvents around that kernel); synchronises on those events. */
float chr_last_kernel_ms(int family);
#define CHR_FAMILY_MORAN 0
#define CHR_FAMILY_GRAM 1
#define CHR_FAMILY_AA 2
#define CHR_FAMILY_KNN 3
#define CHR_FAMILY_PREP 4

#define CHR_FLAG_GEARY 0x1u   /* also accumulate Geary's C                                  */
#define CHR_FLAG_TIMING 0x2u  /* record events around the dominant kernel                   */
#define CHR_FLAG_PRESHIFTED 0x4u /* Moran: columns of X were already shifted to ~zero mean (no in-kernel shift) */
#define CHR_VARIANT_SHIFT 8   /* bits 8..15 of `flags`: kernel variant (0 = default; tuning)  */

/* ---- host side: AnnData's CSR arrays from PAGEABLE host memory to the device -----------------
 * the reference passes adata.X (scipy CSR in numpy arrays) to CPU libraries (core.py:53-59, :126); here it crosses PCIe.
 * A pool of host threads (CHR_HOST_THREADS, default all cores, <= 32) fills a ring of pinned bounce buffers
 * (4 x 32 MB, allocated once) while earlier chunks are in flight.  All pointers named *_host / src_host are HOST
 * pointers; the calls return once the last chunk has been queued on `stream`. */
int chr_host_threads(void);
int chr_host_upload(void* dst_device, const void* src_host, size_t bytes, chr_stream_t stream);
/* the way back for the large results (.obsm['chr_X_pca'], .obsm['chr_aa']; core.py:128,193): device -> host array through the
 * same ring; synchronous (returns when dst_host holds the data), the pool threads first-touch a fresh destination in parallel */
int chr_host_download(void* dst_host, const void* src_device, size_t bytes, chr_stream_t stream);
/* column selection while uploading (pca densifies the SVG columns only, core.py:126; a rank's gene shard under
 * multi-GPU detect_svgs): col_map_host[c] = new column index or -1.  count: sel_indptr_host [n + 1] int64 and the
 * selected nnz; upload: new column indices (int32) and values (float32) of the selected entries, row by row, into
 * device arrays of that length.  data_kind: 0 float32, 1 int32, 2 float64 values in the source. */
int chr_host_csr_select_count(const int64_t* indptr_host, const int32_t* indices_host, int64_t n, int64_t n_genes,
                              const int32_t* col_map_host, int64_t* sel_indptr_host, int64_t* nnz_out_host);
int chr_host_csr_select_upload(const int64_t* indptr_host, const int32_t* indices_host, const void* data_host,
                               int data_kind, int64_t n, const int32_t* col_map_host, const int64_t* sel_indptr_host,
                               int32_t* out_indices_dev, float* out_data_dev, chr_stream_t stream);

/* ---- K1: spatial weights -------------------------------------------------------------
 * replaces  weights.KNN.from_array(points, k=neighbors); w.transform = 'R'
 *           (/root/reference/chrysalis/core.py:64-65; libpysal -> scipy cKDTree.query(k+1))
 * xy: [n][2] float64 coordinates (the caller has already negated y, core.py:62).
 * nbr_out: [n][k] int32, the k nearest other spots of every spot in increasing distance
 * (ties: lower index first).  Row-standardised W is implicit: every weight is 1/k. */
/* bbox_out: device [4] float64 = {min_x, min_y, max_x, max_y} */
int chr_spatial_bbox(const double* xy, int64_t n, double* bbox_out, chr_stream_t stream);
size_t chr_knn_workspace_bytes(int64_t n, int k);
int chr_knn_build(const double* xy, int64_t n, int k, const double* bbox, int32_t* nbr_out,
                  void* workspace, size_t workspace_bytes, chr_stream_t stream);

/* Hilbert-curve key (16 bits per axis over bbox) of every spot: sorting spots by it makes the
 * rows of the dense matrix spatially coherent, so neighbour rows share L1/L2 lines.
 * keys_out: [n] int64. */
int chr_hilbert_keys(const double* xy, int64_t n, const double* bbox, int64_t* keys_out,
                     chr_stream_t stream);

/* ---- K2: preprocessing ---------------------------------------------------------------
 * replaces  sc.pp.filter_genes(min_cells) / sc.pp.normalize_total / sc.pp.log1p / ad.to_df()
 *           (/root/reference/chrysalis/core.py:53-59)
 * CSR counts (spots x genes): indptr [n+1] int64, indices [nnz] int32, data [nnz] float32. */
int chr_csr_gene_nnz(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                     int64_t n_genes, int32_t* gene_nnz /* [n_genes], zeroed by callee */,
                     chr_stream_t stream);
/* compact counts as a loader can deliver them (3 bytes per stored entry instead of 8 across PCIe): column indices as
 * uint16 (n_genes <= 65536), values as uint8 with 255 = "look in the overflow list" (over_pos: positions into the value
 * array, over_val: their values; counts >= 255 are rare).  Expands to the int32 / float32 arrays of the calls below. */
int chr_csr_decode_compact(const uint16_t* cols, const uint8_t* vals, int64_t nnz, const int64_t* over_pos,
                           const float* over_val, int64_t n_over, int32_t* out_indices, float* out_data,
                           chr_stream_t stream);
/* the same with delta-coded columns (2 bytes per stored entry): columns sorted inside a row, deltas[p] = column - previous
 * column of the row (the first entry of a row counts from 0), 255 = "look in the gap list" (gap_pos: sorted positions,
 * gap_val: the true deltas; at ~5 % density deltas >= 255 occur once in a million entries).  indptr: device, [n_rows + 1]. */
int chr_csr_decode_compact_delta(const int64_t* indptr, const uint8_t* deltas, const uint8_t* vals, int64_t n_rows,
                                 int64_t nnz, const int64_t* gap_pos, const int32_t* gap_val, int64_t n_gap,
                                 const int64_t* over_pos, const float* over_val, int64_t n_over, int32_t* out_indices,
                                 float* out_data, chr_stream_t stream);
/* per-spot totals over the kept genes (gene_col[g] >= 0), double accumulation, one warp per spot */
int chr_csr_spot_totals(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                        const int32_t* gene_col, double* totals, chr_stream_t stream);
/* dense[row_of_spot[i]][gene_col[g]] = f(data * scale[i]),  f = log1p if apply_log1p else identity,
 * product formed in float64 and rounded once to float32 (scanpy normalize_total semantics);
 * scale == NULL: no scaling.  dense is [n][ld] float32 and is zero-filled by the callee;
 * row_of_spot may be NULL (identity); genes with gene_col[g] < 0 are dropped.
 * apply_log1p: bit 0 = apply log1p, bit 1 = `dense` was already zero-filled by the caller (lets the
 * 4 n ld byte fill overlap the host-to-device copy of the counts on another stream). */
int chr_csr_densify_log1p(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                          const int32_t* gene_col, const double* scale, const int64_t* row_of_spot,
                          float* dense, int64_t ld, int apply_log1p, chr_stream_t stream);

/* ---- K1b: W as a device plan ------------------------------------------------------------
 * the device representation of the row-standardised kNN weights (core.py:64-65) that the
 * Moran kernel streams against: every unordered pair {i, j} once with coefficient
 * a_ij + a_ji, grouped by quads of 4 consecutive spots, each pair owned by one endpoint, the
 * owners balanced, entries split into rows inside / outside the quad's 112-spot tile.
 * Built once per W (k <= 64; any list: duplicates, self loops, asymmetric kNN).
 * plan: device buffer of chr_moran_plan_bytes(), 256-byte aligned.  info_host (HOST, int32[4],
 * may be NULL): {largest tile slice in 16-byte units (pass to chr_moran_dense_f32), total units,
 * quads, 0}; reading it back synchronises the stream. */
size_t chr_moran_plan_bytes(int64_t n_spots, int k);
size_t chr_moran_plan_workspace_bytes(int64_t n_spots, int k);
int chr_moran_plan_build(const int32_t* nbr, int64_t n_spots, int k, void* plan, size_t plan_bytes,
                         int32_t* info_host, void* workspace, size_t workspace_bytes,
                         chr_stream_t stream);

/* ---- K3/K4: Moran's I (and Geary's C) ------------------------------------------------
 * replaces  the gene loop  esda.moran.Moran(gene_matrix[c], w, permutations=0).I
 *           [+ esda.geary.Geary(...).C]      (/root/reference/chrysalis/core.py:71-76)
 * X: [n_spots][ld] float32 dense expression (ld % 4 == 0, X 16-byte aligned, columns
 *    n_genes..round_up(n_genes,4)-1 readable), nbr: [n_spots][k] int32 (K1 output or any kNN
 *    list), plan + plan_chunk_units: chr_moran_plan_build's output for the same nbr.
 *    With a plan the tile kernel runs (cp.async producer warps + mbarrier pipeline; nbr may be NULL); without one (plan NULL, any k) the
 *    per-spot stream kernel runs from nbr.
 * row_map: NULL, or int32 [n_spots]: spot i takes row row_map[i] of X while W stays in place.  One
 *    call with a random permutation is one draw of esda's permutation inference
 *    (esda.moran.Moran(..., permutations=P): I of np.random.permutation(z)) for every gene at once.
 * out_I / out_C: [n_genes] float64 (out_C may be NULL without CHR_FLAG_GEARY).
 * out_stats: optional [4][n_genes] float64 = {mean, z'z, z'Wz, geary numerator}.
 * One pass over X; per-gene accumulators fp32 over <= 8/16 spots, fp32 per CTA, then fp64. */
size_t chr_moran_workspace_bytes(int64_t n_spots, int64_t n_genes, int k, unsigned flags);
int chr_moran_dense_f32(const float* X, int64_t n_spots, int64_t n_genes, int64_t ld,
                        const int32_t* nbr, int k, const void* plan, int plan_chunk_units,
                        const int32_t* row_map, unsigned flags, double* out_I, double* out_C,
                        double* out_stats, void* workspace, size_t workspace_bytes,
                        chr_stream_t stream);

/* ---- K3/K4 with a block-diagonal W: several samples in one launch ----------------------------
 * replaces  the per-sample  detect_svgs(ad, **kwargs)  loop of  chrysalis.integrate_adatas
 *           (/root/reference/chrysalis/utils.py:210-218): Moran's I is computed PER SAMPLE with that sample's own W,
 *           mean and z'z; here the samples are row blocks of one matrix, W is block diagonal, one launch scores all.
 * X: [n_rows][ld] float32; block b owns rows [block_rows[b], block_rows[b+1]), starts on a multiple of
 *    CHR_BLOCK_ALIGN_ROWS and holds block_n[b] spots followed by padding rows (listed in pad_rows; the call overwrites them
 *    with the per-gene shift value).  plan: chr_moran_plan_build of the block-diagonal neighbour lists over all n_rows rows
 *    (padding rows list themselves k times).  block_rows_host / block_n_host: HOST int64 [n_blocks + 1] / [n_blocks].
 * out_I / out_C: [n_blocks][n_genes] float64.  Synchronises the stream once (two small host arrays are uploaded). */
#define CHR_BLOCK_ALIGN_ROWS 448
size_t chr_moran_blockdiag_workspace_bytes(int64_t n_rows, int64_t n_genes, int n_blocks);
int chr_moran_blockdiag_f32(float* X, int64_t n_rows, int64_t n_genes, int64_t ld, int k, const void* plan,
                            int plan_chunk_units, const int64_t* block_rows_host, const int64_t* block_n_host,
                            int n_blocks, const int32_t* pad_rows, int64_t n_pad, unsigned flags, double* out_I,
                            double* out_C, void* workspace, size_t workspace_bytes, chr_stream_t stream);

/* ---- K3/K4 on an exact square lattice (stencil kernel) -----------------------------------
 * Same statistic, same reference call sites as chr_moran_dense_f32, for spots that sit on a square lattice with the
 * 8-neighbour kNN graph (core.py:64-65 with neighbors=8: Visium HD / Stereo-seq bins).  The kNN adjacency is split as
 * A = B + D: B = the symmetric 8-neighbour stencil between present lattice nodes (evaluated in registers while the
 * matrix streams through shared memory once), D = a short list of corrections at tissue edges and holes.
 * X: [n_rows][ld] float32 in LATTICE LAYOUT (device.lattice_plan): bands of <= CHR_LATTICE_BAND_ROWS lattice rows; inside
 *    a band column after column, each column the band's rows as consecutive matrix rows; n_rows >= n_spots (absent nodes
 *    of a stored column are filler rows, listed in absent_rows - the call overwrites them with the per-gene shift value).
 * bands:  int32 [n_bands][4]  = {first matrix row, lattice x of the first stored column, stored columns, lattice rows}
 * chunks: int32 [n_chunks][4] = {band, first column relative to the band's first stored column, columns, 0} (one CTA each)
 * corr:   int32 [n_corr][4]   = {row_i, row_j, coefficient as float bits, kind}; kind 0: x'Ax' += c x_i x_j,
 *                                kind 1: sum_j (indeg_j - k) x_j += c x_i  (column sums of A differ from k at the edges)
 * last_band_rows: lattice rows of the last band (every other band has CHR_LATTICE_BAND_ROWS); with it the columns are
 *   fetched by TMA (cp.async.bulk.tensor, one box per staged column); 0 = unknown: cp.async producer warps instead
 *   (also selected by kernel variant 2 in `flags`).  Both give bit-identical results.
 * pilot_rows: int32 [n_pilot <= 512] matrix rows of present spots whose mean is the per-gene shift.
 * n_spots = number of present spots (the n of the statistic).  Outputs as chr_moran_dense_f32. */
#define CHR_LATTICE_BAND_ROWS 60
size_t chr_moran_lattice_workspace_bytes(int64_t n_genes, int n_chunks, int64_t n_corr);
int chr_moran_lattice_f32(float* X, int64_t n_rows, int64_t n_genes, int64_t ld, int64_t n_spots, int k,
                          const int32_t* bands, int n_bands, int last_band_rows, const int32_t* chunks, int n_chunks,
                          const int32_t* corr, int64_t n_corr, const int32_t* absent_rows, int64_t n_absent,
                          const int32_t* pilot_rows, int n_pilot, unsigned flags, double* out_I, double* out_C,
                          double* out_stats, void* workspace, size_t workspace_bytes, chr_stream_t stream);

/* The same statistic fed from the CSR COUNTS instead of a dense matrix (north_star: "streams the spots x genes expression
 * matrix (CSR counts or dense log1p)"): fuses filter_genes' column map, normalize_total's per-spot scale, log1p and
 * ad.to_df() (core.py:53-59) into the Moran kernel's producer - the tiles are built in shared memory from the stored
 * entries, no dense matrix exists, HBM sees ~5 bytes per stored entry instead of 4 bytes per spot and gene.
 * indptr / indices / data: device CSR in SPOT order with sorted column indices inside every row, nnz < 2^32;
 * gene_col [n_genes_all]: new column of a kept gene or -1; scale [n_spots] float64 or NULL; apply_log1p as in
 * chr_csr_densify_log1p.  Lattice tables as chr_moran_lattice_f32 plus spot_of_row [n_rows] (-1 = absent node).
 * unsorted_flag (device int32, may be NULL): set to 1 when a row's columns are not strictly increasing - the results are
 * then meaningless and the caller must take the dense path (checked while the entries are prepared, no extra pass). */
size_t chr_moran_lattice_csr_workspace_bytes(int64_t nnz, int64_t n_rows, int64_t n_keep, int n_chunks, int64_t n_corr);
int chr_moran_lattice_csr_f32(const int64_t* indptr, const int32_t* indices, const float* data, int64_t nnz,
                              int64_t n_spots, int64_t n_genes_all, const int32_t* gene_col, int64_t n_keep,
                              const double* scale, int apply_log1p, int64_t n_rows, int k, const int32_t* bands,
                              int n_bands, const int32_t* chunks, int n_chunks, const int32_t* corr, int64_t n_corr,
                              const int32_t* spot_of_row, const int32_t* pilot_rows, int n_pilot, unsigned flags,
                              double* out_I, double* out_C, double* out_stats, int32_t* unsorted_flag,
                              void* workspace, size_t workspace_bytes, chr_stream_t stream);

/* ---- K3p: permutation inference ------------------------------------------------------------
 * replaces  esda.moran.Moran(y, w, permutations=P)'s simulation (np.random.permutation(z), P times per gene;
 *           /root/reference/article/A2_human_lymph_node/morans_i.ipynb:186; core.py:72 hard-codes permutations=0)
 * Draws are counter-based: pi_draw(i) is a keyed Feistel function of (seed, draw, i) (csrc/perm.cuh), evaluated on the
 * device; a draw permutes the spots against the fixed W and is shared by all genes.
 * chr_perm_row_map: pi_draw as a row map for chr_moran_dense_f32 (row_map_out: device int32 [n]).
 * chr_moran_perm_lattice_f32 (square lattice, k = 8; arguments as chr_moran_lattice_f32 plus the two maps of the lattice
 *   layout: spot_of_row [n_rows] (-1 for filler rows) and row_of_spot [n_spots]): the observed I and, folded over all
 *   P draws, out_larger[g] = #{I_sim >= I}, out_sum[g] = sum I_sim, out_sumsq[g] = sum I_sim^2 - all that p_sim,
 *   EI_sim, VI_sim and z_sim need.  Batches of draws are launched with the gene slab as the slowest grid dimension so a
 *   slab that fits L2 is read from HBM once per batch. */
int chr_perm_row_map(int64_t n, uint64_t seed, int draw, int32_t* row_map_out, chr_stream_t stream);
size_t chr_moran_perm_lattice_workspace_bytes(int64_t n_rows, int64_t n_genes, int n_chunks, int64_t n_corr,
                                              int permutations);
int chr_moran_perm_lattice_f32(float* X, int64_t n_rows, int64_t n_genes, int64_t ld, int64_t n_spots, int k,
                               const int32_t* bands, int n_bands, const int32_t* chunks, int n_chunks,
                               const int32_t* corr, int64_t n_corr, const int32_t* absent_rows, int64_t n_absent,
                               const int32_t* pilot_rows, int n_pilot, const int32_t* spot_of_row,
                               const int32_t* row_of_spot, int permutations, uint64_t seed, unsigned flags,
                               double* out_I, int64_t* out_larger, double* out_sum, double* out_sumsq,
                               void* workspace, size_t workspace_bytes, chr_stream_t stream);

/* ---- multi-GPU: all-gather of the per-gene statistics through peer memory -----------------------
 * The one exchange of the gene-sharded Moran path ("a small all-gather of per-gene statistics over NVLink").
 * peer_bufs_dev / signal_pads_dev: device arrays [world] of the peers' receive buffers / uint32 signal pads (a symmetric
 * allocation mapped into every process, e.g. torch.distributed._symmetric_memory).  Every rank writes `count` doubles of
 * `local` to element offset dst_elem0 of EVERY peer's buffer, raises slot signal_slot0 + rank of every peer's pad to
 * `epoch` (non-zero, increasing per call) and waits until its own pad shows `epoch` from every peer.  One launch; ranks must
 * sit on different GPUs. */
int chr_p2p_allgather_f64(const double* local, int64_t count, void* const* peer_bufs_dev, void* const* signal_pads_dev,
                          int rank, int world, int64_t dst_elem0, int signal_slot0, uint32_t epoch,
                          chr_stream_t stream);

/* ---- Local Moran's I (LISA) ---------------------------------------------------------------
 * replaces  esda.moran.Moran_Local(y, w, transformation='r', permutations=P)
 *           (/root/reference/article/A2_human_lymph_node/morans_i.ipynb:150; not called by core.py)
 * z: [n] float32, the variable STANDARDISED by the caller (z = (y - mean) / population std), z2ss = sum z^2.
 * Outputs [n]: lag (W z), Is, quadrant q (1 HH, 2 LH, 3 LL, 4 HL) and, with permutations > 0, the
 * conditional-permutation p_sim, EI_sim, VI_sim per spot (k <= 64). */
int chr_local_moran_f32(const float* z, int64_t n, const int32_t* nbr, int k, double z2ss,
                        int permutations, uint64_t seed, float* lag_out, float* Is_out,
                        int32_t* q_out, float* p_sim_out, float* ei_sim_out, float* vi_sim_out,
                        chr_stream_t stream);

/* ---- K6-K8: PCA of the SVG matrix ------------------------------------------------------
 * replaces  PCA(n_components=n_pcs, svd_solver='arpack', random_state=42).fit_transform(pcs)
 *           (/root/reference/chrysalis/core.py:126-128; sklearn -> scipy ARPACK svds)
 * P: [n][ld] float32 dense SVG matrix (spots x selected genes).  Staged so that a spot-sharded
 * multi-GPU run can all-reduce the column sums and the partial Grams between the calls. */
size_t chr_colsum_workspace_bytes(int64_t n, int64_t s);
/* sums[g] = sum_i P[i][g]  (float64, fixed summation order) */
int chr_colsum_f64(const float* P, int64_t n, int64_t s, int64_t ld, double* sums, void* workspace,
                   size_t workspace_bytes, chr_stream_t stream);
size_t chr_gram_workspace_bytes(int64_t n, int64_t s);
/* C_out[s][s] (float64, symmetric) = (P - mean)^T (P - mean): tcgen05 Gram, operands split into
 * bf16 hi/lo pairs (3 MMAs per product pair, fp32 TMEM accumulators over <= 8192 spots,
 * split-K partials reduced in float64).  mean: [s] float64. */
int chr_gram_centered_bf16x3(const float* P, int64_t n, int64_t s, int64_t ld, const double* mean,
                             double* C_out, unsigned flags, void* workspace, size_t workspace_bytes,
                             chr_stream_t stream);
/* r largest eigenpairs of the symmetric C [s][s]: m-step Lanczos with full reorthogonalisation,
 * Sturm bisection + inverse iteration on the tridiagonal, Ritz vectors orthonormalised.
 * evals: [r] descending Rayleigh quotients, evecs: [r][s], residuals: [r] = |C y - lambda y|.
 * m: Lanczos steps, r <= m <= s (chr_eigh_default_steps gives the starting choice; the caller
 * repeats with a larger m if a residual is too large). */
int chr_eigh_default_steps(int64_t s, int r);
size_t chr_eigh_workspace_bytes(int64_t s, int r, int m);
int chr_eigh_topk_f64(const double* C, int64_t s, int r, int m, double* evals, double* evecs,
                      double* residuals, void* workspace, size_t workspace_bytes, chr_stream_t stream);
/* Y[i][c] = sum_g (P[i][g] - mean[g]) V[c][g]   (scores U*S of the SVD), Y: [n][ldy] float32 */
size_t chr_pca_scores_workspace_bytes(int64_t s, int r);
int chr_pca_scores_f32(const float* P, int64_t n, int64_t s, int64_t ld, const double* mean,
                       const double* V, int r, float* Y, int64_t ldy, void* workspace,
                       size_t workspace_bytes, chr_stream_t stream);

/* ---- K9-K11: archetypal analysis -------------------------------------------------------
 * replaces  archetypes.AA(n_archetypes, n_init=3, max_iter, tol=0.001, random_state=42).fit(X)
 *           (/root/reference/chrysalis/core.py:186-191; archetypes==0.4.2: alternating
 *           penalised NNLS, Cutler & Breiman)
 * X: [n][d] float64 (the first d PCA scores), Z: [n_archetypes][d] float64 archetypes,
 * alphas: [n][n_archetypes] float64.  d <= 64, 2 <= n_archetypes <= 64.  The restart loop,
 * the stopping rule |rss_prev - rss| < tol and the best-of-n_init choice stay on the host
 * (chrysalis_b200/core.py:aa), like they do in the reference's fit(). */
size_t chr_aa_workspace_bytes(int64_t n, int d, int n_archetypes);
/* Z[a] = X[rows[a]]  (one-hot betas0: the initial archetypes are data points); rows: device int64 */
int chr_aa_init_archetypes(const double* X, int64_t n, int d, const int64_t* rows, int n_archetypes,
                           double* Z, chr_stream_t stream);
/* FurthestSum seeding (the initialisation archetypes 0.4.x draws): from the random start `first`
 * greedily pick the point with the largest summed distance to the chosen set, then 10 re-pick
 * rounds.  rows_out: HOST int64[n_archetypes].  Synchronises the stream once per pick. */
size_t chr_aa_furthest_sum_workspace_bytes(int64_t n);
int chr_aa_furthest_sum(const double* X, int64_t n, int d, int n_archetypes, int64_t first,
                        int64_t* rows_out, void* workspace, size_t workspace_bytes,
                        chr_stream_t stream);
/* one iteration, in place:  alphas <- NNLS(X | Z);  Z <- pinv(alphas) X;  betas <- NNLS(Z | X);
 * Z <- betas X;  *rss_out (device) = ||X - alphas Z||_F^2, evaluated as ||X||^2 - 2 <A^T X, Z> + <A^T A, Z Z^T> from the sums
 * the alpha step accumulates anyway (absolute error ~1e-15 ||X||^2, far below the 1e-3 stop rule of the fit; the
 * environment variable CHR_AA_RSS_DIRECT=1 selects the per-spot residual pass instead).  beta_passes_out: HOST int (may be
 * NULL), number of active-set passes of the beta step.  Synchronises the stream (the beta
 * step's termination flag is read back).  warm != 0: `alphas` and the workspace still hold the
 * previous iteration of the same restart; their supports seed the active sets (the NNLS
 * minimisers are unique, so the result is the same - it is only reached in fewer steps). */
int chr_aa_iterate(const double* X, int64_t n, int d, int n_archetypes, double penalty_M, int warm,
                   double* Z, double* alphas, double* rss_out, int* beta_passes_out, void* workspace,
                   size_t workspace_bytes, chr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CHRYSALIS_B200_H */
