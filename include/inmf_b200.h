/*
 * inmf_b200.h -- C ABI of libinmf_b200.so: hand-written sm_100a kernels for pyliger's integrative-NMF
 * hot path (optimize_ALS, online_iNMF, nnlsm_blockpivot).
 *
 * pyliger itself has no FFI: the reference implements this path in numpy
 * (/root/reference/src/pyliger/factorization/).  Every entry point below names the reference call
 * site(s) it replaces.  The host side that mirrors the pyliger Python API lives in pyliger_b200/ and
 * binds these symbols with ctypes (INTEGRATION.md shows the stub a pyliger maintainer would add).
 *
 * Conventions
 *   - extern "C", plain C types only.  Every function returns 0 on success, < 0 on error;
 *     inmf_last_error() returns a thread-local message.
 *   - The caller owns all memory.  Pointers are DEVICE pointers unless stated otherwise.  No hidden
 *     allocations, no host synchronisation: all work is enqueued on `stream` (a cudaStream_t).
 *   - `_f32` / `_f64` suffix = element type T of X values, H, W, V, B and right-hand sides.
 *     Small k x k Gram matrices, A, and scalar accumulators are always double.
 *   - Layout: every dense operand is "entity-major" with the k factors contiguous:
 *       Wt, Vt[s], St[s], B[s], Mt[s] : genes x k (Mt padded to ldm >= k columns)
 *       H, R                          : cells x k
 *     i.e. the transposes of optimize_ALS's k x g matrices and exactly online_iNMF's g x k ones.
 *   - X is CSR over ALL datasets concatenated by rows: rowptr int64[n_total+1], colidx int32[nnz],
 *     vals T[nnz]; a "segment" is one dataset (or one mini-batch window of it).
 *   - seg_desc: device int64[nseg][4] = { out_off, row_base, win_start, win_mod }.  Output row
 *     q in [out_off[s], out_off[s+1]) of segment s reads CSR row
 *         row_base + (win_start + (q - out_off)) % win_mod .
 *     (ALS: win_start = 0, win_mod = n_s.  Online mini-batch t: win_start = t*mb_s mod n_s, the
 *     wrap-around of _online_iNMF.py:511-552.)  seg_desc has nseg+1 rows; only out_off of the last
 *     row is read.
 *   - k <= 64.
 */
#ifndef INMF_B200_H
#define INMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INMF_MAX_K 64
#define INMF_BPP_NSTATS 8 /* not_converged, factorizations, backup_flips, pivot_failures, rounds_sum, rounds_max, columns, deferred (half-warp -> full-warp solver) */

int32_t inmf_version(void);
const char* inmf_last_error(void);
/* Number of kernels this library has launched since it was loaded (host-side counter). */
int64_t inmf_launch_count(void);
/* A CUDA graph captured from calls into this library was replayed: add the n kernels it contains to the counter. */
void inmf_note_graph_launches(int64_t n);

/* Batched block-principal-pivoting NNLS; one problem per column, G shared per segment.
 * Replaces nnlsm_blockpivot + normal_eq_comb (_utilities.py:181-358) for all six call sites
 * (_iNMF_ANLS.py:136,143,149; _online_iNMF.py:315,420,498).
 *   G        double[nseg][k][k]   normal matrices (AtA)
 *   R        T, column c at R + c*ldr (k contiguous values of AtB[:, c])
 *   X        T out, column c at X + c*ldx (X may alias R: in-place solve)
 *   Y        T out or NULL (G x - r, zero on the passive set)
 *   mask     uint64 per column or NULL: final passive set; if (warm & 1) it is also the start set
 *            (the reference's `init` argument, _utilities.py:222-225)
 *   warm     bit 0: warm start from mask; bit 1 (needs ginv_ws): solve a partition on its active-set
 *            complement through G^-1 whenever that is the smaller system (same solution; meant for
 *            well-conditioned G such as the H-update's, whose solutions are mostly passive)
 *            bit 2: Y is not an output but a buffer to CLEAR: the k entries of every solved column are set to 0
 *            (the ALS loop ping-pongs two right-hand-side buffers; the one pass 1 accumulates into next is
 *            cleared here instead of by a separate memset on the critical path)
 *   seg_off  device int64[nseg+1] column ranges;  max_seg_cols = max segment length (host value)
 *   stats    device int64[INMF_BPP_NSTATS], accumulated (caller zeroes)
 *   ws, ws_bytes  device workspace or NULL, sized with inmf_bpp_workspace_bytes():
 *            [nseg][k][k] doubles: G^-1 is formed per segment first, the "all indices passive" solves (the first
 *            round of every cold start on non-negative right-hand sides) become one k x k mat-vec, and passive
 *            sets larger than the register solver's limit are solved on the complement (same NNLS solution);
 *            the rest (int32 counters and per-segment column lists) lets the HALF-WARP solver run first: two
 *            columns per warp, 16 lanes each, partitions of <= 15 unknowns; a column with a larger partition is
 *            put on its segment's list and solved by the full-warp kernel afterwards, from its original start
 *            set.  A workspace of just nseg*k*k doubles is accepted (inverse only, full-warp solver).
 *            warm bit 3 switches the half-warp solver off; warm bit 4 skips forming G^-1 although the workspace
 *            would hold it (small warm-started solves where the k^3 inverse costs more than it saves).
 */
int64_t inmf_bpp_workspace_bytes(int64_t max_seg_cols, int32_t nseg, int32_t k);
int32_t inmf_bpp_f32(const double* G, const float* R, int64_t ldr, float* X, int64_t ldx, float* Y,
                     int64_t ldy, uint64_t* mask, int32_t warm, const int64_t* seg_off, int32_t nseg,
                     int64_t max_seg_cols, int32_t k, int64_t* stats, void* ws, int64_t ws_bytes, void* stream);
int32_t inmf_bpp_f64(const double* G, const double* R, int64_t ldr, double* X, int64_t ldx, double* Y,
                     int64_t ldy, uint64_t* mask, int32_t warm, const int64_t* seg_off, int32_t nseg,
                     int64_t max_seg_cols, int32_t k, int64_t* stats, void* ws, int64_t ws_bytes, void* stream);

/* H-update preparation per segment s (replaces the hstack/AtA of _iNMF_ANLS.py:137 +
 * _utilities.py:212 and _online_iNMF.py:421): Mt[s] = Wt + Vt[s] (zero padded to ldm),
 * Gwv[s] = Mt'Mt, Gvv[s] = Vt'Vt, G[s] = Gwv[s] + lambda*Gvv[s].  Vt may be NULL (projection, V = 0).
 * Outputs are overwritten. */
int32_t inmf_prep_f32(const float* Wt, const float* Vt, float* Mt, int64_t ldm, double* G, double* Gwv,
                      double* Gvv, double lambda, int32_t nseg, int64_t g, int32_t k, void* stream);
int32_t inmf_prep_f64(const double* Wt, const double* Vt, double* Mt, int64_t ldm, double* G, double* Gwv,
                      double* Gvv, double lambda, int32_t nseg, int64_t g, int32_t k, void* stream);

/* R[q, :] = X[row(q), :] * Mt[s]  -- the AtB of the H-update, CSR streamed once
 * (_iNMF_ANLS.py:138 + _utilities.py:217; _online_iNMF.py:422-424, 499-503). */
int32_t inmf_spmm_f32(const int64_t* rowptr, const int32_t* colidx, const float* vals, const float* Mt,
                      int64_t ldm, int64_t g, float* R, int64_t ldr, const int64_t* seg_desc,
                      int32_t nseg, int64_t max_seg_rows, int32_t k, void* stream);
int32_t inmf_spmm_f64(const int64_t* rowptr, const int32_t* colidx, const double* vals, const double* Mt,
                      int64_t ldm, int64_t g, double* R, int64_t ldr, const int64_t* seg_desc,
                      int32_t nseg, int64_t max_seg_rows, int32_t k, void* stream);

/* Blocked form of the two X-streaming passes (the optimised hot kernels; same maths as inmf_spmm /
 * inmf_stats for whole datasets).  X is stored a second and third time as packed (index, value) pairs
 * -- f32: {int32 idx, float val} (8 B), f64: {int64 idx, double val} (16 B) -- grouped so that the
 * gathered dense operand of a CTA is one shared-memory tile:
 *   pass 1  pairs sorted by (dataset, gene block, cell, gene), idx = gene - block start,
 *           dense = Mt (rows = genes), out = R/H rows (ACCUMULATED: caller zeroes)
 *   pass 2  pairs sorted by (dataset, cell block, gene, cell), idx = cell - block start,
 *           dense = H (rows = cells, leading dimension kp), out = St rows (accumulated), and
 *           gram_out[seg] (k x k double) += tile'tile for work items with gram_seg >= 0
 * work: device int64[nwork][8] = {ptr_base, dense_row0, tile_rows, row_begin, row_end, out_row_base,
 * gram_seg, 0}: the CTA loads dense rows [dense_row0, +tile_rows) (leading dimension kp, TMA bulk copy)
 * and, for r in [row_begin, row_end), adds sum_{q in [ptr[ptr_base+r], ptr[ptr_base+r+1])}
 * val_q * tile[idx_q, :] to out[(out_row_base + r) * ldo + 0..k). */
int32_t inmf_blocked_spmm_f32(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                              const float* dense, int32_t kp, float* out, int64_t ldo, int32_t k,
                              int32_t max_tile_rows, double* gram_out, void* stream);
int32_t inmf_blocked_spmm_f64(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                              const double* dense, int32_t kp, double* out, int64_t ldo, int32_t k,
                              int32_t max_tile_rows, double* gram_out, void* stream);

/* Same contract as inmf_blocked_spmm_f32 for 32 < k <= 64, with the dense tile held split in shared memory
 * ([rows x 32] + [rows x remainder]) so that every 8-lane phase of the 128-bit operand loads is one contiguous
 * 128-byte row.  Differences: the pairs' first field is (row of the tile) * 128 (not * kp * 4), and the tile may hold
 * at most (227 KB - 128) / ((32 + rs) * 4) rows, rs = 8, 16 or 32 for kp - 32 <= 8, <= 16, <= 32. */
int32_t inmf_blocked_split_f32(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                               const float* dense, int32_t kp, float* out, int64_t ldo, int32_t k,
                               int32_t max_tile_rows, double* gram_out, void* stream);

/* Same contract and the same formats as inmf_blocked_spmm_f32 for kp == 32: 4 lanes per non-zero with 8 factors each
 * (two bank-swizzled 128-bit loads), 8 non-zeros per warp instruction. */
int32_t inmf_blocked_wide_f32(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                              const float* dense, int32_t kp, float* out, int64_t ldo, int32_t k,
                              int32_t max_tile_rows, double* gram_out, void* stream);

/* HALS update of H, one Gauss-Seidel sweep over the k factors of every cell, in place (iNMF_HALS's _update_H_HALS,
 * _utilities.py:132-148; the W / V sweeps of _iNMF_HALS.py:118-121 are inmf_hals_* with A = H'H, B = X'H):
 *   h_j <- max(1e-16, h_j + (R[c, j] - G[s][j, :] h) / G[s][j, j])     for j = 0 .. k-1,
 * with R = X (W + V_s)' from the pass-1 kernels and G[s] from inmf_prep_*.  Columns as in inmf_bpp_*. */
int32_t inmf_hals_h_f32(const double* G, const float* R, int64_t ldr, float* H, int64_t ldh, const int64_t* seg_off,
                        int32_t nseg, int64_t max_seg_cols, int32_t k, void* stream);
int32_t inmf_hals_h_f64(const double* G, const double* R, int64_t ldr, double* H, int64_t ldh, const int64_t* seg_off,
                        int32_t nseg, int64_t max_seg_cols, int32_t k, void* stream);

/* Sliced (SELL-8) form of the tiled product above, fp32, k <= 64: the default of the two X-streaming passes of an ALS
 * iteration (R = X (W+V_s)' of _iNMF_ANLS.py:135-139; X' H, H' H of :142-152).  The sparse rows of every work item
 * (the rows of the row-ordered blocked format: (cell, gene block) segments in pass 1, (gene, cell block) segments in
 * pass 2) are sorted by length and stored in slices of `height` rows, step-interleaved, two steps per 16 bytes:
 *   pairs4[slice_ptr[s] + jj*height + i] = { off(2jj), val(2jj), off(2jj+1), val(2jj+1) } of row i of slice s,
 *   off = (row of the dense tile) * 128, padding = {0, 0.0f};  slice_ptr int64[nslices+1] in 16-byte units;
 *   rowid int32[nslices*height]: output row of slice row i relative to the work item's out_row_base, -1 = no row.
 *   height = 8: four lanes per row (any kp);  height = 32 (kp == 32 only): one lane per row, the pair stream is not
 *   replicated across lanes and every lane walks the 16-byte chunks of its tile row in a rotated, bank-conflict-free
 *   order;  height = 33 (32 < kp <= 40): 32 rows per slice, one lane per row, [rows x 32] + [rows x 8] tile, the
 *   non-zeros of every row ordered by (tile row mod 4) so that the 32-byte remainder rows of a phase fall in different
 *   bank groups (inmf_sell_fill_f32 with height = 33 builds that order).
 *   work int64[nwork][8] = { slice_begin, slice_end, dense_row0, tile_rows, out_row_base, gram_seg, part, 0 }
 * dense has kp floats per row (kp = 32 for k <= 32, else k rounded up to 4); the tile is held in shared memory as
 * [rows x 32] + [rows x rs], rs = 0 / 8 / 16 / 32 for kp - 32 = 0 / <= 8 / <= 16 / <= 32, so max_tile_rows is bounded by
 * (227 KB - 128) / ((32 + rs) * 4).  out rows are accumulated (caller zeroes). */
int32_t inmf_sell_pass_f32(const int64_t* work, int32_t nwork, const int64_t* slice_ptr, const int32_t* rowid,
                           const void* pairs4, int32_t height, const float* dense, int32_t kp, float* out, int64_t ldo,
                           int32_t k, int32_t max_tile_rows, double* gram_out, int64_t part_stride,
                           int64_t gram_part_stride, void* stream);

/* Pass 2 fused with its all-reduce (multi-GPU, cells sharded over the ranks; the all-reduce of X' H and H' H over the
 * ranks is what makes _iNMF_ANLS.py:142-152 see all cells).  Same arguments as inmf_sell_pass_f32 with ldo == k; partial
 * sums are accumulated in the rank's own `out` / `gram_out` as there.  In addition the work items that add into one
 * (dataset, gene range) share an arrival counter, work[w][6] = counter id and work[w][7] = (first row << 32) | end row of
 * the range (rows relative to out_row_base); the CTA that arrives last at a counter adds the finished range (and, for the
 * range whose items carry gram_seg >= 0, the dataset's k x k block of gram_out) into the MULTICAST copy with
 * multimem.red: the NVLink switch performs the add on every rank's sum buffer while the rest of the pass is still
 * computing, and no separate collective follows.
 *   mc_args -> device struct { float* out_mc; double* gram_mc; int32_t* counters; const int32_t* expected; }:
 *   multicast addresses of the symmetric sum buffers (e.g. torch.distributed._symmetric_memory multicast_ptr + offset;
 *   same layouts as out / gram_out), one zero-initialised counter per range (reset by the last arriver) and the number of
 *   work items per counter.
 * The caller orders the ranks: every rank's sum buffer is zero before any rank launches (one cross-rank barrier), and is
 * read after all ranks' kernels have completed (another). */
int32_t inmf_sell_pass_mc_f32(const int64_t* work, int32_t nwork, const int64_t* slice_ptr, const int32_t* rowid,
                              const void* pairs4, int32_t height, const float* dense, int32_t kp, float* out, int32_t k,
                              int32_t max_tile_rows, double* gram_out, const void* mc_args, void* stream);

/* Deterministic (bit-reproducible) accumulation.  By default the passes add their per-tile partial sums into the
 * outputs with floating-point atomics, whose order depends on the CTA schedule.  With part_stride > 0 in
 * inmf_sell_pass_f32, `out` (and gram_out, gram_part_stride doubles apart) is an array of partial buffers: work item w
 * STORES its rows into partial work[w][6] (one writer per element), and inmf_reduce_parts_* then forms
 * out[i] = part[0][i] + part[1][i] + ... in that fixed order.  inmf_set_deterministic(1) additionally makes
 * inmf_prep_* and inmf_objective_* use one CTA per reduction (their double-precision atomics otherwise make the last
 * bits of the Gram matrices / the objective depend on the schedule); returns the previous setting.  Process-wide. */
int32_t inmf_set_deterministic(int32_t on);
int32_t inmf_reduce_parts_f32(const float* part, int32_t nparts, int64_t stride, int64_t n, float* out, void* stream);
int32_t inmf_reduce_parts_f64(const double* part, int32_t nparts, int64_t stride, int64_t n, double* out, void* stream);

/* Builder of pairs4: copies the rows slice_rows[s*height + i] (indices into ptr of the row-ordered blocked format built
 * by inmf_blockfmt1/2_f32, -1 = no row) into the slices laid out by slice_ptr, zero padded.  height = -32: slices of
 * 32 rows with every row's non-zeros reordered by the parity of (off / 64) for inmf_sell_pass_mixed. */
int32_t inmf_sell_fill_f32(const int64_t* ptr, const void* pairs, const int64_t* slice_rows, const int64_t* slice_ptr,
                           int64_t nslices, int32_t height, void* pairs4, void* stream);

/* Mixed-precision form of inmf_sell_pass_f32 for k <= 32 (the engine's dtype = "mixed", never its fp32 / fp64 modes):
 * the dense operand is bf16, 32 values (64 bytes) per row, zero padded (inmf_to_bf16_rows); pairs carry
 * off = (row of the tile) * 64; X values and the accumulation stay fp32.  Slices of height 32 built by
 * inmf_sell_fill_f32 with height = -32 (every row's non-zeros ordered by the parity of their tile row, so that the
 * two rows sharing a 128-byte shared-memory line are not read in the same phase).  No Gram output. */
int32_t inmf_sell_pass_mixed(const int64_t* work, int32_t nwork, const int64_t* slice_ptr, const int32_t* rowid,
                             const void* pairs4, const void* dense_bf16, float* out, int64_t ldo, int32_t k,
                             int32_t max_tile_rows, void* stream);
/* The same conversion for H (rows [seg_off[s], seg_off[s+1]) = dataset s) fused with the statistics H'H of the
 * UNROUNDED values: gram[s] (k x k, double) += sum_r H[r,:]' H[r,:] (accumulating; caller zeroes). */
int32_t inmf_to_bf16_gram(const float* H, int64_t ld, const int64_t* seg_off, int32_t nseg, int64_t max_seg_rows,
                          int32_t k, void* dst, double* gram, void* stream);
/* dst[r][0..32) = bf16(src[r*ld + 0..k)) rounded to nearest even, zero beyond k. */
int32_t inmf_to_bf16_rows(const float* src, int64_t ld, int64_t rows, int32_t k, void* dst, void* stream);

/* Sufficient statistics of one pass over the cells (accumulating; caller zeroes):
 *   St[s] (g x k) += sum_q X[row(q), :]' H[q, :],   HtH[s] (k x k, double) += sum_q H[q,:]' H[q,:]
 * Replaces the dense X - H@W / vstack products of _iNMF_ANLS.py:144-152 and the H H', X H' of
 * _online_iNMF.py:592-594. */
int32_t inmf_stats_f32(const int64_t* rowptr, const int32_t* colidx, const float* vals, const float* H,
                       int64_t ldh, float* St, double* HtH, int64_t g, const int64_t* seg_desc,
                       int32_t nseg, int64_t max_seg_rows, int32_t k, void* stream);
int32_t inmf_stats_f64(const int64_t* rowptr, const int32_t* colidx, const double* vals, const double* H,
                       int64_t ldh, double* St, double* HtH, int64_t g, const int64_t* seg_desc,
                       int32_t nseg, int64_t max_seg_rows, int32_t k, void* stream);

/* out[s] (k x k double, accumulated) += sum over the rows [out_off[s], out_off[s+1]) of A of a'a  (H'H alone;
 * seg_desc as above, only the out_off column is used). */
int32_t inmf_gram_rows_f32(const float* A, int64_t lda, const int64_t* seg_desc, int32_t nseg, int64_t max_seg_rows,
                           double* out, int32_t k, void* stream);
int32_t inmf_gram_rows_f64(const double* A, int64_t lda, const int64_t* seg_desc, int32_t nseg, int64_t max_seg_rows,
                           double* out, int32_t k, void* stream);

/* ||X_s||_F^2 per segment (double[nseg], overwritten); constant term of the objective. */
int32_t inmf_sqnorm_f32(const int64_t* rowptr, const float* vals, const int64_t* seg_desc, int32_t nseg,
                        double* out, void* stream);
int32_t inmf_sqnorm_f64(const int64_t* rowptr, const double* vals, const int64_t* seg_desc, int32_t nseg,
                        double* out, void* stream);

/* V-update normal equations (_iNMF_ANLS.py:142-146): Gv[s] = (1+lambda) HtH[s],
 * Rv[s][c,:] = St[s][c,:] - Wt[c,:] HtH[s]. */
int32_t inmf_rhs_v_f32(const float* St, const float* Wt, const double* HtH, float* Rv, double* Gv,
                       double lambda, int32_t nseg, int64_t g, int32_t k, void* stream);
int32_t inmf_rhs_v_f64(const double* St, const double* Wt, const double* HtH, double* Rv, double* Gv,
                       double lambda, int32_t nseg, int64_t g, int32_t k, void* stream);
/* W-update normal equations (_iNMF_ANLS.py:149-152): Gw = sum_s HtH[s],
 * Rw[c,:] = sum_s ( St[s][c,:] - Vt[s][c,:] HtH[s] ). */
int32_t inmf_rhs_w_f32(const float* St, const float* Vt, const double* HtH, float* Rw, double* Gw,
                       int32_t nseg, int64_t g, int32_t k, void* stream);
int32_t inmf_rhs_w_f64(const double* St, const double* Vt, const double* HtH, double* Rw, double* Gw,
                       int32_t nseg, int64_t g, int32_t k, void* stream);

/* Objective from statistics (_iNMF_ANLS.py:157-162 without n x g temporaries):
 * obj = sum_s xnorm2[s] - 2<St[s],Mt[s]> + <HtH[s],Gwv[s]> + lambda <HtH[s],Gvv[s]>;  *obj overwritten. */
int32_t inmf_objective_f32(const float* St, const float* Mt, int64_t ldm, const double* HtH,
                           const double* Gwv, const double* Gvv, const double* xnorm2, double lambda,
                           int32_t nseg, int64_t g, int32_t k, double* obj, void* stream);
int32_t inmf_objective_f64(const double* St, const double* Mt, int64_t ldm, const double* HtH,
                           const double* Gwv, const double* Gvv, const double* xnorm2, double lambda,
                           int32_t nseg, int64_t g, int32_t k, double* obj, void* stream);

/* Online statistics epilogue, _update_A_B (_online_iNMF.py:563-601), for nseg datasets at once.
 *   HHt double[nseg][k][k], XHt T[nseg][g][k]: RAW mini-batch sums from inmf_stats
 *   inv_mb double[nseg] (device): 1 / miniBatch_sizes[s];  scale = scale_param;  rollover != 0 when
 *   epoch > 0 and epoch - epoch_prev == 1.  A, A_old are double; B, B_old are T. */
int32_t inmf_update_ab_f32(double* A, double* A_old, float* B, float* B_old, const double* HHt,
                           const float* XHt, const double* inv_mb, double scale, int32_t rollover,
                           int32_t nseg, int64_t g, int32_t k, void* stream);
int32_t inmf_update_ab_f64(double* A, double* A_old, double* B, double* B_old, const double* HHt,
                           const double* XHt, const double* inv_mb, double scale, int32_t rollover,
                           int32_t nseg, int64_t g, int32_t k, void* stream);

/* HALS sweeps of W then every V_s, row (gene) parallel, n_inner times
 * (_update_W_HALS / _update_V_HALS, _utilities.py:102-129; loop _online_iNMF.py:444-460).
 * The first n_prev entries of A/B/Vt are the previous datasets of scenario 2
 * (_online_iNMF.py:450-455): they enter the W sums but their V is not updated. */
int32_t inmf_hals_f32(const double* A, const float* B, float* Wt, float* Vt, double lambda, int32_t n_prev,
                      int32_t nall, int64_t g, int32_t k, int32_t n_inner, void* stream);
int32_t inmf_hals_f64(const double* A, const double* B, double* Wt, double* Vt, double lambda,
                      int32_t n_prev, int32_t nall, int64_t g, int32_t k, int32_t n_inner, void* stream);

/* Dense normal-equation products for the raw-input form of nnlsm_blockpivot (_utilities.py:212-217):
 * A is m x n row-major, B is m x cols row-major.  G (double n x n) = A'A;  Rt (T, cols x n) = (A'B)'. */
int32_t inmf_normal_eq_f32(const float* A, const float* B, int64_t m, int32_t n, int64_t cols, double* G,
                           float* Rt, void* stream);
int32_t inmf_normal_eq_f64(const double* A, const double* B, int64_t m, int32_t n, int64_t cols, double* G,
                           double* Rt, void* stream);

/* Dense-tile (tcgen05 / TMEM) form of the two X-streaming passes, fp32 with 3xTF32 (three_pass != 0,
 * fp32-grade accuracy) or single TF32 arithmetic.  Same maths as inmf_blocked_spmm_f32:
 *   out[out_row0 + r, 0..k) (+)= sum over the tile's chunks of X_tile(128 x 64) * D_chunk(64 x k)
 * inmf_tc_relayout_f32 rewrites the dense operand D (rows x ldd; Mt for pass 1, H for pass 2) into per-chunk
 * UMMA operand tiles (npad x 64, K-major, no swizzle) split into TF32 hi / lo parts;
 * segd = device int64[nseg][3] {first row in D, rows, first chunk index}.
 * inmf_tc_pass_f32: work = device int64[nwork][8] {ptr_index0, dense_chunk0, nchunks, out_row0, tile_rows,
 * atomic, 0, 0}; ptr = int64, 8 entries per (tile, chunk) (16-row bands) + 1; pairs = {uint32 byte offset of
 * (row, k) inside the 128 x 64 A tile, float value}.  Plain stores unless `atomic`. */
int32_t inmf_tc_relayout_f32(const float* D, int64_t ldd, const int64_t* segd, int32_t nseg, int64_t max_seg_chunks,
                             float* out_hi, float* out_lo, int32_t k, int32_t npad, void* stream);
int32_t inmf_tc_pass_f32(const int64_t* work, int32_t nwork, const int64_t* ptr, const void* pairs,
                         const float* dc_hi, const float* dc_lo, float* out, int64_t ldo, int32_t k, int32_t npad,
                         int32_t three_pass, int32_t max_ctas, void* stream);

/* Profiling hook: with a library built with -DTC_DEBUG_TIMERS, inmf_tc_pass_f32 dumps per-role cycle counters
 * (64 int64 per CTA) into this device buffer; NULL (default) disables it.  No effect in normal builds. */
void inmf_tc_debug_buffer(void* device_int64_buffer);

/* Diagnostic: C (128 x n, row-major) = A (128 x 64) * B (64 x n) through one tcgen05 (kind::tf32) tile with
 * the accumulator in TMEM; three_pass != 0 uses the 3xTF32 split (fp32-grade accuracy).  Exercises the
 * shared-memory descriptor / instruction descriptor / TMEM plumbing of the dense-tile passes. */
int32_t inmf_tc_selftest(const float* A, const float* B, float* C, int32_t n, int32_t three_pass, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Producers of scale_data: the O(nnz) passes just before the factorisation.
 * CSR of one dataset (cells x genes): rowptr int64[n_rows + 1], colidx int32[nnz], vals T[nnz].
 * ------------------------------------------------------------------------------------------------------- */

/* normalize -> _normalize_matrix (/root/reference/src/pyliger/preprocessing/_normalization.py:91-99):
 * norm_vals = vals / (row sum of |vals|) (rows with zero norm unchanged: sklearn normalize(norm="l1")),
 * col_sum[g] += norm_vals, col_sumsq[g] += norm_vals^2 (double[genes], caller zeroes them: they become
 * var["norm_sum"], var["norm_sum_sq"]), *n_missing += rows whose plain sum is 0 (_remove_missing_obs,
 * /root/reference/src/pyliger/_utilities.py:74-75).  norm_vals may alias vals. */
int32_t inmf_normalize_rows_f32(const int64_t* rowptr, const int32_t* colidx, const float* vals, int64_t n_rows,
                                float* norm_vals, double* col_sum, double* col_sumsq, int64_t* n_missing,
                                void* stream);
int32_t inmf_normalize_rows_f64(const int64_t* rowptr, const int32_t* colidx, const double* vals, int64_t n_rows,
                                double* norm_vals, double* col_sum, double* col_sumsq, int64_t* n_missing,
                                void* stream);

/* scale_not_center -> _scale_matrix (/root/reference/src/pyliger/preprocessing/_scale.py:92-102): keep the
 * variable genes (colmap int32[genes]: new column index, or -1 = dropped; increasing on the kept genes) and
 * multiply by the per-gene scaler (double[kept genes] = 1 / sqrt(norm_sum_sq / (n_rows - 1))).
 * Two steps around one exclusive scan done by the caller:
 *   inmf_select_count   row_counts[r] = kept non-zeros of row r
 *   inmf_select_scale   writes new_colidx / new_vals at new_rowptr[r].. in the original order. */
int32_t inmf_select_count(const int64_t* rowptr, const int32_t* colidx, int64_t n_rows, const int32_t* colmap,
                          int64_t* row_counts, void* stream);
int32_t inmf_select_scale_f32(const int64_t* rowptr, const int32_t* colidx, const float* vals, int64_t n_rows,
                              const int32_t* colmap, const double* scaler, const int64_t* new_rowptr,
                              int32_t* new_colidx, float* new_vals, void* stream);
int32_t inmf_select_scale_f64(const int64_t* rowptr, const int32_t* colidx, const double* vals, int64_t n_rows,
                              const int32_t* colmap, const double* scaler, const int64_t* new_rowptr,
                              int32_t* new_colidx, double* new_vals, void* stream);

/* Device builders of the two blocked copies of X that inmf_blocked_spmm_* streams (stable counting sorts of the CSR
 * non-zeros; deterministic).  Call with fill = 0 (histogram), scan the counts into ptr, call with fill = 1.
 * Format 1 (gene-blocked CSR, pass 1): seg1 = int64[nseg][4] {own_off, n_own, csr_row0, 0}; owned row i of the H
 * buffer is CSR row csr_row0 + (i - own_off); key(s, b, local_row) = NB * own_off + b * n_own + local_row; gene block
 * b = col / gbs; pairs = {(col - b * gbs) * row_bytes, value}.  fill = 0: cnt_or_ptr[key] += count (caller zeroes);
 * fill = 1: cnt_or_ptr = ptr (exclusive scan of the counts, keys + 1 entries). */
int32_t inmf_blockfmt1_f32(int32_t fill, const int64_t* rowptr, const int32_t* colidx, const float* vals,
                           const int64_t* seg1, int32_t nseg, int64_t n_rows, int32_t NB, int32_t gbs,
                           int32_t row_bytes, int64_t* cnt_or_ptr, void* pairs, void* stream);
int32_t inmf_blockfmt1_f64(int32_t fill, const int64_t* rowptr, const int32_t* colidx, const double* vals,
                           const int64_t* seg1, int32_t nseg, int64_t n_rows, int32_t NB, int32_t gbs,
                           int32_t row_bytes, int64_t* cnt_or_ptr, void* pairs, void* stream);
/* Format 2 (cell-blocked CSC, pass 2): blk2 = int64[nblocks][4] {csr_row0, rows, key_base, 0}; key = key_base + gene;
 * pairs = {(row in block) * row_bytes, value}, cells increasing inside every gene list.  One CTA (W warps) per cell
 * block with W * g uint16 counters in shared memory (W * g * 2 <= 200 KB); offs = uint16[nblocks][W][g] scratch that
 * carries the per-warp offsets from the fill = 0 call (which writes cnt[key]) to the fill = 1 call (which reads ptr). */
int32_t inmf_blockfmt2_f32(int32_t fill, const int64_t* rowptr, const int32_t* colidx, const float* vals,
                           const int64_t* blk2, int32_t nblocks, int32_t g, int32_t row_bytes, int32_t W, int64_t* cnt,
                           const int64_t* ptr, uint16_t* offs, void* pairs, void* stream);
int32_t inmf_blockfmt2_f64(int32_t fill, const int64_t* rowptr, const int32_t* colidx, const double* vals,
                           const int64_t* blk2, int32_t nblocks, int32_t g, int32_t row_bytes, int32_t W, int64_t* cnt,
                           const int64_t* ptr, uint16_t* offs, void* pairs, void* stream);

/* Seeded initial factors without a host round trip: out[i] = i-th value of
 *   np.random.seed(seed); np.random.uniform(0, scale, n)
 * (numpy's legacy MT19937 stream, bit-identical), the draw the reference initialises W, V_i, H_i with
 * (/root/reference/src/pyliger/factorization/_iNMF_ANLS.py:103-108). */
int32_t inmf_mt19937_uniform(uint32_t seed, int64_t n, double scale, double* out, void* stream);

/* ---- quantile_norm (tools/_quantile_norm.py:9-195, clustering/_utilities.py:11-95): the consumer of H ------------ */

/* Max-factor assignment (_quantile_norm.py:110-125): the m selected columns dims[] of H (n x ld, double) are scaled by
 * sqrt(sum(x^2) / (n-1)) (center = 0) or centred and divided by std(ddof = 1) (center = 1); cluster[c] = argmax over
 * the selected columns (numpy semantics).  stats_ws: 2*m doubles of scratch. */
int32_t inmf_qn_max_factor(const double* H, int64_t ld, int64_t n, const int32_t* dims, int32_t m, int32_t center,
                           double* stats_ws, int32_t* cluster, void* stream);

/* Exact K nearest neighbours (Euclidean, self included) of every cell in the space of the selected factors
 * (run_knn, clustering/_utilities.py:11-17): knn int32[n][K], nearest first, ties by the smaller index. */
int32_t inmf_qn_knn(const double* H, int64_t ld, int64_t n, const int32_t* dims, int32_t m, int32_t K, int32_t* knn,
                    void* stream);

/* cluster_vote (clustering/_utilities.py:57-79): for i = 0 .. n-1 in order, cluster[i] = the most frequent label among
 * the K neighbours of cell i (ties: the largest label), in place, so later cells see the labels earlier cells of the
 * same sweep received.  Labels must be < 64. */
int32_t inmf_qn_vote(const int32_t* knn, int32_t K, int64_t n, int32_t* cluster, void* stream);

/* Quantile alignment of one dataset's factor loadings against the reference dataset (_quantile_norm.py:133-183), one
 * (cluster, factor) group per work item: group int64[ngroups][6] = { cluster j, column i, samp2_off, samp2_n, samp1_off,
 * samp1_n }.  memb / off: the cells of a dataset sorted by cluster (off[j] .. off[j+1]).  samp*_n = 0: all members
 * enter the quantiles; otherwise samp[off .. off+n) are positions in the cluster's member list (the first max_sample
 * entries of the reference's np.random.permutation, drawn by the host).  Hk[members, i] is overwritten with the
 * warped values (Hk may be Href: the reference dataset itself is aligned against itself). */
int32_t inmf_qn_align(const int64_t* group, int32_t ngroups, double* Hk, int64_t ldk, const double* Href, int64_t ldr,
                      const int32_t* memb2, const int32_t* off2, const int32_t* memb1, const int32_t* off1,
                      const int32_t* samp, int32_t quantiles, int32_t max_sample, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INMF_B200_H */
