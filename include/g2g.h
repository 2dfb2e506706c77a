/*
 * libg2g -- C ABI of the B200-native Genes2Genes per-gene alignment hot path.
 *
 * The reference (Teichlab/Genes2Genes v0.2.0) is pure Python and has no FFI layer; its seam for
 * this path is the Python API of genes2genes/Main.py (RefQueryAligner.__init__ :196-221,
 * align_all_pairs :434-457).  This header is the boundary a maintainer would bind from that
 * class with ctypes.  Each entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain C types only; every data pointer is a DEVICE pointer owned by the caller;
 *     the library allocates no user-visible memory (workspaces are caller-provided);
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     without synchronising; 0 = success, negative = error (g2g_last_error() has the text);
 *   - thread-compatible: no global state except a thread-local error string;
 *   - sm_100a (B200) only; there is no CPU fallback.
 */
#ifndef G2G_H_
#define G2G_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2G_VERSION 100          /* 0.1.0 */
#define G2G_MAX_BINS 128         /* n_interpolation_points supported: 2..128 */
#define G2G_CELL_TILE 64         /* rows of X per pipeline stage; weight tables are padded to it */

enum {
  G2G_OK = 0,
  G2G_ERR_ARG = -1,              /* bad argument (sizes, alignment, null pointer) */
  G2G_ERR_CUDA = -2,             /* a CUDA runtime / driver call failed */
  G2G_ERR_UNSUPPORTED = -3       /* device is not sm_100 or shape outside compiled range */
};

/* per-gene status bits written by g2g_fixup_trends (flags[g]) */
enum {
  G2G_FLAG_BRANCH_MASK = 0x3,    /* 0 both ~zero, 1 ref ~zero, 2 query ~zero, 3 both expressed (Main.py:344-376) */
  G2G_FLAG_REGION_REF = 0x4,     /* a low-expression region override fired on the reference side */
  G2G_FLAG_REGION_QUERY = 0x8,   /* ... on the query side */
  G2G_FLAG_BAD_SIGMA = 0x10      /* a final std is <= 0 or not finite: the reference raises here
                                    (torch Normal validation, MyFunctions.py:64) */
};

int g2g_version(void);
const char* g2g_last_error(void);
/* 0 when the current device is compute capability 10.x, else G2G_ERR_UNSUPPORTED. */
int g2g_check_device(void);

/* How the interpolation kernels tile `n_bins`: bins are split into `n_groups` groups of
 * `bins_per_group` (last group zero-padded); one row of the weight table holds, per group,
 * `row_floats` floats = [w_0 .. w_{bins_per_group-1}, (0 when that count is odd), valid, window_id(int bits),
 * 0-pad].  window_id = k for p_k <= tau < p_{k+1}, n_bins-1 outside every window, -1 for padding rows, plus
 * 0x10000 when the aligned group of 8 rows the cell belongs to spans more than one window. */
typedef struct {
  int32_t n_groups;
  int32_t bins_per_group;
  int32_t row_floats;            /* multiple of 4 */
  int32_t genes_per_tile;        /* genes handled by one thread block (32 or 64) */
  int64_t n_pad;                 /* n_cells rounded up to G2G_CELL_TILE */
  int64_t cells_per_item;        /* cell range of one work item (multiple of G2G_CELL_TILE); depends
                                    on (n_cells, n_bins) only, never on n_genes => shard-invariant sums */
  int64_t n_split;               /* ceil(n_pad / cells_per_item) */
} g2g_layout_t;

int g2g_layout(int64_t n_cells, int32_t n_bins, g2g_layout_t* out);

/* ---------------------------------------------------------------------------------------------
 * g2g_pack_rows / g2g_pack_csr -- bring an expression matrix into the layout K1 streams.
 * Replaces `adata[:, gene_list]` (Main.py:226-227), the pseudotime sort (TimeSeriesPreprocessor.py:20)
 * and `adata.X.todense()` (Main.py:232,238; TSP.py:28) for device-resident data.
 *   dst[r][g] = src[row_order[r]][gene_idx[g]]   (row_order / gene_idx may be null = identity);
 *   columns n_genes..ld_dst-1 of dst are zero filled.  g2g_pack_csr scatters a CSR matrix
 *   (cells x all genes; indptr i64, indices i32, data f32) into a zeroed dst, mapping source column c
 *   to col_map[c] (-1 = gene not selected; null = identity).
 */
int g2g_pack_rows(const float* src, int64_t n_cells, int64_t n_genes, int64_t ld_src,
                  const int64_t* row_order, const int32_t* gene_idx, float* dst, int64_t ld_dst,
                  void* stream);
int g2g_pack_csr(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_cells,
                 const int64_t* row_order, const int32_t* col_map, float* dst, int64_t ld_dst,
                 void* stream);

/* g2g_csr_to_csc -- the sparse twin of the same step: a CSR matrix (cells x all genes, entries unique per
 * (row, column)) becomes the gene-major CSC matrix g2g_moments_csc streams: selected genes only
 * (col_map[c] = gene or -1; null = identity), cells in time order (row_order[k] = CSR row of the k-th cell;
 * null = identity), row indices ascending inside every column.  A stable counting sort by column (per
 * 1024-row chunk counts, scans, ordered fill): deterministic, no general sort.  col_ptr i64 [n_genes+1];
 * row_idx / values need room for every stored entry of the input; workspace as reported, 4-byte aligned. */
size_t g2g_csr_to_csc_workspace_bytes(int64_t n_cells, int64_t n_genes);
int g2g_csr_to_csc(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_cells,
                   const int64_t* row_order, const int32_t* col_map, int64_t n_genes, int64_t* col_ptr,
                   int32_t* row_idx, float* values, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * g2g_peer_window_* / g2g_peer_push -- the gather of SURVEY.md section 8(e) (reference: the results list
 * `align_all_pairs` fills, Main.py:452-455; here one record block per rank) without a kernel: the
 * gathering rank creates a device window (create: cudaMalloc + an inter-process handle of 64 bytes), the
 * other ranks of the box map it (open: peer access over NVLink / NVSwitch is enabled on first use), and
 * every rank copies its packed records into its slot with one asynchronous device-to-device copy on
 * `stream` (push).  The copy runs on a copy engine, so nothing competes with the persistent kernels of
 * the next batch.  Ordering between ranks (window free / all pushes landed) is the caller's: a stream
 * synchronise plus a barrier.  close: `owner` != 0 frees the window, 0 unmaps a peer's view. */
int g2g_peer_window_create(size_t bytes, void** ptr, void* handle64);
int g2g_peer_window_open(const void* handle64, void** ptr);
int g2g_peer_window_close(void* ptr, int32_t owner);
int g2g_peer_push(void* dst, const void* src, size_t bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * g2g_upload_block / g2g_scatter_rows -- the host->device leg of the same step, for a HOST matrix
 * (`adata.X`, row major f32, cells x all genes), pipelined by row chunk: only the column block
 * [col0, col0+n_cols) a rank needs crosses PCIe (one strided DMA per chunk, cudaMemcpy2DAsync; the
 * source should be pinned), and every uploaded chunk is scattered to its time-sorted position while the
 * next chunk is in flight.  Replaces `adata[:, gene_list]` + the pseudotime sort (Main.py:226-227,
 * TimeSeriesPreprocessor.py:20) without a whole-matrix device copy.
 *   g2g_upload_block: dst[r][c] = host_src[row0 + r][col0 + c], r < n_rows, c < n_cols (dst is a DEVICE
 *     pointer, host_src the only HOST data pointer of the ABI).
 *   g2g_scatter_rows: dst[dst_row[r]][g] = src[r][gene_idx[g]] (gene_idx null = identity), columns
 *     n_genes..ld_dst-1 zero filled; dst_row i64 [n_rows] on the device.
 */
int g2g_upload_block(const float* host_src, int64_t ld_src, int64_t row0, int64_t n_rows, int64_t col0,
                     int64_t n_cols, float* dst, int64_t ld_dst, void* stream);
int g2g_scatter_rows(const float* src, int64_t n_rows, int64_t n_genes, int64_t ld_src,
                     const int64_t* dst_row, const int32_t* gene_idx, float* dst, int64_t ld_dst,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * K0  g2g_weights -- Gaussian-kernel cell weights for one dataset.
 * Replaces TrajectoryInterpolator.compute_abs_timediff_mat / compute_Weight_matrix
 * (TimeSeriesPreprocessor.py:44-51, 122-133) and the window membership test of
 * RefQueryAligner.check_inconsistent_zero_region (Main.py:298-299).
 *   tau[n_cells]        f64 pseudotimes (any order; the sums are order independent)
 *   points[n_bins]      f64 interpolation points (np.linspace(0,1,n_bins), TSP.py:23)
 *   denom[n_bins]       f64 kernel denominators: kernel_WINDOW_SIZE**2 everywhere (TSP.py:127) or
 *                       the adaptive per-bin values (TSP.py:88-119, 125)
 *   wtab                f32 out, [n_groups][n_pad][row_floats]; w = exp(-(tau-p_t)^2/denom_t) computed
 *                       in f64 and rounded once; rows >= n_cells are zero with window -1
 *   sum_w[n_bins]       f64 out, sum over cells of the f64 weights (np.sum(row), TSP.py:164)
 *   workspace           >= g2g_weights_workspace_bytes() bytes, 8-byte aligned
 */
size_t g2g_weights_workspace_bytes(int64_t n_cells, int32_t n_bins);
int g2g_weights(const double* tau, int64_t n_cells, const double* points, const double* denom,
                int32_t n_bins, float* wtab, double* sum_w, void* workspace, size_t workspace_bytes,
                void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1a g2g_pilot_mean -- per-gene shift value a[g] (mean of a fixed strided sample of <= 64 rows).
 * Any value close to the gene's mean works: it only conditions the single-pass variance
 *   sum w (x - xbar)^2 = sum w (x-a)^2 - 2 (xbar-a) (sum w x - a sum w) + (xbar-a)^2 sum w
 * which g2g_fixup_trends evaluates in f64 (the reference takes the variance around the
 * un-weighted whole-gene mean, TSP.py:171-172).
 */
int g2g_pilot_mean(const float* X, int64_t n_cells, int64_t n_genes, int64_t ldx, float* pilot,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1  g2g_moments -- single pass over the expression matrices of up to two datasets.
 * Replaces the per-gene, per-bin reductions of compute_stat (TSP.py:161-174: np.sum(row),
 * np.dot(row, x), np.dot(row, (x - mean)**2)), Utils.csr_mat_col_densify (Utils.py:6-12),
 * np.count_nonzero (Main.py:344,350,363) and the per-window counts of Main.py:298-301.
 *   X            f32 [n_cells][ldx] row major (cells x genes); ldx % 4 == 0 and X 16-byte aligned
 *                (the matrix is streamed with TMA)
 *   wtab, pilot  from g2g_weights / g2g_pilot_mean
 *   part_f64     f64 out [n_split][2*n_bins+1][g_stride]: per split, sum w x per bin,
 *                sum w (x-a)^2 per bin, sum (x-a)
 *   part_i32     i32 out [n_split][n_bins][g_stride]: nnz per window k (p_k <= tau < p_{k+1}) for
 *                k < n_bins-1, and at index n_bins-1 the nnz of cells outside every window
 *   g_stride     >= n_genes
 * Products are f32 (exact inputs, weights rounded once), summed in f32 over at most 64 cells and
 * then in f64, in an order fixed by (n_cells, n_bins) alone.
 */
typedef struct {
  const float* X;
  int64_t n_cells;
  int64_t ldx;
  const float* wtab;             /* g2g_weights table (g2g_moments) or g2g_weights_mma table (g2g_moments_mma) */
  const float* pilot;
  double* part_f64;
  int32_t* part_i32;
  int64_t n_split;               /* partial-sum planes in part_*: 0 = the count of g2g_layout() (g2g_moments);
                                    g2g_moments_mma writes g2g_mma_layout().n_split planes: set it for K3 */
} g2g_side_t;

int g2g_moments(const g2g_side_t* sides, int32_t n_sides, int64_t n_genes, int64_t g_stride,
                int32_t n_bins, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1m g2g_moments_mma -- tensor-core variant of g2g_moments for 30 <= n_bins <= 128 (same reference code
 * replaced: TSP.py:161-174, Main.py:298-301,344-363).  out[genes x bins] = X^T [genes x cells] . W [cells x bins]
 * runs as tcgen05.mma kind::f16 (M = 128 genes = TMEM lanes, N = bins, K = 16 cells) with f32 accumulators
 * in tensor memory.  To keep the f32 tolerance of the path: x and (x - a)^2 are split on the CUDA cores
 * into f16 pairs (hi, lo * 2048) written to tensor memory as the A operand, the weights are split the same
 * way by g2g_weights_mma into a shared-memory B operand table, and three products per moment are kept
 * (hi.hi, lo.hi + hi.lo: 22 bits of each factor); sum (x - a) is added up on the CUDA cores.  The tensor core truncates when it accumulates
 * (measured -1e-7 of the sum per 16-cell step, profiles/r02_umma_ts_probe.txt), so the hi.hi chains are
 * cut every 128 cells: the accumulators are read back and added in f32 round-to-nearest registers, which
 * are added in f64 once per work item.  Values with |x| >= 65504 or |x - a| >= 255 do not fit the f16
 * split: the kernel then sets bit 0 of *status and the caller must redo the batch with g2g_moments.
 *   sides[s].wtab = the table of g2g_weights_mma; part_f64 / part_i32 as for g2g_moments but with
 *   g2g_mma_layout().n_split planes (set sides[s].n_split for K3); part_i32 is zeroed by the call
 *   (counts are added with integer atomics).
 *   status   i32 device word, OR-ed: bit 0 = non-finite accumulator (f16 range exceeded or non-finite input)
 */
typedef struct {
  int32_t n_groups;              /* bin groups (X is streamed once per group): ceil(n_bins / 64) */
  int32_t bins_per_group;
  int32_t n_cols;                /* N of the MMA: bins_per_group rounded up to 16 */
  int32_t reserved;
  int64_t n_pad;                 /* n_cells rounded up to G2G_CELL_TILE */
  int64_t cells_per_item;        /* depends on (n_cells, n_bins) only */
  int64_t n_split;
  int64_t table_bytes;           /* size of the g2g_weights_mma table for this dataset */
} g2g_mma_layout_t;

int g2g_mma_layout(int64_t n_cells, int32_t n_bins, g2g_mma_layout_t* out);
/* g2g_weights for the tensor-core path: same arguments and the same sum_w (bit for bit); `table` receives,
 * per bin group and per 64-cell stage, the f16 B operand of the stage's four K = 16 steps in the canonical
 * K-major no-swizzle core-matrix layout (hi then lo * 2048 per step) followed by 64 window words. */
int g2g_weights_mma(const double* tau, int64_t n_cells, const double* points, const double* denom,
                    int32_t n_bins, void* table, double* sum_w, void* workspace, size_t workspace_bytes,
                    void* stream);
int g2g_moments_mma(const g2g_side_t* sides, int32_t n_sides, int64_t n_genes, int64_t g_stride,
                    int32_t n_bins, int32_t* status, void* stream);
/* The same with `sm_reserve` SMs left without a CTA of the (persistent, one CTA per SM) kernel: a collective of
 * the previous batch running on another stream (the NCCL gather of the result records, whose kernels need a
 * few SMs) can then overlap this pass instead of waiting for it.  Results do not depend on the grid size. */
int g2g_moments_mma_ex(const g2g_side_t* sides, int32_t n_sides, int64_t n_genes, int64_t g_stride,
                       int32_t n_bins, int32_t* status, int32_t sm_reserve, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1s g2g_moments_csc -- sparse variant of g2g_moments: the expression matrix
 * stays in compressed-sparse-column form (gene major), work and bytes are proportional to the number
 * of stored entries.  Same outputs as g2g_moments with shift a = 0 (pilot[] is set to 0 by the call;
 * do not call g2g_pilot_mean): the totals go to split 0 of part_f64 / part_i32, the other splits are
 * zeroed, so g2g_fixup_trends is used unchanged.
 *   col_ptr  i64 [n_genes+1]; row_idx i32 [nnz] = cell index of each entry in the order of tau / wtab
 *   (ascending inside a column for deterministic sums); values f32 [nnz].
 */
typedef struct {
  const int64_t* col_ptr;
  const int32_t* row_idx;
  const float* values;
  int64_t n_cells;
  const float* wtab;
  float* pilot;
  double* part_f64;
  int32_t* part_i32;
} g2g_sparse_side_t;

int g2g_moments_csc(const g2g_sparse_side_t* sides, int32_t n_sides, int64_t n_genes, int64_t g_stride,
                    int32_t n_bins, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3  g2g_fixup_trends -- per gene: finish the moments, apply the low-expression policy, make trends.
 * Replaces the arithmetic of compute_stat (TSP.py:166-177), RefQueryAligner.run_interpolation
 * (Main.py:344-406), check_inconsistent_zero_region / extract_significant_regions_only
 * (Main.py:261-304) and SummaryTimeSeries mean_trend / std_trend (TSP.py:205-206, through the
 * fixed epsilon table: mean(mu + sigma*eps_t) , std(mu + sigma*eps_t)).
 *   side[0] = reference (S), side[1] = query (T): the part_* / pilot / n_cells of g2g_moments
 *   sum_w_{ref,query}[n_bins], points[n_bins], eps_mean[n_bins], eps_std[n_bins]   f64
 *   musigma   f64 out [n_genes][4][n_bins]: intpl_means_S, intpl_stds_S, intpl_means_T, intpl_stds_T
 *   trends    f64 out [n_genes][4][n_bins]: mean_trend_S, std_trend_S, mean_trend_T, std_trend_T
 *   nnz       i32 out [n_genes][2]
 *   flags     i32 out [n_genes]  (G2G_FLAG_*)
 *   workspace >= g2g_fixup_workspace_bytes() bytes, 8-byte aligned (the per-split partial sums of
 *             K1 are first added up in split order into it)
 */
size_t g2g_fixup_workspace_bytes(int64_t g_stride, int32_t n_bins);
int g2g_fixup_trends(const g2g_side_t* sides, const double* sum_w_ref, const double* sum_w_query,
                     const double* points, const double* eps_mean, const double* eps_std,
                     int64_t n_genes, int64_t g_stride, int32_t n_bins, double* musigma,
                     double* trends, int32_t* nnz, int32_t* flags, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4  g2g_cost_dp -- MML match cost + five-state DP + traceback + landscape, one warp per gene.
 * Replaces DP5.compute_cell + MyFunctions.run_dist_compute_v3 (OrgAlign.py:471-503,
 * MyFunctions.py:10-50; closed form in the trend vectors), DP5.init / run_optimal_alignment
 * (OrgAlign.py:272-468), DP5.backtrack (:539-607) and AlignmentLandscape.collate_fwd (:732-747).
 *   trends      f64 [n_genes][4][n_bins] (as written by g2g_fixup_trends)
 *   trans[25]   f64 HOST pointer: transition costs I[to][from], states M,W,V,D,I (FiveStateMachine,
 *               OrgAlign.py:20-97); copied into the launch
 *   start[3]    f64 HOST pointer: -ln(ProbI), -ln(ProbD) (OrgAlign.py:274-276,304,324), -ln(0.99) (:336,345)
 *   c_model     ln(5/(36 sqrt 3)) + ln 45 + 0.5 ln(2 N^2), n_samples = N = 50 (MyFunctions.py:21-29,41,46)
 *   cost_in     optional f64 [n_genes][n_bins][n_bins] (query bin i, ref bin j): use instead of computing
 *   align_str   u8 out [n_genes][2*n_bins], zero padded; align_len i32 out [n_genes]
 *   opt_cost    f64 out [n_genes]
 *   l_states    u8 out [n_genes][(n_bins+1)^2]  argmin over (M,W,V,D,I) per DP cell (:743-747)
 *   optional dense dumps (null to skip): cost_out f64 [n_genes][n_bins][n_bins];
 *   mats_out f64 [n_genes][5][(n_bins+1)^2]; bp_out i8 [n_genes][5][(n_bins+1)^2] (predecessor state,
 *   -1 where the reference defines none); l_out f64 [n_genes][(n_bins+1)^2]
 * All DP arithmetic is f64 adds / compares in the reference's order: bit-exact given equal costs.
 */
int g2g_cost_dp(const double* trends, int64_t n_genes, int32_t n_bins, const double* trans,
                const double* start, double c_model, int32_t n_samples, const double* cost_in,
                uint8_t* align_str, int32_t* align_len, double* opt_cost, uint8_t* l_states,
                double* cost_out, double* mats_out, int8_t* bp_out, double* l_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3+K4  g2g_fixup_cost_dp -- g2g_fixup_trends followed by g2g_cost_dp (compact outputs only) as ONE
 * kernel launch: the production path of RefQueryAligner.align_all_pairs (Main.py:434-457 calling
 * run_interpolation and align_single_pair per gene).  Same arguments and bit-identical outputs as the
 * two calls; no workspace (the split totals of a block's genes stay in shared memory).
 */
int g2g_fixup_cost_dp(const g2g_side_t* sides, const double* sum_w_ref, const double* sum_w_query,
                      const double* points, const double* eps_mean, const double* eps_std,
                      int64_t n_genes, int64_t g_stride, int32_t n_bins, const double* trans,
                      const double* start, double c_model, int32_t n_samples, double* musigma,
                      double* trends, int32_t* nnz, int32_t* flags, uint8_t* align_str,
                      int32_t* align_len, double* opt_cost, uint8_t* l_states, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Aggregate (cell-level) alignment over a gene set -- SURVEY.md 8(f) row f2.
 * g2g_state_histogram: counts[k][cell] = number of selected genes whose landscape state at DP cell
 *   `cell` is k (M,W,V,D,I); replaces the counting loop of ClusterUtils.get_cluster_average_alignments
 *   (ClusterUtils.py:471-490).  l_states u8 [n_genes][(n_bins+1)^2] (g2g_cost_dp output);
 *   gene_mask u8 [n_genes] or null (= all genes); counts i32 out [5][(n_bins+1)^2].
 * g2g_match_counts: counts[i][j] = number of selected genes that match query bin i-1 with reference
 *   bin j-1 (M/W/V steps); replaces ClusterUtils.get_pairwise_match_count_mat (ClusterUtils.py:435-450).
 *   counts i32 out [(n_bins+1)^2].
 */
int g2g_state_histogram(const uint8_t* l_states, int64_t n_genes, int32_t n_bins, const uint8_t* gene_mask,
                        int32_t* counts, void* stream);
int g2g_match_counts(const uint8_t* align_str, const int32_t* align_len, int64_t n_genes, int32_t n_bins,
                     const uint8_t* gene_mask, int32_t* counts, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Alignment-string distance matrices for clustering -- SURVEY.md 8(f) row f4.
 *   strs u8 [n_strings][stride] (g2g_cost_dp's align_str layout: stride = 2*n_bins), lens i32 [n_strings];
 *   dist u16 out [n_strings][n_strings] (integer distances; divide on the host like the reference);
 *   workspace >= g2g_strdist_workspace_bytes(n_strings) bytes, 8-byte aligned.
 * g2g_levenshtein_matrix replaces ClusterUtils.compute_levenshtein_dist_matrix (ClusterUtils.py:21-29;
 *   the reference then divides by max(len)); g2g_hamming_matrix replaces compute_hamming_dist_matrix
 *   (:31-38) on the binary path encoding of binary_encode_alignment_path (:361-377), 2*n_bins bits.
 */
size_t g2g_strdist_workspace_bytes(int64_t n_strings);
int g2g_levenshtein_matrix(const uint8_t* strs, const int32_t* lens, int64_t n_strings, int32_t stride,
                           uint16_t* dist, void* workspace, size_t workspace_bytes, void* stream);
int g2g_hamming_matrix(const uint8_t* strs, const int32_t* lens, int64_t n_strings, int32_t stride,
                       int32_t n_bins, uint16_t* dist, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * g2g_de_statistics -- rank statistics behind matched_region_DE_info, SURVEY.md 8(f) row f1.
 * Replaces the per-pair scipy.stats.wilcoxon / ks_2samp calls of
 * DEAnalyser.get_DE_info_for_matched_regions (Main.py:819-820) up to the p-value lookup: for pair k
 * (gene pair_gene[k], reference bin pair_s[k], query bin pair_t[k]) the two 50-sample bins are rebuilt as
 * mean + eps * std from musigma (g2g_fixup_trends output) and eps f64 [n_bins][50], and
 *   r_plus[k]    sum of the average ranks of |x - y| over positive differences (zeros dropped)
 *   n_nonzero[k] number of non-zero differences;  tie_term[k] = sum over tie groups of t^3 - t
 *   ks_h[k]      max over pooled samples of |#{x <= v} - #{y <= v}|   (KS statistic D = h / 50)
 */
int g2g_de_statistics(const double* musigma, const double* eps, int32_t n_bins, const int32_t* pair_gene,
                      const int32_t* pair_s, const int32_t* pair_t, int64_t n_pairs, double* r_plus,
                      double* tie_term, int32_t* n_nonzero, int32_t* ks_h, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* G2G_H_ */
