// g2g_pack_rows / g2g_pack_csr: bring an expression matrix into the layout K1 streams
// (f32, row major, cells sorted by pseudotime, selected genes only, 16-byte aligned rows).
// Replaces, for the device path, `adata[:, gene_list]` (Main.py:226-227), the time sort
// `adata[np.argsort(adata.obs['time'])]` (TimeSeriesPreprocessor.py:20) and the densification
// `adata.X.todense()` (Main.py:232,238; TSP.py:28).
#include "g2g_common.cuh"
#include <cstring>

namespace {

__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ src, long long n_cells,
                                                        long long n_genes, long long ld_src,
                                                        const long long* __restrict__ row_order,
                                                        const long long* __restrict__ dst_row,
                                                        const int* __restrict__ gene_idx, float* __restrict__ dst,
                                                        long long ld_dst) {
  const long long r = blockIdx.y;
  const long long sr = row_order ? row_order[r] : r;
  const float* in = src + (size_t)sr * ld_src;
  float* out = dst + (size_t)(dst_row ? dst_row[r] : r) * ld_dst;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ld_dst;
       g += (long long)gridDim.x * blockDim.x) {
    float v = 0.0f;
    if (g < n_genes) v = in[gene_idx ? gene_idx[g] : g];
    out[g] = v;
  }
}

// one warp per output row: dst row r = CSR row row_order[r], columns mapped through col_map (-1 = drop)
__global__ void __launch_bounds__(256) pack_csr_kernel(const long long* __restrict__ indptr,
                                                       const int* __restrict__ indices,
                                                       const float* __restrict__ data, long long n_cells,
                                                       const long long* __restrict__ row_order,
                                                       const int* __restrict__ col_map, float* __restrict__ dst,
                                                       long long ld_dst) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_cells) return;
  const int lane = threadIdx.x & 31;
  const long long sr = row_order ? row_order[r] : r;
  float* out = dst + (size_t)r * ld_dst;
  const long long b = indptr[sr], e = indptr[sr + 1];
  for (long long k = b + lane; k < e; k += 32) {
    int c = indices[k];
    if (col_map) c = col_map[c];
    if (c >= 0) out[c] = data[k];
  }
}

// ---- CSR (cells x all genes) -> CSC of the selected genes with time-sorted, ascending row indices ----
// A stable counting sort by column, deterministic without a general sort: the time-sorted rows are cut
// into chunks of kChunkRows; (1) every chunk counts its entries per selected column, (2) a per-column
// exclusive scan over the chunks turns the counts into each chunk's first position inside the column,
// (3) a scan over the columns gives col_ptr, (4) every chunk walks its rows IN ORDER (one block per chunk,
// a block barrier between rows) and appends each entry at its column's running position -- columns are
// unique inside a CSR row, so the running positions need no atomics.
constexpr int kChunkRows = 1024;
constexpr int kCscThreads = 256;

struct CscParams {
  const long long* indptr;
  const int* indices;
  const float* data;
  const long long* row_order;   // CSR row of the k-th cell in time order (null: identity)
  const int* col_map;           // source column -> selected gene (-1: dropped; null: identity)
  long long n_cells;
  int n_genes, n_chunks;
  int* chunk_pos;               // [n_chunks][n_genes]: counts, then first positions
  long long* col_ptr;           // [n_genes + 1]
  int* row_idx;
  float* values;
};

__global__ void __launch_bounds__(kCscThreads) csc_count_kernel(const CscParams p) {
  const int chunk = blockIdx.x;
  int* cnt = p.chunk_pos + (size_t)chunk * p.n_genes;
  const long long r0 = (long long)chunk * kChunkRows;
  const long long r1 = r0 + kChunkRows < p.n_cells ? r0 + kChunkRows : p.n_cells;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long r = r0 + warp; r < r1; r += kCscThreads / 32) {
    const long long sr = p.row_order ? p.row_order[r] : r;
    const long long b = p.indptr[sr], e = p.indptr[sr + 1];
    for (long long k = b + lane; k < e; k += 32) {
      int c = p.indices[k];
      if (p.col_map) c = p.col_map[c];
      if (c >= 0) atomicAdd(&cnt[c], 1);
    }
  }
}

// per column: exclusive scan of the chunk counts (in place) and the column total
__global__ void __launch_bounds__(256) csc_chunk_scan_kernel(const CscParams p, long long* col_count) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.n_genes) return;
  int run = 0;
  for (int k = 0; k < p.n_chunks; ++k) {
    int* slot = p.chunk_pos + (size_t)k * p.n_genes + c;
    const int v = *slot;
    *slot = run;
    run += v;
  }
  col_count[c] = run;
}

// col_ptr = exclusive scan of the column totals (one block; counts sit in col_ptr[1..] on entry)
__global__ void __launch_bounds__(1024) csc_col_scan_kernel(long long* col_ptr, int n_genes) {
  __shared__ long long s_part[1024];
  const int tid = threadIdx.x;
  const int per = (n_genes + 1023) / 1024;
  const int lo = tid * per, hi = min(n_genes, lo + per);
  long long sum = 0;
  for (int c = lo; c < hi; ++c) sum += col_ptr[1 + c];
  s_part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    long long run = 0;
    for (int k = 0; k < 1024; ++k) { const long long v = s_part[k]; s_part[k] = run; run += v; }
    col_ptr[0] = 0;
  }
  __syncthreads();
  long long run = s_part[tid];
  for (int c = lo; c < hi; ++c) { run += col_ptr[1 + c]; col_ptr[1 + c] = run; }   // inclusive sums = col_ptr[c + 1]
}

__global__ void __launch_bounds__(kCscThreads) csc_fill_kernel(const CscParams p) {
  const int chunk = blockIdx.x;
  int* pos = p.chunk_pos + (size_t)chunk * p.n_genes;
  const long long r0 = (long long)chunk * kChunkRows;
  const long long r1 = r0 + kChunkRows < p.n_cells ? r0 + kChunkRows : p.n_cells;
  for (long long r = r0; r < r1; ++r) {     // rows strictly in time order
    const long long sr = p.row_order ? p.row_order[r] : r;
    const long long b = p.indptr[sr], e = p.indptr[sr + 1];
    for (long long k = b + threadIdx.x; k < e; k += kCscThreads) {
      int c = p.indices[k];
      if (p.col_map) c = p.col_map[c];
      if (c >= 0) {
        const long long at = p.col_ptr[c] + pos[c];
        G2G_DEVICE_ASSERT(c < p.n_genes && at >= p.col_ptr[c] && at < p.col_ptr[c + 1]);
        pos[c] += 1;                        // one entry per (row, column): no other thread touches pos[c] in this row
        p.row_idx[at] = (int)r;
        p.values[at] = p.data[k];
      }
    }
    __syncthreads();                        // the next row sees this row's positions
  }
}

}  // namespace

extern "C" size_t g2g_csr_to_csc_workspace_bytes(int64_t n_cells, int64_t n_genes) {
  const int64_t n_chunks = g2g_ceil_div(n_cells, kChunkRows);
  return (size_t)n_chunks * (size_t)n_genes * sizeof(int);
}

extern "C" int g2g_csr_to_csc(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_cells,
                              const int64_t* row_order, const int32_t* col_map, int64_t n_genes, int64_t* col_ptr,
                              int32_t* row_idx, float* values, void* workspace, size_t workspace_bytes, void* stream) {
  G2G_REQUIRE(indptr && indices && data && col_ptr && row_idx && values && workspace, "null pointer");
  G2G_REQUIRE(n_cells >= 1 && n_cells < (1ll << 31) && n_genes >= 1 && n_genes < (1ll << 31), "bad shape");
  G2G_REQUIRE(workspace_bytes >= g2g_csr_to_csc_workspace_bytes(n_cells, n_genes), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  CscParams p;
  p.indptr = reinterpret_cast<const long long*>(indptr); p.indices = indices; p.data = data;
  p.row_order = reinterpret_cast<const long long*>(row_order); p.col_map = col_map;
  p.n_cells = n_cells; p.n_genes = (int)n_genes; p.n_chunks = (int)g2g_ceil_div(n_cells, kChunkRows);
  p.chunk_pos = reinterpret_cast<int*>(workspace);
  p.col_ptr = reinterpret_cast<long long*>(col_ptr); p.row_idx = row_idx; p.values = values;
  G2G_CUDA_TRY(cudaMemsetAsync(workspace, 0, g2g_csr_to_csc_workspace_bytes(n_cells, n_genes), st));
  csc_count_kernel<<<p.n_chunks, kCscThreads, 0, st>>>(p);
  csc_chunk_scan_kernel<<<(unsigned)g2g_ceil_div(n_genes, 256), 256, 0, st>>>(p, p.col_ptr + 1);
  csc_col_scan_kernel<<<1, 1024, 0, st>>>(p.col_ptr, p.n_genes);
  csc_fill_kernel<<<p.n_chunks, kCscThreads, 0, st>>>(p);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}

extern "C" int g2g_pack_rows(const float* src, int64_t n_cells, int64_t n_genes, int64_t ld_src,
                             const int64_t* row_order, const int32_t* gene_idx, float* dst, int64_t ld_dst,
                             void* stream) {
  G2G_REQUIRE(src && dst, "null pointer");
  G2G_REQUIRE(n_cells >= 1 && n_genes >= 1 && ld_dst >= n_genes, "bad shape");
  G2G_REQUIRE(n_cells < 65536ll * 32768ll, "too many cells");
  unsigned bx = (unsigned)g2g_ceil_div(ld_dst, 256);
  if (bx > 64) bx = 64;
  // grid.y is limited to 65535: loop over row blocks
  for (int64_t r0 = 0; r0 < n_cells; r0 += 65535) {
    int64_t rows = n_cells - r0 < 65535 ? n_cells - r0 : 65535;
    dim3 grid(bx, (unsigned)rows);
    pack_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        src, rows, n_genes, ld_src, row_order ? reinterpret_cast<const long long*>(row_order) + r0 : nullptr,
        nullptr, gene_idx, dst + (size_t)r0 * ld_dst, ld_dst);
    if (!row_order) src += (size_t)rows * ld_src;
  }
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}

extern "C" int g2g_scatter_rows(const float* src, int64_t n_rows, int64_t n_genes, int64_t ld_src,
                                const int64_t* dst_row, const int32_t* gene_idx, float* dst, int64_t ld_dst,
                                void* stream) {
  G2G_REQUIRE(src && dst && dst_row, "null pointer");
  G2G_REQUIRE(n_rows >= 1 && n_genes >= 1 && ld_dst >= n_genes, "bad shape");
  unsigned bx = (unsigned)g2g_ceil_div(ld_dst, 256);
  if (bx > 64) bx = 64;
  for (int64_t r0 = 0; r0 < n_rows; r0 += 65535) {
    int64_t rows = n_rows - r0 < 65535 ? n_rows - r0 : 65535;
    dim3 grid(bx, (unsigned)rows);
    pack_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src + (size_t)r0 * ld_src, rows, n_genes, ld_src, nullptr,
                                                             reinterpret_cast<const long long*>(dst_row) + r0,
                                                             gene_idx, dst, ld_dst);
  }
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}

extern "C" int g2g_upload_block(const float* host_src, int64_t ld_src, int64_t row0, int64_t n_rows, int64_t col0,
                                int64_t n_cols, float* dst, int64_t ld_dst, void* stream) {
  G2G_REQUIRE(host_src && dst, "null pointer");
  G2G_REQUIRE(n_rows >= 1 && n_cols >= 1 && row0 >= 0 && col0 >= 0, "bad block");
  G2G_REQUIRE(ld_src >= col0 + n_cols && ld_dst >= n_cols, "block does not fit the leading dimensions");
  G2G_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)ld_dst * sizeof(float), host_src + (size_t)row0 * ld_src + col0,
                                 (size_t)ld_src * sizeof(float), (size_t)n_cols * sizeof(float), (size_t)n_rows,
                                 cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return G2G_OK;
}

// ---- result window on the gathering rank, written by its peers over NVLink (copy engines, no kernel) ----
extern "C" int g2g_peer_window_create(size_t bytes, void** ptr, void* handle64) {
  G2G_REQUIRE(ptr && handle64 && bytes > 0, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  void* p = nullptr;
  G2G_CUDA_TRY(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    G2G_CUDA_TRY(e);
  }
  memcpy(handle64, &h, sizeof(h));
  *ptr = p;
  return G2G_OK;
}

extern "C" int g2g_peer_window_open(const void* handle64, void** ptr) {
  G2G_REQUIRE(ptr && handle64, "bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  G2G_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return G2G_OK;
}

extern "C" int g2g_peer_window_close(void* ptr, int32_t owner) {
  if (!ptr) return G2G_OK;
  if (owner) G2G_CUDA_TRY(cudaFree(ptr));
  else G2G_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return G2G_OK;
}

extern "C" int g2g_peer_push(void* dst, const void* src, size_t bytes, void* stream) {
  G2G_REQUIRE(dst && src, "null pointer");
  if (bytes) G2G_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return G2G_OK;
}

extern "C" int g2g_pack_csr(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n_cells,
                            const int64_t* row_order, const int32_t* col_map, float* dst, int64_t ld_dst,
                            void* stream) {
  G2G_REQUIRE(indptr && indices && data && dst, "null pointer");
  G2G_REQUIRE(n_cells >= 1 && ld_dst >= 1, "bad shape");
  G2G_CUDA_TRY(cudaMemsetAsync(dst, 0, (size_t)n_cells * ld_dst * sizeof(float), (cudaStream_t)stream));
  pack_csr_kernel<<<(unsigned)g2g_ceil_div(n_cells, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(indptr), indices, data, n_cells,
      reinterpret_cast<const long long*>(row_order), col_map, dst, ld_dst);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}
