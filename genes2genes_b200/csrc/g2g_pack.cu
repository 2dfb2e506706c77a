// g2g_pack_rows / g2g_pack_csr: bring an expression matrix into the layout K1 streams
// (f32, row major, cells sorted by pseudotime, selected genes only, 16-byte aligned rows).
// Replaces, for the device path, `adata[:, gene_list]` (Main.py:226-227), the time sort
// `adata[np.argsort(adata.obs['time'])]` (TimeSeriesPreprocessor.py:20) and the densification
// `adata.X.todense()` (Main.py:232,238; TSP.py:28).
#include "g2g_common.cuh"

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

}  // namespace

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
