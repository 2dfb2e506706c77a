// Aggregate (cell-level) alignment support: per-DP-cell histogram of the landscape states over a gene
// set and the matrix of matched (query bin, reference bin) counts.  Replaces the per-gene Python loops
// of ClusterUtils.get_cluster_average_alignments (ClusterUtils.py:455-518, the state counting at
// :478-490) and ClusterUtils.get_pairwise_match_count_mat (:435-450).
#include "g2g_common.cuh"

namespace {

constexpr int kGeneChunk = 128;

// grid (cells / 256, gene chunks); thread = one DP cell; coalesced over cells for each gene
__global__ void __launch_bounds__(256) state_histogram_kernel(const uint8_t* __restrict__ l_states, long long n_genes,
                                                              int n_cells, const uint8_t* __restrict__ mask,
                                                              int* __restrict__ counts) {
  const int cell = blockIdx.x * 256 + threadIdx.x;
  if (cell >= n_cells) return;
  const long long g0 = (long long)blockIdx.y * kGeneChunk;
  const long long g1 = g0 + kGeneChunk < n_genes ? g0 + kGeneChunk : n_genes;
  int c[5] = {0, 0, 0, 0, 0};
  for (long long g = g0; g < g1; ++g) {
    if (mask && !mask[g]) continue;
    const int s = l_states[(size_t)g * n_cells + cell];
#pragma unroll
    for (int k = 0; k < 5; ++k) c[k] += (s == k);
  }
#pragma unroll
  for (int k = 0; k < 5; ++k)
    if (c[k]) atomicAdd(&counts[(size_t)k * n_cells + cell], c[k]);
}

// one thread per gene: walk the alignment string; every M/W/V step lands on DP cell (i, j) =
// (query bin + 1, reference bin + 1) of a matched pair (DEAnalyser.get_matched_time_points, Main.py:716-773)
__global__ void __launch_bounds__(128) match_count_kernel(const uint8_t* __restrict__ align_str,
                                                          const int* __restrict__ align_len, long long n_genes,
                                                          int n_bins, const uint8_t* __restrict__ mask,
                                                          int* __restrict__ counts) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genes || (mask && !mask[g])) return;
  const uint8_t* s = align_str + (size_t)g * 2 * n_bins;
  const int len = align_len[g];
  int i = 0, j = 0;
  for (int k = 0; k < len; ++k) {
    const uint8_t ch = s[k];
    const bool di = (ch == 'M') || (ch == 'V') || (ch == 'I');
    const bool dj = (ch == 'M') || (ch == 'W') || (ch == 'D');
    i += di;
    j += dj;
    if (ch == 'M' || ch == 'W' || ch == 'V') atomicAdd(&counts[i * (n_bins + 1) + j], 1);
  }
}

}  // namespace

extern "C" int g2g_state_histogram(const uint8_t* l_states, int64_t n_genes, int32_t n_bins,
                                   const uint8_t* gene_mask, int32_t* counts, void* stream) {
  G2G_REQUIRE(l_states && counts, "null pointer");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS && n_genes >= 1, "bad shape");
  const int n_cells = (n_bins + 1) * (n_bins + 1);
  cudaStream_t st = (cudaStream_t)stream;
  G2G_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)5 * n_cells * sizeof(int32_t), st));
  dim3 grid((unsigned)g2g_ceil_div(n_cells, 256), (unsigned)g2g_ceil_div(n_genes, kGeneChunk));
  state_histogram_kernel<<<grid, 256, 0, st>>>(l_states, n_genes, n_cells, gene_mask, counts);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}

extern "C" int g2g_match_counts(const uint8_t* align_str, const int32_t* align_len, int64_t n_genes,
                                int32_t n_bins, const uint8_t* gene_mask, int32_t* counts, void* stream) {
  G2G_REQUIRE(align_str && align_len && counts, "null pointer");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS && n_genes >= 1, "bad shape");
  const int n_cells = (n_bins + 1) * (n_bins + 1);
  cudaStream_t st = (cudaStream_t)stream;
  G2G_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)n_cells * sizeof(int32_t), st));
  match_count_kernel<<<(unsigned)g2g_ceil_div(n_genes, 128), 128, 0, st>>>(align_str, align_len, n_genes, n_bins,
                                                                          gene_mask, counts);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}
