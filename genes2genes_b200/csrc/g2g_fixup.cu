// K3  g2g_fixup_trends: one thread per gene.  Finishes the interpolation moments, applies the
// low-expression policy of RefQueryAligner.run_interpolation (Main.py:344-406) and derives the
// mean/std trends of the 50-sample bins through the fixed epsilon table.
#include "g2g_fixup_body.cuh"

namespace {

// Stage 1: sum the per-split partials of K1, one thread per (side, row, gene); coalesced over genes.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const __grid_constant__ ReduceParams p) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.n_genes) return;
  const int T = p.n_bins;
  const int rows_per_side = 3 * T + 1;
  const int side = blockIdx.y / rows_per_side;
  const int row = blockIdx.y % rows_per_side;
  if (row < 2 * T + 1) p.red_f64[side][(size_t)row * p.g_stride + g] = reduce_f64_row<8>(p, side, row, g);
  else p.red_i32[side][(size_t)(row - (2 * T + 1)) * p.g_stride + g] = reduce_i32_row<8>(p, side, row - (2 * T + 1), g);
}

constexpr int kFixWarps = 4;

// Stage 2: one warp per gene.
__global__ void __launch_bounds__(kFixWarps * 32) fixup_kernel(const FixParams p, const ReduceParams rp) {
  const long long g = (long long)blockIdx.x * kFixWarps + (threadIdx.x >> 5);
  if (g >= p.n_genes) return;
  const double* tf[2] = {rp.red_f64[0], rp.red_f64[1]};
  const int32_t* ti[2] = {rp.red_i32[0], rp.red_i32[1]};
  fixup_gene(p, tf, ti, rp.g_stride, g, g, threadIdx.x & 31);
}

}  // namespace

extern "C" size_t g2g_fixup_workspace_bytes(int64_t g_stride, int32_t n_bins) {
  return (size_t)2 * ((size_t)(2 * n_bins + 1) * g_stride * sizeof(double) + (size_t)n_bins * g_stride * sizeof(int32_t));
}

extern "C" int g2g_fixup_trends(const g2g_side_t* sides, const double* sum_w_ref, const double* sum_w_query,
                                const double* points, const double* eps_mean, const double* eps_std,
                                int64_t n_genes, int64_t g_stride, int32_t n_bins, double* musigma,
                                double* trends, int32_t* nnz, int32_t* flags, void* workspace,
                                size_t workspace_bytes, void* stream) {
  G2G_REQUIRE(sides && sum_w_ref && sum_w_query && points && eps_mean && eps_std, "null input pointer");
  G2G_REQUIRE(musigma && trends && nnz && flags, "null output pointer");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS, "n_bins out of range [2, 128]");
  G2G_REQUIRE(n_genes >= 0 && g_stride >= n_genes, "bad gene counts");
  G2G_REQUIRE(workspace && workspace_bytes >= g2g_fixup_workspace_bytes(g_stride, n_bins), "workspace too small");
  G2G_REQUIRE(((uintptr_t)workspace & 7) == 0, "workspace must be 8-byte aligned");
  if (n_genes == 0) return G2G_OK;
  FixParams p;
  ReduceParams rp;
  rp.n_genes = n_genes; rp.g_stride = g_stride; rp.n_bins = n_bins;
  {
    char* w = reinterpret_cast<char*>(workspace);
    const size_t f = (size_t)(2 * n_bins + 1) * g_stride * sizeof(double);
    const size_t i = (size_t)n_bins * g_stride * sizeof(int32_t);
    rp.red_f64[0] = reinterpret_cast<double*>(w);
    rp.red_f64[1] = reinterpret_cast<double*>(w + f);
    rp.red_i32[0] = reinterpret_cast<int32_t*>(w + 2 * f);
    rp.red_i32[1] = reinterpret_cast<int32_t*>(w + 2 * f + i);
  }
  for (int s = 0; s < 2; ++s) {
    G2G_REQUIRE(sides[s].part_f64 && sides[s].part_i32 && sides[s].pilot && sides[s].n_cells >= 1, "bad side");
    g2g_layout_t lay;
    int rc = g2g_layout(sides[s].n_cells, n_bins, &lay);
    if (rc != G2G_OK) return rc;
    rp.part_f64[s] = sides[s].part_f64;
    rp.part_i32[s] = sides[s].part_i32;
    rp.n_split[s] = sides[s].n_split > 0 ? sides[s].n_split : lay.n_split;
    p.side[s].pilot = sides[s].pilot;
    p.side[s].sum_w = s == 0 ? sum_w_ref : sum_w_query;
    p.side[s].n_cells = sides[s].n_cells;
  }
  p.points = points; p.eps_mean = eps_mean; p.eps_std = eps_std;
  p.musigma = musigma; p.trends = trends; p.nnz = nnz; p.flags = flags;
  p.n_genes = n_genes; p.n_bins = n_bins;
  dim3 rgrid((unsigned)g2g_ceil_div(n_genes, 256), (unsigned)(2 * (3 * n_bins + 1)));
  reduce_partials_kernel<<<rgrid, 256, 0, (cudaStream_t)stream>>>(rp);
  G2G_CUDA_TRY(cudaGetLastError());
  fixup_kernel<<<(unsigned)g2g_ceil_div(n_genes, kFixWarps), kFixWarps * 32, 0, (cudaStream_t)stream>>>(p, rp);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}
