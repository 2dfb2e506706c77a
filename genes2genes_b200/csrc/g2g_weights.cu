// K0  g2g_weights     Gaussian-kernel weight table + per-bin weight sums for one dataset
// K1a g2g_pilot_mean  per-gene shift value from a fixed strided row sample
#include "g2g_common.cuh"
#include <math.h>

namespace {

constexpr int kWThreads = 256;

struct WeightsParams {
  const double* tau;
  const double* points;
  const double* denom;
  float* wtab;
  double* sum_w;
  double* partial;        // [n_blocks][T]
  unsigned int* counter;  // zeroed by the launcher
  long long n_cells, n_pad;
  int n_bins, n_groups, bins_per_group, row_floats;
};

// One thread per (padded) cell.  W[t,c] = exp(-(|tau_c - p_t|)^2 / denom_t) in float64
// (TimeSeriesPreprocessor.py:48, 125-127), rounded once to float32 for the table.
__global__ void __launch_bounds__(kWThreads) weights_kernel(const WeightsParams p) {
  __shared__ double s_pts[G2G_MAX_BINS];
  __shared__ double s_den[G2G_MAX_BINS];
  __shared__ double s_red[kWThreads / 32][G2G_MAX_BINS];
  __shared__ bool s_last;
  const int T = p.n_bins;
  for (int t = threadIdx.x; t < T; t += kWThreads) {
    s_pts[t] = p.points[t];
    s_den[t] = p.denom[t];
  }
  __syncthreads();
  const long long c = (long long)blockIdx.x * kWThreads + threadIdx.x;
  const bool in_pad = c < p.n_pad;
  const bool valid = c < p.n_cells;
  const double tau = valid ? p.tau[c] : 0.0;
  // window k: p_k <= tau < p_{k+1}  (half open, Main.py:299); cells outside every window -> T-1
  int win = T - 1;
  if (valid) {
    for (int k = 0; k + 1 < T; ++k)
      if (tau >= s_pts[k] && tau < s_pts[k + 1]) win = k;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int g = 0; g < p.n_groups; ++g) {
    float* row = p.wtab + ((size_t)g * p.n_pad + (size_t)c) * p.row_floats;
    for (int b = 0; b < p.row_floats; ++b) {
      const int t = g * p.bins_per_group + b;
      float out = 0.0f;
      double w = 0.0;
      if (b < p.bins_per_group) {
        if (valid && t < T) {
          double d = fabs(tau - s_pts[t]);
          w = exp(-((d * d) / s_den[t]));
          out = (float)w;
        }
        if (t < T) {
          // deterministic block reduction of the f64 weights: shuffle tree, then fixed warp order
          double r = w;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
          if (lane == 0) s_red[warp][t] = r;
        }
      } else if (b == p.bins_per_group) {
        out = valid ? 1.0f : 0.0f;
      } else if (b == p.bins_per_group + 1) {
        out = __int_as_float(valid ? win : -1);
      }
      if (in_pad) row[b] = out;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += kWThreads) {
    double r = 0.0;
    for (int w = 0; w < kWThreads / 32; ++w) r += s_red[w][t];
    p.partial[(size_t)blockIdx.x * T + t] = r;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int done = atomicAdd(p.counter, 1u);
    s_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {  // the last block to finish sums the per-block partials in block order
    __threadfence();
    for (int t = threadIdx.x; t < T; t += kWThreads) {
      double r = 0.0;
      for (unsigned int b = 0; b < gridDim.x; ++b) r += p.partial[(size_t)b * T + t];
      p.sum_w[t] = r;
    }
  }
}

constexpr int kPilotRows = 64;

// blockDim = (32 genes, 8 row groups)
__global__ void __launch_bounds__(256) pilot_kernel(const float* __restrict__ X, long long n_cells,
                                                    long long n_genes, long long ldx, float* pilot) {
  __shared__ float s[8][33];
  const long long g = (long long)blockIdx.x * 32 + threadIdx.x;
  const int ns = (int)(n_cells < kPilotRows ? n_cells : kPilotRows);
  const long long stride = n_cells / ns;
  float acc = 0.0f;
  if (g < n_genes) {
#pragma unroll 4
    for (int k = threadIdx.y; k < ns; k += 8) acc += X[(size_t)(k * stride) * ldx + g];
  }
  s[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && g < n_genes) {
    float r = 0.0f;
#pragma unroll
    for (int y = 0; y < 8; ++y) r += s[y][threadIdx.x];
    pilot[g] = r / (float)ns;
  }
}

}  // namespace

extern "C" int g2g_layout(int64_t n_cells, int32_t n_bins, g2g_layout_t* out) {
  G2G_REQUIRE(out != nullptr, "null out");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS, "n_bins out of range [2, 128]");
  G2G_REQUIRE(n_cells >= 1, "n_cells must be positive");
  G2GTiling t = g2g_tiling(n_bins);
  out->n_groups = t.n_groups;
  out->bins_per_group = t.bins_per_group;
  out->row_floats = t.row_floats;
  out->genes_per_tile = 32 * t.genes_per_thread;
  out->n_pad = g2g_round_up(n_cells, G2G_CELL_TILE);
  // one work item = cells_per_item rows x one gene tile.  Smaller items balance better over the
  // persistent CTAs but write more partial sums ((5T+2)/cells_per_item of the X bytes streamed);
  // the floor keeps that near 16%, and large inputs are cut into ~64 ranges.
  int64_t min_item = g2g_round_up(6 * (5 * (int64_t)t.bins_per_group + 2), G2G_CELL_TILE);
  int64_t by_n = g2g_round_up(g2g_ceil_div(n_cells, 64), G2G_CELL_TILE);
  int64_t item = by_n > min_item ? by_n : min_item;
  if (item > out->n_pad) item = out->n_pad;
  out->cells_per_item = item;
  out->n_split = g2g_ceil_div(out->n_pad, item);
  return G2G_OK;
}

extern "C" size_t g2g_weights_workspace_bytes(int64_t n_cells, int32_t n_bins) {
  int64_t n_pad = g2g_round_up(n_cells, G2G_CELL_TILE);
  int64_t blocks = g2g_ceil_div(n_pad, kWThreads);
  return (size_t)blocks * n_bins * sizeof(double) + 16;
}

extern "C" int g2g_weights(const double* tau, int64_t n_cells, const double* points, const double* denom,
                           int32_t n_bins, float* wtab, double* sum_w, void* workspace,
                           size_t workspace_bytes, void* stream) {
  G2G_REQUIRE(tau && points && denom && wtab && sum_w && workspace, "null pointer");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS, "n_bins out of range [2, 128]");
  G2G_REQUIRE(n_cells >= 1, "n_cells must be positive");
  G2G_REQUIRE(workspace_bytes >= g2g_weights_workspace_bytes(n_cells, n_bins), "workspace too small");
  G2G_REQUIRE(((uintptr_t)workspace & 7) == 0, "workspace must be 8-byte aligned");
  G2GTiling t = g2g_tiling(n_bins);
  WeightsParams p;
  p.tau = tau; p.points = points; p.denom = denom; p.wtab = wtab; p.sum_w = sum_w;
  p.n_cells = n_cells; p.n_pad = g2g_round_up(n_cells, G2G_CELL_TILE);
  p.n_bins = n_bins; p.n_groups = t.n_groups; p.bins_per_group = t.bins_per_group; p.row_floats = t.row_floats;
  int64_t blocks = g2g_ceil_div(p.n_pad, kWThreads);
  p.counter = reinterpret_cast<unsigned int*>(workspace);
  p.partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 16);
  cudaStream_t st = (cudaStream_t)stream;
  G2G_CUDA_TRY(cudaMemsetAsync(p.counter, 0, 16, st));
  weights_kernel<<<(unsigned)blocks, kWThreads, 0, st>>>(p);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}

extern "C" int g2g_pilot_mean(const float* X, int64_t n_cells, int64_t n_genes, int64_t ldx, float* pilot,
                              void* stream) {
  G2G_REQUIRE(X && pilot, "null pointer");
  G2G_REQUIRE(n_cells >= 1 && n_genes >= 1 && ldx >= n_genes, "bad shape");
  dim3 block(32, 8);
  pilot_kernel<<<(unsigned)g2g_ceil_div(n_genes, 32), block, 0, (cudaStream_t)stream>>>(X, n_cells, n_genes, ldx, pilot);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}
