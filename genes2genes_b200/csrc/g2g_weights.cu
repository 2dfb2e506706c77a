// K0  g2g_weights     Gaussian-kernel weight table + per-bin weight sums for one dataset
// K1a g2g_pilot_mean  per-gene shift value from a fixed strided row sample
#include "g2g_common.cuh"
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

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
  int n_bins, n_groups, bins_per_group, bins_even, row_floats;
  unsigned char* mma_table;   // kMma: [n_groups][n_pad / 64] stages of g2g_mma_stage_bytes(n_cols) bytes
  int n_cols;                 // kMma: N of the MMA (bins_per_group + 1 rounded up to 16)
};

constexpr int kWCells = 64;                 // cells per block
constexpr int kWLanes = kWThreads / kWCells;  // threads per cell, striding over the bins

// W[t,c] = exp(-(|tau_c - p_t|)^2 / denom_t) in float64 (TimeSeriesPreprocessor.py:48, 125-127),
// rounded once to float32 for the table.  Block = 64 cells x 4 bin lanes.  The per-bin sums of the
// f64 weights are reduced in a fixed order: 16-cell partial sums, then 4 partials per block, then
// (last block to finish) the blocks in index order.
//
// kMma = true writes the table of the tensor-core kernel instead (g2g_moments_mma): per bin group and
// 64-cell stage, for each of the four 16-cell K steps the f16 B operand [n_cols x 16] of w_hi = f16(w) and
// of w_lo = f16((w - w_hi) * 2048) in the canonical K-major no-swizzle layout (8-row x 16-byte core matrices:
// element (n, k) at (k/8) * n_cols*16 + (n/8) * 128 + (n%8) * 16 + (k%8) * 2 bytes), then 64 window words
// (window id + 1, 0 for padding rows, bit 16 set when the 16-cell step spans several windows).
template <bool kMma>
__global__ void __launch_bounds__(kWThreads) weights_kernel(const WeightsParams p) {
  __shared__ double s_pts[G2G_MAX_BINS];
  __shared__ double s_den[G2G_MAX_BINS];
  extern __shared__ double s_dyn[];                     // [T][kWCells + 1] weights, then [T][4] partial sums
  double (*s_w)[kWCells + 1] = reinterpret_cast<double (*)[kWCells + 1]>(s_dyn);
  double (*s_part)[4] = reinterpret_cast<double (*)[4]>(s_dyn + (size_t)p.n_bins * (kWCells + 1));
  const int T = p.n_bins;
  __shared__ double s_g[G2G_MAX_BINS];   // exp(-(p_{t+1}^2 - p_t^2) / denom): the tau-independent part of w_{t+1} / w_t
  __shared__ double s_e[G2G_MAX_BINS];   // 2 (Delta_t - Delta_0) / denom: first-order term of the spacing's rounding noise
  __shared__ int s_slow;
  if (threadIdx.x == 0) s_slow = 0;
  for (int t = threadIdx.x; t < T; t += kWThreads) {
    s_pts[t] = p.points[t];
    s_den[t] = p.denom[t];
  }
  __syncthreads();
  // One kernel width for all bins and (nearly) equidistant points -- np.linspace and kernel_WINDOW_SIZE**2, i.e.
  // everything but the adaptive mode: consecutive weights of a cell then differ by the factor
  //   w_{t+1} / w_t = exp(-(p_{t+1}^2 - p_t^2) / denom) * exp(2 tau Delta_0 / denom) * (1 + tau e_t),
  // so a cell needs two exponentials per lane instead of one per bin (f64 exp is ~100 instructions); each lane
  // restarts from the direct formula, so at most T/4 rounding errors of 2^-53 accumulate.
  const double den0 = s_den[0], delta0 = T > 1 ? s_pts[1] - s_pts[0] : 0.0;
  for (int t = threadIdx.x; t + 1 < T; t += kWThreads) {
    const double e = 2.0 * ((s_pts[t + 1] - s_pts[t]) - delta0) / den0;
    s_g[t] = exp(-((s_pts[t + 1] * s_pts[t + 1] - s_pts[t] * s_pts[t]) / den0));
    s_e[t] = e;
    if (s_den[t + 1] != den0 || !(fabs(e) < 1e-9) || !(den0 > 0.0)) s_slow = 1;   // benign race: all write 1
  }
  __syncthreads();
  const bool fast = s_slow == 0;
  const int cl = threadIdx.x % kWCells, bl = threadIdx.x / kWCells;
  const long long c = (long long)blockIdx.x * kWCells + cl;
  const bool in_pad = c < p.n_pad;
  const bool valid = c < p.n_cells;
  const double tau = valid ? p.tau[c] : 0.0;
  if (fast) {
    const int per = (T + kWLanes - 1) / kWLanes;   // contiguous bins per lane
    const int t0 = bl * per, t1 = min(T, t0 + per);
    if (t0 < t1) {
      double w = 0.0;
      if (valid) {
        const double d = fabs(tau - s_pts[t0]);
        w = exp(-((d * d) / den0));
      }
      const double r = exp(2.0 * tau * delta0 / den0);
      s_w[t0][cl] = w;
      for (int t = t0; t + 1 < t1; ++t) {
        w = w * (s_g[t] * r) * fma(tau, s_e[t], 1.0);
        s_w[t + 1][cl] = w;
      }
    }
  } else {
    for (int t = bl; t < T; t += kWLanes) {
      double w = 0.0;
      if (valid) {
        double d = fabs(tau - s_pts[t]);
        w = exp(-((d * d) / s_den[t]));
      }
      s_w[t][cl] = w;
    }
  }
  __syncthreads();
  if constexpr (kMma) {
    const int N = p.n_cols;
    const size_t stage_bytes = g2g_mma_stage_bytes(N);
    const size_t n_stages = (size_t)(p.n_pad / kWCells);
    if ((long long)blockIdx.x * kWCells < p.n_pad) {
      for (int g = 0; g < p.n_groups; ++g) {
        unsigned char* st = p.mma_table + ((size_t)g * n_stages + blockIdx.x) * stage_bytes;
        // one thread per (column n, cell pair): 32 pairs per column
        for (int idx = threadIdx.x; idx < N * 32; idx += kWThreads) {
          const int n = idx >> 5, pr = idx & 31, c0 = pr * 2;          // cells c0, c0 + 1 of the stage
          const int t = g * p.bins_per_group + n;
          double w[2] = {0.0, 0.0};
          if (n < p.bins_per_group && t < T) { w[0] = s_w[t][c0]; w[1] = s_w[t][c0 + 1]; }
          const __half h0 = __double2half(w[0]), h1 = __double2half(w[1]);
          const __half l0 = __double2half((w[0] - (double)__half2float(h0)) * 2048.0);
          const __half l1 = __double2half((w[1] - (double)__half2float(h1)) * 2048.0);
          const int ks = c0 >> 4, k = c0 & 15;
          const size_t off = (size_t)ks * (N * 64) + (size_t)(k >> 3) * (N * 16) + (size_t)(n >> 3) * 128 +
                             (size_t)(n & 7) * 16 + (size_t)(k & 7) * 2;
          *reinterpret_cast<__half2*>(st + off) = __halves2half2(h0, h1);
          *reinterpret_cast<__half2*>(st + off + N * 32) = __halves2half2(l0, l1);
        }
        if (threadIdx.x < kWCells) {   // warps 0, 1: lanes are consecutive cells
          int win = 0;
          if (valid) {
            win = T;   // cells outside every window count at index T - 1
            for (int k = 0; k + 1 < T; ++k)
              if (tau >= s_pts[k] && tau < s_pts[k + 1]) win = k + 1;
          }
          const unsigned same = __match_any_sync(0xffffffffu, win);
          const unsigned group = 0xffffu << ((threadIdx.x & 31) & ~15);
          if ((same & group) != group) win |= G2G_WIN_MIXED;
          reinterpret_cast<int*>(st + (size_t)N * 256)[threadIdx.x] = win;
        }
      }
    }
  } else
  // table rows: [w_0 .. w_{TB-1}, (zero bin when TB is odd), valid, window id, 0 pad] per group
  if (in_pad) {
    for (int g = 0; g < p.n_groups; ++g) {
      float* row = p.wtab + ((size_t)g * p.n_pad + (size_t)c) * p.row_floats;
      for (int b = bl; b < p.row_floats; b += kWLanes) {
        const int t = g * p.bins_per_group + b;
        float out = 0.0f;
        if (b < p.bins_per_group) {
          if (t < T) out = (float)s_w[t][cl];
        } else if (b == p.bins_even) {
          out = valid ? 1.0f : 0.0f;
        } else if (b == p.bins_even + 1) {
          // window k: p_k <= tau < p_{k+1}  (half open, Main.py:299); cells outside every window -> T-1
          int win = -1;
          if (valid) {
            win = T - 1;
            for (int k = 0; k + 1 < T; ++k)
              if (tau >= s_pts[k] && tau < s_pts[k + 1]) win = k;
          }
          // all 32 lanes (consecutive cells) are here together: b depends on the warp only
          const unsigned same = __match_any_sync(0xffffffffu, win);
          const unsigned group = 0xffu << ((threadIdx.x & 31) & ~(G2G_WIN_GROUP - 1));
          if ((same & group) != group) win += G2G_WIN_MIXED;
          out = __int_as_float(win);
        }
        row[b] = out;
      }
    }
  }
  for (int k = threadIdx.x; k < T * 4; k += kWThreads) {
    const int t = k >> 2, q = k & 3;
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < kWCells / 4; ++i) r += s_w[t][q * (kWCells / 4) + i];
    s_part[t][q] = r;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += kWThreads)
    p.partial[(size_t)blockIdx.x * T + t] = ((s_part[t][0] + s_part[t][1]) + s_part[t][2]) + s_part[t][3];
}

// Per-bin totals of the per-block partial sums, in a fixed order: 16 interleaved chains per bin (block b goes
// to chain b % 16, in block order), then the chains in order.  One warp per bin, so that the chains of all
// bins run side by side (one block summing everything took a third of the weights kernel's time at 100 k
// cells); 16 loads in flight per chain, additions strictly in block order.
__global__ void __launch_bounds__(32) weights_sum_kernel(const double* __restrict__ partial, int n_blocks, int T,
                                                         double* __restrict__ sum_w) {
  const int t = blockIdx.x, q = threadIdx.x;
  double r = 0.0;
  if (q < 16) {
    int b = q;
    for (; b + 15 * 16 < n_blocks; b += 16 * 16) {
      double v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = partial[(size_t)(b + 16 * j) * T + t];
#pragma unroll
      for (int j = 0; j < 16; ++j) r += v[j];
    }
    for (; b < n_blocks; b += 16) r += partial[(size_t)b * T + t];
  }
  double total = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) total += __shfl_sync(0xffffffffu, r, k);
  if (q == 0) sum_w[t] = total;
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
  // the floor keeps that near 8%, and large inputs are cut into ~64 ranges
  // (measured on C2: profiles/r01_k1_item_size_sweep.txt).
  int64_t min_item = g2g_round_up(12 * (5 * (int64_t)t.bins_per_group + 2), G2G_CELL_TILE);
  int64_t by_n = g2g_round_up(g2g_ceil_div(n_cells, 64), G2G_CELL_TILE);
  int64_t item = by_n > min_item ? by_n : min_item;
  if (item > out->n_pad) item = out->n_pad;
  out->cells_per_item = item;
  out->n_split = g2g_ceil_div(out->n_pad, item);
  return G2G_OK;
}

extern "C" size_t g2g_weights_workspace_bytes(int64_t n_cells, int32_t n_bins) {
  int64_t n_pad = g2g_round_up(n_cells, G2G_CELL_TILE);
  int64_t blocks = g2g_ceil_div(n_pad, 64);
  return (size_t)blocks * n_bins * sizeof(double) + 16;
}

namespace {
int launch_weights(const double* tau, int64_t n_cells, const double* points, const double* denom, int32_t n_bins,
                   float* wtab, void* mma_table, double* sum_w, void* workspace, size_t workspace_bytes,
                   void* stream) {
  G2G_REQUIRE(tau && points && denom && (wtab || mma_table) && sum_w && workspace, "null pointer");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS, "n_bins out of range [2, 128]");
  G2G_REQUIRE(n_cells >= 1, "n_cells must be positive");
  G2G_REQUIRE(workspace_bytes >= g2g_weights_workspace_bytes(n_cells, n_bins), "workspace too small");
  G2G_REQUIRE(((uintptr_t)workspace & 7) == 0, "workspace must be 8-byte aligned");
  WeightsParams p;
  p.tau = tau; p.points = points; p.denom = denom; p.wtab = wtab; p.sum_w = sum_w;
  p.n_cells = n_cells; p.n_pad = g2g_round_up(n_cells, G2G_CELL_TILE);
  p.n_bins = n_bins;
  p.mma_table = reinterpret_cast<unsigned char*>(mma_table);
  if (mma_table) {
    G2G_REQUIRE(((uintptr_t)mma_table & 127) == 0, "table must be 128-byte aligned");
    G2GMmaTiling t = g2g_mma_tiling(n_bins);
    p.n_groups = t.n_groups; p.bins_per_group = t.bins_per_group; p.n_cols = t.n_cols;
    p.bins_even = 0; p.row_floats = 0;
  } else {
    G2GTiling t = g2g_tiling(n_bins);
    p.n_groups = t.n_groups; p.bins_per_group = t.bins_per_group; p.bins_even = t.bins_even;
    p.row_floats = t.row_floats; p.n_cols = 0;
  }
  int64_t blocks = g2g_ceil_div(p.n_pad, kWCells);
  p.counter = reinterpret_cast<unsigned int*>(workspace);
  p.partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 16);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)n_bins * (kWCells + 1 + 4) * sizeof(double);
  if (mma_table) {
    if (smem > 48 * 1024)
      G2G_CUDA_TRY(cudaFuncSetAttribute(weights_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    weights_kernel<true><<<(unsigned)blocks, kWThreads, smem, st>>>(p);
  } else {
    if (smem > 48 * 1024)
      G2G_CUDA_TRY(cudaFuncSetAttribute(weights_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    weights_kernel<false><<<(unsigned)blocks, kWThreads, smem, st>>>(p);
  }
  G2G_CUDA_TRY(cudaGetLastError());
  weights_sum_kernel<<<(unsigned)n_bins, 32, 0, st>>>(p.partial, (int)blocks, n_bins, sum_w);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}
}  // namespace

extern "C" int g2g_weights(const double* tau, int64_t n_cells, const double* points, const double* denom,
                           int32_t n_bins, float* wtab, double* sum_w, void* workspace,
                           size_t workspace_bytes, void* stream) {
  G2G_REQUIRE(wtab, "null pointer");
  return launch_weights(tau, n_cells, points, denom, n_bins, wtab, nullptr, sum_w, workspace, workspace_bytes, stream);
}

extern "C" int g2g_weights_mma(const double* tau, int64_t n_cells, const double* points, const double* denom,
                               int32_t n_bins, void* table, double* sum_w, void* workspace,
                               size_t workspace_bytes, void* stream) {
  G2G_REQUIRE(table, "null pointer");
  return launch_weights(tau, n_cells, points, denom, n_bins, nullptr, table, sum_w, workspace, workspace_bytes, stream);
}

extern "C" int g2g_mma_layout(int64_t n_cells, int32_t n_bins, g2g_mma_layout_t* out) {
  G2G_REQUIRE(out != nullptr, "null out");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS, "n_bins out of range [2, 128]");
  G2G_REQUIRE(n_cells >= 1, "n_cells must be positive");
  G2GMmaTiling t = g2g_mma_tiling(n_bins);
  out->n_groups = t.n_groups;
  out->bins_per_group = t.bins_per_group;
  out->n_cols = t.n_cols;
  out->reserved = 0;
  out->n_pad = g2g_round_up(n_cells, G2G_CELL_TILE);
  // one work item = 128 genes x cells_per_item cells; a function of the cell count only (shard-invariant
  // sums).  Items are whole 256-cell accumulation chains; about 24 per gene tile and side, at most 4096
  // cells (the per-item f64 partial sums are then 5 % of the bytes streamed).
  int64_t item = g2g_round_up(g2g_ceil_div(n_cells, 24), 256);
  if (item < 1024) item = 1024;
  if (item > 4096) item = 4096;
  if (item > out->n_pad) item = out->n_pad;
  out->cells_per_item = item;
  out->n_split = g2g_ceil_div(out->n_pad, item);
  out->table_bytes = (int64_t)t.n_groups * (out->n_pad / G2G_CELL_TILE) * (int64_t)g2g_mma_stage_bytes(t.n_cols);
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
