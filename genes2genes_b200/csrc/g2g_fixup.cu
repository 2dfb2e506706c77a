// K3  g2g_fixup_trends: one thread per gene.  Finishes the interpolation moments, applies the
// low-expression policy of RefQueryAligner.run_interpolation (Main.py:344-406) and derives the
// mean/std trends of the 50-sample bins through the fixed epsilon table.
#include "g2g_common.cuh"
#include <math.h>

namespace {

struct FixSide {
  const double* part_f64;  // [2T+1][g_stride]  sums over all splits (stage 1 output)
  const int32_t* part_i32; // [T][g_stride]
  const float* pilot;      // [G]
  const double* sum_w;     // [T]
  long long n_cells;
};

struct FixParams {
  FixSide side[2];
  const double* points;
  const double* eps_mean;
  const double* eps_std;
  double* musigma;  // [G][4][T]
  double* trends;   // [G][4][T]
  int32_t* nnz;     // [G][2]
  int32_t* flags;   // [G]
  long long n_genes, g_stride;
  int n_bins;
};

struct Mask128 {
  unsigned long long lo, hi;
  __device__ void set(int b) { if (b < 64) lo |= 1ull << b; else hi |= 1ull << (b - 64); }
  __device__ bool get(int b) const { return b < 64 ? (lo >> b) & 1ull : (hi >> (b - 64)) & 1ull; }
  __device__ bool any() const { return (lo | hi) != 0ull; }
};

// Stage 1: sum the per-split partials of K1 in split order (fixed => deterministic), one thread per
// (side, row, gene); coalesced over genes.
struct ReduceParams {
  const double* part_f64[2];
  const int32_t* part_i32[2];
  double* red_f64[2];
  int32_t* red_i32[2];
  long long n_split[2];
  long long n_genes, g_stride;
  int n_bins;
};

__global__ void __launch_bounds__(256) reduce_partials_kernel(const ReduceParams p) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.n_genes) return;
  const int T = p.n_bins;
  const int rows_per_side = 3 * T + 1;
  const int side = blockIdx.y / rows_per_side;
  const int row = blockIdx.y % rows_per_side;
  const long long ns = p.n_split[side];
  if (row < 2 * T + 1) {
    const size_t plane = (size_t)(2 * T + 1) * p.g_stride;
    const double* src = p.part_f64[side] + (size_t)row * p.g_stride + g;
    double r = 0.0;
    for (long long k = 0; k < ns; ++k) r += src[(size_t)k * plane];
    p.red_f64[side][(size_t)row * p.g_stride + g] = r;
  } else {
    const int q = row - (2 * T + 1);
    const size_t plane = (size_t)T * p.g_stride;
    const int32_t* src = p.part_i32[side] + (size_t)q * p.g_stride + g;
    int r = 0;
    for (long long k = 0; k < ns; ++k) r += src[(size_t)k * plane];
    p.red_i32[side][(size_t)q * p.g_stride + g] = r;
  }
}

// Main.py:293-304 + 261-291 on the window flags (bit k = window [p_k, p_{k+1}) has <= 3 non-zero
// cells); returns the mask of interpolation points whose value appears in the significant-region
// list (np.in1d(points, regions), Main.py:356...).
__device__ Mask128 region_mask(const Mask128& flagged, int T, const double* pts) {
  Mask128 m = {0ull, 0ull};
  int last_flag = -1;
  for (int k = 0; k + 1 < T; ++k)
    if (flagged.get(k)) last_flag = k;
  if (last_flag < 0) return m;
  int run_start = -1;
  for (int k = 0; k + 1 < T; ++k) {
    const bool f = flagged.get(k);
    if (f && run_start < 0) run_start = k;
    const bool next_f = (k + 2 < T) && flagged.get(k + 1);
    if (f && !next_f) {  // run [run_start, k] ends here
      const bool is_last_run = (k == last_flag);
      const bool long_enough = (pts[k + 1] - pts[run_start]) > 0.2;
      // a run that ends at the last flagged window is kept only when it has >= 2 windows (Main.py:283)
      const bool keep = long_enough && (!is_last_run || k > run_start);
      if (keep)
        for (int b = run_start; b <= k + 1; ++b) m.set(b);
      run_start = -1;
    }
  }
  return m;
}

constexpr int kFixWarps = 4;
constexpr int kMaxChunks = G2G_MAX_BINS / 32;

// Stage 2: one warp per gene, lanes over interpolation points.
__global__ void __launch_bounds__(kFixWarps * 32) fixup_kernel(const FixParams p) {
  const long long g = (long long)blockIdx.x * kFixWarps + (threadIdx.x >> 5);
  if (g >= p.n_genes) return;
  const int lane = threadIdx.x & 31;
  const int T = p.n_bins;
  const long long gs = p.g_stride;
  const int n_chunks = (T + 31) / 32;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);

  double mean[2][kMaxChunks], sd[2][kMaxChunks], min_sd[2];
  int nz[2];
  Mask128 flagged[2];
#pragma unroll
  for (int sd_i = 0; sd_i < 2; ++sd_i) {
    const FixSide& s = p.side[sd_i];
    const double a = (double)s.pilot[g];
    const double n = (double)s.n_cells;
    const double delta = s.part_f64[(size_t)(2 * T) * gs + g] / n;  // xbar - a
    double mn = inf;
    int cnt = 0;
    flagged[sd_i].lo = flagged[sd_i].hi = 0ull;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      mean[sd_i][c] = 0.0;
      sd[sd_i][c] = inf;
      if (c < n_chunks) {
        const int t = c * 32 + lane;
        int cw = 0x7fffffff;
        if (t < T) {
          // weighted mean / std of one bin (TimeSeriesPreprocessor.py:164-174) from the single-pass
          // sums A = sum w x, B = sum w (x-a)^2 (a = pilot shift), with d = xbar - a:
          //   mean = A / sum w;   sum w (x - xbar)^2 = B - 2 d (A - a sum w) + d^2 sum w
          const double sw = s.sum_w[t];
          const double A = s.part_f64[(size_t)t * gs + g];
          const double Bq = s.part_f64[(size_t)(T + t) * gs + g];
          double wss = Bq - 2.0 * delta * (A - a * sw) + delta * delta * sw;
          if (wss < 0.0) wss = 0.0;
          mean[sd_i][c] = A / sw;
          sd[sd_i][c] = sqrt(wss / (sw * (n - 1.0) / n)) * (sw / n);
          cw = s.part_i32[(size_t)t * gs + g];
          cnt += cw;
        }
        mn = fmin(mn, sd[sd_i][c]);
        const unsigned bal = __ballot_sync(0xffffffffu, t + 1 < T && cw <= 3);
        if (c < 2) flagged[sd_i].lo |= (unsigned long long)bal << (32 * c);
        else flagged[sd_i].hi |= (unsigned long long)bal << (32 * (c - 2));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    min_sd[sd_i] = mn;
    nz[sd_i] = cnt;
  }

  // policy (Main.py:344-406).  mode per side: 0 estimate, 1 estimate + region override, 2 all bins user-given
  int flags;
  double common = 0.0;
  int mode[2] = {0, 0};
  Mask128 mask[2] = {{0ull, 0ull}, {0ull, 0ull}};
  const bool zr = nz[0] <= 3, zq = nz[1] <= 3;
  if (zr && zq) {
    flags = 0; common = 0.01; mode[0] = 2; mode[1] = 2;
  } else if (zr) {
    flags = 1; common = min_sd[1] / 10.0; mode[0] = 2;
    mask[1] = region_mask(flagged[1], T, p.points);
    if (mask[1].any()) { mode[1] = 1; flags |= G2G_FLAG_REGION_QUERY; }
  } else if (zq) {
    flags = 2; common = min_sd[0] / 10.0; mode[1] = 2;
    mask[0] = region_mask(flagged[0], T, p.points);
    if (mask[0].any()) { mode[0] = 1; flags |= G2G_FLAG_REGION_REF; }
  } else {
    flags = 3;
    mask[0] = region_mask(flagged[0], T, p.points);
    mask[1] = region_mask(flagged[1], T, p.points);
    if (mask[0].any() || mask[1].any()) {
      common = fmin(min_sd[0], min_sd[1]) / 10.0;
      if (mask[0].any()) { mode[0] = 1; flags |= G2G_FLAG_REGION_REF; }
      if (mask[1].any()) { mode[1] = 1; flags |= G2G_FLAG_REGION_QUERY; }
    }
  }

  bool bad = false;
#pragma unroll
  for (int sd_i = 0; sd_i < 2; ++sd_i) {
    double* ms = p.musigma + ((size_t)g * 4 + 2 * sd_i) * T;
    double* tr = p.trends + ((size_t)g * 4 + 2 * sd_i) * T;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int t = c * 32 + lane;
      if (c < n_chunks && t < T) {
        double m = mean[sd_i][c], s = sd[sd_i][c];
        const bool user = mode[sd_i] == 2 || (mode[sd_i] == 1 && mask[sd_i].get(t));
        if (user) { m = 0.0; s = common; }  // TSP.py:175-177
        if (!(s > 0.0) || isinf(s)) bad = true;
        ms[t] = m;
        ms[T + t] = s;
        tr[t] = m + s * p.eps_mean[t];   // np.mean(mu + sigma*eps_t)
        tr[T + t] = s * p.eps_std[t];    // np.std(mu + sigma*eps_t)
      }
    }
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    p.nnz[2 * g] = nz[0];
    p.nnz[2 * g + 1] = nz[1];
    p.flags[g] = flags | (bad ? G2G_FLAG_BAD_SIGMA : 0);
  }
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
    rp.n_split[s] = lay.n_split;
    p.side[s].part_f64 = rp.red_f64[s];
    p.side[s].part_i32 = rp.red_i32[s];
    p.side[s].pilot = sides[s].pilot;
    p.side[s].sum_w = s == 0 ? sum_w_ref : sum_w_query;
    p.side[s].n_cells = sides[s].n_cells;
  }
  p.points = points; p.eps_mean = eps_mean; p.eps_std = eps_std;
  p.musigma = musigma; p.trends = trends; p.nnz = nnz; p.flags = flags;
  p.n_genes = n_genes; p.g_stride = g_stride; p.n_bins = n_bins;
  dim3 rgrid((unsigned)g2g_ceil_div(n_genes, 256), (unsigned)(2 * (3 * n_bins + 1)));
  reduce_partials_kernel<<<rgrid, 256, 0, (cudaStream_t)stream>>>(rp);
  G2G_CUDA_TRY(cudaGetLastError());
  fixup_kernel<<<(unsigned)g2g_ceil_div(n_genes, kFixWarps), kFixWarps * 32, 0, (cudaStream_t)stream>>>(p);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}
