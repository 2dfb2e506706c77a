// K3  g2g_fixup_trends: one thread per gene.  Finishes the interpolation moments, applies the
// low-expression policy of RefQueryAligner.run_interpolation (Main.py:344-406) and derives the
// mean/std trends of the 50-sample bins through the fixed epsilon table.
#include "g2g_common.cuh"
#include <math.h>

namespace {

struct FixSide {
  const double* part_f64;  // [n_split][2T+1][g_stride]
  const int32_t* part_i32; // [n_split][T][g_stride]
  const float* pilot;      // [G]
  const double* sum_w;     // [T]
  long long n_cells;
  long long n_split;
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

__device__ double reduce_f64(const FixSide& s, long long q, long long g, long long g_stride, int T) {
  double r = 0.0;
  const size_t plane = (size_t)(2 * T + 1) * g_stride;
  for (long long k = 0; k < s.n_split; ++k) r += s.part_f64[(size_t)k * plane + (size_t)q * g_stride + g];
  return r;
}
__device__ int reduce_i32(const FixSide& s, long long q, long long g, long long g_stride, int T) {
  int r = 0;
  const size_t plane = (size_t)T * g_stride;
  for (long long k = 0; k < s.n_split; ++k) r += s.part_i32[(size_t)k * plane + (size_t)q * g_stride + g];
  return r;
}

// weighted mean / std of one bin (TimeSeriesPreprocessor.py:164-174), from the shifted sums:
//   mean = a + sum w(x-a) / sum w
//   sum w (x - xbar)^2 = sum w (x-a)^2 - 2 d sum w(x-a) + d^2 sum w,   d = xbar - a
__device__ void bin_stat(const FixSide& s, long long g, long long g_stride, int T, int t, double a,
                         double delta, double& mean, double& sd) {
  const double sw = s.sum_w[t];
  const double A = reduce_f64(s, t, g, g_stride, T);
  const double Bq = reduce_f64(s, T + t, g, g_stride, T);
  const double n = (double)s.n_cells;
  mean = a + A / sw;
  double wss = Bq - 2.0 * delta * A + delta * delta * sw;
  if (wss < 0.0) wss = 0.0;
  sd = sqrt(wss / (sw * (n - 1.0) / n)) * (sw / n);
}

// Main.py:293-304 + 261-291 on the per-window nnz counts; returns the mask of interpolation points
// whose value appears in the significant-region list (np.in1d(points, regions), Main.py:356...).
__device__ Mask128 region_mask(const FixSide& s, long long g, long long g_stride, int T, const double* pts) {
  Mask128 m = {0ull, 0ull};
  int last_flag = -1;
  for (int k = 0; k + 1 < T; ++k)
    if (reduce_i32(s, k, g, g_stride, T) <= 3) last_flag = k;
  if (last_flag < 0) return m;
  int run_start = -1;
  for (int k = 0; k + 1 < T; ++k) {
    const bool f = reduce_i32(s, k, g, g_stride, T) <= 3;
    if (f && run_start < 0) run_start = k;
    const bool next_f = (k + 2 < T) && (reduce_i32(s, k + 1, g, g_stride, T) <= 3);
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

__global__ void __launch_bounds__(128) fixup_kernel(const FixParams p) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.n_genes) return;
  const int T = p.n_bins;
  const long long gs = p.g_stride;
  double a[2], delta[2], min_sd[2];
  int nz[2];
  for (int sd_i = 0; sd_i < 2; ++sd_i) {
    const FixSide& s = p.side[sd_i];
    a[sd_i] = (double)s.pilot[g];
    delta[sd_i] = reduce_f64(s, 2 * T, g, gs, T) / (double)s.n_cells;  // xbar - a
    int c = 0;
    for (int k = 0; k < T; ++k) c += reduce_i32(s, k, g, gs, T);
    nz[sd_i] = c;
    double mn = __longlong_as_double(0x7ff0000000000000LL);
    for (int t = 0; t < T; ++t) {
      double mean, sd;
      bin_stat(s, g, gs, T, t, a[sd_i], delta[sd_i], mean, sd);
      mn = sd < mn ? sd : mn;
    }
    min_sd[sd_i] = mn;
  }
  p.nnz[2 * g] = nz[0];
  p.nnz[2 * g + 1] = nz[1];

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
    mask[1] = region_mask(p.side[1], g, gs, T, p.points);
    if (mask[1].any()) { mode[1] = 1; flags |= G2G_FLAG_REGION_QUERY; }
  } else if (zq) {
    flags = 2; common = min_sd[0] / 10.0; mode[1] = 2;
    mask[0] = region_mask(p.side[0], g, gs, T, p.points);
    if (mask[0].any()) { mode[0] = 1; flags |= G2G_FLAG_REGION_REF; }
  } else {
    flags = 3;
    mask[0] = region_mask(p.side[0], g, gs, T, p.points);
    mask[1] = region_mask(p.side[1], g, gs, T, p.points);
    if (mask[0].any() || mask[1].any()) {
      common = fmin(min_sd[0], min_sd[1]) / 10.0;
      if (mask[0].any()) { mode[0] = 1; flags |= G2G_FLAG_REGION_REF; }
      if (mask[1].any()) { mode[1] = 1; flags |= G2G_FLAG_REGION_QUERY; }
    }
  }

  bool bad = false;
  for (int sd_i = 0; sd_i < 2; ++sd_i) {
    const FixSide& s = p.side[sd_i];
    double* ms = p.musigma + ((size_t)g * 4 + 2 * sd_i) * T;
    double* tr = p.trends + ((size_t)g * 4 + 2 * sd_i) * T;
    for (int t = 0; t < T; ++t) {
      double mean, sd;
      const bool user = mode[sd_i] == 2 || (mode[sd_i] == 1 && mask[sd_i].get(t));
      if (user) { mean = 0.0; sd = common; }  // TSP.py:175-177
      else bin_stat(s, g, gs, T, t, a[sd_i], delta[sd_i], mean, sd);
      if (!(sd > 0.0) || isinf(sd)) bad = true;
      ms[t] = mean;
      ms[T + t] = sd;
      tr[t] = mean + sd * p.eps_mean[t];   // np.mean(mu + sigma*eps_t)
      tr[T + t] = sd * p.eps_std[t];       // np.std(mu + sigma*eps_t)
    }
  }
  if (bad) flags |= G2G_FLAG_BAD_SIGMA;
  p.flags[g] = flags;
}

}  // namespace

extern "C" int g2g_fixup_trends(const g2g_side_t* sides, const double* sum_w_ref, const double* sum_w_query,
                                const double* points, const double* eps_mean, const double* eps_std,
                                int64_t n_genes, int64_t g_stride, int32_t n_bins, double* musigma,
                                double* trends, int32_t* nnz, int32_t* flags, void* stream) {
  G2G_REQUIRE(sides && sum_w_ref && sum_w_query && points && eps_mean && eps_std, "null input pointer");
  G2G_REQUIRE(musigma && trends && nnz && flags, "null output pointer");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS, "n_bins out of range [2, 128]");
  G2G_REQUIRE(n_genes >= 0 && g_stride >= n_genes, "bad gene counts");
  if (n_genes == 0) return G2G_OK;
  FixParams p;
  for (int s = 0; s < 2; ++s) {
    G2G_REQUIRE(sides[s].part_f64 && sides[s].part_i32 && sides[s].pilot && sides[s].n_cells >= 1, "bad side");
    g2g_layout_t lay;
    int rc = g2g_layout(sides[s].n_cells, n_bins, &lay);
    if (rc != G2G_OK) return rc;
    p.side[s].part_f64 = sides[s].part_f64;
    p.side[s].part_i32 = sides[s].part_i32;
    p.side[s].pilot = sides[s].pilot;
    p.side[s].sum_w = s == 0 ? sum_w_ref : sum_w_query;
    p.side[s].n_cells = sides[s].n_cells;
    p.side[s].n_split = lay.n_split;
  }
  p.points = points; p.eps_mean = eps_mean; p.eps_std = eps_std;
  p.musigma = musigma; p.trends = trends; p.nnz = nnz; p.flags = flags;
  p.n_genes = n_genes; p.g_stride = g_stride; p.n_bins = n_bins;
  fixup_kernel<<<(unsigned)g2g_ceil_div(n_genes, 128), 128, 0, (cudaStream_t)stream>>>(p);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}
