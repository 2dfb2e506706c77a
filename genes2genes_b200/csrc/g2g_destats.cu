// Rank statistics of the reference's per-matched-pair DE table.
// For every matched (reference bin s, query bin t) pair of a gene the reference runs
// scipy.stats.wilcoxon and scipy.stats.ks_2samp on the two 50-sample bins (Main.py:791-835, calls at
// :819-820).  The bins are mean + eps * std with the fixed epsilon table, so the rank statistics only
// need the interpolated (mean, std) of the two bins: one warp per pair rebuilds the 2 x 50 samples in
// registers and computes
//   * Wilcoxon signed-rank: r_plus (sum of average ranks of |d| over d > 0), the number of non-zero
//     differences and the tie term sum(t^3 - t);
//   * Kolmogorov-Smirnov: h = max_v |#{x <= v} - #{y <= v}| over the pooled samples (D = h / 50).
// p-values are exact functions of these integers; the host maps them with scipy's own null
// distributions (genes2genes_b200/DEStats.py).
#include "g2g_common.cuh"

namespace {

constexpr int kN = 50;  // samples per bin (TimeSeriesPreprocessor.py:179)

struct DeParams {
  const double* musigma;  // [G][4][T]
  const double* eps;      // [T][50]
  const int* pair_gene;
  const int* pair_s;
  const int* pair_t;
  double* r_plus;
  double* tie_term;
  int* n_nonzero;
  int* ks_h;
  long long n_pairs;
  int n_bins;
};

__global__ void __launch_bounds__(128) de_stats_kernel(const DeParams p) {
  const long long pair = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (pair >= p.n_pairs) return;
  const int lane = threadIdx.x & 31;
  const int T = p.n_bins;
  const int s = p.pair_s[pair], t = p.pair_t[pair];
  const double* ms = p.musigma + (size_t)p.pair_gene[pair] * 4 * T;
  const double muS = ms[s], sdS = ms[T + s], muT = ms[2 * T + t], sdT = ms[3 * T + t];
  // samples k = lane and lane + 32; Normal(mu, sigma).rsample() == loc + eps * scale, no FMA
  double x[2], y[2], d[2];
  bool ok[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int k = lane + 32 * r;
    ok[r] = k < kN;
    const double es = ok[r] ? p.eps[(size_t)s * kN + k] : 0.0, et = ok[r] ? p.eps[(size_t)t * kN + k] : 0.0;
    x[r] = __dadd_rn(muS, __dmul_rn(es, sdS));
    y[r] = __dadd_rn(muT, __dmul_rn(et, sdT));
    d[r] = __dsub_rn(x[r], y[r]);
  }
  // ---- Wilcoxon: average ranks of |d| among the non-zero differences ---------------------------
  int less[2] = {0, 0}, equal[2] = {0, 0};
  int cx_x[2] = {0, 0}, cy_x[2] = {0, 0}, cx_y[2] = {0, 0}, cy_y[2] = {0, 0};
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    for (int j = 0; j < 32; ++j) {
      const int k = j + 32 * rr;
      if (k >= kN) break;
      const double dj = __shfl_sync(0xffffffffu, d[rr], j);
      const double xj = __shfl_sync(0xffffffffu, x[rr], j);
      const double yj = __shfl_sync(0xffffffffu, y[rr], j);
      const double aj = fabs(dj);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const double a = fabs(d[r]);
        if (dj != 0.0) { less[r] += (aj < a); equal[r] += (aj == a); }
        cx_x[r] += (xj <= x[r]);  cy_x[r] += (yj <= x[r]);
        cx_y[r] += (xj <= y[r]);  cy_y[r] += (yj <= y[r]);
      }
    }
  }
  double rp = 0.0, tie = 0.0;
  int nz = 0, h = 0;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (ok[r]) {
      if (d[r] != 0.0) {
        ++nz;
        const double rank = 1.0 + (double)less[r] + 0.5 * (double)(equal[r] - 1);
        if (d[r] > 0.0) rp += rank;
        tie += (double)equal[r] * (double)equal[r] - 1.0;   // sum over groups of t^3 - t
      }
      const int h1 = abs(cx_x[r] - cy_x[r]), h2 = abs(cx_y[r] - cy_y[r]);
      h = max(h, max(h1, h2));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rp += __shfl_xor_sync(0xffffffffu, rp, o);     // ranks are multiples of 0.5: exact in f64
    tie += __shfl_xor_sync(0xffffffffu, tie, o);
    nz += __shfl_xor_sync(0xffffffffu, nz, o);
    h = max(h, __shfl_xor_sync(0xffffffffu, h, o));
  }
  if (lane == 0) {
    p.r_plus[pair] = rp;
    p.tie_term[pair] = tie;
    p.n_nonzero[pair] = nz;
    p.ks_h[pair] = h;
  }
}

}  // namespace

extern "C" int g2g_de_statistics(const double* musigma, const double* eps, int32_t n_bins, const int32_t* pair_gene,
                                 const int32_t* pair_s, const int32_t* pair_t, int64_t n_pairs, double* r_plus,
                                 double* tie_term, int32_t* n_nonzero, int32_t* ks_h, void* stream) {
  G2G_REQUIRE(musigma && eps && pair_gene && pair_s && pair_t, "null input pointer");
  G2G_REQUIRE(r_plus && tie_term && n_nonzero && ks_h, "null output pointer");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS && n_pairs >= 0, "bad shape");
  if (n_pairs == 0) return G2G_OK;
  DeParams p;
  p.musigma = musigma; p.eps = eps; p.pair_gene = pair_gene; p.pair_s = pair_s; p.pair_t = pair_t;
  p.r_plus = r_plus; p.tie_term = tie_term; p.n_nonzero = n_nonzero; p.ks_h = ks_h;
  p.n_pairs = n_pairs; p.n_bins = n_bins;
  de_stats_kernel<<<(unsigned)g2g_ceil_div(n_pairs, 4), 128, 0, (cudaStream_t)stream>>>(p);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}
