// Device code shared by g2g_fixup.cu (stand-alone K3) and g2g_fused.cu (K3 + K4 in one launch).
#pragma once
#include "g2g_common.cuh"
#include <math.h>

namespace {

struct FixSide {
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
  long long n_genes;
  int n_bins;
};

struct Mask128 {
  unsigned long long lo, hi;
  __device__ void set(int b) { if (b < 64) lo |= 1ull << b; else hi |= 1ull << (b - 64); }
  __device__ bool get(int b) const { return b < 64 ? (lo >> b) & 1ull : (hi >> (b - 64)) & 1ull; }
  __device__ bool any() const { return (lo | hi) != 0ull; }
};

// Per-split partial sums of K1 and where their totals go.
struct ReduceParams {
  const double* part_f64[2];   // [n_split][2T+1][g_stride]
  const int32_t* part_i32[2];  // [n_split][T][g_stride]
  double* red_f64[2];          // [2T+1][g_stride]   (unused by the fused kernel: totals stay in shared memory)
  int32_t* red_i32[2];         // [T][g_stride]
  long long n_split[2];
  long long n_genes, g_stride;
  int n_bins;
};

// Sum of one (side, row, gene) over the splits, in split order (fixed => deterministic); the loads
// are issued kBatch at a time (the partial sums sit in L2, latency bound), the additions stay sequential.
template <int kBatch, typename V>
__device__ __forceinline__ V reduce_splits(const V* src, size_t plane, long long ns) {
  V r = 0;
  long long k = 0;
  for (; k + kBatch <= ns; k += kBatch) {
    V v[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) v[j] = src[(size_t)(k + j) * plane];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) r += v[j];
  }
  if (k < ns) {   // tail: predicated loads of one more batch
    V v[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) v[j] = (k + j < ns) ? src[(size_t)(k + j) * plane] : V(0);
#pragma unroll
    for (int j = 0; j < kBatch; ++j) r += v[j];   // + 0 past the end: r is never -0.0, so bits are unchanged
  }
  return r;
}
template <int kBatch>
__device__ __forceinline__ double reduce_f64_row(const ReduceParams& p, int side, int row, long long g) {
  return reduce_splits<kBatch>(p.part_f64[side] + (size_t)row * p.g_stride + g, (size_t)(2 * p.n_bins + 1) * p.g_stride,
                       p.n_split[side]);
}
template <int kBatch>
__device__ __forceinline__ int reduce_i32_row(const ReduceParams& p, int side, int q, long long g) {
  return reduce_splits<kBatch>(p.part_i32[side] + (size_t)q * p.g_stride + g, (size_t)p.n_bins * p.g_stride, p.n_split[side]);
}

// 128-bit helpers for the window masks (bit k = window [p_k, p_{k+1}))
__device__ __forceinline__ Mask128 mask_below(int k) {   // bits 0 .. k-1
  Mask128 m;
  m.lo = k >= 64 ? ~0ull : ((1ull << k) - 1ull);
  m.hi = k <= 64 ? 0ull : (k >= 128 ? ~0ull : ((1ull << (k - 64)) - 1ull));
  return m;
}
__device__ __forceinline__ int mask_msb(const Mask128& m) {   // highest set bit; m must not be empty
  return m.hi ? 127 - __clzll((long long)m.hi) : 63 - __clzll((long long)m.lo);
}
__device__ __forceinline__ int mask_lsb(const Mask128& m) {   // lowest set bit; m must not be empty
  return m.lo ? __ffsll((long long)m.lo) - 1 : 63 + __ffsll((long long)m.hi);
}

// Main.py:293-304 + 261-291 on the window flags (bit k = window [p_k, p_{k+1}) has <= 3 non-zero
// cells); returns the mask of interpolation points whose value appears in the significant-region
// list (np.in1d(points, regions), Main.py:356...).  The reference walks the flagged windows in
// order, groups adjacent ones into runs [a, b], keeps a run when p_{b+1} - p_a > 0.2 -- except that the
// run ending at the last flagged window is kept only when it has >= 2 windows (Main.py:283) -- and
// lists the points a .. b+1 of every kept run.  Here every lane takes the windows t = 32 c + lane, finds
// the bounds of the run its window belongs to with bit scans, decides `keep`, and the warp assembles the
// point mask from ballots: point t is listed iff window t or window t-1 lies in a kept run.  Must be
// called by all 32 lanes; every lane gets the whole mask.
__device__ __forceinline__ Mask128 region_mask(const Mask128& flagged, int T, const double* pts, int lane) {
  Mask128 m = {0ull, 0ull};
  if (!flagged.any()) return m;
  const int last_flag = mask_msb(flagged);
  const Mask128 clear = {~flagged.lo, ~flagged.hi};   // bits >= T-1 are never flagged, so a run always ends
  unsigned carry = 0u;
  const int n_chunks = (T + 31) / 32;
  for (int c = 0; c < n_chunks; ++c) {
    const int t = c * 32 + lane;
    bool kept = false;
    if (t + 1 < T && flagged.get(t)) {
      const Mask128 below = mask_below(t), upto = mask_below(t + 1);
      const Mask128 zb = {clear.lo & below.lo, clear.hi & below.hi};
      const Mask128 za = {clear.lo & ~upto.lo, clear.hi & ~upto.hi};
      const int a = zb.any() ? mask_msb(zb) + 1 : 0;   // first window of the run
      const int b = mask_lsb(za) - 1;                   // last window of the run
      const bool long_enough = (pts[b + 1] - pts[a]) > 0.2;
      kept = long_enough && (b != last_flag || b > a);
    }
    const unsigned k32 = __ballot_sync(0xffffffffu, kept);
    const unsigned pts32 = k32 | (k32 << 1) | carry;
    carry = k32 >> 31;
    if (c < 2) m.lo |= (unsigned long long)pts32 << (32 * c);
    else m.hi |= (unsigned long long)pts32 << (32 * (c - 2));
  }
  return m;
}

constexpr int kMaxChunks = G2G_MAX_BINS / 32;

// One warp finishes one gene; lanes run over the interpolation points.  `tot_f64[s]` / `tot_i32[s]` hold the
// split totals of side s as [row][gs] arrays, the gene at column `gi` (global memory for the stand-alone
// kernel, shared memory in the fused one); `g` is the gene's index in pilot[] and the outputs.
// NCH = compile-time bound on the number of 32-point chunks (ceil(T / 32) <= NCH): the callers that know
// it (the fused kernel, per DP column count) get correspondingly less code.
template <int NCH = kMaxChunks>
__device__ __forceinline__ void fixup_gene(const FixParams& p, const double* const* tot_f64,
                                           const int32_t* const* tot_i32, long long gs, long long gi, long long g,
                                           int lane) {
  const int T = p.n_bins;
  const int n_chunks = (T + 31) / 32;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);

  double mean[2][NCH], sd[2][NCH], min_sd[2];
  int nz[2];
  Mask128 flagged[2];
#pragma unroll
  for (int sd_i = 0; sd_i < 2; ++sd_i) {
    const FixSide& s = p.side[sd_i];
    const double a = (double)s.pilot[g];
    const double n = (double)s.n_cells;
    const double delta = tot_f64[sd_i][(size_t)(2 * T) * gs + gi] / n;  // xbar - a
    double mn = inf;
    int cnt = 0;
    flagged[sd_i].lo = flagged[sd_i].hi = 0ull;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
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
          const double A = tot_f64[sd_i][(size_t)t * gs + gi];
          const double Bq = tot_f64[sd_i][(size_t)(T + t) * gs + gi];
          double wss = Bq - 2.0 * delta * (A - a * sw) + delta * delta * sw;
          if (wss < 0.0) wss = 0.0;
          mean[sd_i][c] = A / sw;
          sd[sd_i][c] = sqrt(wss / (sw * (n - 1.0) / n)) * (sw / n);
          cw = tot_i32[sd_i][(size_t)t * gs + gi];
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
  // a side's low-expression regions matter exactly when that side is expressed (one call site per side
  // keeps the code small; zr / zq are warp-uniform)
  if (!zr) mask[0] = region_mask(flagged[0], T, p.points, lane);
  if (!zq) mask[1] = region_mask(flagged[1], T, p.points, lane);
  if (zr && zq) {
    flags = 0; common = 0.01; mode[0] = 2; mode[1] = 2;
  } else if (zr) {
    flags = 1; common = min_sd[1] / 10.0; mode[0] = 2;
    if (mask[1].any()) { mode[1] = 1; flags |= G2G_FLAG_REGION_QUERY; }
  } else if (zq) {
    flags = 2; common = min_sd[0] / 10.0; mode[1] = 2;
    if (mask[0].any()) { mode[0] = 1; flags |= G2G_FLAG_REGION_REF; }
  } else {
    flags = 3;
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
    for (int c = 0; c < NCH; ++c) {
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
