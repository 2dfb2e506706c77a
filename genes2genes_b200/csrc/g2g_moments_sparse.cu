// K1s  g2g_moments_csc: sparse (gene-major, CSC) variant of g2g_moments.  Work and bytes are
// proportional to the number of non-zeros: a zero contributes nothing to
//     A = sum_c w[t,c] x_c,   B = sum_c w[t,c] x_c^2   (shift a = 0),   C = sum_c x_c
// nor to the per-window non-zero counts, so only stored entries are visited.  Replaces the same
// reference code as K1 (TimeSeriesPreprocessor.py:161-174, Main.py:298-301,344-363) without ever
// densifying the matrix (the reference calls .todense(): Main.py:232,238; TSP.py:28).
//
// One CTA of four warps per (gene, bin group).  Thread k takes non-zeros k, k+128, ... of the gene's
// column, gathers the 64-byte weight row of each non-zero's cell (L2-resident table) and accumulates
// TBE bins x {A,B} in f32 registers (<= ~nnz/128 terms each), then the CTA reduces in f64 in a fixed
// order (warp butterfly, then warps in index order): results are independent of scheduling and of
// which other genes are in the batch.
#include "g2g_common.cuh"

namespace {

constexpr int kSpWarps = 4;
constexpr int kSpThreads = kSpWarps * 32;

struct SpSide {
  const long long* col_ptr;
  const int* row_idx;
  const float* values;
  const float* wtab;
  float* pilot;
  double* part_f64;
  int32_t* part_i32;
  long long n_pad;
  long long n_split;
};

struct SpParams {
  SpSide side[2];
  long long n_genes, g_stride;
  int n_sides, n_bins, n_groups, bins_per_group;
};

template <int TBE>
__global__ void __launch_bounds__(kSpThreads) moments_csc_kernel(const SpParams p) {
  constexpr int RF = (TBE + 2 + 3) / 4 * 4;
  constexpr int NQ = 2 * TBE + 1;
  __shared__ double s_red[kSpWarps][NQ];
  __shared__ int s_cnt[G2G_MAX_BINS];
  const long long g = blockIdx.x;
  const int bg = blockIdx.y % p.n_groups;
  const SpSide& s = p.side[blockIdx.y / p.n_groups];
  const int T = p.n_bins, TB = p.bins_per_group;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < T; k += kSpThreads) s_cnt[k] = 0;
  __syncthreads();

  float accA[TBE], accB[TBE], sx = 0.0f;
#pragma unroll
  for (int t = 0; t < TBE; ++t) { accA[t] = 0.0f; accB[t] = 0.0f; }
  int cur_win = -1, cnt = 0;
  const long long p0 = s.col_ptr[g], p1 = s.col_ptr[g + 1];
  const float* wbase = s.wtab + (size_t)bg * s.n_pad * RF;

  auto visit = [&](int c, float x) {
    const float4* wrow = reinterpret_cast<const float4*>(wbase + (size_t)c * RF);
    float w[RF];
#pragma unroll
    for (int q = 0; q < RF / 4; ++q) {
      const float4 v = __ldg(wrow + q);
      w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
    const float y = x * x;
#pragma unroll
    for (int t = 0; t < TBE; ++t) {
      accA[t] = fmaf(w[t], x, accA[t]);
      accB[t] = fmaf(w[t], y, accB[t]);
    }
    sx += x;
    if (x != 0.0f) {  // explicit zeros may be stored; they are not "non-zero" (np.count_nonzero)
      int win = __float_as_int(w[TBE + 1]);
      if (win >= G2G_WIN_MIXED / 2) win -= G2G_WIN_MIXED;
      if (win != cur_win) {
        if (cnt && cur_win >= 0) atomicAdd(&s_cnt[cur_win], cnt);
        cur_win = win;
        cnt = 0;
      }
      ++cnt;
    }
  };

  long long k = p0 + tid;
  for (; k + kSpThreads < p1; k += 2 * kSpThreads) {  // two independent gathers in flight per thread
    const int c0 = s.row_idx[k], c1 = s.row_idx[k + kSpThreads];
    const float x0 = s.values[k], x1 = s.values[k + kSpThreads];
    visit(c0, x0);
    visit(c1, x1);
  }
  if (k < p1) visit(s.row_idx[k], s.values[k]);
  if (cnt && cur_win >= 0) atomicAdd(&s_cnt[cur_win], cnt);

  // f64 reduction: butterfly inside the warp, then the four warps in index order
  double red[NQ];
#pragma unroll
  for (int t = 0; t < TBE; ++t) { red[t] = (double)accA[t]; red[TBE + t] = (double)accB[t]; }
  red[2 * TBE] = (double)sx;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) red[q] += __shfl_xor_sync(0xffffffffu, red[q], o);
    if (lane == 0) s_red[warp][q] = red[q];
  }
  __syncthreads();
  double* pf = s.part_f64;
  for (int q = tid; q < NQ; q += kSpThreads) {
    const double sum = ((s_red[0][q] + s_red[1][q]) + s_red[2][q]) + s_red[3][q];
    int row = -1;
    if (q < TBE) { const int t = bg * TB + q; if (q < TB && t < T) row = t; }
    else if (q < 2 * TBE) { const int t = bg * TB + (q - TBE); if (q - TBE < TB && t < T) row = T + t; }
    else if (bg == 0) row = 2 * T;
    if (row >= 0) pf[(size_t)row * p.g_stride + g] = sum;   // split 0 holds the totals
  }
  if (bg == 0) {
    for (int q = tid; q < T; q += kSpThreads) s.part_i32[(size_t)q * p.g_stride + g] = s_cnt[q];
    if (tid == 0) s.pilot[g] = 0.0f;
  }
}

template <int TBE>
int launch_sparse(const SpParams& p, cudaStream_t stream) {
  dim3 grid((unsigned)p.n_genes, (unsigned)(p.n_sides * p.n_groups));
  moments_csc_kernel<TBE><<<grid, kSpThreads, 0, stream>>>(p);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}

#define G2G_SP_CASE(TBV) case TBV: return launch_sparse<TBV>(p, st);

}  // namespace

extern "C" int g2g_moments_csc(const g2g_sparse_side_t* sides, int32_t n_sides, int64_t n_genes, int64_t g_stride,
                               int32_t n_bins, void* stream) {
  G2G_REQUIRE(sides != nullptr && n_sides >= 1 && n_sides <= 2, "n_sides must be 1 or 2");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS, "n_bins out of range [2, 128]");
  G2G_REQUIRE(n_genes >= 1 && g_stride >= n_genes && n_genes < (1ll << 31), "bad gene counts");
  G2GTiling t = g2g_tiling(n_bins);
  SpParams p;
  p.n_genes = n_genes; p.g_stride = g_stride; p.n_sides = n_sides; p.n_bins = n_bins;
  p.n_groups = t.n_groups; p.bins_per_group = t.bins_per_group;
  cudaStream_t st = (cudaStream_t)stream;
  for (int s = 0; s < n_sides; ++s) {
    const g2g_sparse_side_t& in = sides[s];
    G2G_REQUIRE(in.col_ptr && in.row_idx && in.values && in.wtab && in.pilot && in.part_f64 && in.part_i32,
                "null pointer in side");
    G2G_REQUIRE(in.n_cells >= 1 && in.n_cells < (1ll << 31), "n_cells out of range");
    G2G_REQUIRE(((uintptr_t)in.wtab & 15) == 0, "wtab must be 16-byte aligned");
    g2g_layout_t lay;
    int rc = g2g_layout(in.n_cells, n_bins, &lay);
    if (rc != G2G_OK) return rc;
    p.side[s].col_ptr = reinterpret_cast<const long long*>(in.col_ptr);
    p.side[s].row_idx = in.row_idx; p.side[s].values = in.values; p.side[s].wtab = in.wtab;
    p.side[s].pilot = in.pilot; p.side[s].part_f64 = in.part_f64; p.side[s].part_i32 = in.part_i32;
    p.side[s].n_pad = lay.n_pad; p.side[s].n_split = lay.n_split;
    // the fix-up stage sums n_split partial planes: plane 0 gets the totals, the rest are zero
    if (lay.n_split > 1) {
      G2G_CUDA_TRY(cudaMemsetAsync(in.part_f64 + (size_t)(2 * n_bins + 1) * g_stride, 0,
                                   (size_t)(lay.n_split - 1) * (2 * n_bins + 1) * g_stride * sizeof(double), st));
      G2G_CUDA_TRY(cudaMemsetAsync(in.part_i32 + (size_t)n_bins * g_stride, 0,
                                   (size_t)(lay.n_split - 1) * n_bins * g_stride * sizeof(int32_t), st));
    }
  }
  switch (t.bins_even) {
    G2G_SP_CASE(2) G2G_SP_CASE(4) G2G_SP_CASE(6) G2G_SP_CASE(8) G2G_SP_CASE(10) G2G_SP_CASE(12) G2G_SP_CASE(14)
    G2G_SP_CASE(16) G2G_SP_CASE(18) G2G_SP_CASE(20) G2G_SP_CASE(22) G2G_SP_CASE(24) G2G_SP_CASE(26) G2G_SP_CASE(28)
    G2G_SP_CASE(30) G2G_SP_CASE(32) G2G_SP_CASE(34) G2G_SP_CASE(36) G2G_SP_CASE(38) G2G_SP_CASE(40) G2G_SP_CASE(42)
    G2G_SP_CASE(44) G2G_SP_CASE(46) G2G_SP_CASE(48) G2G_SP_CASE(50) G2G_SP_CASE(52) G2G_SP_CASE(54) G2G_SP_CASE(56)
    default:
      g2g_set_error("bins_even %d has no compiled kernel", t.bins_even);
      return G2G_ERR_UNSUPPORTED;
  }
}
