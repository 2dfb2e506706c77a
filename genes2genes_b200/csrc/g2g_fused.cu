// K3 + K4 in one launch: split totals, per-gene fix-up (g2g_fixup_trends) and cost + DP + traceback
// (g2g_cost_dp) for the production path, where nothing but the compact results is wanted.  Each block
// serves kWarpsPerBlock * GPW adjacent genes: (1) all threads add up K1's per-split partial sums of those
// genes into shared memory (coalesced over the adjacent genes, same order as reduce_partials_kernel);
// (2) one warp per gene finishes it (fixup_gene) and writes musigma / trends / nnz / flags; (3) the first
// kWarpsPerBlock warps align GPW genes each (dp_genes).  Results are bit-identical to the two separate calls.
#include "g2g_fixup_body.cuh"
#include "g2g_dp_body.cuh"

namespace {

struct FusedParams {
  FixParams fix;
  ReduceParams red;
  DpParams dp;
};

__host__ __device__ inline size_t totals_smem_bytes(int T, int genes_per_block) {
  return (size_t)2 * genes_per_block * ((size_t)(2 * T + 1) * sizeof(double) + (size_t)T * sizeof(int32_t));
}

// Block = one warp per gene of the block (GPC = WPB * GPW warps): all of them add up the
// split totals and finish one gene each; the first WPB warps then align GPW genes each.
// WPB = kWarpsPerBlock (4) in general.  kWideBlockWarps (17) for two DP columns per lane: the DP needs 122
// registers, i.e. 16 resident warps per SM in 4-warp blocks -- 2368 genes per wave on 148 SMs, so a 2500-gene
// shard (C4 over 8 GPUs) pays a second, nearly empty wave.  One 17-warp block per SM (544 threads, capped at 96
// registers by the scheduler that gets five of the warps; 96 bytes of spills) holds 17 x 148 = 2516 genes.
constexpr int kWideBlockWarps = 17;
template <int B, int LW, int WPB>
__global__ void __launch_bounds__(WPB * (32 / LW) * 32) fixup_dp_kernel(const __grid_constant__ FusedParams q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int GPW = 32 / LW;
  constexpr int GPC = WPB * GPW;   // genes per block = warps per block
  constexpr int kThreads = GPC * 32;
  constexpr int kLoads = B == 1 ? 24 : 12;   // loads in flight per thread in the split reduction (register room)
  const int T = q.fix.n_bins;
  const int rows_f = 2 * T + 1;
  double* tot_f = reinterpret_cast<double*>(smem_raw);                  // [2][2T+1][GPC]
  int32_t* tot_i = reinterpret_cast<int32_t*>(tot_f + (size_t)2 * rows_f * GPC);  // [2][T][GPC]
  const long long g0 = (long long)blockIdx.x * GPC;
  G2G_DEVICE_ASSERT(g0 < q.fix.n_genes && blockDim.x == kThreads);

  const int rows = rows_f + T;
  for (int idx = threadIdx.x; idx < 2 * rows * GPC; idx += kThreads) {
    const int gl = idx % GPC, r = idx / GPC;
    const int side = r / rows, row = r - side * rows;
    const long long g = g0 + gl;
    if (g < q.fix.n_genes) {
      if (row < rows_f) tot_f[((size_t)side * rows_f + row) * GPC + gl] = reduce_f64_row<kLoads>(q.red, side, row, g);
      else tot_i[((size_t)side * T + (row - rows_f)) * GPC + gl] = reduce_i32_row<kLoads>(q.red, side, row - rows_f, g);
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* tf[2] = {tot_f, tot_f + (size_t)rows_f * GPC};
  const int32_t* ti[2] = {tot_i, tot_i + (size_t)T * GPC};
  if (g0 + warp < q.fix.n_genes) fixup_gene<(B < kMaxChunks ? B : kMaxChunks)>(q.fix, tf, ti, GPC, warp, g0 + warp, lane);
  __syncthreads();   // trends are visible to the block; the DP scratch below reuses the totals' shared memory
  if (warp < WPB) dp_genes<B, false, LW, WPB>(q.dp, smem_raw);
}

template <int WPB, int GPW>
size_t fused_smem_bytes(int T) {
  size_t smem = dp_smem_per_warp(T) * WPB * GPW;
  const size_t tot = totals_smem_bytes(T, WPB * GPW);
  return tot > smem ? tot : smem;
}

template <int B, int LW, int WPB>
int launch_fused_mode(const FusedParams& q, cudaStream_t stream) {
  constexpr int GPW = 32 / LW;
  const size_t smem = fused_smem_bytes<WPB, GPW>(q.fix.n_bins);
  if (smem > 48 * 1024) {
    G2G_CUDA_TRY(cudaFuncSetAttribute(fixup_dp_kernel<B, LW, WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const unsigned blocks = (unsigned)g2g_ceil_div(q.fix.n_genes, WPB * GPW);
  fixup_dp_kernel<B, LW, WPB><<<blocks, WPB * GPW * 32, smem, stream>>>(q);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}

// Two DP columns per lane: one 17-warp block per SM exactly when the launch fits one wave of those but not one
// wave of the 4-warp blocks (16 x SMs < genes <= 17 x SMs).  Measured at T = 50 on 148 SMs: 2500 genes 0.184 ms
// against 0.228 ms (the 4-warp blocks' second wave holds 132 genes and still lasts one whole DP, 0.09 ms); with
// more waves the wide blocks lose (96 registers, five warps on one scheduler: 0.184 ms per wave of 2516 genes
// against 0.139 ms per wave of 2368; 10 000 genes 0.685 against 0.586 ms).  A function of the launch's gene count
// and the device only -- the records do not depend on the block shape (same per-gene arithmetic and order).
inline bool wide_blocks_fit_one_wave(const FusedParams& q) {
#ifdef G2G_DP_NARROW_BLOCKS
  return false;   // profiling builds: always the 4-warp blocks
#endif
  const size_t smem = fused_smem_bytes<kWideBlockWarps, 1>(q.fix.n_bins);
  int dev = 0, sms = 0, smem_max = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return false;
  if (sms <= 0 || smem + 1024 > (size_t)smem_max) return false;
  const long long narrow_wave = 4ll * kWarpsPerBlock * sms;   // 4 resident 4-warp blocks per SM (122 -> 128 registers)
  const long long wide_wave = (long long)kWideBlockWarps * sms;
  return q.fix.n_genes > narrow_wave && q.fix.n_genes <= wide_wave;
}

template <int B>
int launch_fused(const FusedParams& q, cudaStream_t stream) {
  if (B == 1) {  // short series: several genes per warp, as in g2g_cost_dp
    if (q.fix.n_bins + 1 <= 8) return launch_fused_mode<1, 8, kWarpsPerBlock>(q, stream);
    if (q.fix.n_bins + 1 <= 16) return launch_fused_mode<1, 16, kWarpsPerBlock>(q, stream);
  }
  if (B == 2 && wide_blocks_fit_one_wave(q)) return launch_fused_mode<2, 32, kWideBlockWarps>(q, stream);
  return launch_fused_mode<B, 32, kWarpsPerBlock>(q, stream);
}

}  // namespace

extern "C" int g2g_fixup_cost_dp(const g2g_side_t* sides, const double* sum_w_ref, const double* sum_w_query,
                                 const double* points, const double* eps_mean, const double* eps_std,
                                 int64_t n_genes, int64_t g_stride, int32_t n_bins, const double* trans,
                                 const double* start, double c_model, int32_t n_samples, double* musigma,
                                 double* trends, int32_t* nnz, int32_t* flags, uint8_t* align_str,
                                 int32_t* align_len, double* opt_cost, uint8_t* l_states, void* stream) {
  G2G_REQUIRE(sides && sum_w_ref && sum_w_query && points && eps_mean && eps_std, "null input pointer");
  G2G_REQUIRE(trans && start, "trans/start are required host pointers");
  G2G_REQUIRE(musigma && trends && nnz && flags, "null output pointer");
  G2G_REQUIRE(align_str && align_len && opt_cost && l_states, "null output pointer");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS, "n_bins out of range [2, 128]");
  G2G_REQUIRE(n_genes >= 0 && g_stride >= n_genes, "bad gene counts");
  if (n_genes == 0) return G2G_OK;
  FusedParams q;
  q.red.n_genes = n_genes; q.red.g_stride = g_stride; q.red.n_bins = n_bins;
  for (int s = 0; s < 2; ++s) {
    G2G_REQUIRE(sides[s].part_f64 && sides[s].part_i32 && sides[s].pilot && sides[s].n_cells >= 1, "bad side");
    g2g_layout_t lay;
    int rc = g2g_layout(sides[s].n_cells, n_bins, &lay);
    if (rc != G2G_OK) return rc;
    q.red.part_f64[s] = sides[s].part_f64;
    q.red.part_i32[s] = sides[s].part_i32;
    q.red.red_f64[s] = nullptr;
    q.red.red_i32[s] = nullptr;
    q.red.n_split[s] = sides[s].n_split > 0 ? sides[s].n_split : lay.n_split;
    q.fix.side[s].pilot = sides[s].pilot;
    q.fix.side[s].sum_w = s == 0 ? sum_w_ref : sum_w_query;
    q.fix.side[s].n_cells = sides[s].n_cells;
  }
  q.fix.points = points; q.fix.eps_mean = eps_mean; q.fix.eps_std = eps_std;
  q.fix.musigma = musigma; q.fix.trends = trends; q.fix.nnz = nnz; q.fix.flags = flags;
  q.fix.n_genes = n_genes; q.fix.n_bins = n_bins;
  DpParams& p = q.dp;
  p.trends = trends; p.cost_in = nullptr; p.align_str = align_str; p.align_len = align_len;
  p.opt_cost = opt_cost; p.l_states = l_states; p.cost_out = nullptr; p.mats_out = nullptr;
  p.bp_out = nullptr; p.l_out = nullptr; p.n_genes = n_genes; p.n_bins = n_bins;
  for (int k = 0; k < 25; ++k) p.trans[k] = trans[k];
  for (int k = 0; k < 3; ++k) p.start[k] = start[k];
  p.half_n = n_samples / 2.0;
  p.inv_4n = 1.0 / (4.0 * n_samples);
  p.two_c = 2.0 * c_model;
  cudaStream_t st = (cudaStream_t)stream;
  switch ((n_bins + 1 + 31) / 32) {
    case 1: return launch_fused<1>(q, st);
    case 2: return launch_fused<2>(q, st);
    case 3: return launch_fused<3>(q, st);
    case 4: return launch_fused<4>(q, st);
    default: return launch_fused<5>(q, st);
  }
}
