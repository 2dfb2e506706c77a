// K4  g2g_cost_dp: MML match cost + five-state (M,W,V,D,I) DP + traceback + landscape.
//
// One warp per gene.  Lane l owns the B = ceil((T+1)/32) DP columns [l*B, l*B+B); at step s it
// fills row i = s - l + 1 of its columns (a skewed wavefront: lane l-1 finished row i one step
// earlier), receiving the five values of its left neighbour's last column by warp shuffle.
// All DP arithmetic is float64 adds/compares in the reference's order
// ((prev + cost) + transition, first minimum wins; OrgAlign.py:344-450), so given equal costs the
// five matrices, back-pointers, string and landscape are bit-identical to the reference.
#include "g2g_dp_body.cuh"

namespace {

template <int B, bool FULL, int LW>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) dp_kernel(const __grid_constant__ DpParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  dp_genes<B, FULL, LW>(p, smem_raw);
}

template <int B, bool FULL, int LW>
int launch_dp_mode(const DpParams& p, cudaStream_t stream) {
  constexpr int GPW = 32 / LW;
  size_t smem = dp_smem_per_warp(p.n_bins) * kWarpsPerBlock * GPW;
  if (smem > 48 * 1024) {
    G2G_CUDA_TRY(cudaFuncSetAttribute(dp_kernel<B, FULL, LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  unsigned blocks = (unsigned)g2g_ceil_div(p.n_genes, kWarpsPerBlock * GPW);
  dp_kernel<B, FULL, LW><<<blocks, kWarpsPerBlock * 32, smem, stream>>>(p);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}

template <int B>
int launch_dp(const DpParams& p, cudaStream_t stream) {
  const bool full = p.cost_in || p.cost_out || p.mats_out || p.bp_out || p.l_out;
  if (B == 1 && !full) {  // short series: several genes per warp
    if (p.n_bins + 1 <= 8) return launch_dp_mode<1, false, 8>(p, stream);
    if (p.n_bins + 1 <= 16) return launch_dp_mode<1, false, 16>(p, stream);
  }
  return full ? launch_dp_mode<B, true, 32>(p, stream) : launch_dp_mode<B, false, 32>(p, stream);
}

}  // namespace

extern "C" int g2g_cost_dp(const double* trends, int64_t n_genes, int32_t n_bins, const double* trans,
                           const double* start, double c_model, int32_t n_samples, const double* cost_in,
                           uint8_t* align_str, int32_t* align_len, double* opt_cost, uint8_t* l_states,
                           double* cost_out, double* mats_out, int8_t* bp_out, double* l_out, void* stream) {
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS, "n_bins out of range [2, 128]");
  G2G_REQUIRE(n_genes >= 0, "negative n_genes");
  G2G_REQUIRE(trans && start, "trans/start are required host pointers");
  G2G_REQUIRE(align_str && align_len && opt_cost && l_states, "null output pointer");
  G2G_REQUIRE(trends || cost_in, "need trends or cost_in");
  G2G_REQUIRE(trends != nullptr, "trends is required (its values are unused only when cost_in is given)");
  if (n_genes == 0) return G2G_OK;
  DpParams p;
  p.trends = trends; p.cost_in = cost_in; p.align_str = align_str; p.align_len = align_len;
  p.opt_cost = opt_cost; p.l_states = l_states; p.cost_out = cost_out; p.mats_out = mats_out;
  p.bp_out = bp_out; p.l_out = l_out; p.n_genes = n_genes; p.n_bins = n_bins;
  for (int k = 0; k < 25; ++k) p.trans[k] = trans[k];
  for (int k = 0; k < 3; ++k) p.start[k] = start[k];
  p.half_n = n_samples / 2.0;
  p.inv_4n = 1.0 / (4.0 * n_samples);
  p.two_c = 2.0 * c_model;
  cudaStream_t st = (cudaStream_t)stream;
  int nC = n_bins + 1;
  int B = (nC + 31) / 32;
  switch (B) {
    case 1: return launch_dp<1>(p, st);
    case 2: return launch_dp<2>(p, st);
    case 3: return launch_dp<3>(p, st);
    case 4: return launch_dp<4>(p, st);
    default: return launch_dp<5>(p, st);
  }
}
