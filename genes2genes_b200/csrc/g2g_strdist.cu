// Alignment-string distance matrices for clustering.
// Replaces ClusterUtils.compute_levenshtein_dist_matrix (ClusterUtils.py:21-29: Levenshtein.distance
// for every ordered pair of alignment strings) and ClusterUtils.compute_hamming_dist_matrix
// (:31-38 with binary_encode_alignment_path :361-377).  Integer distances are exact; the host divides by
// max(len) / the encoding length like the reference.
//
// Levenshtein: Myers' bit-vector algorithm (block form, up to 4 x 64-bit words = 256 characters),
// one thread per unordered pair; the per-string match masks are built once.
#include "g2g_common.cuh"

namespace {

constexpr int kMaxWords = 4;   // strings up to 256 characters (2 * G2G_MAX_BINS)
constexpr int kAlphabet = 5;   // M, W, V, D, I

__device__ __forceinline__ int letter_code(uint8_t ch) {
  return ch == 'M' ? 0 : ch == 'W' ? 1 : ch == 'V' ? 2 : ch == 'D' ? 3 : 4;
}

// peq[g][c][w]: bit i of word w set when character 64*w + i of string g is letter c
__global__ void __launch_bounds__(128) build_peq_kernel(const uint8_t* __restrict__ strs, const int* __restrict__ lens,
                                                        long long n, int stride, int words,
                                                        unsigned long long* __restrict__ peq) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  unsigned long long m[kAlphabet][kMaxWords];
#pragma unroll
  for (int c = 0; c < kAlphabet; ++c)
#pragma unroll
    for (int w = 0; w < kMaxWords; ++w) m[c][w] = 0ull;
  const uint8_t* s = strs + (size_t)g * stride;
  const int len = lens[g];
  for (int i = 0; i < len; ++i) {
    const int c = letter_code(s[i]);
#pragma unroll
    for (int cc = 0; cc < kAlphabet; ++cc)
#pragma unroll
      for (int w = 0; w < kMaxWords; ++w)
        if (cc == c && w == (i >> 6)) m[cc][w] |= 1ull << (i & 63);
  }
  for (int c = 0; c < kAlphabet; ++c)
    for (int w = 0; w < words; ++w) peq[((size_t)g * kAlphabet + c) * words + w] = m[c][w];
}

template <int W>
__global__ void __launch_bounds__(256) levenshtein_kernel(const uint8_t* __restrict__ strs, const int* __restrict__ lens,
                                                          long long n, int stride,
                                                          const unsigned long long* __restrict__ peq,
                                                          uint16_t* __restrict__ out) {
  const long long i = (long long)blockIdx.y * 16 + (threadIdx.x >> 4);   // pattern
  const long long j = (long long)blockIdx.x * 16 + (threadIdx.x & 15);   // text
  if (i >= n || j >= n || j < i) return;
  const int m = lens[i], nt = lens[j];
  int score = m;
  if (m == 0) {
    score = nt;
  } else {
    unsigned long long Peq[kAlphabet][W], Pv[W], Mv[W];
#pragma unroll
    for (int c = 0; c < kAlphabet; ++c)
#pragma unroll
      for (int w = 0; w < W; ++w) Peq[c][w] = peq[((size_t)i * kAlphabet + c) * W + w];
#pragma unroll
    for (int w = 0; w < W; ++w) { Pv[w] = ~0ull; Mv[w] = 0ull; }
    const int last = (m - 1) >> 6;                       // block holding the last pattern character
    const unsigned long long top_last = 1ull << ((m - 1) & 63);
    const uint8_t* t = strs + (size_t)j * stride;
    for (int k = 0; k < nt; ++k) {
      const int c = letter_code(t[k]);
      int hin = 1;                                       // row 0 of the edit matrix grows by one per text char
#pragma unroll
      for (int w = 0; w < W; ++w) {
        if (w > last) break;
        unsigned long long Eq = 0ull;
#pragma unroll
        for (int cc = 0; cc < kAlphabet; ++cc) Eq = (cc == c) ? Peq[cc][w] : Eq;
        const unsigned long long pv = Pv[w], mv = Mv[w];
        const unsigned long long Xv = Eq | mv;
        if (hin < 0) Eq |= 1ull;
        const unsigned long long Xh = (((Eq & pv) + pv) ^ pv) | Eq;
        unsigned long long Ph = mv | ~(Xh | pv);
        unsigned long long Mh = pv & Xh;
        const unsigned long long top = (w == last) ? top_last : (1ull << 63);
        const int hout = (Ph & top) ? 1 : ((Mh & top) ? -1 : 0);
        Ph <<= 1;
        Mh <<= 1;
        if (hin < 0) Mh |= 1ull; else if (hin > 0) Ph |= 1ull;
        Pv[w] = Mh | ~(Xv | Ph);
        Mv[w] = Ph & Xv;
        hin = hout;
      }
      score += hin;                                      // vertical delta leaving the last block
    }
  }
  out[(size_t)i * n + j] = (uint16_t)score;
  out[(size_t)j * n + i] = (uint16_t)score;
}

// binary path encoding (ClusterUtils.py:361-377): reference-side bits (M,W -> 1; D -> 0) followed by
// query-side bits (M,V -> 1; I -> 0), n_bins bits each
__global__ void __launch_bounds__(128) binary_encode_kernel(const uint8_t* __restrict__ strs,
                                                            const int* __restrict__ lens, long long n, int stride,
                                                            int n_bins, unsigned long long* __restrict__ enc) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  unsigned long long e[kMaxWords] = {0ull, 0ull, 0ull, 0ull};
  const uint8_t* s = strs + (size_t)g * stride;
  int is = 0, it = n_bins;
  for (int k = 0; k < lens[g]; ++k) {
    const uint8_t ch = s[k];
    const bool on_s = (ch == 'M' || ch == 'W' || ch == 'D'), on_t = (ch == 'M' || ch == 'V' || ch == 'I');
    if (on_s && is < n_bins) {
      if (ch != 'D') {
#pragma unroll
        for (int w = 0; w < kMaxWords; ++w) if (w == (is >> 6)) e[w] |= 1ull << (is & 63);
      }
      ++is;
    }
    if (on_t && it < 2 * n_bins) {
      if (ch != 'I') {
#pragma unroll
        for (int w = 0; w < kMaxWords; ++w) if (w == (it >> 6)) e[w] |= 1ull << (it & 63);
      }
      ++it;
    }
  }
#pragma unroll
  for (int w = 0; w < kMaxWords; ++w) enc[(size_t)g * kMaxWords + w] = e[w];
}

__global__ void __launch_bounds__(256) hamming_kernel(const unsigned long long* __restrict__ enc, long long n,
                                                      uint16_t* __restrict__ out) {
  const long long i = (long long)blockIdx.y * 16 + (threadIdx.x >> 4);
  const long long j = (long long)blockIdx.x * 16 + (threadIdx.x & 15);
  if (i >= n || j >= n) return;
  int d = 0;
#pragma unroll
  for (int w = 0; w < kMaxWords; ++w) d += __popcll(enc[(size_t)i * kMaxWords + w] ^ enc[(size_t)j * kMaxWords + w]);
  out[(size_t)i * n + j] = (uint16_t)d;
}

}  // namespace

extern "C" size_t g2g_strdist_workspace_bytes(int64_t n_strings) {
  return (size_t)n_strings * kAlphabet * kMaxWords * sizeof(unsigned long long);
}

extern "C" int g2g_levenshtein_matrix(const uint8_t* strs, const int32_t* lens, int64_t n_strings, int32_t stride,
                                      uint16_t* dist, void* workspace, size_t workspace_bytes, void* stream) {
  G2G_REQUIRE(strs && lens && dist && workspace, "null pointer");
  G2G_REQUIRE(n_strings >= 1 && stride >= 1 && stride <= 64 * kMaxWords, "strings longer than 256 characters");
  G2G_REQUIRE(workspace_bytes >= g2g_strdist_workspace_bytes(n_strings), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int words = (stride + 63) / 64;
  unsigned long long* peq = reinterpret_cast<unsigned long long*>(workspace);
  build_peq_kernel<<<(unsigned)g2g_ceil_div(n_strings, 128), 128, 0, st>>>(strs, lens, n_strings, stride, words, peq);
  G2G_CUDA_TRY(cudaGetLastError());
  dim3 grid((unsigned)g2g_ceil_div(n_strings, 16), (unsigned)g2g_ceil_div(n_strings, 16));
  switch (words) {
    case 1: levenshtein_kernel<1><<<grid, 256, 0, st>>>(strs, lens, n_strings, stride, peq, dist); break;
    case 2: levenshtein_kernel<2><<<grid, 256, 0, st>>>(strs, lens, n_strings, stride, peq, dist); break;
    case 3: levenshtein_kernel<3><<<grid, 256, 0, st>>>(strs, lens, n_strings, stride, peq, dist); break;
    default: levenshtein_kernel<4><<<grid, 256, 0, st>>>(strs, lens, n_strings, stride, peq, dist); break;
  }
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}

extern "C" int g2g_hamming_matrix(const uint8_t* strs, const int32_t* lens, int64_t n_strings, int32_t stride,
                                  int32_t n_bins, uint16_t* dist, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  G2G_REQUIRE(strs && lens && dist && workspace, "null pointer");
  G2G_REQUIRE(n_strings >= 1 && stride >= 1 && n_bins >= 1 && n_bins <= G2G_MAX_BINS, "bad shape");
  G2G_REQUIRE(workspace_bytes >= (size_t)n_strings * kMaxWords * sizeof(unsigned long long), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* enc = reinterpret_cast<unsigned long long*>(workspace);
  binary_encode_kernel<<<(unsigned)g2g_ceil_div(n_strings, 128), 128, 0, st>>>(strs, lens, n_strings, stride, n_bins, enc);
  G2G_CUDA_TRY(cudaGetLastError());
  dim3 grid((unsigned)g2g_ceil_div(n_strings, 16), (unsigned)g2g_ceil_div(n_strings, 16));
  hamming_kernel<<<grid, 256, 0, st>>>(enc, n_strings, dist);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}
