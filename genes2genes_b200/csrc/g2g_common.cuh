// Shared helpers for the libg2g kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "g2g.h"

void g2g_set_error(const char* fmt, ...);

#define G2G_CUDA_TRY(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      g2g_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return G2G_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define G2G_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) {                                               \
      g2g_set_error("%s: %s", __func__, msg);                    \
      return G2G_ERR_ARG;                                        \
    }                                                            \
  } while (0)

// In-kernel bounds checks of a debug build (`make debug`: -DG2G_DEBUG, device-side assert); nothing in the
// shipped library.  compute-sanitizer is not available on the GPU pool, these are the substitute.
#ifdef G2G_DEBUG
#include <assert.h>
#define G2G_DEVICE_ASSERT(cond) assert(cond)
#else
#define G2G_DEVICE_ASSERT(cond) ((void)0)
#endif

static inline int64_t g2g_round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
static inline int64_t g2g_ceil_div(int64_t v, int64_t m) { return (v + m - 1) / m; }

// Window-id slot of a weight-table row: the window k (p_k <= tau < p_{k+1}), -1 for padding rows;
// K0 adds G2G_WIN_MIXED when the aligned group of G2G_WIN_GROUP cells the row belongs to (one warp's
// share of a K1 stage) spans more than one window.
#define G2G_WIN_GROUP 8
#define G2G_WIN_MIXED 0x10000

// Bin tiling shared by K0 (table layout) and K1 (register tile): see g2g_layout_t.
struct G2GTiling {
  int n_groups, bins_per_group, bins_even, row_floats, genes_per_thread;
};
static inline G2GTiling g2g_tiling(int n_bins) {
  G2GTiling t;
  t.n_groups = (int)g2g_ceil_div(n_bins, 56);
  t.bins_per_group = (int)g2g_ceil_div(n_bins, t.n_groups);
  t.bins_even = (t.bins_per_group + 1) & ~1;  // K1 packs bins in pairs (FFMA2); odd counts get a zero bin
  t.row_floats = (int)g2g_round_up(t.bins_even + 2, 4);
  t.genes_per_thread = t.bins_even <= 28 ? 2 : 1;
  return t;
}

// Tiling of the tensor-core interpolation kernel (g2g_moments_mma / g2g_weights_mma): bins in groups of
// <= 64, N = group size rounded up to the MMA granularity of 16.
struct G2GMmaTiling {
  int n_groups, bins_per_group, n_cols;
};
static inline G2GMmaTiling g2g_mma_tiling(int n_bins) {
  G2GMmaTiling t;
  t.n_groups = (int)g2g_ceil_div(n_bins, 64);
  t.bins_per_group = (int)g2g_ceil_div(n_bins, t.n_groups);
  t.n_cols = (int)g2g_round_up(t.bins_per_group, 16);
  return t;
}
// bytes of one 64-cell stage of the g2g_weights_mma table: 4 K steps x (hi + lo) x [n_cols x 16] f16 + 64 window words
__host__ __device__ static inline size_t g2g_mma_stage_bytes(int n_cols) { return (size_t)n_cols * 256 + 256; }
