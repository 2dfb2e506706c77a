// K1  g2g_moments: single pass over the expression matrix (cells x genes, f32) of up to two
// datasets, producing per (gene, bin) the shifted weighted moments
//     A = sum_c w[t,c] x_c,   B = sum_c w[t,c] (x_c - a)^2,   C = sum_c (x_c - a)
// and the non-zero counts per pseudotime window.  Replaces the per-gene pandas/numpy loop of
// TimeSeriesPreprocessor.compute_stat (TSP.py:161-174) and the counts of Main.py:298-301,344-363.
//
// Structure (B200): persistent CTAs (two per SM while the register tile allows) walk a static list of
// work items (gene tile x cell range).  [64 cells x GT genes] tiles of X are streamed with TMA
// (cp.async.bulk.tensor.2d) and the matching 64 rows of the weight table with a 1-D bulk copy into a
// ring of shared-memory stages guarded by mbarriers; the warp that finishes a stage last (a wrapping
// shared-memory counter) re-arms it from a cursor kept decoded in shared memory, so there is no producer
// warp, no blocking hand-off and no per-tile index arithmetic.  Eight warps each take 8 cells of a stage;
// every thread owns GX genes x TB bins x {A,B} float accumulators in registers (an FFMA2 register tile:
// the weights are warp-broadcast LDS.128, X is a conflict-free LDS, two bins per packed FFMA2, and with
// GX = 2 the d / d^2 / C arithmetic of the two genes is packed too).  The kernel is bound by FP32
// instruction issue, so the loop is kept to FFMA2 plus the fewest other instructions: non-zero counting
// is one FMNMX per value, window bookkeeping happens per 8-cell group (K0 flags the groups that straddle
// a window boundary).  Every kFlushStages stages, and at the end of an item, the eight warps' f32
// partials are summed in f64 in a fixed order through a shared staging area, so results do not depend
// on scheduling; the item's f64 sums go to its own slot of `part_f64`.
#include "g2g_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kThreads = kConsumerWarps * 32;
constexpr int kCellTile = G2G_CELL_TILE;              // rows per stage
constexpr int kCellsPerWarp = kCellTile / kConsumerWarps;
static_assert(kCellsPerWarp == G2G_WIN_GROUP, "K0 flags mixed-window groups per warp share of a stage");
constexpr int kFlushStages = 16;                       // f32 run length per thread = 16 stages * 8 cells = 128 cells
constexpr int kMaxSides = 2;
enum { kCurItem, kCurSide, kCurTilesLeft, kCurGene0, kCurRow, kCurGroup, kCursorInts };

struct MomSide {
  const float* wtab;     // [n_groups][n_pad][RF]
  const float* pilot;    // [G]
  double* part_f64;      // [n_split][2T+1][g_stride]
  int32_t* part_i32;     // [n_split][T][g_stride]
  long long n_pad;
  long long cells_per_item;
  int n_split;
  int item_base;         // first item index of this side
};

struct MomParams {
  CUtensorMap tmap[kMaxSides];
  MomSide side[kMaxSides];
  long long n_genes, g_stride;
  int n_sides, n_bins, n_groups, bins_per_group, n_gtiles, n_items, n_stages;
};

// ---- raw PTX helpers (mbarrier + TMA) ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Wrapping counter: returns the old value and stores old >= limit ? 0 : old + 1.  The arrival that sees
// `limit` has also reset the counter, and ptxas emits a bare ATOMS.INC for it (an ATOMS.ADD gets wrapped
// in warp-aggregation code even when a single lane runs it).
// acq_rel at CTA scope: the arrival that completes a stage (and goes on to overwrite it) is ordered after the
// other warps' shared-memory reads of that stage, which their own arrivals released.
__device__ __forceinline__ unsigned smem_atomic_inc_wrap(int* addr, unsigned limit) {
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared.inc.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(addr)), "r"(limit) : "memory");
  return old;
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// CTAs per SM the register / shared-memory budget is sized for (2 measured faster than 1 on B200,
// profiles/r01_k1_item_size_sweep.txt)
constexpr int kMomMinBlocks = 2;

template <int TBE, int GX>
struct Cfg {
  static constexpr int NP = TBE / 2;                                  // packed bin pairs per thread
  static constexpr int GT = 32 * GX;                                  // genes per CTA tile
  static constexpr int RF = (TBE + 2 + 3) / 4 * 4;                    // floats per weight-table row
  static constexpr int NQ = 2 * TBE + 1;                              // accumulator rows (A bins, B bins, C)
  static constexpr int XBYTES = kCellTile * GT * 4;
  static constexpr int WBYTES = kCellTile * RF * 4;
  static constexpr int STAGE_BYTES = XBYTES + WBYTES;
  // the cross-warp reduction goes through a shared staging area, in kPasses slices of kRows rows
  // two CTAs per SM while the register tile leaves room (<= 128 registers without spills)
  static constexpr int kMinBlocks = (kMomMinBlocks >= 2 && 2 * TBE * GX <= 64) ? 2 : 1;
  static constexpr int kStagingCap = (kMinBlocks >= 2 ? 16 : 64) * 1024;
  static constexpr int kPasses = (kConsumerWarps * NQ * GT * 4 + kStagingCap - 1) / kStagingCap;
  static constexpr int kRows = (NQ + kPasses - 1) / kPasses;
  static constexpr int STAGING_BYTES = kConsumerWarps * kRows * GT * 4;
};

struct Item {
  int side, bg, split, gt, n_tiles;
  long long row0;
};

__device__ __forceinline__ Item decode_item(const MomParams& p, int item) {
  Item it;
  it.side = (p.n_sides > 1 && item >= p.side[1].item_base) ? 1 : 0;
  const MomSide& s = p.side[it.side];
  int r = item - s.item_base;
  it.gt = r % p.n_gtiles;
  r /= p.n_gtiles;
  it.split = r % s.n_split;
  it.bg = r / s.n_split;
  it.row0 = (long long)it.split * s.cells_per_item;
  long long rows = s.n_pad - it.row0;
  if (rows > s.cells_per_item) rows = s.cells_per_item;
  it.n_tiles = (int)(rows / kCellTile);
  return it;
}

// acc{lo,hi} += w2{lo,hi} * x   -- one packed FFMA2 (two bins of one gene); the scalar x is a
// broadcast operand in SASS (FFMA2 Rd, Ra.F32x2.HI_LO, Rb.F32, Rc.F32x2.HI_LO)
__device__ __forceinline__ void ffma2_bcast(unsigned long long& acc, unsigned long long w2, float x) {
  asm("{\n.reg .b64 xx;\nmov.b64 xx, {%2, %2};\nfma.rn.f32x2 %0, %1, xx, %0;\n}" : "+l"(acc) : "l"(w2), "f"(x));
}
// 1 where x != 0 (NaN counts as non-zero, like np.count_nonzero), else 0, in one FMNMX: the bit
// pattern of min(|x|, smallest denormal) is the integer 0 or 1 (f32 denormals are not flushed here)
#ifdef __CUDA_FTZ
#error "g2g_moments.cu must be compiled without -ftz=true / --use_fast_math: nonzero_flag() relies on f32 denormals"
#endif
__device__ __forceinline__ int nonzero_flag(float x) { return __float_as_int(fminf(fabsf(x), __int_as_float(1))); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float lo_f(unsigned long long v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi_f(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }

template <int TBE, int GX>
__global__ void __launch_bounds__(kThreads, Cfg<TBE, GX>::kMinBlocks) moments_kernel(const __grid_constant__ MomParams p) {
  using C = Cfg<TBE, GX>;
  constexpr int NP = C::NP;
  // small register tiles: all 8 cells of a stage straight-line; the 128-register ones keep 2 (no spills)
  constexpr int kCellUnroll = (2 * TBE * GX <= 56) ? kCellsPerWarp : (2 * TBE * GX <= 64) ? 2 : 1;
  extern __shared__ __align__(128) unsigned char smem[];
  const int n_stages = p.n_stages;
  // layout: [stages][X tile | W slice] | staging | acc64 [NQ][GT] | cnt [n_bins][GT] | barriers, counters
  unsigned char* stage_base = smem;
  float* staging = reinterpret_cast<float*>(smem + (size_t)n_stages * C::STAGE_BYTES);
  double* acc64 = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(staging) + C::STAGING_BYTES);
  int* cnt_s = reinterpret_cast<int*>(acc64 + C::NQ * C::GT);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(cnt_s + (size_t)p.n_bins * C::GT);
  int* done = reinterpret_cast<int*>(full_bar + n_stages);  // warps finished with each stage
  int* cursor = done + n_stages;                             // decoded position of the next tile to load

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tid = threadIdx.x;

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      done[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int k = tid; k < p.n_bins * C::GT; k += kThreads) cnt_s[k] = 0;
  // (no barrier here: thread 0 goes straight on to issue the first loads, the block meets below)

  // Load the next tile of this CTA's item sequence into `slot`.  Only one thread runs this at a
  // time: the prologue, then whichever warp finished the previous occupant of the slot last.  The
  // cursor in shared memory holds the decoded position so that the common case is a few adds.
  auto cursor_set_item = [&](int item) {
    cursor[kCurItem] = item;
    if (item >= p.n_items) return;
    const Item it = decode_item(p, item);
    cursor[kCurSide] = it.side;
    cursor[kCurTilesLeft] = it.n_tiles;
    cursor[kCurGene0] = it.gt * C::GT;
    cursor[kCurRow] = (int)it.row0;
    cursor[kCurGroup] = it.bg;
  };
  auto issue_next = [&](int slot_i) {
    if (cursor[kCurItem] >= p.n_items) return;
    const int side = cursor[kCurSide], row = cursor[kCurRow];
    const MomSide& s = p.side[side];
    unsigned char* dst = stage_base + (size_t)slot_i * C::STAGE_BYTES;
    mbar_arrive_expect_tx(&full_bar[slot_i], C::STAGE_BYTES);
    tma_load_2d(dst, &p.tmap[side], cursor[kCurGene0], row, &full_bar[slot_i]);
    bulk_load_1d(dst + C::XBYTES, s.wtab + ((size_t)cursor[kCurGroup] * s.n_pad + (size_t)row) * C::RF, C::WBYTES,
                 &full_bar[slot_i]);
    const int left = cursor[kCurTilesLeft] - 1;
    if (left == 0) {
      cursor_set_item(cursor[kCurItem] + gridDim.x);
    } else {
      cursor[kCurTilesLeft] = left;
      cursor[kCurRow] = row + kCellTile;
    }
  };
  if (tid == 0) {
    cursor_set_item(blockIdx.x);
    for (int s = 0; s < n_stages; ++s) issue_next(s);
    __threadfence_block();
  }
  __syncthreads();

  // ---- register tile: GX genes x TBE bins x {A, B} as packed f32 pairs, plus C per gene ----------
  unsigned long long accA[GX][NP], accB[GX][NP];
  float accC[GX];
  unsigned long long accC2 = 0ull;   // GX == 2: C of both genes as one packed accumulator
  int cnt[GX];
#pragma unroll
  for (int x = 0; x < GX; ++x) {
    accC[x] = 0.0f;
    cnt[x] = 0;
#pragma unroll
    for (int k = 0; k < NP; ++k) { accA[x][k] = 0ull; accB[x][k] = 0ull; }
  }
  int slot = 0;
  uint32_t phase = 0;

  // f32 -> f64 reduction across the eight warps (all threads of the CTA take part).  Every warp
  // stores its partial sums to the staging area; thread k then adds the eight values of slot k in
  // warp order to the CTA's f64 accumulator -- a fixed summation order.  `final` also adds the f64
  // accumulator, writes the item's result to global memory and clears it.
  bool acc64_used = false;   // a non-final flush of the current item has left sums in acc64
  auto reduce_flush = [&](bool final, const Item& it) {
    const MomSide& s = p.side[it.side];
    const int T = p.n_bins, TB = p.bins_per_group;
    const long long g0 = (long long)it.gt * C::GT;
    double* pf = s.part_f64 + (size_t)it.split * (2 * T + 1) * p.g_stride;
#pragma unroll
    for (int pass = 0; pass < C::kPasses; ++pass) {
      const int r0 = pass * C::kRows;
      float* mine = staging + (size_t)warp * C::kRows * C::GT + lane * GX;
#pragma unroll
      for (int k = 0; k < NP; ++k) {
#pragma unroll
        for (int x = 0; x < GX; ++x) {
          const int qa = 2 * k, qb = TBE + 2 * k;
          if (qa >= r0 && qa < r0 + C::kRows) mine[(qa - r0) * C::GT + x] = lo_f(accA[x][k]);
          if (qa + 1 >= r0 && qa + 1 < r0 + C::kRows) mine[(qa + 1 - r0) * C::GT + x] = hi_f(accA[x][k]);
          if (qb >= r0 && qb < r0 + C::kRows) mine[(qb - r0) * C::GT + x] = lo_f(accB[x][k]);
          if (qb + 1 >= r0 && qb + 1 < r0 + C::kRows) mine[(qb + 1 - r0) * C::GT + x] = hi_f(accB[x][k]);
        }
      }
      if (2 * TBE >= r0 && 2 * TBE < r0 + C::kRows) {
#pragma unroll
        for (int x = 0; x < GX; ++x)
          mine[(2 * TBE - r0) * C::GT + x] = GX == 2 ? (x == 0 ? lo_f(accC2) : hi_f(accC2)) : accC[x];
      }
      __syncthreads();
      const int n_slots = ((C::NQ - r0 < C::kRows) ? (C::NQ - r0) : C::kRows) * C::GT;
      for (int k = tid; k < n_slots; k += kThreads) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) sum += (double)staging[(size_t)w * C::kRows * C::GT + k];
        const int slot64 = r0 * C::GT + k;
        if (acc64_used) sum += acc64[slot64];
        if (final) {
          const int q = slot64 / C::GT, gl = slot64 % C::GT;
          const long long g = g0 + gl;
          int row = -1;
          if (q < TBE) { int t = it.bg * TB + q; if (q < TB && t < T) row = t; }
          else if (q < 2 * TBE) { int t = it.bg * TB + (q - TBE); if (q - TBE < TB && t < T) row = T + t; }
          else if (it.bg == 0) row = 2 * T;
          if (row >= 0 && g < p.n_genes) pf[(size_t)row * p.g_stride + g] = sum;
        } else {
          acc64[slot64] = sum;
        }
      }
      __syncthreads();
    }
    acc64_used = !final;
    accC2 = 0ull;
#pragma unroll
    for (int x = 0; x < GX; ++x) {
      accC[x] = 0.0f;
#pragma unroll
      for (int k = 0; k < NP; ++k) { accA[x][k] = 0ull; accB[x][k] = 0ull; }
    }
  };

  // pilot (shift) values of the item after the current one are fetched one item ahead
  auto load_pilot = [&](int item_i, float* dst) {
    if (item_i < p.n_items) {
      const Item nx = decode_item(p, item_i);
#pragma unroll
      for (int x = 0; x < GX; ++x) {
        long long g = (long long)nx.gt * C::GT + lane * GX + x;
        dst[x] = g < p.n_genes ? p.side[nx.side].pilot[g] : 0.0f;
      }
    }
  };
  float a_next[GX];
#pragma unroll
  for (int x = 0; x < GX; ++x) a_next[x] = 0.0f;
  load_pilot(blockIdx.x, a_next);

  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    const Item it = decode_item(p, item);
    const MomSide& s = p.side[it.side];
    const bool do_counts = (it.bg == 0);
    float a[GX];
#pragma unroll
    for (int x = 0; x < GX; ++x) a[x] = a_next[x];
    load_pilot(item + gridDim.x, a_next);
    const unsigned long long neg_a2 = pack2(-a[0], -a[GX - 1]);
    int cur_win = -1;     // window the running counts `cnt` belong to
    int since_flush = 0;
    auto flush_counts = [&]() {
      if (do_counts && cur_win >= 0) {
#pragma unroll
        for (int x = 0; x < GX; ++x)
          if (cnt[x]) atomicAdd(&cnt_s[cur_win * C::GT + lane * GX + x], cnt[x]);
      }
#pragma unroll
      for (int x = 0; x < GX; ++x) cnt[x] = 0;
    };
    for (int tile = 0; tile < it.n_tiles; ++tile) {
      mbar_wait(&full_bar[slot], phase);
      const float* xt = reinterpret_cast<const float*>(stage_base + (size_t)slot * C::STAGE_BYTES);
      const float* wt = xt + kCellTile * C::GT;
      const int cell0 = warp * kCellsPerWarp;
      // counts: K0 marks the 8-cell groups whose cells do not all lie in one window (with
      // time-sorted input only the few groups on a window boundary); everywhere else the loop
      // below counts straight into `cnt` for that window
      const int win0 = __float_as_int(wt[cell0 * C::RF + TBE + 1]);
      const bool mixed = win0 >= G2G_WIN_MIXED / 2;  // warp-uniform
      if (do_counts && !mixed && win0 != cur_win) { flush_counts(); cur_win = win0; }
#pragma unroll(kCellUnroll)
      for (int cc = 0; cc < kCellsPerWarp; ++cc) {
        const int cell = cell0 + cc;
        const ulonglong2* wrow = reinterpret_cast<const ulonglong2*>(wt + cell * C::RF);
        unsigned long long w2[C::RF / 2];
#pragma unroll
        for (int q = 0; q < C::RF / 4; ++q) {
          ulonglong2 v = wrow[q];
          w2[2 * q] = v.x;
          w2[2 * q + 1] = v.y;
        }
        const float valid = lo_f(w2[NP]);
        if constexpr (GX == 2) {
          // the two genes of a lane side by side: d, y and C as packed f32x2 operations
          const unsigned long long x2 = *reinterpret_cast<const unsigned long long*>(xt + cell * C::GT + lane * 2);
          const unsigned long long d2 = add2(x2, neg_a2);
          const unsigned long long y2 = mul2(d2, d2);
          const float xs[2] = {lo_f(x2), hi_f(x2)}, ys[2] = {lo_f(y2), hi_f(y2)};
#pragma unroll
          for (int x = 0; x < 2; ++x) {
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              ffma2_bcast(accA[x][k], w2[k], xs[x]);
              ffma2_bcast(accB[x][k], w2[k], ys[x]);
            }
            cnt[x] += nonzero_flag(xs[x]);
          }
          ffma2_bcast(accC2, d2, valid);
        } else {
          const float xv = xt[cell * C::GT + lane];
          const float d = xv - a[0];
          const float y = d * d;
#pragma unroll
          for (int k = 0; k < NP; ++k) {
            ffma2_bcast(accA[0][k], w2[k], xv);
            ffma2_bcast(accB[0][k], w2[k], y);
          }
          accC[0] = fmaf(valid, d, accC[0]);
          cnt[0] += nonzero_flag(xv);
        }
      }
      if (do_counts && mixed) {  // rare: take the loop's counts back, then count cell by cell
        for (int cc = 0; cc < kCellsPerWarp; ++cc) {
#pragma unroll
          for (int x = 0; x < GX; ++x) cnt[x] -= nonzero_flag(xt[(cell0 + cc) * C::GT + lane * GX + x]);
        }
        for (int cc = 0; cc < kCellsPerWarp; ++cc) {
          const int cell = cell0 + cc;
          int win = __float_as_int(wt[cell * C::RF + TBE + 1]);
          if (win >= G2G_WIN_MIXED / 2) win -= G2G_WIN_MIXED;
          if (win != cur_win) { flush_counts(); cur_win = win; }
#pragma unroll
          for (int x = 0; x < GX; ++x) cnt[x] += nonzero_flag(xt[cell * C::GT + lane * GX + x]);
        }
      }
      __syncwarp();  // every lane's shared-memory reads of this stage have returned (values are in registers)
      if (lane == 0) {
        // last warp out refills the stage (its arrival also wraps the counter back to 0); the proxy fence orders
        // the generic-proxy reads of the stage before the async-proxy (TMA / bulk copy) writes that reuse it
        if (smem_atomic_inc_wrap(&done[slot], kConsumerWarps - 1) == kConsumerWarps - 1) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue_next(slot);
        }
      }
      if (++slot == n_stages) { slot = 0; phase ^= 1; }
      if (++since_flush == kFlushStages && tile + 1 < it.n_tiles) { reduce_flush(false, it); since_flush = 0; }
    }
    flush_counts();
    reduce_flush(true, it);   // its barriers also order the count atomics before the write-out below
    if (do_counts) {
      const int T = p.n_bins;
      const long long g0 = (long long)it.gt * C::GT;
      int32_t* pi = s.part_i32 + (size_t)it.split * T * p.g_stride;
      for (int k = tid; k < T * C::GT; k += kThreads) {
        const int q = k / C::GT, gl = k % C::GT;
        const long long g = g0 + gl;
        if (g < p.n_genes) pi[(size_t)q * p.g_stride + g] = cnt_s[k];
        cnt_s[k] = 0;
      }
      __syncthreads();
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    G2G_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) {
      g2g_set_error("cuTensorMapEncodeTiled not available from the driver");
      return G2G_ERR_CUDA;
    }
    cached = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *out = cached;
  return G2G_OK;
}

template <int TBE, int GX>
int launch_moments(MomParams& p, const g2g_side_t* sides, cudaStream_t stream) {
  using C = Cfg<TBE, GX>;
  EncodeTiledFn encode;
  int rc = get_encode_fn(&encode);
  if (rc != G2G_OK) return rc;
  for (int s = 0; s < p.n_sides; ++s) {
    cuuint64_t gdim[2] = {(cuuint64_t)p.n_genes, (cuuint64_t)sides[s].n_cells};
    cuuint64_t gstride[1] = {(cuuint64_t)sides[s].ldx * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)C::GT, (cuuint32_t)kCellTile};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.tmap[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(sides[s].X), gdim,
                        gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      g2g_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
      return G2G_ERR_CUDA;
    }
  }
  p.n_gtiles = (int)g2g_ceil_div(p.n_genes, C::GT);
  int items = 0;
  for (int s = 0; s < p.n_sides; ++s) {
    p.side[s].item_base = items;
    items += p.n_groups * p.side[s].n_split * p.n_gtiles;
  }
  p.n_items = items;

  int dev = 0, n_sm = 0;
  G2G_CUDA_TRY(cudaGetDevice(&dev));
  G2G_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const size_t fixed = (size_t)C::STAGING_BYTES + (size_t)C::NQ * C::GT * 8 + (size_t)p.n_bins * C::GT * 4 +
                       8 * 8 + (8 + kCursorInts) * 4 + 128;
  // as many stages (<= 8, >= 2) as fit the shared memory of kMomMinBlocks CTAs per SM
  int stages = 8;
  const size_t budget = (size_t)(C::kMinBlocks >= 2 ? 110 : 200) * 1024;
  while (stages > 2 && fixed + (size_t)stages * C::STAGE_BYTES > budget) --stages;
  p.n_stages = stages;
  const size_t smem = fixed + (size_t)stages * C::STAGE_BYTES;
  G2G_CUDA_TRY(cudaFuncSetAttribute(moments_kernel<TBE, GX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  G2G_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, moments_kernel<TBE, GX>, kThreads, smem));
  if (occ < 1) {
    g2g_set_error("moments kernel does not fit on an SM (smem %zu bytes)", smem);
    return G2G_ERR_UNSUPPORTED;
  }
  int grid = n_sm * occ;
  if (grid > items) grid = items;
  moments_kernel<TBE, GX><<<grid, kThreads, smem, stream>>>(p);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}

#define G2G_CASE2(TBV) case TBV: return launch_moments<TBV, 2>(p, sides, stream);
#define G2G_CASE1(TBV) case TBV: return launch_moments<TBV, 1>(p, sides, stream);

int dispatch_moments(MomParams& p, const g2g_side_t* sides, int bins_even, cudaStream_t stream) {
  switch (bins_even) {
    G2G_CASE2(2) G2G_CASE2(4) G2G_CASE2(6) G2G_CASE2(8) G2G_CASE2(10) G2G_CASE2(12) G2G_CASE2(14)
    G2G_CASE2(16) G2G_CASE2(18) G2G_CASE2(20) G2G_CASE2(22) G2G_CASE2(24) G2G_CASE2(26) G2G_CASE2(28)
    G2G_CASE1(30) G2G_CASE1(32) G2G_CASE1(34) G2G_CASE1(36) G2G_CASE1(38) G2G_CASE1(40) G2G_CASE1(42)
    G2G_CASE1(44) G2G_CASE1(46) G2G_CASE1(48) G2G_CASE1(50) G2G_CASE1(52) G2G_CASE1(54) G2G_CASE1(56)
    default:
      g2g_set_error("bins_even %d has no compiled kernel", bins_even);
      return G2G_ERR_UNSUPPORTED;
  }
}

}  // namespace

extern "C" int g2g_moments(const g2g_side_t* sides, int32_t n_sides, int64_t n_genes, int64_t g_stride,
                           int32_t n_bins, void* stream) {
  G2G_REQUIRE(sides != nullptr && n_sides >= 1 && n_sides <= kMaxSides, "n_sides must be 1 or 2");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS, "n_bins out of range [2, 128]");
  G2G_REQUIRE(n_genes >= 1 && g_stride >= n_genes, "bad gene counts");
  G2G_REQUIRE(n_genes < (1ll << 31), "n_genes too large");
  G2GTiling t = g2g_tiling(n_bins);
  MomParams p;
  p.n_genes = n_genes; p.g_stride = g_stride; p.n_sides = n_sides; p.n_bins = n_bins; p.n_groups = t.n_groups;
  p.bins_per_group = t.bins_per_group;
  for (int s = 0; s < n_sides; ++s) {
    const g2g_side_t& in = sides[s];
    G2G_REQUIRE(in.X && in.wtab && in.pilot && in.part_f64 && in.part_i32, "null pointer in side");
    G2G_REQUIRE(in.n_cells >= 1 && in.n_cells < (1ll << 31), "n_cells out of range");
    G2G_REQUIRE(in.ldx >= n_genes && in.ldx % 4 == 0, "ldx must be >= n_genes and a multiple of 4 (16-byte rows for TMA)");
    G2G_REQUIRE(((uintptr_t)in.X & 15) == 0, "X must be 16-byte aligned");
    G2G_REQUIRE(((uintptr_t)in.wtab & 15) == 0, "wtab must be 16-byte aligned");
    g2g_layout_t lay;
    int rc = g2g_layout(in.n_cells, n_bins, &lay);
    if (rc != G2G_OK) return rc;
    p.side[s].wtab = in.wtab; p.side[s].pilot = in.pilot;
    p.side[s].part_f64 = in.part_f64; p.side[s].part_i32 = in.part_i32;
    p.side[s].n_pad = lay.n_pad; p.side[s].cells_per_item = lay.cells_per_item;
    p.side[s].n_split = (int)lay.n_split;
  }
  return dispatch_moments(p, sides, t.bins_even, (cudaStream_t)stream);
}
