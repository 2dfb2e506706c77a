// K1m  g2g_moments_mma: the interpolation pass of g2g_moments on the tensor cores (30 <= n_bins).
//
//     out[gene][bin] = sum_c  X[c][gene] * W[c][bin]          (TimeSeriesPreprocessor.py:161-174)
//
// is a [genes x cells] . [cells x bins] contraction.  Here it runs as tcgen05.mma kind::f16 with M = 128
// genes (one TMEM lane per gene), N = bins (+ the `valid` column) and K = 16 cells per instruction:
//
//   * X tiles [64 cells x 128 genes] f32 arrive by TMA, the matching B operand slice of the weight table
//     (f16 hi / lo, canonical K-major layout written by g2g_weights_mma) by a bulk copy, into a ring of
//     kStages shared-memory stages;
//   * converter warps (thread = gene) read their gene's 16 values of a K step, form y = (x - a)^2 and
//     split x and y into f16 pairs (hi, lo * 2048) -- 22 significant bits -- which they store to tensor
//     memory (tcgen05.st) as the A operands; they also count the non-zero values per pseudotime window
//     and add up x - a (the un-weighted sum needs no tensor core);
//   * one thread issues six MMAs per K step:  hhA += xh.wh, hhB += yh.wh, loA += xl.wh + xh.wl,
//     loB += yl.wh + yh.wl  (f32 accumulators in TMEM; the lo.lo product, 2^-22 of the result, is dropped);
//   * the tensor core truncates when it adds (profiles/r02_umma_ts_probe.txt: -1e-7 of the sum per step),
//     so the hh chains are cut every kFlushSteps = 8 steps (128 cells): flusher warps read the accumulator
//     (two buffers alternate) and add it, round to nearest, to f32 registers; at the end of a work item they
//     add the lo accumulator / 2048 in f64 and write the item's partial sums.
//
// The first moment is accumulated un-shifted on purpose: the truncation error is proportional to the sum
// of the magnitudes added, and only for sum w x is that the bin's own mean (with x - a it would be |a - mean|,
// which dwarfs the mean of a weakly expressed bin).
// Outputs as g2g_moments, except that there are g2g_mma_layout().n_split planes and that the window
// counts are added to part_i32 with integer atomics (the call zeroes it first).
#include "g2g_common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace {

constexpr int kGT = 128;               // genes per tile = TMEM lanes
constexpr int kStageCells = G2G_CELL_TILE;
constexpr int kStepCells = 16;         // K of one kind::f16 MMA
constexpr int kStepsPerStage = kStageCells / kStepCells;
constexpr int kFlushSteps = 8;         // K steps per first-level accumulation chain.  Measured on a C4 block, max error of the
                                       // means / stds vs the f64 oracle: 16 steps 0.585 ms / 2.9e-6, 8 steps 0.595 ms / 1.0e-6
                                       // (the f32 FFMA2 kernel: 1.1e-6), 4 steps 0.611 ms / 3.8e-7
constexpr int kStages = 4;
constexpr int kABufs = 4;              // A operand staging buffers in TMEM (32 columns each)
constexpr int kConvGroups = 4;         // converter groups of 4 warps (one warp per 32 TMEM lanes); the pipeline is
                                       // latency bound: 2 / 3 / 4 groups run a C4 block in 0.73 / 0.65 / 0.59 ms
constexpr int kFlushWarps = 8;         // 2 per lane quarter: A' half and B half of an accumulator buffer
// Warp roles, lowest to highest warp id = least to most latency critical (the SM's warp arbiter prefers
// the highest warp id among the eligible warps of a scheduler): flushers, converters, then one warp group
// with the TMEM allocator, the TMA producer and the three MMA issuers (one thread each, two MMAs per step:
// issuing the six MMAs of a step from one thread costs more than the tensor core needs to execute them).
constexpr int kFirstFlushWarp = 0, kFirstConvWarp = kFirstFlushWarp + kFlushWarps;
constexpr int kAllocWarp = kFirstConvWarp + 4 * kConvGroups, kProducerWarp = kAllocWarp + 1;
// three MMA issuers (the allocator warp doubles as one): second-moment lo chain, first-moment lo chain, hi.hi chains
constexpr int kMmaLoBWarp = kAllocWarp, kMmaLoAWarp = kAllocWarp + 2, kMmaWarp = kAllocWarp + 3;
constexpr int kIssuers = 3;
constexpr int kWarps = kAllocWarp + 4;
constexpr int kThreads = kWarps * 32;
constexpr int kXBytes = kStageCells * kGT * 4;
constexpr int kMaxSides = 2;
constexpr float kLoScale = 2048.0f;
#ifndef G2G_MMA_ABLATE
#define G2G_MMA_ABLATE 0   // profiling builds only: 1 no MMAs, 2 no conversion arithmetic, 4 no accumulator reads, 8 no tcgen05.st
#endif

struct MmaSide {
  const unsigned char* table;   // [n_groups][n_pad/64] stages
  const float* pilot;
  double* part_f64;
  int32_t* part_i32;
  long long n_pad, cells_per_item, n_cells;
  int n_split, item_base;
};

struct MmaParams {
  CUtensorMap tmap[kMaxSides];
  MmaSide side[kMaxSides];
  long long n_genes, g_stride;
  int32_t* status;
  int n_sides, n_bins, n_groups, bins_per_group, n_gtiles, n_items;
};

struct Item {
  int side, bg, split, gt, n_tiles;
  long long row0;
};

__device__ __forceinline__ Item decode_item(const MmaParams& p, int item) {
  Item it;
  it.side = (p.n_sides > 1 && item >= p.side[1].item_base) ? 1 : 0;
  const MmaSide& s = p.side[it.side];
  int r = item - s.item_base;
  // the bin group varies fastest: the (at most two) passes over an X tile of a T > 64 run are made by
  // neighbouring CTAs at the same time, so the second read of the tile is served by L2, not by HBM
  it.bg = r % p.n_groups;
  r /= p.n_groups;
  it.gt = r % p.n_gtiles;
  it.split = r / p.n_gtiles;
  it.row0 = (long long)it.split * s.cells_per_item;
  long long rows = s.n_pad - it.row0;
  if (rows > s.cells_per_item) rows = s.cells_per_item;
  it.n_tiles = (int)(rows / kStageCells);
  return it;
}

// ---- PTX helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// wait of a role that is not latency critical (flushers): the probe may suspend the thread for up to a few
// microseconds (suspend-time hint), so a waiting warp takes next to no issue slots from the converters
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP_R:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WAIT_DONE_R;\n"
      "bra WAIT_LOOP_R;\n"
      "WAIT_DONE_R:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(4000u)
      : "memory");
}
// Blocking hardware barrier among `threads` threads (id 1..15; 0 is __syncthreads): a waiting warp issues
// nothing, unlike a warp polling an mbarrier.  Roles whose warps all wait for the same mbarrier let one warp
// poll it and meet the others here.
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
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
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]   (M = 128, kind::f16, f32 accumulate)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// the mbarrier gets one arrival when every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle (8-row x 16-byte core matrices, rows 16 B apart);
// LBO = distance of the two 16-byte K chunks of a K = 16 step, SBO = distance of consecutive 8-row blocks
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: D = f32 (bit 4), A = B = f16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t instr_desc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// packed f32x2 arithmetic (two cells of one gene per instruction)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float lo_f(unsigned long long v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi_f(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }
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
// f16x2 of a packed f32 pair (low element in the low half)
__device__ __forceinline__ uint32_t cvt2_f16x2(unsigned long long v) {
  uint32_t r;
  asm("{\n.reg .f32 flo, fhi;\nmov.b64 {flo, fhi}, %1;\ncvt.rn.f16x2.f32 %0, fhi, flo;\n}" : "=r"(r) : "l"(v));
  return r;
}
// packed f32 pair (h.lo - v.lo, h.hi - v.hi) of an f16x2 h and a packed f32 pair v: two mixed-precision
// subtractions (SASS FHADD, f16 operand read from its half of the register) -- no conversion, negation or move
__device__ __forceinline__ unsigned long long half2_minus(uint32_t h, unsigned long long v) {
  unsigned long long r;
  asm("{\n.reg .b16 lo, hi;\n.reg .f32 vlo, vhi, dlo, dhi;\nmov.b32 {lo, hi}, %1;\nmov.b64 {vlo, vhi}, %2;\n"
      "sub.rn.f32.f16 dlo, lo, vlo;\nsub.rn.f32.f16 dhi, hi, vhi;\nmov.b64 %0, {dlo, dhi};\n}"
      : "=l"(r)
      : "r"(h), "l"(v));
  return r;
}
__device__ __forceinline__ unsigned long long pack2_asm(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
// hi / lo f16 split of a packed pair: hi = f16(v), lo = f16((v - hi) * 2048), formed as (hi - v) * -2048;
// (hi - v) is exact in f32
__device__ __forceinline__ void split2(unsigned long long v, uint32_t& hi, uint32_t& lo) {
  hi = cvt2_f16x2(v);
  lo = cvt2_f16x2(mul2(half2_minus(hi, v), pack2(-kLoScale, -kLoScale)));
}
#ifdef __CUDA_FTZ
#error "g2g_moments_mma.cu must be compiled without -ftz=true / --use_fast_math: nonzero_flag() relies on f32 denormals"
#endif
__device__ __forceinline__ int nonzero_flag(float x) { return __float_as_int(fminf(fabsf(x), __int_as_float(1))); }

// Register budget per role (setmaxnreg, whole warp groups of 4 warps): the kernel is launched with
// 65536 / kThreads registers per thread; the issuer / producer group and the converters give registers
// back, the flushers (N second-level accumulators per thread) take them.
constexpr int kRegsFlush = 104, kRegsConv = 64, kRegsMisc = 40;
static_assert((kFlushWarps * kRegsFlush + 4 * kConvGroups * kRegsConv + 4 * kRegsMisc) * 32 <= 65536 / kThreads / 8 * 8 * kThreads,
              "register pool");
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// Shared memory: [kStages][X tile 32 KB | B slice N*256 | 64 window words] | barriers
template <int N>
struct Smem {
  static constexpr int kBBytes = N * 256 + 256;            // = g2g_mma_stage_bytes(N)
  static constexpr int kStageBytes = kXBytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 2 * kConvGroups * kGT * 8;
};

template <int N>
__global__ void __launch_bounds__(kThreads, 1) moments_mma_kernel(const __grid_constant__ MmaParams p) {
  using S = Smem<N>;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* stage_full = bars;                       // [kStages]  TMA bytes landed
  uint64_t* stage_empty = stage_full + kStages;      // [kStages]  MMAs of the stage's 4 steps done
  uint64_t* a_full = stage_empty + kStages;          // [kABufs]   converters wrote the A operands (4 warps)
  uint64_t* a_empty = a_full + kABufs;               // [kABufs]   MMAs that read them done
  uint64_t* acc_full = a_empty + kABufs;             // [2]        a 128-cell chain is complete
  uint64_t* acc_empty = acc_full + 2;                // [2]        flushers have read it (kFlushWarps)
  uint64_t* lo_full = acc_empty + 2;                 // [1]        item complete
  uint64_t* lo_empty = lo_full + 1;                  // [1]
  uint64_t* done_bar = lo_empty + 1;                 // [1]        every MMA (and commit) of this CTA has retired
  uint64_t* c_full = done_bar + 1;                   // [2]        converters posted an item's sum (x - a)
  uint64_t* c_empty = c_full + 2;                    // [2]        flushers took it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_empty + 2);
  double* c_s = reinterpret_cast<double*>(smem + S::kBarOffset + 256);   // [2][kConvGroups][kGT]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&stage_full[s], 1); mbar_init(&stage_empty[s], kIssuers); }
    for (int b = 0; b < kABufs; ++b) { mbar_init(&a_full[b], 4); mbar_init(&a_empty[b], kIssuers); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kFlushWarps); }
    mbar_init(lo_full, 2);
    mbar_init(lo_empty, kFlushWarps);
    mbar_init(done_bar, kIssuers);
    for (int b = 0; b < 2; ++b) { mbar_init(&c_full[b], 4 * kConvGroups); mbar_init(&c_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kAllocWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  // TMEM columns: accumulator buffer b: [b*2N, b*2N + N) = A' (first moment), [.. + N, .. + 2N) = B (second);
  // lo accumulators at 4N (A') and 5N (B); A operand staging buffers of 32 columns from 6N
  constexpr uint32_t kLoCol = 4 * N, kACol = 6 * N;
  static_assert(kACol + 32 * kABufs <= 512, "TMEM columns");

  if (warp >= kAllocWarp) reg_dec<kRegsMisc>();
  if (warp == kProducerWarp) {
    // ================= producer: TMA loads of the X tiles, bulk copies of the B slices ==================
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        const MmaSide& s = p.side[it.side];
        const size_t n_stage_rows = (size_t)(s.n_pad / kStageCells);
        for (int tile = 0; tile < it.n_tiles; ++tile) {
          mbar_wait(&stage_empty[stage], ph ^ 1);
          unsigned char* dst = smem + (size_t)stage * S::kStageBytes;
          const long long row = it.row0 + (long long)tile * kStageCells;
          mbar_arrive_expect_tx(&stage_full[stage], S::kStageBytes);
          tma_load_2d(dst, &p.tmap[it.side], it.gt * kGT, (int)row, &stage_full[stage]);
          bulk_load_1d(dst + kXBytes, s.table + ((size_t)it.bg * n_stage_rows + (size_t)(row / kStageCells)) * S::kBBytes,
                       S::kBBytes, &stage_full[stage]);
          if (++stage == kStages) { stage = 0; ph ^= 1; }
        }
      }
    }
  }
  if (warp == kMmaWarp || warp == kMmaLoAWarp || warp == kMmaLoBWarp) {
    // ================= MMA issuers ========================================================================
    // kMmaWarp:    hhA (+)= xh.wh, hhB (+)= yh.wh   (chains of kFlushSteps steps, two accumulator buffers)
    // kMmaLoAWarp: loA (+)= xl.wh + xh.wl           (one chain per work item)
    // kMmaLoBWarp: loB (+)= yl.wh + yh.wl
    // All wait for the same A operands and B slices; a buffer / stage is free when all three have committed.
    if (lane == 0) {
      const bool hi_role = warp == kMmaWarp;
      const uint32_t idesc = instr_desc(kGT, N);
      int stage = 0;
      uint32_t sph = 0;
      uint32_t n_step = 0, n_chain = 0, n_item = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n_item) {
        const Item it = decode_item(p, item);
        const int total = it.n_tiles * kStepsPerStage;
        if (!hi_role) mbar_wait(lo_empty, (n_item & 1) ^ 1);
        const uint32_t lo_off = warp == kMmaLoBWarp ? 16u : 0u;   // A operands: x at +0 / +8, y at +16 / +24
        int ks = 0;
        for (int tile = 0; tile < it.n_tiles; ++tile) {
          mbar_wait(&stage_full[stage], sph);
          // descriptors of the stage's first step; later steps / the lo half are constant offsets (>> 4)
          const uint64_t d0 = smem_desc(smem_u32(smem + (size_t)stage * S::kStageBytes + kXBytes), N * 16, 128);
#pragma unroll
          for (int s4 = 0; s4 < kStepsPerStage; ++s4, ++ks, ++n_step) {
            const uint32_t ab = n_step % kABufs;
            const uint64_t d_hi = d0 + (uint64_t)((s4 * (N * 64)) >> 4);
            const uint64_t d_lo = d0 + (uint64_t)((s4 * (N * 64) + N * 32) >> 4);
            const uint32_t a0 = tbase + kACol + ab * 32;      // xh | xl | yh | yl, 8 columns each
            if (hi_role) {
              const uint32_t accb = n_chain & 1;
              const int in_chain = ks % kFlushSteps;
              if (in_chain == 0) mbar_wait(&acc_empty[accb], ((n_chain >> 1) & 1) ^ 1);
              mbar_wait(&a_full[ab], (n_step / kABufs) & 1);
              tc_fence_after();
              const uint32_t hh = tbase + accb * (2 * N);
              if (!(G2G_MMA_ABLATE & 1)) {
                umma_ts(hh, a0, d_hi, idesc, in_chain > 0);
                umma_ts(hh + N, a0 + 16, d_hi, idesc, in_chain > 0);
              }
              umma_commit(&a_empty[ab]);
              if (in_chain == kFlushSteps - 1 || ks + 1 == total) {
                umma_commit(&acc_full[accb]);
                ++n_chain;
              }
            } else {
              mbar_wait(&a_full[ab], (n_step / kABufs) & 1);
              tc_fence_after();
              const uint32_t lo = tbase + kLoCol + (lo_off ? N : 0);
              if (!(G2G_MMA_ABLATE & 1)) {
                umma_ts(lo, a0 + lo_off + 8, d_hi, idesc, ks > 0);
                umma_ts(lo, a0 + lo_off, d_lo, idesc, 1);
              }
              umma_commit(&a_empty[ab]);
              if (ks + 1 == total) umma_commit(lo_full);
            }
          }
          umma_commit(&stage_empty[stage]);
          if (++stage == kStages) { stage = 0; sph ^= 1; }
        }
      }
      // no barrier arrival of this CTA may still be in flight when its shared memory is released
      umma_commit(done_bar);
      mbar_wait(done_bar, 0);
    }
  } else if (warp >= kFirstFlushWarp && warp < kFirstConvWarp) {
    // ================= flushers: TMEM accumulators -> f32 registers (RN) -> f64 partial sums ============
    reg_inc<kRegsFlush>();
    const int q = warp & 3, h = (warp - kFirstFlushWarp) >> 2;   // lane quarter, half (0: A', 1: B)
    const uint32_t lane_base = tbase + ((uint32_t)(q * 32) << 16);
    const int T = p.n_bins, TB = p.bins_per_group;
    float acc[N];
    uint32_t n_chain = 0, n_item = 0;
    bool bad = false;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n_item) {
      const Item it = decode_item(p, item);
      const int total = it.n_tiles * kStepsPerStage;
      const int chains = (total + kFlushSteps - 1) / kFlushSteps;
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = 0.0f;
      for (int c = 0; c < chains; ++c, ++n_chain) {
        const uint32_t accb = n_chain & 1;
        if (warp == kFirstFlushWarp) mbar_wait_relaxed(&acc_full[accb], (n_chain >> 1) & 1);
        named_bar_sync(1, kFlushWarps * 32);
        tc_fence_after();
        const uint32_t src = lane_base + accb * (2 * N) + h * N;
#pragma unroll
        for (int j = 0; j < ((G2G_MMA_ABLATE & 4) ? 8 : N); j += 8) {
          uint32_t v[8];
          tmem_ld8(src + j, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[j + e] += __uint_as_float(v[e]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[accb]);
      }
      // item end: add the lo accumulator, write the f64 partial sums of this (item, gene)
      if (warp == kFirstFlushWarp) {
        mbar_wait_relaxed(lo_full, n_item & 1);
        mbar_wait_relaxed(&c_full[n_item & 1], (n_item >> 1) & 1);
      }
      named_bar_sync(1, kFlushWarps * 32);
      tc_fence_after();
      const MmaSide& s = p.side[it.side];
      const long long g = (long long)it.gt * kGT + q * 32 + lane;
      G2G_DEVICE_ASSERT(it.split >= 0 && it.split < s.n_split && it.gt * kGT < p.g_stride + kGT);
      double* pf = s.part_f64 + (size_t)it.split * (2 * T + 1) * p.g_stride + g;
#pragma unroll
      for (int j = 0; j < N; j += 8) {
        uint32_t v[8];
        tmem_ld8(lane_base + kLoCol + h * N + j, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int col = j + e;
          const double val = (double)acc[col] + (double)__uint_as_float(v[e]) * (1.0 / kLoScale);
          const int t = it.bg * TB + col;
          if (col < TB && t < T && g < p.n_genes) {
            pf[(size_t)(h * T + t) * p.g_stride] = val;
            if (!isfinite(val)) bad = true;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(lo_empty);
      if (h == 0) {   // sum (x - a): the converter groups' shares, added in group order
        const uint32_t par = n_item & 1;
        double c = 0.0;
#pragma unroll
        for (int k = 0; k < kConvGroups; ++k) c += c_s[(par * kConvGroups + k) * kGT + q * 32 + lane];
        if (it.bg == 0 && g < p.n_genes) pf[(size_t)(2 * T) * p.g_stride] = c;
        __syncwarp();
        if (lane == 0) mbar_arrive(&c_empty[par]);
      }
    }
    if (bad) atomicOr(p.status, 1);
  } else if (warp >= kFirstConvWarp && warp < kAllocWarp) {
    // ================= converters: x -> (xh, xl, yh, yl) in tensor memory, window counts ================
    reg_dec<kRegsConv>();
    const int cg = (warp - kFirstConvWarp) >> 2, q = warp & 3;
    const int gl = q * 32 + lane;
    const uint32_t lane_base = tbase + ((uint32_t)(q * 32) << 16);
    const int T = p.n_bins;
    int stage = 0;
    uint32_t sph = 0, n_step = 0, n_item = 0;
    float a_next = 0.0f;
    {
      if ((int)blockIdx.x < p.n_items) {
        const Item nx = decode_item(p, blockIdx.x);
        const long long g = (long long)nx.gt * kGT + gl;
        a_next = g < p.n_genes ? p.side[nx.side].pilot[g] : 0.0f;
      }
    }
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const Item it = decode_item(p, item);
      const MmaSide& s = p.side[it.side];
      const long long g = (long long)it.gt * kGT + gl;
      const float a = a_next;
      const unsigned long long neg_a2 = pack2(-a, -a);
      if (item + (int)gridDim.x < p.n_items) {
        const Item nx = decode_item(p, item + gridDim.x);
        const long long gn = (long long)nx.gt * kGT + gl;
        a_next = gn < p.n_genes ? p.side[nx.side].pilot[gn] : 0.0f;
      }
      const bool do_counts = it.bg == 0 && g < p.n_genes;
      int32_t* pi = s.part_i32 + (size_t)it.split * T * p.g_stride + g;
      int cur_win = 0, cnt = 0;   // window word: id + 1 (0 = padding rows)
      double c_acc = 0.0;         // this thread's share of sum (x - a) over the item
      const long long n_cells = s.n_cells;
      auto flush_counts = [&]() {
        if (do_counts && cur_win > 0 && cnt) atomicAdd(pi + (size_t)(cur_win - 1) * p.g_stride, cnt);
        cnt = 0;
      };
      for (int tile = 0; tile < it.n_tiles; ++tile) {
        bool waited = false;
        const unsigned char* st = smem + (size_t)stage * S::kStageBytes;
        const float* xt = reinterpret_cast<const float*>(st);
        const int* win = reinterpret_cast<const int*>(st + kXBytes + N * 256);
#pragma unroll 1
        for (int s4 = 0; s4 < kStepsPerStage; ++s4, ++n_step) {
          if ((int)(n_step % kConvGroups) != cg) continue;
          if (!waited) {
            if (q == 0) mbar_wait(&stage_full[stage], sph);
            named_bar_sync(2 + cg, 128);
            waited = true;
          }
          const float* xs = xt + (s4 * kStepCells) * kGT + gl;
          int nz = 0;
          unsigned long long csum2 = 0ull;
          const long long step_row = it.row0 + (long long)tile * kStageCells + s4 * kStepCells;
          const bool all_valid = step_row + kStepCells <= n_cells;
          const uint32_t ab = n_step % kABufs;
          const uint32_t dst = lane_base + kACol + ab * 32;
          // two halves of 8 cells (4 TMEM columns per operand each): 24 live values instead of 48
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t xh[4], xl[4], yh[4], yl[4];
#pragma unroll
            for (int j = 0; j < ((G2G_MMA_ABLATE & 2) ? 1 : 4); ++j) {
              const float x0 = xs[(8 * half + 2 * j) * kGT], x1 = xs[(8 * half + 2 * j + 1) * kGT];
              nz += nonzero_flag(x0) + nonzero_flag(x1);
              const unsigned long long x2 = pack2_asm(x0, x1);
              const unsigned long long d2 = add2(x2, neg_a2);
              csum2 = add2(csum2, d2);
              split2(x2, xh[j], xl[j]);
              split2(mul2(d2, d2), yh[j], yl[j]);
            }
            if (half == 0) {   // the buffer's previous occupants have been consumed by the MMAs
              if (q == 0) mbar_wait(&a_empty[ab], ((n_step / kABufs) & 1) ^ 1);
              named_bar_sync(2 + cg, 128);
              tc_fence_after();
            }
            if (!(G2G_MMA_ABLATE & 8)) {
              tmem_st4(dst + 4 * half, xh);
              tmem_st4(dst + 8 + 4 * half, xl);
              tmem_st4(dst + 16 + 4 * half, yh);
              tmem_st4(dst + 24 + 4 * half, yl);
            }
          }
          float csum = lo_f(csum2) + hi_f(csum2);
          if (!all_valid) {   // last step of a dataset: rows past n_cells are zero filled, leave them out
            csum = 0.0f;
            for (int c = 0; c < kStepCells; ++c)
              if (step_row + c < n_cells) csum += xs[c * kGT] - a;
          }
          c_acc += (double)csum;
          // window counts (uniform per step except at the few steps a window boundary falls into)
          const int w0 = win[s4 * kStepCells];
          if (!(w0 & G2G_WIN_MIXED)) {
            if (w0 != cur_win) { flush_counts(); cur_win = w0; }
            cnt += nz;
          } else {
            for (int c = 0; c < kStepCells; ++c) {
              const int w = win[s4 * kStepCells + c] & (G2G_WIN_MIXED - 1);
              if (w != cur_win) { flush_counts(); cur_win = w; }
              cnt += nonzero_flag(xs[c * kGT]);
            }
          }
          if (!(G2G_MMA_ABLATE & 8)) tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[ab]);
        }
        if (++stage == kStages) { stage = 0; sph ^= 1; }
      }
      flush_counts();
      {   // hand this group's share of sum (x - a) to the flushers
        const uint32_t par = n_item & 1;
        mbar_wait(&c_empty[par], ((n_item >> 1) & 1) ^ 1);
        c_s[(par * kConvGroups + cg) * kGT + gl] = c_acc;
        __syncwarp();
        if (lane == 0) mbar_arrive(&c_full[par]);
      }
      ++n_item;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kAllocWarp) tmem_dealloc(tbase, 512);
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

template <int N>
int launch_mma(MmaParams& p, const g2g_side_t* sides, int sm_reserve, cudaStream_t stream) {
  EncodeTiledFn encode;
  int rc = get_encode_fn(&encode);
  if (rc != G2G_OK) return rc;
  for (int s = 0; s < p.n_sides; ++s) {
    cuuint64_t gdim[2] = {(cuuint64_t)p.n_genes, (cuuint64_t)sides[s].n_cells};
    cuuint64_t gstride[1] = {(cuuint64_t)sides[s].ldx * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kGT, (cuuint32_t)kStageCells};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.tmap[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(sides[s].X), gdim,
                        gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      g2g_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
      return G2G_ERR_CUDA;
    }
  }
  p.n_gtiles = (int)g2g_ceil_div(p.n_genes, kGT);
  int items = 0;
  for (int s = 0; s < p.n_sides; ++s) {
    p.side[s].item_base = items;
    items += p.n_groups * p.side[s].n_split * p.n_gtiles;
  }
  p.n_items = items;
  int dev = 0, n_sm = 0;
  G2G_CUDA_TRY(cudaGetDevice(&dev));
  G2G_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = Smem<N>::kTotal;
  G2G_CUDA_TRY(cudaFuncSetAttribute(moments_mma_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = n_sm - sm_reserve;   // one CTA per SM: the kernel owns all 512 TMEM columns
  if (grid < 1) grid = 1;
  if (grid > items) grid = items;
  moments_mma_kernel<N><<<grid, kThreads, smem, stream>>>(p);
  G2G_CUDA_TRY(cudaGetLastError());
  return G2G_OK;
}

}  // namespace

extern "C" int g2g_moments_mma(const g2g_side_t* sides, int32_t n_sides, int64_t n_genes, int64_t g_stride,
                               int32_t n_bins, int32_t* status, void* stream) {
  return g2g_moments_mma_ex(sides, n_sides, n_genes, g_stride, n_bins, status, 0, stream);
}

extern "C" int g2g_moments_mma_ex(const g2g_side_t* sides, int32_t n_sides, int64_t n_genes, int64_t g_stride,
                                  int32_t n_bins, int32_t* status, int32_t sm_reserve, void* stream) {
  G2G_REQUIRE(sm_reserve >= 0 && sm_reserve < 128, "sm_reserve out of range");
  G2G_REQUIRE(sides != nullptr && n_sides >= 1 && n_sides <= kMaxSides, "n_sides must be 1 or 2");
  G2G_REQUIRE(n_bins >= 2 && n_bins <= G2G_MAX_BINS, "n_bins out of range [2, 128]");
  G2G_REQUIRE(n_genes >= 1 && g_stride >= n_genes, "bad gene counts");
  G2G_REQUIRE(n_genes < (1ll << 31), "n_genes too large");
  G2G_REQUIRE(status != nullptr, "null status pointer");
  G2GMmaTiling t = g2g_mma_tiling(n_bins);
  MmaParams p;
  p.n_genes = n_genes; p.g_stride = g_stride; p.n_sides = n_sides; p.n_bins = n_bins; p.n_groups = t.n_groups;
  p.bins_per_group = t.bins_per_group;
  p.status = status;
  cudaStream_t st = (cudaStream_t)stream;
  for (int s = 0; s < n_sides; ++s) {
    const g2g_side_t& in = sides[s];
    G2G_REQUIRE(in.X && in.wtab && in.pilot && in.part_f64 && in.part_i32, "null pointer in side");
    G2G_REQUIRE(in.n_cells >= 1 && in.n_cells < (1ll << 31), "n_cells out of range");
    G2G_REQUIRE(in.ldx >= n_genes && in.ldx % 4 == 0, "ldx must be >= n_genes and a multiple of 4 (16-byte rows for TMA)");
    G2G_REQUIRE(((uintptr_t)in.X & 15) == 0, "X must be 16-byte aligned");
    G2G_REQUIRE(((uintptr_t)in.wtab & 127) == 0, "the weight table must be 128-byte aligned");
    g2g_mma_layout_t lay;
    int rc = g2g_mma_layout(in.n_cells, n_bins, &lay);
    if (rc != G2G_OK) return rc;
    p.side[s].table = reinterpret_cast<const unsigned char*>(in.wtab);
    p.side[s].pilot = in.pilot;
    p.side[s].part_f64 = in.part_f64; p.side[s].part_i32 = in.part_i32;
    p.side[s].n_pad = lay.n_pad; p.side[s].cells_per_item = lay.cells_per_item;
    p.side[s].n_split = (int)lay.n_split;
    p.side[s].n_cells = in.n_cells;
    G2G_CUDA_TRY(cudaMemsetAsync(in.part_i32, 0, (size_t)lay.n_split * n_bins * g_stride * sizeof(int32_t), st));
  }
  switch (t.n_cols) {
    case 16: return launch_mma<16>(p, sides, sm_reserve, st);
    case 32: return launch_mma<32>(p, sides, sm_reserve, st);
    case 48: return launch_mma<48>(p, sides, sm_reserve, st);
    case 64: return launch_mma<64>(p, sides, sm_reserve, st);
    default:
      g2g_set_error("n_cols %d has no compiled kernel", t.n_cols);
      return G2G_ERR_UNSUPPORTED;
  }
}
