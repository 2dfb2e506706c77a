// Probe for the tensor-core variant of K1 (g2g_moments): tcgen05.mma kind::f16 with the A operand
// (expression values, one TMEM lane per gene) in tensor memory and the B operand (kernel weights,
// K-major, no swizzle) in shared memory; f32 accumulators in TMEM.
//
// Answers, on a B200:
//   E1  are the TMEM-A / smem-B layouts and descriptors right?   (D vs an exact f64 product of the
//       f16-rounded operands)
//   E2  how does the f32 accumulation inside the tensor core round?  (long chains of positive terms:
//       sign and size of the error against the exact sum of the same f16 products)
//   E3  error of the three-term split  x = xh + xl/2048, w = wh + wl/2048  (xh*wh + (xl*wh + xh*wl)/2048)
//       against the f64 product of the original f32 data
//   E4  issue rate: cycles per 16-cell step of 6 MMAs (N = 64, 112) with A in TMEM
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_ts_probe umma_ts_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes, rows 16 B apart; SBO between 8-row blocks, LBO
// between the two 16-byte K chunks of a K=16 step.
__host__ __device__ inline uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;
}
__host__ __device__ inline uint32_t instr_desc_f16(int M, int N) {
  return (1u << 4) /* D = f32 */ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int kM = 128;

// mode 0: hi only (D = rn16(A) * rn16(B));  mode 1: three-term split, D_hh in columns [0,N), D_lo in [N,2N)
// A: [128][K] f32 row major, B: [N][K] f32 row major; out: [128][2N] f32
template <int N>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B, int K,
                                                    int mode, float* __restrict__ out, long long* cycles, int reps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __half* b_hi = reinterpret_cast<__half*>(smem);               // [2 chunks][N/8][8 rows][8 halves] = N*32 bytes
  __half* b_lo = reinterpret_cast<__half*>(smem + N * 32);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t d_hh = tbase, d_lo = tbase + N;           // accumulators
  const uint32_t a_col = 4 * N;                            // A staging: xh at +0, xl at +8
  const uint32_t idesc = instr_desc_f16(kM, N);
  const uint64_t desc_hi = smem_desc(smem_u32(b_hi), N * 16, 128);
  const uint64_t desc_lo = smem_desc(smem_u32(b_lo), N * 16, 128);
  uint32_t phase = 0;
  long long t0 = 0, t1 = 0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    // A: this thread's row, 16 values -> packed pairs
    uint32_t ah[8], al[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x0 = A[(size_t)tid * K + k0 + 2 * j], x1 = A[(size_t)tid * K + k0 + 2 * j + 1];
      __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
      ah[j] = pack_h2(__half2float(h0), __half2float(h1));
      al[j] = pack_h2((x0 - __half2float(h0)) * 2048.0f, (x1 - __half2float(h1)) * 2048.0f);
    }
    tmem_st8(lane_addr + a_col, ah);
    tmem_st8(lane_addr + a_col + 8, al);
    tmem_st_wait();
    // B tile for this step: element (n, k) at chunk (k/8), block (n/8), row (n%8), e (k%8)
    for (int idx = tid; idx < N * 16; idx += 128) {
      int n = idx / 16, k = idx % 16;
      float w = B[(size_t)n * K + k0 + k];
      __half wh = __float2half_rn(w);
      size_t off = (size_t)(k / 8) * (N * 8) + (size_t)(n / 8) * 64 + (n % 8) * 8 + (k % 8);
      b_hi[off] = wh;
      b_lo[off] = __float2half_rn((w - __half2float(wh)) * 2048.0f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t acc = k0 > 0;
      umma_ts_f16(d_hh, tbase + a_col, desc_hi, idesc, acc);
      if (mode == 1) {
        umma_ts_f16(d_lo, tbase + a_col + 8, desc_hi, idesc, acc);   // xl * wh
        umma_ts_f16(d_lo, tbase + a_col, desc_lo, idesc, 1);          // xh * wl
      }
      umma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
  }
  // read back
  for (int c = 0; c < 2 * N; c += 16) {
    uint32_t v[16];
    tmem_ld16(lane_addr + c, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) out[(size_t)tid * 2 * N + c + j] = __uint_as_float(v[j]);
  }
  // E4: issue-rate loop: `reps` steps of 6 MMAs on whatever is in TMEM / smem
  tc_fence_before();
  __syncthreads();
  if (tid == 0 && reps > 0) {
    tc_fence_after();
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      umma_ts_f16(tbase + 0 * N, tbase + a_col, desc_hi, idesc, 1);
      umma_ts_f16(tbase + 1 * N, tbase + a_col + 8, desc_hi, idesc, 1);
      umma_ts_f16(tbase + 1 * N, tbase + a_col, desc_lo, idesc, 1);
      umma_ts_f16(tbase + 2 * N, tbase + a_col + 16, desc_hi, idesc, 1);
      umma_ts_f16(tbase + 3 * N, tbase + a_col + 24, desc_hi, idesc, 1);
      umma_ts_f16(tbase + 3 * N, tbase + a_col + 16, desc_lo, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, phase);
    t1 = clock64();
    cycles[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

static float h2f_round(float x) { return __half2float(__float2half_rn(x)); }

template <int N>
void run(int K, int mode, int data, int reps) {
  std::vector<float> A((size_t)kM * K), B((size_t)N * K), out((size_t)kM * 2 * N);
  srand(1234 + K + mode * 7 + data);
  for (auto& v : A) {
    float u = rand() / (float)RAND_MAX;
    v = data == 0 ? (u * 4.0f - 2.0f) : (data == 1 ? (0.5f + u * 3.0f) : (u < 0.5f ? 0.0f : log1pf(u * 40.0f)));
  }
  for (size_t n = 0; n < (size_t)N; ++n)
    for (int k = 0; k < K; ++k) {
      float u = rand() / (float)RAND_MAX;
      B[n * K + k] = data == 0 ? (u * 2.0f - 1.0f) : expf(-(u * 3.0f));
    }
  float *dA, *dB, *dO;
  long long* dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, out.size() * 4));
  CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0, out.size() * 4));
  CK(cudaMemset(dC, 0, 8));
  probe_kernel<N><<<1, 128, N * 64 + 1024>>>(dA, dB, K, mode, dO, dC, reps);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
  long long cyc = 0;
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  // references
  double max_rel_h = 0, sum_rel_h = 0, sum_signed_h = 0, max_rel_full = 0, sum_signed_full = 0, max_abs_layout = 0;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < N; ++n) {
      double exact_h = 0, exact_full = 0, absum = 0;
      for (int k = 0; k < K; ++k) {
        double a = A[(size_t)m * K + k], b = B[(size_t)n * K + k];
        exact_h += (double)h2f_round((float)a) * (double)h2f_round((float)b);
        exact_full += a * b;
        absum += fabs(a * b);
      }
      double got_h = out[(size_t)m * 2 * N + n];
      double got = got_h + (mode == 1 ? (double)out[(size_t)m * 2 * N + N + n] / 2048.0 : 0.0);
      double scale = absum > 0 ? absum : 1.0;
      double eh = (got_h - exact_h) / scale, ef = (got - exact_full) / scale;
      max_rel_h = fmax(max_rel_h, fabs(eh)); sum_rel_h += fabs(eh); sum_signed_h += eh;
      max_rel_full = fmax(max_rel_full, fabs(ef)); sum_signed_full += ef;
      max_abs_layout = fmax(max_abs_layout, fabs(got_h - exact_h));
    }
  const double cnt = (double)kM * N;
  printf("N=%3d K=%5d mode=%d data=%d | hh vs exact f16 products: max %.3e mean|.| %.3e mean signed %+.3e (abs max %.3e) |"
         " full vs f64(f32 data): max %.3e mean signed %+.3e",
         N, K, mode, data, max_rel_h, sum_rel_h / cnt, sum_signed_h / cnt, max_abs_layout, max_rel_full,
         sum_signed_full / cnt);
  if (reps > 0) printf(" | %d steps x 6 MMAs: %.1f cycles/step", reps, (double)cyc / reps);
  printf("\n");
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC);
}

int main() {
  CK(cudaFuncSetAttribute(probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 64 + 1024));
  CK(cudaFuncSetAttribute(probe_kernel<112>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 64 + 1024));
  printf("E1 layout check (signed data, short K; errors relative to sum |a b| should be ~1e-7 or below)\n");
  run<64>(16, 0, 0, 0);
  run<64>(64, 0, 0, 0);
  run<112>(64, 0, 0, 0);
  printf("E2 accumulation rounding (positive data; 'mean signed' < 0 with size ~ steps/2 * 6e-8 means truncation)\n");
  for (int K : {256, 1024, 4096, 16384}) run<64>(K, 0, 1, 0);
  printf("E3 three-term split vs f64 of the f32 data\n");
  for (int K : {256, 1024, 4096}) run<64>(K, 1, 1, 0);
  for (int K : {256, 1024, 4096}) run<64>(K, 1, 2, 0);
  run<112>(1024, 1, 2, 0);
  printf("E4 issue rate\n");
  run<64>(64, 1, 1, 2000);
  run<112>(64, 1, 1, 2000);
  return 0;
}
