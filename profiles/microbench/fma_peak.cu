// Micro-benchmark: peak FFMA vs FFMA2 (fma.rn.f32x2) issue rate on sm_100a, per SM per clock.
// Used to set the compute ceiling of the K1 register tile.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, float x) {
  asm volatile("{\n.reg .b64 xx;\nmov.b64 xx, {%2, %2};\nfma.rn.f32x2 %0, %1, xx, %0;\n}" : "+l"(acc) : "l"(a), "f"(x));
}
__device__ __forceinline__ void ffma2_full(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float x = seed + threadIdx.x * 1e-9f;
  if (MODE == 0) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = i;
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = seed * (i + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fmaf(w[i & 7], x, acc[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    unsigned long long acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = i;
    unsigned long long w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = (unsigned long long)__float_as_uint(seed * (i + 1)) * 0x100000001ull;
    unsigned long long xx = (unsigned long long)__float_as_uint(x) * 0x100000001ull + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 1) ffma2(acc[i], w[i & 3], x); else ffma2_full(acc[i], w[i & 3], xx);
      }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
  }
}

template <int MODE>
void run(const char* name, int threads, int blocks_per_sm) {
  int dev = 0, nsm = 0, khz = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  float* out;
  cudaMalloc(&out, sizeof(float) * nsm * blocks_per_sm * threads);
  int iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<nsm * blocks_per_sm, threads>>>(out, 100, 1.0f);
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a);
    k<MODE><<<nsm * blocks_per_sm, threads>>>(out, iters, 1.0f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  double fmas = (double)nsm * blocks_per_sm * threads * iters * 32.0;
  double per_clk_sm = fmas / (best * 1e-3) / nsm / (khz * 1e3);
  printf("%-28s threads/block %4d blocks/SM %d : %.3f ms  %.1f TFLOP/s  %.1f FMA/clk/SM (at max clock %d MHz)\n", name,
         threads, blocks_per_sm, best, 2 * fmas / (best * 1e-3) / 1e12, per_clk_sm, khz / 1000);
  cudaFree(out);
}

int main() {
  for (int bps = 1; bps <= 2; ++bps) {
    for (int th : {128, 256, 512, 1024}) {
      run<0>("FFMA (3-reg)", th, bps);
      run<1>("FFMA2 (scalar-broadcast b)", th, bps);
      run<2>("FFMA2 (packed b)", th, bps);
    }
  }
  return 0;
}
