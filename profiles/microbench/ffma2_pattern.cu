// Micro-benchmark: the exact FFMA2 operand pattern of K1's inner loop (7 weight pairs x {x0,x1,y0,y1}
// -> 28 distinct 64-bit accumulators) without any memory traffic, to get the in-situ FMA-pipe cost of
// one FFMA2, versus the same arithmetic written with scalar FFMA.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, float x) {
  asm volatile("{\n.reg .b64 xx;\nmov.b64 xx, {%2, %2};\nfma.rn.f32x2 %0, %1, xx, %0;\n}" : "+l"(acc) : "l"(a), "f"(x));
}

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float x0 = seed + threadIdx.x * 1e-9f, x1 = x0 * 1.1f, y0 = x0 * x0, y1 = x1 * x1;
  if (MODE == 0) {
    unsigned long long acc[4][7];
    unsigned long long w[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) w[i] = (unsigned long long)__float_as_uint(seed * (i + 1)) * 0x100000001ull + i;
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int i = 0; i < 7; ++i) acc[g][i] = g * 7 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        ffma2(acc[0][i], w[i], x0); ffma2(acc[1][i], w[i], x1);
        ffma2(acc[2][i], w[i], y0); ffma2(acc[3][i], w[i], y1);
      }
      x0 += 1e-7f; x1 += 1e-7f; y0 += 1e-7f; y1 += 1e-7f;   // 4 FADD like the d / y updates
    }
    unsigned long long s = 0;
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int i = 0; i < 7; ++i) s += acc[g][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
  } else {
    float acc[4][14], w[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) w[i] = seed * (i + 1);
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int i = 0; i < 14; ++i) acc[g][i] = g * 14 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 14; ++i) {
        acc[0][i] = fmaf(w[i], x0, acc[0][i]); acc[1][i] = fmaf(w[i], x1, acc[1][i]);
        acc[2][i] = fmaf(w[i], y0, acc[2][i]); acc[3][i] = fmaf(w[i], y1, acc[3][i]);
      }
      x0 += 1e-7f; x1 += 1e-7f; y0 += 1e-7f; y1 += 1e-7f;
    }
    float s = 0;
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int i = 0; i < 14; ++i) s += acc[g][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

template <int MODE>
void run(const char* name, int threads, int bps) {
  int nsm = 0, khz = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * nsm * bps * threads);
  const int iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<nsm * bps, threads>>>(out, 100, 1.0f);
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a);
    k<MODE><<<nsm * bps, threads>>>(out, iters, 1.0f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  const double fmas = (double)nsm * bps * threads * iters * 56.0;
  const double warps_per_smsp = threads * bps / 128.0;
  const double cyc_per_iter_smsp = best * 1e-3 * khz * 1e3 / iters / 1.0;   // wall cycles per loop iteration
  printf("%-22s %4d thr x %d CTA/SM (%.0f warps/SMSP): %.3f ms, %.1f FMA/clk/SM, %.1f scheduler cycles per 56-FMA iteration per warp-slot\n",
         name, threads, bps, warps_per_smsp, best, fmas / (best * 1e-3) / nsm / (khz * 1e3),
         cyc_per_iter_smsp / warps_per_smsp);
  cudaFree(out);
}

int main() {
  for (int th : {128, 256, 512}) {
    run<0>("K1 pattern, FFMA2", th, 1);
    run<1>("K1 pattern, FFMA", th, 1);
  }
  run<0>("K1 pattern, FFMA2", 256, 2);
  run<1>("K1 pattern, FFMA", 256, 2);
  return 0;
}
