// ubench_ffma2.cu — what is the FP32 roofline of the CUDA-core path on B200?
//   (a) scalar FFMA stream            (the figure bench.py's roofline uses: spi_b200_fp32_peak)
//   (b) packed FFMA2 stream           (fma.rn.f32x2, sm_100+): same lanes or twice the rate?
//   (c) scalar FFMA + ALU-pipe filler (FMNMX) 1:1 — does a second pipe issue alongside a saturated FMA pipe?
//   (d) FFMA2 + FMNMX filler 1:1      — do the issue slots FFMA2 frees let other work through?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_ffma2.bin tools/ubench_ffma2.cu
#include <cuda_runtime.h>

#include <cstdio>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, float seed, float* sink) {
  const float m = 0.999f, b = 0.001f;
  float r = 0.f;
  if (MODE == 0 || MODE == 2) {
    float a[8], f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + i; f[i] = seed - i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 16; u++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          a[i] = fmaf(a[i], m, b);
          if (MODE == 2) f[i] = fminf(f[i], a[(i + 3) & 7]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) r += a[i] + f[i];
  } else {
    float2 a[8];
    float f[8];
    const float2 m2 = make_float2(m, m), b2 = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = make_float2(seed + i, seed - i); f[i] = seed - i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 16; u++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          a[i] = __ffma2_rn(a[i], m2, b2);
          if (MODE == 3) f[i] = fminf(f[i], a[(i + 3) & 7].x);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) r += a[i].x + a[i].y + f[i];
  }
  if (r == 123456.789f) sink[0] = r;
}

template <int MODE> int run(const char* name, int flops_per_inst, int sms) {
  float* sink;
  CK(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int grid = sms * 8, threads = 256, iters = 8192;
  k<MODE><<<grid, threads>>>(iters / 4, 1.f, sink);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    k<MODE><<<grid, threads>>>(iters, 1.f, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double inst = 8.0 * 16.0 * iters * (double)grid * threads;   // FMA-type thread instructions
  printf("%-28s %8.3f ms  %7.2f TFLOP/s  %6.3f FMA-inst/clk/SMSP@1.965GHz\n", name, best,
         inst * flops_per_inst / (best * 1e-3) / 1e12, inst / 32.0 / (best * 1e-3) / 1.965e9 / (sms * 4));
  cudaFree(sink);
  return 0;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d\n", sms);
  run<0>("FFMA", 2, sms);
  run<1>("FFMA2", 4, sms);
  run<2>("FFMA + FMNMX 1:1", 2, sms);
  run<3>("FFMA2 + FMNMX 1:1", 4, sms);
  return 0;
}
