// fp_peak.cu -- build-time microbenchmark of the non-tensor FP pipes (SURVEY 8d asks for measured
// FP32/FP64 FMA peaks next to the roofline): dependent-chain latency and multi-warp throughput
// of DFMA / FFMA / MUFU.RCP64H / LDS.64 on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp_peak fp_peak.cu && ./fp_peak
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int MODE, int ILP>
__global__ void __launch_bounds__(1024) k_chain(double *out, long long *cyc, int iters, double seed) {
  double a[ILP];
  float fa[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) { a[k] = seed + k + threadIdx.x * 1e-9; fa[k] = (float)a[k]; }
  const double m = 1.0 + seed * 1e-12, c = seed * 1e-13;
  __shared__ double sm[1024];
  sm[threadIdx.x] = seed;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      if (MODE == 0) a[k] = fma(a[k], m, c);                    // DFMA
      if (MODE == 1) fa[k] = fmaf(fa[k], (float)m, (float)c);   // FFMA
      if (MODE == 2) { double x; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a[k])); a[k] = x; }
      if (MODE == 3) { a[k] = sm[((int)__double2hiint(a[k]) + threadIdx.x) & 1023]; }  // dependent LDS.64
      if (MODE == 4) a[k] = a[k] + c;                           // DADD
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += a[k] + fa[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int ILP>
static void run(const char *name, int blocks, int threads, int iters, double flop_per_op) {
  double *out; long long *cyc, h = 0;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaMalloc(&cyc, sizeof(long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_chain<MODE, ILP><<<blocks, threads>>>(out, cyc, 10, 1.5);
  cudaEventRecord(e0);
  k_chain<MODE, ILP><<<blocks, threads>>>(out, cyc, iters, 1.5);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double ops = (double)blocks * threads * iters * ILP;
  printf("%-12s ILP=%d grid=%dx%d : %.2f cycles/iter (block 0), %.3f Tops/s, %.2f TFLOP/s\n", name, ILP, blocks,
         threads, (double)h / iters, ops / (ms * 1e-3) / 1e12, ops * flop_per_op / (ms * 1e-3) / 1e12);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  const int it = 20000;
  run<0, 1>("DFMA", 1, 32, it, 2);
  run<0, 2>("DFMA", 1, 32, it, 2);
  run<0, 4>("DFMA", 1, 32, it, 2);
  run<0, 8>("DFMA", 1, 32, it, 2);
  run<0, 1>("DFMA", 1, 128, it, 2);
  run<0, 4>("DFMA", 1, 128, it, 2);
  run<0, 8>("DFMA", 148 * 2, 1024, it, 2);
  run<4, 8>("DADD", 148 * 2, 1024, it, 1);
  run<1, 1>("FFMA", 1, 32, it, 2);
  run<1, 8>("FFMA", 148 * 2, 1024, it, 2);
  run<2, 1>("RCP64H", 1, 32, it, 1);
  run<2, 4>("RCP64H", 1, 32, it, 1);
  run<2, 4>("RCP64H", 148 * 2, 1024, it / 4, 1);
  run<3, 1>("LDS.64dep", 1, 32, it, 1);
  return 0;
}
