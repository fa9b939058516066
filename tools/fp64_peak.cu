// fp64_peak.cu -- microbenchmark for the FP64 roofline denominators on B200 (sm_100a):
//   (1) DFMA pipe peak, (2) DMMA m8n8k4 tensor peak, (3) both issued from the same warps.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
// Prints one JSON line; bench.py reads the committed result from profiles/.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k_peak(double *out, int iters, double seed) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x * 1e-3;
  double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;   // dmma accumulators (2 regs each x 2 mmas)
  const double b = 1.0000001, c = 1e-9;
  const double ma = 1.0 + threadIdx.x * 1e-6, mb = 0.5;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1) : "d"(ma), "d"(mb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c2), "+d"(c3) : "d"(ma), "d"(mb));
      }
    }
  }
  double s = c0 + c1 + c2 + c3;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(int iters, double *d_out, int grid) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_peak<MODE><<<grid, 256>>>(d_out, iters / 10, 1.0);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_peak<MODE><<<grid, 256>>>(d_out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best * 1e-3;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int grid = p.multiProcessorCount * 8;
  double *d_out; cudaMalloc(&d_out, sizeof(double) * grid * 256);
  const int iters = 20000;
  const double threads = (double)grid * 256;
  const double t0 = run<0>(iters, d_out, grid);
  const double t1 = run<1>(iters, d_out, grid);
  const double t2 = run<2>(iters, d_out, grid);
  const double fma_flops = threads * iters * 64.0 * 2.0;                  // 64 DFMA per iter
  const double mma_flops = (threads / 32.0) * iters * 16.0 * (8 * 8 * 4 * 2.0);  // 16 mma per iter per warp
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, "
         "\"both_dfma_tflops\": %.2f, \"both_dmma_tflops\": %.2f, \"both_total_tflops\": %.2f}\n",
         p.name, p.multiProcessorCount, fma_flops / t0 * 1e-12, mma_flops / t1 * 1e-12,
         fma_flops / t2 * 1e-12, mma_flops / t2 * 1e-12, (fma_flops + mma_flops) / t2 * 1e-12);
  return 0;
}
