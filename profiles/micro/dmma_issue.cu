// Micro-benchmark (round 2): DMMA.8x8x4 issue rate per SM as a function of warps per SM and of
// independent accumulators per warp (register operands only, no memory traffic).
#include <cuda_runtime.h>
#include <cstdio>

template <int NACC>
__global__ void dmma_kernel(int iters, double* out) {
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  if (s == 1.2345) out[0] = s;
}

template <int NACC>
void run(int sms, int clk, double* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps : {1, 2, 4, 8, 10, 12, 16, 20, 24, 32}) {
    const int iters = 4000;
    dmma_kernel<NACC><<<sms, warps * 32>>>(10, out);
    cudaEventRecord(e0);
    dmma_kernel<NACC><<<sms, warps * 32>>>(iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double clks = ms * 1e-3 * clk * 1e3;
    const double dmma_per_sm = (double)warps * iters * NACC;
    printf("acc/warp %2d warps/SM %2d: %7.3f clk per DMMA per SM, %6.1f clk per DMMA per warp, %6.2f TFLOP/s chip\n", NACC, warps,
           clks / dmma_per_sm, clks / (iters * (double)NACC), dmma_per_sm * 512.0 * sms / (ms * 1e-3) / 1e12);
  }
}

int main() {
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, 8);
  run<25>(sms, clk, out);
  run<10>(sms, clk, out);
  run<5>(sms, clk, out);
  run<1>(sms, clk, out);
  return 0;
}
