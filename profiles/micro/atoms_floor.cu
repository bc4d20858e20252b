// Micro-benchmark (round 2): issue cost of shared-memory reductions with NO other memory traffic.
// Addresses are computed in registers; `ways` lanes share each bank (distinct addresses), so an
// instruction needs `ways` wavefronts.  Variants: red.add 1 (ATOMS.POPC.INC), red.add v (ATOMS.ADD),
// plain st.shared (STS) and ld.shared (LDS) for comparison.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

template <int OP>
__global__ void __launch_bounds__(1024, 1) floor_kernel(int iters, int ways, int active, unsigned long long* out) {
  extern __shared__ uint32_t hist[];
  for (int i = threadIdx.x; i < 40960; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(hist);
  // lane -> bank (lane % (32/ways)), row lane / (32/ways): `ways` distinct addresses per bank
  const int nb = 32 / ways;
  uint32_t a = sbase + 4u * ((lane % nb) + 32 * (lane / nb) + 1024 * warp);
  uint32_t acc = 0;
  if (lane < active) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const uint32_t addr = a + 128u * ((i + u) & 7);  // same banks, rotating rows
        if (OP == 0) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
        if (OP == 1) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(addr | 1u) : "memory");
        if (OP == 2) asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(addr) : "memory");
        if (OP == 3) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); acc += v; }
      }
    }
  }
  __syncthreads();
  if (acc == 0x1234567u) out[0] = acc;
}

int main() {
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  unsigned long long* out; cudaMalloc(&out, 8);
  const int smem = 40960 * 4;
  cudaFuncSetAttribute(floor_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(floor_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(floor_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(floor_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  const char* names[4] = {"ATOMS.POPC.INC", "ATOMS.ADD", "STS", "LDS"};
  printf("%-16s %5s %7s %12s\n", "op", "ways", "active", "clk/instr/SM");
  for (int op = 0; op < 4; ++op)
    for (int active : {32, 16}) 
    for (int ways : {1, 2, 4, 8}) {
      auto launch = [&](int it) {
        if (op == 0) floor_kernel<0><<<sms, 1024, smem>>>(it, ways, active, out);
        if (op == 1) floor_kernel<1><<<sms, 1024, smem>>>(it, ways, active, out);
        if (op == 2) floor_kernel<2><<<sms, 1024, smem>>>(it, ways, active, out);
        if (op == 3) floor_kernel<3><<<sms, 1024, smem>>>(it, ways, active, out);
      };
      launch(10);
      cudaEventRecord(e0); launch(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double instr_per_sm = 32.0 * iters * 16;
      printf("%-16s %5d %7d %12.3f\n", names[op], ways, active, ms * 1e-3 * clk * 1e3 / instr_per_sm);
    }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); return 1; }
  return 0;
}
