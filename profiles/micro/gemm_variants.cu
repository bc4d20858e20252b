// Micro-benchmark (round 2): FP64 DMMA GEMM tile kernels for the 400x400 fit.  C[b] = A[b] B[b]
// (row-major n x n, n = 400) with an 80x80 CTA tile and different warp decompositions; measures the
// lone-product latency (batch = 1, with and without split-K) and the saturated rate (batch = 100).
//   WM x WN warps over the tile (warp tile (80/WM) x (80/WN)), KG k-groups (each takes every KG-th
//   k-step of a chunk; partial sums folded through shared memory), BK = k chunk, NST = cp.async stages,
//   PF = fragments of the next k-step are loaded before the current k-step's DMMAs are issued.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

constexpr int BT = 80;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(double* s, const double* g) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <int WM, int WN, int KG, int BK, int NST, bool PF>
struct Cfg {
  static constexpr int THREADS = 32 * WM * WN * KG;
  static constexpr int MI = BT / 8 / WM, NJ = BT / 8 / WN;  // 8x8 blocks per warp
  static constexpr int LDA = BK + 4;                        // A tile [80][BK] k contiguous
  static constexpr int LDB = BT + 4;                        // B tile [BK][80] n contiguous
  static constexpr int A_ELEMS = BT * LDA, B_ELEMS = BK * LDB;
  static constexpr int STAGE = A_ELEMS + B_ELEMS;
  static constexpr size_t SMEM = (size_t)NST * STAGE * sizeof(double);
};

// grid = (25 tiles, batch, ksplit); partial != nullptr: every z writes its own slab (reduced by reduce_kernel)
template <int WM, int WN, int KG, int BK, int NST, bool PF>
__global__ void __launch_bounds__(32 * WM * WN * KG)
gemm_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int n, int ksplit,
            double* __restrict__ partial) {
  using P = Cfg<WM, WN, KG, BK, NST, PF>;
  extern __shared__ double smem[];
  const int tiles_n = n / BT;
  const int m0 = (blockIdx.x / tiles_n) * BT, n0 = (blockIdx.x % tiles_n) * BT;
  const size_t nn = (size_t)n * n;
  const double* Ab = A + blockIdx.y * nn;
  const double* Bb = B + blockIdx.y * nn;
  const int chunks = n / BK;
  const int c_begin = chunks * blockIdx.z / ksplit, c_end = chunks * (blockIdx.z + 1) / ksplit;
  double* out = partial ? partial + ((size_t)blockIdx.y * ksplit + blockIdx.z) * nn : C + blockIdx.y * nn;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int kgroup = warp / (WM * WN), wq = warp % (WM * WN);
  const int rbase = (wq / WN) * (BT / WM), cbase = (wq % WN) * (BT / WN);
  double acc[P::MI][P::NJ][2];
#pragma unroll
  for (int i = 0; i < P::MI; ++i)
#pragma unroll
    for (int j = 0; j < P::NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  auto issue = [&](int c, int stage) {
    double* As = smem + (size_t)stage * P::STAGE;
    double* Bs = As + P::A_ELEMS;
    const int k0 = c * BK;
    for (int e = threadIdx.x; e < BT * (BK / 2); e += P::THREADS) {
      const int row = e / (BK / 2), seg = e % (BK / 2);
      cp_async16(As + row * P::LDA + seg * 2, Ab + (size_t)(m0 + row) * n + k0 + seg * 2);
    }
    for (int e = threadIdx.x; e < BK * (BT / 2); e += P::THREADS) {
      const int row = e / (BT / 2), seg = e % (BT / 2);
      cp_async16(Bs + row * P::LDB + seg * 2, Bb + (size_t)(k0 + row) * n + n0 + seg * 2);
    }
  };
  const int n_chunks = c_end - c_begin;
#pragma unroll
  for (int s = 0; s < NST - 1; ++s) {
    if (s < n_chunks) issue(c_begin + s, s);
    cp_commit();
  }
  constexpr int KSTEPS = BK / 4 / KG;
  for (int i = 0; i < n_chunks; ++i) {
    cp_wait<NST - 2>();
    __syncthreads();
    const int nxt = i + NST - 1;
    if (nxt < n_chunks) issue(c_begin + nxt, nxt % NST);
    cp_commit();
    const double* As = smem + (size_t)(i % NST) * P::STAGE;
    const double* Bs = As + P::A_ELEMS;
    double a[2][P::MI], b[2][P::NJ];
    auto load_frag = [&](int kq, int buf) {
      const int kk = (kq * KG + kgroup) * 4;
#pragma unroll
      for (int ii = 0; ii < P::MI; ++ii) a[buf][ii] = As[(rbase + 8 * ii + g) * P::LDA + kk + tg];
#pragma unroll
      for (int jj = 0; jj < P::NJ; ++jj) b[buf][jj] = Bs[(kk + tg) * P::LDB + cbase + 8 * jj + g];
    };
    if (PF) load_frag(0, 0);
#pragma unroll
    for (int kq = 0; kq < KSTEPS; ++kq) {
      const int cur = PF ? (kq & 1) : 0;
      if (PF) { if (kq + 1 < KSTEPS) load_frag(kq + 1, cur ^ 1); }
      else load_frag(kq, 0);
#pragma unroll
      for (int ii = 0; ii < P::MI; ++ii)
#pragma unroll
        for (int jj = 0; jj < P::NJ; ++jj) dmma884(acc[ii][jj][0], acc[ii][jj][1], a[cur][ii], b[cur][jj]);
    }
  }
  cp_wait<0>();
  if (KG > 1) {
    __syncthreads();
    double* park = smem;  // [KG-1][MI*NJ*2][32*WM*WN]
    constexpr int TW = 32 * WM * WN;
    const int tq = threadIdx.x % TW;
    if (kgroup > 0) {
#pragma unroll
      for (int ii = 0; ii < P::MI; ++ii)
#pragma unroll
        for (int jj = 0; jj < P::NJ; ++jj) {
          park[(((kgroup - 1) * P::MI * P::NJ + ii * P::NJ + jj) * 2 + 0) * TW + tq] = acc[ii][jj][0];
          park[(((kgroup - 1) * P::MI * P::NJ + ii * P::NJ + jj) * 2 + 1) * TW + tq] = acc[ii][jj][1];
        }
    }
    __syncthreads();
    if (kgroup != 0) return;
    for (int kg = 1; kg < KG; ++kg)
#pragma unroll
      for (int ii = 0; ii < P::MI; ++ii)
#pragma unroll
        for (int jj = 0; jj < P::NJ; ++jj) {
          acc[ii][jj][0] += park[(((kg - 1) * P::MI * P::NJ + ii * P::NJ + jj) * 2 + 0) * TW + tq];
          acc[ii][jj][1] += park[(((kg - 1) * P::MI * P::NJ + ii * P::NJ + jj) * 2 + 1) * TW + tq];
        }
  }
#pragma unroll
  for (int ii = 0; ii < P::MI; ++ii)
#pragma unroll
    for (int jj = 0; jj < P::NJ; ++jj)
      *reinterpret_cast<double2*>(out + (size_t)(m0 + rbase + 8 * ii + g) * n + n0 + cbase + 8 * jj + 2 * tg) =
          make_double2(acc[ii][jj][0], acc[ii][jj][1]);
}

__global__ void reduce_kernel(const double* __restrict__ partial, double* __restrict__ C, size_t nn, int ksplit) {
  const size_t b = blockIdx.y;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x) {
    double v = 0;
    for (int z = 0; z < ksplit; ++z) v += partial[(b * ksplit + z) * nn + e];
    C[b * nn + e] = v;
  }
}
__global__ void naive_kernel(const double* A, const double* B, double* C, int n) {
  const int r = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double s = 0;
  for (int k = 0; k < n; ++k) s += A[(size_t)r * n + k] * B[(size_t)k * n + c];
  C[(size_t)r * n + c] = s;
}

static double *dA, *dB, *dC, *dP, *dRef;
static int g_sms, g_clk;

template <int WM, int WN, int KG, int BK, int NST, bool PF>
void bench(const char* name) {
  using P = Cfg<WM, WN, KG, BK, NST, PF>;
  const int n = 400;
  const size_t nn = (size_t)n * n;
  auto kern = gemm_kernel<WM, WN, KG, BK, NST, PF>;
  size_t smem = P::SMEM;
  const size_t park = (size_t)(KG - 1) * P::MI * P::NJ * 2 * 32 * WM * WN * sizeof(double);
  if (park > smem) smem = park;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, P::THREADS, smem);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](int batch, int ksplit, int reps) -> float {
    float best = 1e9;
    for (int r = 0; r < reps + 2; ++r) {
      cudaEventRecord(e0);
      kern<<<dim3(25, batch, ksplit), P::THREADS, smem>>>(dA, dB, dC, n, ksplit, ksplit > 1 ? dP : nullptr);
      if (ksplit > 1) reduce_kernel<<<dim3(40, batch), 256>>>(dP, dC, nn, ksplit);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r >= 2 && ms < best) best = ms;
    }
    return best * 1e3f;
  };
  // correctness (batch 1, ksplit 1 and 5)
  double maxerr = 0;
  for (int ks : {1, 5}) {
    cudaMemset(dC, 0, nn * 8);
    kern<<<dim3(25, 1, ks), P::THREADS, smem>>>(dA, dB, dC, n, ks, ks > 1 ? dP : nullptr);
    if (ks > 1) reduce_kernel<<<dim3(40, 1), 256>>>(dP, dC, nn, ks);
    std::vector<double> h(nn), ref(nn);
    cudaMemcpy(h.data(), dC, nn * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ref.data(), dRef, nn * 8, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < nn; ++i) maxerr = fmax(maxerr, fabs(h[i] - ref[i]));
  }
  const float t1 = run(1, 1, 10), t1k5 = run(1, 5, 10), t1k10 = run(1, 10, 10), t4 = run(4, 1, 10), t4k3 = run(4, 3, 10), t8 = run(8, 1, 10), t8k2 = run(8, 2, 10);
  const float t100 = run(100, 1, 5);
  cudaError_t err = cudaGetLastError();
  printf("%-34s thr %4d regs %3d occ %d err %.1e | b1 %6.1f  b1k5 %6.1f  b1k10 %6.1f | b4 %6.1f b4k3 %6.1f | b8 %6.1f b8k2 %6.1f | b100 %7.1f us = %5.2f TF/s %s\n",
         name, P::THREADS, fa.numRegs, occ, maxerr, t1, t1k5, t1k10, t4, t4k3, t8, t8k2, t100, 100 * 2.0 * 400 * 400 * 400 / (t100 * 1e-6) / 1e12,
         err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main(int argc, char** argv) {
  const bool quick = argc > 1;
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&g_clk, cudaDevAttrClockRate, 0);
  const int n = 400, batch = 100;
  const size_t nn = (size_t)n * n;
  cudaMalloc(&dA, batch * nn * 8); cudaMalloc(&dB, batch * nn * 8); cudaMalloc(&dC, batch * nn * 8);
  cudaMalloc(&dP, (size_t)batch * 2 * nn * 8 + 16 * nn * 8); cudaMalloc(&dRef, nn * 8);
  std::vector<double> h(batch * nn);
  for (auto& v : h) v = (rand() % 2001 - 1000) * 1e-3;
  cudaMemcpy(dA, h.data(), batch * nn * 8, cudaMemcpyHostToDevice);
  for (auto& v : h) v = (rand() % 2001 - 1000) * 1e-3;
  cudaMemcpy(dB, h.data(), batch * nn * 8, cudaMemcpyHostToDevice);
  naive_kernel<<<dim3((n + 127) / 128, n), 128>>>(dA, dB, dRef, n);
  printf("columns: batch b, split-K k (with a separate reduce launch); times in us\n");
  //      WM WN KG BK NST PF
  bench<2, 2, 1, 16, 3, false>("4w 40x40 bk16 st3 (round 1)");
  bench<5, 2, 1, 16, 3, true>("10w 16x40 bk16 st3 prefetch");
  bench<2, 2, 2, 16, 3, false>("8w 40x40 kg2 bk16 st3");
  if (quick) return 0;
  bench<2, 2, 1, 16, 3, true>("4w 40x40 bk16 st3 prefetch");
  bench<2, 2, 1, 32, 3, true>("4w 40x40 bk32 st3 prefetch");
  bench<2, 2, 2, 16, 3, false>("8w 40x40 kg2 bk16 st3");
  bench<2, 2, 2, 32, 3, true>("8w 40x40 kg2 bk32 st3 prefetch");
  bench<2, 2, 4, 32, 3, false>("16w 40x40 kg4 bk32 st3");
  bench<5, 2, 1, 16, 3, false>("10w 16x40 bk16 st3");
  bench<5, 2, 1, 16, 3, true>("10w 16x40 bk16 st3 prefetch");
  bench<5, 2, 1, 32, 3, true>("10w 16x40 bk32 st3 prefetch");
  bench<2, 5, 1, 16, 3, true>("10w 40x16 bk16 st3 prefetch");
  bench<5, 2, 2, 32, 3, true>("20w 16x40 kg2 bk32 st3 prefetch");
  bench<5, 1, 1, 16, 3, true>("5w 16x80 bk16 st3 prefetch");
  bench<1, 5, 1, 16, 3, true>("5w 80x16 bk16 st3 prefetch");
  bench<5, 5, 1, 32, 3, false>("25w 16x16 bk32 st3");
  bench<2, 2, 1, 16, 4, true>("4w 40x40 bk16 st4 prefetch");
  return 0;
}
