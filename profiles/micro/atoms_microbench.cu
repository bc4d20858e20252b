// Micro-benchmark (round 2): cost of shared-memory reductions (red.shared.add.u32 -> ATOMS.POPC.INC /
// ATOMS.ADD) under the address patterns of count_lg_kernel and of candidate re-layouts.  Address
// sets are generated on the host with the statistics of the bench workload (BASELINE config 3:
// 13 % gaps per row, 34 % mutated sites, log-normal pair lengths, 4 rate categories, K = 100).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_microbench atoms_microbench.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

constexpr int K = 100, S = 20, L = 300, R = 4, NPAIR = 512;
constexpr uint32_t SKIP = 0xffffffffu;  // lane does not take part
constexpr int NINSTR = 4096;             // warp-instructions per pattern

__global__ void __launch_bounds__(1024, 1) atoms_kernel(const uint32_t* __restrict__ addr, int n_instr, int reps,
                                                       int mode, unsigned long long* out) {
  extern __shared__ uint32_t hist[];
  for (int i = threadIdx.x; i < K * 401 + 64; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(hist);
  for (int r = 0; r < reps; ++r) {
    // every warp walks the instruction list from its own starting point
    int i = (warp * 97 + blockIdx.x * 13 + r * 31) % n_instr;
    for (int it = 0; it < n_instr; it += 16) {
      uint32_t a[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        int idx = i + u; if (idx >= n_instr) idx -= n_instr;
        a[u] = __ldg(addr + (size_t)idx * 32 + lane);
      }
      if (mode == 0) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (a[u] != SKIP) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(sbase + 4u * a[u]) : "memory");
      } else if (mode == 1) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (a[u] != SKIP) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(sbase + 4u * a[u]), "r"(a[u] | 1u) : "memory");
      } else {  // loads only: the floor of this harness
        uint32_t x = 0;
#pragma unroll
        for (int u = 0; u < 16; ++u) x ^= a[u];
        if (x == 0x12345u) hist[0] = x;
      }
      i += 16; if (i >= n_instr) i -= n_instr;
    }
  }
  __syncthreads();
  unsigned long long s = 0;
  for (int i = threadIdx.x; i < K * 401 + 64; i += blockDim.x) s += hist[i];
  if (s == 12345ull) out[0] = s;
}

struct Data {
  std::vector<uint8_t> a, b;  // [NPAIR][L]
  std::vector<double> t;
  std::vector<int> cat;       // [L] sorted
  std::vector<int> bk;        // [NPAIR][R]
  std::vector<int> order;     // pairs sorted by t
};

static int quant(double t, const std::vector<double>& q) {
  if (t < q[0] || t > q[K - 1]) return -1;
  int lo = (int)(std::lower_bound(q.begin(), q.end(), t) - q.begin());
  if (lo == 0) return 0;
  return (t / q[lo - 1] - 1 < q[lo] / t - 1) ? lo - 1 : lo;
}

int main(int argc, char** argv) {
  std::mt19937_64 rng(1234);
  std::uniform_real_distribution<double> U(0, 1);
  std::lognormal_distribution<double> LN(std::log(0.52), 0.95);
  std::vector<double> q(K);
  for (int i = 0; i < K; ++i) q[i] = 0.03 * std::pow(1.1, i - 50);
  const double rates[4] = {0.15, 0.5, 1.0, 2.35};
  Data d;
  d.a.resize(NPAIR * L); d.b.resize(NPAIR * L); d.t.resize(NPAIR); d.cat.resize(L); d.bk.resize(NPAIR * R);
  for (int s = 0; s < L; ++s) d.cat[s] = (int)(U(rng) * R);
  std::sort(d.cat.begin(), d.cat.end());
  for (int p = 0; p < NPAIR; ++p) {
    d.t[p] = LN(rng);
    for (int r = 0; r < R; ++r) d.bk[p * R + r] = quant(d.t[p] * rates[r], q);
    for (int s = 0; s < L; ++s) {
      int x = (int)(U(rng) * 20), y = x;
      if (U(rng) < 0.34) y = (int)(U(rng) * 20);
      if (U(rng) < 0.13) x = 20;
      if (U(rng) < 0.13) y = 20;
      d.a[p * L + s] = x; d.b[p * L + s] = y;
    }
  }
  d.order.resize(NPAIR);
  for (int p = 0; p < NPAIR; ++p) d.order[p] = p;
  std::sort(d.order.begin(), d.order.end(), [&](int u, int v) { return d.t[u] < d.t[v]; });
  const int nch = (L + 15) / 16;
  auto cell = [&](int p, int s, int stride, int& diag) -> int {  // -1 = skipped
    if (s >= L) return -1;
    int x = d.a[p * L + s], y = d.b[p * L + s], b = d.bk[p * R + d.cat[s]];
    if (x == 20 || y == 20 || b < 0) return -1;
    diag = (x == y);
    return b * stride + 20 * x + y;
  };
  struct Pat { const char* name; std::vector<uint32_t> addr; int mode; double sites_per_instr; };
  std::vector<Pat> pats;
  auto rnd = [&](int n) { return (int)(U(rng) * n); };
  const int JUNK400 = K * 400, JUNK401 = K * 401;
  {  // P0: ideal, lane-private banks
    Pat p{"ideal: 32 distinct banks", {}, 0, 32};
    for (int i = 0; i < NINSTR; ++i) for (int l = 0; l < 32; ++l) p.addr.push_back((rnd(1000) * 32 + l));
    pats.push_back(p);
    Pat p2 = p; p2.name = "ideal, variable addend (ATOMS.ADD)"; p2.mode = 1; pats.push_back(p2);
    Pat p3 = p; p3.name = "harness floor: address loads only"; p3.mode = 2; pats.push_back(p3);
  }
  {  // P1: the current kernel: lanes = 32 consecutive 16-site chunks (pair-major), site k of each
    Pat p{"current kernel (chunk-major lanes, one junk word)", {}, 0, 32};
    for (int i = 0; i < NINSTR; ++i) {
      int i0 = rnd(NPAIR * nch - 32), k = rnd(16);
      for (int l = 0; l < 32; ++l) {
        int it = i0 + l, pr = it / nch, c = it % nch, dg;
        int ce = cell(pr, c * 16 + k, 400, dg);
        p.addr.push_back(ce < 0 ? JUNK400 : ce);
      }
    }
    pats.push_back(p);
  }
  {  // P2: lanes = 32 consecutive 4-site words of ONE (pair, category) run stream sorted by t: single bucket
    Pat comb{"single bucket per instr: combined, junk in a free bank", {}, 0, 32};
    Pat diag{"single bucket per instr: diagonal only (others -> free-bank junk)", {}, 0, 32};
    Pat off{"single bucket per instr: off-diagonal lanes only (others inactive)", {}, 0, 0};
    double offs = 0;
    for (int i = 0; i < NINSTR; ++i) {
      int b = rnd(K);
      int junk = JUNK400 + 32 + ((b & 1) ? 17 : 1);
      for (int l = 0; l < 32; ++l) {
        int x = rnd(20), y = x;
        if (U(rng) < 0.34) y = rnd(20);
        bool gap = (U(rng) < 0.13) || (U(rng) < 0.13);
        int ce = b * 400 + 20 * x + y;
        comb.addr.push_back(gap ? junk : ce);
        diag.addr.push_back((gap || x != y) ? junk : ce);
        off.addr.push_back((gap || x == y) ? SKIP : ce);
        if (!gap && x != y) offs += 1;
      }
    }
    off.sites_per_instr = offs / NINSTR;
    pats.push_back(comb); pats.push_back(diag); pats.push_back(off);
  }
  {  // P3: off-diagonal sites compacted per lane over 16 sites (loop until all lanes are done), single bucket
    Pat p{"off-diagonal, per-lane compaction over 16 sites (single bucket)", {}, 0, 0};
    Pat pm{"off-diagonal, per-lane compaction over 16 sites (lanes in ~7 buckets)", {}, 0, 0};
    double n_sites = 0; int n_instr = 0; double n_sites_m = 0; int n_instr_m = 0;
    while (n_instr < NINSTR) {
      int b = rnd(K);
      std::vector<std::vector<int>> lists(32), lists_m(32);
      for (int l = 0; l < 32; ++l) {
        int bm = (b + l / 5) % K;
        for (int s = 0; s < 16; ++s) {
          int x = rnd(20), y = x;
          if (U(rng) < 0.34) y = rnd(20);
          bool gap = (U(rng) < 0.13) || (U(rng) < 0.13);
          if (!gap && x != y) { lists[l].push_back(b * 400 + 20 * x + y); lists_m[l].push_back(bm * 400 + 20 * x + y); }
        }
      }
      size_t mx = 0;
      for (auto& v : lists) mx = std::max(mx, v.size());
      for (size_t it = 0; it < mx; ++it) {
        for (int l = 0; l < 32; ++l) {
          bool has = lists[l].size() > it;
          p.addr.push_back(has ? lists[l][it] : SKIP);
          pm.addr.push_back(has ? lists_m[l][it] : SKIP);
          if (has) n_sites += 1;
        }
        ++n_instr;
      }
    }
    p.sites_per_instr = pm.sites_per_instr = n_sites / n_instr;
    p.addr.resize((size_t)NINSTR * 32); pm.addr.resize((size_t)NINSTR * 32);
    pats.push_back(p); pats.push_back(pm);
  }
  {  // P4: lanes = 32 pairs (sorted by t) at the same site: few buckets
    Pat p{"lanes = 32 t-sorted pairs at one site: combined", {}, 0, 32};
    Pat pd{"lanes = 32 t-sorted pairs at one site: diagonal only", {}, 0, 32};
    for (int i = 0; i < NINSTR; ++i) {
      int g = rnd(NPAIR / 32), s = rnd(L);
      for (int l = 0; l < 32; ++l) {
        int dg = 0, ce = cell(d.order[g * 32 + l], s, 400, dg);
        p.addr.push_back(ce < 0 ? JUNK400 : ce);
        pd.addr.push_back((ce < 0 || !dg) ? JUNK400 : ce);
      }
    }
    pats.push_back(p); pats.push_back(pd);
  }
  {  // P5: uniformly random cells (worst case)
    Pat p{"uniform random cells", {}, 0, 32};
    for (int i = 0; i < NINSTR * 32; ++i) p.addr.push_back(rnd(K * 400));
    pats.push_back(p);
  }
  (void)JUNK401;
  uint32_t* dev; unsigned long long* out;
  cudaMalloc(&dev, (size_t)NINSTR * 32 * 4); cudaMalloc(&out, 8);
  const int smem = (K * 401 + 64) * 4;
  cudaFuncSetAttribute(atoms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("SMs %d, clock %d kHz\n", sms, clk);
  printf("%-72s %10s %10s %12s\n", "pattern", "clk/instr", "sites/clk", "wavefronts*");
  for (auto& p : pats) {
    // host-side wavefront model: max over banks of distinct addresses
    double wf = 0;
    for (int i = 0; i < NINSTR; ++i) {
      std::vector<uint32_t> v;
      for (int l = 0; l < 32; ++l) if (p.addr[i * 32 + l] != SKIP) v.push_back(p.addr[i * 32 + l]);
      std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end());
      int cnt[32] = {0}, m = 0;
      for (uint32_t a : v) m = std::max(m, ++cnt[a & 31]);
      wf += m;
    }
    wf /= NINSTR;
    cudaMemcpy(dev, p.addr.data(), (size_t)NINSTR * 32 * 4, cudaMemcpyHostToDevice);
    const int reps = 8;
    atoms_kernel<<<sms, 1024, smem>>>(dev, NINSTR, 1, p.mode, out);
    cudaEventRecord(e0);
    atoms_kernel<<<sms, 1024, smem>>>(dev, NINSTR, reps, p.mode, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); return 1; }
    // per SM: 32 warps * reps * NINSTR instructions
    const double instr_per_sm = 32.0 * reps * NINSTR;
    const double clk_per_instr = ms * 1e-3 * clk * 1e3 / instr_per_sm;
    printf("%-72s %10.3f %10.2f %12.3f\n", p.name, clk_per_instr, p.sites_per_instr / clk_per_instr, wf);
  }
  return 0;
}
