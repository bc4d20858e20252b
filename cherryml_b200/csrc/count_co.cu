// Co-transition counting for sm_100a: counting sort of the pairs by time bucket, then a
// persistent streaming kernel with a per-CTA shared-memory histogram of the "near-diagonal"
// cells of ONE bucket at a time.
//
// Reference semantics (songlab-cal/CherryML v0.2.0): the per-contact loop of
// counting/_count_co_transitions.cpp:358-383 (cherry++), :469-531 (cherry / edge) ==
// counting/_count_co_transitions.py:96-224: for a pair of sequences (a, b) at distance t and
// a contacting site pair (i, j): bucket = quantization_idx(t); if all four residues are in
// the alphabet, the cell [(a_i, a_j) -> (b_i, b_j)] gets one count (the 0.25 / 0.5 weights
// and the mirrored cells are applied afterwards by cherry_symmetrize_co, exactly).
//
// Layout: a family's rows are stored CONTACT-PAIRED -- bytes 2c and
// 2c+1 of a row are the residues at sites i_c and j_c of contact c, padded with the skip
// code S to a multiple of 16 bytes -- so the kernel never gathers: the item stream is two
// rows read front to back, 4 bytes per (pair, contact).
//
// Why sort by bucket: the histogram of one bucket has S^4 = 160 000 cells (640 KB), K of
// them never fit in shared memory, and global reductions on the hot cells serialise in L2.
// A cherry has ONE bucket, so after a counting sort of the pairs a CTA walks a contiguous
// run of pairs of the same bucket and only needs one bucket's hot cells on chip.  Hot cells:
// transitions in which at most one of the two sites changed (measured 89 % of the valid
// items on Pfam-like data): I[xi][yi][xj] (site j unchanged; includes the diagonal) and
// J[xj][yj][xi] (site i unchanged), 2*S^3 = 16 000 uint32 = 64 KB.  The remaining items
// (both sites changed) are spread over 144 400 cells per bucket and go to L2 reductions.
#include "common.cuh"

namespace {

constexpr int kCoThreads = 512;
constexpr int kCoStages = 3;
constexpr int kCoRegionBytes = 20 * 1024;   // per stage: this many bytes of a-rows and of b-rows
constexpr int kCoMaxPairsPerStage = 64;     // two warp passes of the producer
constexpr int kCoSmemLimit = 227 * 1024;

// ------------------------------------------------------------------ counting sort by bucket
// ws[0..K]      : bucket_start (exclusive prefix of the bucket sizes), ws[K+1] = n_pairs;
//                 bucket K collects the pairs outside the grid
// ws[K+2..2K+3] : running cursors of the scatter pass (scratch)
__global__ void co_bucket_hist_kernel(const uint8_t* __restrict__ tab, int r_pad, int64_t n_pairs,
                                      int K, int32_t* __restrict__ ws) {
  __shared__ int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pairs;
       p += (int64_t)gridDim.x * blockDim.x) {
    int b = tab[p * r_pad];
    if (b >= K) b = K;
    atomicAdd(&sh[b], 1);
  }
  __syncthreads();
  if ((int)threadIdx.x <= K && sh[threadIdx.x]) atomicAdd(&ws[K + 2 + threadIdx.x], sh[threadIdx.x]);
}

__global__ void co_bucket_scan_kernel(int K, int32_t* __restrict__ ws) {
  // one thread: K <= 254
  int acc = 0;
  for (int b = 0; b <= K; ++b) {
    const int c = ws[K + 2 + b];
    ws[b] = acc;
    ws[K + 2 + b] = acc;
    acc += c;
  }
  ws[K + 1] = acc;
}

__global__ void co_bucket_scatter_kernel(const uint8_t* __restrict__ tab, int r_pad, int64_t n_pairs,
                                         int K, int32_t* __restrict__ ws, int32_t* __restrict__ order) {
  __shared__ int cnt[256];
  __shared__ int base[256];
  const int64_t per_block = (n_pairs + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = blockIdx.x * per_block;
  const int64_t p1 = (p0 + per_block < n_pairs) ? p0 + per_block : n_pairs;
  // chunks of blockDim pairs keep the output of a block roughly in input order
  for (int64_t c0 = p0; c0 < p1; c0 += blockDim.x) {
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t p = c0 + threadIdx.x;
    int b = -1, rank = 0;
    if (p < p1) {
      b = tab[p * r_pad];
      if (b >= K) b = K;
      rank = atomicAdd(&cnt[b], 1);
    }
    __syncthreads();
    if ((int)threadIdx.x <= K && cnt[threadIdx.x])
      base[threadIdx.x] = atomicAdd(&ws[K + 2 + threadIdx.x], cnt[threadIdx.x]);
    __syncthreads();
    if (b >= 0) order[base[b] + rank] = (int32_t)p;
    __syncthreads();
  }
}

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map); 16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct CoStageMeta {
  int nbytes;  // bytes per region in this stage; -1 = no more work
  int bucket;
};

// One (pair, contact): bytes (xi, xj) of row a and (yi, yj) of row b in the low 16 bits of
// ha / hb.  `hist` = shared histogram (I then J), counts_b = this bucket's global cells.
template <bool SMEM>
__device__ __forceinline__ void co_contact(uint32_t ha, uint32_t hb, uint32_t S, uint32_t S3,
                                           uint32_t hist, uint32_t* __restrict__ counts_b) {
  const uint32_t xi = ha & 0xffu, xj = (ha >> 8) & 0xffu;
  const uint32_t yi = hb & 0xffu, yj = (hb >> 8) & 0xffu;
  if (xi >= S || xj >= S || yi >= S || yj >= S) return;
  if (SMEM && xj == yj) {
    const uint32_t idx = (xi * S + yi) * S + xj;
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist + idx * 4u) : "memory");
  } else if (SMEM && xi == yi) {
    const uint32_t idx = S3 + (xj * S + yj) * S + xi;
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist + idx * 4u) : "memory");
  } else {
    const uint32_t n = S * S;
    atomicAdd(counts_b + (size_t)(xi * S + xj) * n + (yi * S + yj), 1u);
  }
}

template <bool SMEM>
__device__ __forceinline__ void co_flush(uint32_t* hist, int S, uint32_t* __restrict__ counts_b) {
  if (!SMEM) return;
  const int S3 = S * S * S, n = S * S;
  for (int i = threadIdx.x; i < 2 * S3; i += kCoThreads) {
    const uint32_t v = hist[i];
    if (v == 0) continue;
    hist[i] = 0;
    int r = i < S3 ? i : i - S3;
    const int c = r % S;
    r /= S;
    const int q = r % S, p = r / S;
    // I[p=xi][q=yi][c=xj]: (xi,xj)->(yi,xj);  J[p=xj][q=yj][c=xi]: (xi,xj)->(xi,yj)
    const int s = i < S3 ? p * S + c : c * S + p;
    const int e = i < S3 ? q * S + c : c * S + q;
    atomicAdd(counts_b + (size_t)s * n + e, v);
  }
}

// Persistent kernel: CTA c owns the sorted pairs [c*n_valid/G, (c+1)*n_valid/G).  Warp 0
// is also the producer: it walks its range, packs up to 64 pairs of ONE bucket into a stage
// (a-rows back to back in region A, b-rows at the same offsets in region B) with one bulk
// copy per row, and publishes {bytes, bucket}.  All warps consume: item = 8 bytes of region
// A + the same 8 bytes of region B = 4 contacts; no per-item pair lookup is needed because
// the whole stage has one bucket and the padding bytes are skip codes.
template <bool SMEM>
__global__ void __launch_bounds__(kCoThreads, 1)
count_co_sorted_kernel(const uint8_t* __restrict__ msa, const cherry_fam_desc* __restrict__ fams,
                       const int32_t* __restrict__ pair_a, const int32_t* __restrict__ pair_b,
                       const int32_t* __restrict__ pair_fam, const int32_t* __restrict__ order,
                       const int32_t* __restrict__ bucket_start, int K, int S,
                       uint32_t* __restrict__ counts) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[kCoStages];
  __shared__ CoStageMeta meta[kCoStages];
  __shared__ int sbstart[CHERRY_MAX_BUCKETS + 2];

  const int tid = threadIdx.x;
  const int S3 = S * S * S;
  const size_t cells = (size_t)S * S * S * S;
  uint8_t* stage_base = smem;  // kCoStages * 2 * kCoRegionBytes
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem + (size_t)kCoStages * 2 * kCoRegionBytes);
  const uint32_t hist_s = smem_u32(hist);

  for (int i = tid; i <= K + 1; i += kCoThreads) sbstart[i] = bucket_start[i];
  if (SMEM)
    for (int i = tid; i < 2 * S3; i += kCoThreads) hist[i] = 0;
  if (tid == 0) {
    for (int s = 0; s < kCoStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t n_valid = sbstart[K];
  const int r0 = (int)(n_valid * blockIdx.x / gridDim.x);
  const int r1 = (int)(n_valid * (blockIdx.x + 1) / gridDim.x);

  // producer state (warp 0, uniform across its lanes)
  int pos = r0, pb = 0;
  auto produce = [&](int s) {
    const int lane = tid & 31;
    const uint32_t bar = smem_u32(&bars[s]);
    if (pos >= r1) {
      if (lane == 0) {
        meta[s].nbytes = -1;
        meta[s].bucket = -1;
        mbar_arrive_expect_tx(bar, 0);
      }
      return;
    }
    while (sbstart[pb + 1] <= pos) ++pb;  // bucket of position pos
    const int limit = min(r1, sbstart[pb + 1]);
    const uint32_t regA = smem_u32(stage_base + (size_t)s * 2 * kCoRegionBytes);
    const uint32_t regB = regA + kCoRegionBytes;
    int used = 0;
#pragma unroll 1
    for (int pass = 0; pass < kCoMaxPairsPerStage / 32; ++pass) {
      const int idx = pos + lane;
      const bool in = idx < limit;
      int stride = 0;
      const uint8_t *ra = nullptr, *rb = nullptr;
      if (in) {
        const int o = __ldg(order + idx);
        const cherry_fam_desc* fd = fams + __ldg(pair_fam + o);
        stride = fd->row_stride;
        const uint8_t* base = msa + fd->msa_off;
        ra = base + (int64_t)__ldg(pair_a + o) * stride;
        rb = base + (int64_t)__ldg(pair_b + o) * stride;
      }
      int incl = stride;  // inclusive warp scan of the strides
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      const bool take = in && (used + incl <= kCoRegionBytes);
      const uint32_t tmask = __ballot_sync(0xffffffffu, take);
      const int ntake = __popc(tmask);  // a prefix of the lanes (strides are positive)
      if (take) {
        const uint32_t off = (uint32_t)(used + incl - stride);
        bulk_g2s(regA + off, ra, (uint32_t)stride, bar);
        bulk_g2s(regB + off, rb, (uint32_t)stride, bar);
      }
      const int taken_bytes = __shfl_sync(0xffffffffu, incl, ntake > 0 ? ntake - 1 : 0);
      if (ntake > 0) used += taken_bytes;
      pos += ntake;
      if (ntake < 32) break;
    }
    __syncwarp();
    if (lane == 0) {
      meta[s].nbytes = used;
      meta[s].bucket = pb;
      mbar_arrive_expect_tx(bar, 2u * (uint32_t)used);
    }
  };

  if (tid < 32)
    for (int s = 0; s < kCoStages; ++s) produce(s);

  int cur_bucket = -1;
  for (int k = 0;; ++k) {
    const int s = k % kCoStages;
    mbar_wait(smem_u32(&bars[s]), (uint32_t)((k / kCoStages) & 1));
    const int nbytes = meta[s].nbytes, bucket = meta[s].bucket;
    if (nbytes < 0) break;
    if (bucket != cur_bucket) {
      // the end-of-stage barrier below already ordered all increments of the old bucket
      if (cur_bucket >= 0) {
        co_flush<SMEM>(hist, S, counts + (size_t)cur_bucket * cells);
        __syncthreads();
      }
      cur_bucket = bucket;
    }
    uint32_t* __restrict__ counts_b = counts + (size_t)bucket * cells;
    const uint8_t* regA = stage_base + (size_t)s * 2 * kCoRegionBytes;
    const uint8_t* regB = regA + kCoRegionBytes;
    const int n_items = nbytes >> 3;
    for (int i = tid; i < n_items; i += kCoThreads) {
      const uint2 a = *reinterpret_cast<const uint2*>(regA + 8 * i);
      const uint2 b = *reinterpret_cast<const uint2*>(regB + 8 * i);
      co_contact<SMEM>(a.x, b.x, S, S3, hist_s, counts_b);
      co_contact<SMEM>(a.x >> 16, b.x >> 16, S, S3, hist_s, counts_b);
      co_contact<SMEM>(a.y, b.y, S, S3, hist_s, counts_b);
      co_contact<SMEM>(a.y >> 16, b.y >> 16, S, S3, hist_s, counts_b);
    }
    __syncthreads();  // everyone is done with this stage's buffers and meta
    if (tid < 32) produce(s);
  }
  __syncthreads();
  if (cur_bucket >= 0) co_flush<SMEM>(hist, S, counts + (size_t)cur_bucket * cells);
}

}  // namespace

extern "C" {

int cherry_sort_pairs_by_bucket(const uint8_t* tab, int r_pad, int64_t n_pairs, int K,
                                int32_t* order, int32_t* ws, void* stream) {
  if (!tab || !order || !ws) return cherry::fail(CHERRY_EINVAL, "sort_pairs_by_bucket: null pointer");
  if (K <= 0 || K > CHERRY_MAX_BUCKETS)
    return cherry::fail(CHERRY_ELIMIT, "sort_pairs_by_bucket: K=%d outside 1..%d", K, CHERRY_MAX_BUCKETS);
  if (r_pad <= 0 || n_pairs < 0 || n_pairs > 0x7fffffff)
    return cherry::fail(CHERRY_EINVAL, "sort_pairs_by_bucket: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  CHERRY_CUDA(cudaMemsetAsync(ws, 0, sizeof(int32_t) * 2 * (K + 2), st));
  if (n_pairs == 0) return 0;
  int blocks = (int)((n_pairs + 255) / 256);
  const int cap = cherry::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  co_bucket_hist_kernel<<<blocks, 256, 0, st>>>(tab, r_pad, n_pairs, K, ws);
  CHERRY_LAUNCH_CHECK("co_bucket_hist_kernel");
  co_bucket_scan_kernel<<<1, 1, 0, st>>>(K, ws);
  CHERRY_LAUNCH_CHECK("co_bucket_scan_kernel");
  co_bucket_scatter_kernel<<<blocks, 256, 0, st>>>(tab, r_pad, n_pairs, K, ws, order);
  CHERRY_LAUNCH_CHECK("co_bucket_scatter_kernel");
  return 0;
}

int cherry_count_co(const uint8_t* msa, const cherry_fam_desc* fams, const int32_t* pair_a,
                    const int32_t* pair_b, const int32_t* pair_fam, const int32_t* order,
                    const int32_t* bucket_start, int64_t n_pairs, int max_row_stride, int K, int S,
                    uint32_t* counts, void* stream) {
  if (!msa || !fams || !pair_a || !pair_b || !pair_fam || !order || !bucket_start || !counts)
    return cherry::fail(CHERRY_EINVAL, "count_co: null pointer argument");
  if (K <= 0 || K > CHERRY_MAX_BUCKETS)
    return cherry::fail(CHERRY_ELIMIT, "count_co: K=%d outside 1..%d", K, CHERRY_MAX_BUCKETS);
  if (S <= 0 || S > 64) return cherry::fail(CHERRY_ELIMIT, "count_co: S=%d outside 1..64", S);
  if (max_row_stride <= 0 || max_row_stride % 16 != 0)
    return cherry::fail(CHERRY_EINVAL, "count_co: max_row_stride must be a positive multiple of 16");
  if (max_row_stride > kCoRegionBytes)
    return cherry::fail(CHERRY_ELIMIT, "count_co: a row of %d bytes (%d contacts) exceeds the %d-byte stage",
                        max_row_stride, max_row_stride / 2, kCoRegionBytes);
  if (n_pairs == 0) return 0;
  const size_t stage_bytes = (size_t)kCoStages * 2 * kCoRegionBytes;
  const size_t hist_bytes = 2 * (size_t)S * S * S * sizeof(uint32_t);
  const bool smem_hist = stage_bytes + hist_bytes + 4096 <= (size_t)kCoSmemLimit;
  const size_t dyn = stage_bytes + (smem_hist ? hist_bytes : 0);
  static bool attr_set[64] = {false};
  int dev = 0;
  CHERRY_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    CHERRY_CUDA(cudaFuncSetAttribute(count_co_sorted_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kCoSmemLimit - 4096));
    CHERRY_CUDA(cudaFuncSetAttribute(count_co_sorted_kernel<false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kCoSmemLimit - 4096));
    attr_set[dev] = true;
  }
  const int grid = cherry::sm_count();
  if (smem_hist)
    count_co_sorted_kernel<true><<<grid, kCoThreads, dyn, (cudaStream_t)stream>>>(
        msa, fams, pair_a, pair_b, pair_fam, order, bucket_start, K, S, counts);
  else
    count_co_sorted_kernel<false><<<grid, kCoThreads, dyn, (cudaStream_t)stream>>>(
        msa, fams, pair_a, pair_b, pair_fam, order, bucket_start, K, S, counts);
  CHERRY_LAUNCH_CHECK("count_co_sorted_kernel");
  return 0;
}

}  // extern "C"
