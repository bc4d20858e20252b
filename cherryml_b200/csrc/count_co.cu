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

constexpr int kCoConsumerWarps = 24;
constexpr int kCoProducerWarps = 2;
constexpr int kCoThreads = (kCoConsumerWarps + kCoProducerWarps) * 32;
constexpr int kCoStages = 4;
constexpr int kCoRegionBytes = 16 * 1024;   // per stage: this many bytes of a-rows and of b-rows
constexpr int kCoMaxPairsPerStage = 32 * kCoProducerWarps;  // one warp pass per producer warp
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

// Scatter: besides the pair index, every sorted position gets a 16-byte record with
// everything the counting kernel's producer needs -- {byte offset of row a, (offset of row
// b - offset of row a) / 16, row stride} -- so that the producer's only global access is one
// coalesced load per lane (the indirections pair -> family -> descriptor happen here, in
// parallel over all pairs).
__global__ void co_bucket_scatter_kernel(const uint8_t* __restrict__ tab, int r_pad, int64_t n_pairs,
                                         int K, const cherry_fam_desc* __restrict__ fams,
                                         const int32_t* __restrict__ pair_a,
                                         const int32_t* __restrict__ pair_b,
                                         const int32_t* __restrict__ pair_fam,
                                         int32_t* __restrict__ ws, int32_t* __restrict__ order,
                                         cherry_co_rec* __restrict__ recs) {
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
    if (b >= 0) {
      const int dst = base[b] + rank;
      order[dst] = (int32_t)p;
      if (recs) {
        const cherry_fam_desc* fd = fams + pair_fam[p];
        const int64_t stride = fd->row_stride;
        const int64_t oa = fd->msa_off + pair_a[p] * stride, ob = fd->msa_off + pair_b[p] * stride;
        cherry_co_rec r;
        r.off_a = oa;
        r.delta_b16 = (int32_t)((ob - oa) / 16);
        r.stride = (int32_t)stride;
        recs[dst] = r;
      }
    }
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// 16-byte asynchronous copy global -> shared (LDGSTS), bypassing L1.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// The mbarrier gets one arrival (already counted at init: .noinc) once all cp.async issued
// by this thread so far have completed.
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_sync() {  // named barrier 1: the consumer warps only
  asm volatile("bar.sync 1, %0;" ::"n"(kCoConsumerWarps * 32) : "memory");
}

struct CoStageMeta {
  int nbytes;  // bytes per region in this stage; -1 = no more work
  int bucket;
};

// One (pair, contact): (xi, xj) = bytes (0, 1) of ha, (yi, yj) = bytes (0, 1) of hb (the upper
// halves are ignored).  The shared histogram has S1 = S+1 states per axis, so a skip byte
// (== S) on a near-diagonal item lands in a junk cell that the flush drops -- no validity
// test on the hot path; only the rare both-sites-changed items test it before their L2
// reduction.  I at hist, J at hist + 4*S1^3; row4 = 4*S1, plane4 = 4*S1*S1.
template <bool SMEM>
__device__ __forceinline__ void co_contact(uint32_t ha, uint32_t hb, uint32_t S, uint32_t row4,
                                           uint32_t plane4, uint32_t hist, uint32_t histJ,
                                           uint32_t* __restrict__ counts_b) {
  const uint32_t d = ha ^ hb;
  const bool eqi = (d & 0x00ffu) == 0, eqj = (d & 0xff00u) == 0;
  if (SMEM && (eqi || eqj)) {
    // eqj: I[xi][yi][xj] -> (p, q, r) = (a0, b0, a1);  else J[xj][yj][xi] -> (a1, b1, a0)
    const uint32_t w = __byte_perm(ha, hb, eqj ? 0x4041u : 0x4150u);  // bytes: r, q, p, (unused)
    const uint32_t lo = __dp4a(w, 0x00000004u | (row4 << 8), eqj ? hist : histJ);  // 4r + row4*q + base
    const uint32_t addr = __byte_perm(w, 0, 0x4442u) * plane4 + lo;
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
  } else {
    const uint32_t xi = ha & 0xffu, xj = (ha >> 8) & 0xffu;
    const uint32_t yi = hb & 0xffu, yj = (hb >> 8) & 0xffu;
    if (xi < S && xj < S && yi < S && yj < S) {
      const uint32_t n = S * S;
      atomicAdd(counts_b + (size_t)(xi * S + xj) * n + (yi * S + yj), 1u);
    }
  }
}

template <bool SMEM>
__device__ __forceinline__ void co_flush(uint32_t* hist, int S, int tid, uint32_t* __restrict__ counts_b) {
  if (!SMEM) return;
  const int S1 = S + 1, T = S1 * S1 * S1, n = S * S;
  for (int i = tid; i < 2 * T; i += kCoConsumerWarps * 32) {
    const uint32_t v = hist[i];
    if (v == 0) continue;
    hist[i] = 0;
    int r = i < T ? i : i - T;
    const int c = r % S1;
    r /= S1;
    const int q = r % S1, p = r / S1;
    if (c >= S || q >= S || p >= S) continue;  // junk cells (an item with a skip byte)
    // I[p=xi][q=yi][c=xj]: (xi,xj)->(yi,xj);  J[p=xj][q=yj][c=xi]: (xi,xj)->(xi,yj)
    const int s = i < T ? p * S + c : c * S + p;
    const int e = i < T ? q * S + c : c * S + q;
    atomicAdd(counts_b + (size_t)s * n + e, v);
  }
}

// Persistent, warp-specialised: CTA c owns the sorted pairs [c*n_valid/G, (c+1)*n_valid/G).
// The last two warps are producers: they walk the range, pack up to 64 pairs of ONE bucket
// into a stage (a-rows back to back in region A, b-rows at the same offsets in region B)
// with 16-byte cp.async copies, and publish {bytes, bucket}; full/empty mbarriers decouple
// them from the consumer warps by kCoStages stages.  (v2 used one cp.async.bulk per row: a
// 304-byte bulk copy per lane serialises in the uniform datapath, 7 us per 39 KB stage.)  Consumers: item = 8 bytes of region A + the
// same 8 bytes of region B = 4 contacts; no per-item pair lookup is needed because the whole
// stage has one bucket and padding bytes are skip codes.
template <bool SMEM>
__global__ void __launch_bounds__(kCoThreads, 1)
count_co_sorted_kernel(const uint8_t* __restrict__ msa, const cherry_co_rec* __restrict__ recs,
                       const int32_t* __restrict__ bucket_start, int K, int S,
                       uint32_t* __restrict__ counts) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) unsigned long long full_bar[kCoStages];
  __shared__ __align__(8) unsigned long long empty_bar[kCoStages];
  __shared__ CoStageMeta meta[kCoStages];
  __shared__ int sbstart[CHERRY_MAX_BUCKETS + 2];

  const int tid = threadIdx.x;
  const int S1 = S + 1, T = S1 * S1 * S1;
  const size_t cells = (size_t)S * S * S * S;
  uint8_t* stage_base = smem;  // kCoStages * 2 * kCoRegionBytes
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem + (size_t)kCoStages * 2 * kCoRegionBytes);

  for (int i = tid; i <= K + 1; i += kCoThreads) sbstart[i] = bucket_start[i];
  if (SMEM)
    for (int i = tid; i < 2 * T; i += kCoThreads) hist[i] = 0;
  if (tid == 0) {
    for (int s = 0; s < kCoStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 2 * kCoProducerWarps * 32);
      mbar_init(smem_u32(&empty_bar[s]), kCoConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t n_valid = sbstart[K];
  const int r0 = (int)(n_valid * blockIdx.x / gridDim.x);
  const int r1 = (int)(n_valid * (blockIdx.x + 1) / gridDim.x);

  if (tid >= kCoConsumerWarps * 32) {
    // ------------------------------------------------------------ producer warps
    // Both producer warps walk the range in lockstep (same loads, same scan, no
    // communication); warp w issues the copies of pass w of every stage.  Copies are
    // 16-byte cp.async (LDGSTS), one warp instruction per 512 bytes of a row; each thread
    // then hands its copies to the stage's mbarrier (cp.async.mbarrier.arrive.noinc).
    const int lane = tid & 31, pw = (tid >> 5) - kCoConsumerWarps;
    int pos = r0, pb = 0;
    for (int k = 0;; ++k) {
      const int s = k % kCoStages;
      if (k >= kCoStages) mbar_wait(smem_u32(&empty_bar[s]), (uint32_t)(((k / kCoStages) - 1) & 1));
      const uint32_t bar = smem_u32(&full_bar[s]);
      if (pos >= r1) {
        if (pw == 0 && lane == 0) {
          meta[s].nbytes = -1;
          meta[s].bucket = -1;
        }
        __syncwarp();
        mbar_arrive(bar);  // two arrivals per producer thread and stage, as below
        mbar_arrive(bar);
        break;
      }
      while (sbstart[pb + 1] <= pos) ++pb;  // bucket of position pos
      const int limit = min(r1, sbstart[pb + 1]);
      const uint32_t regA = smem_u32(stage_base + (size_t)s * 2 * kCoRegionBytes);
      const uint32_t regB = regA + kCoRegionBytes;
      if (pos + kCoMaxPairsPerStage + lane < r1)  // next stage's records: into L2 now
        asm volatile("prefetch.global.L2 [%0];" ::"l"(recs + pos + kCoMaxPairsPerStage + lane));
      int used = 0;
#pragma unroll 1
      for (int pass = 0; pass < kCoProducerWarps; ++pass) {
        const int idx = pos + lane;
        const bool in = idx < limit;
        int stride = 0;
        const uint8_t* ra = msa;
        int delta16 = 0;
        if (in) {
          const int4 rec = __ldg(reinterpret_cast<const int4*>(recs) + idx);
          ra = msa + (((int64_t)(uint32_t)rec.y << 32) | (uint32_t)rec.x);
          delta16 = rec.z;
          stride = rec.w;
        }
        int incl = stride;  // inclusive warp scan of the strides
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += v;
        }
        const bool take = in && (used + incl <= kCoRegionBytes);
        const int ntake = __popc(__ballot_sync(0xffffffffu, take));  // a prefix of the lanes
        if (pass == pw) {
          // a quad of lanes per pair: 64 contiguous bytes (two full sectors) of each row per
          // step, 8 pairs per sub-step; the row addresses come from the lane holding the pair
          const uint32_t off = (uint32_t)(used + incl - stride);
          const int q = lane >> 2, c0 = (lane & 3) * 16;
#pragma unroll 1
          for (int g0 = 0; g0 < ntake; g0 += 8) {
            const int g = g0 + q;
            const uint64_t ga = __shfl_sync(0xffffffffu, (uint64_t)ra, g & 31);
            const int gd = __shfl_sync(0xffffffffu, delta16, g & 31);
            const int gs = __shfl_sync(0xffffffffu, stride, g & 31);
            const uint32_t go = __shfl_sync(0xffffffffu, off, g & 31);
            if (g < ntake) {
              const uint8_t* pa = reinterpret_cast<const uint8_t*>(ga) + c0;
              const uint8_t* pbrow = pa + (int64_t)gd * 16;
              uint32_t da = regA + go + c0;
              for (int c = c0; c < gs; c += 64) {
                cp_async16(da, pa);
                cp_async16(da + kCoRegionBytes, pbrow);
                da += 64; pa += 64; pbrow += 64;
              }
            }
          }
        }
        if (ntake > 0) used += __shfl_sync(0xffffffffu, incl, ntake - 1);
        pos += ntake;
        if (ntake < 32) break;
      }
      if (pw == 0 && lane == 0) {
        meta[s].nbytes = used;
        meta[s].bucket = pb;
      }
      __syncwarp();
      cp_async_mbar_arrive_noinc(bar);  // fires when this thread's copies have landed
      mbar_arrive(bar);                 // orders the meta store; count = 2 per producer thread
    }
    return;
  }

  // -------------------------------------------------------------- consumer warps
  const uint32_t hist_s = smem_u32(hist), histJ_s = hist_s + 4u * T;
  const uint32_t row4 = 4u * S1, plane4 = 4u * S1 * S1;
  int cur_bucket = -1;
  for (int k = 0;; ++k) {
    const int s = k % kCoStages;
    mbar_wait(smem_u32(&full_bar[s]), (uint32_t)((k / kCoStages) & 1));
    const int nbytes = meta[s].nbytes, bucket = meta[s].bucket;
    if (nbytes < 0) break;
    if (bucket != cur_bucket) {
      if (cur_bucket >= 0) {
        consumer_sync();  // all increments of the old bucket are done
        co_flush<SMEM>(hist, S, tid, counts + (size_t)cur_bucket * cells);
        consumer_sync();
      }
      cur_bucket = bucket;
    }
    uint32_t* __restrict__ counts_b = counts + (size_t)bucket * cells;
    const uint8_t* regA = stage_base + (size_t)s * 2 * kCoRegionBytes;
    const uint8_t* regB = regA + kCoRegionBytes;
    const int n_items = nbytes >> 3;
    for (int i = tid; i < n_items; i += kCoConsumerWarps * 32) {
      const uint2 a = *reinterpret_cast<const uint2*>(regA + 8 * i);
      const uint2 b = *reinterpret_cast<const uint2*>(regB + 8 * i);
      co_contact<SMEM>(a.x, b.x, S, row4, plane4, hist_s, histJ_s, counts_b);
      co_contact<SMEM>(a.x >> 16, b.x >> 16, S, row4, plane4, hist_s, histJ_s, counts_b);
      co_contact<SMEM>(a.y, b.y, S, row4, plane4, hist_s, histJ_s, counts_b);
      co_contact<SMEM>(a.y >> 16, b.y >> 16, S, row4, plane4, hist_s, histJ_s, counts_b);
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(smem_u32(&empty_bar[s]));  // this warp is done with the stage
  }
  consumer_sync();
  if (cur_bucket >= 0) co_flush<SMEM>(hist, S, tid, counts + (size_t)cur_bucket * cells);
}

}  // namespace

extern "C" {

int cherry_sort_pairs_by_bucket(const uint8_t* tab, int r_pad, int64_t n_pairs, int K,
                                const cherry_fam_desc* fams, const int32_t* pair_a,
                                const int32_t* pair_b, const int32_t* pair_fam, int32_t* order,
                                cherry_co_rec* recs, int32_t* ws, void* stream) {
  if (!tab || !order || !ws) return cherry::fail(CHERRY_EINVAL, "sort_pairs_by_bucket: null pointer");
  if (recs && (!fams || !pair_a || !pair_b || !pair_fam))
    return cherry::fail(CHERRY_EINVAL, "sort_pairs_by_bucket: recs needs fams, pair_a, pair_b, pair_fam");
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
  co_bucket_scatter_kernel<<<blocks, 256, 0, st>>>(tab, r_pad, n_pairs, K, fams, pair_a, pair_b, pair_fam,
                                                  ws, order, recs);
  CHERRY_LAUNCH_CHECK("co_bucket_scatter_kernel");
  return 0;
}

int cherry_count_co(const uint8_t* msa, const cherry_co_rec* recs, const int32_t* bucket_start,
                    int64_t n_pairs, int max_row_stride, int K, int S, uint32_t* counts,
                    void* stream) {
  if (!msa || !recs || !bucket_start || !counts)
    return cherry::fail(CHERRY_EINVAL, "count_co: null pointer argument");
  if (K <= 0 || K > CHERRY_MAX_BUCKETS)
    return cherry::fail(CHERRY_ELIMIT, "count_co: K=%d outside 1..%d", K, CHERRY_MAX_BUCKETS);
  if (S <= 0 || S > 62) return cherry::fail(CHERRY_ELIMIT, "count_co: S=%d outside 1..62", S);
  if (max_row_stride <= 0 || max_row_stride % 16 != 0)
    return cherry::fail(CHERRY_EINVAL, "count_co: max_row_stride must be a positive multiple of 16");
  if (max_row_stride > kCoRegionBytes)
    return cherry::fail(CHERRY_ELIMIT, "count_co: a row of %d bytes (%d contacts) exceeds the %d-byte stage",
                        max_row_stride, max_row_stride / 2, kCoRegionBytes);
  if (n_pairs == 0) return 0;
  const size_t stage_bytes = (size_t)kCoStages * 2 * kCoRegionBytes;
  const size_t hist_bytes = 2 * (size_t)(S + 1) * (S + 1) * (S + 1) * sizeof(uint32_t);
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
        msa, recs, bucket_start, K, S, counts);
  else
    count_co_sorted_kernel<false><<<grid, kCoThreads, dyn, (cudaStream_t)stream>>>(
        msa, recs, bucket_start, K, S, counts);
  CHERRY_LAUNCH_CHECK("count_co_sorted_kernel");
  return 0;
}

}  // extern "C"
