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
constexpr int kCoProducerWarps = 4;
constexpr int kCoThreads = (kCoConsumerWarps + kCoProducerWarps) * 32;
constexpr int kCoMaxStages = 8;
constexpr int kCoMaxPairsPerStage = 64;     // two passes of a producer warp
constexpr int kCoMaxRowBytes = 16 * 1024;   // longest supported row (8192 contacts per family)
constexpr int kCoRegionTarget = 18 * 1024;  // aim for regions of about this size ...
constexpr int kCoStageBudget = 144 * 1024;  // ... within this much shared memory for all stages
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
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes
// (or the hint expires) instead of polling -- polling warps were taking a third of the
// issue slots from the counting warps.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
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
  int nbytes;        // bytes per region in this stage
  int n_pairs;       // pairs in this stage
  int bucket_first;  // bucket of the first / last pair (equal in all but ~K stages of a launch)
  int bucket_last;
};

// Two (pair, contact) items = one 32-bit word of row a (bytes xi0 xj0 xi1 xj1) and the same
// word of row b.  An item goes to the shared histogram when at most one site changed:
// I[xi][yi][xj] (site j unchanged) at hist, J[xi][xj][yj] (site i unchanged) at histJ =
// hist + 4*S1^3.  The shared histogram has S1 = S+1 states per axis and those two cells name
// all four bytes (the unchanged site's byte stands for both rows), so an item with a skip
// byte lands in a junk cell that the flush drops: no validity test on the hot path.
// Branch-free: one PRMT gathers the contact's four bytes (xi, yi, xj, yj), one IDP.4A with
// the table's coefficient bytes gives 4*S1*mid + 4*low + base, one IMAD adds 4*S1^2*xi; an
// item with BOTH sites changed is redirected to one junk word (ATOMS.POPC.INC cannot be
// predicated; equal addresses merge), and if all four bytes are valid (6 % of the items on
// Pfam-like data) its bit is set in the returned mask: those go to L2 after the fast path
// of the whole 8-byte item, once.  w4[c] returns the gathered bytes for that slow path.
//   vmask = (0x80 - S) * 0x01010101: (byte + 0x80 - S) has bit 7 set iff byte == S.
struct CoConst {
  uint32_t vmask, coefI, coefJ, plane4, hist, histJ, junk;
};

// Returns the slow-path flags of the word's two contacts at bits 7 and 23.  The per-byte flags are formed
// for the whole word at once (round 2: the fast path is bound by issue slots, 17.5 instructions per contact;
// this form and the 16 x 8-bit dot product below save three of them):
//   nz   bit 7 of a byte set iff the two rows differ in that byte,
//   both bit 7 / 23 set iff BOTH sites of contact 0 / 1 changed (the item leaves the shared histogram),
//   slow = both and no skip byte among the contact's four bytes.
__device__ __forceinline__ uint32_t co_word(uint32_t wa, uint32_t wb, const CoConst& k, uint32_t* w4) {
  const uint32_t d = wa ^ wb;
  const uint32_t v = ((wa + k.vmask) | (wb + k.vmask)) & 0x80808080u;
  const uint32_t nz = (((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d) & 0x80808080u;
  const uint32_t both = nz & (nz >> 8);
  const uint32_t slow = both & ~(v | (v >> 8)) & 0x00800080u;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const bool eqj = (nz & (0x8000u << (16 * c))) == 0;
    const bool far = (both & (0x80u << (16 * c))) != 0;
    const uint32_t g = __byte_perm(wa, wb, c == 0 ? 0x5140u : 0x7362u);  // bytes: xi, yi, xj, yj
    w4[c] = g;
    const uint32_t lo = __dp4a(g, eqj ? k.coefI : k.coefJ, eqj ? k.hist : k.histJ);
    uint32_t addr = __dp2a_lo(k.plane4, g, lo);  // + plane4 * xi: plane4 in the low 16 bits, byte 0 of g
    if (far) addr = k.junk;
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
  }
  return slow;
}

// Both sites changed, all four bytes valid: one L2 reduction.  g = bytes (xi, yi, xj, yj);
// cell = (xi*S + xj) * S^2 + yi*S + yj.
__device__ __forceinline__ void co_slow(uint32_t g, uint32_t S, uint32_t coef_e, uint32_t* __restrict__ counts_b) {
  const uint32_t xi = g & 0xffu, xj = __byte_perm(g, 0, 0x4442u);
  const uint32_t e = __dp4a(g, coef_e, 0u);  // yi*S + yj
  atomicAdd(counts_b + (uint32_t)((xi * S + xj) * (S * S) + e), 1u);  // < S^4 <= 2^24: 32-bit offset
}
// The same with the bucket folded into a 32-bit ELEMENT index from the start of the count tensor
// (K * S^4 <= 254 * 62^4 < 2^32): one widening multiply-add forms the address, where pointer + offset took an
// add with carry and a shift with carry (the loop over the slow items is the other half of the fast path's
// instruction count).  coef_x = S^3 | S^2 << 16 for the 16 x 8-bit dot product over (xi, xj) in w = g's
// bytes 0 and 2 moved to bytes 0 and 1.
__device__ __forceinline__ void co_slow_idx(uint32_t g, uint32_t coef_x, uint32_t coef_e, uint32_t bucket_off,
                                            uint32_t* __restrict__ counts) {
  const uint32_t e = __dp4a(g, coef_e, bucket_off);                      // bucket*S^4 + yi*S + yj
  const uint32_t idx = __dp2a_lo(coef_x, __byte_perm(g, 0, 0x4420u), e);  // + xi*S^3 + xj*S^2
  atomicAdd(counts + idx, 1u);
}

template <bool SMEM>
__device__ __forceinline__ void co_flush(uint32_t* hist, int S, int tid, uint32_t* __restrict__ counts_b) {
  if (!SMEM) return;
  const int S1 = S + 1, T = S1 * S1 * S1, n = S * S;
  for (int i = tid; i < 2 * T + 1; i += kCoConsumerWarps * 32) {
    const uint32_t v = hist[i];
    if (v == 0) continue;
    hist[i] = 0;
    if (i >= 2 * T) continue;  // the junk word
    int r = i < T ? i : i - T;
    const int c = r % S1;
    r /= S1;
    const int q = r % S1, p = r / S1;
    if (c >= S || q >= S || p >= S) continue;  // junk cells (an item with a skip byte)
    // I[p=xi][q=yi][c=xj]: (xi,xj)->(yi,xj);  J[p=xi][q=xj][c=yj]: (xi,xj)->(xi,yj)
    const int s = i < T ? p * S + c : p * S + q;
    const int e = i < T ? q * S + c : p * S + c;
    atomicAdd(counts_b + (size_t)s * n + e, v);
  }
}

// A valid item of a pair whose bucket is not the one in shared memory (only in the ~K stages
// of a launch that straddle a bucket boundary), or any valid item when there is no shared
// histogram: straight to L2.
__device__ __forceinline__ void co_direct(uint32_t ha, uint32_t hb, uint32_t S, uint32_t* __restrict__ counts_b) {
  const uint32_t xi = ha & 0xffu, xj = (ha >> 8) & 0xffu;
  const uint32_t yi = hb & 0xffu, yj = (hb >> 8) & 0xffu;
  if (xi < S && xj < S && yi < S && yj < S)
    atomicAdd(counts_b + (size_t)(xi * S + xj) * (S * S) + (yi * S + yj), 1u);
}

// Persistent, warp-specialised.  CTA c owns the sorted pairs [r0, r1) = [c*n_valid/grid,
// (c+1)*n_valid/grid); stage k of the CTA is the G consecutive pairs starting at r0 + k*G
// (G = pairs per stage <= 64, chosen by the host so that G rows of the longest family fit a
// region).  Stage boundaries depend on nothing but k, so the kCoProducerWarps producer warps
// work independently: warp w fills stages w, w + P, ...: one coalesced 16-byte record load
// per lane and pass (= per pair), a warp scan of the row strides (rows are packed back to
// back: a-rows in region A, b-rows at the same offsets in region B), then 16-byte cp.async
// copies with a quad of lanes per pair (64 contiguous bytes of each row per step), handed
// to the stage's mbarrier with cp.async.mbarrier.arrive.noinc.
// Consumers see the CTA's stages as ONE stream of 256-byte chunks (32 items of 8 bytes of
// region A + the same 8 bytes of region B = 4 contacts per lane); consumer warp w takes
// chunks w, w + C, ... of that stream, walking from stage to stage on its own (wait full,
// arrive empty), so no lane idles at a stage's end beyond its last partial chunk.  No
// per-item pair lookup: a stage has one bucket and padding bytes are skip codes.  The ~K
// stages of a launch that straddle a bucket boundary look the pair up per item and send
// the items of the other buckets straight to L2.
// History: v2 issued one cp.async.bulk per row -- a 304-byte bulk copy per lane serialises
// in the uniform datapath (7 us per 39 KB stage); v3's two lock-step producer warps were
// bound by their own dependent-issue latency; v4 split every stage over all consumer
// threads (1.6 items per thread and stage: rounding up idled a fifth of the lanes).
template <bool SMEM>
__global__ void __launch_bounds__(kCoThreads, 1)
count_co_sorted_kernel(const uint8_t* __restrict__ msa, const cherry_co_rec* __restrict__ recs,
                       const int32_t* __restrict__ bucket_start, int K, int S, int G, int n_stages,
                       int n_producers, int region_bytes, uint32_t* __restrict__ counts) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) unsigned long long full_bar[kCoMaxStages];
  __shared__ __align__(8) unsigned long long empty_bar[kCoMaxStages];
  __shared__ CoStageMeta meta[kCoMaxStages];
  __shared__ int pair_end[kCoMaxStages][kCoMaxPairsPerStage];  // inclusive prefix of the row strides
  __shared__ int pair_bucket[kCoMaxStages][kCoMaxPairsPerStage];
  __shared__ int sbstart[CHERRY_MAX_BUCKETS + 2];

  const int tid = threadIdx.x;
  const int S1 = S + 1, T = S1 * S1 * S1;
  const size_t cells = (size_t)S * S * S * S;
  uint8_t* stage_base = smem;  // n_stages * 2 * region_bytes
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem + (size_t)n_stages * 2 * region_bytes);

  for (int i = tid; i <= K + 1; i += kCoThreads) sbstart[i] = bucket_start[i];
  if (SMEM)
    for (int i = tid; i < 2 * T + 1; i += kCoThreads) hist[i] = 0;
  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 2 * 32);  // one producer warp, two arrivals per thread
      mbar_init(smem_u32(&empty_bar[s]), kCoConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t n_valid = sbstart[K];
  const int r0 = (int)(n_valid * blockIdx.x / gridDim.x);
  const int r1 = (int)(n_valid * (blockIdx.x + 1) / gridDim.x);
  const int n_k = (r1 - r0 + G - 1) / G;
  const int lane = tid & 31;

  if (tid >= kCoConsumerWarps * 32) {
    // ------------------------------------------------------------ producer warps
    const int pw = (tid >> 5) - kCoConsumerWarps;
    if (pw >= n_producers) return;
    const int q = lane >> 2, c0 = (lane & 3) * 16;
    int pb = 0;
    // n_producers <= n_stages: the producer of stage k has waited for stage k - P - n_stages
    // to be consumed, which then implies stage k - 2*n_stages was -- an mbarrier parity wait
    // is only meaningful when the waiter is less than two phases ahead.
    for (int k = pw; k < n_k; k += n_producers) {
      const int s = k % n_stages;
      const int pos = r0 + k * G;
      const int n = min(G, r1 - pos);
      // this stage's records first (the empty-slot wait below then overlaps the loads) ...
      int stride[2] = {0, 0}, delta16[2] = {0, 0}, incl[2], lb[2];
      const uint8_t* ra[2] = {msa, msa};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (32 * h + lane < n) {
          const int4 rec = __ldg(reinterpret_cast<const int4*>(recs) + pos + 32 * h + lane);
          ra[h] = msa + (((int64_t)(uint32_t)rec.y << 32) | (uint32_t)rec.x);
          delta16[h] = rec.z;
          stride[h] = rec.w;
        }
      }
      // ... and the next one's into L2
      if (pos + n_producers * G + 2 * lane < r1)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(recs + pos + n_producers * G + 2 * lane));
      while (sbstart[pb + 1] <= pos) ++pb;  // bucket of the first pair
      int carry = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        lb[h] = pb;  // bucket of this lane's pair
        if (32 * h + lane < n)
          while (sbstart[lb[h] + 1] <= pos + 32 * h + lane) ++lb[h];
        int v = stride[h];  // inclusive warp scan of the strides
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, v, d);
          if (lane >= d) v += u;
        }
        incl[h] = carry + v;
        carry = __shfl_sync(0xffffffffu, incl[h], 31);
      }
      const int used = carry;
      const int last_bucket = __shfl_sync(0xffffffffu, n > 32 ? lb[1] : lb[0], (n - 1) & 31);
      if (k >= n_stages) mbar_wait(smem_u32(&empty_bar[s]), (uint32_t)(((k / n_stages) - 1) & 1));
      const uint32_t bar = smem_u32(&full_bar[s]);
      const uint32_t regA = smem_u32(stage_base + (size_t)s * 2 * region_bytes);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t off = (uint32_t)(incl[h] - stride[h]);
        const int nh = min(32, n - 32 * h);
#pragma unroll 1
        for (int g0 = 0; g0 < nh; g0 += 8) {
          const int g = g0 + q;
          const uint64_t ga = __shfl_sync(0xffffffffu, (uint64_t)ra[h], g & 31);
          const int gd = __shfl_sync(0xffffffffu, delta16[h], g & 31);
          const int gs = __shfl_sync(0xffffffffu, stride[h], g & 31);
          const uint32_t go = __shfl_sync(0xffffffffu, off, g & 31);
          if (g < nh) {
            const uint8_t* pa = reinterpret_cast<const uint8_t*>(ga) + c0;
            const uint8_t* pbrow = pa + (int64_t)gd * 16;
            uint32_t da = regA + go + c0;
            for (int c = c0; c < gs; c += 64) {
              cp_async16(da, pa);
              cp_async16(da + region_bytes, pbrow);
              da += 64; pa += 64; pbrow += 64;
            }
          }
        }
        pair_end[s][32 * h + lane] = incl[h];
        pair_bucket[s][32 * h + lane] = lb[h];
      }
      if (lane == 0) {
        meta[s].nbytes = used;
        meta[s].n_pairs = n;
        meta[s].bucket_first = pb;
        meta[s].bucket_last = last_bucket;
      }
      __syncwarp();
      cp_async_mbar_arrive_noinc(bar);  // fires when this thread's copies have landed
      mbar_arrive(bar);                 // orders the meta stores
    }
    return;
  }

  // -------------------------------------------------------------- consumer warps
  CoConst kc;
  kc.hist = smem_u32(hist);
  kc.histJ = kc.hist + 4u * T;
  kc.junk = kc.hist + 8u * T;
  kc.plane4 = 4u * S1 * S1;
  kc.coefI = ((4u * S1) << 8) | (4u << 16);   // bytes (xi, yi, xj, yj): I = [xi][yi][xj]
  kc.coefJ = ((4u * S1) << 16) | (4u << 24);  // J = [xi][xj][yj]
  kc.vmask = (0x80u - (uint32_t)S) * 0x01010101u;
  const uint32_t coef_e = ((uint32_t)S << 8) | (1u << 24);  // yi*S + yj
  // S^3 and S^2 as the two 16-bit coefficients of the (xi, xj) dot product: S <= 40; larger S keeps co_slow
  const bool wide_ok = S <= 40;
  const uint32_t coef_x = (uint32_t)(S * S * S) | ((uint32_t)(S * S) << 16);
  const int warp = tid >> 5;
  int cur_bucket = -1;
  int c = warp;  // this warp's next chunk, relative to the first chunk of stage k
  for (int k = 0; k < n_k; ++k) {
    const int s = k % n_stages;
    mbar_wait(smem_u32(&full_bar[s]), (uint32_t)((k / n_stages) & 1));
    const CoStageMeta m = meta[s];
    if (m.bucket_first != cur_bucket) {
      // every consumer warp walks every stage in order, so they all get here for stage k
      if (cur_bucket >= 0) {
        consumer_sync();  // all increments of the old bucket are done
        co_flush<SMEM>(hist, S, tid, counts + (size_t)cur_bucket * cells);
        consumer_sync();
      }
      cur_bucket = m.bucket_first;
    }
    const int n_items = m.nbytes >> 3, n_chunks = (n_items + 31) >> 5;
    const uint8_t* regA = stage_base + (size_t)s * 2 * region_bytes;
    const uint8_t* regB = regA + region_bytes;
    uint32_t* __restrict__ counts_b = counts + (size_t)cur_bucket * cells;
    const uint32_t bucket_off = (uint32_t)cur_bucket * (uint32_t)cells;  // element index (used when wide_ok)
    const bool mixed = m.bucket_first != m.bucket_last;
    for (; c < n_chunks; c += kCoConsumerWarps) {
      const int i = c * 32 + lane;
      if (i >= n_items) continue;
      const uint2 a = *reinterpret_cast<const uint2*>(regA + 8 * i);
      const uint2 b = *reinterpret_cast<const uint2*>(regB + 8 * i);
      uint32_t w4[4];
      if (SMEM && !mixed) {
        // flags of the four contacts at bits 7, 23 (word x) and 8, 24 (word y)
        uint32_t slow = co_word(a.x, b.x, kc, w4);
        slow += co_word(a.y, b.y, kc, w4 + 2) << 1;
        // lanes with both-sites-changed items loop over them (typically 0-2 of the 4): the
        // warp runs max-over-lanes iterations instead of four guarded blocks
        while (slow) {
          const uint32_t g = (slow & 0x80u) ? w4[0] : (slow & 0x100u) ? w4[2] : (slow & 0x800000u) ? w4[1] : w4[3];
          slow &= slow - 1;  // lowest flag first: bit 7, 8, 23, 24 -- the order of the selection above
          if (wide_ok) co_slow_idx(g, coef_x, coef_e, bucket_off, counts);
          else co_slow(g, S, coef_e, counts_b);
        }
      } else {
        int g = 0;  // the pair this item belongs to: first g with pair_end[g] > 8*i
        while (pair_end[s][g] <= 8 * i) ++g;
        const int bucket = pair_bucket[s][g];
        if (SMEM && bucket == cur_bucket) {
          const uint32_t sx = co_word(a.x, b.x, kc, w4), sy = co_word(a.y, b.y, kc, w4 + 2);
          if (sx & 0x80u) co_slow(w4[0], S, coef_e, counts_b);
          if (sx & 0x800000u) co_slow(w4[1], S, coef_e, counts_b);
          if (sy & 0x80u) co_slow(w4[2], S, coef_e, counts_b);
          if (sy & 0x800000u) co_slow(w4[3], S, coef_e, counts_b);
        } else {
          uint32_t* __restrict__ cb = counts + (size_t)bucket * cells;
          co_direct(a.x, b.x, S, cb);
          co_direct(a.x >> 16, b.x >> 16, S, cb);
          co_direct(a.y, b.y, S, cb);
          co_direct(a.y >> 16, b.y >> 16, S, cb);
        }
      }
    }
    c -= n_chunks;
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&empty_bar[s]));  // this warp is done with the stage
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
  if (max_row_stride > kCoMaxRowBytes)
    return cherry::fail(CHERRY_ELIMIT, "count_co: a row of %d bytes (%d contacts) exceeds the %d-byte stage",
                        max_row_stride, max_row_stride / 2, kCoMaxRowBytes);
  if (n_pairs == 0) return 0;
  // G pairs per stage (one producer-warp pass), packed rows: a region holds G rows of the
  // longest family; as many stages as fit the budget.
  int G = kCoRegionTarget / max_row_stride;
  if (G > kCoMaxPairsPerStage) G = kCoMaxPairsPerStage;
  if (G < 1) G = 1;
  const int region_bytes = G * max_row_stride;
  int n_stages = kCoStageBudget / (2 * region_bytes);
  if (n_stages > kCoMaxStages) n_stages = kCoMaxStages;
  const int n_producers = n_stages < kCoProducerWarps ? n_stages : kCoProducerWarps;
  const size_t stage_bytes = (size_t)n_stages * 2 * region_bytes;
  const size_t hist_bytes = (2 * (size_t)(S + 1) * (S + 1) * (S + 1) + 4) * sizeof(uint32_t);  // + junk word
  const bool smem_hist = stage_bytes + hist_bytes + 8192 <= (size_t)kCoSmemLimit;
  const size_t dyn = stage_bytes + (smem_hist ? hist_bytes : 0);
  static bool attr_set[64] = {false};
  int dev = 0;
  CHERRY_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    CHERRY_CUDA(cudaFuncSetAttribute(count_co_sorted_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kCoSmemLimit - 8192));
    CHERRY_CUDA(cudaFuncSetAttribute(count_co_sorted_kernel<false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kCoSmemLimit - 8192));
    attr_set[dev] = true;
  }
  const int grid = cherry::sm_count();
  if (smem_hist)
    count_co_sorted_kernel<true><<<grid, kCoThreads, dyn, (cudaStream_t)stream>>>(
        msa, recs, bucket_start, K, S, G, n_stages, n_producers, region_bytes, counts);
  else
    count_co_sorted_kernel<false><<<grid, kCoThreads, dyn, (cudaStream_t)stream>>>(
        msa, recs, bucket_start, K, S, G, n_stages, n_producers, region_bytes, counts);
  CHERRY_LAUNCH_CHECK("count_co_sorted_kernel");
  return 0;
}

}  // extern "C"
