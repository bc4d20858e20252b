// FastCherries on the device: divide-and-pair cherries by normalised Hamming distance, then
// quantised branch lengths and site-rate categories by coordinate ascent over a table of
// log transition probabilities.
//
// Replaces the reference program cherryml/phylogeny_estimation/FastCherries/
//   pairing_algorithms.cpp:15-175      (fc_pair_kernel)
//   branch_length_estimation.cpp:10-241 (fc_ble_kernel)
// run once per family by fast_cherries.cpp:224-255.  One CTA per family in both kernels.
//
// Parity contract: the pairing is integer work and is reproduced exactly, including the
// order in which cherries are emitted (post-order of the recursion: it fixes the order of the
// cherries in the output tree AND the summation order of the site-rate sums below) and the
// std::mt19937 + libstdc++ uniform_int_distribution pivot draws.  The two binary searches
// compare fp64 sums; each sum is accumulated by ONE thread strictly in the reference's order
// (sites ascending for a cherry, cherries ascending for a site), so given the same table the
// decisions are bit-identical.  T[x][y] + T[y][x] of the reference's inner loops is folded
// into the table (fp64 addition is commutative, so the folded table is exactly symmetric).
#include <cstdint>

#include "common.cuh"

namespace {

constexpr int kPairThreads = 256;
constexpr int kPairWarps = kPairThreads / 32;
constexpr int kBleThreads = 512;

// ------------------------------------------------------------------ std::mt19937 in shared memory
struct Mt {
  uint32_t s[624];
  int pos;
};

__device__ void mt_seed(Mt* m, uint32_t seed) {
  m->s[0] = seed;
  for (int i = 1; i < 624; ++i) m->s[i] = 1812433253u * (m->s[i - 1] ^ (m->s[i - 1] >> 30)) + (uint32_t)i;
  m->pos = 624;
}

__device__ uint32_t mt_next(Mt* m) {
  if (m->pos >= 624) {
    for (int i = 0; i < 624; ++i) {
      const uint32_t y = (m->s[i] & 0x80000000u) | (m->s[(i + 1) % 624] & 0x7fffffffu);
      uint32_t v = m->s[(i + 397) % 624] ^ (y >> 1);
      if (y & 1u) v ^= 0x9908b0dfu;
      m->s[i] = v;
    }
    m->pos = 0;
  }
  uint32_t y = m->s[m->pos++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// std::uniform_int_distribution<size_t>(0, n - 1) on a 32-bit engine, libstdc++ (GCC >= 11):
// Lemire's nearly-divisionless method.
__device__ uint32_t mt_uniform(Mt* m, uint32_t n) {
  unsigned long long product = (unsigned long long)mt_next(m) * n;
  uint32_t low = (uint32_t)product;
  if (low < n) {
    const uint32_t threshold = (0u - n) % n;
    while (low < threshold) {
      product = (unsigned long long)mt_next(m) * n;
      low = (uint32_t)product;
    }
  }
  return (uint32_t)(product >> 32);
}

// ------------------------------------------------------------------ distances
// One warp: (#sites where both residues are valid and differ, #sites where both are valid)
// between two rows of n_chunks 16-byte chunks.  Result valid in every lane.
__device__ __forceinline__ void warp_hamming(const uint4* __restrict__ row, const uint4* __restrict__ pivot,
                                             int n_chunks, uint32_t skip4, int lane, int* dist, int* count) {
  int inv_bits = 0, ne_bits = 0;
  for (int c = lane; c < n_chunks; c += 32) {
    const uint4 a = row[c];
    const uint4 p = pivot[c];
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
    const uint32_t pw[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const uint32_t inv = __vcmpeq4(aw[w], skip4) | __vcmpeq4(pw[w], skip4);
      const uint32_t ne = __vcmpne4(aw[w], pw[w]) & ~inv;
      inv_bits += __popc(inv);
      ne_bits += __popc(ne);
    }
  }
  inv_bits = __reduce_add_sync(0xffffffffu, inv_bits);
  ne_bits = __reduce_add_sync(0xffffffffu, ne_bits);
  *dist = ne_bits >> 3;
  *count = n_chunks * 16 - (inv_bits >> 3);
}

// pairing_algorithms.cpp:33-39: dist * -1.0 / count, 0 when no site is valid in both rows.
__device__ __forceinline__ double neg_hamming(int dist, int count) {
  return count == 0 ? 0.0 : ((double)dist * -1.0) / (double)count;
}

struct Frame {
  int start, len, nx, ux, phase;
};

struct PairShared {
  Mt mt;
  double wd[kPairWarps];
  int wi[kPairWarps];
  int wcnt[kPairWarps];
  int action, start, len, pivot, nx;
};

// Distances of the rows idx[start .. start+len) to `pivot`; optionally stored; returns (in
// every thread) the position of the first minimum (== the reference's strict '<' scan).
__device__ int distance_pass(const uint8_t* __restrict__ msa, long long msa_off, int stride, int n_chunks,
                             uint32_t skip4, const int* __restrict__ idx, int start, int len, int pivot,
                             double* store, PairShared* sh) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4* prow = reinterpret_cast<const uint4*>(msa + msa_off + (long long)pivot * stride);
  double best = 1.0;  // distances are <= 0
  int best_i = 0x7fffffff;
  for (int i = warp; i < len; i += kPairWarps) {
    const int r = idx[start + i];
    int dist, count;
    warp_hamming(reinterpret_cast<const uint4*>(msa + msa_off + (long long)r * stride), prow, n_chunks, skip4, lane,
                 &dist, &count);
    const double d = neg_hamming(dist, count);
    if (store != nullptr && lane == 0) store[start + i] = d;
    if (d < best) {
      best = d;
      best_i = i;
    }
  }
  if (lane == 0) {
    sh->wd[warp] = best;
    sh->wi[warp] = best_i;
  }
  __syncthreads();
  double b = sh->wd[0];
  int bi = sh->wi[0];
#pragma unroll
  for (int w = 1; w < kPairWarps; ++w) {
    const double d = sh->wd[w];
    const int i = sh->wi[w];
    if (d < b || (d == b && i < bi)) {
      b = d;
      bi = i;
    }
  }
  __syncthreads();
  return bi;
}

__global__ void __launch_bounds__(kPairThreads)
fc_pair_kernel(const uint8_t* __restrict__ msa, const cherry_fc_family* __restrict__ fams, int S, uint32_t seed,
               int32_t* __restrict__ pair_a, int32_t* __restrict__ pair_b, int32_t* __restrict__ unpaired,
               int* __restrict__ idx_all, int* __restrict__ idx2_all, double* __restrict__ d1_all,
               uint8_t* __restrict__ flag_all, Frame* __restrict__ frames_all) {
  __shared__ PairShared sh;
  const cherry_fc_family fam = fams[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = fam.n_seqs, stride = fam.row_stride, n_chunks = stride >> 4;
  const long long msa_off = fam.msa_off;
  // scratch of this family: one slot per sequence (+ 2 frames per family)
  int* idx = idx_all + fam.seq_off;
  int* idx2 = idx2_all + fam.seq_off;
  double* d1 = d1_all + fam.seq_off;
  uint8_t* flag = flag_all + fam.seq_off;
  Frame* stack = frames_all + fam.seq_off + 2 * (long long)blockIdx.x;
  const uint32_t skip4 = 0x01010101u * (uint32_t)S;

  for (int i = tid; i < N; i += kPairThreads) idx[i] = i;
  int sp = -1, ret = -1, n_out = 0;  // thread 0 only
  if (tid == 0) {
    mt_seed(&sh.mt, seed);
    if (N > 0) {
      sp = 0;
      stack[0] = Frame{0, N, 0, -1, 0};
    }
  }
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      sh.action = 0;
      while (sp >= 0) {
        Frame f = stack[sp];
        if (f.phase == 0) {
          if (f.len <= 2) {  // base cases, pairing_algorithms.cpp:85-94
            if (f.len == 2) {
              pair_a[fam.cherry_off + n_out] = idx[f.start];
              pair_b[fam.cherry_off + n_out] = idx[f.start + 1];
              ++n_out;
              ret = -1;
            } else {
              ret = f.len == 1 ? idx[f.start] : -1;
            }
            --sp;
            continue;
          }
          sh.action = 1;
          sh.start = f.start;
          sh.len = f.len;
          sh.pivot = idx[f.start + (int)mt_uniform(&sh.mt, (uint32_t)f.len)];
          break;
        }
        if (f.phase == 1) {  // the close-to-x half returned
          stack[sp].ux = ret;
          stack[sp].phase = 2;
          stack[sp + 1] = Frame{f.start + f.nx, f.len - f.nx, 0, -1, 0};
          ++sp;
          continue;
        }
        // both halves returned: pair the two left-over leaves, or hand one up (:156-162)
        if (f.ux >= 0 && ret >= 0) {
          pair_a[fam.cherry_off + n_out] = f.ux;
          pair_b[fam.cherry_off + n_out] = ret;
          ++n_out;
          ret = -1;
        } else if (f.ux >= 0) {
          ret = f.ux;
        }
        --sp;
      }
    }
    __syncthreads();
    if (sh.action == 0) break;
    const int start = sh.start, len = sh.len;
    // x = the sequence farthest from the random pivot; y = the one farthest from x (:99-113)
    int bi = distance_pass(msa, msa_off, stride, n_chunks, skip4, idx, start, len, sh.pivot, nullptr, &sh);
    const int x = idx[start + bi];
    bi = distance_pass(msa, msa_off, stride, n_chunks, skip4, idx, start, len, x, d1, &sh);
    const int y = idx[start + bi];
    // closer to x than to y (ties to x), y itself always on the y side (:62-77, :124-131)
    {
      const uint4* prow = reinterpret_cast<const uint4*>(msa + msa_off + (long long)y * stride);
      for (int i = warp; i < len; i += kPairWarps) {
        const int r = idx[start + i];
        int dist, count;
        warp_hamming(reinterpret_cast<const uint4*>(msa + msa_off + (long long)r * stride), prow, n_chunks, skip4,
                     lane, &dist, &count);
        if (lane == 0) flag[start + i] = (d1[start + i] >= neg_hamming(dist, count)) && r != y;
      }
    }
    __syncthreads();
    // stable partition of idx[start .. start+len) into the x side then the y side
    int mine = 0;
    for (int i = tid; i < len; i += kPairThreads) mine += flag[start + i];
    mine = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0) sh.wcnt[warp] = mine;
    __syncthreads();
    int nx = 0;
#pragma unroll
    for (int w = 0; w < kPairWarps; ++w) nx += sh.wcnt[w];
    __syncthreads();
    int placed_x = 0, placed_y = 0;
    for (int base = 0; base < len; base += kPairThreads) {
      const int i = base + tid;
      const bool in = i < len;
      const bool fx = in && flag[start + i];
      const unsigned bx = __ballot_sync(0xffffffffu, fx);
      const unsigned bin = __ballot_sync(0xffffffffu, in);
      if (lane == 0) sh.wcnt[warp] = __popc(bx) | (__popc(bin) << 16);
      __syncthreads();
      int before_x = 0, before_in = 0, tot_x = 0, tot_in = 0;
#pragma unroll
      for (int w = 0; w < kPairWarps; ++w) {
        const int v = sh.wcnt[w];
        if (w < warp) {
          before_x += v & 0xffff;
          before_in += v >> 16;
        }
        tot_x += v & 0xffff;
        tot_in += v >> 16;
      }
      const unsigned lt = (1u << lane) - 1u;
      const int rank_x = before_x + __popc(bx & lt);
      const int rank_in = before_in + __popc(bin & lt);
      if (in) {
        const int dst = fx ? start + placed_x + rank_x : start + nx + placed_y + (rank_in - rank_x);
        idx2[dst] = idx[start + i];
      }
      placed_x += tot_x;
      placed_y += tot_in - tot_x;
      __syncthreads();
    }
    for (int i = tid; i < len; i += kPairThreads) idx[start + i] = idx2[start + i];
    if (tid == 0) {
      stack[sp].nx = nx;
      stack[sp].phase = 1;
      stack[sp + 1] = Frame{start, nx, 0, -1, 0};
      ++sp;
    }
  }
  if (tid == 0) unpaired[blockIdx.x] = ret;
}

// ------------------------------------------------------------------ branch lengths / site rates

struct BleArgs {
  const uint8_t* msa;
  const cherry_fc_family* fams;
  const int32_t* pair_a;
  const int32_t* pair_b;
  const double2* pair_k;  // [K-1][R][S][S]: {sym[k], sym[k+1]} -- one 16-byte gather per site
  const double2* pair_r;  // [K][R-1][S][S]: {sym[.][r], sym[.][r+1]}
  const double* priors;   // [R]
  const double* weights;  // [R] cumulative weights of the initial gamma bins
  int32_t* len_idx;       // [cherries]
  int32_t* site_cat;      // [sites]
  int32_t* iters;         // [families]
  long long* total;       // scratch [sites]
  int32_t* rank;          // scratch [sites]
  int32_t* cat_by_rank;   // scratch [sites]
  uint8_t* cat8;          // scratch: site categories as bytes, 16-byte aligned per family (fc_cat8_offset)
  int S, K, R, max_iters;
};

__device__ __forceinline__ int byte_of(const uint4& v, int b) {
  const uint32_t w = b < 4 ? v.x : b < 8 ? v.y : b < 12 ? v.z : v.w;
  return (w >> (8 * (b & 3))) & 0xff;
}

// Byte-sized copy of a family's site categories: 16-byte aligned, row_stride entries.
__device__ __host__ __forceinline__ long long fc_cat8_offset(const cherry_fc_family& fam, int f) {
  return (((long long)fam.site_off + 15) & ~15LL) + 32LL * f;
}

// get_branch_lengths, branch_length_estimation.cpp:64-108, one thread per cherry.  Returns
// whether any length of this thread changed (compare != 0).  The sums run over the sites in
// ascending order; a site that is invalid in either row adds +0.0 (which leaves an fp64 sum
// unchanged) instead of being skipped, so eight table gathers can be in flight at once.
__device__ int ble_lengths(const BleArgs& a, const cherry_fc_family& fam, bool compare) {
  const int n_cherries = fam.n_seqs >> 1, n_chunks = fam.row_stride >> 4;
  const int S = a.S, SS = a.S * a.S, RSS = a.R * SS;
  const uint4* cat16 = reinterpret_cast<const uint4*>(a.cat8 + fc_cat8_offset(fam, blockIdx.x));
  int changed = 0;
  for (int c = threadIdx.x; c < n_cherries; c += blockDim.x) {
    const uint4* ra =
        reinterpret_cast<const uint4*>(a.msa + fam.msa_off + (long long)a.pair_a[fam.cherry_off + c] * fam.row_stride);
    const uint4* rb =
        reinterpret_cast<const uint4*>(a.msa + fam.msa_off + (long long)a.pair_b[fam.cherry_off + c] * fam.row_stride);
    int low = 0, high = a.K - 1;
    while (low < high) {
      const int mid = low + (high - low) / 2;
      const double2* t0 = a.pair_k + (long long)mid * RSS;
      double ll_m = 0.0, ll_m1 = 0.0;
      for (int ch = 0; ch < n_chunks; ++ch) {
        const uint4 va = ra[ch], vb = rb[ch], vc = cat16[ch];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          double2 v[8];
          bool ok[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int b = half * 8 + k;
            const int x = byte_of(va, b), y = byte_of(vb, b);
            ok[k] = x != S && y != S;
            v[k] = t0[ok[k] ? byte_of(vc, b) * SS + x * S + y : 0];
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            ll_m += ok[k] ? v[k].x : 0.0;
            ll_m1 += ok[k] ? v[k].y : 0.0;
          }
        }
      }
      if (ll_m > ll_m1) {
        high = mid;
      } else {
        low = mid + 1;
      }
    }
    if (compare && a.len_idx[fam.cherry_off + c] != low) changed = 1;
    a.len_idx[fam.cherry_off + c] = low;
  }
  return changed;
}

// get_site_rates, branch_length_estimation.cpp:110-148, one thread per site; cherries in
// ascending order, four in flight.
__device__ void ble_rates(const BleArgs& a, const cherry_fc_family& fam) {
  const int n_cherries = fam.n_seqs >> 1;
  const int S = a.S, SS = a.S * a.S, RSS = (a.R - 1) * SS;
  const uint8_t* base = a.msa + fam.msa_off;
  const int32_t* pa = a.pair_a + fam.cherry_off;
  const int32_t* pb = a.pair_b + fam.cherry_off;
  const int32_t* li = a.len_idx + fam.cherry_off;
  uint8_t* cat8 = a.cat8 + fc_cat8_offset(fam, blockIdx.x);
  for (int j = threadIdx.x; j < fam.n_sites; j += blockDim.x) {
    int low = 0, high = a.R - 1;
    while (low < high) {
      const int mid = low + (high - low) / 2;
      const double2* t0 = a.pair_r + (long long)mid * SS;
      double ll_m = a.priors[mid], ll_m1 = a.priors[mid + 1];
      int c = 0;
      for (; c + 4 <= n_cherries; c += 4) {
        int x[4], y[4], l[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          x[u] = base[(long long)pa[c + u] * fam.row_stride + j];
          y[u] = base[(long long)pb[c + u] * fam.row_stride + j];
          l[u] = li[c + u];
        }
        double2 v[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ok[u] = x[u] != S && y[u] != S;
          v[u] = t0[ok[u] ? (long long)l[u] * RSS + x[u] * S + y[u] : 0];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ll_m += ok[u] ? v[u].x : 0.0;
          ll_m1 += ok[u] ? v[u].y : 0.0;
        }
      }
      for (; c < n_cherries; ++c) {
        const int x = base[(long long)pa[c] * fam.row_stride + j];
        const int y = base[(long long)pb[c] * fam.row_stride + j];
        if (x != S && y != S) {
          const double2 v = t0[(long long)li[c] * RSS + x * S + y];
          ll_m += v.x;
          ll_m1 += v.y;
        }
      }
      if (ll_m > ll_m1) {
        high = mid;
      } else {
        low = mid + 1;
      }
    }
    a.site_cat[fam.site_off + j] = low;
    cat8[j] = (uint8_t)low;
  }
}

// pair_k[k][r][c] = {sym[k][r][c], sym[k+1][r][c]}, pair_r[k][r][c] = {sym[k][r][c], sym[k][r+1][c]}
__global__ void fc_pair_tables_kernel(const double* __restrict__ sym, int K, int R, int SS, double2* pair_k,
                                      double2* pair_r) {
  const long long n = (long long)K * R * SS;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % SS);
    const int r = (int)((i / SS) % R);
    const int k = (int)(i / ((long long)SS * R));
    const double v = sym[i];
    if (k + 1 < K) pair_k[i] = make_double2(v, sym[i + (long long)R * SS]);
    if (r + 1 < R) pair_r[((long long)k * (R - 1) + r) * SS + c] = make_double2(v, sym[i + SS]);
  }
}

__global__ void __launch_bounds__(kBleThreads) fc_ble_kernel(BleArgs a) {
  extern __shared__ unsigned short cnt[];  // [(S + 1)][blockDim.x]
  const cherry_fc_family fam = a.fams[blockIdx.x];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = fam.n_seqs, L = fam.n_sites, S = a.S;
  if (N < 2 || L == 0) {
    if (tid == 0) a.iters[blockIdx.x] = 0;
    return;
  }
  long long* total = a.total + fam.site_off;
  int32_t* rank = a.rank + fam.site_off;
  int32_t* cat_by_rank = a.cat_by_rank + fam.site_off;
  int32_t* site_cat = a.site_cat + fam.site_off;
  // initial categories from the per-site diversity (branch_length_estimation.cpp:10-62)
  for (int j = tid; j < L; j += nt) {
    for (int k = 0; k <= S; ++k) cnt[k * nt + tid] = 0;
    const uint8_t* col = a.msa + fam.msa_off + j;
    for (int i = 0; i < N; ++i) cnt[col[(long long)i * fam.row_stride] * nt + tid] += 1;
    long long non_missing = 0, tot = 0;
    for (int k = 0; k < S; ++k) non_missing += cnt[k * nt + tid];
    for (int k = 0; k < S; ++k) tot += (non_missing - cnt[k * nt + tid]) * (long long)cnt[k * nt + tid];
    total[j] = tot;
  }
  __syncthreads();
  for (int j = tid; j < L; j += nt) {  // position of (total, site) in ascending order
    const long long tj = total[j];
    int r = 0;
    for (int i = 0; i < L; ++i) {
      const long long ti = total[i];
      r += (ti < tj) || (ti == tj && i < j);
    }
    rank[j] = r;
  }
  if (tid == 0) {
    int rc = 0;
    for (int i = 0; i < L; ++i) {
      rc += (double)i >= (double)(int)round(a.weights[rc] * (double)L);
      cat_by_rank[i] = rc;
    }
  }
  __syncthreads();
  uint8_t* cat8 = a.cat8 + fc_cat8_offset(fam, blockIdx.x);
  for (int j = tid; j < fam.row_stride; j += nt) {
    const int cat = j < L ? cat_by_rank[rank[j]] : 0;
    if (j < L) site_cat[j] = cat;
    cat8[j] = (uint8_t)cat;
  }
  __syncthreads();
  // coordinate ascent (branch_length_estimation.cpp:186-227)
  ble_lengths(a, fam, false);
  __syncthreads();
  int iters = 0, budget = a.max_iters;
  bool match = false;
  while (!match && budget) {
    ble_rates(a, fam);
    __syncthreads();
    const int changed = ble_lengths(a, fam, true);
    match = __syncthreads_or(changed) == 0;
    --budget;
    ++iters;
  }
  if (tid == 0) a.iters[blockIdx.x] = iters;
}

// FastCherries layout (all sequences in file order, natural columns) -> LG counting layout (rows
// in cherry order, partners adjacent; columns at dest[site]: sorted by site-rate category, every
// category padded to 4 sites; the output is pre-filled with the skip code).  One CTA per family.
__global__ void fc_relayout_lg_kernel(const uint8_t* __restrict__ msa_in, const cherry_fc_family* __restrict__ fc_fams,
                                      const cherry_fam_desc* __restrict__ out_fams,
                                      const int32_t* __restrict__ pair_a, const int32_t* __restrict__ pair_b,
                                      const int32_t* __restrict__ dest, uint8_t* __restrict__ msa_out) {
  const cherry_fc_family in = fc_fams[blockIdx.x];
  const cherry_fam_desc out = out_fams[blockIdx.x];
  const int n_cherries = in.n_seqs >> 1, L = in.n_sites;
  const int32_t* d = dest + in.site_off;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  for (int row = warp; row < 2 * n_cherries; row += n_warps) {
    const int c = row >> 1;
    const int src_row = (row & 1) ? pair_b[in.cherry_off + c] : pair_a[in.cherry_off + c];
    const uint8_t* src = msa_in + in.msa_off + (long long)src_row * in.row_stride;
    uint8_t* dst = msa_out + out.msa_off + (long long)row * out.row_stride;
    for (int j = lane; j < L; j += 32) dst[d[j]] = src[j];
  }
}

constexpr size_t kPairBytesPerSeq = 4 + 4 + 8 + 1 + sizeof(Frame);  // idx, idx2, d1, flag, frame
constexpr size_t kBleBytesPerSite = 8 + 4 + 4;

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

size_t cherry_fc_scratch_bytes(int64_t total_seqs, int64_t total_sites, int n_fams, int K, int R, int S) {
  const size_t seqs = (size_t)total_seqs + 2 * (size_t)n_fams + 16;
  const size_t pair = 256 + align256(seqs * 4) * 2 + align256(seqs * 8) + align256(seqs) + align256(seqs * sizeof(Frame));
  const size_t ble = 256 + align256((size_t)total_sites * 8) + 2 * align256((size_t)total_sites * 4) +
                     align256((size_t)total_sites + 32 * (size_t)n_fams + 64) +
                     2 * align256((size_t)K * R * S * S * sizeof(double2));
  return pair > ble ? pair : ble;
}

int cherry_fc_pair(const uint8_t* msa, const cherry_fc_family* fams, int n_fams, int64_t total_seqs, int S,
                   uint32_t seed, int32_t* pair_a, int32_t* pair_b, int32_t* unpaired, void* scratch,
                   size_t scratch_bytes, void* stream) {
  if (n_fams == 0) return CHERRY_OK;
  if (!msa || !fams || !pair_a || !pair_b || !unpaired || !scratch) return cherry::fail(CHERRY_EINVAL, "null pointer");
  if (S < 1 || S > 254) return cherry::fail(CHERRY_ELIMIT, "S=%d out of range", S);
  if (scratch_bytes < cherry_fc_scratch_bytes(total_seqs, 0, n_fams, 0, 0, 0))
    return cherry::fail(CHERRY_EINVAL, "scratch too small");
  const size_t seqs = (size_t)total_seqs + 2 * (size_t)n_fams + 16;
  char* p = reinterpret_cast<char*>(scratch);
  p = reinterpret_cast<char*>(align256(reinterpret_cast<size_t>(p)));
  int* idx = reinterpret_cast<int*>(p);
  p += align256(seqs * 4);
  int* idx2 = reinterpret_cast<int*>(p);
  p += align256(seqs * 4);
  double* d1 = reinterpret_cast<double*>(p);
  p += align256(seqs * 8);
  uint8_t* flag = reinterpret_cast<uint8_t*>(p);
  p += align256(seqs);
  Frame* frames = reinterpret_cast<Frame*>(p);
  fc_pair_kernel<<<n_fams, kPairThreads, 0, (cudaStream_t)stream>>>(msa, fams, S, seed, pair_a, pair_b, unpaired, idx,
                                                                     idx2, d1, flag, frames);
  CHERRY_LAUNCH_CHECK("fc_pair_kernel");
  return CHERRY_OK;
}

int cherry_fc_relayout_lg(const uint8_t* msa_in, const cherry_fc_family* fc_fams, const cherry_fam_desc* out_fams,
                          int n_fams, const int32_t* pair_a, const int32_t* pair_b, const int32_t* dest,
                          uint8_t* msa_out, void* stream) {
  if (n_fams == 0) return CHERRY_OK;
  if (!msa_in || !fc_fams || !out_fams || !pair_a || !pair_b || !dest || !msa_out)
    return cherry::fail(CHERRY_EINVAL, "null pointer");
  fc_relayout_lg_kernel<<<n_fams, 256, 0, (cudaStream_t)stream>>>(msa_in, fc_fams, out_fams, pair_a, pair_b, dest,
                                                                  msa_out);
  CHERRY_LAUNCH_CHECK("fc_relayout_lg_kernel");
  return CHERRY_OK;
}

int cherry_fc_ble(const uint8_t* msa, const cherry_fc_family* fams, int n_fams, int64_t total_sites, int S,
                  const int32_t* pair_a, const int32_t* pair_b, const double* sym_table, int K, int R,
                  const double* priors, const double* init_weights, int max_iters, int32_t* len_idx,
                  int32_t* site_cat, int32_t* iters, void* scratch, size_t scratch_bytes, void* stream) {
  if (n_fams == 0) return CHERRY_OK;
  if (!msa || !fams || !pair_a || !pair_b || !sym_table || !priors || !init_weights || !len_idx || !site_cat ||
      !iters || !scratch)
    return cherry::fail(CHERRY_EINVAL, "null pointer");
  if (S < 1 || S > 32) return cherry::fail(CHERRY_ELIMIT, "FastCherries supports up to 32 states, got %d", S);
  if (K < 1 || R < 1 || R > 255 || max_iters < 0) return cherry::fail(CHERRY_EINVAL, "bad K/R/max_iters");
  if (scratch_bytes < cherry_fc_scratch_bytes(0, total_sites, n_fams, K, R, S))
    return cherry::fail(CHERRY_EINVAL, "scratch too small");
  char* p = reinterpret_cast<char*>(align256(reinterpret_cast<size_t>(scratch)));
  BleArgs a;
  a.msa = msa;
  a.fams = fams;
  a.pair_a = pair_a;
  a.pair_b = pair_b;
  a.priors = priors;
  a.weights = init_weights;
  a.len_idx = len_idx;
  a.site_cat = site_cat;
  a.iters = iters;
  a.total = reinterpret_cast<long long*>(p);
  p += align256((size_t)total_sites * 8);
  a.rank = reinterpret_cast<int32_t*>(p);
  p += align256((size_t)total_sites * 4);
  a.cat_by_rank = reinterpret_cast<int32_t*>(p);
  p += align256((size_t)total_sites * 4);
  a.cat8 = reinterpret_cast<uint8_t*>(p);
  p += align256((size_t)total_sites + 32 * (size_t)n_fams + 64);
  double2* pair_k = reinterpret_cast<double2*>(p);
  p += align256((size_t)K * R * S * S * sizeof(double2));
  double2* pair_r = reinterpret_cast<double2*>(p);
  a.pair_k = pair_k;
  a.pair_r = pair_r;
  fc_pair_tables_kernel<<<2 * cherry::sm_count(), 256, 0, (cudaStream_t)stream>>>(sym_table, K, R, S * S, pair_k,
                                                                                  pair_r);
  CHERRY_LAUNCH_CHECK("fc_pair_tables_kernel");
  a.S = S;
  a.K = K;
  a.R = R;
  a.max_iters = max_iters;
  const size_t smem = (size_t)(S + 1) * kBleThreads * sizeof(unsigned short);
  fc_ble_kernel<<<n_fams, kBleThreads, smem, (cudaStream_t)stream>>>(a);
  CHERRY_LAUNCH_CHECK("fc_ble_kernel");
  return CHERRY_OK;
}

}  // extern "C"
