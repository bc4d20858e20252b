// Tree log-likelihood by Felsenstein pruning in log space, one launch per (tree, model).
//
// Replaces the dynamic programme of dp_likelihood_computation, reference
// cherryml/evaluation/_likelihood.py:239-326.  A "unit" is an independently evolving site
// (c = 1, Su = S states) or a pair of contacting sites (c = 2, Su = S*S states, state index
// S*i + j).  Thread (g, s) of a CTA owns state s of the CTA's g-th unit and walks the tree in
// post-order; the running sum of child messages of the open node at depth d lives in
// acc[d][g][s], which only that thread touches, so the children of a node are accumulated in
// the reference's order (tree.children order) without synchronisation.  The message of child v:
//     log(max(0, sum_s' P_v[s][s'] * exp(dp_v[s'] - max dp_v) * obs_v[s'])) + max dp_v
// with P_v = expm(branch length * site rate * Q), computed beforehand by cherry_expm_batched and
// handed over TRANSPOSED (Pt[k][s] = P[s][k], cherry_tree_ll_transpose): thread s walks column s
// of Pt, so a warp reads 32 consecutive doubles per step of the dot product instead of 32
// different rows; the sum over k keeps its order.
#include <cstdint>

#include "common.cuh"

namespace {

struct LlArgs {
  const cherry_ll_node* nodes;
  const int32_t* p_index;  // [n_nodes][n_cats]
  const double* P;         // [n_matrices][Su][Su], transposed: P[m][k][s] = expm(...)[s][k]
  const uint8_t* obs;      // [n_leaves][n_units][c]
  const int32_t* unit_cat; // [n_units]
  const double* pi;        // [Su]
  double* acc;             // [gridDim.x][max_depth + 1][G][Su]
  double* ll_out;          // [n_units]
  int n_nodes, n_cats, S, c, Su, n_units, G, max_depth, staged;
};

constexpr int kLlBatch = 16;

// one commit group: entries k0 .. k0+15 of this thread's column -> its slots of staging buffer `buf`
__device__ __forceinline__ void ll_stage_batch(uint32_t stage_u32, uint32_t slot_stride, const double* col, int Su,
                                               int k0, int buf) {
#pragma unroll
  for (int j = 0; j < kLlBatch; ++j)
    if (k0 + j < Su)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(stage_u32 + (uint32_t)(buf * kLlBatch + j) * slot_stride),
                   "l"(col + (size_t)(k0 + j) * Su)
                   : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void tree_ll_kernel(LlArgs a) {
  extern __shared__ double w[];  // [G][Su], then [G][32] partial maxima, then (Su > 64) the staging slots
  double* red = w + (size_t)a.G * a.Su;
  double* stage = red + (size_t)a.G * 32 + threadIdx.x;  // [2][kLlBatch][blockDim.x], this thread's column
  const uint32_t stage_u32 = (uint32_t)__cvta_generic_to_shared(stage);
  const uint32_t slot_stride = blockDim.x * (uint32_t)sizeof(double);
  const int Su = a.Su, S = a.S;
  const int g = threadIdx.x / Su, s = threadIdx.x - g * Su;
  const int u = blockIdx.x * a.G + g;
  const bool live = g < a.G && u < a.n_units;
  const int cat = live ? a.unit_cat[u] : 0;
  double* acc = a.acc + ((size_t)blockIdx.x * (a.max_depth + 1) * a.G + (live ? g : 0)) * Su + s;
  const size_t acc_stride = (size_t)a.G * Su;
  double* wg = w + (live ? g : 0) * Su;
  for (int i = 0; i < a.n_nodes; ++i) {
    const cherry_ll_node node = a.nodes[i];
    const bool is_leaf = node.flags & 1;
    const bool is_root = i == a.n_nodes - 1;
    // column s of the edge's matrix; the index does not depend on the running sums, so its load
    // overlaps the reduction below
    const double* col = a.P + (is_root || !live ? 0 : (size_t)a.p_index[(size_t)i * a.n_cats + cat] * Su * Su) + s;
    double m = 0.0;
    if (a.staged && live && !is_leaf && !is_root) {
      ll_stage_batch(stage_u32, slot_stride, col, Su, 0, 0);
      ll_stage_batch(stage_u32, slot_stride, col, Su, kLlBatch, 1);
    }
    if (!is_leaf) {
      // dp of this node is complete: w = exp(dp - max dp) (observation vector of an internal node = ones)
      const double v = live ? acc[(size_t)node.depth * acc_stride] : 0.0;
      if (live) wg[s] = v;
      __syncthreads();
      double mx = -INFINITY;
      if (Su > 64) {
        // two-stage maximum: 32 strided partial maxima per unit, then everyone reads those
        if (live && s < 32) {
          double r = wg[s];
          for (int k = s + 32; k < Su; k += 32) r = fmax(r, wg[k]);
          red[g * 32 + s] = r;
        }
        __syncthreads();
        if (live)
          for (int k = 0; k < 32; ++k) mx = fmax(mx, red[g * 32 + k]);
      } else {
        if (live)
          for (int k = 0; k < Su; ++k) mx = fmax(mx, wg[k]);
        __syncthreads();
      }
      if (live) wg[s] = exp(v - mx);
      __syncthreads();
      if (live) {
        double sum = 0.0;
        if (is_root) {
          if (s == 0) {
            for (int k = 0; k < Su; ++k) sum += a.pi[k] * wg[k];
            a.ll_out[u] = log(fmax(sum, 0.0)) + mx;
          }
        } else {
          if (a.staged) {
            // Two batches of 16 column entries in flight per thread through cp.async (a warp's
            // 256-byte segments come from L2 every time: the matrices of one tree do not fit in
            // L1, so the loop is bound by how many bytes are outstanding); every thread stages
            // and consumes its own slots, and the sum keeps its order in k.
            // (the first two batches were issued at the top of the node, before the reduction)
            const int nb = (Su + kLlBatch - 1) / kLlBatch;
            for (int b = 0; b < nb; ++b) {
              if (b + 1 < nb) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");
              } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
              }
              const int k0 = b * kLlBatch;
              const double* mine = stage + (size_t)(b & 1) * kLlBatch * blockDim.x;
#pragma unroll
              for (int j = 0; j < kLlBatch; ++j)
                if (k0 + j < Su) sum += mine[(size_t)j * blockDim.x] * wg[k0 + j];
              if (b + 2 < nb) ll_stage_batch(stage_u32, slot_stride, col, Su, k0 + 2 * kLlBatch, b & 1);
            }
          } else {
            for (int k = 0; k < Su; ++k) sum += col[(size_t)k * Su] * wg[k];
          }
          m = log(fmax(sum, 0.0)) + mx;
        }
      }
      __syncthreads();
    } else if (live) {
      // leaf: dp = 0, observation = one-hot, or every state compatible with the known residues
      const uint8_t* ob = a.obs + ((size_t)node.obs_row * a.n_units + u) * a.c;
      double sum = 0.0;
      if (is_root) {  // a single-node tree
        if (s == 0) {
          if (a.c == 1) {
            const int x = ob[0];
            for (int k = 0; k < Su; ++k) sum += (x == S || x == k) ? a.pi[k] : 0.0;
          } else {
            const int x = ob[0], y = ob[1];
            for (int k = 0; k < Su; ++k) sum += ((x == S || x == k / S) && (y == S || y == k % S)) ? a.pi[k] : 0.0;
          }
          a.ll_out[u] = log(fmax(sum, 0.0));
        }
      } else {
        if (a.c == 1) {
          const int x = ob[0];
          if (x != S) {
            sum = col[(size_t)x * Su];
          } else {
            for (int k = 0; k < Su; ++k) sum += col[(size_t)k * Su];
          }
        } else {
          const int x = ob[0], y = ob[1];
          if (x != S && y != S) {
            sum = col[(size_t)(x * S + y) * Su];
          } else if (x != S) {
            for (int k = 0; k < S; ++k) sum += col[(size_t)(x * S + k) * Su];
          } else if (y != S) {
            for (int k = 0; k < S; ++k) sum += col[(size_t)(k * S + y) * Su];
          } else {
            for (int k = 0; k < Su; ++k) sum += col[(size_t)k * Su];
          }
        }
        m = log(fmax(sum, 0.0));
      }
    }
    if (live && !is_root) {
      double* dst = acc + (size_t)(node.depth - 1) * acc_stride;
      *dst = (node.flags & 2) ? m : *dst + m;
    }
  }
}

// out[m][k][s] = in[m][s][k], 32 x 32 tiles through shared memory (both sides coalesced)
__global__ void transpose_matrices_kernel(const double* __restrict__ in, double* __restrict__ out, int Su) {
  __shared__ double tile[32][33];
  const size_t base = (size_t)blockIdx.z * Su * Su;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < Su && c < Su) tile[j][threadIdx.x] = in[base + (size_t)r * Su + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = c0 + j, c = r0 + threadIdx.x;
    if (r < Su && c < Su) out[base + (size_t)r * Su + c] = tile[threadIdx.x][j];
  }
}

}  // namespace

extern "C" {

int cherry_tree_ll_transpose(const double* P, int n_matrices, int Su, double* Pt, void* stream) {
  if (n_matrices == 0) return CHERRY_OK;
  if (!P || !Pt || P == Pt) return cherry::fail(CHERRY_EINVAL, "null or aliased pointer");
  if (n_matrices < 0 || Su < 1) return cherry::fail(CHERRY_EINVAL, "bad sizes");
  const int tiles = (Su + 31) / 32;
  for (int m0 = 0; m0 < n_matrices; m0 += 65535) {
    const int nm = n_matrices - m0 < 65535 ? n_matrices - m0 : 65535;
    transpose_matrices_kernel<<<dim3(tiles, tiles, nm), dim3(32, 8), 0, (cudaStream_t)stream>>>(
        P + (size_t)m0 * Su * Su, Pt + (size_t)m0 * Su * Su, Su);
  }
  CHERRY_LAUNCH_CHECK("transpose_matrices_kernel");
  return CHERRY_OK;
}

int cherry_tree_ll_units_per_block(int S, int c) {
  const int Su = c == 2 ? S * S : S;
  const int G = 256 / Su;
  return G < 1 ? 1 : G;
}

size_t cherry_tree_ll_scratch_bytes(int S, int c, int n_units, int max_depth) {
  const int Su = c == 2 ? S * S : S;
  const int G = cherry_tree_ll_units_per_block(S, c);
  const size_t blocks = (size_t)(n_units + G - 1) / G;
  return blocks * (size_t)(max_depth + 1) * G * Su * sizeof(double);
}

int cherry_tree_log_likelihood(const cherry_ll_node* nodes, int n_nodes, const int32_t* p_index, int n_cats,
                               const double* P, const uint8_t* obs, const int32_t* unit_cat, const double* pi,
                               int S, int c, int n_units, int max_depth, void* scratch, size_t scratch_bytes,
                               double* ll_out, void* stream) {
  if (n_units == 0) return CHERRY_OK;
  if (!nodes || !p_index || !P || !obs || !unit_cat || !pi || !scratch || !ll_out)
    return cherry::fail(CHERRY_EINVAL, "null pointer");
  if (n_nodes < 1 || n_cats < 1 || (c != 1 && c != 2) || S < 1 || max_depth < 0)
    return cherry::fail(CHERRY_EINVAL, "bad sizes");
  const int Su = c == 2 ? S * S : S;
  if (Su > 1024) return cherry::fail(CHERRY_ELIMIT, "more than 1024 unit states (%d)", Su);
  if (scratch_bytes < cherry_tree_ll_scratch_bytes(S, c, n_units, max_depth))
    return cherry::fail(CHERRY_EINVAL, "scratch too small");
  LlArgs a;
  a.nodes = nodes;
  a.p_index = p_index;
  a.P = P;
  a.obs = obs;
  a.unit_cat = unit_cat;
  a.pi = pi;
  a.acc = reinterpret_cast<double*>(scratch);
  a.ll_out = ll_out;
  a.n_nodes = n_nodes;
  a.n_cats = n_cats;
  a.S = S;
  a.c = c;
  a.Su = Su;
  a.n_units = n_units;
  a.G = cherry_tree_ll_units_per_block(S, c);
  a.max_depth = max_depth;
  const int threads = (a.G * Su + 31) / 32 * 32;
  const int blocks = (n_units + a.G - 1) / a.G;
  size_t smem = (size_t)a.G * (Su + 32) * sizeof(double);
  const size_t staging = (size_t)2 * kLlBatch * threads * sizeof(double);
  a.staged = Su > 64 && smem + staging <= 227 * 1024;
  if (a.staged) {
    smem += staging;
    const int rc = cherry::check_cuda(
        cudaFuncSetAttribute(tree_ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
        "cudaFuncSetAttribute(tree_ll_kernel)");
    if (rc != 0) return rc;
  }
  tree_ll_kernel<<<blocks, threads, smem, (cudaStream_t)stream>>>(a);
  CHERRY_LAUNCH_CHECK("tree_ll_kernel");
  return CHERRY_OK;
}

}  // extern "C"
