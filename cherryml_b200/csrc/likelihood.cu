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
// with P_v = expm(branch length * site rate * Q), computed beforehand by cherry_expm_batched.
#include <cstdint>

#include "common.cuh"

namespace {

struct LlArgs {
  const cherry_ll_node* nodes;
  const int32_t* p_index;  // [n_nodes][n_cats]
  const double* P;         // [n_matrices][Su][Su]
  const uint8_t* obs;      // [n_leaves][n_units][c]
  const int32_t* unit_cat; // [n_units]
  const double* pi;        // [Su]
  double* acc;             // [gridDim.x][max_depth + 1][G][Su]
  double* ll_out;          // [n_units]
  int n_nodes, n_cats, S, c, Su, n_units, G, max_depth;
};

__global__ void tree_ll_kernel(LlArgs a) {
  extern __shared__ double w[];  // [G][Su]
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
    double m = 0.0;
    if (!is_leaf) {
      // dp of this node is complete: w = exp(dp - max dp) (observation vector of an internal node = ones)
      const double v = live ? acc[(size_t)node.depth * acc_stride] : 0.0;
      if (live) wg[s] = v;
      __syncthreads();
      double mx = -INFINITY;
      if (live)
        for (int k = 0; k < Su; ++k) mx = fmax(mx, wg[k]);
      __syncthreads();
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
          const double* row = a.P + ((size_t)a.p_index[(size_t)i * a.n_cats + cat] * Su + s) * Su;
          for (int k = 0; k < Su; ++k) sum += row[k] * wg[k];
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
        const double* row = a.P + ((size_t)a.p_index[(size_t)i * a.n_cats + cat] * Su + s) * Su;
        if (a.c == 1) {
          const int x = ob[0];
          if (x != S) {
            sum = row[x];
          } else {
            for (int k = 0; k < Su; ++k) sum += row[k];
          }
        } else {
          const int x = ob[0], y = ob[1];
          if (x != S && y != S) {
            sum = row[x * S + y];
          } else if (x != S) {
            for (int k = 0; k < S; ++k) sum += row[x * S + k];
          } else if (y != S) {
            for (int k = 0; k < S; ++k) sum += row[k * S + y];
          } else {
            for (int k = 0; k < Su; ++k) sum += row[k];
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

}  // namespace

extern "C" {

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
  tree_ll_kernel<<<blocks, threads, (size_t)a.G * Su * sizeof(double), (cudaStream_t)stream>>>(a);
  CHERRY_LAUNCH_CHECK("tree_ll_kernel");
  return CHERRY_OK;
}

}  // extern "C"
