/* CPU ORACLE (C) for transition counting on integer-encoded families.
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and
 * bench.py's CPU-baseline legs load this library.
 *
 * A scalar restatement of the reference's inner loops (songlab-cal/CherryML v0.2.0) over the
 * same encoded arrays the GPU consumes, so that mid-size inputs can be checked in seconds:
 *   quantization_idx        counting/_count_transitions.cpp:295-307 (== cherryml/utils.py:35-56)
 *   LG per-site loop        counting/_count_transitions.cpp:368-381 (cherry: += 0.5 twice),
 *                           :452-466 (edge: += 1.0)
 *   co per-contact loop     counting/_count_co_transitions.cpp:358-383 (+= 0.25 four times),
 *                           :469-496 (edge: += 0.5 twice)
 * As in the reference the bucket is recomputed for EVERY (pair, site) from t * rate, and
 * the accumulators are doubles.
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libcount_oracle.so oracle/count_encoded.c
 */
#include <stdint.h>
#include <stddef.h>

typedef struct {
  int64_t msa_off;
  int32_t row_stride, n_chunks, aux_off, aux_cnt, rate_off, n_rates;
} fam_desc;

int oracle_quantization_idx(double t, const double* q, int K) {
  if (t < q[0] || t > q[K - 1]) return -1;
  int lo = 0, hi = K; /* std::lower_bound */
  while (lo < hi) {
    int mid = lo + (hi - lo) / 2;
    if (q[mid] < t) lo = mid + 1; else hi = mid;
  }
  if (lo == 0) return 0;
  double left = q[lo - 1], right = q[lo];
  volatile double el = t / left - 1;   /* volatile: no extended precision / contraction */
  volatile double er = right / t - 1;
  return (el < er) ? lo - 1 : lo;
}

/* counts: double [K][S][S], accumulated into. */
void oracle_count_lg(const uint8_t* msa, const fam_desc* fams, const int32_t* pair_a,
                     const int32_t* pair_b, const double* pair_t, const int32_t* pair_fam,
                     int64_t n_pairs, const double* rate_vals, const uint16_t* group_cat,
                     const double* grid, int K, int S, int directed, double* counts) {
  for (int64_t p = 0; p < n_pairs; ++p) {
    const fam_desc* fd = &fams[pair_fam[p]];
    const uint8_t* ra = msa + fd->msa_off + (int64_t)pair_a[p] * fd->row_stride;
    const uint8_t* rb = msa + fd->msa_off + (int64_t)pair_b[p] * fd->row_stride;
    for (int j = 0; j < fd->row_stride; ++j) {
      unsigned x = ra[j], y = rb[j];
      if (x >= (unsigned)S || y >= (unsigned)S) continue;
      double rate = rate_vals[fd->rate_off + group_cat[fd->aux_off + j / 4]];
      volatile double tt = pair_t[p] * rate;
      int b = oracle_quantization_idx(tt, grid, K);
      if (b < 0) continue;
      double* c = counts + (size_t)b * S * S;
      if (directed) {
        c[x * S + y] += 1.0;
      } else {
        c[x * S + y] += 0.5;
        c[y * S + x] += 0.5;
      }
    }
  }
}

/* counts: double [K][S*S][S*S], accumulated into.  Rows are contact-paired: bytes 2k, 2k+1
 * of a row are the residues at the two sites of the family's k-th contact (the encoder's
 * layout, cherryml_b200/counting/_ingest.py contact_paired_rows); `contacts` is unused. */
void oracle_count_co(const uint8_t* msa, const fam_desc* fams, const int32_t* pair_a,
                     const int32_t* pair_b, const double* pair_t, const int32_t* pair_fam,
                     int64_t n_pairs, const int32_t* contacts, const double* grid, int K, int S,
                     int directed, double* counts) {
  const size_t n = (size_t)S * S;
  for (int64_t p = 0; p < n_pairs; ++p) {
    const fam_desc* fd = &fams[pair_fam[p]];
    int b = oracle_quantization_idx(pair_t[p], grid, K);
    if (b < 0) continue;
    const uint8_t* ra = msa + fd->msa_off + (int64_t)pair_a[p] * fd->row_stride;
    const uint8_t* rb = msa + fd->msa_off + (int64_t)pair_b[p] * fd->row_stride;
    double* c = counts + (size_t)b * n * n;
    for (int k = 0; k < fd->aux_cnt; ++k) {
      (void)contacts;
      unsigned xi = ra[2 * k], xj = ra[2 * k + 1], yi = rb[2 * k], yj = rb[2 * k + 1];
      if (xi >= (unsigned)S || xj >= (unsigned)S || yi >= (unsigned)S || yj >= (unsigned)S) continue;
      size_t s = xi * S + xj, e = yi * S + yj, sr = xj * S + xi, er = yj * S + yi;
      if (directed) {
        c[s * n + e] += 0.5;
        c[sr * n + er] += 0.5;
      } else {
        c[s * n + e] += 0.25;
        c[e * n + s] += 0.25;
        c[sr * n + er] += 0.25;
        c[er * n + sr] += 0.25;
      }
    }
  }
}
