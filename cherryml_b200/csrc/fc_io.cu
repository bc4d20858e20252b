// Host-side text I/O of the FastCherries stage, multithreaded (one family per task):
//   cherry_fc_read_msas     MSA files -> residue rows + family descriptors + sequence names
//                           (read_msa of FastCherries/io_helpers.cpp:35-74)
//   cherry_fc_write_outputs cherries / length indices / site categories -> the files the
//                           reference stage leaves behind: <family>.txt tree (star of cherries,
//                           phylogeny_estimation/_fast_cherries.py:121-141 + io/_tree.py write_tree),
//                           .newick, site rates (io_helpers.cpp:91-103), likelihood, .profiling.
// Numbers are formatted exactly like the reference's writers: '%.17f' where the C++ program
// prints (and the tree's branch lengths take the same round trip through that text), Python's
// repr(float) where its Python wrapper prints.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

struct IoErr {
  std::string msg;
};

std::string slurp(const char* path) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) throw IoErr{std::string("cannot open ") + path + ": " + strerror(errno)};
  struct stat st;
  std::string out;
  if (fstat(fd, &st) == 0 && st.st_size > 0) out.resize((size_t)st.st_size);
  size_t got = 0;
  while (got < out.size()) {
    ssize_t r = read(fd, &out[got], out.size() - got);
    if (r < 0) {
      if (errno == EINTR) continue;
      close(fd);
      throw IoErr{std::string("cannot read ") + path};
    }
    if (r == 0) break;
    got += (size_t)r;
  }
  close(fd);
  out.resize(got);
  return out;
}

void spill(const char* path, const std::string& text) {
  int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
  if (fd < 0) throw IoErr{std::string("cannot write ") + path + ": " + strerror(errno)};
  size_t put = 0;
  while (put < text.size()) {
    ssize_t r = write(fd, text.data() + put, text.size() - put);
    if (r < 0) {
      if (errno == EINTR) continue;
      close(fd);
      throw IoErr{std::string("cannot write ") + path};
    }
    put += (size_t)r;
  }
  close(fd);
}

// Python's repr(float): shortest digits that round-trip, exponent form iff the decimal point
// would sit more than 16 digits right or more than 3 zeros left of the first digit.
template <typename T>
void py_repr_t(T x, std::string* out) {
  if (x == 0.0) {
    *out += std::signbit(x) ? "-0.0" : "0.0";
    return;
  }
  if (std::isnan(x)) {
    *out += "nan";
    return;
  }
  if (std::isinf(x)) {
    *out += x < 0 ? "-inf" : "inf";
    return;
  }
  char buf[64];
  const auto res = std::to_chars(buf, buf + sizeof(buf) - 1, x, std::chars_format::scientific);
  *res.ptr = 0;
  const char* p = buf;
  if (*p == '-') {
    *out += '-';
    ++p;
  }
  std::string digits;
  const char* e = p;
  while (e < res.ptr && *e != 'e') {
    if (*e != '.') digits += *e;
    ++e;
  }
  const int exp10 = atoi(e + 1);
  const int decpt = exp10 + 1, nd = (int)digits.size();
  if (decpt > 16 || decpt < -3) {
    *out += digits[0];
    if (nd > 1) {
      *out += '.';
      out->append(digits, 1, std::string::npos);
    }
    char eb[16];
    snprintf(eb, sizeof(eb), "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
    *out += eb;
  } else if (decpt <= 0) {
    *out += "0.";
    out->append((size_t)(-decpt), '0');
    *out += digits;
  } else if (decpt >= nd) {
    *out += digits;
    out->append((size_t)(decpt - nd), '0');
    *out += ".0";
  } else {
    out->append(digits, 0, (size_t)decpt);
    *out += '.';
    out->append(digits, (size_t)decpt, std::string::npos);
  }
}

// Positive multiples of 1/4 (what count matrices hold) without going through to_chars / printf.
// Python repr: "<int>.0|.25|.5|.75" below 1e16; "%g": the same digits without a trailing ".0"
// as long as they fit in 6 significant digits.  Returns false when the general path must be used.
bool quarter_fast(double v, int cpp_style, std::string* out) {
  if (!(v > 0.0) || v >= 4503599627370496.0) return false;
  const double ipart = std::floor(v);
  const double f4 = (v - ipart) * 4.0;
  const int fi = (int)f4;
  if ((double)fi != f4) return false;
  const unsigned long long ip = (unsigned long long)ipart;
  if (cpp_style) {
    const unsigned long long limit = fi == 0 ? 1000000ull : fi == 2 ? 100000ull : 10000ull;
    if (ip >= limit) return false;
  } else if (ip >= 10000000000000000ull) {
    return false;
  }
  char buf[24];
  int n = 0;
  unsigned long long t = ip;
  do {
    buf[n++] = (char)('0' + t % 10);
    t /= 10;
  } while (t);
  while (n) *out += buf[--n];
  static const char* suffix_py[4] = {".0", ".25", ".5", ".75"};
  static const char* suffix_g[4] = {"", ".25", ".5", ".75"};
  *out += cpp_style ? suffix_g[fi] : suffix_py[fi];
  return true;
}

void py_repr(double x, std::string* out) { py_repr_t<double>(x, out); }

void fixed17(double x, std::string* out) {
  char buf[400];
  const int n = snprintf(buf, sizeof(buf), "%.17f", x);
  out->append(buf, (size_t)n);
}

double through_fixed17(double x) {
  char buf[400];
  snprintf(buf, sizeof(buf), "%.17f", x);
  return strtod(buf, nullptr);
}

template <typename F>
void parallel_for(int n, int n_threads, F&& body, std::string* first_error) {
  std::atomic<int> next{0};
  std::atomic<bool> failed{false};
  std::vector<std::string> errors((size_t)std::max(1, n_threads));
  auto worker = [&](int tid) {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n || failed.load()) return;
      try {
        body(i);
      } catch (const IoErr& err) {
        errors[(size_t)tid] = err.msg;
        failed.store(true);
        return;
      } catch (const std::exception& err) {
        errors[(size_t)tid] = err.what();
        failed.store(true);
        return;
      }
    }
  };
  n_threads = std::max(1, std::min(n_threads, n));
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  for (const std::string& e : errors)
    if (!e.empty()) {
      *first_error = e;
      return;
    }
}

struct ParsedMsa {
  std::string text;
  std::vector<std::pair<size_t, size_t>> names, seqs;  // (offset, length) into text
};

}  // namespace

extern "C" {

int cherry_fc_read_msas(const char* const* paths, int n_fams, const char* const* states, int n_states,
                        int n_threads, int pinned, cherry_fc_msas** out_ptr) {
  if (!paths || !states || !out_ptr || n_fams < 0) return cherry::fail(CHERRY_EINVAL, "null pointer");
  if (n_states < 1 || n_states > 254) return cherry::fail(CHERRY_ELIMIT, "number of states out of range");
  uint8_t lut[256];
  memset(lut, n_states, sizeof(lut));
  for (int i = 0; i < n_states; ++i) {
    if (!states[i] || strlen(states[i]) != 1) return cherry::fail(CHERRY_EINVAL, "states must be single characters");
    lut[(unsigned char)states[i][0]] = (uint8_t)i;
  }
  cherry::keep_large_buffers_on_heap();
  std::vector<ParsedMsa> parsed((size_t)n_fams);
  std::string err;
  parallel_for(n_fams, n_threads, [&](int f) {
    ParsedMsa& m = parsed[(size_t)f];
    m.text = slurp(paths[f]);
    const std::string& t = m.text;
    size_t pos = 0;
    const size_t n = t.size();
    while (pos < n) {
      const char* nl = (const char*)memchr(t.data() + pos, '\n', n - pos);
      const size_t end = nl ? (size_t)(nl - t.data()) : n;
      if (end > pos && t[pos] == '>') {
        if (end + 1 >= n) break;  // a name without a sequence line: the reference drops it
        const size_t s0 = end + 1;
        const char* nl2 = s0 < n ? (const char*)memchr(t.data() + s0, '\n', n - s0) : nullptr;
        const size_t s1 = nl2 ? (size_t)(nl2 - t.data()) : n;
        m.names.push_back({pos + 1, end - pos - 1});
        m.seqs.push_back({s0, s1 - s0});
        pos = s1 + 1;
      } else {
        pos = end + 1;
      }
    }
    for (const auto& s : m.seqs)
      if (s.second != m.seqs[0].second)
        throw IoErr{std::string("MSA ") + paths[f] + ": sequences of different lengths"};
    if (m.names.size() > 65535) throw IoErr{std::string("MSA ") + paths[f] + ": more than 65535 sequences"};
  }, &err);
  if (!err.empty()) return cherry::fail(CHERRY_EINVAL, "%s", err.c_str());

  cherry_fc_msas* r = (cherry_fc_msas*)calloc(1, sizeof(cherry_fc_msas));
  if (!r) return cherry::fail(CHERRY_EINVAL, "out of memory");
  r->n_fams = n_fams;
  r->pinned = 0;
  r->fams = (cherry_fc_family*)calloc((size_t)std::max(1, n_fams), sizeof(cherry_fc_family));
  int64_t off = 0, cherry_off = 0, site_off = 0, seq_off = 0, name_bytes = 0;
  for (int f = 0; f < n_fams; ++f) {
    const ParsedMsa& m = parsed[(size_t)f];
    const int64_t n = (int64_t)m.names.size(), L = n ? (int64_t)m.seqs[0].second : 0;
    const int64_t stride = std::max<int64_t>(16, (L + 15) / 16 * 16);
    cherry_fc_family& d = r->fams[f];
    d.msa_off = off;
    d.n_seqs = (int32_t)n;
    d.row_stride = (int32_t)stride;
    d.n_sites = (int32_t)L;
    d.cherry_off = (int32_t)cherry_off;
    d.site_off = (int32_t)site_off;
    d.seq_off = (int32_t)seq_off;
    off += n * stride;
    cherry_off += n / 2;
    site_off += L;
    seq_off += n;
    for (const auto& nm : m.names) name_bytes += (int64_t)nm.second;
  }
  if (cherry_off > INT32_MAX || site_off > INT32_MAX || seq_off > INT32_MAX) {
    free(r->fams);
    free(r);
    return cherry::fail(CHERRY_ELIMIT, "batch too large for 32-bit offsets: split the families");
  }
  r->msa_bytes = off;
  r->total_seqs = seq_off;
  r->total_sites = site_off;
  r->total_cherries = cherry_off;
  const size_t alloc = (size_t)std::max<int64_t>(64, off);
  if (pinned) r->msa = (uint8_t*)cherry::pinned_alloc(alloc);
  if (r->msa) {
    r->pinned = 1;
  } else {
    void* p = nullptr;
    if (posix_memalign(&p, 64, alloc) != 0) p = nullptr;
    r->msa = (uint8_t*)p;
  }
  r->name_blob = (char*)malloc((size_t)std::max<int64_t>(1, name_bytes));
  r->name_off = (int64_t*)malloc((size_t)(seq_off + 1) * sizeof(int64_t));
  if (!r->msa || !r->name_blob || !r->name_off) {
    cherry_fc_free_msas(r);
    return cherry::fail(CHERRY_EINVAL, "out of memory");
  }
  {
    int64_t nb = 0, s = 0;
    for (int f = 0; f < n_fams; ++f)
      for (const auto& nm : parsed[(size_t)f].names) {
        r->name_off[s++] = nb;
        nb += (int64_t)nm.second;
      }
    r->name_off[s] = nb;
  }
  const uint8_t skip = (uint8_t)n_states;
  parallel_for(n_fams, n_threads, [&](int f) {
    const ParsedMsa& m = parsed[(size_t)f];
    const cherry_fc_family& d = r->fams[f];
    for (int i = 0; i < d.n_seqs; ++i) {
      uint8_t* row = r->msa + d.msa_off + (int64_t)i * d.row_stride;
      const unsigned char* s = (const unsigned char*)m.text.data() + m.seqs[(size_t)i].first;
      for (int j = 0; j < d.n_sites; ++j) row[j] = lut[s[j]];
      memset(row + d.n_sites, skip, (size_t)(d.row_stride - d.n_sites));
      memcpy(r->name_blob + r->name_off[d.seq_off + i], m.text.data() + m.names[(size_t)i].first,
             m.names[(size_t)i].second);
    }
  }, &err);
  *out_ptr = r;
  return CHERRY_OK;
}

void cherry_fc_free_msas(cherry_fc_msas* r) {
  if (!r) return;
  if (r->msa) {
    if (r->pinned) cherry::pinned_free(r->msa); else free(r->msa);
  }
  free(r->fams);
  free(r->name_blob);
  free(r->name_off);
  free(r);
}

int cherry_fc_write_outputs(const cherry_fc_msas* m, const int32_t* pair_a, const int32_t* pair_b,
                            const int32_t* unpaired, const int32_t* len_idx, const int32_t* site_cat,
                            const double* grid, int K, const double* cats, int R, const char* const* tree_paths,
                            const char* const* newick_paths, const char* const* site_rate_paths,
                            const char* const* likelihood_paths, const char* const* profiling_paths,
                            const double* profiling /* [n_fams][4]: pairing, ble, cpp, total seconds */,
                            int n_threads) {
  if (!m || !pair_a || !pair_b || !unpaired || !len_idx || !site_cat || !grid || !cats || !tree_paths ||
      !site_rate_paths)
    return cherry::fail(CHERRY_EINVAL, "null pointer");
  cherry::keep_large_buffers_on_heap();
  std::string err;
  parallel_for(m->n_fams, n_threads, [&](int f) {
    const cherry_fc_family& d = m->fams[f];
    const int n_cherries = d.n_seqs / 2, L = d.n_sites;
    const int32_t* pa = pair_a + d.cherry_off;
    const int32_t* pb = pair_b + d.cherry_off;
    const int32_t* li = len_idx + d.cherry_off;
    const int32_t* sc = site_cat + d.site_off;
    for (int c = 0; c < n_cherries; ++c)
      if (li[c] < 0 || li[c] >= K || pa[c] < 0 || pa[c] >= d.n_seqs || pb[c] < 0 || pb[c] >= d.n_seqs)
        throw IoErr{"cherry_fc_write_outputs: index out of range"};
    // fast_cherries.cpp:268-279: rates to mean 1 (left-to-right sum), lengths absorb the factor
    double sum = 0.0;
    for (int j = 0; j < L; ++j) {
      if (sc[j] < 0 || sc[j] >= R) throw IoErr{"cherry_fc_write_outputs: category out of range"};
      sum += cats[sc[j]];
    }
    const double mean = L ? sum / (double)L : 1.0;
    auto name = [&](int row) {
      const int64_t a = m->name_off[d.seq_off + row], b = m->name_off[d.seq_off + row + 1];
      return std::string(m->name_blob + a, (size_t)(b - a));
    };
    std::string text;
    if (site_rate_paths[f]) {
      text = std::to_string(L) + " sites\n";
      for (int j = 0; j < L; ++j) {
        fixed17(cats[sc[j]] / mean, &text);
        text += ' ';
      }
      spill(site_rate_paths[f], text);
    }
    const int u = unpaired[f];
    const int n_nodes = 1 + 3 * n_cherries + (u >= 0 ? 1 : 0);
    std::string tree = std::to_string(n_nodes) + " nodes\nroot\n";
    std::string edges = std::to_string(n_nodes - 1) + " edges\n";  // insertion order of the reference's DFS
    std::string newick = "(";
    for (int c = 0; c < n_cherries; ++c) {
      const double half = through_fixed17(grid[li[c]] * mean) / 2.0;
      const std::string inner = "internal-" + std::to_string(c), na = name(pa[c]), nb = name(pb[c]);
      tree += inner + "\n" + na + "\n" + nb + "\n";
      std::string hs;
      py_repr(half, &hs);
      edges += "root " + inner + " 1.0\n" + inner + " " + na + " " + hs + "\n" + inner + " " + nb + " " + hs + "\n";
      char g[64];
      snprintf(g, sizeof(g), "%g", half);
      if (c) newick += ',';
      newick += "(" + na + ":" + g + "," + nb + ":" + g + ")" + inner + ":1";
    }
    if (u >= 0) {
      if (u >= d.n_seqs) throw IoErr{"cherry_fc_write_outputs: unpaired index out of range"};
      tree += name(u) + "\n";
      edges += "root " + name(u) + " 1.0\n";
      if (n_cherries) newick += ',';
      newick += name(u) + ":1";
    }
    newick += ");";
    spill(tree_paths[f], tree + edges);
    if (newick_paths && newick_paths[f]) spill(newick_paths[f], newick);
    if (likelihood_paths && likelihood_paths[f]) spill(likelihood_paths[f], "0.0");
    if (profiling_paths && profiling_paths[f] && profiling) {
      static const char* keys[4] = {"pairing_time: ", "ble_time: ", "cpp_time: ", "total_time: "};
      std::string p;
      for (int k = 0; k < 4; ++k) {
        p += keys[k];
        py_repr(profiling[(size_t)f * 4 + k], &p);
        if (k < 3) p += '\n';
      }
      spill(profiling_paths[f], p);
    }
  }, &err);
  if (!err.empty()) return cherry::fail(CHERRY_EINVAL, "%s", err.c_str());
  return CHERRY_OK;
}


// Branch lengths and site rates exactly as the counting stage reads them back from the files
// cherry_fc_write_outputs writes: pair_t[c] = the cherry's path length -- the tree's two edges of
// round_trip('%.17f', grid[len_idx] * mean) / 2 each, printed with repr and parsed again (through
// float32 when float32_lengths, like the C++ counting program's std::stof) and added;
// rate_table[f][r] = round_trip('%.17f', cats[r] / mean_f), the value a site of category r has in
// the site-rates file.  This is what lets FastCherries hand its result to the counting kernels in
// memory with counts identical to the route through the text files.
int cherry_fc_lengths_and_rates(const cherry_fc_family* fams, int n_fams, const int32_t* len_idx,
                                const int32_t* site_cat, const double* grid, int K, const double* cats, int R,
                                int float32_lengths, double* pair_t, double* rate_table, int n_threads) {
  if (!fams || !len_idx || !site_cat || !grid || !cats || !pair_t || !rate_table)
    return cherry::fail(CHERRY_EINVAL, "null pointer");
  std::string err;
  parallel_for(n_fams, n_threads, [&](int f) {
    const cherry_fc_family& d = fams[f];
    const int n_cherries = d.n_seqs / 2, L = d.n_sites;
    const int32_t* sc = site_cat + d.site_off;
    double sum = 0.0;
    for (int j = 0; j < L; ++j) {
      if (sc[j] < 0 || sc[j] >= R) throw IoErr{"cherry_fc_lengths_and_rates: category out of range"};
      sum += cats[sc[j]];
    }
    const double mean = L ? sum / (double)L : 1.0;
    for (int r = 0; r < R; ++r) rate_table[(size_t)f * R + r] = through_fixed17(cats[r] / mean);
    std::vector<double> by_index((size_t)K, -1.0);
    std::string text;
    for (int c = 0; c < n_cherries; ++c) {
      const int k = len_idx[d.cherry_off + c];
      if (k < 0 || k >= K) throw IoErr{"cherry_fc_lengths_and_rates: length index out of range"};
      if (by_index[(size_t)k] < 0) {
        const double half = through_fixed17(grid[k] * mean) / 2.0;
        text.clear();
        py_repr(half, &text);
        const double edge = float32_lengths ? (double)strtof(text.c_str(), nullptr) : strtod(text.c_str(), nullptr);
        by_index[(size_t)k] = (0.0 + edge) + (0.0 + edge);  // the two leaf-to-parent distances of the traversal
      }
      pair_t[d.cherry_off + c] = by_index[(size_t)k];
    }
  }, &err);
  if (!err.empty()) return cherry::fail(CHERRY_EINVAL, "%s", err.c_str());
  return CHERRY_OK;
}

// Host metadata of the LG counting batch for FastCherries results, with the layout rules of the
// ingest (process_family in ingest.cu): categories = the site_cat values present in the family,
// ascending; sites keep their order inside a category; every category is padded to 4 sites; the
// row stride is a multiple of 16.  Two calls: out arrays NULL -> only the sizes (sizes[0] = encoded
// residue bytes, [1] = group_cat entries, [2] = rate values, [3] = tiles, [4] = r_pad,
// [5] = (pair, site) items); then with the arrays allocated.
int cherry_fc_count_layout(const cherry_fc_family* fams, int n_fams, const int32_t* site_cat,
                           const double* rate_table, int R, int chunks_per_tile, int32_t* dest,
                           uint16_t* group_cat, double* rate_vals, cherry_fam_desc* out_fams, cherry_tile* tiles,
                           int64_t* sizes, int n_threads) {
  if (!fams || !site_cat || !sizes || R < 1) return cherry::fail(CHERRY_EINVAL, "null pointer");
  const bool fill = dest && group_cat && rate_vals && out_fams && tiles && rate_table;
  std::vector<int64_t> stride((size_t)n_fams), n_rates((size_t)n_fams), n_tiles((size_t)n_fams);
  std::string err;
  parallel_for(n_fams, n_threads, [&](int f) {
    const cherry_fc_family& d = fams[f];
    std::vector<int64_t> cnt((size_t)R, 0);
    for (int j = 0; j < d.n_sites; ++j) {
      const int c = site_cat[d.site_off + j];
      if (c < 0 || c >= R) throw IoErr{"cherry_fc_count_layout: category out of range"};
      ++cnt[(size_t)c];
    }
    int64_t total = 0, present = 0;
    for (int r = 0; r < R; ++r)
      if (cnt[(size_t)r]) {
        total += (cnt[(size_t)r] + 3) / 4 * 4;
        ++present;
      }
    stride[(size_t)f] = std::max<int64_t>(16, (total + 15) / 16 * 16);
    n_rates[(size_t)f] = present ? present : 1;
    const int64_t per_tile = std::max<int64_t>(1, chunks_per_tile / std::max<int64_t>(1, stride[(size_t)f] / 16));
    const int64_t np = d.n_seqs / 2;
    n_tiles[(size_t)f] = (np + per_tile - 1) / per_tile;
  }, &err);
  if (!err.empty()) return cherry::fail(CHERRY_EINVAL, "%s", err.c_str());
  std::vector<int64_t> msa_off((size_t)n_fams), aux_off((size_t)n_fams), rate_off((size_t)n_fams),
      tile_off((size_t)n_fams), pair_off((size_t)n_fams);
  int64_t msa_bytes = 0, n_aux = 0, n_rate = 0, n_tile = 0, n_pair = 0, examined = 0, max_rates = 1;
  for (int f = 0; f < n_fams; ++f) {
    const int64_t np = fams[f].n_seqs / 2;
    msa_off[(size_t)f] = msa_bytes;
    aux_off[(size_t)f] = n_aux;
    rate_off[(size_t)f] = n_rate;
    tile_off[(size_t)f] = n_tile;
    pair_off[(size_t)f] = n_pair;
    msa_bytes += 2 * np * stride[(size_t)f];
    n_aux += stride[(size_t)f] / 4;
    n_rate += n_rates[(size_t)f];
    n_tile += n_tiles[(size_t)f];
    n_pair += np;
    examined += np * fams[f].n_sites;
    max_rates = std::max(max_rates, n_rates[(size_t)f]);
  }
  sizes[0] = msa_bytes;
  sizes[1] = n_aux;
  sizes[2] = n_rate;
  sizes[3] = n_tile;
  sizes[4] = (max_rates + 3) / 4 * 4;
  sizes[5] = examined;
  if (n_aux > INT32_MAX || n_rate > INT32_MAX || n_tile > INT32_MAX || n_pair > INT32_MAX)
    return cherry::fail(CHERRY_ELIMIT, "batch too large for 32-bit indices (split the families)");
  if (!fill) return CHERRY_OK;
  parallel_for(n_fams, n_threads, [&](int f) {
    const cherry_fc_family& d = fams[f];
    const int32_t* sc = site_cat + d.site_off;
    std::vector<int64_t> cnt((size_t)R, 0), start((size_t)R, 0);
    std::vector<int> index_of((size_t)R, -1);
    for (int j = 0; j < d.n_sites; ++j) ++cnt[(size_t)sc[j]];
    uint16_t* gc = group_cat + aux_off[(size_t)f];
    for (int64_t g = 0; g < stride[(size_t)f] / 4; ++g) gc[g] = 0;
    double* rv = rate_vals + rate_off[(size_t)f];
    int64_t total = 0;
    int present = 0;
    for (int r = 0; r < R; ++r) {
      if (!cnt[(size_t)r]) continue;
      start[(size_t)r] = total;
      index_of[(size_t)r] = present;
      const int64_t padded = (cnt[(size_t)r] + 3) / 4 * 4;
      for (int64_t g = total / 4; g < (total + padded) / 4; ++g) gc[g] = (uint16_t)present;
      rv[present] = rate_table[(size_t)f * R + r];
      total += padded;
      ++present;
    }
    if (!present) rv[0] = 1.0;
    for (int j = 0; j < d.n_sites; ++j) dest[d.site_off + j] = (int32_t)start[(size_t)sc[j]]++;
    cherry_fam_desc& o = out_fams[f];
    o.msa_off = msa_off[(size_t)f];
    o.row_stride = (int32_t)stride[(size_t)f];
    o.n_chunks = (int32_t)(stride[(size_t)f] / 16);
    o.aux_off = (int32_t)aux_off[(size_t)f];
    o.aux_cnt = (int32_t)(stride[(size_t)f] / 4);
    o.rate_off = (int32_t)rate_off[(size_t)f];
    o.n_rates = (int32_t)n_rates[(size_t)f];
    const int64_t np = d.n_seqs / 2;
    const int64_t per_tile = std::max<int64_t>(1, chunks_per_tile / std::max<int64_t>(1, stride[(size_t)f] / 16));
    int64_t ti = tile_off[(size_t)f];
    for (int64_t b = 0; b < np; b += per_tile) {
      cherry_tile& tl = tiles[ti++];
      tl.fam = f;
      tl.pair_begin = (int32_t)(pair_off[(size_t)f] + b);
      tl.n_pairs = (int32_t)std::min<int64_t>(per_tile, np - b);
      tl.reserved = 0;
    }
  }, &err);
  if (!err.empty()) return cherry::fail(CHERRY_EINVAL, "%s", err.c_str());
  return CHERRY_OK;
}

// ------------------------------------------------------------------ count matrices (result.txt)
// The reference's two writers: io/_count_matrices.py:66-81 (pandas to_csv, repr floats) and the
// C++ program's writer (counting/_count_transitions.cpp:524-548, ostream << double = "%g").

int cherry_write_count_matrices(const char* path, const double* q, int K, const char* const* states, int S,
                                const double* counts, int cpp_style, int n_threads) {
  if (!path || !q || !states || !counts || K < 0 || S < 1) return cherry::fail(CHERRY_EINVAL, "null pointer");
  cherry::keep_large_buffers_on_heap();
  std::string header;
  if (cpp_style) {
    header = "\t";
    for (int i = 0; i < S; ++i) header += std::string(states[i]) + "\t";
  } else {
    for (int i = 0; i < S; ++i) header += "\t" + std::string(states[i]);
  }
  header += "\n";
  std::vector<std::string> blocks((size_t)K);
  std::string err;
  parallel_for(K, n_threads, [&](int k) {
    std::string& b = blocks[(size_t)k];
    b.reserve((size_t)S * S * 5 + header.size() + 64);
    char buf[64];
    if (cpp_style) {
      b.append(buf, (size_t)snprintf(buf, sizeof(buf), "%g", q[k]));
    } else {
      py_repr(q[k], &b);
    }
    b += '\n';
    b += header;
    const double* m = counts + (size_t)k * S * S;
    for (int i = 0; i < S; ++i) {
      b += states[i];
      for (int j = 0; j < S; ++j) {
        const double v = m[(size_t)i * S + j];
        b += '\t';
        if (v == 0.0 && !std::signbit(v)) {
          b += cpp_style ? "0" : "0.0";
        } else if (quarter_fast(v, cpp_style, &b)) {
          // counts are multiples of 1/4: integer digits + one of four suffixes
        } else if (cpp_style) {
          b.append(buf, (size_t)snprintf(buf, sizeof(buf), "%g", v));
        } else {
          py_repr(v, &b);
        }
      }
      b += '\n';
    }
  }, &err);
  if (!err.empty()) return cherry::fail(CHERRY_EINVAL, "%s", err.c_str());
  try {
    int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) throw IoErr{std::string("cannot write ") + path + ": " + strerror(errno)};
    auto put = [&](const std::string& t) {
      size_t done = 0;
      while (done < t.size()) {
        ssize_t r = write(fd, t.data() + done, t.size() - done);
        if (r < 0) {
          if (errno == EINTR) continue;
          close(fd);
          throw IoErr{std::string("cannot write ") + path};
        }
        done += (size_t)r;
      }
    };
    put(std::to_string(K) + " matrices\n" + std::to_string(S) + " states\n");
    for (const std::string& b : blocks) put(b);
    close(fd);
  } catch (const IoErr& e) {
    return cherry::fail(CHERRY_EINVAL, "%s", e.msg.c_str());
  }
  return CHERRY_OK;
}

namespace {
struct CountFile {
  std::string text;
  std::vector<size_t> line_start;  // offsets of the lines of text.strip().split("\n"), + end sentinel
  size_t end = 0;
};

bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

void index_lines(CountFile* f) {
  const std::string& t = f->text;
  size_t b = 0, e = t.size();
  while (b < e && is_ws(t[b])) ++b;
  while (e > b && is_ws(t[e - 1])) --e;
  f->end = e;
  size_t s = b;
  while (s <= e) {
    f->line_start.push_back(s);
    const char* nl = s < e ? (const char*)memchr(t.data() + s, '\n', e - s) : nullptr;
    if (!nl) break;
    s = (size_t)(nl - t.data()) + 1;
  }
}

// "<n> <word>" (after strip) -> n, or -1
long long header_number(const CountFile& f, size_t line, const char* word) {
  if (line >= f.line_start.size()) return -1;
  const size_t a = f.line_start[line];
  const size_t b = line + 1 < f.line_start.size() ? f.line_start[line + 1] - 1 : f.end;
  std::string ln(f.text.data() + a, b - a);
  while (!ln.empty() && is_ws(ln.back())) ln.pop_back();
  size_t st = 0;
  while (st < ln.size() && is_ws(ln[st])) ++st;
  const size_t sp = ln.find(' ', st);
  if (sp == std::string::npos || ln.substr(sp + 1) != word) return -1;
  char* endp = nullptr;
  const long long n = strtoll(ln.c_str() + st, &endp, 10);
  if (endp != ln.c_str() + sp || n < 0) return -1;
  return n;
}
}  // namespace

int cherry_read_count_matrices_header(const char* path, int* K_out, int* S_out) {
  if (!path || !K_out || !S_out) return cherry::fail(CHERRY_EINVAL, "null pointer");
  try {
    int fd = open(path, O_RDONLY);
    if (fd < 0) throw IoErr{std::string("cannot open ") + path + ": " + strerror(errno)};
    char buf[256];
    const ssize_t n = read(fd, buf, sizeof(buf) - 1);
    close(fd);
    CountFile f;
    f.text.assign(buf, n > 0 ? (size_t)n : 0);
    // only the first two lines matter here; cut at the second newline
    size_t nl1 = f.text.find('\n');
    size_t nl2 = nl1 == std::string::npos ? nl1 : f.text.find('\n', nl1 + 1);
    if (nl2 != std::string::npos) f.text.resize(nl2);
    index_lines(&f);
    const long long K = header_number(f, 0, "matrices"), S = header_number(f, 1, "states");
    if (K < 0) throw IoErr{std::string("In file ") + path + ", expected line '[num_matrices] matrices'"};
    if (S < 0) throw IoErr{std::string("In file ") + path + ", expected line '[num_states] states'"};
    *K_out = (int)K;
    *S_out = (int)S;
  } catch (const IoErr& e) {
    return cherry::fail(CHERRY_EINVAL, "%s", e.msg.c_str());
  }
  return CHERRY_OK;
}

int cherry_read_count_matrices(const char* path, int K, int S, double* q, double* counts, char* states_out,
                               size_t states_cap, int n_threads) {
  if (!path || !q || !counts || !states_out) return cherry::fail(CHERRY_EINVAL, "null pointer");
  cherry::keep_large_buffers_on_heap();
  CountFile f;
  try {
    f.text = slurp(path);
  } catch (const IoErr& e) {
    return cherry::fail(CHERRY_EINVAL, "%s", e.msg.c_str());
  }
  index_lines(&f);
  if ((long long)f.line_start.size() < 2 + (long long)K * (S + 2))
    return cherry::fail(CHERRY_EINVAL, "count matrices file %s is truncated", path);
  auto line = [&](size_t i, const char** a, const char** b) {
    *a = f.text.data() + f.line_start[i];
    *b = f.text.data() + (i + 1 < f.line_start.size() ? f.line_start[i + 1] - 1 : f.end);
  };
  std::vector<std::string> state_names((size_t)K ? (size_t)S : 0);
  std::string err;
  parallel_for(K, n_threads, [&](int k) {
    const size_t l0 = 2 + (size_t)k * (S + 2);
    const char *a, *b;
    line(l0, &a, &b);
    {
      std::string tok(a, (size_t)(b - a));
      char* endp = nullptr;
      q[k] = strtod(tok.c_str(), &endp);
      while (*endp && is_ws(*endp)) ++endp;
      if (endp == tok.c_str() || *endp) throw IoErr{"count matrices: bad quantization point '" + tok + "'"};
    }
    line(l0 + 1, &a, &b);
    int n_hdr = 0;
    for (const char* p = a; p < b;) {
      while (p < b && is_ws(*p)) ++p;
      const char* s0 = p;
      while (p < b && !is_ws(*p)) ++p;
      if (p > s0) {
        if (k == K - 1 && n_hdr < S) state_names[(size_t)n_hdr].assign(s0, (size_t)(p - s0));
        ++n_hdr;
      }
    }
    if (n_hdr != S)
      throw IoErr{"Error reading count matrices file: expected " + std::to_string(S) + " states in a header line, found " +
                  std::to_string(n_hdr)};
    double* m = counts + (size_t)k * S * S;
    for (int i = 0; i < S; ++i) {
      line(l0 + 2 + (size_t)i, &a, &b);
      const char* p = a;
      while (p < b && is_ws(*p)) ++p;
      while (p < b && !is_ws(*p)) ++p;  // the row label
      int got = 0;
      while (p < b) {
        while (p < b && is_ws(*p)) ++p;
        if (p >= b) break;
        char* endp = nullptr;
        const double v = strtod(p, &endp);  // the buffer ends with a NUL (std::string), tokens end at whitespace
        if (endp == p) throw IoErr{"Could not read count matrices: bad number"};
        if (got < S) m[(size_t)i * S + got] = v;
        ++got;
        p = endp;
      }
      if (got != S) throw IoErr{"Could not read count matrices. Matrix " + std::to_string(k) + " is ragged"};
    }
  }, &err);
  if (!err.empty()) return cherry::fail(CHERRY_EINVAL, "%s", err.c_str());
  std::string blob;
  for (int i = 0; i < S && K > 0; ++i) {
    blob += state_names[(size_t)i];
    blob += '\n';
  }
  if (blob.size() + 1 > states_cap) return cherry::fail(CHERRY_EINVAL, "states buffer too small");
  memcpy(states_out, blob.c_str(), blob.size() + 1);
  return CHERRY_OK;
}


// Labelled square table (rate matrices, masks): "\t<col>...\n" then "<row>\t<v>\t<v>...\n" with
// every fp64 number printed as the shortest string that round-trips -- what pandas' to_csv /
// str(numpy.float64) print (reference io/_rate_matrix.py:37-52).
int cherry_write_labelled_matrix(const char* path, const char* const* states, int S, const double* data,
                                 int n_threads) {
  if (!path || !states || !data || S < 1) return cherry::fail(CHERRY_EINVAL, "null pointer");
  cherry::keep_large_buffers_on_heap();
  std::vector<std::string> rows((size_t)S);
  std::string err;
  parallel_for(S, n_threads, [&](int i) {
    std::string& b = rows[(size_t)i];
    b.reserve((size_t)S * 24 + 32);
    b += states[i];
    for (int j = 0; j < S; ++j) {
      b += '\t';
      py_repr(data[(size_t)i * S + j], &b);
    }
    b += '\n';
  }, &err);
  if (!err.empty()) return cherry::fail(CHERRY_EINVAL, "%s", err.c_str());
  std::string text;
  for (int j = 0; j < S; ++j) text += "\t" + std::string(states[j]);
  text += '\n';
  for (const std::string& r : rows) text += r;
  try {
    spill(path, text);
  } catch (const IoErr& e) {
    return cherry::fail(CHERRY_EINVAL, "%s", e.msg.c_str());
  }
  return CHERRY_OK;
}

}  // extern "C"
