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
void py_repr(double x, std::string* out) {
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

}  // extern "C"
