// Native, multithreaded ingest for counting (host code): text directories -> the encoded
// batch the counting kernels read (see DESIGN.md "Data layout in HBM").
//
// What it replaces in the reference (songlab-cal/CherryML v0.2.0): the text readers and the
// pairing traversal of the C++ counting programs --
//   read_tree / read_msa / read_site_rates / read_contact_map
//                      counting/_count_transitions.cpp:209-293, _count_co_transitions.cpp:209-293
//   _dfs (cherry++), the cherry / edge loops
//                      counting/_count_transitions.cpp:316-390, 444-506
// with the accept/reject behaviour of the Python readers (io/_tree.py:214-265,
// io/_msa.py:51-73, io/_site_rates.py:5-26, io/_contact_map.py:6-28) that the stage
// functions use.  One family = one task; `n_threads` workers pull families from an atomic
// counter (the reference's parallelism is one MPI rank per family stripe, .cpp:624-629).
// Branch lengths: `float32_branch_lengths` parses them with strtof exactly like the C++
// program's std::stof (.cpp:247); otherwise strtod (Python float()).
#include <algorithm>
#include <mutex>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "common.cuh"

namespace {

constexpr int kTargetChunksPerTile = 16384;  // keep in sync with counting/_ingest.py
constexpr int kTargetItemsPerCoTile = 8192;

struct Err {
  std::string msg;
};
[[noreturn]] void die(const std::string& m) { throw Err{m}; }

std::string read_file(const std::string& path) {
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) die("cannot open " + path + ": " + strerror(errno));
  struct stat st;
  std::string out;
  if (fstat(fd, &st) == 0 && st.st_size > 0) out.resize((size_t)st.st_size);
  size_t got = 0;
  while (got < out.size()) {
    ssize_t r = read(fd, &out[got], out.size() - got);
    if (r < 0) {
      if (errno == EINTR) continue;
      close(fd);
      die("cannot read " + path + ": " + strerror(errno));
    }
    if (r == 0) break;
    got += (size_t)r;
  }
  close(fd);
  out.resize(got);
  return out;
}

struct Span {
  const char* p;
  size_t n;
  std::string str() const { return std::string(p, n); }
  bool eq(const char* s) const { return strlen(s) == n && memcmp(p, s, n) == 0; }
};

bool is_py_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// Python's  text.strip().split("\n")
std::vector<Span> strip_lines(const std::string& text) {
  size_t b = 0, e = text.size();
  while (b < e && is_py_space(text[b])) ++b;
  while (e > b && is_py_space(text[e - 1])) --e;
  std::vector<Span> lines;
  const char* base = text.data();
  size_t s = b;
  while (s <= e) {
    const char* nl = s < e ? (const char*)memchr(base + s, '\n', e - s) : nullptr;
    const size_t stop = nl ? (size_t)(nl - base) : e;
    lines.push_back(Span{base + s, stop - s});
    if (!nl) break;
    s = stop + 1;
  }
  return lines;
}

// Python's  line.split(" ")
std::vector<Span> split_space(Span line) {
  std::vector<Span> out;
  size_t s = 0;
  for (size_t i = 0; i <= line.n; ++i) {
    if (i == line.n || line.p[i] == ' ') {
      out.push_back(Span{line.p + s, i - s});
      s = i + 1;
    }
  }
  return out;
}

bool parse_int(Span t, long long* v) {
  if (t.n == 0 || t.n > 30) return false;
  char buf[32];
  memcpy(buf, t.p, t.n);
  buf[t.n] = 0;
  char* end = nullptr;
  errno = 0;
  long long x = strtoll(buf, &end, 10);
  if (errno || end != buf + t.n) return false;
  *v = x;
  return true;
}

bool parse_double(Span t, bool as_float32, double* v) {
  if (t.n == 0 || t.n > 120) return false;
  char buf[128];
  memcpy(buf, t.p, t.n);
  buf[t.n] = 0;
  char* end = nullptr;
  if (as_float32) {
    float f = strtof(buf, &end);
    *v = (double)f;
  } else {
    *v = strtod(buf, &end);
  }
  return end == buf + t.n;
}

// "<n> <word>" header line
long long header_count(Span line, const char* word, const std::string& what) {
  std::vector<Span> t = split_space(line);
  long long n = 0;
  if (t.size() != 2 || !t[1].eq(word) || !parse_int(t[0], &n) || n < 0)
    die(what + " (found: '" + line.str() + "')");
  return n;
}

// Open-addressing hash map from a byte span (a name inside a file buffer) to an int: no
// allocation per key, which is what dominated the std::unordered_map<std::string, ...> version.
class SpanMap {
 public:
  explicit SpanMap(size_t expected) {
    size_t cap = 16;
    while (cap < 2 * expected + 2) cap <<= 1;
    keys_.assign(cap, Span{nullptr, 0});
    vals_.assign(cap, -1);
    mask_ = cap - 1;
  }
  // returns the value already stored for `k`, or stores `v` and returns -1
  int insert(Span k, int v) {
    size_t i = hash(k) & mask_;
    while (vals_[i] >= 0) {
      if (keys_[i].n == k.n && memcmp(keys_[i].p, k.p, k.n) == 0) return vals_[i];
      i = (i + 1) & mask_;
    }
    keys_[i] = k;
    vals_[i] = v;
    return -1;
  }
  void set(Span k, int v) {  // insert or overwrite (dict semantics: the last one wins)
    size_t i = hash(k) & mask_;
    while (vals_[i] >= 0) {
      if (keys_[i].n == k.n && memcmp(keys_[i].p, k.p, k.n) == 0) { vals_[i] = v; return; }
      i = (i + 1) & mask_;
    }
    keys_[i] = k;
    vals_[i] = v;
  }
  int find(Span k) const {
    size_t i = hash(k) & mask_;
    while (vals_[i] >= 0) {
      if (keys_[i].n == k.n && memcmp(keys_[i].p, k.p, k.n) == 0) return vals_[i];
      i = (i + 1) & mask_;
    }
    return -1;
  }

 private:
  static size_t hash(Span k) {
    uint64_t h = 1469598103934665603ull;  // FNV-1a
    for (size_t i = 0; i < k.n; ++i) h = (h ^ (unsigned char)k.p[i]) * 1099511628211ull;
    return (size_t)(h ^ (h >> 29));
  }
  std::vector<Span> keys_;
  std::vector<int> vals_;
  size_t mask_;
};

struct TreeData {
  std::string text;         // the file; names are spans into it
  std::vector<Span> names;
  std::vector<int> first_child, next_sibling, last_child;  // children in edge-line order
  std::vector<double> length;                              // of the edge to the parent
  std::vector<int> parent;
  std::vector<int> n_children;
};

// "u v length" with exactly two single spaces (Python's line.split(" ") giving 3 tokens)
bool split3(Span line, Span* u, Span* v, Span* w) {
  const char* sp1 = (const char*)memchr(line.p, ' ', line.n);
  if (!sp1) return false;
  const size_t rest = line.n - (size_t)(sp1 + 1 - line.p);
  const char* sp2 = (const char*)memchr(sp1 + 1, ' ', rest);
  if (!sp2) return false;
  const size_t rest2 = line.n - (size_t)(sp2 + 1 - line.p);
  if (memchr(sp2 + 1, ' ', rest2)) return false;
  *u = Span{line.p, (size_t)(sp1 - line.p)};
  *v = Span{sp1 + 1, (size_t)(sp2 - sp1 - 1)};
  *w = Span{sp2 + 1, rest2};
  return true;
}

void parse_tree(const std::string& path, bool f32, TreeData* tp) {
  TreeData& t = *tp;
  t.text = read_file(path);
  const std::vector<Span> lines = strip_lines(t.text);
  const long long n = header_count(lines[0], "nodes", "Tree file: " + path + " should start with '[num_nodes] nodes'");
  if ((long long)lines.size() < n + 2) die("Tree file: " + path + " is truncated");
  SpanMap index((size_t)n);
  t.names.reserve((size_t)n);
  for (long long i = 1; i <= n; ++i)
    if (index.insert(lines[(size_t)i], (int)t.names.size()) < 0)  // a repeated name is one node
      t.names.push_back(lines[(size_t)i]);
  const size_t nn = t.names.size();
  t.first_child.assign(nn, -1);
  t.next_sibling.assign(nn, -1);
  t.last_child.assign(nn, -1);
  t.length.assign(nn, 0.0);
  t.parent.assign(nn, -1);
  t.n_children.assign(nn, 0);
  const long long m = header_count(lines[(size_t)n + 1], "edges",
                                   "Tree file: " + path + " should have line '[num_edges] edges' at position " +
                                       std::to_string(n + 1));
  if ((long long)lines.size() != n + m + 2)
    die("Tree file: " + path + " should have " + std::to_string(m) + " edges, but it has " +
        std::to_string((long long)lines.size() - n - 2) + " edges instead.");
  for (long long i = n + 2; i < n + 2 + m; ++i) {
    Span su, sv, sw;
    double len = 0.0;
    if (!split3(lines[(size_t)i], &su, &sv, &sw) || !parse_double(sw, f32, &len))
      die("Tree file: " + path + " should have line '[u] [v] [length]' at position " + std::to_string(i) +
          ", but it had line: '" + lines[(size_t)i].str() + "'");
    const int u = index.find(su), v = index.find(sv);
    if (u < 0 || v < 0)
      die("In Tree file " + path + ": " + su.str() + " and " + sv.str() + " should be nodes in the tree");
    if (t.parent[(size_t)v] >= 0)
      die("Node " + sv.str() + " already has a parent, cannot also have parent " + su.str() +
          " - graph is not a tree.");
    t.parent[(size_t)v] = u;
    t.length[(size_t)v] = len;
    if (t.last_child[(size_t)u] < 0) t.first_child[(size_t)u] = v; else t.next_sibling[(size_t)t.last_child[(size_t)u]] = v;
    t.last_child[(size_t)u] = v;
    ++t.n_children[(size_t)u];
  }
}

struct Pair {
  int a, b;  // node indices
  double t;
};

std::vector<Pair> extract_pairs(const TreeData& t, const std::string& mode, const std::string& path) {
  std::vector<Pair> pairs;
  const int n = (int)t.names.size();
  if (mode == "cherry++") {
    int root = -1, n_roots = 0;
    for (int v = 0; v < n; ++v)
      if (t.parent[(size_t)v] < 0) {
        if (n_roots++ == 0) root = v;
      }
    if (n_roots != 1) die("Tree " + path + " should have one root, but found " + std::to_string(n_roots));
    // iterative post-order; res[v] = (unmatched leaf or -1, its distance to v)
    std::vector<int> res_leaf((size_t)n, -1);
    std::vector<double> res_dist((size_t)n, 0.0);
    std::vector<std::pair<int, bool>> stack;
    stack.push_back({root, false});
    std::vector<int> leaves_under, kids;
    std::vector<double> dists_under;
    pairs.reserve((size_t)n / 4 + 1);
    while (!stack.empty()) {
      auto [v, expanded] = stack.back();
      stack.pop_back();
      if (t.first_child[(size_t)v] < 0) {
        res_leaf[(size_t)v] = v;
        res_dist[(size_t)v] = 0.0;
        continue;
      }
      if (!expanded) {
        stack.push_back({v, true});
        kids.clear();
        for (int c = t.first_child[(size_t)v]; c >= 0; c = t.next_sibling[(size_t)c]) kids.push_back(c);
        for (auto it = kids.rbegin(); it != kids.rend(); ++it) stack.push_back({*it, false});
        continue;
      }
      leaves_under.clear();
      dists_under.clear();
      for (int c = t.first_child[(size_t)v]; c >= 0; c = t.next_sibling[(size_t)c]) {
        if (res_leaf[(size_t)c] >= 0) {
          leaves_under.push_back(res_leaf[(size_t)c]);
          dists_under.push_back(res_dist[(size_t)c] + t.length[(size_t)c]);
        }
      }
      for (size_t i = 0; i + 1 < leaves_under.size(); i += 2)
        pairs.push_back(Pair{leaves_under[i], leaves_under[i + 1], dists_under[i] + dists_under[i + 1]});
      if (leaves_under.size() % 2 == 0) {
        res_leaf[(size_t)v] = -1;
      } else {
        res_leaf[(size_t)v] = leaves_under.back();
        res_dist[(size_t)v] = dists_under.back();
      }
    }
    size_t n_leaves = 0;
    for (int v = 0; v < n; ++v) n_leaves += t.first_child[(size_t)v] < 0 ? 1 : 0;
    if (pairs.size() != n_leaves / 2)
      die("cherry++ produced " + std::to_string(pairs.size()) + " pairs for " + std::to_string(n_leaves) + " leaves");
  } else if (mode == "cherry") {
    for (int v = 0; v < n; ++v) {
      if (t.n_children[(size_t)v] != 2) continue;
      const int c0 = t.first_child[(size_t)v], c1 = t.next_sibling[(size_t)c0];
      if (t.first_child[(size_t)c0] < 0 && t.first_child[(size_t)c1] < 0)
        pairs.push_back(Pair{c0, c1, t.length[(size_t)c0] + t.length[(size_t)c1]});
    }
  } else if (mode == "edge") {
    for (int v = 0; v < n; ++v)
      for (int c = t.first_child[(size_t)v]; c >= 0; c = t.next_sibling[(size_t)c])
        pairs.push_back(Pair{v, c, t.length[(size_t)c]});
  } else {
    die("Unknown edge_or_cherry: '" + mode + "'");
  }
  return pairs;
}

struct MsaData {
  std::string text;
  std::vector<Span> lines;  // name / sequence lines alternate
  SpanMap index{0};         // name -> index of its sequence line
};

void parse_msa(const std::string& path, MsaData* m) {
  m->text = read_file(path);
  m->lines = strip_lines(m->text);
  const std::vector<Span>& lines = m->lines;
  if (lines.size() % 2 != 0) die("The MSA at " + path + " should have an even number of lines");
  m->index = SpanMap(lines.size() / 2);
  for (size_t i = 0; i + 1 < lines.size(); i += 2) {
    if (lines[i].n == 0 || lines[i].p[0] != '>')
      die("MSA at " + path + ": at line " + std::to_string(i) + " expected '>[seq_name]' but found " + lines[i].str());
    m->index.set(Span{lines[i].p + 1, lines[i].n - 1}, (int)(i + 1));
  }
}

std::vector<double> parse_site_rates(const std::string& path) {
  const std::string text = read_file(path);
  const std::vector<Span> lines = strip_lines(text);
  const long long n = header_count(lines[0], "sites", "Site rates file: " + path +
                                                         " should start with line '[num_sites] sites'");
  std::vector<double> out;
  if (lines.size() < 2) {
    if (n != 0) die("Could nor read site rates in file: " + path);
    return out;
  }
  for (const Span& tok : split_space(lines[1])) {
    double v;
    if (!parse_double(tok, false, &v)) die("Could nor read site rates in file: " + path);
    out.push_back(v);
  }
  if ((long long)out.size() != n)
    die("Site rates file: " + path + " was supposed to have " + std::to_string(n) + " sites, but it has " +
        std::to_string(out.size()));
  return out;
}

// (i, j) with map[i][j] == '1', i < j, j - i >= min_dist, row-major
std::vector<int32_t> parse_contacts(const std::string& path, int min_dist, long long* n_sites) {
  const std::string text = read_file(path);
  const std::vector<Span> lines = strip_lines(text);
  const long long n = header_count(lines[0], "sites", "Contact map file should start with line '[num_sites] sites'");
  if ((long long)lines.size() != n + 1)
    die("Contact Map at: " + path + " should have " + std::to_string(n) + " rows, but has " +
        std::to_string(lines.size() - 1));
  size_t total = 0;
  for (long long i = 1; i <= n; ++i) total += lines[(size_t)i].n;
  if (total != (size_t)(n * n)) die("Contact Map at: " + path + " is not square");
  std::vector<int32_t> out;
  // the reference joins the rows before reshaping: index the concatenation
  long long flat = 0;
  for (long long r = 1; r <= n; ++r) {
    const Span& ln = lines[(size_t)r];
    for (size_t c = 0; c < ln.n; ++c, ++flat) {
      if (ln.p[c] != '1') continue;
      const long long i = flat / n, j = flat % n;
      if (j > i && j - i >= min_dist) {
        out.push_back((int32_t)i);
        out.push_back((int32_t)j);
      }
    }
  }
  *n_sites = n;
  return out;
}

struct FamilyOut {
  std::vector<uint8_t> rows;  // n_rows * stride
  int n_rows = 0, stride = 16;
  std::vector<int32_t> pair_a, pair_b;
  std::vector<double> pair_t;
  std::vector<double> rate_vals;
  std::vector<uint16_t> group_cat;  // LG
  std::vector<int32_t> contacts;    // co: flattened (i, j)
  int aux_cnt = 0;
  long long items_per_pair = 0;
};

struct Job {
  int kind;  // 0 LG, 1 co
  std::string tree_dir, msa_dir, third_dir, mode;
  std::vector<std::string> families;
  uint8_t lut[256];
  int S;
  bool f32;
  int min_dist;
};

// rows (node indices in order of first use) + pair row indices
void rows_for_pairs(const std::vector<Pair>& pairs, int n_nodes, std::vector<int>* row_nodes, FamilyOut* out) {
  std::vector<int> row_of((size_t)n_nodes, -1);
  out->pair_a.resize(pairs.size());
  out->pair_b.resize(pairs.size());
  out->pair_t.resize(pairs.size());
  for (size_t i = 0; i < pairs.size(); ++i) {
    for (int side = 0; side < 2; ++side) {
      const int node = side ? pairs[i].b : pairs[i].a;
      if (row_of[(size_t)node] < 0) {
        row_of[(size_t)node] = (int)row_nodes->size();
        row_nodes->push_back(node);
      }
      (side ? out->pair_b : out->pair_a)[i] = row_of[(size_t)node];
    }
    out->pair_t[i] = pairs[i].t;
  }
}

void process_family(const Job& job, const std::string& fam, FamilyOut* out) {
  const std::string tree_path = job.tree_dir + "/" + fam + ".txt";
  TreeData tree;
  parse_tree(tree_path, job.f32, &tree);
  const std::vector<Pair> pairs = extract_pairs(tree, job.mode, tree_path);
  std::vector<int> row_nodes;
  rows_for_pairs(pairs, (int)tree.names.size(), &row_nodes, out);
  MsaData msa;
  parse_msa(job.msa_dir + "/" + fam + ".txt", &msa);
  // sequences of the rows
  std::vector<Span> seqs(row_nodes.size());
  for (size_t r = 0; r < row_nodes.size(); ++r) {
    const int li = msa.index.find(tree.names[(size_t)row_nodes[r]]);
    if (li < 0)
      die("Family " + fam + ": node '" + tree.names[(size_t)row_nodes[r]].str() + "' of the tree is not in the MSA");
    seqs[r] = msa.lines[(size_t)li];
  }
  const size_t n_rows = seqs.size();
  const size_t L_msa = n_rows ? seqs[0].n : 0;
  for (const Span& s : seqs)
    if (s.n != L_msa) die("Family " + fam + ": sequences in the MSA have different lengths");
  out->n_rows = (int)n_rows;
  const uint8_t skip = (uint8_t)job.S;
  if (job.kind == 0) {
    std::vector<double> rates = parse_site_rates(job.third_dir + "/" + fam + ".txt");
    if (n_rows && L_msa > rates.size())
      die("Family " + fam + ": MSA has " + std::to_string(L_msa) + " sites but there are only " +
          std::to_string(rates.size()) + " site rates");
    if (n_rows) rates.resize(L_msa);  // the reference indexes site_rates by MSA position
    const size_t L = rates.size();
    out->items_per_pair = (long long)L;
    if (L == 0) {
      out->rate_vals = {1.0};
      out->group_cat.assign(4, 0);
      out->stride = 16;
      out->rows.assign(n_rows * 16, skip);
    } else {
      // categories = distinct rate values in ascending order; columns sorted by category
      // (stable), every category padded to a multiple of 4 sites
      std::vector<double> vals(rates);
      std::sort(vals.begin(), vals.end());
      vals.erase(std::unique(vals.begin(), vals.end()), vals.end());
      if (vals.size() > 65535) die("more than 65535 distinct site rates in one family");
      std::vector<int> cat(L);
      std::vector<long long> counts(vals.size(), 0);
      for (size_t j = 0; j < L; ++j) {
        cat[j] = (int)(std::lower_bound(vals.begin(), vals.end(), rates[j]) - vals.begin());
        ++counts[(size_t)cat[j]];
      }
      std::vector<long long> starts(vals.size(), 0);
      long long total = 0;
      for (size_t c = 0; c < vals.size(); ++c) {
        starts[c] = total;
        total += (counts[c] + 3) / 4 * 4;
      }
      const long long stride = std::max<long long>(16, (total + 15) / 16 * 16);
      std::vector<long long> next(starts);
      std::vector<int> dest(L);
      for (size_t j = 0; j < L; ++j) dest[j] = (int)next[(size_t)cat[j]]++;
      out->stride = (int)stride;
      out->group_cat.assign((size_t)stride / 4, 0);
      for (size_t c = 0; c < vals.size(); ++c)
        for (long long g = starts[c] / 4; g < (starts[c] + (counts[c] + 3) / 4 * 4) / 4; ++g)
          out->group_cat[(size_t)g] = (uint16_t)c;
      out->rate_vals = std::move(vals);
      out->rows.assign(n_rows * (size_t)stride, skip);
      // four rows per pass over the column permutation (dest[] is loaded once per four stores)
      size_t r = 0;
      for (; r + 4 <= n_rows; r += 4) {
        uint8_t* row = out->rows.data() + r * (size_t)stride;
        const unsigned char* s0 = reinterpret_cast<const unsigned char*>(seqs[r].p);
        const unsigned char* s1 = reinterpret_cast<const unsigned char*>(seqs[r + 1].p);
        const unsigned char* s2 = reinterpret_cast<const unsigned char*>(seqs[r + 2].p);
        const unsigned char* s3 = reinterpret_cast<const unsigned char*>(seqs[r + 3].p);
        for (size_t j = 0; j < L; ++j) {
          uint8_t* d = row + dest[j];
          d[0] = job.lut[s0[j]];
          d[(size_t)stride] = job.lut[s1[j]];
          d[2 * (size_t)stride] = job.lut[s2[j]];
          d[3 * (size_t)stride] = job.lut[s3[j]];
        }
      }
      for (; r < n_rows; ++r) {
        uint8_t* row = out->rows.data() + r * (size_t)stride;
        const unsigned char* s = reinterpret_cast<const unsigned char*>(seqs[r].p);
        for (size_t j = 0; j < L; ++j) row[dest[j]] = job.lut[s[j]];
      }
    }
    out->aux_cnt = out->stride / 4;
  } else {
    long long n_sites = 0;
    out->contacts = parse_contacts(job.third_dir + "/" + fam + ".txt", job.min_dist, &n_sites);
    const size_t P = out->contacts.size() / 2;
    const size_t L = n_rows ? L_msa : (size_t)n_sites;
    for (size_t c = 0; c < 2 * P; ++c)
      if ((size_t)out->contacts[c] >= L) die("Family " + fam + ": contact map is larger than the MSA");
    const size_t stride = std::max<size_t>(16, (2 * P + 15) / 16 * 16);
    out->stride = (int)stride;
    out->rows.assign(n_rows * stride, skip);
    for (size_t r = 0; r < n_rows; ++r) {
      uint8_t* row = out->rows.data() + r * stride;
      const unsigned char* s = reinterpret_cast<const unsigned char*>(seqs[r].p);
      for (size_t c = 0; c < P; ++c) {
        row[2 * c] = job.lut[s[out->contacts[2 * c]]];
        row[2 * c + 1] = job.lut[s[out->contacts[2 * c + 1]]];
      }
    }
    out->rate_vals = {1.0};
    out->aux_cnt = (int)P;
    out->items_per_pair = (long long)P;
  }
}

template <typename T>
T* alloc_array(size_t n) {
  void* p = nullptr;
  if (posix_memalign(&p, 64, std::max<size_t>(64, n * sizeof(T))) != 0) return nullptr;
  return reinterpret_cast<T*>(p);
}

int run_ingest(Job& job, int n_threads, int pinned, cherry_ingest_result** out_ptr) {
  const int F = (int)job.families.size();
  cherry::keep_large_buffers_on_heap();
  std::vector<FamilyOut> fam_out((size_t)F);
  std::atomic<int> next{0};
  std::atomic<bool> failed{false};
  std::string first_error;
  std::atomic_flag err_lock = ATOMIC_FLAG_INIT;
  auto worker = [&]() {
    for (;;) {
      const int f = next.fetch_add(1);
      if (f >= F || failed.load()) return;
      try {
        process_family(job, job.families[(size_t)f], &fam_out[(size_t)f]);
      } catch (const Err& e) {
        if (!failed.exchange(true)) {
          while (err_lock.test_and_set()) {}
          first_error = e.msg;
          err_lock.clear();
        }
        return;
      } catch (const std::exception& e) {
        if (!failed.exchange(true)) {
          while (err_lock.test_and_set()) {}
          first_error = std::string("ingest: ") + e.what();
          err_lock.clear();
        }
        return;
      }
    }
  };
  if (n_threads < 1) n_threads = 1;
  if (n_threads > F) n_threads = F > 0 ? F : 1;
  std::vector<std::thread> threads;
  for (int i = 1; i < n_threads; ++i) threads.emplace_back(worker);
  worker();
  for (auto& t : threads) t.join();
  if (failed.load()) return cherry::fail(CHERRY_EINVAL, "%s", first_error.c_str());

  // ---- concatenate (offsets first, then parallel copies)
  cherry_ingest_result* R = reinterpret_cast<cherry_ingest_result*>(calloc(1, sizeof(cherry_ingest_result)));
  if (!R) return cherry::fail(CHERRY_EINVAL, "ingest: out of memory");
  R->kind = job.kind;
  R->n_fams = F;
  int64_t msa_bytes = 0, n_pairs = 0, n_rates = 0, n_aux = 0, n_tiles = 0, examined = 0;
  int max_rates = 1, max_stride = 16;
  std::vector<int64_t> off_msa((size_t)F), off_pair((size_t)F), off_rate((size_t)F), off_aux((size_t)F),
      off_tile((size_t)F);
  std::vector<int> per_tile((size_t)F);
  for (int f = 0; f < F; ++f) {
    const FamilyOut& o = fam_out[(size_t)f];
    off_msa[(size_t)f] = msa_bytes; off_pair[(size_t)f] = n_pairs; off_rate[(size_t)f] = n_rates;
    off_aux[(size_t)f] = n_aux; off_tile[(size_t)f] = n_tiles;
    msa_bytes += (int64_t)o.rows.size();
    const int64_t np = (int64_t)o.pair_a.size();
    n_pairs += np;
    n_rates += (int64_t)o.rate_vals.size();
    n_aux += job.kind == 0 ? (int64_t)o.group_cat.size() : (int64_t)o.contacts.size() / 2;
    per_tile[(size_t)f] = job.kind == 0 ? std::max(1, kTargetChunksPerTile / std::max(1, o.stride / 16))
                                        : std::max(1, kTargetItemsPerCoTile / std::max(1, o.aux_cnt));
    n_tiles += (np + per_tile[(size_t)f] - 1) / per_tile[(size_t)f];
    examined += np * o.items_per_pair;
    max_rates = std::max(max_rates, (int)o.rate_vals.size());
    max_stride = std::max(max_stride, o.stride);
  }
  if (n_pairs > 0x7fffffff || n_aux > 0x7fffffff || n_rates > 0x7fffffff || n_tiles > 0x7fffffff) {
    free(R);
    return cherry::fail(CHERRY_ELIMIT, "ingest: batch too large for 32-bit indices (split the families)");
  }
  R->msa_bytes = std::max<int64_t>(16, msa_bytes);
  R->pinned = 0;
  if (pinned) {
    void* p = cherry::pinned_alloc((size_t)R->msa_bytes);  // nullptr without a device: pageable memory
    if (p) {
      R->msa = reinterpret_cast<uint8_t*>(p);
      R->pinned = 1;
    }
  }
  if (!R->msa) R->msa = alloc_array<uint8_t>((size_t)R->msa_bytes);
  R->fams = alloc_array<cherry_fam_desc>((size_t)F);
  R->n_pairs = n_pairs;
  R->pair_a = alloc_array<int32_t>((size_t)n_pairs);
  R->pair_b = alloc_array<int32_t>((size_t)n_pairs);
  R->pair_t = alloc_array<double>((size_t)n_pairs);
  R->pair_fam = alloc_array<int32_t>((size_t)n_pairs);
  R->n_rate_vals = n_rates;
  R->rate_vals = alloc_array<double>((size_t)n_rates);
  R->n_aux = n_aux;
  R->aux = job.kind == 0 ? (void*)alloc_array<uint16_t>((size_t)n_aux) : (void*)alloc_array<int32_t>((size_t)n_aux * 2);
  R->n_tiles = (int)n_tiles;
  R->tiles = alloc_array<cherry_tile>((size_t)n_tiles);
  R->r_pad = job.kind == 0 ? (max_rates + 3) / 4 * 4 : 4;
  R->n_items_examined = examined;
  R->max_row_stride = max_stride;
  if (!R->msa || !R->fams || !R->pair_a || !R->pair_b || !R->pair_t || !R->pair_fam || !R->rate_vals || !R->aux ||
      !R->tiles) {
    cherry_ingest_free(R);
    return cherry::fail(CHERRY_EINVAL, "ingest: out of memory");
  }
  if (msa_bytes == 0) memset(R->msa, 0, 16);
  std::atomic<int> next2{0};
  auto copier = [&]() {
    for (;;) {
      const int f = next2.fetch_add(1);
      if (f >= F) return;
      const FamilyOut& o = fam_out[(size_t)f];
      cherry_fam_desc& d = R->fams[f];
      d.msa_off = off_msa[(size_t)f];
      d.row_stride = o.stride;
      d.n_chunks = o.stride / 16;
      d.aux_off = (int32_t)off_aux[(size_t)f];
      d.aux_cnt = o.aux_cnt;
      d.rate_off = (int32_t)off_rate[(size_t)f];
      d.n_rates = (int32_t)o.rate_vals.size();
      if (!o.rows.empty()) memcpy(R->msa + off_msa[(size_t)f], o.rows.data(), o.rows.size());
      const size_t np = o.pair_a.size(), po = (size_t)off_pair[(size_t)f];
      if (np) {
        memcpy(R->pair_a + po, o.pair_a.data(), np * sizeof(int32_t));
        memcpy(R->pair_b + po, o.pair_b.data(), np * sizeof(int32_t));
        memcpy(R->pair_t + po, o.pair_t.data(), np * sizeof(double));
        for (size_t i = 0; i < np; ++i) R->pair_fam[po + i] = f;
      }
      memcpy(R->rate_vals + off_rate[(size_t)f], o.rate_vals.data(), o.rate_vals.size() * sizeof(double));
      if (job.kind == 0) {
        memcpy(reinterpret_cast<uint16_t*>(R->aux) + off_aux[(size_t)f], o.group_cat.data(),
               o.group_cat.size() * sizeof(uint16_t));
      } else if (!o.contacts.empty()) {
        memcpy(reinterpret_cast<int32_t*>(R->aux) + 2 * off_aux[(size_t)f], o.contacts.data(),
               o.contacts.size() * sizeof(int32_t));
      }
      int64_t ti = off_tile[(size_t)f];
      for (size_t b = 0; b < np; b += (size_t)per_tile[(size_t)f]) {
        cherry_tile& tl = R->tiles[ti++];
        tl.fam = f;
        tl.pair_begin = (int32_t)(po + b);
        tl.n_pairs = (int32_t)std::min<size_t>((size_t)per_tile[(size_t)f], np - b);
        tl.reserved = 0;
      }
    }
  };
  threads.clear();
  for (int i = 1; i < n_threads; ++i) threads.emplace_back(copier);
  copier();
  for (auto& t : threads) t.join();
  *out_ptr = R;
  return 0;
}

int fill_job(Job& job, int kind, const char* tree_dir, const char* msa_dir, const char* third_dir,
             const char* const* families, int n_fams, const char* const* states, int n_states, const char* mode,
             int f32, int min_dist) {
  if (!tree_dir || !msa_dir || !third_dir || (!families && n_fams > 0) || !states || !mode)
    return cherry::fail(CHERRY_EINVAL, "ingest: null pointer argument");
  if (n_fams < 0 || n_states <= 0 || n_states > 254) return cherry::fail(CHERRY_EINVAL, "ingest: bad sizes");
  job.kind = kind;
  job.tree_dir = tree_dir;
  job.msa_dir = msa_dir;
  job.third_dir = third_dir;
  job.mode = mode;
  if (job.mode.rfind("cherry++__", 0) == 0) job.mode = "cherry++";
  job.S = n_states;
  job.f32 = f32 != 0;
  job.min_dist = min_dist;
  memset(job.lut, n_states, sizeof(job.lut));  // every byte that is not a state is the skip code
  for (int i = 0; i < n_states; ++i) {
    if (!states[i] || strlen(states[i]) != 1)
      return cherry::fail(CHERRY_EINVAL, "ingest: states must be single one-byte characters");
    job.lut[(unsigned char)states[i][0]] = (uint8_t)i;
  }
  job.families.reserve((size_t)n_fams);
  for (int i = 0; i < n_fams; ++i) job.families.emplace_back(families[i]);
  return 0;
}

}  // namespace

extern "C" {

int cherry_ingest_lg(const char* tree_dir, const char* msa_dir, const char* site_rates_dir,
                     const char* const* families, int n_fams, const char* const* states, int n_states,
                     const char* edge_or_cherry, int float32_branch_lengths, int n_threads, int pinned,
                     cherry_ingest_result** out) {
  if (!out) return cherry::fail(CHERRY_EINVAL, "ingest_lg: null out");
  Job job;
  int rc = fill_job(job, 0, tree_dir, msa_dir, site_rates_dir, families, n_fams, states, n_states,
                    edge_or_cherry, float32_branch_lengths, 0);
  if (rc) return rc;
  return run_ingest(job, n_threads, pinned, out);
}

int cherry_ingest_co(const char* tree_dir, const char* msa_dir, const char* contact_map_dir,
                     const char* const* families, int n_fams, const char* const* states, int n_states,
                     const char* edge_or_cherry, int minimum_distance, int float32_branch_lengths,
                     int n_threads, int pinned, cherry_ingest_result** out) {
  if (!out) return cherry::fail(CHERRY_EINVAL, "ingest_co: null out");
  Job job;
  int rc = fill_job(job, 1, tree_dir, msa_dir, contact_map_dir, families, n_fams, states, n_states,
                    edge_or_cherry, float32_branch_lengths, minimum_distance);
  if (rc) return rc;
  return run_ingest(job, n_threads, pinned, out);
}

void cherry_ingest_free(cherry_ingest_result* r) {
  if (!r) return;
  if (r->msa) {
    if (r->pinned) cherry::pinned_free(r->msa); else free(r->msa);
  }
  free(r->fams); free(r->pair_a); free(r->pair_b); free(r->pair_t); free(r->pair_fam);
  free(r->rate_vals); free(r->aux); free(r->tiles);
  free(r);
}

}  // extern "C"
