// common.hpp -- host-side containers and helpers shared by the hierarchy builder and the CUDA side.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace ngb {

using i64 = int64_t;
using i32 = int32_t;

struct Error : std::runtime_error {
  explicit Error(const std::string &m) : std::runtime_error(m) {}
};

// Block CSR in NGSolve SparseMatrix<Mat<bh,bw>> layout (sorted columns, row-major AoS blocks).
struct HostBsr {
  i64 nrows = 0, ncols = 0;
  int bh = 1, bw = 1;
  std::vector<i64> rowptr;
  std::vector<i32> col;
  std::vector<double> val;
  i64 nnz() const { return rowptr.empty() ? 0 : rowptr.back(); }
  int bs() const { return bh * bw; }
};

// tiny fork-join helper for the host setup code (the reference uses ngcore ParallelForRange there)
inline int host_threads()
{
  static int n = [] {
    const char *e = std::getenv("NGSAMG_B200_HOST_THREADS");
    int v = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(v, 64));
  }();
  return n;
}

template <class F>
inline void parallel_for(i64 n, F &&f, i64 grain = 4096)
{
  int nt = host_threads();
  if (nt <= 1 || n < 2 * grain) { f(i64(0), n); return; }
  nt = (int)std::min<i64>(nt, (n + grain - 1) / grain);
  std::vector<std::thread> th;
  i64 chunk = (n + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    i64 lo = t * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([&f, lo, hi] { f(lo, hi); });
  }
  for (auto &t : th) t.join();
}

// NGSolve-Flags-like option bag: "ngs_amg_" prefixed keys, per-level lists via "<key>_spec".
struct Flags {
  std::map<std::string, std::string> kv;
  static std::string strip(std::string k)
  {
    const char *pre[] = {"ngs_amg_", "NgsAMG_"};
    for (auto p : pre)
      if (k.rfind(p, 0) == 0) return k.substr(std::strlen(p));
    return k;
  }
  void set(const std::string &k, const std::string &v) { kv[strip(k)] = v; }
  bool has(const std::string &k) const { return kv.count(k) > 0; }
  std::string str(const std::string &k, const std::string &def) const
  {
    auto it = kv.find(k);
    return it == kv.end() ? def : it->second;
  }
  double num(const std::string &k, double def) const
  {
    auto it = kv.find(k);
    if (it == kv.end()) return def;
    if (it->second == "True" || it->second == "true") return 1;
    if (it->second == "False" || it->second == "false") return 0;
    return std::atof(it->second.c_str());
  }
  bool flag(const std::string &k, bool def) const { return num(k, def ? 1 : 0) != 0; }
  // SpecOpt<T> semantics (src/base/utils/SpecOpt.hpp:47-64): "<key>" is the default,
  // "<key>_spec" a list for levels 0,1,...; levels past the list fall back to the default.
  std::string spec(const std::string &k, int level, const std::string &def) const
  {
    auto it = kv.find(k + "_spec");
    if (it != kv.end()) {
      std::vector<std::string> items;
      std::string cur;
      for (char c : it->second) {
        if (c == ',' || c == ' ' || c == ';') { if (!cur.empty()) items.push_back(cur); cur.clear(); }
        else if (c != '[' && c != ']' && c != '\'' && c != '"') cur.push_back(c);
      }
      if (!cur.empty()) items.push_back(cur);
      if (level < (int)items.size()) return items[level];
    }
    return str(k, def);
  }
};

// ---- hierarchy builder (coarsen.cpp) -----------------------------------------------------------
struct CoarsenOptions {
  int max_per_row = 3;       // sp_max_per_row (H1: 3, elasticity 3d: 1+DIM)      h1_impl.hpp:319-324
  double min_frac = 0.08;    // sp_min_frac
  double omega = 1.0;        // sp_omega
  bool smooth = true;        // prol_type != "piecewise"
  int rounds = 3;            // spw_rounds (aggregates <= 2^rounds)
  double soc_thresh = 0.25;  // scalRelThresh
};

// Build the prolongation P (fine n x coarse nc) of one level from the level matrix.
// bf = fine block size, bc = coarse block size (bf==bc except elasticity level 0: 3 -> 6).
// xyz: fine vertex coordinates (n x 3) or empty; cxyz: out, coarse vertex coordinates (elasticity).
// vmap out: fine vertex -> coarse vertex (-1 = Dirichlet / dropped).
struct ParCoarsen;  // par.hpp: sharing classes for the multi-rank (class-respecting) mode, nullptr = single rank
void build_prolongation(const HostBsr &A, const uint8_t *free_mask, int bc, const std::vector<double> &xyz,
                        const CoarsenOptions &opt, HostBsr &P, std::vector<i32> &vmap, std::vector<double> &cxyz,
                        const ParCoarsen *par = nullptr);

void host_transpose(const HostBsr &A, HostBsr &T);

// dense.cpp: CalcPseudoInverseTryNormal(Mat<N,N>&) of the reference restated (utils_denseLA.hpp:1237-1569, utils_denseLA.cpp:458-555); m: n x n row-major, in place
void block_pinv(int n, double *m);
// dense.cpp: RegularizeMatrix of the elasticity preconditioners on one diagonal block of the coarsest matrix (elasticity_pc_impl.hpp:711-763)
void block_regularize(int n, double *m, int dim);

}  // namespace ngb
