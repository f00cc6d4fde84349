// C++ host-side mirror test: a 3D 7-point Poisson problem through include/ngsamg_b200.hpp (the way code written against the
// reference's BaseAMGPC / CGSolver would look).  Exit code 0 = pass, 3 = "no CUDA device" (expected on the CPU-only box).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "ngsamg_b200.hpp"

int main()
{
  const int N = 24;
  const int64_t n = (int64_t)N * N * N;
  std::vector<int64_t> rp(n + 1, 0);
  std::vector<int32_t> ci;
  std::vector<double> v;
  std::vector<uint8_t> freed(n, 1);
  auto id = [&](int x, int y, int z) { return (int32_t)(x + N * (y + N * z)); };
  for (int z = 0; z < N; z++)
    for (int y = 0; y < N; y++)
      for (int x = 0; x < N; x++) {
        const int64_t i = id(x, y, z);
        if (x == 0) freed[i] = 0;
        const int d[7][3] = {{0, 0, -1}, {0, -1, 0}, {-1, 0, 0}, {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (auto &o : d) {
          const int a = x + o[0], b = y + o[1], c = z + o[2];
          if (a < 0 || b < 0 || c < 0 || a >= N || b >= N || c >= N) continue;
          ci.push_back(id(a, b, c));
          v.push_back((o[0] == 0 && o[1] == 0 && o[2] == 0) ? 6.0 : -1.0);
        }
        rp[i + 1] = (int64_t)ci.size();
      }
  amg::SparseMat A(n, n, 1, 1, rp, ci, v);
  try {
    amg::BaseAMGPC pc("NgsAMG.h1_scal", A, &freed, {{"ngs_amg_max_coarse_size", "30"}});
    pc.FinalizeLevel();
    std::vector<double> b(n, 1.0), x(n, 0.0), y(n, 1.0);
    for (int64_t i = 0; i < n; i++) if (!freed[i]) b[i] = 0.0;
    pc.Mult(b.data(), x.data());
    pc.MultAdd(2.0, b.data(), y.data());                       // y = 1 + 2 C b
    double e = 0;
    for (int64_t i = 0; i < n; i++) e = std::fmax(e, std::fabs(y[i] - 1.0 - 2.0 * x[i]));
    amg::CGSolver cg(A, pc, 60, 1e-8);
    std::vector<double> u(n, 0.0);
    cg.Solve(b.data(), u.data());
    // true residual on the free rows
    double rn = 0, bn = 0;
    for (int64_t i = 0; i < n; i++) {
      if (!freed[i]) continue;
      double r = b[i];
      for (int64_t k = rp[i]; k < rp[i + 1]; k++) r -= v[k] * u[ci[k]];
      rn += r * r; bn += b[i] * b[i];
    }
    auto Ac = amg::RestrictMatrix(A, pc.GetProlongation(0));
    auto A1 = pc.GetLevelMatrix(1);
    bool same = Ac.rowptr == A1.rowptr && Ac.col == A1.col;
    std::printf("levels=%zu OC=%.3f multadd_err=%.2e cg_its=%d rel_res=%.2e rap_pattern_same=%d\n", pc.GetNLevels(), pc.GetOC(), e,
                cg.iterations, std::sqrt(rn / bn), (int)same);
    bool threw = false;
    try { pc.GetSmoother((int)pc.GetNLevels() - 1); } catch (const amg::Exception &) { threw = true; }
    if (!(e < 1e-12 && cg.iterations < 40 && std::sqrt(rn / bn) < 1e-5 && same && threw)) return 1;
  } catch (const amg::Exception &ex) {
    std::printf("amg::Exception: %s\n", ex.what());
    return std::strstr(ex.what(), "no CUDA device") ? 3 : 2;
  }
  return 0;
}
