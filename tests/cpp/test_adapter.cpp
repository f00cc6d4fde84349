// test_adapter.cpp -- the reference-side adapter (include/ngsamg_b200_ngsolve.hpp) compiled against the NGSolve stand-in and driven the
// way NGSolve drives a registered preconditioner: look "NgsAMG.h1_scal" up in the registry, InitLevel(freedofs), FinalizeLevel(mat),
// then Mult / MultAdd / MultTrans through the BaseMatrix interface and a PCG.  `test_adapter --list` only prints the registered names
// (no device needed); without arguments it needs a GPU and checks the adapter against direct C-ABI calls.
#include "ngcomp_standin.hpp"
#include "ngsamg_b200_ngsolve.hpp"

#include <cmath>
#include <cstring>

NGSAMG_B200_REGISTER_PRECONDITIONERS();

using namespace ngla;

static std::shared_ptr<SparseMatrix<double>> laplace3d(int n, std::shared_ptr<ngcore::BitArray> &free)
{
  const size_t N = size_t(n) * n * n;
  auto id = [n](int i, int j, int k) { return (size_t(k) * n + j) * n + i; };
  ngcore::Array<int> per(N);
  for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    int c = 1;
    c += (i > 0) + (i < n - 1) + (j > 0) + (j < n - 1) + (k > 0) + (k < n - 1);
    per[id(i, j, k)] = c;
  }
  auto A = std::make_shared<SparseMatrix<double>>(per);
  free = std::make_shared<ngcore::BitArray>(N);
  for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    const size_t r = id(i, j, k);
    auto cols = A->GetRowIndices(r);
    auto vals = A->GetRowValues(r);
    int p = 0;
    auto put = [&](size_t c, double v) { cols[p] = int(c); vals(p) = v; p++; };
    if (k > 0) put(id(i, j, k - 1), -1.0);
    if (j > 0) put(id(i, j - 1, k), -1.0);
    if (i > 0) put(id(i - 1, j, k), -1.0);
    put(r, 6.0);
    if (i < n - 1) put(id(i + 1, j, k), -1.0);
    if (j < n - 1) put(id(i, j + 1, k), -1.0);
    if (k < n - 1) put(id(i, j, k + 1), -1.0);
    if (i > 0) free->SetBit(r);                          // Dirichlet on the face i = 0
  }
  return A;
}

static double norm(const BaseVector &v) { double s = 0; auto f = v.FVDouble(); for (size_t i = 0; i < f.Size(); i++) s += f(i) * f(i); return std::sqrt(s); }

int main(int argc, char **argv)
{
  auto &reg = ngcomp::GetPreconditionerClasses();
  if (argc > 1 && !std::strcmp(argv[1], "--list")) {
    for (auto &e : reg) std::printf("%s\n", e.first.c_str());
    return (reg.count("NgsAMG.h1_scal") && reg.count("NgsAMG.elast_3d")) ? 0 : 1;
  }
  try {
    const int n = 17;
    std::shared_ptr<ngcore::BitArray> free;
    auto A = laplace3d(n, free);
    ngcore::Flags flags;
    flags.SetFlag("ngs_amg_max_coarse_size", 20.0).SetFlag("ngs_amg_sm_type", "gs").SetFlag("ngs_amg_some_unknown_flag", true);
    auto pc = reg.at("NgsAMG.h1_scal")(std::make_shared<ngcomp::BilinearForm>(), flags, "b200");
    pc->InitLevel(free);
    pc->FinalizeLevel(A.get());
    if (&pc->GetAMatrix() != A.get()) { std::printf("GetAMatrix is not the finest matrix\n"); return 1; }
    const size_t N = A->Height();
    auto b = pc->CreateColVector(), x = pc->CreateRowVector(), y = pc->CreateRowVector(), z = pc->CreateRowVector();
    auto fb = b->FVDouble();
    for (size_t i = 0; i < N; i++) fb(i) = free->Test(i) ? std::sin(0.37 * double(i)) : 0.0;
    pc->Mult(*b, *x);                                    // through ngcomp::Preconditioner
    pc->GetMatrix().Mult(*b, *y);                        // through the AMGMatrix
    auto pcb = std::dynamic_pointer_cast<amg::H1ScalB200>(pc);
    // the same cycle through the C ABI directly
    std::vector<double> direct(N, 0.0);
    if (ngsamg_b200_apply(pcb->GetAMGMatrix()->Handle(), fb.Data(), direct.data())) { std::printf("%s\n", ngsamg_b200_last_error()); return 1; }
    double d1 = 0, d2 = 0;
    for (size_t i = 0; i < N; i++) { d1 = std::max(d1, std::fabs(x->FVDouble()(i) - direct[i])); d2 = std::max(d2, std::fabs(y->FVDouble()(i) - direct[i])); }
    // MultAdd: z = 0.5 b-image added onto x  ->  z == 1.5 x ;  MultTrans aliases Mult
    *z = *x;
    pc->MultAdd(0.5, *b, *z);
    double d3 = 0;
    for (size_t i = 0; i < N; i++) d3 = std::max(d3, std::fabs(z->FVDouble()(i) - 1.5 * x->FVDouble()(i)));
    pcb->MultTrans(*b, *y);
    double d4 = 0;
    for (size_t i = 0; i < N; i++) d4 = std::max(d4, std::fabs(y->FVDouble()(i) - x->FVDouble()(i)));
    // PCG on the device, checked by the true residual computed with the stand-in's own SpMV
    auto u = pc->CreateRowVector(), r = pc->CreateRowVector();
    const int its = pcb->SolveCG(*b, *u, 1e-10, 100);
    *r = *b;
    A->MultAdd(-1.0, *u, *r);
    auto fr = r->FVDouble();
    for (size_t i = 0; i < N; i++) if (!free->Test(i)) fr(i) = 0.0;
    const double rel = norm(*r) / norm(*b);
    std::printf("adapter: levels=%d ndof0=%zu OC=%.3f  |Mult-abi|=%.2e |AMGMatrix-abi|=%.2e |MultAdd|=%.2e |MultTrans|=%.2e  pcg its=%d true rel.res=%.2e\n",
                pcb->GetAMGMatrix()->GetNLevels(0), pcb->GetAMGMatrix()->GetNDof(0, 0), pcb->GetAMGMatrix()->GetOC(), d1, d2, d3, d4, its, rel);
    const double xn = norm(*x);
    if (!(d1 == 0.0 && d2 == 0.0 && d3 <= 1e-14 * xn && d4 == 0.0 && its > 0 && its < 40 && rel < 1e-8)) return 1;
    // error behaviour: a matrix of the wrong entry type must throw ngcore::Exception, like the reference
    bool thrown = false;
    try { auto pc3 = reg.at("NgsAMG.elast_3d")(nullptr, flags, "e"); pc3->FinalizeLevel(A.get()); } catch (const ngcore::Exception &) { thrown = true; }
    if (!thrown) { std::printf("no exception for a scalar matrix handed to elast_3d\n"); return 1; }
  } catch (const std::exception &e) {
    std::printf("exception: %s\n", e.what());
    return 2;
  }
  return 0;
}
