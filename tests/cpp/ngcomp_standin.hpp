// ngcomp_standin.hpp -- TEST INFRASTRUCTURE: the handful of NGSolve types include/ngsamg_b200_ngsolve.hpp touches beyond what
// oracle/ref_pin/ngs_standin.hpp already models (Flags, AutoVector/VVector, ngcomp::Preconditioner, BilinearForm and the preconditioner
// registry), with NGSolve's signatures, so that the reference-side adapter is compiled and exercised in an image without NGSolve.
#pragma once
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../oracle/ref_pin/ngs_standin.hpp"

namespace ngcore {
// ngcore::Flags: typed (name, value) lists with positional access (flags.hpp)
class Flags {
  std::vector<std::pair<std::string, double>> num;
  std::vector<std::pair<std::string, std::string>> str;
  std::vector<std::pair<std::string, bool>> def;
public:
  Flags &SetFlag(const std::string &n, double v) { num.emplace_back(n, v); return *this; }
  Flags &SetFlag(const std::string &n, const std::string &v) { str.emplace_back(n, v); return *this; }
  Flags &SetFlag(const std::string &n, const char *v) { str.emplace_back(n, std::string(v)); return *this; }
  Flags &SetFlag(const std::string &n, bool v = true) { def.emplace_back(n, v); return *this; }
  int GetNNumFlags() const { return int(num.size()); }
  int GetNStringFlags() const { return int(str.size()); }
  int GetNDefineFlags() const { return int(def.size()); }
  double GetNumFlag(int i, std::string &name) const { name = num[i].first; return num[i].second; }
  const std::string &GetStringFlag(int i, std::string &name) const { name = str[i].first; return str[i].second; }
  bool GetDefineFlag(int i, std::string &name) const { name = def[i].first; return def[i].second; }
  double GetNumFlag(const std::string &n, double dflt) const { for (auto &e : num) if (e.first == n) return e.second; return dflt; }
};
}  // namespace ngcore

namespace ngla {
template <class T> class VVector : public BaseVector {
public:
  explicit VVector(size_t n) : BaseVector(n, sizeof(T) / sizeof(double)) {}
};
using AutoVector = std::unique_ptr<BaseVector>;
// the rest of the BaseMatrix interface the adapter overrides (basematrix.hpp)
class BaseMatrixFull : public BaseMatrix {
public:
  virtual void MultTrans(const BaseVector &x, BaseVector &y) const = 0;
  virtual void MultTransAdd(double s, const BaseVector &x, BaseVector &y) const = 0;
  virtual AutoVector CreateRowVector() const = 0;
  virtual AutoVector CreateColVector() const = 0;
  virtual bool IsComplex() const { return false; }
};
}  // namespace ngla
#define NGSAMG_B200_BASEMATRIX ngla::BaseMatrixFull
#define NGSAMG_B200_NGSOLVE_STANDIN 1

namespace ngcomp {
using ngcore::Flags;
class BilinearForm {};
// ngcomp::Preconditioner (preconditioner.hpp): a BaseMatrix that NGSolve drives through InitLevel / FinalizeLevel / Update
class Preconditioner : public ngla::BaseMatrixFull {
protected:
  std::shared_ptr<BilinearForm> bfa;
  Flags flags;
  std::string name;
public:
  Preconditioner(std::shared_ptr<BilinearForm> abfa, const Flags &f, const std::string aname) : bfa(abfa), flags(f), name(aname) {}
  virtual void InitLevel(std::shared_ptr<ngcore::BitArray> freedofs) = 0;
  virtual void FinalizeLevel(const ngla::BaseMatrix *mat) = 0;
  virtual void Update() = 0;
  virtual const ngla::BaseMatrix &GetMatrix() const = 0;
  virtual const ngla::BaseMatrix &GetAMatrix() const = 0;
  ngla::AutoVector CreateRowVector() const override { return std::make_unique<ngla::VVector<double>>(size_t(VHeight())); }
  ngla::AutoVector CreateColVector() const override { return std::make_unique<ngla::VVector<double>>(size_t(VHeight())); }
};
// GetPreconditionerClasses() / RegisterPreconditioner<T> (preconditioner.hpp): name -> creator
using PCCreator = std::function<std::shared_ptr<Preconditioner>(std::shared_ptr<BilinearForm>, const Flags &, const std::string)>;
inline std::map<std::string, PCCreator> &GetPreconditionerClasses()
{
  static std::map<std::string, PCCreator> reg;
  return reg;
}
template <class PRECOND> struct RegisterPreconditioner {
  explicit RegisterPreconditioner(const std::string &label)
  {
    GetPreconditionerClasses()[label] = [](std::shared_ptr<BilinearForm> b, const Flags &f, const std::string n) {
      return std::static_pointer_cast<Preconditioner>(std::make_shared<PRECOND>(b, f, n));
    };
  }
};
}  // namespace ngcomp
