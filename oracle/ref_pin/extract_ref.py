#!/usr/bin/env python3
"""Cut single function definitions out of the reference's sources AT BUILD TIME.

TEST INFRASTRUCTURE ONLY (oracle/).  The reference (an NGSolve add-on) cannot be built in this image, and no translation
unit of the hot path compiles on its own (every one includes NGSolve's comp.hpp).  What CAN be compiled is the text of
the individual functions of the path, against a small stand-in for the NGSolve containers they use
(oracle/ref_pin/ngs_standin.hpp, written for this repository).  This script locates each function in the tree under
/root/reference by an anchor pattern, cuts it out with a brace matcher that understands comments and literals, and
writes it to oracle/_ref/frag/<name>.inc.  Nothing of the reference is stored in this repository: oracle/_ref/ is
git-ignored, and oracle/Makefile deletes the fragments again once the library is compiled.

usage: extract_ref.py <reference root> <output dir>
"""
import os
import re
import sys

# name, file (relative to the reference root), anchor regex (first match wins, searched from `after` if given),
# how the fragment starts: "template" = walk back to the closest preceding line that starts a template header,
# "line" = the anchor line itself
FRAGMENTS = [
    # --- sparse matrix products (RAP) -------------------------------------------------------------------------
    ("timer_transpose", "src/base/linalg/utils_sparseMM.cpp", r"timer_hack_TransposeSPMImpl \(\)", "line", None),
    ("transpose", "src/base/linalg/utils_sparseMM.cpp", r"^\s*TransposeSPMImpl \(SparseMatTM<H, W> const &mat\)", "template", None),
    ("timer_matmult", "src/base/linalg/utils_sparseMM.cpp", r"timer_hack_MatMultABImpl \(int nr\)", "line", None),
    ("matmult", "src/base/linalg/utils_sparseMM.cpp", r"^MatMultABImpl \(SparseMatTM<A, B> const &mata,", "template", None),
    ("timer_restrict", "src/base/linalg/utils_sparseMM.hpp", r"timer_hack_restrictspm2 \(\)", "line", None),
    ("restrict", "src/base/linalg/utils_sparseMM.hpp", r"^RestrictMatrix \(SparseMatTM<W, H> const &PT,", "template", None),
    # --- sequential Gauss-Seidel ------------------------------------------------------------------------------
    ("gss3_setup", "src/base/smoothers/gssmoother.cpp", r"^void GSS3<TM> :: SetUp \(", "template", None),
    ("gss3_calcdiags", "src/base/smoothers/gssmoother.cpp", r"^void GSS3<TM> :: CalcDiags \(", "template", None),
    ("gss3_rhs", "src/base/smoothers/gssmoother.cpp", r"^void GSS3<TM> :: SmoothRHSInternal \(", "template", None),
    ("gss3_res", "src/base/smoothers/gssmoother.cpp", r"^void GSS3<TM> :: SmoothRESInternal \(", "template", None),
    ("gss3_smooth", "src/base/smoothers/gssmoother.cpp", r"^void GSS3<TM> :: Smooth \(BaseVector", "template", None),
    ("gss3_smoothback", "src/base/smoothers/gssmoother.cpp", r"^void GSS3<TM> :: SmoothBack \(BaseVector", "template", None),
    # --- pseudo inverse of a diagonal block ----------------------------------------------------------------------
    ("la_reltol", "src/base/utils/utils_denseLA.hpp", r"^constexpr T RelZeroTol\(\)", "template", None),
    ("la_abstol", "src/base/utils/utils_denseLA.hpp", r"^constexpr T AbsZeroTol\(\)", "template", None),
    ("la_regtm", "src/base/utils/utils_denseLA.hpp", r"^template<int IMIN, int N, int NN> INLINE void RegTM \(Mat<NN,NN,double> & m, double maxadd = -1\)", "line", None),
    ("la_nzblock", "src/base/utils/utils_denseLA.hpp", r"^CallOnNonZeroDiagonalBlock \(int const &n, TGETETR mat,", "template", None),
    ("la_nzblock_mat", "src/base/utils/utils_denseLA.hpp", r"^CallOnNonZeroDiagonalBlock \(Mat<N, N, TSCAL> &mat,", "template", None),
    ("la_trydirect_simple", "src/base/utils/utils_denseLA.cpp", r"^TryDirectInverse_simple \(FlatMatrix<TSCAL> A, LocalHeap & lh\)", "template", None),
    ("la_trydirect", "src/base/utils/utils_denseLA.hpp", r"^TryDirectInverse \(FlatMatrix<TSCAL> A, LocalHeap & lh\)", "template", None),
    ("la_pinv_tol", "src/base/utils/utils_denseLA.hpp", r"^CalcPseudoInverseWithTolNonZeroBlock \(FlatMatrix<TSCAL>  M,", "template", None),
    ("la_pinv_mat", "src/base/utils/utils_denseLA.hpp", r"^CalcPseudoInverseTryNormal \(Mat<N, N, TSCAL>       &mat,", "template", None),
    ("la_pinv_scal", "src/base/utils/utils_denseLA.hpp", r"^CalcPseudoInverseTryNormal \(TSCAL &mat, LocalHeap &lh,", "template", None),
    # --- smoother protocol (in-class definitions of BaseSmoother / ProxySmoother) ------------------------------
    ("bs_smoothsymm", "src/base/smoothers/base_smoother.hpp", r"^\s*virtual void SmoothSymm \(BaseVector", "line", None),
    ("bs_smoothk", "src/base/smoothers/base_smoother.hpp", r"^\s*virtual void SmoothK \(int k", "line", None),
    ("bs_smoothbackk", "src/base/smoothers/base_smoother.hpp", r"^\s*virtual void SmoothBackK \(int k", "line", None),
    ("bs_smoothsymmk", "src/base/smoothers/base_smoother.hpp", r"^\s*virtual void SmoothSymmK \(int k", "line", None),
    ("bs_calcresiduum", "src/base/smoothers/base_smoother.hpp", r"^\s*virtual void CalcResiduum\(BaseVector const &x,", "line", None),
    ("proxy_smooth", "src/base/smoothers/base_smoother.hpp", r"^\s*virtual void Smooth \(BaseVector &x, const BaseVector &b,", "line",
     r"^class ProxySmoother"),
    ("proxy_smoothback", "src/base/smoothers/base_smoother.hpp", r"^\s*virtual void SmoothBack \(BaseVector &x, const BaseVector &b,", "line",
     r"^class ProxySmoother"),
    ("rich_ctor", "src/base/smoothers/base_smoother.cpp", r"^RichardsonSmoother :: RichardsonSmoother \(", "line", None),
    ("rich_smooth", "src/base/smoothers/base_smoother.cpp", r"^void RichardsonSmoother :: Smooth \(", "line", None),
    ("rich_smoothback", "src/base/smoothers/base_smoother.cpp", r"^void RichardsonSmoother :: SmoothBack \(", "line", None),
    ("jacobi_ctor", "src/base/smoothers/base_smoother.cpp", r"^JacobiSmoother<TM>::JacobiSmoother \(", "template", None),
    # --- rigid-body transport between vertices (elasticity prolongation blocks) ---------------------------------------
    ("el_calcq", "src/elasticity/elasticity_energy_impl.hpp", r"^INLINE void EpsEpsEnergy<DIM, TVD, TED>::CalcQ\(const Vec<DIM>& t, TM& Q,", "template", None),
    # --- grid transfer -----------------------------------------------------------------------------------------
    ("prol_f2c", "src/base/coarsening/dof_map.cpp", r"^TransferF2C \(BaseVector const \*x_fine,", "template", r"^timer_hack_prol_c2f"),
    ("prol_addc2f", "src/base/coarsening/dof_map.cpp", r"^AddC2F \(double fac, BaseVector \*x_fine, BaseVector const \*x_coarse\) const", "template",
     r"^timer_hack_prol_c2f"),
    # --- multigrid cycles --------------------------------------------------------------------------------------
    ("amg_smoothw", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: SmoothW \(", "line", None),
    ("amg_smoothbs", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: SmoothBS \(", "line", None),
    ("amg_smoothv", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: SmoothV \(", "line", None),
    ("amg_smoothvfrom", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: SmoothVFromLevel \(", "line", None),
    ("amg_getoc", "src/base/solve/amg_matrix.cpp", r"^\s*Array<double> AMGMatrix :: GetOC \(\) const", "line", None),
    ("bs_getnops", "src/base/smoothers/base_smoother.hpp", r"^\s*virtual size_t GetNOps \(\) const$", "line", None),
    ("bs_getanze", "src/base/smoothers/base_smoother.hpp", r"^\s*virtual size_t GetANZE \(\) const$", "line", None),
    ("proxy_getnops", "src/base/smoothers/base_smoother.cpp", r"^size_t ProxySmoother :: GetNOps \(\) const", "line", None),
    ("proxy_getanze", "src/base/smoothers/base_smoother.cpp", r"^size_t ProxySmoother :: GetANZE \(\) const", "line", None),
    ("amg_smooth", "src/base/solve/amg_matrix.hpp", r"^\s*INLINE void Smooth \(BaseVector & x, const BaseVector & b\) const", "line", None),
    ("amg_mult", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: Mult \(const BaseVector & b, BaseVector & x\) const", "line", None),
    ("amg_multtrans", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: MultTrans \(const BaseVector & b, BaseVector & x\) const", "line", None),
    ("amg_multadd", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: MultAdd \(double s, const BaseVector & b, BaseVector & x\) const", "line", None),
    ("amg_multtransadd", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: MultTransAdd \(double s, const BaseVector & b, BaseVector & x\) const", "line", None),
    # ===== multi-rank path (compiled against the threaded MPI stand-in, ngs_standin_mpi.hpp) =======================
    # --- small utilities of the reference's own tree ---------------------------------------------------------------
    ("u_find_sorted", "src/base/utils/utils_arrays_tables.hpp", r"^INLINE size_t find_in_sorted_array \(const T & elem, FlatArray<T> a\)$", "template", None),
    ("u_merge3", "src/base/utils/utils_arrays_tables.hpp", r"^INLINE void merge_arrays \(T1& tab_in, Array<T2> & out, T3 lam_comp\)", "template", None),
    ("u_tabtrait", "src/base/utils/utils_arrays_tables.hpp", r"^template<class T> struct tab_scal_trait", "line", None),
    ("u_merge2", "src/base/utils/utils_arrays_tables.hpp", r"^INLINE Array<typename tab_scal_trait<T1>::type> merge_arrays \(T1& tab_in, T3 lam_comp\)", "template", None),
    ("u_allreduce_dofdata", "src/base/distributed/mpiwrap_extension.hpp", r"^void MyAllReduceDofData \(const ParallelDofs", "template", None),
    # --- DCCMap: DISTRIBUTED -> CONCENTRATED -> CUMULATED ----------------------------------------------------------
    ("dcc_alloc", "src/base/linalg/dcc_map.cpp", r"^AllocMPIStuff \(\)", "template", None),
    ("dcc_start_d2c", "src/base/linalg/dcc_map.cpp", r"^StartDIS2CO \(BaseVector & vec\) const", "template", None),
    ("dcc_apply_d2c", "src/base/linalg/dcc_map.cpp", r"^ApplyDIS2CO \(BaseVector & vec\) const", "template", None),
    ("dcc_finish_d2c", "src/base/linalg/dcc_map.cpp", r"^FinishDIS2CO \(\) const", "template", None),
    ("dcc_start_c2c", "src/base/linalg/dcc_map.cpp", r"^StartCO2CU \(BaseVector & vec\) const", "template", None),
    ("dcc_apply_c2c", "src/base/linalg/dcc_map.cpp", r"^ApplyCO2CU \(BaseVector & vec\) const", "template", None),
    ("dcc_finish_c2c", "src/base/linalg/dcc_map.cpp", r"^FinishCO2CU \(\) const", "template", None),
    ("dcc_wait_d2c", "src/base/linalg/dcc_map.cpp", r"^WaitD2C \(\) const", "template", None),
    ("dcc_iterate", "src/base/linalg/dcc_map.cpp", r"^iterate_buf_vec \(int block_size", "template", None),
    ("dcc_buffer_g", "src/base/linalg/dcc_map.cpp", r"^BufferG \(BaseVector & vec\) const", "template", None),
    ("dcc_apply_m", "src/base/linalg/dcc_map.cpp", r"^ApplyM \(BaseVector & vec\) const", "template", None),
    ("dcc_buffer_m", "src/base/linalg/dcc_map.cpp", r"^BufferM \(BaseVector & vec\) const", "template", None),
    ("dcc_apply_g", "src/base/linalg/dcc_map.cpp", r"^ApplyG \(BaseVector & vec\) const", "template", None),
    ("dcc_masters", "src/base/linalg/dcc_map.cpp", r"^CalcDOFMasters \(\)$", "template", r"^BasicDCCMap<TSCAL>::"),
    # --- CtrMap: contraction of a distributed level onto the group master ---------------------------------------------
    ("ctr_timer_f2c", "src/base/coarsening/dof_contract.cpp", r"timer_hack_ctr_f2c \(\)", "line", None),
    ("ctr_timer_c2f", "src/base/coarsening/dof_contract.cpp", r"timer_hack_ctr_c2f \(\)", "line", None),
    ("ctr_f2c", "src/base/coarsening/dof_contract.cpp", r"^void CtrMap<TV> :: TransferF2C \(", "template", None),
    ("ctr_addf2c", "src/base/coarsening/dof_contract.cpp", r"^void CtrMap<TV> :: AddF2C \(", "template", None),
    ("ctr_c2f", "src/base/coarsening/dof_contract.cpp", r"^void CtrMap<TV> :: TransferC2F \(", "template", None),
    ("ctr_addc2f", "src/base/coarsening/dof_contract.cpp", r"^void CtrMap<TV> :: AddC2F \(", "template", None),
    ("ctr_setup_mpi", "src/base/coarsening/dof_contract.cpp", r"^void CtrMap<TV> :: SetUpMPIStuff \(\)", "template", None),
    ("ctr_timer_mat", "src/base/coarsening/dof_contract.cpp", r"timer_hack_ctrmat \(int nr\)", "line", None),
    ("ctr_assemble", "src/base/coarsening/dof_contract.cpp", r"^CtrMap<TV> :: DoAssembleMatrix \(", "template", None),
    # --- hybrid matrix: A = M + G ----------------------------------------------------------------------------------
    ("hyb_decompose", "src/base/linalg/hybrid_matrix.cpp", r"^DecomposeSparseMatrixHybrid \(shared_ptr<SparseMatrix<TM>> anA,", "template", None),
    ("hyb_multadd", "src/base/linalg/hybrid_matrix.cpp", r"^MultAdd \(double s, const BaseVector & x, BaseVector & y\) const", "template", None),
    ("hyb_mult", "src/base/linalg/hybrid_matrix.cpp", r"^Mult \(const BaseVector & x, BaseVector & y\) const", "template", None),
    # --- modified diagonal ------------------------------------------------------------------------------------------
    ("rdg_generic", "src/base/smoothers/hybrid_smoother_utils.hpp", r"^CalcHybridSmootherRDGItGeneric\(size_t", "template", None),
    ("rdg", "src/base/smoothers/hybrid_smoother_utils.hpp", r"^CalcHybridSmootherRDG\(size_t", "template", None),
    ("hyb_calcmoddiag", "src/base/smoothers/hybrid_smoother.cpp", r"^CalcModDiag \(shared_ptr<BitArray> free\)", "template", None),
    # --- GSS3 range sweeps (in-class), GSS4, HybridGSSmoother --------------------------------------------------------
    ("gss3_r_smooth", "src/base/smoothers/gssmoother.hpp", r"^\s*virtual void Smooth \(size_t first, size_t next, BaseVector &x, const BaseVector &b\) const", "line", None),
    ("gss3_r_smoothback", "src/base/smoothers/gssmoother.hpp", r"^\s*virtual void SmoothBack \(size_t first, size_t next, BaseVector &x, const BaseVector &b\) const", "line", None),
    ("gss3_r_smoothres", "src/base/smoothers/gssmoother.hpp", r"^\s*virtual void SmoothRES \(size_t first, size_t next, BaseVector &x, BaseVector &res\) const", "line", None),
    ("gss3_r_smoothbackres", "src/base/smoothers/gssmoother.hpp", r"^\s*virtual void SmoothBackRES \(size_t first, size_t next, BaseVector &x, BaseVector &res\) const", "line", None),
    ("gss4_smooth", "src/base/smoothers/gssmoother.hpp", r"^\s*INLINE void Smooth \(BaseVector &x, const BaseVector &b\) const", "line", None),
    ("gss4_smoothback", "src/base/smoothers/gssmoother.hpp", r"^\s*INLINE void SmoothBack \(BaseVector &x, const BaseVector &b\) const", "line", None),
    ("gss4_smoothres", "src/base/smoothers/gssmoother.hpp", r"^\s*INLINE void SmoothRES \(BaseVector &x, BaseVector &res\) const", "line", None),
    ("gss4_smoothbackres", "src/base/smoothers/gssmoother.hpp", r"^\s*INLINE void SmoothBackRES \(BaseVector &x, BaseVector &res\) const", "line", None),
    ("gss4_ctor_repl", "src/base/smoothers/gssmoother.cpp", r"^GSS4<TM> :: GSS4 \(shared_ptr<SparseMatrix<TM>> A, FlatArray<TM> repl_diag", "template", None),
    ("gss4_iterate", "src/base/smoothers/gssmoother.cpp", r"^INLINE void GSS4<TM> :: iterate_rows", "template", None),
    ("gss4_setup", "src/base/smoothers/gssmoother.cpp", r"^void GSS4<TM> :: SetUp \(", "template", None),
    ("gss4_res", "src/base/smoothers/gssmoother.cpp", r"^void GSS4<TM> :: SmoothRESInternal \(", "template", None),
    ("gss4_rhs", "src/base/smoothers/gssmoother.cpp", r"^void GSS4<TM> :: SmoothRHSInternal \(", "template", None),
    ("hgs_finalize", "src/base/smoothers/gssmoother.cpp", r"^Finalize \(\)", "template", r"^/\*\* HybridGSSmoother \*\*/"),
    ("hgs_stage_rhs", "src/base/smoothers/gssmoother.cpp", r"^SmoothStageRHS \(SMOOTH_STAGE        const &stage,", "template", r"^/\*\* HybridGSSmoother \*\*/"),
    ("hgs_stage_res", "src/base/smoothers/gssmoother.cpp", r"^SmoothStageRes \(SMOOTH_STAGE        const &stage,", "template", r"^/\*\* HybridGSSmoother \*\*/"),
    # --- HybridBaseSmoother: the stage protocol around the exchanges -----------------------------------------------
    ("hbs_start_d2c", "src/base/smoothers/hybrid_base_smoother.cpp", r"^StartDIS2CO \(BaseVector &vec\) const", "template", None),
    ("hbs_finish_d2c", "src/base/smoothers/hybrid_base_smoother.cpp", r"^FinishDIS2CO \(BaseVector &vec\) const", "template", None),
    ("hbs_start_c2c", "src/base/smoothers/hybrid_base_smoother.cpp", r"^StartCO2CU \(BaseVector &vec\) const", "template", None),
    ("hbs_finish_c2c", "src/base/smoothers/hybrid_base_smoother.cpp", r"^FinishCO2CU \(BaseVector &vec\) const", "template", None),
    ("hbs_smooth", "src/base/smoothers/hybrid_base_smoother.cpp", r"^Smooth \(BaseVector       &x,", "template", None),
    ("hbs_smoothback", "src/base/smoothers/hybrid_base_smoother.cpp", r"^SmoothBack \(BaseVector       &x,", "template", None),
    ("hbs_impl", "src/base/smoothers/hybrid_base_smoother.cpp", r"^SmoothImpl \(BaseVector       &x,", "template", None),
    ("hbs_impl_res", "src/base/smoothers/hybrid_base_smoother.cpp", r"^SmoothImplRES \(BaseVector       &x,", "template", None),
    ("hbs_impl_rhs", "src/base/smoothers/hybrid_base_smoother.cpp", r"^SmoothImplRHS \(BaseVector       &x,", "template", None),
    ("hbs_stages", "src/base/smoothers/hybrid_base_smoother.cpp", r"^CallStageKernelsImpl\(BaseVector       &x,", "template", None),
]


def match_braces(text, pos):
    """index one past the brace that closes the first '{' at or after pos; comments, string and char literals are skipped"""
    depth, i, n, seen = 0, pos, len(text), False
    while i < n:
        c = text[i]
        two = text[i:i + 2]
        if two == "//":
            i = text.index("\n", i) if "\n" in text[i:] else n
            continue
        if two == "/*":
            i = text.index("*/", i) + 2
            continue
        if c == '"' or c == "'":
            q, i = c, i + 1
            while text[i] != q:
                i += 2 if text[i] == "\\" else 1
            i += 1
            continue
        if c == "{":
            depth, seen = depth + 1, True
        elif c == "}":
            depth -= 1
            if seen and depth == 0:
                return i + 1
        elif c == ";" and not seen:
            raise ValueError("declaration, not a definition")
        i += 1
    raise ValueError("unbalanced braces")


def extract(root, name, rel, anchor, start, after):
    path = os.path.join(root, rel)
    lines = open(path).read().split("\n")
    first = 0
    if after:
        first = next(i for i, ln in enumerate(lines) if re.search(after, ln))
    a = next(i for i in range(first, len(lines)) if re.search(anchor, lines[i]))
    s = a
    if start == "template":
        while not re.match(r"\s*template\s*<", lines[s]):
            s -= 1
            if a - s > 6:
                raise ValueError("%s: no template header above %s:%d" % (name, rel, a + 1))
        while s > 0 and re.match(r"\s*template\s*<", lines[s - 1]):     # member templates of class templates: two headers
            s -= 1
    text = "\n".join(lines)
    off = sum(len(ln) + 1 for ln in lines[:s])
    aoff = sum(len(ln) + 1 for ln in lines[:a])
    end = match_braces(text, aoff)
    body = text[off:end]
    l1 = s + body.count("\n") + 1
    return "// %s:%d-%d -- cut from the reference at build time by oracle/ref_pin/extract_ref.py, not stored in git\n%s\n" % (
        rel, s + 1, l1, body), (rel, s + 1, l1)


# second, independent group (own library _ref/libngsamg_ref_bgs.so, own stand-in ngs_standin_bgs.hpp): the two update routines of the
# block Gauss-Seidel smoother
FRAGMENTS_BGS = [
    ("bgs_richardson", "src/base/smoothers/loc_block_gssmoother_impl.hpp",
     r"^\s*INLINE void BSmoother2<TM>::BSBlock :: RichardsonUpdate \(double omega, FlatVector<TV> smallsol, FlatVector<TV> bigsol,", "template", None),
    ("bgs_richardson_res", "src/base/smoothers/loc_block_gssmoother_impl.hpp",
     r"^\s*INLINE void BSmoother2<TM>::BSBlock :: RichardsonUpdate_RES \(double omega, FlatVector<TV> smallupdate, FlatVector<TV> bigsol,", "template", None),
    # block set-up: off-block rows, dense diagonal block and its inverse
    ("bgs_setfromspmat", "src/base/smoothers/loc_block_gssmoother_impl.hpp",
     r"^\s*INLINE void BSmoother2<TM>::BSBlock :: SetFromSPMat \(const SparseMatrixTM<TM> & A, FlatArray<int> dofs, LocalHeap & lh, bool pinv,", "template", None),
    # block order and the smoother's flag protocol
    ("bgs_iterate", "src/base/smoothers/loc_block_gssmoother_impl.hpp", r"^\s*INLINE void BSmoother2<TM> :: IterateBlocks \(FlatArray<int> groups, bool reverse, TLAM smooth_block\) const", "template", None),
    ("bgs_smoothwo", "src/base/smoothers/loc_block_gssmoother_impl.hpp", r"^\s*INLINE void BSmoother2<TM> :: SmoothWO \(FlatArray<int> groups, BaseVector & x, const BaseVector & b,", "template", None),
    ("bgs_smoothsimple", "src/base/smoothers/loc_block_gssmoother_impl.hpp", r"^\s*INLINE void BSmoother2<TM> :: SmoothSimple \(FlatArray<int> groups, BaseVector & x, const BaseVector & b, int steps, bool reverse, bool symm\) const", "template", None),
    ("bgs_smoothressimple", "src/base/smoothers/loc_block_gssmoother_impl.hpp", r"^\s*INLINE void BSmoother2<TM> :: SmoothRESSimple \(FlatArray<int> groups, BaseVector & x, BaseVector & res, int steps, bool reverse, bool symm\) const", "template", None),
]


def main():
    root, out = sys.argv[1], sys.argv[2]
    group = sys.argv[3] if len(sys.argv) > 3 else "main"
    os.makedirs(out, exist_ok=True)
    index = []
    for name, rel, anchor, start, after in (FRAGMENTS_BGS if group == "bgs" else FRAGMENTS):
        frag, where = extract(root, name, rel, anchor, start, after)
        with open(os.path.join(out, name + ".inc"), "w") as f:
            f.write(frag)
        index.append("%-18s %s:%d-%d" % ((name,) + where))
    with open(os.path.join(out, "INDEX.txt"), "w") as f:
        f.write("\n".join(index) + "\n")
    with open(os.path.join(out, "INDEX.inc"), "w") as f:       # the same list as a C string literal (ref_fragment_index())
        f.write("\n".join('"%s\\n"' % ln for ln in index) + "\n")
    print("\n".join(index))


if __name__ == "__main__":
    main()
