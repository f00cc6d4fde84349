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
    # --- grid transfer -----------------------------------------------------------------------------------------
    ("prol_f2c", "src/base/coarsening/dof_map.cpp", r"^TransferF2C \(BaseVector const \*x_fine,", "template", r"^timer_hack_prol_c2f"),
    ("prol_addc2f", "src/base/coarsening/dof_map.cpp", r"^AddC2F \(double fac, BaseVector \*x_fine, BaseVector const \*x_coarse\) const", "template",
     r"^timer_hack_prol_c2f"),
    # --- multigrid cycles --------------------------------------------------------------------------------------
    ("amg_smoothw", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: SmoothW \(", "line", None),
    ("amg_smoothbs", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: SmoothBS \(", "line", None),
    ("amg_smoothv", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: SmoothV \(", "line", None),
    ("amg_smoothvfrom", "src/base/solve/amg_matrix.cpp", r"^\s*void AMGMatrix :: SmoothVFromLevel \(", "line", None),
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
    text = "\n".join(lines)
    off = sum(len(ln) + 1 for ln in lines[:s])
    aoff = sum(len(ln) + 1 for ln in lines[:a])
    end = match_braces(text, aoff)
    body = text[off:end]
    l1 = s + body.count("\n") + 1
    return "// %s:%d-%d -- cut from the reference at build time by oracle/ref_pin/extract_ref.py, not stored in git\n%s\n" % (
        rel, s + 1, l1, body), (rel, s + 1, l1)


def main():
    root, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    index = []
    for name, rel, anchor, start, after in FRAGMENTS:
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
