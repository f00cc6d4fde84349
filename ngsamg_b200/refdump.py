"""Reader / writer for the reference's on-disk debug dumps -- the interchange format between a real NgsAMG run and this library.

With `ngs_amg_log_level = "debug"` the reference writes its hierarchy as text files:

    ngs_amg_mat_l<level>_rk<rank>.out     level matrices          print_bmat, src/base/factory/base_factory.cpp:437-444, 506-522
    SP_semi_aux_rk_<rank>_l_<level>.out   smoothed prolongations  src/base/factory/vertex_factory_impl.hpp:2422-2428

both through `print_tm_spmat` (src/base/utils/utils_io.hpp:102-132):

  * block matrices (Mat<H,W>), H text lines per block row:
        "Row <k, width 6>: " { "<col, width 4>: " { "<v, width 4> " } * W  " | " } ...
        "          : "        { "    : "          { "<v> " } * W            " | " } ...     (block rows 1 .. H-1)
        "Row <k>: (empty)"
  * scalar matrices: NGSolve's own `operator<<` of SparseMatrix<double>:  "Row <k>:   <col>: <v>   <col>: <v> ..."

Values are printed with the stream's default precision (6 significant digits), so a dump pins the INTEGER side of the hierarchy
(sparsity patterns, DOF maps -- what BASELINE.json wants bit-exact) and the values to ~1e-6.  `read_*` turns such files into
`ngsamg_b200.SparseMatrix` objects that can be injected (`h1_scal(A, free, prolongations=[...])`) or compared
(`compare_patterns`); `write_*` / `dump_hierarchy` emit OUR hierarchy in the same format and file naming, so a plain `diff` against the
files of a reference run works.  Nothing here needs a device.
"""
import os
import re

import numpy as np

from . import SparseMatrix

_ROW = re.compile(r"^\s*Row\s+(\d+)\s*:(.*)$")
_CONT = re.compile(r"^\s*:(.*)$")


def parse_spmat(text, ncols=None, nrows=None):
    """text of one print_tm_spmat dump -> SparseMatrix (block shape inferred).  ncols / nrows: sizes if known (a dump does not state
    them: trailing empty columns are invisible)."""
    rows = {}          # row -> list of (col, [block row 0 values], [block row 1 values], ...)
    cur = None
    for line in text.splitlines():
        if not line.strip():
            continue
        m = _ROW.match(line)
        if m:
            k, rest = int(m.group(1)), m.group(2)
            cur = k
            rows[k] = []
            if "(empty)" in rest:
                continue
            if "|" in rest:        # block format: entries separated by '|', "<col>: v v v"
                for chunk in rest.split("|"):
                    chunk = chunk.strip()
                    if not chunk:
                        continue
                    c, vals = chunk.split(":", 1)
                    rows[k].append([int(c), [float(v) for v in vals.split()]])
            else:                  # scalar format: "<col>: <v>" pairs
                toks = rest.replace(":", " : ").split()
                if len(toks) % 3:
                    raise ValueError("cannot parse row %d of the dump: %r" % (k, line))
                for i in range(0, len(toks), 3):
                    c, colon, v = toks[i], toks[i + 1], toks[i + 2]
                    if colon != ":":
                        raise ValueError("cannot parse row %d of the dump: %r" % (k, line))
                    rows[k].append([int(c), [float(v)]])
            continue
        m = _CONT.match(line)
        if m and cur is not None and "|" in line:   # further block rows of the current row
            chunks = [c.strip() for c in m.group(1).split("|")]
            chunks = [c for c in chunks if c]
            if len(chunks) != len(rows[cur]):
                raise ValueError("row %d: continuation line has %d blocks, the first line %d" % (cur, len(chunks), len(rows[cur])))
            for ent, chunk in zip(rows[cur], chunks):
                vals = chunk.split(":", 1)[1] if ":" in chunk else chunk
                ent.append([float(v) for v in vals.split()])
            continue
        # anything else (headers such as " prol : ") is ignored
    n = (max(rows) + 1) if rows else 0
    if nrows is not None:
        if nrows < n:
            raise ValueError("dump has %d rows, expected %d" % (n, nrows))
        n = nrows
    bh = bw = 1
    for ents in rows.values():
        if ents:
            bh, bw = len(ents[0]) - 1, len(ents[0][1])
            break
    rowptr = np.zeros(n + 1, np.int64)
    for k, ents in rows.items():
        rowptr[k + 1] = len(ents)
    np.cumsum(rowptr, out=rowptr)
    nnz = int(rowptr[-1])
    col = np.zeros(nnz, np.int32)
    val = np.zeros((nnz, bh, bw))
    for k, ents in rows.items():
        p = int(rowptr[k])
        for ent in ents:
            if len(ent) - 1 != bh or any(len(r) != bw for r in ent[1:]):
                raise ValueError("row %d: inconsistent block shape in the dump" % k)
            col[p] = ent[0]
            val[p] = np.asarray(ent[1:])
            p += 1
    mc = int(col.max()) + 1 if nnz else 0
    if ncols is None:
        ncols = mc
    elif ncols < mc:
        raise ValueError("dump references column %d, expected at most %d columns" % (mc - 1, ncols))
    return SparseMatrix(n, ncols, bh, bw, rowptr, col, val.reshape(-1))


def read_spmat(path, ncols=None, nrows=None):
    with open(path) as f:
        return parse_spmat(f.read(), ncols=ncols, nrows=nrows)


def format_spmat(M, precision=6):
    """SparseMatrix -> text in the reference's print_tm_spmat format.  precision=6 reproduces the reference's default stream
    precision; precision=17 is lossless."""
    fmt = "%%.%dg" % precision
    out = []
    val = M.val.reshape(-1, M.bh, M.bw) if M.nnz else np.zeros((0, M.bh, M.bw))
    scalar = M.bh == 1 and M.bw == 1
    for k in range(M.nrows):
        lo, hi = int(M.rowptr[k]), int(M.rowptr[k + 1])
        if scalar:
            out.append("Row %d:" % k + "".join("   %d: %s" % (M.col[j], fmt % val[j, 0, 0]) for j in range(lo, hi)))
            continue
        if lo == hi:
            out.append("Row %6d: (empty)" % k)
            continue
        for kh in range(M.bh):
            line = ("Row %6d: " % k) if kh == 0 else "          : "
            for j in range(lo, hi):
                line += ("%4d: " % M.col[j]) if kh == 0 else "    : "
                line += "".join("%4s " % (fmt % val[j, kh, jw]) for jw in range(M.bw))
                line += " | "
            out.append(line)
    return "\n".join(out) + "\n"


def write_spmat(path, M, precision=6):
    with open(path, "w") as f:
        f.write(format_spmat(M, precision))


def mat_file(level, rank=0):
    return "ngs_amg_mat_l%d_rk%d.out" % (level, rank)


def prol_file(level, rank=0):
    return "SP_semi_aux_rk_%d_l_%d.out" % (rank, level)


def load_hierarchy(directory, rank=0):
    """read every level matrix / prolongation dump of `rank` found in `directory` -> (mats, prols): mats[l] = A_l, prols[l] = P_l
    (level l+1 -> l; column count taken from the next level matrix when that dump exists)."""
    mats, prols = [], []
    l = 0
    while os.path.exists(os.path.join(directory, mat_file(l, rank))):
        mats.append(read_spmat(os.path.join(directory, mat_file(l, rank))))
        l += 1
    l = 0
    while os.path.exists(os.path.join(directory, prol_file(l, rank))):
        nc = mats[l + 1].nrows if l + 1 < len(mats) else None
        nr = mats[l].nrows if l < len(mats) else None
        prols.append(read_spmat(os.path.join(directory, prol_file(l, rank)), ncols=nc, nrows=nr))
        l += 1
    return mats, prols


def dump_hierarchy(pc, directory, rank=0, precision=6):
    """write the hierarchy of a finalized preconditioner in the reference's file naming and format"""
    os.makedirs(directory, exist_ok=True)
    n = pc.GetNLevels()
    for l in range(n):
        write_spmat(os.path.join(directory, mat_file(l, rank)), pc.GetLevelMatrix(l), precision)
    for l in range(n - 1):
        write_spmat(os.path.join(directory, prol_file(l, rank)), pc.GetProlongation(l), precision)


def compare_patterns(M1, M2):
    """None if the two matrices have the same shape and sparsity pattern, else a short description of the first difference"""
    if (M1.nrows, M1.bh, M1.bw) != (M2.nrows, M2.bh, M2.bw):
        return "shapes differ: %dx(%dx%d) vs %dx(%dx%d)" % (M1.nrows, M1.bh, M1.bw, M2.nrows, M2.bh, M2.bw)
    if not np.array_equal(M1.rowptr, M2.rowptr):
        k = int(np.flatnonzero(np.diff(M1.rowptr) != np.diff(M2.rowptr))[0])
        return "row %d has %d entries vs %d" % (k, M1.rowptr[k + 1] - M1.rowptr[k], M2.rowptr[k + 1] - M2.rowptr[k])
    if not np.array_equal(M1.col, M2.col):
        j = int(np.flatnonzero(M1.col != M2.col)[0])
        k = int(np.searchsorted(M1.rowptr, j, side="right") - 1)
        return "row %d: column %d vs %d" % (k, M1.col[j], M2.col[j])
    return None
