"""CPU tests of ngsamg_b200/refdump.py: the reader / writer of the reference's debug text dumps (print_tm_spmat,
src/base/utils/utils_io.hpp:102-132; files written at base_factory.cpp:506-522 and vertex_factory_impl.hpp:2422-2428)."""
import os

import numpy as np
import pytest

import ngsamg_b200 as ng
from ngsamg_b200 import refdump as RD
from ngsamg_b200 import synthetic as S
from helpers import host_hierarchy, to_oracle, to_product
from oracle import oracle as O

# what print_tm_spmat writes for a 2 x 3 block-row matrix of Mat<3,6> blocks (setw(6) row number, setw(4) columns and values,
# " | " after every block, H text lines per block row, "(empty)" rows), typed out from the format code
REF_BLOCK_DUMP = """Row      0:    0:    1    0    0    0  0.5 -0.25  |    2:    1    0    0    0 -0.5 0.25  | 
          :     :    0    1    0 -0.5    0 0.125  |     :    0    1    0  0.5    0 -0.125  | 
          :     :    0    0    1 0.25 -0.125    0  |     :    0    0    1 -0.25 0.125    0  | 
Row      1: (empty)
"""

# NGSolve's operator<< of a SparseMatrix<double>
REF_SCALAR_DUMP = """Row 0:   0: 2.66667   1: -0.333333   3: -1e-05
Row 1:   0: -0.333333   1: 2.66667
Row 2:
Row 3:   0: -1e-05   3: 1
"""


def test_parse_reference_block_format():
    P = RD.parse_spmat(REF_BLOCK_DUMP)
    assert (P.nrows, P.ncols, P.bh, P.bw) == (2, 3, 3, 6)
    assert list(P.rowptr) == [0, 2, 2] and list(P.col) == [0, 2]
    v = P.val.reshape(2, 3, 6)
    assert np.array_equal(v[0, :, :3], np.eye(3)) and np.array_equal(v[1, :, :3], np.eye(3))
    assert np.allclose(v[0, :, 3:], [[0, 0.5, -0.25], [-0.5, 0, 0.125], [0.25, -0.125, 0]])
    assert np.allclose(v[1, :, 3:], -v[0, :, 3:])


def test_parse_reference_scalar_format():
    A = RD.parse_spmat(REF_SCALAR_DUMP)
    assert (A.nrows, A.ncols, A.bh, A.bw) == (4, 4, 1, 1)
    assert list(A.rowptr) == [0, 3, 5, 5, 7] and list(A.col) == [0, 1, 3, 0, 1, 0, 3]
    assert np.allclose(A.val, [2.66667, -0.333333, -1e-05, -0.333333, 2.66667, -1e-05, 1])
    with pytest.raises(ValueError):
        RD.parse_spmat(REF_SCALAR_DUMP, ncols=3)


@pytest.mark.parametrize("shape", [(1, 1), (3, 3), (3, 6), (6, 3), (6, 6), (2, 3)])
def test_round_trip(shape):
    import scipy.sparse as sp
    bh, bw = shape
    rng = np.random.default_rng(7)
    pat = sp.random(40, 25, density=0.15, random_state=rng, format="csr")
    pat.sort_indices()
    M = ng.SparseMatrix(40, 25, bh, bw, pat.indptr, pat.indices, rng.standard_normal((pat.nnz, bh, bw)))
    # lossless
    M2 = RD.parse_spmat(RD.format_spmat(M, precision=17), ncols=25, nrows=40)
    assert RD.compare_patterns(M, M2) is None and np.array_equal(M.val, M2.val)
    # the reference's default precision: pattern exact, values to 6 digits
    M3 = RD.parse_spmat(RD.format_spmat(M), ncols=25, nrows=40)
    assert RD.compare_patterns(M, M3) is None and np.allclose(M.val, M3.val, rtol=1e-5)
    # the writer reproduces the reference's layout token by token on the hand-typed sample
    if shape == (3, 6):
        P = RD.parse_spmat(REF_BLOCK_DUMP)
        assert RD.format_spmat(P).split() == REF_BLOCK_DUMP.split()


def test_hierarchy_files_round_trip(tmp_path):
    """a hierarchy written in the reference's file naming is read back level by level; the injected prolongations reproduce the
    same coarse sparsity patterns (the integer side a dump pins bit-exactly)"""
    p = S.poisson3d_kuhn(9)
    A = ng.SparseMatrix(p["n"], p["n"], 1, 1, p["rowptr"], p["col"], p["val"])
    prols = host_hierarchy(A, p["free"], max_coarse=30)
    assert len(prols) >= 2
    mats = [A]
    for P in prols:
        Po = to_oracle(P)
        mats.append(to_product(O.restrict_matrix(O.transpose(Po), to_oracle(mats[-1]), Po)))
    d = str(tmp_path)
    for l, M in enumerate(mats):
        RD.write_spmat(os.path.join(d, RD.mat_file(l)), M)
    for l, P in enumerate(prols):
        RD.write_spmat(os.path.join(d, RD.prol_file(l)), P)
    mats2, prols2 = RD.load_hierarchy(d)
    assert len(mats2) == len(mats) and len(prols2) == len(prols)
    for a, b in zip(mats + prols, mats2 + prols2):
        assert RD.compare_patterns(a, b) is None and a.ncols == b.ncols
        assert np.allclose(a.val, b.val, rtol=2e-5, atol=1e-12)
    # Galerkin patterns from the re-read prolongations == patterns of the dumped coarse matrices
    cur = to_oracle(mats2[0])
    for l, P in enumerate(prols2):
        Po = to_oracle(P)
        cur = O.restrict_matrix(O.transpose(Po), cur, Po)
        assert RD.compare_patterns(to_product(cur), mats2[l + 1]) is None
    assert RD.compare_patterns(mats2[0], mats2[1]) is not None
