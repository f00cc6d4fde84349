"""oracle/oracle_bgs.py -- CPU restatement of the reference's block Gauss-Seidel smoother (sm_type = bgs).  TEST INFRASTRUCTURE ONLY.

Follows src/base/smoothers/loc_block_gssmoother_impl.hpp of /root/reference:
  * BSBlock::RichardsonUpdate      (:244-268)  smallrhs = b_B - A_{B,off} x - DB x_B ; x_B += omega * DB^-1 smallrhs
  * BSBlock::RichardsonUpdate_RES  (:516-541)  upd = omega * DB^-1 res_B ; x_B += upd ; res -= A_{:,B} upd   (transposed rows: symmetric A)
  * BSmoother2::IterateBlocks      (:618-651)  blocks ascending, or descending when `reverse`
  * BSmoother2::SmoothWO           (:656-669)  res_updated && update_res -> SmoothRESSimple, else SmoothSimple (+ res = b - A x)
  * blocks: GetGSBlocks (src/base/precond/amg_pc_vertex_impl.hpp:1171-1269): block cv = { v : vmap[v] == cv }, ascending; vertices with
    vmap == -1 are in no block and are not smoothed.
omega = 1 (SmoothSimple / SmoothRESSimple pass 1.0).  PIN: the update routines, the block order and the flag protocol are checked against the
reference's OWN code -- RichardsonUpdate, RichardsonUpdate_RES, IterateBlocks, SmoothWO, SmoothSimple, SmoothRESSimple cut out of
/root/reference at build time and compiled against oracle/ref_pin/ngs_standin_bgs.hpp into oracle/_ref/libngsamg_ref_bgs.so
(tests/test_oracle_bgs.py: every flag combination, forward / reverse, scalar and 3x3 blocks, <= 1e-13; both sides use numpy's dense block
inverses), and so is the block set-up BSBlock::SetFromSPMat (sorted dofnrs, off-block rows, D_B; its CalcInverse is NGSolve's and replaced by a
Gauss-Jordan in the stand-in: 1e-11).  Unpinned: GetGSBlocks (a six-line grouping of the vertices by coarse vertex) and the pseudo-inverse variant."""
import numpy as np


class BlockGS:
    def __init__(self, A_scipy, b, block_of):
        """A_scipy: scalar CSR matrix of the level (n*b rows); b: dofs per vertex; block_of[v]: block of vertex v (-1 = none)"""
        self.A = A_scipy.tocsr()
        self.b = int(b)
        block_of = np.asarray(block_of).astype(np.int64)
        nb = int(block_of.max()) + 1 if len(block_of) else 0
        self.blocks = []
        for k in range(nb):
            verts = np.flatnonzero(block_of == k)                       # ascending (TableCreator order)
            if len(verts) == 0:
                continue
            dofs = (verts[:, None] * self.b + np.arange(self.b)[None, :]).ravel()
            DB = self.A[dofs][:, dofs].toarray()
            self.blocks.append((dofs, DB, np.linalg.inv(DB)))

    def _order(self, reverse):
        return reversed(self.blocks) if reverse else self.blocks

    def smooth_simple(self, x, b, reverse=False):                      # SmoothSimple, one step
        for dofs, DB, DBinv in self._order(reverse):
            r = b[dofs] - self.A[dofs] @ x                              # b_B - A_{B,:} x  (off-block rows + diag * x_B)
            x[dofs] += DBinv @ r

    def smooth_res_simple(self, x, res, reverse=False):                # SmoothRESSimple, one step
        AT = self.A.T.tocsr()
        for dofs, DB, DBinv in self._order(reverse):
            upd = DBinv @ res[dofs]
            x[dofs] += upd
            res -= AT[dofs].T @ upd                                     # res -= A_{:,B} upd through the transposed rows

    def smooth(self, x, b, res, res_updated, update_res, x_zero, reverse=False):   # SmoothWO
        if res_updated and update_res:
            self.smooth_res_simple(x, res, reverse)
        else:
            self.smooth_simple(x, b, reverse)
            if update_res:
                res[:] = b - self.A @ x


def vcycle(levels, prols, smoothers, coarse_solve, b):
    """AMGMatrix::SmoothV (src/base/solve/amg_matrix.cpp:160-307) with arbitrary smoother objects: levels = scalar CSR matrices,
    prols = scalar CSR prolongations, smoothers[l].smooth(x, b, res, ru, ur, xz, reverse)."""
    nl = len(levels)
    xs, bs, rs = [None] * nl, [None] * nl, [None] * nl
    bs[0] = np.array(b, dtype=np.float64)
    for l in range(nl - 1):
        xs[l] = np.zeros_like(bs[l])
        rs[l] = bs[l].copy()
        smoothers[l].smooth(xs[l], bs[l], rs[l], True, True, True, False)
        bs[l + 1] = prols[l].T @ rs[l]
    xs[nl - 1] = coarse_solve(bs[nl - 1])
    for l in range(nl - 2, -1, -1):
        xs[l] += prols[l] @ xs[l + 1]
        smoothers[l].smooth(xs[l], bs[l], rs[l], False, False, False, True)
    return xs[0]
