"""Model of the second-generation cluster panel kernel (lu_panel_cl2_kernel, csrc/gpu/lu_kernels.cu) in exact-FMA Python: columns
in blocks of eight, a column step updates only the columns left in its block, the columns right of the block take the block's eight
rank-1 updates at its end, and the pivot rows -- which arrive stale in those columns -- are rebuilt from the pulled row and the
row's own multipliers.  The claim the kernel rests on: every element sees the same sequence of fused multiply-adds as in the
right-looking loop of the first kernels (and of the reference's Crout loop, common/linalg.f90:166-280), so the factors are
bit-identical -- with interchanges in every column, exact ties (last row wins) and panels that are not a multiple of eight wide.
The GPU kernels themselves are compared bit for bit on the device (scripts/lu_ab.py, profiles/r02g_summary.md)."""
from fractions import Fraction as F

import numpy as np
import pytest


def fma(a, b, c):
    return float(F(a) * F(b) + F(c))   # one rounding, as the hardware FMA


def right_looking(A, vv):
    A, vv = A.copy(), vv.copy()
    m, nb = A.shape
    piv = []
    for j in range(nb):
        best, bi = -1., -1
        for i in range(j, m):
            v = vv[i] * abs(A[i, j])
            if v > best or (v == best and i > bi):   # ties -> LAST row (linalg.f90:242)
                best, bi = v, i
        p = bi
        piv.append(p)
        if p != j:
            A[[j, p]] = A[[p, j]]
            vv[p] = vv[j]
        inv = 1.0 / A[j, j]
        for i in range(j + 1, m):
            l = A[i, j] * inv
            A[i, j] = l
            for c in range(j + 1, nb):
                A[i, c] = fma(-l, A[j, c], A[i, c])
    return A, piv


def delayed(A, vv, IB=8):
    A, vv = A.copy(), vv.copy()
    m, nb = A.shape
    piv = []
    U = np.zeros((IB, nb))
    for jj in range(nb):
        blk0 = jj & ~(IB - 1)
        blkend = min(blk0 + IB, nb)
        best, bi = -1., -1
        for i in range(jj, m):
            v = vv[i] * abs(A[i, jj])
            if v > best or (v == best and i > bi):
                best, bi = v, i
        p = bi
        piv.append(p)
        pulled, oldj, vvj = A[p].copy(), A[jj].copy(), vv[jj]
        for c in range(blkend, nb):                      # the pivot row in the delayed columns (s_U of the kernel)
            u = pulled[c]
            for b in range(jj - blk0):
                u = fma(-pulled[blk0 + b], U[b, c], u)
            U[jj - blk0, c] = u
        if p != jj:                                      # whole-row interchange, stale delayed columns and all
            A[jj], A[p], vv[p] = pulled, oldj, vvj
        inv = 1.0 / pulled[jj]
        for i in range(jj + 1, m):
            l = A[i, jj] * inv
            A[i, jj] = l
            for c in range(jj + 1, blkend):
                A[i, c] = fma(-l, pulled[c], A[i, c])
        if jj + 1 == blkend and blkend < nb:             # end of a complete block
            for b in range(IB):
                A[blk0 + b, blkend:] = U[b, blkend:]
            for i in range(blkend, m):
                for c in range(blkend, nb):
                    v = A[i, c]
                    for b in range(IB):
                        v = fma(-A[i, blk0 + b], U[b, c], v)
                    A[i, c] = v
    return A, piv


@pytest.mark.parametrize("m, nb, seed", [(40, 20, 1), (30, 16, 2), (70, 27, 3), (25, 8, 4), (33, 7, 5)])
def test_delayed_updates_give_the_right_looking_factors_bit_for_bit(m, nb, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, nb))
    vv = 1. / np.abs(rng.standard_normal(m) + 2.)
    A[5], vv[5] = A[9], vv[9]                            # an exact tie of vv * |a| in every column
    R, p1 = right_looking(A, vv)
    D, p2 = delayed(A, vv)
    assert p1 == p2
    assert any(p != j for j, p in enumerate(p1))         # interchanges did happen
    assert np.array_equal(R.view(np.uint64), D.view(np.uint64))
