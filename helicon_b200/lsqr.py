"""``scipy.sparse.linalg.lsqr`` over the CUDA operator of one candidate -- the solver ``refine_tilt_psi_dy`` uses for
its unbounded systems (SLR:711-714: ``lsqr(A, b, atol=1e-6, btol=1e-6)``).

The reference hands scipy an explicit float32 CSR matrix; scipy wraps it with ``aslinearoperator`` and runs Paige &
Saunders' Golub-Kahan recurrence on it.  Here the SAME scipy routine runs on a ``LinearOperator`` whose two products
are the batch's projector kernels (``hb2_batch_apply_forward`` / ``hb2_batch_apply_adjoint``: explicit-row CSR kernels
or the matrix-free projectors, plus the symmetry-row kernels), so the recurrence, its float32 / float64 mix (u, v, w in
float32, x in float64), the stopping rules (istop 1...7) and the iteration limit 2n are scipy's own; only the summation
order inside a product differs (device row layout, padded rows carry no equation and are masked).  There is no CPU
product: without the CUDA library ``engine`` raises before this module is reached.

The refinement is not the grid-search hot path (one system per Gauss-Newton step), so the O(n) vector updates stay on
the host with scipy; the O(nnz) products -- all of the arithmetic that scales with the matrix -- run on the GPU.
"""

from __future__ import annotations

import numpy as np


def real_row_mask(batch, c=0):
    """Rows of the padded device layout that carry an equation: explicit batches keep their rows in front of the
    padding of the last pseudo view; the planner's symmetry rows follow the padded data block."""
    nd_pad, tot = batch.rows_padded(c)
    real = np.zeros(tot, dtype=bool)
    if hasattr(batch, "m_rows"):  # engine.ExplicitBatch
        real[: batch.m_rows + batch.m_sym_explicit] = True
    else:
        real[batch.data_row_index(c)[0]] = True
    real[nd_pad:] = True
    return real


def rhs_full(batch, c=0):
    """[b_data (padded layout); 0 for every symmetry row] as float32 (SLR:690-698)."""
    nd_pad, tot = batch.rows_padded(c)
    b = np.zeros(tot, dtype=np.float32)
    b[:nd_pad] = batch.rhs_padded(c)
    return b


def operator(batch, c=0, real=None, counter=None):
    """float32 ``LinearOperator`` [A_data; A_hsym] of candidate ``c`` in the padded row layout."""
    from scipy.sparse.linalg import LinearOperator

    real = real_row_mask(batch, c) if real is None else real
    realf = real.astype(np.float32)
    tot, n = len(real), batch.n

    def matvec(x):
        if counter is not None:
            counter[0] += 1
        return batch.apply_forward(c, np.asarray(x, dtype=np.float32).reshape(n)) * realf

    def rmatvec(y):
        if counter is not None:
            counter[1] += 1
        return batch.apply_adjoint(c, np.asarray(y, dtype=np.float32).reshape(tot) * realf)

    return LinearOperator((tot, n), matvec=matvec, rmatvec=rmatvec, dtype=np.float32)


def solve_lsqr(batch, c=0, atol=1e-6, btol=1e-6, conlim=1e8, iter_lim=None, info=None):
    """``lsqr(A, b, atol, btol)[0]`` (SLR:713-714) for candidate ``c`` of ``batch``; ``info`` (a dict) receives istop,
    itn, r1norm, anorm, acond, arnorm, xnorm and the number of device products."""
    from scipy.sparse.linalg import lsqr

    real = real_row_mask(batch, c)
    count = [0, 0]
    op = operator(batch, c, real, count)
    b = rhs_full(batch, c) * real.astype(np.float32)
    res = lsqr(op, b, atol=atol, btol=btol, conlim=conlim, iter_lim=iter_lim)
    if info is not None:
        info.update(istop=int(res[1]), itn=int(res[2]), r1norm=float(res[3]), anorm=float(res[5]), acond=float(res[6]),
                    arnorm=float(res[7]), xnorm=float(res[8]), forward_products=count[0], adjoint_products=count[1])
    return res[0]
