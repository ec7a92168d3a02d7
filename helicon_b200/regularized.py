"""The sklearn-model branch of ``solve_equations`` (SLR:272-342): ElasticNet / Lasso / Ridge on the stacked system
``[A_data; A_hsym] x = [b_data; 0]`` with ``fit_intercept=True``.

The reference hands the explicit sparse matrix to scikit-learn's coordinate descent (``selection="random"``, ``tol=1e-2``,
``max_iter=200``): a SEQUENTIAL sweep over the unknowns, stochastic (numpy's global RNG) and loosely converged -- its
result moves by ~1e-1 (rel-L2 of x) and ~5e-3 (score) against the minimiser of its own objective
(tests/golden/model_*.npz, oracle/make_golden_models.py).  Here the same strictly convex objectives are minimised with
the matrix-free CUDA operator (``Batch.apply_forward`` / ``apply_adjoint``: the projector kernels of the LSMR path;
nothing is materialised): accelerated proximal gradient (FISTA with adaptive restart) for the l1 / non-negative models,
conjugate gradients on the normal equations for the unconstrained ridge.  The minimiser is unique, so parity with the
reference is STATISTICAL by construction: the result lies inside the reference's own convergence band and reaches an
objective value <= the reference's (both asserted in tests/test_gpu_models.py).

Objectives (scikit-learn's, with the intercept profiled out: X_c, y_c are column- / mean-centred over the m real rows):
  ElasticNet / Lasso:  1/(2m) ||y_c - X_c w||^2 + alpha*l1_ratio*||w||_1 + alpha*(1-l1_ratio)/2*||w||^2   [w >= 0 if positive]
  Ridge:               ||y_c - X_c w||^2 + alpha*||w||^2                                                   [w >= 0 if positive]
The host drives the iteration (vector updates on n ~ 1e5...1e6 unknowns in numpy); every X / X^T product runs on the GPU.
"""

from __future__ import annotations

import numpy as np

MODELS = ("elasticnet", "lasso", "ridge")


class _CenteredOperator:
    """X_c = (I - 11^T/m) X restricted to the REAL rows of one candidate (padded rows of the device layout carry no
    equation and are masked out), y_c likewise."""

    def __init__(self, batch, c, row_keep=None):
        self.batch, self.c = batch, int(c)
        nd_pad, tot = batch.rows_padded(c)
        pidx, kk, jj = batch.data_row_index(c)
        real = np.zeros(tot, dtype=bool)
        if row_keep is not None:  # half sets: only the data rows whose pixel is kept
            pidx = pidx[np.asarray(row_keep, dtype=bool)[kk * batch.problem.D2 + jj]]
        real[pidx] = True
        real[nd_pad:] = True
        self.real, self.m, self.n = real, int(real.sum()), batch.n
        self.realf = real.astype(np.float32)
        y = np.zeros(tot, dtype=np.float64)
        y[:nd_pad] = batch.rhs_padded(c)
        y[~real] = 0.0
        self.ybar = y.sum() / self.m
        self.yc = np.where(real, y - self.ybar, 0.0)
        self.xbar = batch.apply_adjoint(c, self.realf).astype(np.float64) / self.m  # column means of X
        self.napply = 1

    def X(self, w):
        r = self.batch.apply_forward(self.c, w.astype(np.float32)).astype(np.float64)
        r -= float(self.xbar @ w)
        r[~self.real] = 0.0
        self.napply += 1
        return r

    def XT(self, r):
        r = np.where(self.real, r, 0.0)
        g = self.batch.apply_adjoint(self.c, r.astype(np.float32)).astype(np.float64)
        self.napply += 1
        return g - self.xbar * r.sum()

    def lipschitz(self, iters=20, seed=0):
        """Largest eigenvalue of X_c^T X_c by power iteration (a few percent above is all FISTA needs)."""
        v = np.random.default_rng(seed).standard_normal(self.n)
        v /= np.linalg.norm(v)
        lam = 1.0
        for _ in range(iters):
            z = self.XT(self.X(v))
            lam = float(np.linalg.norm(z))
            if lam == 0:
                return 1.0
            v = z / lam
        return lam * 1.05


def _enet_gap(op, w, R, alpha_s, beta_s, positive):
    """scikit-learn's duality gap of the (un-normalised) elastic-net problem
    1/2||y - Xw||^2 + alpha_s||w||_1 + beta_s/2||w||^2 (linear_model/_cd_fast.pyx: enet_coordinate_descent)."""
    XtA = op.XT(R) - beta_s * w
    dual_norm = float(XtA.max()) if positive else float(np.abs(XtA).max())
    R2, w2 = float(R @ R), float(w @ w)
    if dual_norm > alpha_s:
        const = alpha_s / dual_norm
        gap = 0.5 * (R2 + R2 * const * const)
    else:
        const, gap = 1.0, R2
    return gap + alpha_s * float(np.abs(w).sum()) - const * float(R @ op.yc) + 0.5 * beta_s * (1 + const * const) * w2


def solve_model(batch, c, algorithm, positive, row_keep=None, gap_factor=0.05, max_iter=4000, info=None):
    """coef_ of the reference's model for candidate ``c`` of a created (not necessarily solved) batch, float32 in the
    reference's unknown order.  ``gap_factor``: stop at gap_factor x the duality gap scikit-learn itself accepts
    (tol=1e-2 times ||y_c||^2), i.e. 20x tighter than the reference by default."""
    model = algorithm.get("model")
    if model not in MODELS:
        raise NotImplementedError(f"helicon_b200: algorithm model {model!r} is not implemented on the CUDA path "
                                  "(lsq, elasticnet, lasso, ridge are; 'lreg' / 'ard' need the dense matrix)")
    alpha = float(algorithm.get("alpha", 1 if model == "ridge" else 1e-4))
    l1_ratio = 1.0 if model == "lasso" else (0.0 if model == "ridge" else float(algorithm.get("l1_ratio", 0.5)))
    op = _CenteredOperator(batch, c, row_keep)
    while True:
        w, stats = (_ridge(op, alpha, positive, max_iter) if model == "ridge"
                    else _enet(op, alpha, l1_ratio, positive, gap_factor, max_iter))
        if np.any(w) or model == "ridge":
            break
        alpha *= 0.1  # SLR:331-335: an all-zero coef_ is refitted with a ten times smaller alpha
    if info is not None:
        info.update(stats, alpha=alpha, l1_ratio=l1_ratio, operator_applies=op.napply, rows=op.m)
    return w.astype(np.float32)


def _enet(op, alpha, l1_ratio, positive, gap_factor, max_iter):
    m = op.m
    alpha_s, beta_s = alpha * l1_ratio * m, alpha * (1 - l1_ratio) * m  # scikit-learn's un-normalised form
    L = op.lipschitz() + beta_s
    step = 1.0 / L
    w = np.zeros(op.n)
    z, t = w.copy(), 1.0
    tol_gap = gap_factor * 1e-2 * float(op.yc @ op.yc)
    gap, it = np.inf, 0
    for it in range(1, max_iter + 1):
        Rz = op.yc - op.X(z)
        grad = -op.XT(Rz) + beta_s * z
        w_new = z - step * grad
        w_new = np.sign(w_new) * np.maximum(np.abs(w_new) - step * alpha_s, 0.0)
        if positive:
            np.maximum(w_new, 0.0, out=w_new)
        if float((z - w_new) @ (w_new - w)) > 0:  # O'Donoghue-Candes gradient restart
            t_new, z = 1.0, w_new.copy()
        else:
            t_new = 0.5 * (1 + np.sqrt(1 + 4 * t * t))
            z = w_new + ((t - 1) / t_new) * (w_new - w)
        w, t = w_new, t_new
        if it % 10 == 0:
            gap = _enet_gap(op, w, op.yc - op.X(w), alpha_s, beta_s, positive)
            if gap <= tol_gap:
                break
    return w, dict(iterations=it, gap=float(gap), gap_tolerance_sklearn=1e-2 * float(op.yc @ op.yc), lipschitz=L)


def _ridge(op, alpha, positive, max_iter):
    """min ||y_c - X_c w||^2 + alpha ||w||^2: CG on (X_c^T X_c + alpha I) w = X_c^T y_c; projected FISTA if positive."""
    rhs = op.XT(op.yc)
    if not positive:
        w = np.zeros(op.n)
        r = rhs.copy()
        p, rs = r.copy(), float(r @ r)
        rs0, it = rs, 0
        for it in range(1, max_iter + 1):
            Ap = op.XT(op.X(p)) + alpha * p
            a = rs / float(p @ Ap)
            w += a * p
            r -= a * Ap
            rs_new = float(r @ r)
            if rs_new <= 1e-12 * rs0:  # relative residual 1e-6 (sklearn's own sparse_cg runs at tol=1e-2)
                break
            p = r + (rs_new / rs) * p
            rs = rs_new
        return w, dict(iterations=it, residual=float(np.sqrt(rs_new / max(rs0, 1e-300))))
    L = op.lipschitz() + alpha
    step = 1.0 / L
    w = np.zeros(op.n)
    z, t, it, dw = w.copy(), 1.0, 0, np.inf
    for it in range(1, max_iter + 1):
        grad = op.XT(op.X(z)) + alpha * z - rhs
        w_new = np.maximum(z - step * grad, 0.0)
        if float((z - w_new) @ (w_new - w)) > 0:
            t_new, z = 1.0, w_new.copy()
        else:
            t_new = 0.5 * (1 + np.sqrt(1 + 4 * t * t))
            z = w_new + ((t - 1) / t_new) * (w_new - w)
        dw = float(np.linalg.norm(w_new - w)) / max(float(np.linalg.norm(w_new)), 1e-300)
        w, t = w_new, t_new
        if dw <= 1e-7 and it > 10:
            break
    return w, dict(iterations=it, last_relative_update=dw, lipschitz=L)
