"""Shared test helpers (CSR fixtures, tolerances)."""
import json
import os

import numpy as np
from scipy.sparse import csr_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def csr_from(d, prefix="A"):
    return csr_matrix(
        (d[prefix + "_data"], d[prefix + "_indices"], d[prefix + "_indptr"]),
        shape=tuple(int(v) for v in d[prefix + "_shape"]),
    )


def canon(A):
    A = A.tocsr().copy()
    A.sum_duplicates()
    A.sort_indices()
    return A


def csr_equal(A, B, tol=0.0):
    """Same shape, same sparsity pattern (explicit zeros included), values within tol."""
    A, B = canon(A), canon(B)
    if A.shape != B.shape:
        return False, f"shape {A.shape} vs {B.shape}"
    if not np.array_equal(A.indptr, B.indptr):
        return False, "indptr differs"
    if not np.array_equal(A.indices, B.indices):
        return False, "indices differ"
    d = float(np.max(np.abs(A.data - B.data))) if A.nnz else 0.0
    if d > tol:
        return False, f"max |dA| = {d}"
    return True, ""


def cases(kind, interpolation=None):
    out = []
    for k, v in manifest().items():
        if k.startswith("_"):
            continue
        if v["kind"] == kind and (interpolation is None or v["interpolation"] == interpolation):
            out.append(k)
    return sorted(out)


def oracle_permutation_floor(details, nperm=3):
    """The oracle's (= the reference's scipy call's) own reproducibility at its stopping point: the same equations in
    permuted row order (same maths, other float32 summation order) -> max rel-L2 change of x."""
    from scipy.optimize import lsq_linear
    from scipy.sparse import vstack

    A_d, b_d, A_s, x0 = details["A_data"], details["b_data"], details["A_hsym"], details["x"]
    A = vstack((A_d, A_s)).tocsr() if A_s is not None else A_d.tocsr()
    b = np.concatenate((b_d, np.zeros(A.shape[0] - len(b_d), np.float32)))
    lb, ub = (0.0, float(np.max(b_d))) if details["positive"] else (-np.inf, np.inf)
    floor = 0.0
    for seed in range(nperm):
        p = np.random.default_rng(seed).permutation(A.shape[0])
        x = lsq_linear(A[p].tocsr(), b[p], bounds=(lb, ub), tol=1e-2, max_iter=200, lsmr_maxiter=1000, lsmr_tol="auto").x
        floor = max(floor, float(np.linalg.norm(x.astype(np.float32) - x0) / np.linalg.norm(x0)))
    return floor
