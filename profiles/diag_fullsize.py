"""Diagnostic (round 2): where does the CUDA path leave the oracle at the big shapes?  Operator apply (forward / adjoint)
against the oracle's CSR and fixed-iteration LSMR iterates at 1, 2, 5, 10, 20 iterations.
Usage: python profiles/diag_fullsize.py fixed_384_l3p52"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scipy.sparse import vstack  # noqa: E402

from oracle import denovo3d_oracle as O  # noqa: E402
from helicon_b200.engine import Batch, Problem  # noqa: E402
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec  # noqa: E402

name = sys.argv[1]
d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
a = d["args"]
apix, twist, rise, csym = float(a[0]), float(a[1]), float(a[2]), int(a[3])
so, L3, D2, L2, D3 = int(a[5]), int(a[6]), int(a[7]), int(a[8]), int(a[9])
img = d["image"]
rise_px = rise / apix
prob = Problem(img, 1.0, D2, L2, D3, 0.0, D3 // 2 - 1)
n3 = L3 * prob.ndisk
target = min(MAX_EQUATIONS, int(max(D2 * L2, n3) * so))
batch = Batch(prob, L3, [CandidateSpec(twist, rise_px, csym, target, target, False)])
t0 = time.time()
A_d, b_d, pid = O.build_A_data_matrix_fast(img, 1.0, twist, rise_px, csym, D2, L2, D3, 0, L3, target)
A_s, b_s = O.build_A_helical_sym_matrix(L3, D3, D3, twist, rise_px, csym, 0.0, D3 // 2 - 1, target, "nn")
print(f"oracle build {time.time() - t0:.0f}s rows {A_d.shape[0]}+{A_s.shape[0]} nnz {A_d.nnz}", flush=True)
pidx, kk, jj = batch.data_row_index(0)
nd_pad, tot = batch.rows_padded(0)
print("rows gpu", len(pidx), tot - nd_pad, "b equal", np.array_equal(batch.rhs_padded(0)[pidx], b_d))
rng = np.random.default_rng(5)
x = rng.standard_normal(batch.n).astype(np.float32)
y = batch.apply_forward(0, x)
yd = A_d.astype(np.float64) @ x.astype(np.float64)
ys = (A_s @ x).astype(np.float32)
dd = np.abs(y[:nd_pad][pidx] - yd)
print("forward data: max|d| %.3e (scale %.3e) rows with |d|>1e-3*scale: %d" % (dd.max(), np.abs(yd).max(), int((dd > 1e-3 * np.abs(yd).max()).sum())))
print("forward sym : equal", np.array_equal(y[nd_pad:], ys), "n differing", int((y[nd_pad:] != ys).sum()))
u = np.zeros(tot, np.float32)
ur = rng.standard_normal(len(pidx) + A_s.shape[0]).astype(np.float32)
u[:nd_pad][pidx] = ur[:len(pidx)]
u[nd_pad:] = ur[len(pidx):]
g = batch.apply_adjoint(0, u)
gd = A_d.astype(np.float64).T @ ur[:len(pidx)].astype(np.float64)
gs = A_s.astype(np.float64).T @ ur[len(pidx):].astype(np.float64)
g2 = np.abs(g - (gd + gs))
print("adjoint: max|d| %.3e (scale %.3e) voxels with |d|>1e-3*scale: %d" % (g2.max(), np.abs(gd + gs).max(), int((g2 > 1e-3 * np.abs(gd + gs).max()).sum())))
A = vstack((A_d, A_s)).tocsr()
b = np.concatenate((b_d, np.zeros(A_s.shape[0], np.float32)))
for iters in (1, 2, 5, 10, 20):
    xr = O.lsmr_mixed(A, b, fixed_iters=iters)[0]
    res = batch.solve(fixed_iters=iters, check_every=iters)
    xg = batch.x(0)
    print(f"iters {iters}: rel-L2 {np.linalg.norm(xg - xr) / np.linalg.norm(xr):.3e}  |x| gpu {np.linalg.norm(xg):.6e} ref {np.linalg.norm(xr):.6e}", flush=True)
batch.close(); prob.close()
