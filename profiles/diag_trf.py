"""Diagnostic (round 2): bounded-branch trace of ONE candidate on the GPU next to the oracle's restated scipy TRF
(oracle.trf_linear_restated, pinned to scipy bit for bit) started from the GPU's own LSMR solution and from scipy's.
Usage: python profiles/diag_trf.py N twist rise"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import warnings  # noqa: E402

warnings.filterwarnings("ignore")
from scipy.sparse import vstack  # noqa: E402
from scipy.sparse.linalg import lsmr  # noqa: E402

import bench  # noqa: E402
from oracle import denovo3d_oracle as O  # noqa: E402
from helicon_b200.engine import Batch, Problem  # noqa: E402
from helicon_b200.grid import derive_geometry  # noqa: E402
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec  # noqa: E402

N, twist, rise = int(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3])
apix = 1.3
g = derive_geometry(N, N, apix, rise, rise, N * apix, 0.0, N * apix, 3 * rise, apix, 0, -1)
img = bench.synthetic_filament(n=N)
D2, L2, D3, L3 = g["D2"], g["L2"], g["D3"], g["L3"]
rise_px = rise / apix
prob = Problem(img, 1.0, D2, L2, D3, 0.0, D3 // 2 - 1)
n3 = L3 * prob.ndisk
target = min(MAX_EQUATIONS, int(max(D2 * L2, n3) * g["sym_oversample"]))
batch = Batch(prob, L3, [CandidateSpec(twist, rise_px, 1, target, target, True)])
res = batch.solve()
tr, nit, status = batch.trf_trace(0)
print("GPU: itn", res[0]["itn"], "trf nit", nit, "status", status, "score", res[0]["score"])
for i, r in enumerate(tr):
    print("  gpu it %2d cost %.6f g_norm %.4e inner %d kind %d vals %.5e %.5e %.5e change %.5e" % ((i,) + tuple(r)))
x_gpu = batch.x(0)
A_d, b_d, pid = O.build_A_data_matrix_fast(img, 1.0, twist, rise_px, 1, D2, L2, D3, 0, L3, target)
A_s, b_s = O.build_A_helical_sym_matrix(L3, D3, D3, twist, rise_px, 1, 0.0, D3 // 2 - 1, target, "nn")
A = vstack((A_d, A_s)).tocsr()
b = np.concatenate((b_d, b_s)).astype(np.float32)
t0 = time.time()
r0 = lsmr(A, b, maxiter=1000, atol=1e-4, btol=1e-4)
print("scipy lsmr itn", r0[2], "%.0fs" % (time.time() - t0), flush=True)
# the GPU's unbounded LSMR solution: solve again without the positive rule
batch2 = Batch(prob, L3, [CandidateSpec(twist, rise_px, 1, target, target, False)])
res2 = batch2.solve()
x_lsq_gpu = batch2.x(0).astype(np.float64)
print("GPU lsmr itn", res2[0]["itn"], "rel-L2 of the two LSMR solutions", np.linalg.norm(x_lsq_gpu - r0[0]) / np.linalg.norm(r0[0]))
ub = float(np.max(b_d))
for name, xl in (("scipy x_lsq", r0[0]), ("GPU x_lsq", x_lsq_gpu)):
    trace = []
    xo = O.trf_linear_restated(A, b, xl.copy(), 0.0, ub, trace=trace)
    xo = xo[0] if isinstance(xo, tuple) else xo
    sc = O.cosine_similarity(A_d.dot(np.asarray(xo).astype(np.float32)), b_d)
    print("oracle TRF from", name, ": nit", len(trace), "score", float(sc), "rel-L2 vs GPU x", np.linalg.norm(np.asarray(xo) - x_gpu) / np.linalg.norm(xo))
    for t in trace:
        print("  cpu it %2d cost %.6f g_norm %.4e inner %d kind %s p %.5e r %.5e" % (t["it"], t["cost"], t["g_norm"], t["inner_itn"], t["kind"], t.get("p_value") or 0, t.get("r_value") or 0))
batch.close(); batch2.close(); prob.close()
