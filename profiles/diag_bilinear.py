"""Sensitivity of fixed-iteration LSMR iterates to the float32 summation order (norm mode, row layout) on a trilinear
system: explicit rows vs the matrix-free operator.  usage: python profiles/diag_bilinear.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_bilinear import _image
from helicon_b200.bilinear import BilinearBatch
from helicon_b200.engine import ExplicitBatch, Problem
from helicon_b200.planner import CandidateSpec

ni = int(sys.argv[1]) if len(sys.argv) > 1 else 30
N, L3 = 48, 8
img = _image(N, seed=9)
prob = Problem(img, 1.0, N, N, N, 0.0, N // 2 - 1)
sp = CandidateSpec(-2.1, 2.9, 1, 0, 30000, False)
xs = {}
for nm in (1, 0):
    eb = ExplicitBatch(prob, L3, sp, interpolation="linear")
    r = eb.solve(fixed_iters=ni, check_every=ni, norm_mode=nm)
    xs[f"explicit nm{nm}"] = (eb.x(0), float(r["score"][0]), r[0].copy())
    eb.close()
    bb = BilinearBatch(prob, L3, [sp])
    r = bb.solve(fixed_iters=ni, check_every=ni, norm_mode=nm)
    xs[f"matrix-free nm{nm}"] = (bb.x(0), float(r["score"][0]), r[0].copy())
    bb.close()
keys = list(xs)
for k in keys:
    r = xs[k][2]
    print(f"{k:18s} score {xs[k][1]:.7f} normr {float(r['normr']):.6f} normar {float(r['normar']):.6e} normA {float(r['normA']):.5f} normx {float(r['normx']):.5f}")
for i, a in enumerate(keys):
    for b in keys[i + 1:]:
        rel = float(np.linalg.norm(xs[a][0] - xs[b][0]) / np.linalg.norm(xs[b][0]))
        print(f"  {a:18s} vs {b:18s}: rel-L2(x) {rel:.2e} |dscore| {abs(xs[a][1] - xs[b][1]):.2e}")
prob.close()
