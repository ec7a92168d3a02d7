"""A batch of candidates whose every odd h is a column->slice tie view (rise 3.5 px) for ncu captures of the tie kernels.
usage: python profiles/prof_tie.py [NC]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from helicon_b200.engine import Batch, Problem
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 16
img = bench.synthetic_filament()
tasks = bench.grid_tasks()
g = tasks[0].geom
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
tw = np.linspace(-3.0, -0.2, nc)
batch = Batch(prob, g["L3"], [CandidateSpec(t, 4.55 / g["apix3d"], 1, target, target, False) for t in tw])
batch.solve(fixed_iters=2, check_every=2)
res = batch.solve(fixed_iters=4, check_every=4, profile=1)
print(batch.timing())
