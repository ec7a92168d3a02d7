import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
from helicon_b200.engine import Batch, Problem
from helicon_b200.grid import build_tasks
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec
N=512
img = bench.synthetic_filament(n=N, apix=1.3, diameter=0.3 * N * 1.3)
tasks, _ = build_tasks(N, N, 1.3, np.array([-1.3,-1.7]), np.array([4.8, 4.85,4.9,4.95]), (1,), 3, None, 0.0, None, 0, -1, 0)
g = tasks[0].geom
tl = [t for t in tasks if t.geom["L3"] == g["L3"]]
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
batch = Batch(prob, g["L3"], [CandidateSpec(t.twist, t.rise / g["apix3d"], 1, target, target, False) for t in tl])
res = batch.solve(fixed_iters=4, check_every=4, profile=1)
print(len(tl), batch.timing())
