"""Small fixed workload for ncu captures: a cfg2 batch (256x256, L3=12) of NC
candidates, NI LSMR iterations.  usage: python profiles/prof_run.py [NC] [NI]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from helicon_b200.engine import Batch, Problem
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ni = int(sys.argv[2]) if len(sys.argv) > 2 else 6
img = bench.synthetic_filament()
tasks = bench.grid_tasks()
g = tasks[0].geom
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
sel = tasks[20000:20000 + nc]
batch = Batch(prob, g["L3"], [CandidateSpec(t.twist, t.rise / g["apix3d"], 1, target, target, False) for t in sel])
res = batch.solve(fixed_iters=ni, check_every=ni, profile=1)
print("itn", res["itn"][:4], "score", res["score"][:4], batch.timing())
batch.close(); prob.close()
