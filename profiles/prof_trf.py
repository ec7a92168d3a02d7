"""cfg2 batch with the reference's default positive-constraint rule (bounded TRF branch). usage: python profiles/prof_trf.py [NC]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from helicon_b200.engine import Batch, Problem
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec, positive_rule

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 32
img = bench.synthetic_filament()
tasks = bench.grid_tasks()
g = tasks[0].geom
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
sel = tasks[20000:20000 + nc]
batch = Batch(prob, g["L3"], [CandidateSpec(t.twist, t.rise / g["apix3d"], 1, target, target,
                                            positive_rule(-1, t.rise / g["apix3d"], t.twist, g["L3"])) for t in sel])
res = batch.solve()
tm = batch.timing()
print("itn", res["itn"][:4], "trf_nit", res["trf_nit"][:8], "score", res["score"][:4])
print({k: (round(float(v), 1) if isinstance(v, float) else v) for k, v in tm.items()})
batch.close(); prob.close()
