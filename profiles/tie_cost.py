"""Cost of the exact tie paths at the cfg2 shape: a batch of 100 candidates whose rise makes h*rise_pixel a half-integer
(rise 4.75 A at 1.3 A/px: tie views h = +-13, +-39), one whose twist puts views on 30-degree multiples (twist -1.2: h =
+-25), and a regular batch.  usage: python profiles/tie_cost.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from helicon_b200.engine import Batch, Problem
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec

img = bench.synthetic_filament()
tasks = bench.grid_tasks()
g = tasks[0].geom
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
tw = np.linspace(-3.0, -0.2, 100)
cases = {"regular (rise 4.73)": [(t, 4.73) for t in tw], "z ties (rise 4.75)": [(t, 4.75) for t in tw],
         "z ties on every odd h (rise 4.55 = 3.5 px)": [(t, 4.55) for t in tw[::4]],
         "xy ties (twist -1.2, -1.5, -2.0, -2.5, -3.0)": [(t, r) for t in (-1.2, -1.5, -2.0, -2.5, -3.0) for r in np.linspace(4.41, 5.09, 20)]}
for name, cl in cases.items():
    batch = Batch(prob, g["L3"], [CandidateSpec(t, r / g["apix3d"], 1, target, target, False) for t, r in cl])
    batch.solve(fixed_iters=2, check_every=2)
    res = batch.solve(fixed_iters=12, check_every=12, profile=1)
    tm = batch.timing()
    nv = int(batch.plan.cands["view_count"].sum())
    print(f"{name}: {len(cl)} candidates, {nv} view slots, tie views {batch.plan.n_tie}, exact-map angles {len(batch.nvalid) - len(batch.plan.angles)} | "
          f"per candidate-pass (us): fwd_data {tm['fwd_data_ms']/13/len(cl)*1e3:.1f} adj {tm['adj_ms']/13/len(cl)*1e3:.1f} "
          f"lsmr/iter {tm['lsmr_ms']/13/len(cl)*1e3:.1f}", flush=True)
    batch.close()
