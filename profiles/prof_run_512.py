"""512x512 (BASELINE config 3 shape) batch for ncu captures / launch timing.  usage: python profiles/prof_run_512.py [n_twists] [n_rises]"""
import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
from helicon_b200.engine import Batch, Problem
from helicon_b200.grid import build_tasks
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec
N = int(os.environ.get("PROF_N", "512"))
img = bench.synthetic_filament(n=N, apix=1.3, diameter=0.3 * N * 1.3)
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 2
NR = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tasks, _ = build_tasks(N, N, 1.3, np.array([-1.3, -1.7][:NT]), np.array([4.8, 4.85, 4.9, 4.95][:NR]), (1,), 3, None, 0.0, None, 0, -1, 0)
g = tasks[0].geom
tl = [t for t in tasks if t.geom["L3"] == g["L3"]]
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
batch = Batch(prob, g["L3"], [CandidateSpec(t.twist, t.rise / g["apix3d"], 1, target, target, False) for t in tl])
batch.solve(fixed_iters=2, check_every=2)
res = batch.solve(fixed_iters=4, check_every=4, profile=1)
tm = batch.timing()
print(len(tl), "candidates; per launch (ms): fwd_data", tm["fwd_data_ms"] / tm["fwd_data_launches"], "adj", tm["adj_ms"] / tm["adj_launches"],
      "views", batch.plan.cands["view_count"].tolist(), "angles", len(batch.nvalid))
import time
for prof in (0, 1):
    t0 = time.perf_counter()
    batch.solve(fixed_iters=50, check_every=50, profile=prof)
    dt = time.perf_counter() - t0
    tm = batch.timing()
    print(f"profile={prof}: wall {dt*1e3/51:.3f} ms per iteration; lsmr_ms/51 = {tm['lsmr_ms']/51:.3f}; fwd_data per launch {tm['fwd_data_ms']/max(1,tm['fwd_data_launches']):.3f} fwd_sym {tm['fwd_sym_ms']/max(1,tm['fwd_data_launches']):.3f} adj {tm['adj_ms']/max(1,tm['adj_launches']):.3f} update {tm['update_ms']/max(1,tm['update_launches']):.3f} scalar {tm['scalar_ms']/51:.3f}")
