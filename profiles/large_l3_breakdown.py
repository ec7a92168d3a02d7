"""Kernel-class breakdown of the LSMR phase at the large-L3 shapes (BASELINE configs 3 and 5), where the tile adjoint
(L3P <= 16) does not apply.  usage: python profiles/large_l3_breakdown.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from helicon_b200.engine import Batch, Problem
from helicon_b200.grid import build_tasks
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec

for name, N, tw, rises in (("cfg5 384 rise 21", 384, -170.0, [21.0, 21.5]), ("cfg5 384 rise 45", 384, 34.0, [45.0, 45.5]),
                           ("cfg3 512 rise 4.8", 512, -1.3, [4.8, 4.85]), ("cfg3 512 rise 19", 512, -58.7, [19.0, 19.3])):
    img = bench.synthetic_filament(n=N, apix=1.3, diameter=0.3 * N * 1.3)
    tasks, _ = build_tasks(N, N, 1.3, np.array([tw]), np.array(rises), (1,), 3, None, 0.0, None, 0, -1, 0)
    g = tasks[0].geom
    tl = [t for t in tasks if t.geom["L3"] == g["L3"]]
    prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
    n3 = g["L3"] * prob.ndisk
    target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
    batch = Batch(prob, g["L3"], [CandidateSpec(t.twist, t.rise / g["apix3d"], 1, target, target, False) for t in tl])
    batch.solve(fixed_iters=2, check_every=2)  # warm-up (lazy kernel loading would land in the first timed launch)
    res = batch.solve(fixed_iters=10, check_every=10, profile=1)
    tm = batch.timing()
    nc = len(tl)
    per = lambda k: tm[k] / 11 / nc * 1e3
    print(f"{name}: L3={g['L3']} n={n3} views/cand={batch.plan.cands['view_count'].tolist()} MC={batch.plan.MC} | per candidate-pass (us): "
          f"fwd_data {per('fwd_data_ms'):.0f} fwd_sym {per('fwd_sym_ms'):.0f} adj {per('adj_ms'):.0f} update {per('update_ms'):.0f}", flush=True)
    batch.close(); prob.close()
