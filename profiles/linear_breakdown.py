"""Where the time of one trilinear candidate goes (cfg2 shape): row build vs solve.  usage: python profiles/linear_breakdown.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from helicon_b200.engine import ExplicitBatch, Problem
from helicon_b200.grid import build_tasks
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
img = bench.synthetic_filament(n=N, apix=1.3)
tasks, _ = build_tasks(N, N, 1.3, np.array([-1.23]), np.array([4.7]), (1,), 3, None, 0.0, None, 0, -1, 0)
g = tasks[0].geom
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], g["D3i"] / 2, g["D3"] // 2 - 1)
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
spec = CandidateSpec(-1.23, 4.7 / 1.3, 1, target, target, False)
for rep in range(2):
    t0 = time.perf_counter()
    b = ExplicitBatch(prob, g["L3"], spec, interpolation="linear")
    t1 = time.perf_counter()
    res = b.solve(profile=1)
    t2 = time.perf_counter()
    tm = b.timing()
    print(f"rep {rep}: build {t1-t0:.3f} s (rows {b.m_rows} nnz {b.nnz} sym rows {b.m_sym_explicit}), solve {t2-t1:.3f} s, itn {res[0]['itn']}, "
          f"lsmr {tm['lsmr_ms']:.0f} ms: fwd_data {tm['fwd_data_ms']:.0f} fwd_sym {tm['fwd_sym_ms']:.0f} adj {tm['adj_ms']:.0f} update {tm['update_ms']:.0f}", flush=True)
    b.close()
