"""Matrix-free trilinear batch at the cfg2 shape (256x256, L3 = 12): NC grid candidates of one twist row,
positive_constraint = 0.  usage: python profiles/prof_trilinear.py [NC] [fixed iterations, 0 = to convergence] [positive]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from helicon_b200.bilinear import BilinearBatch
from helicon_b200.engine import Problem
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ni = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pos = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
img = bench.synthetic_filament()
tasks = bench.grid_tasks()
g = tasks[0].geom
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1, interpolation="linear")
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
sel = tasks[20000:20000 + nc]
for rep in range(2):
    t0 = time.perf_counter()
    batch = BilinearBatch(prob, g["L3"], [CandidateSpec(t.twist, t.rise / g["apix3d"], 1, target, target, pos) for t in sel])
    t1 = time.perf_counter()
    kw = dict(fixed_iters=ni, check_every=ni) if ni else {}
    res = batch.solve(profile=1, **kw)
    t2 = time.perf_counter()
    tm = batch.timing()
    its = float(res["itn"].sum())
    print(f"rep {rep}: nc={nc} maps={len(batch.maps)} (regular {batch.n_regular_maps}) views/cand={batch.cand_nview.mean():.1f} "
          f"sym rows/cand={batch.m_sym.mean():.0f} data rows/cand={batch.plan.cand_n_data_rows.mean():.0f}")
    print(f"   setup {1e3 * (t1 - t0) / nc:.2f} ms/cand, solve {1e3 * (t2 - t1) / nc:.2f} ms/cand -> {nc / (t2 - t0):.2f} cand/s; "
          f"mean itn {its / nc:.1f}, trf {res['trf_nit'].mean():.1f}, score[:3] {res['score'][:3]}")
    print("   per candidate-iteration (us): " + " ".join(
        f"{k}={1e3 * tm[k] / max(its, 1):.2f}" for k in ("fwd_data_ms", "fwd_sym_ms", "adj_ms", "update_ms", "norm_ms", "scalar_ms", "lsmr_ms")) +
        f" | trf_ms={tm['trf_ms']:.1f} score_ms={tm['score_ms']:.2f}")
    batch.close()
prob.close()
