"""Why do small batches cost more per candidate?  cfg2 batches of NC = 50 / 100 / 200 grid candidates (whole twist rows,
default positive rule) built and solved to convergence with per-class device timing; prints set-up wall time and the
device time per candidate, per phase.  usage: python profiles/small_batch_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from helicon_b200.engine import Batch, Problem
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec, positive_rule

img = bench.synthetic_filament()
tasks = bench.grid_tasks()
g = tasks[0].geom
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
sel = tasks[20000:20200]


def specs(ts):
    return [CandidateSpec(t.twist, t.rise / g["apix3d"], 1, target, target,
                          positive_rule(-1, t.rise / g["apix3d"], t.twist, g["L3"])) for t in ts]


for rep in range(2):  # the first round warms the allocator / module load
    for nc in (200, 100, 50):
        tot = dict(setup=0.0, wall=0.0, lsmr=0.0, trf=0.0, score=0.0, itn=0, fwd=0.0, adj=0.0, upd=0.0, sym=0.0, trfit=0)
        for i0 in range(0, 200, nc):
            t0 = time.perf_counter()
            b = Batch(prob, g["L3"], specs(sel[i0:i0 + nc]))
            t1 = time.perf_counter()
            res = b.solve(profile=1)
            t2 = time.perf_counter()
            tm = b.timing()
            b.close()
            tot["setup"] += t1 - t0; tot["wall"] += t2 - t1
            tot["lsmr"] += tm["lsmr_ms"]; tot["trf"] += tm["trf_ms"]; tot["score"] += tm["score_ms"]
            tot["fwd"] += tm["fwd_data_ms"]; tot["adj"] += tm["adj_ms"]; tot["upd"] += tm["update_ms"]; tot["sym"] += tm["fwd_sym_ms"]
            tot["itn"] += int(res["itn"].sum()); tot["trfit"] += int(res["trf_nit"].sum())
        if rep:
            it = tot["itn"]
            print(f"batches of {nc:3d}: set-up {1e3 * tot['setup'] / 200:.2f} ms/cand, solve wall {1e3 * tot['wall'] / 200:.2f} ms/cand "
                  f"(lsmr {tot['lsmr'] / 200:.2f} trf {tot['trf'] / 200:.2f} score {tot['score'] / 200:.3f}); per candidate-iteration: "
                  f"lsmr {1e3 * tot['lsmr'] / it:.2f} us = fwd {1e3 * tot['fwd'] / it:.2f} + sym {1e3 * tot['sym'] / it:.2f} + adj "
                  f"{1e3 * tot['adj'] / it:.2f} + update {1e3 * tot['upd'] / it:.2f} + rest; itn {it} trf iterations {tot['trfit']}", flush=True)
prob.close()
