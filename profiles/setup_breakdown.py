"""Host/GPU setup breakdown of one cfg2 batch: where the non-LSMR time of a step goes.
usage: python profiles/setup_breakdown.py [NC] [POSITIVE]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import bench
from helicon_b200 import _lib
from helicon_b200.engine import Batch, Problem
from helicon_b200.planner import MAX_EQUATIONS, BatchPlan, CandidateSpec, positive_rule

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 200
pos = int(sys.argv[2]) if len(sys.argv) > 2 else 0
img = bench.synthetic_filament()
tasks = bench.grid_tasks()
g = tasks[0].geom
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
for rep, start in enumerate((20000, 20000 + nc, 31000)):
    sel = tasks[start:start + nc]
    specs = [CandidateSpec(t.twist, t.rise / g["apix3d"], 1, target, target,
                           positive_rule(pos, t.rise / g["apix3d"], t.twist, g["L3"])) for t in sel]
    t0 = time.perf_counter()
    plan = BatchPlan(prob.s, prob.D2, prob.L2, g["L3"], specs)
    t1 = time.perf_counter()
    lib = _lib.load()
    tt = {}
    for nm in ("hb2_batch_begin", "hb2_batch_create"):
        def wrap(f, nm=nm):
            def w(*a):
                q0 = time.perf_counter(); r = f(*a); tt[nm] = time.perf_counter() - q0; return r
            return w
        setattr(lib, nm + "_orig", getattr(lib, nm + "_orig", None) or getattr(lib, nm))
    class L:  # proxy timing the two setup calls
        def __getattr__(self, k):
            f = getattr(lib, k)
            if k in ("hb2_batch_begin", "hb2_batch_create"):
                def w(*a):
                    q0 = time.perf_counter(); r = f(*a); tt[k] = time.perf_counter() - q0; return r
                return w
            return f
    _orig = _lib.require_gpu
    _lib.require_gpu = lambda: L()
    import helicon_b200.engine as E
    batch = Batch(prob, g["L3"], specs)
    _lib.require_gpu = _orig
    t2 = time.perf_counter()
    print("   " + " ".join(f"{k} {1e3*v:.0f} ms" for k, v in tt.items()))
    res = batch.solve()
    t3 = time.perf_counter()
    tm = batch.timing()
    print(f"rep {rep}: plan-stage1 {1e3*(t1-t0):.0f} ms | Batch() total {1e3*(t2-t1):.0f} ms | solve() {1e3*(t3-t2):.0f} ms "
          f"(lsmr {tm['lsmr_ms']:.0f} trf {tm['trf_ms']:.0f} score {tm['score_ms']:.1f}) | angles {len(batch.plan.angles)} "
          f"views {len(batch.plan.views)} pairs {len(batch.plan.pairs)}")
    itn = res["itn"]
    print("   itn min/mean/max", itn.min(), itn.mean(), itn.max(), " per-twist-group max:",
          [int(itn[i:i + 50].max()) for i in range(0, nc, 50)], "trf_nit mean", res["trf_nit"].mean())
    print("   sym rows mean", res["n_sym_rows"].mean(), "flags", np.unique(res["flags"], return_counts=True))
    batch.close()
prob.close()
