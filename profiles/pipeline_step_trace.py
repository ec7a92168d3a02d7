"""Per-step wall-clock trace of the double-buffered batch pipeline (six 200-candidate steps of cfg2): when each batch
becomes ready, how long its solve takes.  usage: python profiles/pipeline_step_trace.py"""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, warnings; warnings.filterwarnings("ignore")
import torch
import bench
from helicon_b200.engine import Batch, Problem
from helicon_b200.grid import BatchPipeline
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec
img = bench.synthetic_filament(); tasks = bench.grid_tasks(); g = tasks[0].geom
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
def specs(step): return [CandidateSpec(t.twist, t.rise / g["apix3d"], 1, target, target, False) for t in tasks[step*200:(step+1)*200]]
pipe = BatchPipeline(device=0, pipelined=True)
chunks=[specs(s) for s in range(6)]
torch.cuda.synchronize(); T0=time.perf_counter(); last=T0
for i,batch in pipe.run(prob, g["L3"], chunks):
    t1=time.perf_counter()
    res=batch.solve(profile=int(os.environ.get("PROF","0")))
    t2=time.perf_counter()
    tm=batch.timing()
    batch.close()
    t3=time.perf_counter()
    print(f"step {i}: wait-for-batch {1e3*(t1-last):.0f} ms | solve call {1e3*(t2-t1):.0f} ms (lsmr {tm['lsmr_ms']:.0f}) | close {1e3*(t3-t2):.0f} ms | itn max {res['itn'].max()}")
    last=t3
torch.cuda.synchronize(); print("total", time.perf_counter()-T0)
