"""A/B of the side-stream overlap of the symmetry-row forward (HB2_OVERLAP=0/1, read at the first solve of the process):
a cfg2 batch of NC candidates, NI fixed LSMR iterations, profile=0 (no per-launch events) and profile=1; prints the
device time of the LSMR phase per candidate-iteration and a checksum of the scores (must not change).
usage: HB2_OVERLAP=0|1 python profiles/prof_overlap.py [NC] [NI]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from helicon_b200.engine import Batch, Problem
from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ni = int(sys.argv[2]) if len(sys.argv) > 2 else 60
img = bench.synthetic_filament()
tasks = bench.grid_tasks()
g = tasks[0].geom
prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
n3 = g["L3"] * prob.ndisk
target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
sel = tasks[20000:20000 + nc]
batch = Batch(prob, g["L3"], [CandidateSpec(t.twist, t.rise / g["apix3d"], 1, target, target, False) for t in sel])
for prof in (0, 0, 0, 1):
    res = batch.solve(fixed_iters=ni, check_every=ni, profile=prof)
    tm = batch.timing()
    print(f"HB2_OVERLAP={os.environ.get('HB2_OVERLAP', '1')} profile={prof}: lsmr {tm['lsmr_ms']:.2f} ms = "
          f"{1e3 * tm['lsmr_ms'] / (nc * ni):.2f} us per candidate-iteration; fwd_data {tm['fwd_data_ms']:.1f} fwd_sym "
          f"{tm['fwd_sym_ms']:.1f} adj {tm['adj_ms']:.1f} update {tm['update_ms']:.1f} norm {tm['norm_ms']:.1f}; "
          f"score checksum {float(np.sum(res['score'].astype(np.float64))):.9f}")
batch.close(); prob.close()
