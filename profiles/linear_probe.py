"""Trilinear (explicit-row) path at BASELINE config sizes: a few candidates of cfg1 (200x200) and cfg2 (256x256) through
search_grid(interpolation="linear"): candidates/s, iterations, rows / entries.  usage: python profiles/linear_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from helicon_b200.grid import search_grid

for name, N, twists, rises in (("cfg1 200x200", 200, np.array([-1.23, -1.19]), np.array([4.7, 4.8])),
                               ("cfg2 256x256", 256, np.array([-1.23, -1.19]), np.array([4.7, 4.8]))):
    img = bench.synthetic_filament(n=N, apix=1.3)
    for interp in ("nn", "linear"):
        search_grid(img, 1.3, twists[:1], rises[:1], positive_constraint=0, interpolation=interp)  # warm-up
        t0 = time.perf_counter()
        out = search_grid(img, 1.3, twists, rises, positive_constraint=0, interpolation=interp)
        dt = time.perf_counter() - t0
        print(f"{name} {interp}: {out['n_candidates']} candidates in {dt:.2f} s = {out['n_candidates']/dt:.2f} cand/s; "
              f"iterations {out['itn'].ravel().tolist()}; scores {np.round(out['scores'].ravel(), 5).tolist()}", flush=True)
