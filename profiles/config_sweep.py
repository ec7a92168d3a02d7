"""Throughput of small batches at the shapes of BASELINE configs 1, 3 and 5 (parity-test configurations, not bench
lines): full solve + score through search_grid().  usage: python profiles/config_sweep.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from helicon_b200.grid import search_grid

def run(name, N, apix, twists, rises, csyms, **kw):
    img = bench.synthetic_filament(n=N, apix=apix, diameter=0.3 * N * apix)
    search_grid(img, apix, twists[:1], rises[:2], csyms=csyms[:1], **kw)  # warm-up
    t0 = time.perf_counter()
    out = search_grid(img, apix, twists, rises, csyms=csyms, **kw)
    dt = time.perf_counter() - t0
    ok = np.isfinite(out["scores"])
    print(f"{name}: {out['n_candidates']} candidates in {dt:.2f} s = {out['n_candidates']/dt:.1f} cand/s; "
          f"iterations mean {out['itn'][ok].mean():.0f} max {out['itn'][ok].max()}; best score {np.nanmax(out['scores']):.5f} at "
          f"{out['top'][0]['twist']:.3f} deg / {out['top'][0]['rise']:.3f} A csym {out['top'][0]['csym']}; "
          f"flags {np.unique(out['flags'][ok])}", flush=True)

run("cfg1 200x200 (100 twists x 10 rises)", 200, 1.3, np.linspace(-2.19, -0.21, 100), np.linspace(4.5, 4.95, 10), (1,), positive_constraint=0)
run("cfg3 512x512 csym 1-3 (4 twists x 4 rises)", 512, 1.3, np.linspace(-3.0, -0.5, 4), np.linspace(4.5, 5.0, 4), (1, 2, 3), positive_constraint=0)
run("cfg5 384x384 rise 20-60 (6 twists x 4 rises)", 384, 1.3, np.linspace(-170.0, 170.0, 6), np.linspace(21.0, 58.0, 4), (1,), positive_constraint=0)
