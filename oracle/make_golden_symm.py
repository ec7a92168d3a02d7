"""Golden fixtures for the post-solve products: runs the UNMODIFIED reference
(helicon.apply_helical_symmetry, lib/transforms.py:58-165) in the build container.
Usage: python oracle/make_golden_symm.py   ->  tests/golden/symm_*.npz"""
import os
import sys
import warnings

os.environ.setdefault("HELION_CACHE_DIR", "/tmp/helicon_cache_golden")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
sys.path.insert(0, "/root/reference/src")
warnings.filterwarnings("ignore")
import numpy as np  # noqa: E402
import helicon  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# name, asym-unit shape, apix, twist, rise, csym, new_size, new_apix, seed
CASES = [
    ("symm_a", (6, 20, 20), 2.0, -7.3, 4.8, 1, (40, 24, 24), 2.0, 1),
    ("symm_c3_rescale", (8, 24, 24), 5.0, 27.5, 11.0, 3, (50, 30, 30), 2.5, 2),
    ("symm_same_size", (16, 16, 16), 1.0, 45.0, 3.0, 1, (16, 16, 16), 1.0, 3),
    ("symm_odd", (6, 18, 18), 1.3, -1.2, 4.75, 1, (33, 21, 21), 1.3, 4),
]
for name, shp, apix, twist, rise, csym, new_size, new_apix, seed in CASES:
    rng = np.random.default_rng(seed)
    data = rng.random(shp).astype(np.float32)
    zz, yy, xx = np.mgrid[0:shp[0], 0:shp[1], 0:shp[2]]
    data *= (((yy - shp[1] // 2) ** 2 + (xx - shp[2] // 2) ** 2) < (shp[1] // 2 - 1) ** 2).astype(np.float32)
    data[0] *= 0.001  # a slice below the 1 % profile threshold
    out = helicon.apply_helical_symmetry(data=data, apix=apix, twist_degree=twist, rise_angstrom=rise, csym=csym,
                                         new_size=new_size, new_apix=new_apix, cpu=1)
    xp = np.sum(out, axis=2).T
    yp = np.sum(out, axis=1).T
    np.savez_compressed(os.path.join(OUT, name + ".npz"), data=data, args=np.array([apix, twist, rise, csym, new_apix]),
                        new_size=np.array(new_size), out=np.asarray(out, dtype=np.float32), x_proj=xp, y_proj=yp)
    print(name, out.shape, float(np.abs(out).max()))

# ---- pipeline.process_one_task (pipeline.py:85-497) on small synthetic filaments -------------------------------------
from helicon.webApps.denovo3D import pipeline as RP  # noqa: E402
from helicon.webApps.denovo3D import utils as RU  # noqa: E402

TASKS = [
    # name, N, apix, twist, rise, csym, positive_constraint, thresh_fraction, seed
    ("task_a", 48, 5.0, -2.4, 9.6, 1, 0, -1, 11),
    ("task_c2_thresh", 40, 4.0, 31.0, 8.3, 2, 0, 0.05, 12),
    ("task_tiez", 40, 4.0, 31.0, 8.2, 2, 0, 0.05, 12),  # h*rise_px half-integer: a rounding tie the reference resolves by coordinate noise
]
for name, N, apix, twist, rise, csym, pc, tf, seed in TASKS:
    np.random.seed(seed)
    img = RU.simulate_helical_projection(n=12, twist=twist, rise=rise, csym=csym, helical_diameter=0.5 * N * apix,
                                         ball_radius=1.5 * apix, polymer=1, planarity=0.9, ny=N, nx=N, apix=apix)
    img = np.ascontiguousarray(img, dtype=np.float32)
    kw = dict(ti=0, ntasks=1, data=img.copy(), imageFile="synthetic", imageIndex=1, twist=twist, rise=rise,
              rise_range=(rise, rise), csym=csym, tilt=0, tilt_range=(0, 0), psi=0, psi_range=0, dy=0, dy_range=0,
              apix2d_orig=apix, denoise="", low_pass=0, transpose=0, horizontalize=0, target_apix3d=0,
              target_apix2d=apix, thresh_fraction=tf, positive_constraint=pc, tube_length=-1, tube_diameter=N * apix,
              tube_diameter_inner=0, reconstruct_length=3 * rise, sym_oversample=-1, interpolation="nn", fsc_test=0,
              return_3d=True, score_metric="cosine", algorithm=dict(model="lsq"), verbose=0)
    score, rd, meta = RP.process_one_task(**kw)
    xp, yp, zs, (rec3d, _, _), D2, D3, L2, L3 = rd
    np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img,
                        args=np.array([apix, twist, rise, csym, pc, tf]), score=np.float32(score), x_proj=xp, y_proj=yp,
                        z_sections=zs, rec3d=rec3d, geom=np.array([D2, D3, L2, L3]), data_orig=meta[0])
    print(name, float(score), xp.shape, yp.shape, zs.shape, rec3d.shape, (D2, D3, L2, L3))
