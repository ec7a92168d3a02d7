"""Goldens for the tilted / refined task wrapper and for helicon.transform_map from the UNMODIFIED reference (needs
/root/reference): process_one_task with tilt/psi/dy != 0 (explicit-row solve + transform_map of the display volume,
pipeline.py:428-447), process_one_task with a dy refinement range, and transform_map on a small random volume.
Usage: python oracle/make_golden_task_tilt.py.  TEST INFRASTRUCTURE ONLY."""
import os
import sys
import warnings

os.environ.setdefault("HELION_CACHE_DIR", "/tmp/helicon_cache_golden")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import helicon  # noqa: E402
from helicon.webApps.denovo3D import pipeline as RP  # noqa: E402
from helicon.webApps.denovo3D import solver_linear_regression as S  # noqa: E402
from helicon.webApps.denovo3D import utils as RU  # noqa: E402
from make_golden import OUT  # noqa: E402

rng = np.random.default_rng(5)
vol = rng.random((12, 10, 8)).astype(np.float32)
tm = {}
for i, kw in enumerate([dict(tilt=12.5, psi=-7.0, dy=1.3), dict(rot=30.0, tilt=-4.0, dx=0.7, dz=-1.1),
                        dict(scale=0.9, psi=15.0), dict()]):
    tm[f"case{i}_args"] = np.array([kw.get(k, d) for k, d in (("scale", 1.0), ("rot", 0), ("tilt", 0), ("psi", 0),
                                                            ("dx", 0), ("dy", 0), ("dz", 0))], dtype=np.float64)
    tm[f"case{i}"] = helicon.transform_map(vol, **kw)
np.savez_compressed(os.path.join(OUT, "transform_map.npz"), vol=vol, **tm)

TASKS = [
    # name, N, apix, twist, rise, csym, pc, seed, tilt, psi, dy (A), tilt_range, psi_range, dy_range
    ("task_tilt", 40, 6.0, -3.1, 9.6, 1, 0, 11, 3.0, -1.5, 4.0, (0, 0), 0, 0),
    ("task_refine_dy", 40, 6.0, -3.1, 9.6, 1, 0, 11, 0.0, 0.0, 0.0, (0, 0), 0, 1.5),
]
for name, N, apix, twist, rise, csym, pc, seed, tilt, psi, dy, trng, prng, drng in TASKS:
    np.random.seed(seed)
    img = RU.simulate_helical_projection(n=12, twist=twist, rise=rise, csym=csym, helical_diameter=0.5 * N * apix,
                                         ball_radius=1.5 * apix, polymer=1, planarity=0.9, ny=N, nx=N, apix=apix)
    img = np.ascontiguousarray(img, dtype=np.float32)
    if hasattr(S.lsq_reconstruct, "_refined_params"):
        del S.lsq_reconstruct._refined_params
    S.build_A_data_matrix.clear_cache()
    S.build_A_helical_sym_matrix.clear_cache()
    kw = dict(ti=0, ntasks=1, data=img.copy(), imageFile="synthetic", imageIndex=1, twist=twist, rise=rise,
              rise_range=(rise, rise), csym=csym, tilt=tilt, tilt_range=trng, psi=psi, psi_range=prng, dy=dy, dy_range=drng,
              apix2d_orig=apix, denoise="", low_pass=0, transpose=0, horizontalize=0, target_apix3d=0,
              target_apix2d=apix, thresh_fraction=-1, positive_constraint=pc, tube_length=-1, tube_diameter=N * apix,
              tube_diameter_inner=0, reconstruct_length=3 * rise, sym_oversample=-1, interpolation="nn", fsc_test=0,
              return_3d=True, score_metric="cosine", algorithm=dict(model="lsq"), verbose=0)
    score, rd, meta = RP.process_one_task(**kw)
    xp, yp, zs, (rec3d, h1, h2), D2, D3, L2, L3 = rd
    np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img,
                        args=np.array([apix, twist, rise, csym, pc, -1, tilt, psi, dy, trng[0], trng[1], prng, drng]),
                        score=np.float32(score), x_proj=xp, y_proj=yp, z_sections=zs, rec3d=rec3d,
                        geom=np.array([D2, D3, L2, L3]))
    print(name, float(score), xp.shape, rec3d.shape, (D2, D3, L2, L3))
