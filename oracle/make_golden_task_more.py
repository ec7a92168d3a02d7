"""More pipeline.process_one_task goldens from the UNMODIFIED reference (needs /root/reference): trilinear interpolation
(the app's default mode) and fsc_test = 2 (even/odd half sets) through the whole task wrapper.
Usage: python oracle/make_golden_task_more.py.  TEST INFRASTRUCTURE ONLY."""
import os
import sys
import warnings

os.environ.setdefault("HELION_CACHE_DIR", "/tmp/helicon_cache_golden")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
from helicon.webApps.denovo3D import pipeline as RP  # noqa: E402
from helicon.webApps.denovo3D import utils as RU  # noqa: E402
from make_golden import OUT  # noqa: E402

TASKS = [
    # name, N, apix, twist, rise, csym, positive_constraint, thresh_fraction, seed, interpolation, fsc_test
    ("task_linear", 48, 5.0, -2.4, 9.6, 1, 0, -1, 11, "linear", 0),
    ("task_fsc2", 48, 5.0, -2.4, 9.6, 1, 0, -1, 11, "nn", 2),
]
for name, N, apix, twist, rise, csym, pc, tf, seed, interp, fsc in TASKS:
    np.random.seed(seed)
    img = RU.simulate_helical_projection(n=12, twist=twist, rise=rise, csym=csym, helical_diameter=0.5 * N * apix,
                                         ball_radius=1.5 * apix, polymer=1, planarity=0.9, ny=N, nx=N, apix=apix)
    img = np.ascontiguousarray(img, dtype=np.float32)
    kw = dict(ti=0, ntasks=1, data=img.copy(), imageFile="synthetic", imageIndex=1, twist=twist, rise=rise,
              rise_range=(rise, rise), csym=csym, tilt=0, tilt_range=(0, 0), psi=0, psi_range=0, dy=0, dy_range=0,
              apix2d_orig=apix, denoise="", low_pass=0, transpose=0, horizontalize=0, target_apix3d=0,
              target_apix2d=apix, thresh_fraction=tf, positive_constraint=pc, tube_length=-1, tube_diameter=N * apix,
              tube_diameter_inner=0, reconstruct_length=3 * rise, sym_oversample=-1, interpolation=interp, fsc_test=fsc,
              return_3d=True, score_metric="cosine", algorithm=dict(model="lsq"), verbose=0)
    score, rd, meta = RP.process_one_task(**kw)
    xp, yp, zs, (rec3d, h1, h2), D2, D3, L2, L3 = rd
    extra = {} if h1 is None else dict(half1=h1, half2=h2)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img,
                        args=np.array([apix, twist, rise, csym, pc, tf]), score=np.float32(score), x_proj=xp, y_proj=yp,
                        z_sections=zs, rec3d=rec3d, geom=np.array([D2, D3, L2, L3]), data_orig=meta[0],
                        interpolation=np.int64(interp == "linear"), fsc_test=np.int64(fsc), **extra)
    print(name, float(score), xp.shape, rec3d.shape, (D2, D3, L2, L3))
