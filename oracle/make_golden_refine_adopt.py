"""Golden that pins WHICH branch lsq_reconstruct(refine_tilt_psi_dy_range=...) takes in the UNMODIFIED reference
(SLR:372-437): for model "lsq" solve_equations returns score=None, so the refined x / score / _refined_params are always
adopted.  Same image as refine_dy_40.  Usage: python oracle/make_golden_refine_adopt.py.  TEST INFRASTRUCTURE ONLY."""
import os
import sys
import warnings

os.environ.setdefault("HELION_CACHE_DIR", "/tmp/helicon_cache_golden")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
from helicon.webApps.denovo3D import solver_linear_regression as S  # noqa: E402
from make_golden import OUT  # noqa: E402

d = np.load(os.path.join(OUT, "refine_dy_40.npz"))
apix, twist, rise, csym, L3, so, pc, mi = d["args"]
img = d["image"]
N = img.shape[0]
kw = dict(scale2d_to_3d=1.0, twist_degree=float(twist), rise_pixel=float(rise / apix), csym=int(csym),
          positive_constraint=int(pc), reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
          reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so), interpolation="nn",
          algorithm=dict(model="lsq"), cpu=1)
S.build_A_data_matrix.clear_cache()
S.build_A_helical_sym_matrix.clear_cache()
(rec0, _, _), s0 = S.lsq_reconstruct(projection_image=img, **kw)
if hasattr(S.lsq_reconstruct, "_refined_params"):
    del S.lsq_reconstruct._refined_params
rng = dict(tilt=5.0, psi=5.0, dy=2.0, max_iter=2)
(rec1, h1, h2), s1 = S.lsq_reconstruct(projection_image=img, refine_tilt_psi_dy_range=rng, **kw)
rp = getattr(S.lsq_reconstruct, "_refined_params", {})
np.savez_compressed(os.path.join(OUT, "refine_adopt_40.npz"), image=img, args=d["args"],
                    range=np.array([rng["tilt"], rng["psi"], rng["dy"], rng["max_iter"]], dtype=np.float64),
                    score_base=np.float64(s0), score=np.float64(s1), rec3d=rec1.astype(np.float32),
                    refined=np.array([rp.get("tilt", np.nan), rp.get("psi", np.nan), rp.get("dy", np.nan)], dtype=np.float64))
print("base", float(s0), "with refine", float(s1), "params", rp, "rel change", float(np.linalg.norm(rec1 - rec0) / np.linalg.norm(rec0)))
