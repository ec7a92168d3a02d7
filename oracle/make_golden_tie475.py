"""Golden solve at the headline's TRUE parameters (rise 4.75 A at 1.3 A/px -> rise_pixel = 95/26, so h = 13 gives the
half-integer z shift 47.5: a column->slice tie view) on a 96x96 synthetic filament, run through the
UNMODIFIED reference.  Usage: python oracle/make_golden_tie475.py -> tests/golden/solve_nn_unb_96_tie475.npz"""
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
from make_golden import synth_image, OUT  # noqa: E402

name, N, apix, twist, rise, csym, pc, so, L3 = "solve_nn_unb_96_tie475", 96, 1.3, -1.2, 4.75, 1, 0, 2, 12
img = synth_image(N, apix, twist=twist, rise=rise, csym=csym, diameter=80.0)
S.build_A_data_matrix.clear_cache()
(rec3d, _, _), score = S.lsq_reconstruct(
    projection_image=img, scale2d_to_3d=1.0, twist_degree=twist, rise_pixel=rise / apix, csym=csym,
    positive_constraint=pc, reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
    reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=L3, sym_oversample=so, interpolation="nn",
    algorithm=dict(model="lsq"), cpu=1)
np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img,
                    args=np.array([apix, twist, rise, csym, pc, so, L3], dtype=np.float64), rec3d=rec3d,
                    score=np.float64(score))
print(name, rec3d.shape, float(score))
