"""Golden grid search: the UNMODIFIED reference's lsq_reconstruct over a small (twist x rise) grid on a 64x64 synthetic
filament -> scores of every candidate (tests/golden/grid_64.npz).  The GPU grid driver must reproduce the scores to
1e-5, the best (twist, rise) and the top-K ordering (BASELINE.json north_star).
Usage: python oracle/make_golden_grid.py"""
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

N, apix = 64, 4.0625
true_twist, true_rise = -1.37, 4.71
img = synth_image(N, apix, twist=true_twist, rise=true_rise, csym=1, seed=5)
twists = np.array([-2.13, -1.71, -1.37, -1.03, -0.67])
rises = np.array([4.31, 4.71, 5.13, 5.57])
L3, so = 4, 10
scores = np.zeros((len(twists), len(rises)), dtype=np.float32)
for a, tw in enumerate(twists):
    for b, ri in enumerate(rises):
        S.build_A_data_matrix.clear_cache()
        (_, _, _), sc = S.lsq_reconstruct(
            projection_image=img, scale2d_to_3d=1.0, twist_degree=float(tw), rise_pixel=float(ri / apix), csym=1,
            positive_constraint=0, reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
            reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=L3, sym_oversample=so, interpolation="nn",
            algorithm=dict(model="lsq"), cpu=1)
        scores[a, b] = sc
        print(tw, ri, float(sc), flush=True)
np.savez_compressed(os.path.join(OUT, "grid_64.npz"), image=img, apix=apix, twists=twists, rises=rises,
                    L3=L3, sym_oversample=so, scores=scores)
print("best", np.unravel_index(np.argmax(scores), scores.shape))
