"""Half-set (fsc_test) goldens from the UNMODIFIED reference (needs /root/reference; run in the build container):
lsq_reconstruct(fsc_test = 1..4) -> full / half-1 / half-2 reconstructions and the combined score
(solver_linear_regression.py:175-203 split_A_b, :441-482, :527-547).  Mode 1 shuffles with the global numpy RNG,
so the fixture stores the seed set right before the call.  Usage: python oracle/make_golden_fsc.py
TEST INFRASTRUCTURE ONLY."""
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
from make_golden import OUT, synth_image  # noqa: E402

# (name, N, apix, twist, rise_A, csym, positive_constraint, sym_oversample, L3, fsc_test, seed)
CASES = [
    ("fsc_mode1_48", 48, 5.4, -3.5, 9.5, 1, 0, 2, 6, 1, 123),
    ("fsc_mode2_48", 48, 5.4, -3.5, 9.5, 1, 0, 2, 6, 2, 0),
    ("fsc_mode3_48_c2", 48, 5.4, 27.0, 12.0, 2, 0, 2, 8, 3, 0),
    ("fsc_mode4_32", 32, 8.125, -1.2, 4.75, 1, 0, 4, 2, 4, 0),
]

for name, N, apix, twist, rise, csym, pc, so, L3, mode, seed in CASES:
    img = synth_image(N, apix, twist=twist, rise=rise, csym=csym)
    S.build_A_data_matrix.clear_cache()
    np.random.seed(seed)
    (rec, h1, h2), score = S.lsq_reconstruct(
        projection_image=img, scale2d_to_3d=1.0, twist_degree=twist, rise_pixel=rise / apix, csym=csym,
        positive_constraint=pc, reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
        reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=L3, sym_oversample=so, interpolation="nn",
        fsc_test=mode, algorithm=dict(model="lsq"), cpu=1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img,
                        args=np.array([apix, twist, rise, csym, pc, so, L3, mode, seed], dtype=np.float64),
                        rec3d=rec, half1=h1, half2=h2, score=np.float64(score))
    print(name, rec.shape, float(score))
