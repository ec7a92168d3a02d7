"""fsc_test goldens on the explicit-row paths from the UNMODIFIED reference (needs /root/reference): half-set solves with
trilinear interpolation and with a tilted candidate.  Usage: python oracle/make_golden_fsc_explicit.py.
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

# (name, N, apix, twist, rise_A, csym, sym_oversample, L3, interpolation, tilt, psi, dy, fsc_test, seed)
CASES = [
    ("fsc_lin_mode2_40", 40, 6.5, 27.0, 12.0, 2, 2, 10, "linear", 0.0, 0.0, 0.0, 2, 0),
    ("fsc_tilt_mode3_48", 48, 5.4, -3.5, 9.5, 1, 2, 6, "nn", 3.0, 2.0, 0.7, 3, 0),
    ("fsc_tilt_mode1_48", 48, 5.4, -3.5, 9.5, 1, 2, 6, "nn", 3.0, 2.0, 0.7, 1, 77),
]
for name, N, apix, twist, rise, csym, so, L3, interp, tilt, psi, dy, mode, seed in CASES:
    img = synth_image(N, apix, twist=twist, rise=rise, csym=csym)
    S.build_A_data_matrix.clear_cache()
    S.build_A_helical_sym_matrix.clear_cache()
    np.random.seed(seed)
    (rec, h1, h2), score = S.lsq_reconstruct(
        projection_image=img, scale2d_to_3d=1.0, twist_degree=twist, rise_pixel=rise / apix, csym=csym, tilt_degree=tilt,
        psi_degree=psi, dy_pixel=dy, positive_constraint=0, reconstruct_diameter_2d_pixel=N,
        reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=L3, sym_oversample=so,
        interpolation=interp, fsc_test=mode, algorithm=dict(model="lsq"), cpu=1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img,
                        args=np.array([apix, twist, rise, csym, so, L3, tilt, psi, dy, mode, seed], dtype=np.float64),
                        linear=np.int64(interp == "linear"), rec3d=rec, half1=h1, half2=h2, score=np.float64(score))
    print(name, rec.shape, float(score))
