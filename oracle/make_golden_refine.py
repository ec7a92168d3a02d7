"""refine_tilt_psi_dy goldens from the UNMODIFIED reference (needs /root/reference): a synthetic filament projected with
a known in-plane shift/tilt, refined from (0, 0, 0).  Usage: python oracle/make_golden_refine.py.  TEST INFRASTRUCTURE ONLY."""
import os
import sys
import warnings

os.environ.setdefault("HELION_CACHE_DIR", "/tmp/helicon_cache_golden")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
from scipy.ndimage import shift as nd_shift  # noqa: E402
from helicon.webApps.denovo3D import solver_linear_regression as S  # noqa: E402
from make_golden import OUT, synth_image  # noqa: E402

# (name, N, apix, twist, rise_A, csym, L3, sym_oversample, positive_constraint, image shift in y (px), max_iter)
CASES = [
    ("refine_dy_40", 40, 6.5, -3.5, 9.5, 1, 6, 2, 0, 1.3, 3),
    ("refine_dy_32_pos", 32, 8.125, 27.0, 12.0, 2, 6, 2, 1, -0.8, 2),
]
for name, N, apix, twist, rise, csym, L3, so, pc, sy, mi in CASES:
    img = synth_image(N, apix, twist=twist, rise=rise, csym=csym)
    img = np.ascontiguousarray(nd_shift(img, (sy, 0.0), order=1), dtype=np.float32)
    S.build_A_data_matrix.clear_cache()
    S.build_A_helical_sym_matrix.clear_cache()
    out = S.refine_tilt_psi_dy(
        projection_image=img, scale2d_to_3d=1.0, twist_degree=twist, rise_pixel=rise / apix, csym=csym,
        reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N,
        reconstruct_diameter_3d_inner_pixel=0, reconstruct_length_3d_pixel=L3, sym_oversample=so, interpolation="nn",
        x_init=None, max_iter=mi, positive_constraint=pc, verbose=0)
    tilt, psi, dy, x, score = out
    np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img,
                        args=np.array([apix, twist, rise, csym, L3, so, pc, mi], dtype=np.float64),
                        out=np.array([tilt, psi, dy, score], dtype=np.float64), x=np.asarray(x, dtype=np.float64))
    print(name, tilt, psi, dy, score)
