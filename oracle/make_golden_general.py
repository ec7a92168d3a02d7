"""General-orientation (tilt/psi/dy != 0) and trilinear goldens from the UNMODIFIED reference (needs /root/reference;
run in the build container): build_A_data_matrix rows and full lsq_reconstruct solves.
Usage: python oracle/make_golden_general.py.  TEST INFRASTRUCTURE ONLY."""
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
from make_golden import OUT, csr_parts, synth_image  # noqa: E402

rng = np.random.default_rng(5)
# (name, N, s, twist, rise_px, csym, D2, L2, D3, D3i, L3, min_lines, interpolation, tilt, psi, dy)
DATA = [
    ("gen_data_nn_tilt", 24, 1.0, -7.3, 2.4, 1, 20, 22, 20, 0, 6, 10**7, "nn", 5.0, -3.0, 1.5),
    ("gen_data_nn_c2_stop", 24, 1.0, 33.0, 3.1, 2, 24, 24, 24, 0, 8, 1500, "nn", -8.0, 0.0, 0.0),
    ("gen_data_nn_s05_dy", 32, 0.5, -1.2, 1.9, 1, 32, 32, 16, 0, 4, 10**7, "nn", 0.0, 0.0, -0.75),
    ("gen_data_lin_tilt", 24, 1.0, -7.3, 2.4, 1, 20, 22, 20, 0, 6, 10**7, "linear", 5.0, -3.0, 1.5),
    ("gen_data_lin_psi_inner", 24, 1.0, 12.5, 2.2, 3, 24, 24, 24, 6, 6, 10**7, "linear", 0.0, 4.0, 0.0),
]
for name, N, s, twist, rise, csym, D2, L2, D3, D3i, L3, mpl, interp, tilt, psi, dy in DATA:
    img = rng.random((N, N)).astype(np.float32)
    A, b, pid = S.build_A_data_matrix.__wrapped__(
        image=img, scale2d_to_3d=s, twist_degree=twist, rise_pixel=rise, csym=csym, tilt_degree=tilt, psi_degree=psi,
        dy_pixel=dy, reconstruct_diameter_2d_pixel=D2, reconstruct_length_2d_pixel=L2, reconstruct_diameter_3d_pixel=D3,
        reconstruct_diameter_3d_inner_pixel=D3i, reconstruct_length_3d_pixel=L3, min_projection_lines=mpl,
        interpolation=interp, verbose=0, cpu=1)
    d = dict(image=img, args=np.array([s, twist, rise, csym, D2, L2, D3, D3i, L3, mpl, tilt, psi, dy], dtype=np.float64),
             b=b, b_pid=pid, linear=np.int64(interp == "linear"))
    d.update(csr_parts(A, "A"))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, A.shape, A.nnz)

# (name, N, apix, twist, rise_A, csym, positive_constraint, sym_oversample, L3, tilt, psi, dy)
SOLVE = [
    ("gen_solve_nn_tilt_48", 48, 5.4, -3.5, 9.5, 1, 0, 2, 6, 3.0, 2.0, 0.7),
    ("gen_solve_nn_tilt_48_pos", 48, 5.4, -3.5, 9.5, 1, 1, 2, 6, 3.0, 2.0, 0.7),
    ("gen_solve_nn_dy_32", 32, 8.125, -1.2, 4.75, 1, 0, 4, 2, 0.0, 0.0, 1.0),
]
for name, N, apix, twist, rise, csym, pc, so, L3, tilt, psi, dy in SOLVE:
    img = synth_image(N, apix, twist=twist, rise=rise, csym=csym)
    S.build_A_data_matrix.clear_cache()
    (rec, _, _), score = S.lsq_reconstruct(
        projection_image=img, scale2d_to_3d=1.0, twist_degree=twist, rise_pixel=rise / apix, csym=csym, tilt_degree=tilt,
        psi_degree=psi, dy_pixel=dy, positive_constraint=pc, reconstruct_diameter_2d_pixel=N,
        reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=L3,
        sym_oversample=so, interpolation="nn", algorithm=dict(model="lsq"), cpu=1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img,
                        args=np.array([apix, twist, rise, csym, pc, so, L3, tilt, psi, dy], dtype=np.float64), rec3d=rec,
                        score=np.float64(score))
    print(name, rec.shape, float(score))

# trilinear full solves: (name, N, apix, twist, rise_A, csym, positive_constraint, sym_oversample, L3, tilt, psi, dy)
SOLVE_LIN = [
    ("gen_solve_lin_48", 48, 5.4, -3.5, 9.5, 1, 0, 2, 8, 0.0, 0.0, 0.0),
    ("gen_solve_lin_48_pos", 48, 5.4, -3.5, 9.5, 1, 1, 2, 8, 0.0, 0.0, 0.0),
    ("gen_solve_lin_40_c2_tilt", 40, 6.5, 27.0, 12.0, 2, 0, 2, 10, 2.0, -1.0, 0.5),
]
for name, N, apix, twist, rise, csym, pc, so, L3, tilt, psi, dy in SOLVE_LIN:
    img = synth_image(N, apix, twist=twist, rise=rise, csym=csym)
    S.build_A_data_matrix.clear_cache()
    S.build_A_helical_sym_matrix.clear_cache()
    (rec, _, _), score = S.lsq_reconstruct(
        projection_image=img, scale2d_to_3d=1.0, twist_degree=twist, rise_pixel=rise / apix, csym=csym, tilt_degree=tilt,
        psi_degree=psi, dy_pixel=dy, positive_constraint=pc, reconstruct_diameter_2d_pixel=N,
        reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=L3,
        sym_oversample=so, interpolation="linear", algorithm=dict(model="lsq"), cpu=1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img,
                        args=np.array([apix, twist, rise, csym, pc, so, L3, tilt, psi, dy], dtype=np.float64), rec3d=rec,
                        score=np.float64(score))
    print(name, rec.shape, float(score))

# trilinear symmetry rows: (name, nz, ny, nx, twist, rise_px, csym, rmin, rmax, min_pairs)
HSYM_LIN = [
    ("gen_hsym_lin_c2_stop", 10, 24, 24, 41.0, 1.43, 2, 0, 11, 2500),
    ("gen_hsym_lin_inner", 12, 20, 20, -7.3, 2.37, 1, 3, 9, 10**7),
    ("gen_hsym_lin_tie", 10, 20, 20, 30.0, 2.5, 1, 0, 9, 10**7),
]
for name, nz, ny, nx, twist, rise, csym, rmin, rmax, msp in HSYM_LIN:
    A, b = S.build_A_helical_sym_matrix.__wrapped__(nz, ny, nx, twist, rise, csym, rmin, rmax, msp, "linear", verbose=0)
    d = dict(args=np.array([nz, ny, nx, twist, rise, csym, rmin, rmax, msp], dtype=np.float64))
    d.update(csr_parts(A, "A"))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, A.shape, A.nnz)
