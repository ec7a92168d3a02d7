"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build
container (needs /root/reference; it does not travel to the GPU box, the
fixtures do).  Usage:  python oracle/make_golden.py

Every fixture stores the inputs and the reference's outputs for one call of a
hot-path function so that the oracle restatement
(oracle/denovo3d_oracle.py) and the CUDA path can both be pinned to them.
"""

import os
import sys
import json
import warnings

os.environ.setdefault("HELION_CACHE_DIR", "/tmp/helicon_cache_golden")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
sys.path.insert(0, "/root/reference/src")
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import helicon  # noqa: E402
from helicon.webApps.denovo3D import solver_linear_regression as S  # noqa: E402
from helicon.webApps.denovo3D import utils as U  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def csr_parts(A, prefix):
    A = A.tocsr().copy()
    A.sum_duplicates()
    A.sort_indices()
    return {
        prefix + "_indptr": A.indptr.astype(np.int64),
        prefix + "_indices": A.indices.astype(np.int32),
        prefix + "_data": A.data.astype(np.float32),
        prefix + "_shape": np.array(A.shape, dtype=np.int64),
    }


def synth_image(N, apix, seed=7, twist=-1.2, rise=4.75, csym=1, n=30, diameter=100.0, ball=3.0, polymer=1, planarity=0.9):
    np.random.seed(seed)
    img = U.simulate_helical_projection(
        n=n, twist=twist, rise=rise, csym=csym, helical_diameter=diameter, ball_radius=ball,
        polymer=polymer, planarity=planarity, ny=N, nx=N, apix=apix,
    )
    return np.ascontiguousarray(img, dtype=np.float32)


# (name, N, s, twist, rise_px, csym, D2, L2, D3, D3i, L3, min_lines, interpolation)
DATA_CASES = [
    ("data_nn_a", 24, 1.0, -7.3, 2.4, 1, 20, 22, 20, 0, 6, 10**7, "nn"),
    ("data_nn_csym2", 24, 1.0, 33.0, 3.1, 2, 24, 24, 24, 0, 8, 10**7, "nn"),
    ("data_nn_s05", 32, 0.5, -1.2, 1.9, 1, 32, 32, 16, 0, 4, 10**7, "nn"),
    ("data_nn_earlystop", 24, 1.0, -7.3, 2.4, 1, 24, 24, 24, 0, 6, 900, "nn"),
    ("data_nn_inner", 24, 1.0, 12.5, 2.2, 3, 24, 24, 24, 6, 6, 10**7, "nn"),
    ("data_nn_tie30", 16, 1.0, 30.0, 2.0, 1, 12, 12, 12, 0, 8, 10**7, "nn"),
    ("data_nn_tiez", 24, 1.0, -3.7, 3.5, 1, 24, 24, 24, 0, 8, 10**7, "nn"),
    ("data_nn_reftest", 8, 1.0, 30.0, 2.0, 1, 4, 4, 4, 0, 4, 10, "nn"),
    ("data_lin_a", 24, 1.0, -7.3, 2.4, 1, 20, 22, 20, 0, 6, 10**7, "linear"),
    ("data_lin_csym2", 24, 1.0, 33.0, 3.1, 2, 24, 24, 24, 0, 8, 10**7, "linear"),
]

# (name, nz, ny, nx, twist, rise_px, csym, rmin, rmax, min_pairs, interpolation)
HSYM_CASES = [
    ("hsym_nn_a", 8, 16, 16, -7.3, 2.37, 1, 0, 7, 10**7, "nn"),
    ("hsym_nn_csym2_stop", 6, 20, 20, 41.0, 1.43, 2, 0, 9, 3000, "nn"),
    ("hsym_nn_inner", 8, 16, 16, -1.2, 0.95, 1, 3, 7, 10**7, "nn"),
    ("hsym_nn_tie", 8, 16, 16, 30.0, 2.5, 1, 0, 7, 10**7, "nn"),
    ("hsym_nn_reftest", 8, 8, 8, 30.0, 2.0, 1, 0, 3, 10, "nn"),
    ("hsym_lin_a", 10, 24, 24, -17.3, 3.37, 1, 0, 11, 10**7, "linear"),
]

# (name, N, apix, twist, rise_A, csym, positive_constraint, sym_oversample, interpolation, L3)
SOLVE_CASES = [
    ("solve_nn_unb_32", 32, 8.125, -1.2, 4.75, 1, 0, 4, "nn", 2),
    ("solve_nn_pos_32", 32, 8.125, -1.2, 4.75, 1, -1, 4, "nn", 2),
    ("solve_nn_unb_48_t35", 48, 5.4, -3.5, 9.5, 1, 0, 2, "nn", 6),
    ("solve_nn_pos_48_t35", 48, 5.4, -3.5, 9.5, 1, 1, 2, "nn", 6),
    ("solve_nn_unb_48_c2", 48, 5.4, 27.0, 12.0, 2, 0, 2, "nn", 8),
    ("solve_nn_unb_64", 64, 4.0625, -1.2, 4.75, 1, 0, 10, "nn", 4),
]


def main():
    rng = np.random.default_rng(0)
    manifest = {}

    # helpers ---------------------------------------------------------------
    from scipy.stats import qmc

    hal = {}
    for n in [1, 2, 3, 5, 10, 15, 59, 73, 147, 435, 1000]:
        hal[str(n)] = qmc.Halton(d=1, scramble=False).integers(l_bounds=0, u_bounds=n, n=n).ravel().astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "halton.npz"), **hal)
    pairs = {}
    for i, (tw, ri, cs, nz) in enumerate([(30, 5, 1, 20), (30, 5, 2, 20), (-1.2, 3.654, 1, 12), (41.0, 1.43, 3, 6), (179.5, 7.1, 4, 30)]):
        res = S.sorted_hsym_csym_pairs(tw, ri, cs, nz)
        arr = np.array([[e[0], e[1], e[2], e[3], e[4], e[5][0][0], e[5][0][1], e[5][1][0], e[5][1][1]] for e in res], dtype=np.float64)
        pairs[f"case{i}_args"] = np.array([tw, ri, cs, nz], dtype=np.float64)
        pairs[f"case{i}"] = arr
    np.savez_compressed(os.path.join(OUT, "hsym_pairs.npz"), **pairs)
    masks = {}
    for i, (nz, ny, nx, rmin, rmax) in enumerate([(2, 8, 8, 0, 3), (3, 16, 16, 3, 7), (1, 20, 20, 2.5, 9), (2, 12, 16, 0, -1)]):
        masks[f"case{i}_args"] = np.array([nz, ny, nx, rmin, rmax], dtype=np.float64)
        masks[f"case{i}"] = helicon.get_cylindrical_mask(nz, ny, nx, rmin=rmin, rmax=rmax)
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **masks)
    bp = {}
    for i, (N, s, D2, L2) in enumerate([(8, 1.0, 4, 6), (12, 0.5, 8, 10), (10, 1.0, -1, -1)]):
        img = rng.random((N, N)).astype(np.float32)
        (X, Y, Z), pv = S.back_project_2d_coords_to_3d_coords(img, s, D2, L2)
        bp[f"case{i}_args"] = np.array([N, s, D2, L2], dtype=np.float64)
        bp[f"case{i}_img"] = img
        bp[f"case{i}_X"], bp[f"case{i}_Y"], bp[f"case{i}_Z"], bp[f"case{i}_pix"] = X, Y, Z, pv
    np.savez_compressed(os.path.join(OUT, "backproject.npz"), **bp)

    # data rows ---------------------------------------------------------------
    for name, N, s, twist, rise, csym, D2, L2, D3, D3i, L3, mpl, interp in DATA_CASES:
        img = rng.random((N, N)).astype(np.float32)
        A, b, pid = S.build_A_data_matrix.__wrapped__(
            image=img, scale2d_to_3d=s, twist_degree=twist, rise_pixel=rise, csym=csym, tilt_degree=0, psi_degree=0,
            dy_pixel=0, reconstruct_diameter_2d_pixel=D2, reconstruct_length_2d_pixel=L2,
            reconstruct_diameter_3d_pixel=D3, reconstruct_diameter_3d_inner_pixel=D3i,
            reconstruct_length_3d_pixel=L3, min_projection_lines=mpl, interpolation=interp, verbose=0, cpu=1,
        )
        d = dict(image=img, args=np.array([s, twist, rise, csym, D2, L2, D3, D3i, L3, mpl], dtype=np.float64), b=b, b_pid=pid)
        d.update(csr_parts(A, "A"))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        manifest[name] = dict(kind="data", interpolation=interp, shape=list(A.shape), nnz=int(A.nnz))
        print(name, A.shape, A.nnz)

    # symmetry rows -----------------------------------------------------------
    for name, nz, ny, nx, twist, rise, csym, rmin, rmax, msp, interp in HSYM_CASES:
        A, b = S.build_A_helical_sym_matrix.__wrapped__(nz, ny, nx, twist, rise, csym, rmin, rmax, msp, interp, verbose=0)
        d = dict(args=np.array([nz, ny, nx, twist, rise, csym, rmin, rmax, msp], dtype=np.float64))
        d.update(csr_parts(A, "A"))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        manifest[name] = dict(kind="hsym", interpolation=interp, shape=list(A.shape), nnz=int(A.nnz))
        print(name, A.shape, A.nnz)

    # full solve + score -----------------------------------------------------
    for name, N, apix, twist, rise, csym, pc, so, interp, L3 in SOLVE_CASES:
        img = synth_image(N, apix, twist=twist, rise=rise, csym=csym)
        S.build_A_data_matrix.clear_cache()
        (rec3d, _, _), score = S.lsq_reconstruct(
            projection_image=img, scale2d_to_3d=1.0, twist_degree=twist, rise_pixel=rise / apix, csym=csym,
            positive_constraint=pc, reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
            reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=L3, sym_oversample=so,
            interpolation=interp, algorithm=dict(model="lsq"), cpu=1,
        )
        d = dict(image=img, args=np.array([apix, twist, rise, csym, pc, so, L3], dtype=np.float64), rec3d=rec3d,
                 score=np.float64(score))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        manifest[name] = dict(kind="solve", interpolation=interp, score=float(score), shape=list(rec3d.shape))
        print(name, rec3d.shape, float(score))

    manifest["_versions"] = dict(
        helicon=helicon.__version__, numpy=np.__version__, scipy=__import__("scipy").__version__,
        numba=__import__("numba").__version__,
    )
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
