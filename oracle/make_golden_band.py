"""Reproducibility band of the reference itself for the trilinear solve goldens (needs /root/reference): the same
scipy solve with the equations permuted (float32 LSMR noise, amplified on the bounded branch by its data-dependent
decisions).  Adds band_dscore / band_relx / band_itn to tests/golden/gen_solve_lin_*.npz; the GPU parity test accepts
max(north-star tolerance, 2 x band).  Usage: python oracle/make_golden_band.py.  TEST INFRASTRUCTURE ONLY."""
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
from scipy.optimize import lsq_linear  # noqa: E402
from scipy.sparse import vstack  # noqa: E402
from helicon.webApps.denovo3D import solver_linear_regression as S  # noqa: E402
from make_golden import OUT  # noqa: E402

for name in ["gen_solve_lin_48", "gen_solve_lin_48_pos", "gen_solve_lin_40_c2_tilt"]:
    path = os.path.join(OUT, name + ".npz")
    d = dict(np.load(path))
    apix, twist, rise, csym, pc, so, L3, tilt, psi, dy = d["args"]
    img = d["image"]
    N, L3, csym = img.shape[0], int(L3), int(csym)
    n3 = int(np.count_nonzero(helicon.get_cylindrical_mask(L3, N, N, rmin=0, rmax=N // 2 - 1)))
    target = min(2**26, int(max(N * N, n3) * so))
    A_d, b_d, _ = S.build_A_data_matrix.__wrapped__(
        image=img, scale2d_to_3d=1.0, twist_degree=twist, rise_pixel=rise / apix, csym=csym, tilt_degree=tilt,
        psi_degree=psi, dy_pixel=dy, reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
        reconstruct_diameter_3d_pixel=N, reconstruct_diameter_3d_inner_pixel=0, reconstruct_length_3d_pixel=L3,
        min_projection_lines=target, interpolation="linear", verbose=0, cpu=1)
    A_s, b_s = S.build_A_helical_sym_matrix.__wrapped__(L3, N, N, twist, rise / apix, csym, 0, N // 2 - 1, target, "linear")
    A = vstack((A_d, A_s)).tocsr()
    b = np.concatenate((b_d, b_s))
    bounds = (0.0, float(b_d.max())) if pc > 0 else (-np.inf, np.inf)
    solve = lambda A_, b_: lsq_linear(A_, b_, bounds=bounds, tol=1e-2, max_iter=200, lsmr_maxiter=1000, lsmr_tol="auto")
    sc = lambda x: float(helicon.cosine_similarity(A_d.dot(x.astype(np.float32)), b_d))
    r0 = solve(A, b)
    x0 = r0.x.astype(np.float32)
    assert abs(sc(x0) - float(d["score"])) < 1e-7, (sc(x0), float(d["score"]))
    ds, dx, its = [], [], [r0.unbounded_sol[2] if pc <= 0 else r0.nit]
    for seed in range(4):
        p = np.random.default_rng(seed).permutation(A.shape[0])
        r = solve(A[p].tocsr(), b[p])
        x = r.x.astype(np.float32)
        ds.append(abs(sc(x) - sc(x0)))
        dx.append(float(np.linalg.norm(x - x0) / np.linalg.norm(x0)))
        its.append(r.unbounded_sol[2] if pc <= 0 else r.nit)
    d["band_dscore"], d["band_relx"], d["band_itn"] = np.float64(max(ds)), np.float64(max(dx)), np.array(its)
    np.savez_compressed(path, **d)
    print(name, "band dscore", max(ds), "relx", max(dx), "iterations", its)
