"""Goldens for the sklearn-model branch of solve_equations (SLR:272-342: ElasticNet / Lasso / Ridge with selection="random",
tol=1e-2, max_iter=200) from the UNMODIFIED reference (needs /root/reference and scikit-learn).  The reference's coordinate
descent is stochastic (np.random global state) and loosely converged, so next to its result the script stores the value
of sklearn's own objective at that result and -- as the convergence yardstick -- the same model refitted with tol=1e-8,
max_iter=20000, cyclic selection (the minimiser the stochastic run approximates).
Usage: python oracle/make_golden_models.py.  TEST INFRASTRUCTURE ONLY."""
import os
import sys
import warnings

os.environ.setdefault("HELION_CACHE_DIR", "/tmp/helicon_cache_golden")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
from scipy.sparse import vstack  # noqa: E402
from helicon.webApps.denovo3D import solver_linear_regression as S  # noqa: E402
from make_golden import OUT, synth_image  # noqa: E402

# name, N, apix, twist, rise_A, csym, L3, sym_oversample, positive_constraint, algorithm
CASES = [
    ("model_enet_40", 40, 6.5, -3.5, 9.5, 1, 6, 2, 0, dict(model="elasticnet")),
    ("model_enet_40_pos", 40, 6.5, -3.5, 9.5, 1, 6, 2, 1, dict(model="elasticnet", alpha=1e-4, l1_ratio=0.5)),
    ("model_lasso_32", 32, 8.125, 27.0, 12.0, 2, 6, 2, 0, dict(model="lasso", alpha=1e-4)),
    ("model_ridge_32", 32, 8.125, 27.0, 12.0, 2, 6, 2, 0, dict(model="ridge", alpha=1)),
    ("model_ridge_32_pos", 32, 8.125, 27.0, 12.0, 2, 6, 2, 1, dict(model="ridge", alpha=1)),
]


def objective(model, A, b, w, alpha, l1_ratio):
    """sklearn's objective with the fitted intercept profiled out (fit_intercept=True)."""
    r = b - A @ w
    r = r - r.mean()
    m = A.shape[0]
    if model == "ridge":
        return float(r @ r + alpha * (w @ w))
    return float(r @ r / (2 * m) + alpha * l1_ratio * np.abs(w).sum() + 0.5 * alpha * (1 - l1_ratio) * (w @ w))


for name, N, apix, twist, rise, csym, L3, so, pc, alg in CASES:
    img = synth_image(N, apix, twist=twist, rise=rise, csym=csym)
    S.build_A_data_matrix.clear_cache()
    S.build_A_helical_sym_matrix.clear_cache()
    np.random.seed(7)
    kw = dict(scale2d_to_3d=1.0, twist_degree=twist, rise_pixel=rise / apix, csym=csym, positive_constraint=pc,
              reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N,
              reconstruct_length_3d_pixel=L3, sym_oversample=so, interpolation="nn", algorithm=alg)
    (rec3d, _, _), score = S.lsq_reconstruct(img, **kw)
    # the system itself, for the objective values and the tight refit
    target = int(max(N * N, L3 * int(np.count_nonzero(rec3d[0] == rec3d[0]))) * so)
    mask = S.helicon.get_cylindrical_mask(nz=L3, ny=N, nx=N, rmin=0, rmax=N // 2 - 1)
    x_ref = rec3d[mask].astype(np.float64)
    n_eq = int(max(N * N, int(mask.sum())) * so)
    A_d, b_d, _ = S.build_A_data_matrix(image=img, scale2d_to_3d=1.0, twist_degree=twist, rise_pixel=rise / apix, csym=csym,
                                        tilt_degree=0, psi_degree=0, dy_pixel=0, reconstruct_diameter_2d_pixel=N,
                                        reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N,
                                        reconstruct_diameter_3d_inner_pixel=0, reconstruct_length_3d_pixel=L3,
                                        min_projection_lines=n_eq, interpolation="nn", verbose=0)
    A_s, b_s = S.build_A_helical_sym_matrix(nz=L3, ny=N, nx=N, twist_degree=twist, rise_pixel=rise / apix, csym=csym, rmin=0,
                                            rmax=N // 2 - 1, min_sym_pairs=n_eq, interpolation="nn", verbose=0)
    A = vstack((A_d, A_s)).tocsr().astype(np.float64)
    b = np.concatenate((b_d, b_s)).astype(np.float64)
    model = alg["model"]
    alpha = alg.get("alpha", 1 if model == "ridge" else 1e-4)
    l1r = 1.0 if model == "lasso" else alg.get("l1_ratio", 0.5)
    positive = pc > 0
    if model == "ridge":
        from sklearn.linear_model import Ridge
        tight = Ridge(alpha=alpha, fit_intercept=True, positive=positive, tol=1e-10, max_iter=20000)
    else:
        from sklearn.linear_model import ElasticNet
        tight = ElasticNet(alpha=alpha, l1_ratio=l1r, fit_intercept=True, positive=positive, selection="cyclic", tol=1e-9,
                           max_iter=20000)
    tight.fit(A, b)
    x_tight = tight.coef_.astype(np.float64)
    f_ref, f_tight = objective(model, A, b, x_ref, alpha, l1r), objective(model, A, b, x_tight, alpha, l1r)
    pred = A_d @ x_tight.astype(np.float32)
    score_tight = float(S.helicon.cosine_similarity(pred, b_d))
    rel = float(np.linalg.norm(x_ref - x_tight) / np.linalg.norm(x_tight))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img,
                        args=np.array([apix, twist, rise, csym, L3, so, pc, alpha, l1r], dtype=np.float64),
                        model=np.array(model), score=np.float32(score), rec3d=rec3d, x_tight=x_tight,
                        score_tight=np.float32(score_tight), f_ref=f_ref, f_tight=f_tight, rel_ref_vs_tight=rel)
    print(f"{name}: score {float(score):.6f} tight {score_tight:.6f}; objective ref {f_ref:.8e} tight {f_tight:.8e}; "
          f"rel-L2(x_ref, x_tight) {rel:.2e}; nnz {np.count_nonzero(x_ref)}/{len(x_ref)}")
