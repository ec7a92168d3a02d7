"""How reproducible is the reference's DEFAULT (bounded TRF) path at the BASELINE cfg2 shape?  (test infrastructure; needs
/root/reference.)  For a seeded sample of cfg2 grid candidates the reference's equations (its own builders) are solved by
the reference's scipy call (SLR:258-270) twice: in the reference's row order and in a permuted row order (same maths,
different float32 summation order).  Stored per candidate: LSMR iterations, TRF iterations and score of both runs ->
tests/golden/bounded_band_cfg2.npz.  The GPU test holds the CUDA path's deviation from run 0 against the deviation of
run 1 from run 0 (tests/test_gpu_fullsize_parity.py::test_bounded_path_statistics_vs_reference_band).

scipy's trf_linear exits on `cost_change < 1e-2 * cost` after 13-25 outer iterations of 1-3 inner LSMR iterations each,
far from the bounded optimum: whether an outer iteration takes a Newton, reflected or anti-gradient step and whether it is
the last one are discontinuous decisions, so 1e-3 differences of the LSMR start (the reference's own float32 noise)
occasionally end on another branch, with score changes of 1e-3...2e-2.

Usage: python oracle/make_golden_bounded_band.py <index> ...   (writes /tmp/gold/bb_<index>.npz)
       python oracle/make_golden_bounded_band.py merge
"""
import glob
import importlib
import os
import sys
import warnings

os.environ.setdefault("HELION_CACHE_DIR", "/tmp/helicon_cache_bb_%d" % os.getpid())
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
os.environ.setdefault("OMP_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")
import numpy as np  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
NCAND = 16


def candidates():
    import bench

    rng = np.random.default_rng(314)
    tw = rng.choice(len(bench.TWISTS), NCAND, replace=False)
    ri = rng.integers(0, len(bench.RISES), NCAND)
    return [(float(np.round(bench.TWISTS[a], 6)), float(bench.RISES[b])) for a, b in zip(tw, ri)]


def run(i):
    import helicon
    from helicon.webApps.denovo3D import solver_linear_regression as S
    from scipy.optimize import lsq_linear
    from scipy.sparse import vstack
    from make_golden_fullsize import geometry, image_for
    from oracle import denovo3d_oracle as O

    twist, rise = candidates()[i]
    N = 256
    g = geometry(N, rise)
    img = image_for(N)
    rise_px = rise / g["apix3d"]
    mask = O.cylindrical_mask(g["L3"], g["D3"], g["D3"], 0, g["D3"] // 2 - 1)
    n3 = int(mask.sum())
    target = min(O.MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
    A_data, b_data, _ = S.build_A_data_matrix.__wrapped__(
        image=img, scale2d_to_3d=g["s"], twist_degree=twist, rise_pixel=rise_px, csym=1, tilt_degree=0, psi_degree=0,
        dy_pixel=0, reconstruct_diameter_2d_pixel=g["D2"], reconstruct_length_2d_pixel=g["L2"],
        reconstruct_diameter_3d_pixel=g["D3"], reconstruct_diameter_3d_inner_pixel=0,
        reconstruct_length_3d_pixel=g["L3"], min_projection_lines=target, interpolation="nn", verbose=0, cpu=1)
    A_hsym, b_hsym = S.build_A_helical_sym_matrix.__wrapped__(g["L3"], g["D3"], g["D3"], twist, rise_px, 1, 0,
                                                              g["D3"] // 2 - 1, target, "nn", verbose=0)
    A = vstack((A_data, A_hsym)).tocsr()
    b = np.concatenate((b_data, b_hsym))
    LL = importlib.import_module("scipy.optimize._lsq.lsq_linear")
    real = LL.lsmr
    rows = []
    for seed in (-1, 0):
        itns = []

        def spy(*a, **k):
            r = real(*a, **k)
            itns.append(int(r[2]))
            return r

        LL.lsmr = spy
        p = np.arange(A.shape[0]) if seed < 0 else np.random.default_rng(100 + seed).permutation(A.shape[0])
        res = lsq_linear(A[p].tocsr(), b[p], bounds=(0.0, float(np.max(b_data))), tol=1e-2, max_iter=200,
                         lsmr_maxiter=1000, lsmr_tol="auto", verbose=0)
        LL.lsmr = real
        score = float(helicon.cosine_similarity(A_data.dot(res.x.astype(np.float32)), b_data))
        rows.append([itns[0], res.nit, score])
        print(i, twist, rise, "seed", seed, rows[-1], flush=True)
    np.savez(f"/tmp/gold/bb_{i}.npz", cand=np.array([twist, rise]), rows=np.array(rows, dtype=np.float64))


if __name__ == "__main__":
    if sys.argv[1] == "merge":
        cands, rows = [], []
        for i in range(NCAND):
            d = np.load(f"/tmp/gold/bb_{i}.npz")
            cands.append(d["cand"]); rows.append(d["rows"])
        np.savez_compressed(os.path.join(OUT, "bounded_band_cfg2.npz"), cand=np.array(cands), rows=np.array(rows))
        r = np.array(rows)
        print("reference run 1 vs run 0: |dscore|", np.abs(r[:, 1, 2] - r[:, 0, 2]).tolist())
    else:
        for a in sys.argv[1:]:
            run(int(a))
