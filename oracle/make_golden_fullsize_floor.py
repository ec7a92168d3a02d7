"""The reference's OWN reproducibility floor at the BASELINE shapes (companion of make_golden_fullsize.py; test
infrastructure, needs /root/reference).  For each full-size case the reference's builders give A, b; the reference's
scipy call (SLR:258-270) is repeated on the same equations in permuted row order (same maths, different float32
summation order in A^T u) and the changes of the stopping iteration, the TRF outer iterations, the score and x against
the unpermuted golden are stored in tests/golden/full_floor.npz -- the band the GPU parity test is held to.

Usage: python oracle/make_golden_fullsize_floor.py <N> <twist> <rise>   (writes /tmp/gold/floor_<N>_<twist>.npz)
       python oracle/make_golden_fullsize_floor.py merge                 (collects them into tests/golden/full_floor.npz)
"""
import glob
import os
import sys
import warnings

os.environ.setdefault("HELION_CACHE_DIR", "/tmp/helicon_cache_golden_floor")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
os.environ.setdefault("OMP_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")
import numpy as np  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
NPERM = 3


def run(N, twist, rise):
    import helicon
    from helicon.webApps.denovo3D import solver_linear_regression as S
    from scipy.optimize import lsq_linear
    from scipy.sparse import vstack
    from make_golden_fullsize import APIX, REF_CASES, geometry, image_for
    from oracle import denovo3d_oracle as O

    g = geometry(N, rise)
    img = image_for(N)
    rise_px = rise / g["apix3d"]
    mask = O.cylindrical_mask(g["L3"], g["D3"], g["D3"], 0, g["D3"] // 2 - 1)
    n3 = int(np.count_nonzero(mask))
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
    names = {(v[0], v[1], v[4]): k for k, v in REF_CASES.items()}
    out = {}
    for pc in (0, -1):
        gold = np.load(os.path.join(OUT, names[(N, twist, pc)] + ".npz"))
        lb, ub = (0.0, float(np.max(b_data))) if pc else (-np.inf, np.inf)
        rows = []
        for seed in range(NPERM):
            p = np.random.default_rng(100 + seed).permutation(A.shape[0])
            res = lsq_linear(A[p].tocsr(), b[p], bounds=(lb, ub), tol=1e-2, max_iter=200, lsmr_maxiter=1000,
                             lsmr_tol="auto", verbose=0)
            x = res.x.astype(np.float32)
            score = float(helicon.cosine_similarity(A_data.dot(x), b_data))
            rel = float(np.linalg.norm(x - gold["x"]) / np.linalg.norm(gold["x"]))
            rows.append([res.nit, score - float(gold["score"]), rel])
            print(names[(N, twist, pc)], "perm", seed, "nit", res.nit, "gold itn/trf", int(gold["itn"]), int(gold["trf_nit"]),
                  "dscore", score - float(gold["score"]), "rel", rel, flush=True)
        out[names[(N, twist, pc)]] = np.array(rows, dtype=np.float64)
    np.savez(f"/tmp/gold/floor_{N}_{twist}.npz", **out)


if __name__ == "__main__":
    if sys.argv[1] == "merge":
        allv = {}
        for f in sorted(glob.glob("/tmp/gold/floor_*.npz")):
            d = np.load(f)
            for k in d.files:
                allv[k] = d[k]   # rows: [nit (LSMR itn when unbounded, TRF outer iterations when bounded), dscore, rel-L2(x)]
        np.savez_compressed(os.path.join(OUT, "full_floor.npz"), **allv)
        for k, v in allv.items():
            print(k, v.tolist())
    else:
        run(int(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3]))
