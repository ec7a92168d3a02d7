"""Full-size goldens at the BASELINE config shapes (test infrastructure; needs /root/reference, which does not travel).

Part A -- cfg1 (200x200) and cfg2 (256x256), the bench's own synthetic filament, two candidates each, unbounded
(positive_constraint=0) AND the reference-default rule (positive_constraint=-1 -> bounded TRF): the UNMODIFIED
reference's ``lsq_reconstruct`` (SLR:31-547) is run with spies on scipy's ``lsmr`` / ``trf_linear`` (they only record
``itn``/``istop``/``nit``; nothing is altered).  Stored: image, args, score, itn, istop, trf nit, x = rec3d[mask] (float32).

Part B -- shapes that force the other kernel templates (512x512 -> uint32 maps; 384x384 with rise >= 21 A -> L3P > 16):
the reference needs hours there, so the oracle's ``lsmr_mixed(fixed_iters=20)`` iterate (scipy's executed precision
map; pinned to scipy bit for bit in tests/test_oracle_golden.py; here with exactly rounded norms,
and the iterate with scipy's OpenBLAS float32 norms stored beside it -- see lsmr_mixed's docstring) on rows from
``build_A_data_matrix_fast`` (pinned bit-exact to the reference's builder, tests/test_oracle_golden.py) and the
reference's own ``build_A_helical_sym_matrix``.  Stored: a strided sample of x, ||x||, score of the iterate, and the
noise floor of that iterate under a row permutation (same maths, other float32 summation order).

Usage:  python oracle/make_golden_fullsize.py [case ...]      (one process per case is fine; they are independent)
        HB2_BAND_TAG=t8 OMP_NUM_THREADS=8 python oracle/make_golden_fullsize.py case   (re-run under another BLAS
        configuration -> /tmp/gold/band_*.npz);  python oracle/make_golden_fullsize.py merge_band -> full_band.npz
"""
import os
import sys
import time
import warnings

os.environ.setdefault("HELION_CACHE_DIR", "/tmp/helicon_cache_golden_full")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
os.environ.setdefault("OMP_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.join(HERE, ".."))
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
APIX = 1.3

# name: (N, twist, rise_A, csym, positive_constraint)
REF_CASES = {
    "full_cfg1_true_unb": (200, -1.2, 4.75, 1, 0),
    "full_cfg1_true_pos": (200, -1.2, 4.75, 1, -1),
    "full_cfg1_b_unb": (200, -2.03, 4.62, 1, 0),
    "full_cfg1_b_pos": (200, -2.03, 4.62, 1, -1),
    "full_cfg2_true_unb": (256, -1.2, 4.75, 1, 0),
    "full_cfg2_true_pos": (256, -1.2, 4.75, 1, -1),
    "full_cfg2_b_unb": (256, -2.03, 4.62, 1, 0),
    "full_cfg2_b_pos": (256, -2.03, 4.62, 1, -1),
    # trilinear interpolation (the app's default) at the benchmark shape: the matrix-free trilinear path (bilinear.py)
    "full_cfg2_b_lin_unb": (256, -2.03, 4.62, 1, 0, "linear"),
    "full_cfg1_b_lin_unb": (200, -2.03, 4.62, 1, 0, "linear"),
    "full_cfg1_b_lin_pos": (200, -2.03, 4.62, 1, -1, "linear"),
}
# name: (N, twist, rise_A, csym, fixed_iters)
FIXED_CASES = {
    "fixed_512_u32": (512, -1.37, 4.8137, 1, 20),
    "fixed_512_c3": (512, 21.3, 9.7, 3, 20),
    "fixed_384_l3p52": (384, 47.3, 21.7, 1, 20),
    "fixed_384_l3p104": (384, -101.9, 44.3, 1, 20),
}
STRIDE = 16


def geometry(N, rise):
    from helicon_b200.grid import derive_geometry  # pipeline.py:253-349 (pinned by the task_* goldens)

    return derive_geometry(N, N, APIX, rise, rise, N * APIX, 0.0, N * APIX, 3 * rise, APIX, 0, -1)


def image_for(N, twist=-1.2, rise=4.75):
    import bench

    return bench.synthetic_filament(n=N, apix=APIX, twist=twist, rise=rise)


def run_reference(name):
    from helicon.webApps.denovo3D import solver_linear_regression as S
    import importlib

    LL = importlib.import_module("scipy.optimize._lsq.lsq_linear")

    N, twist, rise, csym, pc = REF_CASES[name][:5]
    interp = REF_CASES[name][5] if len(REF_CASES[name]) > 5 else "nn"
    g = geometry(N, rise)
    img = image_for(N)
    spy = dict(lsmr=[], trf=[])
    real_lsmr, real_trf = LL.lsmr, LL.trf_linear

    def lsmr_spy(*a, **k):
        r = real_lsmr(*a, **k)
        spy["lsmr"].append((int(r[2]), int(r[1])))
        return r

    def trf_spy(*a, **k):
        r = real_trf(*a, **k)
        spy["trf"].append((int(r.nit), int(r.status)))
        return r

    LL.lsmr, LL.trf_linear = lsmr_spy, trf_spy
    S.build_A_data_matrix.clear_cache()
    S.build_A_helical_sym_matrix.clear_cache()
    t0 = time.time()
    (rec3d, _, _), score = S.lsq_reconstruct(
        projection_image=img, scale2d_to_3d=g["s"], twist_degree=twist, rise_pixel=rise / g["apix3d"], csym=csym,
        positive_constraint=pc, reconstruct_diameter_3d_inner_pixel=g["D3i"], reconstruct_diameter_2d_pixel=g["D2"],
        reconstruct_length_2d_pixel=g["L2"], reconstruct_diameter_3d_pixel=g["D3"], reconstruct_length_3d_pixel=g["L3"],
        sym_oversample=g["sym_oversample"], interpolation=interp, algorithm=dict(model="lsq"), cpu=1)
    dt = time.time() - t0
    LL.lsmr, LL.trf_linear = real_lsmr, real_trf
    from oracle import denovo3d_oracle as O

    mask = O.cylindrical_mask(g["L3"], g["D3"], g["D3"], g["D3i"] / 2, g["D3"] // 2 - 1)
    assert float(np.abs(rec3d[~mask]).max()) == 0.0
    itn, istop = spy["lsmr"][0]
    nit, status = spy["trf"][0] if spy["trf"] else (0, 0)
    if BAND_TAG:  # the same reference call under another BLAS configuration (threads / kernel): only the deltas are kept
        gold = np.load(os.path.join(OUT, name + ".npz"))
        xg = rec3d[mask].astype(np.float32)
        rel = float(np.linalg.norm(xg - gold["x"]) / np.linalg.norm(gold["x"]))
        np.savez(f"/tmp/gold/band_{name}_{BAND_TAG}.npz", row=np.array(
            [itn - int(gold["itn"]), nit - int(gold["trf_nit"]), float(score) - float(gold["score"]), rel], dtype=np.float64))
        print("band", BAND_TAG, name, "d itn", itn - int(gold["itn"]), "d trf", nit - int(gold["trf_nit"]), "dscore",
              float(score) - float(gold["score"]), "rel", rel, f"{dt:.1f}s", flush=True)
        return
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), image=img,
        args=np.array([APIX, twist, rise, csym, pc, g["sym_oversample"], g["L3"], g["D2"], g["L2"], g["D3"]], dtype=np.float64),
        x=rec3d[mask].astype(np.float32), score=np.float64(score), itn=np.int64(itn), istop=np.int64(istop),
        trf_nit=np.int64(nit), trf_status=np.int64(status), seconds=np.float64(dt))
    print(name, rec3d.shape, "score", float(score), "itn", itn, "istop", istop, "trf", nit, status, f"{dt:.1f}s", flush=True)


def run_fixed(name):
    from helicon.webApps.denovo3D import solver_linear_regression as S
    from oracle import denovo3d_oracle as O
    from scipy.sparse import vstack

    N, twist, rise, csym, iters = FIXED_CASES[name]
    g = geometry(N, rise)
    img = image_for(N, twist=twist if abs(twist) < 5 else -1.2, rise=rise)
    D2, L2, D3, L3 = g["D2"], g["L2"], g["D3"], g["L3"]
    rise_px = rise / g["apix3d"]
    mask = O.cylindrical_mask(L3, D3, D3, 0, D3 // 2 - 1)
    n3 = int(np.count_nonzero(mask))
    target = min(O.MAX_EQUATIONS, int(max(D2 * L2, n3) * g["sym_oversample"]))
    t0 = time.time()
    A_data, b_data, b_pid = O.build_A_data_matrix_fast(img, g["s"], twist, rise_px, csym, D2, L2, D3, 0, L3, target)
    t1 = time.time()
    A_hsym, b_hsym = S.build_A_helical_sym_matrix.__wrapped__(L3, D3, D3, twist, rise_px, csym, 0, D3 // 2 - 1, target,
                                                              "nn", verbose=0)
    t2 = time.time()
    A = vstack((A_data, A_hsym)).tocsr()
    b = np.concatenate((b_data, b_hsym)).astype(np.float32)
    x = O.lsmr_mixed(A, b, fixed_iters=iters, norm="exact")[0]   # exactly rounded norms (what the CUDA path computes)
    t3 = time.time()
    xb = O.lsmr_mixed(A, b, fixed_iters=iters, norm="blas")[0]   # scipy as executed HERE: OpenBLAS float32 sdot
    blas_rel = float(np.linalg.norm(xb - x) / np.linalg.norm(x))
    # the iterate's own float32 noise floor: same equations, permuted order
    perm = np.random.default_rng(1).permutation(A.shape[0])
    xp = O.lsmr_mixed(A[perm], b[perm], fixed_iters=iters, norm="exact")[0]
    floor = float(np.linalg.norm(x - xp) / np.linalg.norm(x))
    x32 = x.astype(np.float32)
    score = O.cosine_similarity(A_data.dot(x32), b_data)
    score_p = O.cosine_similarity(A_data.dot(xp.astype(np.float32)), b_data)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        args=np.array([APIX, twist, rise, csym, 0, g["sym_oversample"], L3, D2, L2, D3, iters, STRIDE], dtype=np.float64),
        image=img,
        x_sample=x32[::STRIDE].copy(), x_sample_blas=xb.astype(np.float32)[::STRIDE].copy(), blas_rel=np.float64(blas_rel), x_norm=np.float64(np.linalg.norm(x)), x_sum=np.float64(x.sum()),
        n=np.int64(len(x)), m_data=np.int64(A_data.shape[0]), m_sym=np.int64(A_hsym.shape[0]),
        nnz_data=np.int64(A_data.nnz), score=np.float64(score), score_perm=np.float64(score_p), floor=np.float64(floor),
        b_sum=np.float64(b_data.astype(np.float64).sum()), pid_sum=np.int64(b_pid.astype(np.int64).sum()))
    print(name, "n", len(x), "rows", A_data.shape[0], A_hsym.shape[0], "nnz", A_data.nnz, "score", float(score),
          "floor", floor, "blas-vs-exact", blas_rel, f"build {t1 - t0:.0f}s sym {t2 - t1:.0f}s lsmr {t3 - t2:.0f}s", flush=True)


def merge_band():
    """tests/golden/full_band.npz: per case the rows [d itn, d trf nit, d score, rel-L2(x)] of the UNMODIFIED reference
    re-run under other BLAS configurations (numpy's float32 norm = OpenBLAS sdot: its float32 accumulation order, and with
    it the LSMR scalars, depends on the CPU kernel OpenBLAS dispatches and on its thread count)."""
    import glob

    out = {}
    for f in sorted(glob.glob("/tmp/gold/band_*.npz")):
        base = os.path.basename(f)[5:-4]
        case, tag = base.rsplit("_", 1)
        out.setdefault(case, []).append((tag, np.load(f)["row"]))
    flat = {}
    for case, rows in out.items():
        flat[case] = np.stack([r for _, r in rows])
        flat[case + "_tags"] = np.array([t for t, _ in rows])
        print(case, [(t, r.tolist()) for t, r in rows])
    np.savez_compressed(os.path.join(OUT, "full_band.npz"), **flat)


BAND_TAG = os.environ.get("HB2_BAND_TAG", "")

if __name__ == "__main__":
    if sys.argv[1:] == ["merge_band"]:
        merge_band()
        sys.exit(0)
    names = sys.argv[1:] or list(REF_CASES) + list(FIXED_CASES)
    for nm in names:
        (run_reference if nm in REF_CASES else run_fixed)(nm)
