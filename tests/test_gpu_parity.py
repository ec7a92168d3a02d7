"""GPU parity tests: the CUDA path (through the C ABI) against the committed
reference outputs (tests/golden) and against the CPU oracle on seeded inputs.

Tolerances (SURVEY.md sections 7/8c, BASELINE.json north_star):
  * row sets / matrices: bit-exact (integer index work);
  * operator application: float32 round-off (summation order differs);
  * score: 1e-5 absolute;
  * reconstructed slices: reported against the reference's own reproducibility
    floor (a row permutation moves the reference's loosely converged LSMR
    iterate by 2e-4..8e-4 rel-L2); asserted < 5e-3 at the stopping
    point and < 1e-4 at a fixed iteration count.
"""
import numpy as np
import pytest
from scipy.sparse import vstack

from oracle import denovo3d_oracle as O
from tests.helpers import cases, csr_equal, csr_from, load, oracle_permutation_floor  # noqa: F401

pytestmark = pytest.mark.gpu

# rounding ties are resolved exactly from the reference's coordinate tables: in-plane ties (view angles of
# 30/60/... degrees, or s = 0.5) by per-column exact maps, column->slice ties by per-sample slices
XY_TIE_EXACT = {"data_nn_tie30", "data_nn_reftest", "data_nn_inner"}  # data_nn_s05: a view with BOTH tie kinds stays flagged
Z_TIE_EXACT = {"data_nn_tiez", "data_nn_csym2"}


@pytest.fixture(scope="module")
def solver():
    from helicon_b200 import solver_linear_regression as S

    return S


def _data_args(d):
    s, twist, rise, csym, D2, L2, D3, D3i, L3, mpl = d["args"]
    return dict(image=d["image"], scale2d_to_3d=float(s), twist_degree=float(twist), rise_pixel=float(rise),
                csym=int(csym), tilt_degree=0, psi_degree=0, dy_pixel=0, reconstruct_diameter_2d_pixel=int(D2),
                reconstruct_length_2d_pixel=int(L2), reconstruct_diameter_3d_pixel=int(D3),
                reconstruct_diameter_3d_inner_pixel=int(D3i), reconstruct_length_3d_pixel=int(L3),
                min_projection_lines=int(mpl), interpolation="nn")


@pytest.mark.parametrize("name", cases("data", "nn"))
def test_data_rows_vs_reference(name):
    """Row sets, b and pixel ids bit-exact against the reference.  Geometries whose
    rounding decisions follow the reference's last-bit coordinate noise
    must at least be FLAGGED (HB2_FLAG_TIE_*); for them equality is reported only."""
    from helicon_b200.engine import Batch, Problem
    from helicon_b200.planner import CandidateSpec

    d = load(name)
    s, twist, rise, csym, D2, L2, D3, D3i, L3, mpl = d["args"]
    prob = Problem(d["image"], float(s), int(D2), int(L2), int(D2), D3i / 2, int(D3) // 2 - 1)
    batch = Batch(prob, int(L3), [CandidateSpec(twist, rise, int(csym), int(mpl), -1, False)])
    used = np.unique(batch.plan.views["angle"])
    flagged = int(batch.tie[used].sum()) > 0 or bool(batch.plan.cand_tie_z[0])
    A, b, pid = batch.data_csr(0)
    ref = csr_from(d)
    ok, why = csr_equal(A, ref)
    print(f"{name}: flagged={flagged} rows gpu={A.shape[0]} ref={ref.shape[0]} identical={ok}")
    batch.close(); prob.close()
    if name in XY_TIE_EXACT:
        assert int(batch.tie.sum()) > 0 and not flagged  # tie angles present, every view on them replaced by exact maps
    if name in Z_TIE_EXACT:
        assert batch.plan.has_ties and not flagged
    if not flagged:
        assert ok, why
        assert np.array_equal(b, d["b"]) and b.dtype == np.float32
        assert np.array_equal(pid, d["b_pid"]) and pid.dtype == np.int32


def test_data_rows_enough_unflagged_cases():
    assert len(cases("data", "nn")) >= 8


@pytest.mark.parametrize("name", cases("hsym", "nn"))
def test_hsym_rows_bit_exact_vs_reference(solver, name):
    d = load(name)
    nz, ny, nx, twist, rise, csym, rmin, rmax, msp = d["args"]
    A, b = solver.build_A_helical_sym_matrix(int(nz), int(ny), int(nx), float(twist), float(rise), int(csym), rmin,
                                             int(rmax), int(msp), "nn")
    ok, why = csr_equal(A, csr_from(d))
    assert ok, why
    # csr_equal compares row by row, so the reference's row ORDER is checked too
    assert b.dtype == np.float32 and not b.any() and len(b) == A.shape[0]


def _make_batch(img, twist, rise_px, csym, L3, so, specs_extra=()):
    from helicon_b200.engine import Batch, Problem
    from helicon_b200.planner import CandidateSpec, MAX_EQUATIONS

    N = img.shape[0]
    prob = Problem(img, 1.0, N, N, N, 0.0, N // 2 - 1)
    target = min(MAX_EQUATIONS, int(max(N * N, L3 * prob.ndisk) * so))
    specs = [CandidateSpec(twist, rise_px, csym, target, target, False)]
    for tw, ri in specs_extra:
        specs.append(CandidateSpec(tw, ri, csym, target, target, False))
    return prob, Batch(prob, L3, specs), target


def _oracle_system(img, twist, rise_px, csym, L3, target):
    N = img.shape[0]
    A_d, b_d, pid = O.build_A_data_matrix(img, 1.0, twist, rise_px, csym, 0, 0, 0, N, N, N, 0, L3, target, "nn")
    A_s, b_s = O.build_A_helical_sym_matrix(L3, N, N, twist, rise_px, csym, 0.0, N // 2 - 1, target, "nn")
    return A_d, b_d, A_s


@pytest.mark.parametrize("name", ["solve_nn_unb_48_t35", "solve_nn_unb_48_c2"])
def test_operator_forward_adjoint_vs_oracle_csr(name):
    d = load(name)
    apix, twist, rise, csym, pc, so, L3 = d["args"]
    img = d["image"]
    prob, batch, target = _make_batch(img, float(twist), float(rise / apix), int(csym), int(L3), int(so))
    A_d, b_d, A_s = _oracle_system(img, float(twist), float(rise / apix), int(csym), int(L3), target)
    pidx, kk, jj = batch.data_row_index(0)
    nd_pad, tot = batch.rows_padded(0)
    assert len(pidx) == A_d.shape[0] and tot - nd_pad == A_s.shape[0]
    assert np.array_equal(batch.rhs_padded(0)[pidx], b_d)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(batch.n).astype(np.float32)
    y = batch.apply_forward(0, x)
    yd_ref = A_d.astype(np.float64) @ x.astype(np.float64)
    ys_ref = A_s.astype(np.float64) @ x.astype(np.float64)
    scale = np.abs(yd_ref).max()
    assert np.abs(y[:nd_pad][pidx] - yd_ref).max() <= 2e-6 * scale * np.sqrt(img.shape[0])
    pad_only = np.ones(nd_pad, bool); pad_only[pidx] = False
    assert not y[:nd_pad][pad_only].any()  # padded (non-existent) rows stay exactly zero
    # the symmetry rows are stored in internal voxel order per pair round; sidx[r] = stored position of the reference's row r
    sidx = batch.sym_row_order(0)
    assert np.array_equal(np.sort(sidx), np.arange(A_s.shape[0]))
    assert np.array_equal(y[nd_pad:][sidx], (A_s @ x).astype(np.float32))  # a-b: one rounding, bit-exact
    # adjoint
    u = np.zeros(tot, np.float32)
    u_real = rng.standard_normal(len(pidx) + A_s.shape[0]).astype(np.float32)
    u[:nd_pad][pidx] = u_real[: len(pidx)]
    u[nd_pad:][sidx] = u_real[len(pidx):]
    g = batch.apply_adjoint(0, u)
    g_ref = vstack((A_d, A_s)).astype(np.float64).T @ u_real.astype(np.float64)
    assert np.abs(g - g_ref).max() <= 2e-6 * np.abs(g_ref).max() * np.sqrt(img.shape[0])
    # <Ax,y> = <x,A^T y>
    lhs = float(np.dot(y.astype(np.float64), u.astype(np.float64)))
    rhs = float(np.dot(x.astype(np.float64), g.astype(np.float64)))
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)
    batch.close(); prob.close()


@pytest.mark.parametrize("name", ["solve_nn_unb_32", "solve_nn_unb_48_t35", "solve_nn_unb_48_c2", "solve_nn_unb_64",
                                  "solve_nn_unb_96_tie475"])
def test_unbounded_solve_vs_reference_golden(solver, name):
    d = load(name)
    apix, twist, rise, csym, pc, so, L3 = d["args"]
    img = d["image"]
    N = img.shape[0]
    (rec, h1, h2), score, info = solver.lsq_reconstruct(
        img, 1.0, float(twist), float(rise / apix), int(csym), positive_constraint=int(pc),
        reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N,
        reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so), interpolation="nn", return_info=True)
    ref = d["rec3d"]
    rel = float(np.linalg.norm(rec - ref) / np.linalg.norm(ref))
    dscore = abs(float(score) - float(d["score"]))
    print(f"{name}: itn={info['res']['itn']} istop={info['res']['istop']} score={float(score):.7f} "
          f"ref={float(d['score']):.7f} |dscore|={dscore:.2e} rel-L2(x)={rel:.2e}")
    assert rec.dtype == np.float32 and rec.shape == ref.shape and h1 is None and h2 is None
    if name.endswith("tie475"):  # rise_pixel*13 = 47.5: tie views h = +-13, resolved exactly (oracle/make_golden_tie475.py)
        assert info["res"]["flags"] & 16 and not info["res"]["flags"] & 2
    if name == "solve_nn_unb_32":  # twist*h = -30 deg at h = 25: in-plane tie views, resolved by exact per-column maps
        assert not info["res"]["flags"] & 3
    if info["res"]["flags"] & 3:  # tie-flagged geometry: a few samples may land in a neighbouring voxel
        assert dscore <= 2e-4
    else:
        # the bound on x is tied to the reference's OWN reproducibility at its stopping point, measured here: the same
        # equations in permuted row order through the same scipy call (north star 1e-4 where it allows)
        _, _, det = O.lsq_reconstruct(
            img, 1.0, float(twist), float(rise / apix), int(csym), positive_constraint=int(pc),
            reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N,
            reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so), interpolation="nn", return_details=True,
            fast=not name.endswith("tie475"))
        floor = oracle_permutation_floor(det)
        print(f"{name}: the reference's own row-permutation floor at the stopping point: {floor:.2e}")
        assert dscore <= 1e-5
        assert rel <= max(1e-4, 4 * floor)


@pytest.mark.parametrize("name", ["solve_nn_unb_48_t35"])
def test_fixed_iteration_parity_vs_oracle_lsmr(name):
    """Same iteration count => x within the oracle's own float32 round-off growth:
    bit-level agreement early (5 iterations), and within a few times the oracle's
    row-permutation noise floor later (rounding differences are amplified by the
    loss of orthogonality in the Lanczos process)."""
    d = load(name)
    apix, twist, rise, csym, pc, so, L3 = d["args"]
    img = d["image"]
    prob, batch, target = _make_batch(img, float(twist), float(rise / apix), int(csym), int(L3), int(so))
    A_d, b_d, A_s = _oracle_system(img, float(twist), float(rise / apix), int(csym), int(L3), target)
    A = vstack((A_d, A_s)).tocsr()
    b = np.concatenate((b_d, np.zeros(A_s.shape[0], np.float32)))
    relf = lambda a, r: float(np.linalg.norm(a - r) / np.linalg.norm(r))
    for iters in (5, 40):
        x_ref = O.lsmr_mixed(A, b, fixed_iters=iters)[0]
        # the oracle's own reproducibility floor at this iteration count: same maths, rows permuted
        floor = 0.0
        for seed in range(3):
            p = np.random.default_rng(seed).permutation(A.shape[0])
            floor = max(floor, relf(O.lsmr_mixed(A[p].tocsr(), b[p], fixed_iters=iters)[0], x_ref))
        res = batch.solve(fixed_iters=iters, check_every=iters)
        assert res[0]["itn"] == iters
        rel = relf(batch.x(0), x_ref)
        print(f"{name}: fixed {iters} iterations rel-L2(x)={rel:.2e} (oracle row-permutation floor {floor:.2e})")
        assert rel < (2e-6 if iters == 5 else max(1e-4, 4 * floor))
    batch.close(); prob.close()


def test_batched_candidates_equal_single_candidate_solves():
    d = load("solve_nn_unb_48_t35")
    apix, twist, rise, csym, pc, so, L3 = d["args"]
    img = d["image"]
    extra = [(-3.1, 9.0 / apix), (-4.2, 10.1 / apix), (12.5, 9.5 / apix)]
    prob, batch, target = _make_batch(img, float(twist), float(rise / apix), int(csym), int(L3), int(so), extra)
    res = batch.solve()
    xs = [batch.x(c) for c in range(batch.nc)]
    batch.close()
    singles = [(float(twist), float(rise / apix))] + extra
    for c, (tw, ri) in enumerate(singles):
        p2, b2, _ = _make_batch(img, tw, ri, int(csym), int(L3), int(so))
        r2 = b2.solve()
        assert r2[0]["itn"] == res[c]["itn"] and r2[0]["score"] == res[c]["score"]
        assert np.array_equal(b2.x(0), xs[c])  # deterministic: identical bits
        b2.close(); p2.close()
    prob.close()
    assert len({float(r["score"]) for r in res}) == len(res)


def test_reference_test_shapes(solver):
    """Mirrors the structural checks of the reference's tests/test_denovo3D_solver.py:178-260."""
    np.random.seed(42)
    image = np.random.rand(12, 12).astype(np.float32)
    for csym, inner in ((1, 0), (2, 0), (1, 2)):
        (rec3d, h1, h2), score = solver.lsq_reconstruct(
            projection_image=image, scale2d_to_3d=1.0, twist_degree=30, rise_pixel=2, csym=csym,
            reconstruct_diameter_2d_pixel=8, reconstruct_length_2d_pixel=8, reconstruct_diameter_3d_pixel=8,
            reconstruct_diameter_3d_inner_pixel=inner, reconstruct_length_3d_pixel=8, interpolation="nn",
            positive_constraint=0, verbose=0)
        assert isinstance(rec3d, np.ndarray) and rec3d.dtype == np.float32 and rec3d.shape == (8, 8, 8)
        assert h1 is None and h2 is None and np.all(np.isfinite(rec3d))
        assert isinstance(score, (float, np.floating))
    A, b, pid = solver.build_A_data_matrix(
        image=np.eye(8, dtype=np.float32), scale2d_to_3d=1.0, twist_degree=30, rise_pixel=2, csym=1, tilt_degree=0,
        psi_degree=0, dy_pixel=0, reconstruct_diameter_2d_pixel=4, reconstruct_length_2d_pixel=4,
        reconstruct_diameter_3d_pixel=4, reconstruct_diameter_3d_inner_pixel=0, reconstruct_length_3d_pixel=4,
        min_projection_lines=10, interpolation="nn", verbose=0)
    from scipy.sparse import csr_matrix

    assert isinstance(A, csr_matrix) and len(b) == A.shape[0] == len(pid) and A.shape[1] > 0
    # a tilted candidate and trilinear interpolation run on explicit GPU-built rows (tests/test_denovo3D_solver.py:160-176)
    for kw in (dict(tilt_degree=5, interpolation="nn"), dict(interpolation="linear")):
        (rec3d, _, _), score = solver.lsq_reconstruct(
            image, 1.0, 30, 2, reconstruct_diameter_2d_pixel=8, reconstruct_length_2d_pixel=8,
            reconstruct_diameter_3d_pixel=8, reconstruct_length_3d_pixel=8, positive_constraint=0, **kw)
        assert rec3d.shape == (8, 8, 8) and np.all(np.isfinite(rec3d))
    with pytest.raises(NotImplementedError):  # what is still outside the CUDA path fails loudly
        solver.lsq_reconstruct(image, 1.0, 30, 2, reconstruct_diameter_3d_pixel=8, reconstruct_length_3d_pixel=8,
                               algorithm=dict(model="no_such_model"))
    with pytest.raises(ValueError):
        solver.lsq_reconstruct(image, 1.0, 30, 2, reconstruct_diameter_3d_pixel=8, reconstruct_length_3d_pixel=8,
                               score_metric="no_such_metric")


@pytest.mark.gpu
@pytest.mark.parametrize("interp", ["nn", "linear"])
def test_2d_score_metrics_on_the_reprojection(solver, interp):
    """score_metric ssim / ms_ssim / mutual_information / composite (SLR:494-525): the reprojection is scattered back to
    the image by pixel id and compared in 2-D.  The metric implementations restate scikit-image (absent here and in
    the reference's environment of this container -> parity unpinned); pinned here: the scatter equals the oracle's
    prediction image, composite = mean of the four, and the metrics see a near-perfect reprojection."""
    from helicon_b200 import imageprep as M

    d = load("solve_nn_unb_48_t35")
    apix, twist, rise, csym, pc, so, L3 = d["args"]
    img = d["image"]
    N = img.shape[0]
    kw = dict(scale2d_to_3d=1.0, twist_degree=float(twist), rise_pixel=float(rise / apix), csym=int(csym),
              positive_constraint=0, reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
              reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so),
              interpolation=interp)
    (rec_c, _, _), s_cos = solver.lsq_reconstruct(img, **kw)
    vals = {}
    for m in ("ssim", "ms_ssim", "mutual_information", "composite"):
        (rec, _, _), vals[m] = solver.lsq_reconstruct(img, score_metric=m, **kw)
        assert np.array_equal(rec, rec_c)
        assert np.isfinite(vals[m])
    assert 0.5 < vals["ssim"] <= 1 and 0.5 < vals["ms_ssim"] <= 1 and 0 < vals["mutual_information"] <= 1
    assert abs(vals["composite"] - np.mean([float(s_cos), vals["ssim"], vals["ms_ssim"], vals["mutual_information"]])) < 1e-6
    # the oracle's reprojection image through the same metric: inherits only the solve's reproducibility floor
    (rec_o, _, _), _, det = O.lsq_reconstruct(img, return_details=True, **kw)
    pred = det["A_data"].dot(det["res"].x.astype(np.float32))
    D2, L2 = kw["reconstruct_diameter_2d_pixel"], kw["reconstruct_length_2d_pixel"]
    ny, nx = img.shape
    p2 = np.zeros((L2, D2), np.float32)
    p2.ravel()[det["b_pid"]] = pred
    ref2 = img[ny // 2 - D2 // 2: ny // 2 + D2 // 2, nx // 2 - L2 // 2: nx // 2 + L2 // 2].T
    assert abs(M.ssim_score(p2, ref2) - vals["ssim"]) < 2e-3


def _oracle_bounded_spread(img, kw, n_perm=4):
    """Oracle bounded solve + the same solve with the equations permuted: the reference's own
    reproducibility band (float32 LSMR noise is amplified by the data-dependent TRF decisions)."""
    from scipy.optimize import lsq_linear

    (rec_o, _, _), score_o, det = O.lsq_reconstruct(img, return_details=True, **kw)
    A = vstack((det["A_data"], det["A_hsym"])).tocsr()
    b = np.concatenate((det["b_data"], np.zeros(det["A_hsym"].shape[0], np.float32)))
    ub = float(det["b_data"].max())
    x0 = det["res"].x
    sc = lambda x: float(O.cosine_similarity(det["A_data"].dot(x.astype(np.float32)), det["b_data"]))
    nits, dsc, dx = [det["res"].nit], [0.0], [0.0]
    for seed in range(n_perm):
        p = np.random.default_rng(seed).permutation(A.shape[0])
        r2 = lsq_linear(A[p].tocsr(), b[p], bounds=(0.0, ub), tol=1e-2, max_iter=200, lsmr_maxiter=1000, lsmr_tol="auto")
        nits.append(r2.nit); dsc.append(abs(sc(r2.x) - sc(x0))); dx.append(float(np.linalg.norm(r2.x - x0) / np.linalg.norm(x0)))
    return rec_o, float(score_o), nits, max(dsc), max(dx)


BOUNDED_CASES = {
    "golden_48_t35": ("solve_nn_pos_48_t35", None),
    "golden_32": ("solve_nn_pos_32", None),
    "fresh_40": (None, dict(N=40, L3=4, twist=-2.37, rise_px=1.31, so=4, seed=11)),
}


@pytest.mark.parametrize("case", sorted(BOUNDED_CASES))
def test_bounded_solve_within_reference_reproducibility_band(solver, case):
    """positive constraint -> scipy's bounded TRF branch (float64 outer iterations started from the float32 LSMR
    iterate).  The reference does NOT reproduce itself to 1e-5/1e-4 on this path: permuting its equations moves the
    score by 5e-5..2e-2 and x by 1e-3..3e-1 on these cases (the termination test and the 3-way step choice are
    data-dependent).  The CUDA path must land inside that band (x2), respect the bounds, and take a number of outer
    iterations the reference also takes."""
    gname, spec = BOUNDED_CASES[case]
    if gname:
        d = load(gname)
        apix, twist, rise, csym, pc, so, L3 = d["args"]
        img = d["image"]; N = img.shape[0]
        kw = dict(scale2d_to_3d=1.0, twist_degree=float(twist), rise_pixel=float(rise / apix), csym=int(csym),
                  positive_constraint=int(pc), reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
                  reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so),
                  interpolation="nn")
    else:
        rng = np.random.default_rng(spec["seed"]); N = spec["N"]
        yy, xx = np.mgrid[0:N, 0:N]
        img = np.zeros((N, N), np.float32)
        for _ in range(25):
            cy, cx = rng.uniform(0.3 * N, 0.7 * N), rng.uniform(0, N)
            img += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 5.0).astype(np.float32)
        kw = dict(scale2d_to_3d=1.0, twist_degree=spec["twist"], rise_pixel=spec["rise_px"], csym=1, positive_constraint=1,
                  reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N,
                  reconstruct_length_3d_pixel=spec["L3"], sym_oversample=spec["so"], interpolation="nn")
    (rec, _, _), score, info = solver.lsq_reconstruct(img, return_info=True, **kw)
    rec_o, score_o, nits, band_s, band_x = _oracle_bounded_spread(img, kw)
    if gname:
        assert abs(score_o - float(load(gname)["score"])) < 1e-6  # the oracle IS the reference here
    r = info["res"]
    rel = float(np.linalg.norm(rec - rec_o) / np.linalg.norm(rec_o))
    dscore = abs(float(score) - score_o)
    print(f"{case}: gpu lsmr itn={r['itn']} trf_nit={r['trf_nit']} flags={r['flags']} score={float(score):.7f} "
          f"oracle={score_o:.7f} |dscore|={dscore:.2e} (band {band_s:.2e}) rel-L2(x)={rel:.2e} (band {band_x:.2e}) "
          f"oracle nits={nits}")
    assert r["flags"] & 4 and r["trf_nit"] > 0
    assert rec.min() >= 0.0 and rec.max() <= float(img.max()) + 1e-6  # bounds respected
    tie = bool(r["flags"] & 3)
    assert dscore <= max(1e-5, 2 * band_s) * (5 if tie else 1)
    assert rel <= max(2e-3, 2 * band_x) * (5 if tie else 1)
    assert min(nits) - 1 <= r["trf_nit"] <= max(nits) + 1


@pytest.mark.parametrize("env", [dict(HB2_FWD_BAND="0"), dict(HB2_NO_ADJ_TILE="1"), dict(HB2_VOXEL_ORDER="0"),
                                 dict(HB2_VOXEL_ORDER="0", HB2_FWD_BAND="0")])
def test_alternative_kernel_paths_agree_with_default(env, monkeypatch):
    """The default forward band path (TMA-staged voxel bands in shared memory + partial ray sums) against the gather
    kernel, the tile adjoint against the (voxel, quad) fallback, and the band-column-major voxel order against round
    1's row-major tiles, on the same batch: operator applies to float32 round-off, solve scores to 2e-6, stopping
    iteration within 2."""
    d = load("solve_nn_unb_64")
    apix, twist, rise, csym, pc, so, L3 = d["args"]
    img = d["image"]
    extra = [(float(twist) * 1.7, float(rise / apix) * 1.1), (float(twist) * 0.4, float(rise / apix))]
    rng = np.random.default_rng(9)

    def run():
        prob, batch, target = _make_batch(img, float(twist), float(rise / apix), int(csym), int(L3), int(so), extra)
        x = rng.standard_normal(batch.n).astype(np.float32)
        out = []
        for c in range(batch.nc):
            y = batch.apply_forward(c, x)
            g = batch.apply_adjoint(c, y)
            nd_pad, _ = batch.rows_padded(c)
            # symmetry rows in the reference's order (their stored order follows the internal voxel order)
            out.append((np.concatenate([y[:nd_pad], y[nd_pad:][batch.sym_row_order(c)]]), g))
        res = batch.solve()
        xs = [batch.x(c) for c in range(batch.nc)]
        batch.close(); prob.close()
        return out, res.copy(), xs

    rng = np.random.default_rng(9)
    base, res0, x0 = run()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(9)
    alt, res1, x1 = run()
    for (y0, g0), (y1, g1) in zip(base, alt):
        assert np.abs(y0 - y1).max() <= 2e-6 * np.abs(y0).max() * 8
        assert np.abs(g0 - g1).max() <= 2e-6 * np.abs(g0).max() * 8
    # the variants group the partial sums of the norms differently, so alpha/beta may differ in the last bit and the
    # stopping test can fire an iteration earlier or later
    assert np.abs(res0["itn"].astype(int) - res1["itn"].astype(int)).max() <= 2
    assert np.abs(res0["score"] - res1["score"]).max() <= 2e-6
    for a, b in zip(x0, x1):
        # a different summation order moves the loosely converged LSMR iterate like a row permutation of the
        # reference does (2e-4..8e-4 at N=64, more on small cases): same bound as at the stopping point
        assert np.linalg.norm(a - b) <= 5e-3 * np.linalg.norm(a)


def test_float64_band_forward_agrees_with_the_gather_kernel(solver, monkeypatch):
    """Bounded branch: the float64 instantiation of the forward band kernel (default) against the float64 gather kernel
    (HB2_FWD_BAND=1) on the same bounded solve -- float64 sums in another order: same TRF iteration count, x to 1e-9."""
    d = load("solve_nn_unb_64")
    apix, twist, rise, csym, pc, so, L3 = d["args"]
    img = d["image"]
    N = img.shape[0]
    kw = dict(positive_constraint=1, reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
              reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so), return_info=True)
    (r0, _, _), s0, i0 = solver.lsq_reconstruct(img, 1.0, float(twist), float(rise / apix), int(csym), **kw)
    monkeypatch.setenv("HB2_FWD_BAND", "1")
    (r1, _, _), s1, i1 = solver.lsq_reconstruct(img, 1.0, float(twist), float(rise / apix), int(csym), **kw)
    assert i0["res"]["trf_nit"] > 0 and i0["res"]["trf_nit"] == i1["res"]["trf_nit"]
    assert np.linalg.norm(r0 - r1) <= 1e-6 * np.linalg.norm(r1) and abs(float(s0) - float(s1)) <= 1e-6


def test_grid_scores_best_and_topk_vs_reference_grid():
    """North-star criterion on a small grid: 5 twists x 4 rises on a 64x64 filament solved by the REFERENCE
    (tests/golden/grid_64.npz, oracle/make_golden_grid.py) vs ONE batched GPU solve of the 20 candidates:
    scores within 1e-5, identical best (twist, rise), identical top-K ordering (ties closer than 1e-5 may swap)."""
    from helicon_b200.engine import Batch, Problem
    from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec

    d = load("grid_64")
    img, apix, L3, so = d["image"], float(d["apix"]), int(d["L3"]), int(d["sym_oversample"])
    N = img.shape[0]
    prob = Problem(img, 1.0, N, N, N, 0.0, N // 2 - 1)
    target = min(MAX_EQUATIONS, int(max(N * N, L3 * prob.ndisk) * so))
    specs = [CandidateSpec(float(tw), float(ri / apix), 1, target, target, False) for tw in d["twists"] for ri in d["rises"]]
    batch = Batch(prob, L3, specs)
    res = batch.solve()
    batch.close(); prob.close()
    got = res["score"].reshape(d["scores"].shape)
    ref = d["scores"]
    flagged = (res["flags"] & 3) != 0
    print("max |dscore|", float(np.abs(got - ref).max()), "flagged", int(flagged.sum()))
    assert np.abs(got - ref)[~flagged.reshape(ref.shape)].max() <= 1e-5
    assert np.abs(got - ref).max() <= 2e-4
    assert np.unravel_index(np.argmax(got), got.shape) == np.unravel_index(np.argmax(ref), ref.shape)
    order_ref = np.argsort(-ref.ravel(), kind="stable")
    order_got = np.argsort(-got.ravel(), kind="stable")
    for k in range(10):  # top-10: same candidate unless the reference's own scores are closer than 1e-5 at that rank
        if order_ref[k] != order_got[k]:
            assert abs(ref.ravel()[order_ref[k]] - ref.ravel()[order_got[k]]) <= 1e-5, k


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["fsc_mode1_48", "fsc_mode2_48", "fsc_mode3_48_c2", "fsc_mode4_32"])
def test_half_set_solves_vs_reference_golden(solver, name):
    """lsq_reconstruct(fsc_test=1..4): full + two half-set solves as three masked candidates of one batch."""
    d = load(name)
    apix, twist, rise, csym, pc, so, L3, mode, seed = d["args"]
    img = d["image"]
    N = img.shape[0]
    np.random.seed(int(seed))
    (rec, h1, h2), score, info = solver.lsq_reconstruct(
        img, 1.0, float(twist), float(rise / apix), int(csym), positive_constraint=int(pc),
        reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N,
        reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so), interpolation="nn", fsc_test=int(mode),
        return_info=True)
    relf = lambda a, r: float(np.linalg.norm(a - r) / np.linalg.norm(r))
    rels = [relf(rec, d["rec3d"]), relf(h1, d["half1"]), relf(h2, d["half2"])]
    dscore = abs(float(score) - float(d["score"]))
    print(f"{name}: itn={[int(r['itn']) for r in info['all_res']]} |dscore|={dscore:.2e} rel-L2={['%.1e' % r for r in rels]}")
    # the halves keep disjoint, complementary row sets: zero support where the other half has its data
    assert h1.shape == rec.shape == h2.shape and h1.dtype == np.float32
    if info["res"]["flags"] & 3:
        assert dscore <= 2e-4
    else:
        assert dscore <= 1e-5
        assert max(rels) < 5e-3


GEN_DATA = ["gen_data_nn_tilt", "gen_data_nn_c2_stop", "gen_data_nn_s05_dy", "gen_data_lin_tilt", "gen_data_lin_psi_inner",
            "data_lin_a", "data_lin_csym2"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", GEN_DATA)
def test_explicit_rows_vs_reference(solver, name):
    """build_A_data_matrix outside the grid-search case (tilt/psi/dy != 0 and/or trilinear interpolation): rows built
    on the GPU (k_exp_rows) -- same row set, b, pixel ids; nn pattern bit-exact, trilinear weights to 1e-6."""
    d = load(name)
    a = d["args"]
    s, twist, rise, csym, D2, L2, D3, D3i, L3, mpl = a[:10]
    tilt, psi, dy = (a[10:13] if len(a) > 10 else (0.0, 0.0, 0.0))
    linear = bool(int(d["linear"])) if "linear" in d else name.startswith("data_lin")
    A, b, pid = solver.build_A_data_matrix(
        image=d["image"], scale2d_to_3d=float(s), twist_degree=float(twist), rise_pixel=float(rise), csym=int(csym),
        tilt_degree=float(tilt), psi_degree=float(psi), dy_pixel=float(dy), reconstruct_diameter_2d_pixel=int(D2),
        reconstruct_length_2d_pixel=int(L2), reconstruct_diameter_3d_pixel=int(D3),
        reconstruct_diameter_3d_inner_pixel=int(D3i), reconstruct_length_3d_pixel=int(L3), min_projection_lines=int(mpl),
        interpolation="linear" if linear else "nn")
    ref = csr_from(d)
    ok, why = csr_equal(A, ref, tol=1e-6 if linear else 0.0)
    print(f"{name}: rows gpu={A.shape[0]} ref={ref.shape[0]} nnz gpu={A.nnz} ref={ref.nnz} identical={ok} {why}")
    assert ok, why
    assert np.array_equal(b, d["b"]) and b.dtype == np.float32
    assert np.array_equal(pid, d["b_pid"]) and pid.dtype == np.int32


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gen_solve_nn_tilt_48", "gen_solve_nn_dy_32", "gen_solve_nn_tilt_48_pos"])
def test_general_orientation_solve_vs_reference_golden(solver, name):
    """lsq_reconstruct with tilt/psi/dy != 0: explicit rows + the batch's LSMR / bounded branch / score."""
    d = load(name)
    apix, twist, rise, csym, pc, so, L3, tilt, psi, dy = d["args"]
    img = d["image"]
    N = img.shape[0]
    (rec, h1, h2), score, info = solver.lsq_reconstruct(
        img, 1.0, float(twist), float(rise / apix), int(csym), tilt_degree=float(tilt), psi_degree=float(psi),
        dy_pixel=float(dy), positive_constraint=int(pc), reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
        reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so),
        interpolation="nn", return_info=True)
    ref = d["rec3d"]
    rel = float(np.linalg.norm(rec - ref) / np.linalg.norm(ref))
    dscore = abs(float(score) - float(d["score"]))
    r = info["res"]
    print(f"{name}: itn={r['itn']} istop={r['istop']} trf_nit={r['trf_nit']} flags={r['flags']} score={float(score):.7f} "
          f"ref={float(d['score']):.7f} |dscore|={dscore:.2e} rel-L2(x)={rel:.2e}")
    assert rec.shape == ref.shape and rec.dtype == np.float32
    if int(pc) > 0:  # bounded branch: inside the reference's own reproducibility band (see the bounded-solve test)
        assert r["flags"] & 4 and dscore <= 2e-3 and rel < 5e-2
    else:
        assert dscore <= 1e-5 and rel < 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hsym_lin_a", "gen_hsym_lin_c2_stop", "gen_hsym_lin_inner", "gen_hsym_lin_tie"])
def test_trilinear_symmetry_rows_vs_reference(solver, name):
    """build_A_helical_sym_matrix(interpolation="linear"): rows built on the GPU (k_lsym_*), same rows in the same
    order, weights to 1e-6 (float64 products cast to float32, incl. the reference's corner-110 expression)."""
    d = load(name)
    nz, ny, nx, twist, rise, csym, rmin, rmax, msp = d["args"]
    A, b = solver.build_A_helical_sym_matrix(int(nz), int(ny), int(nx), float(twist), float(rise), int(csym), float(rmin),
                                             float(rmax), int(msp), "linear")
    ref = csr_from(d)
    ok, why = csr_equal(A, ref, tol=1e-6)
    print(f"{name}: rows gpu={A.shape[0]} ref={ref.shape[0]} nnz gpu={A.nnz} ref={ref.nnz} identical={ok} {why}")
    assert ok, why
    assert b.shape == (ref.shape[0],) and not b.any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gen_solve_lin_48", "gen_solve_lin_40_c2_tilt", "gen_solve_lin_48_pos"])
def test_trilinear_solve_vs_reference_golden(solver, name):
    """lsq_reconstruct(interpolation="linear"), also tilted: explicit trilinear data + symmetry rows, LSMR / bounded
    branch / score of the batch."""
    d = load(name)
    apix, twist, rise, csym, pc, so, L3, tilt, psi, dy = d["args"]
    img = d["image"]
    N = img.shape[0]
    (rec, h1, h2), score, info = solver.lsq_reconstruct(
        img, 1.0, float(twist), float(rise / apix), int(csym), tilt_degree=float(tilt), psi_degree=float(psi),
        dy_pixel=float(dy), positive_constraint=int(pc), reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
        reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so),
        interpolation="linear", return_info=True)
    ref = d["rec3d"]
    rel = float(np.linalg.norm(rec - ref) / np.linalg.norm(ref))
    dscore = abs(float(score) - float(d["score"]))
    r = info["res"]
    print(f"{name}: itn={r['itn']} istop={r['istop']} trf_nit={r['trf_nit']} flags={r['flags']} score={float(score):.7f} "
          f"ref={float(d['score']):.7f} |dscore|={dscore:.2e} rel-L2(x)={rel:.2e}")
    assert rec.shape == ref.shape and rec.dtype == np.float32
    # the reference's own reproducibility band on this system (same scipy solve, equations permuted; measured by
    # oracle/make_golden_band.py): the trilinear systems are ill-conditioned enough that float32 LSMR noise moves
    # the stopped iterate by more than the north-star tolerances, so the bar is max(tolerance, 2 x band)
    band_s, band_x, band_it = float(d["band_dscore"]), float(d["band_relx"]), d["band_itn"]
    print(f"    reference band: |dscore| {band_s:.2e} rel-L2 {band_x:.2e} iterations {band_it.tolist()}")
    assert dscore <= max(1e-5, 2 * band_s) and rel <= max(5e-3, 2 * band_x)
    it = r["trf_nit"] if int(pc) > 0 else r["itn"]
    # the number of TRF iterations is as chaotic as the iterate (data-dependent step kinds) and 5 permutations sample the
    # reference's own spread only coarsely: the bound is a factor, not +-2
    assert band_it.min() // 2 <= it <= 2 * band_it.max()
    if int(pc) > 0:
        assert r["flags"] & 4 and rec.min() >= 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["fsc_lin_mode2_40", "fsc_tilt_mode3_48", "fsc_tilt_mode1_48"])
def test_half_sets_on_explicit_rows_vs_reference_golden(solver, name):
    """fsc_test on the explicit-row paths (trilinear / tilted): full + two half-set solves, the builder drops the other
    half's rows after the early stop (oracle/make_golden_fsc_explicit.py)."""
    d = load(name)
    apix, twist, rise, csym, so, L3, tilt, psi, dy, mode, seed = d["args"]
    img = d["image"]
    N = img.shape[0]
    linear = bool(int(d["linear"]))
    np.random.seed(int(seed))
    (rec, h1, h2), score, info = solver.lsq_reconstruct(
        img, 1.0, float(twist), float(rise / apix), int(csym), tilt_degree=float(tilt), psi_degree=float(psi),
        dy_pixel=float(dy), positive_constraint=0, reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
        reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so),
        interpolation="linear" if linear else "nn", fsc_test=int(mode), return_info=True)
    relf = lambda a, r: float(np.linalg.norm(a - r) / np.linalg.norm(r))
    rels = [relf(rec, d["rec3d"]), relf(h1, d["half1"]), relf(h2, d["half2"])]
    dscore = abs(float(score) - float(d["score"]))
    print(f"{name}: itn={[int(r['itn']) for r in info['all_res']]} |dscore|={dscore:.2e} rel-L2={['%.1e' % r for r in rels]}")
    # trilinear systems carry the reference's own permutation band (see test_trilinear_solve_vs_reference_golden)
    assert dscore <= (1e-4 if linear else 1e-5)
    assert max(rels) <= (3e-2 if linear else 5e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["refine_dy_40", "refine_dy_32_pos"])
def test_refine_tilt_psi_dy_vs_reference_golden(solver, name):
    """refine_tilt_psi_dy (SLR:550-841): the reference's Gauss-Newton loop with every build / prediction / solve on the
    GPU.  The reference solves with LSQR(1e-6) -- here scipy's LSQR over the CUDA operator -- and, when the positive rule
    fires, lsq_linear at 1e-10 -- here the batch's LSMR + TRF state machines; that case is compared at the level of the
    refinement's own convergence thresholds (tol_tilt = 0.05 deg, tol_psi = 0.1 deg, tol_dy = 0.05 px)."""
    d = load(name)
    apix, twist, rise, csym, L3, so, pc, mi = d["args"]
    img = d["image"]
    N = img.shape[0]
    tilt, psi, dy, x, score = solver.refine_tilt_psi_dy(
        projection_image=img, scale2d_to_3d=1.0, twist_degree=float(twist), rise_pixel=float(rise / apix), csym=int(csym),
        reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N,
        reconstruct_diameter_3d_inner_pixel=0, reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so),
        interpolation="nn", x_init=None, max_iter=int(mi), positive_constraint=int(pc), verbose=0)
    rt, rp, rdy, rs = d["out"]
    relx = float(np.linalg.norm(x - d["x"]) / np.linalg.norm(d["x"]))
    print(f"{name}: tilt {tilt:.5f} ({rt:.5f}) psi {psi:.5f} ({rp:.5f}) dy {dy:.5f} ({rdy:.5f}) score {score:.6f} ({rs:.6f}) "
          f"rel-L2(x) {relx:.2e}")
    assert isinstance(tilt, float) and isinstance(psi, float) and isinstance(dy, float) and isinstance(score, float)
    assert isinstance(x, np.ndarray) and np.all(np.isfinite(x))
    assert abs(tilt - rt) <= 0.05 and abs(psi - rp) <= 0.1 and abs(dy - rdy) <= 0.05
    assert abs(score - rs) <= 2e-3 and relx <= 5e-2
    if int(pc) == 0:
        # unbounded systems: scipy's LSQR over the CUDA operator, as the reference runs it (lsqr.py) -- measured
        # parameters equal to 1e-7, score to 1e-7, rel-L2(x) 3.5e-5 (the device-resident LSMR of round 1: ~1e-3)
        assert abs(tilt - rt) <= 1e-4 and abs(psi - rp) <= 1e-4 and abs(dy - rdy) <= 1e-4
        assert abs(score - rs) <= 1e-5 and relx <= 1e-3


@pytest.mark.gpu
def test_lsq_reconstruct_with_refine_range(solver):
    """lsq_reconstruct(refine_tilt_psi_dy_range=...) (SLR:372-437): base solve, refinement, the refined one adopted."""
    d = load("refine_dy_40")
    apix, twist, rise, csym, L3, so, pc, mi = d["args"]
    img = d["image"]
    N = img.shape[0]
    kw = dict(positive_constraint=int(pc), reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N,
              reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so), interpolation="nn")
    (rec0, _, _), s0 = solver.lsq_reconstruct(img, 1.0, float(twist), float(rise / apix), int(csym), **kw)
    (rec1, h1, h2), s1 = solver.lsq_reconstruct(img, 1.0, float(twist), float(rise / apix), int(csym),
                                               refine_tilt_psi_dy_range=dict(tilt=5.0, psi=5.0, dy=2.0, max_iter=2), **kw)
    assert rec1.shape == rec0.shape and h1 is None and h2 is None and np.isfinite(rec1).all()
    assert np.isfinite(float(s1)) and hasattr(solver.lsq_reconstruct, "_refined_params")


@pytest.mark.gpu
@pytest.mark.parametrize("interp,tilt", [("nn", 0.0), ("linear", 0.0), ("nn", 2.5)])
def test_nonsquare_image_with_cropped_region_vs_oracle(solver, interp, tilt):
    """A non-square image (40 x 56) with a cropped reconstruction region (D2 = 32, L2 = 48), csym 2, an inner diameter:
    the data rows of all three paths (matrix-free maps, explicit trilinear, explicit tilted) and the solve of the
    nearest-neighbour one against the oracle run in the test."""
    rng = np.random.default_rng(21)
    img = rng.random((40, 56)).astype(np.float32)
    kw = dict(scale2d_to_3d=1.0, twist_degree=17.3, rise_pixel=2.9, csym=2, tilt_degree=tilt, psi_degree=0.0, dy_pixel=0.0,
              reconstruct_diameter_2d_pixel=32, reconstruct_length_2d_pixel=48, reconstruct_diameter_3d_pixel=32,
              reconstruct_diameter_3d_inner_pixel=6, reconstruct_length_3d_pixel=8, min_projection_lines=10**7,
              interpolation=interp)
    A, b, pid = solver.build_A_data_matrix(image=img, **kw)
    Ao, bo, pido = O.build_A_data_matrix(img, 1.0, 17.3, 2.9, 2, tilt, 0.0, 0.0, 32, 48, 32, 6, 8, 10**7, interp)
    ok, why = csr_equal(A, Ao, tol=1e-6 if interp == "linear" else 0.0)
    assert ok, why
    assert np.array_equal(b, bo) and np.array_equal(pid, pido)
    if interp == "nn" and tilt == 0.0:
        skw = dict(scale2d_to_3d=1.0, twist_degree=17.3, rise_pixel=2.9, csym=2, positive_constraint=0,
                   reconstruct_diameter_3d_inner_pixel=6, reconstruct_diameter_2d_pixel=32, reconstruct_length_2d_pixel=48,
                   reconstruct_diameter_3d_pixel=32, reconstruct_length_3d_pixel=8, sym_oversample=2, interpolation="nn")
        from tests.helpers import oracle_permutation_floor

        (rec, _, _), score = solver.lsq_reconstruct(img, **skw)
        (rec_o, _, _), score_o, det = O.lsq_reconstruct(img, return_details=True, **skw)
        rel = float(np.linalg.norm(rec - rec_o) / np.linalg.norm(rec_o))
        floor = oracle_permutation_floor(det)
        print(f"non-square solve: score {float(score):.7f} vs {float(score_o):.7f}, rel-L2 {rel:.2e} "
              f"(the oracle's own row-permutation floor {floor:.2e})")
        assert abs(float(score) - float(score_o)) <= 1e-5 and rel <= max(1e-4, 4 * floor)


def test_threaded_lsq_reconstruct_equals_serial(solver):
    """The reference calls the solver from a ThreadPoolExecutor (app.py:2473); ctypes releases the GIL, so several
    batches are alive on the default stream at once.  Only one of them may own the stream's arena (two live
    batches must never be handed the same device range): results must equal the serial run bit for bit."""
    from concurrent.futures import ThreadPoolExecutor

    d = load("solve_nn_unb_48_t35")
    apix, twist, rise, csym, pc, so, L3 = d["args"]
    img = d["image"]
    N = img.shape[0]
    cands = [(float(twist) + 0.37 * i, float(rise / apix) * (1 + 0.03 * (i % 3)), int(pc) if i % 2 else 1) for i in range(8)]

    def run(c):
        tw, ri, pcc = c
        (rec, _, _), score = solver.lsq_reconstruct(
            img, 1.0, tw, ri, int(csym), positive_constraint=pcc, reconstruct_diameter_2d_pixel=N,
            reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N, reconstruct_length_3d_pixel=int(L3),
            sym_oversample=int(so), interpolation="nn")
        return rec, float(score)

    serial = [run(c) for c in cands]
    for rep in range(2):
        with ThreadPoolExecutor(max_workers=4) as ex:
            par = list(ex.map(run, cands))
        for (r0, s0), (r1, s1) in zip(serial, par):
            assert s0 == s1 and np.array_equal(r0, r1)


def test_lsq_reconstruct_refine_adoption_vs_reference_golden(solver):
    """oracle/make_golden_refine_adopt.py: for model 'lsq' the reference ALWAYS adopts the refined solution (its first
    solve returns score=None, SLR:270/424) -- here the refined score (0.98415) is LOWER than the base score (0.98655)
    and is still the one returned, with _refined_params set."""
    d = load("refine_adopt_40")
    apix, twist, rise, csym, L3, so, pc, mi = d["args"]
    img = d["image"]
    N = img.shape[0]
    tr, pr, dr, mit = d["range"]
    if hasattr(solver.lsq_reconstruct, "_refined_params"):
        del solver.lsq_reconstruct._refined_params
    (rec, h1, h2), score = solver.lsq_reconstruct(
        img, 1.0, float(twist), float(rise / apix), int(csym), positive_constraint=int(pc),
        reconstruct_diameter_2d_pixel=N, reconstruct_length_2d_pixel=N, reconstruct_diameter_3d_pixel=N,
        reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so), interpolation="nn",
        refine_tilt_psi_dy_range=dict(tilt=float(tr), psi=float(pr), dy=float(dr), max_iter=int(mit)))
    rp = solver.lsq_reconstruct._refined_params
    got = np.array([rp["tilt"], rp["psi"], rp["dy"]])
    rel = float(np.linalg.norm(rec - d["rec3d"]) / np.linalg.norm(d["rec3d"]))
    print(f"refine adoption: score {float(score):.6f} ref {float(d['score']):.6f} (base {float(d['score_base']):.6f}) "
          f"params {got} ref {d['refined']} rel-L2 {rel:.2e}")
    assert float(d["score"]) < float(d["score_base"])           # the golden really pins the 'lower score adopted' branch
    assert abs(float(score) - float(d["score"])) <= 2e-4       # LSMR here vs LSQR there at 1e-6
    assert np.all(np.abs(got - d["refined"]) <= 2e-3)
    assert rel <= 2e-2
