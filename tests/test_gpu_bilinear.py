"""Matrix-free trilinear rows (helicon_b200.bilinear, csrc/hb2_bilinear.cuh) against the explicit GPU-built rows
(engine.ExplicitBatch), which are pinned to the reference's matrices (tests/test_gpu_parity.py:
test_explicit_rows_vs_reference, test_trilinear_symmetry_rows_vs_reference): same row set, right-hand side and pixel
ids; the operator and its transpose equal to float32 round-off; batched solves equal to the single-candidate explicit
solves within the reference's own reproducibility band."""
import numpy as np
import pytest


def _image(N, seed=5):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:N, 0:N]
    img = np.zeros((N, N), np.float32)
    for _ in range(14):
        cy, cx = rng.uniform(N * 0.3, N * 0.7), rng.uniform(1, N - 1)
        img += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 5.0).astype(np.float32)
    return img


# N, L3, twist, rise_pixel, csym, min_projection_lines, min_sym_pairs, inner diameter
CASES = {
    "generic": (40, 6, 23.7, 2.31, 1, 0, 30000, 0),
    "early_stop": (40, 6, -31.3, 1.77, 1, 4000, 9000, 0),
    "integer_rise": (36, 8, 17.2, 3.0, 1, 0, 20000, 0),
    "half_integer_rise_c2": (36, 5, 41.0, 2.5, 2, 0, 20000, 0),
    "twist90_c4": (32, 4, 90.0, 1.9, 4, 0, 15000, 0),
    "inner_L12": (48, 12, -1.2, 3.6538461538461537, 1, 0, 40000, 10),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_matrix_free_trilinear_rows_equal_explicit_rows(name):
    from helicon_b200.bilinear import BilinearBatch
    from helicon_b200.engine import ExplicitBatch, Problem
    from helicon_b200.planner import CandidateSpec

    N, L3, twist, rise, csym, mpl, msp, inner = CASES[name]
    img = _image(N)
    prob = Problem(img, 1.0, N, N, N, inner / 2, N // 2 - 1, interpolation="linear")  # row-major tiles, as lsq_reconstruct / search_grid create it
    spec = CandidateSpec(twist, rise, csym, mpl, msp, False)
    other = CandidateSpec(twist * 0.93 + 1.0, rise * 1.07, 1, mpl, msp, False)
    eb = ExplicitBatch(prob, L3, spec, interpolation="linear")
    bb = BilinearBatch(prob, L3, [other, spec])  # the candidate under test is NOT the first of its batch
    c = 1
    try:
        A, b, pid = eb.data_csr(0)
        idx, kk, jj = bb.data_row_index(c)
        print(f"{name}: rows explicit={A.shape[0]} matrix-free={len(idx)} maps={len(bb.maps)} (regular {bb.n_regular_maps}) "
              f"views={int(bb.cand_nview[c])} sym rows={int(bb.m_sym[c])}")
        assert len(idx) == A.shape[0]
        assert np.array_equal((kk * N + jj).astype(np.int32), pid)
        nd, tot = bb.rows_padded(c)
        rhs = bb.rhs_padded(c)
        assert np.array_equal(rhs[idx], b)
        rest = np.ones(nd, dtype=bool)
        rest[idx] = False
        assert not rhs[rest].any()
        assert int(bb.plan.cand_n_data_rows[c]) == A.shape[0]
        As, _ = eb.sym_csr(0)
        Bs, _ = bb.sym_csr(c)
        ms = 0 if As is None else As.shape[0]
        assert int(bb.m_sym[c]) == ms
        if ms:
            assert abs(As - Bs).max() == 0.0
        rng = np.random.default_rng(1)
        x = rng.normal(size=bb.n).astype(np.float32)
        y = bb.apply_forward(c, x)
        yref = A @ x
        scale = float(np.abs(yref).max())
        err = float(np.abs(y[idx] - yref).max()) / scale
        off = bb.sym_row_offset(c)
        assert not y[:nd][rest[:nd] & (np.arange(nd) < off)].any()
        errs = 0.0
        if ms:
            ys = As @ x
            errs = float(np.abs(y[off:off + ms] - ys).max()) / float(np.abs(ys).max())
        # transpose
        yv = rng.normal(size=tot).astype(np.float32)
        xa = bb.apply_adjoint(c, yv)
        xref = A.T @ yv[idx] + (As.T @ yv[off:off + ms] if ms else 0.0)
        erra = float(np.abs(xa - xref).max()) / float(np.abs(xref).max())
        print(f"    forward rel err {err:.2e} (sym rows {errs:.2e}), adjoint rel err {erra:.2e}")
        assert err < 2e-6 and errs < 2e-6 and erra < 2e-6
    finally:
        eb.close(); bb.close(); prob.close()


@pytest.mark.gpu
@pytest.mark.parametrize("positive", [False, True])
def test_matrix_free_trilinear_batched_solve_equals_explicit_solves(positive):
    from helicon_b200.bilinear import BilinearBatch
    from helicon_b200.engine import ExplicitBatch, Problem
    from helicon_b200.planner import CandidateSpec

    N, L3 = 48, 8
    img = _image(N, seed=9)
    prob = Problem(img, 1.0, N, N, N, 0.0, N // 2 - 1)  # band-column-major voxel order (the nearest-neighbour default)
    specs = [CandidateSpec(tw, rs, 1, 0, 30000, positive) for tw, rs in ((-2.1, 2.9), (-1.4, 3.3), (-3.0, 3.0))]
    bb = BilinearBatch(prob, L3, specs)
    try:
        res = bb.solve()
        for c, sp in enumerate(specs):
            eb = ExplicitBatch(prob, L3, sp, interpolation="linear")
            try:
                r1 = eb.solve()
                x0, x1 = eb.x(0), bb.x(c)
                rel = float(np.linalg.norm(x1 - x0) / np.linalg.norm(x0))
                ds = abs(float(res["score"][c]) - float(r1["score"][0]))
                print(f"cand {c}: itn {int(res['itn'][c])} / {int(r1['itn'][0])} trf {int(res['trf_nit'][c])} / {int(r1['trf_nit'][0])} "
                      f"score {float(res['score'][c]):.7f} / {float(r1['score'][0]):.7f} |d|={ds:.2e} rel-L2(x)={rel:.2e}")
                assert int(res["n_data_rows"][c]) == eb.m_rows
                assert abs(int(res["itn"][c]) - int(r1["itn"][0])) <= max(3, int(r1["itn"][0]) // 10)
                # two float32 LSMR runs on operators that agree to round-off: inside the reference's own reproducibility band
                # for trilinear systems (oracle/make_golden_band.py: |dscore| 2e-5 ... 1.2e-4 unbounded, 4.6e-2 bounded)
                # (bounded: the TRF step kinds are data-dependent and the iterate is chaotic on these ill-conditioned systems --
                # the reference's own band is 4.6e-2 in score; the reference goldens gen_solve_lin_48_pos / task_linear pin it)
                # (unbounded: with the Halton-duplicate views computed separately or not, |dscore| moves between 6e-5 and 3e-4
                # on this noisy test image -- two LSMR runs with different float32 summation orders)
                assert ds <= (5e-2 if positive else 1e-3) and rel <= (0.5 if positive else 2e-2)
            finally:
                eb.close()
    finally:
        bb.close(); prob.close()


@pytest.mark.gpu
def test_matrix_free_trilinear_fixed_iterations_equal_explicit_rows():
    """The same 5 LSMR iterations on the matrix-free operator and on the explicit rows (solver modes of the kernels: 1/beta
    scaling, duplicate views, partial norms): x agrees to float32 round-off.  (Later iterates of these ill-conditioned
    trilinear systems amplify any change of summation order: at 30 iterations the explicit path itself moves by 5e-4 in
    rel-L2 between its two norm modes, and the reference's converged solution by 2e-2 when its equations are permuted --
    tests/golden/gen_solve_lin_48.npz band_relx; profiles/diag_bilinear.py.)"""
    from helicon_b200.bilinear import BilinearBatch
    from helicon_b200.engine import ExplicitBatch, Problem
    from helicon_b200.planner import CandidateSpec

    N, L3 = 48, 8
    img = _image(N, seed=9)
    prob = Problem(img, 1.0, N, N, N, 0.0, N // 2 - 1, interpolation="linear")
    specs = [CandidateSpec(tw, rs, 1, 0, 30000, False) for tw, rs in ((-2.1, 2.9), (-1.4, 3.3))]
    bb = BilinearBatch(prob, L3, specs)
    try:
        res = bb.solve(fixed_iters=5, check_every=5)
        for c, sp in enumerate(specs):
            eb = ExplicitBatch(prob, L3, sp, interpolation="linear")
            try:
                r1 = eb.solve(fixed_iters=5, check_every=5)
                x0, x1 = eb.x(0), bb.x(c)
                rel = float(np.linalg.norm(x1 - x0) / np.linalg.norm(x0))
                ds = abs(float(res["score"][c]) - float(r1["score"][0]))
                print(f"cand {c}: 5 iterations, score {float(res['score'][c]):.7f} / {float(r1['score'][0]):.7f} |d|={ds:.2e} rel-L2(x)={rel:.2e}")
                assert int(res["itn"][c]) == int(r1["itn"][0]) == 5
                assert ds <= 2e-6 and rel <= 1e-5
            finally:
                eb.close()
    finally:
        bb.close(); prob.close()


@pytest.mark.gpu
def test_matrix_free_trilinear_tile_adjoint_agrees_with_gather_adjoint(monkeypatch):
    """Default adjoint (k_adj_bil_tile: map runs, weights and un-blended row windows staged in shared memory by TMA) against
    the plain gather kernel (k_adj_bil) on the same batch: operator applies to float32 round-off, 5 fixed LSMR iterations
    to 1e-5."""
    from helicon_b200.bilinear import BilinearBatch
    from helicon_b200.engine import Problem
    from helicon_b200.planner import CandidateSpec

    N, L3 = 64, 12
    img = _image(N, seed=11)
    specs = [CandidateSpec(tw, rs, cs, 0, 60000, False) for tw, rs, cs in ((-1.9, 3.65, 1), (27.3, 3.1, 1), (58.0, 3.5, 2))]

    def run():
        prob = Problem(img, 1.0, N, N, N, 0.0, N // 2 - 1, interpolation="linear")
        bb = BilinearBatch(prob, L3, specs)
        rng = np.random.default_rng(4)
        x = rng.normal(size=bb.n).astype(np.float32)
        out = []
        for c in range(bb.nc):
            y = bb.apply_forward(c, x)
            out.append((y, bb.apply_adjoint(c, y)))
        res = bb.solve(fixed_iters=5, check_every=5)
        xs = [bb.x(c) for c in range(bb.nc)]
        bb.close(); prob.close()
        return out, res.copy(), xs

    base, res0, x0 = run()
    monkeypatch.setenv("HB2_NO_BIL_TILE", "1")
    alt, res1, x1 = run()
    for c, ((y0, g0), (y1, g1)) in enumerate(zip(base, alt)):
        ey = float(np.abs(y0 - y1).max() / np.abs(y0).max()); eg = float(np.abs(g0 - g1).max() / np.abs(g0).max())
        rel = float(np.linalg.norm(x0[c] - x1[c]) / np.linalg.norm(x1[c]))
        print(f"cand {c}: forward {ey:.2e} adjoint {eg:.2e}; 5 iterations rel-L2(x) {rel:.2e} |dscore| {abs(float(res0['score'][c]) - float(res1['score'][c])):.2e}")
        assert ey < 2e-6 and eg < 2e-6 and rel < 1e-5
