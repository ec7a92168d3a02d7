"""Full-size property tests (BASELINE.json configs 2, 3 and 5 shapes): the oracle cannot run these sizes in seconds, so
the CUDA path is checked through size-independent properties of the operator it applies matrix-free --
<A x, y> = <x, A^T y> (forward and adjoint kernels are transposes of each other, incl. symmetry rows), padded rows
stay zero, a batched solve equals single-candidate solves, scores are finite and in (0, 1]."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _image(N, seed=3):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:N, 0:N]
    img = np.zeros((N, N), np.float32)
    for _ in range(40):
        cy, cx = rng.uniform(0.3 * N, 0.7 * N), rng.uniform(0, N)
        img += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * (N / 60) ** 2)).astype(np.float32)
    return img


CONFIGS = [
    # name, N, apix, [(twist, rise, csym)], fixed iterations
    ("cfg2_256", 256, 1.3, [(-1.2, 4.75, 1), (-2.93, 4.43, 1)], 6),
    ("cfg3_512_csym", 512, 1.3, [(-3.1, 4.8, 1), (27.3, 9.1, 3), (-58.7, 19.3, 6)], 4),
    ("cfg5_384_large_rise", 384, 1.3, [(-137.4, 21.7, 1), (55.1, 41.3, 1)], 4),
]


@pytest.mark.parametrize("name,N,apix,cands,iters", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_adjoint_identity_and_batched_solve_at_full_size(name, N, apix, cands, iters):
    from helicon_b200.engine import Batch, Problem
    from helicon_b200.grid import build_tasks
    from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec

    img = _image(N)
    rng = np.random.default_rng(1)
    by_l3 = {}
    for tw, ri, cs in cands:
        tasks, _ = build_tasks(N, N, apix, [tw], [ri], csyms=(cs,), reconstruct_length_rise=3)
        t = tasks[0]
        by_l3.setdefault((t.geom["L3"], t.geom["D2"], t.geom["L2"], t.geom["D3"], t.geom["s"]), []).append(t)
    for (L3, D2, L2, D3, s), tl in by_l3.items():
        prob = Problem(img, s, D2, L2, D3, 0.0, D3 // 2 - 1)
        n3 = L3 * prob.ndisk
        specs = []
        for t in tl:
            target = min(MAX_EQUATIONS, int(max(D2 * L2, n3) * t.geom["sym_oversample"]))
            specs.append(CandidateSpec(t.twist, t.rise / t.geom["apix3d"], t.csym, target, target, False))
        batch = Batch(prob, L3, specs)
        for c in range(batch.nc):
            nd_pad, tot = batch.rows_padded(c)
            x = rng.standard_normal(batch.n).astype(np.float32)
            y = batch.apply_forward(c, x)
            pidx, _, _ = batch.data_row_index(c)
            pad_only = np.ones(nd_pad, bool)
            pad_only[pidx] = False
            assert not y[:nd_pad][pad_only].any()
            u = np.zeros(tot, np.float32)
            u[:nd_pad][pidx] = rng.standard_normal(len(pidx)).astype(np.float32)
            u[nd_pad:] = rng.standard_normal(tot - nd_pad).astype(np.float32)
            g = batch.apply_adjoint(c, u)
            lhs = float(np.dot(y.astype(np.float64), u.astype(np.float64)))
            rhs = float(np.dot(x.astype(np.float64), g.astype(np.float64)))
            assert abs(lhs - rhs) <= 2e-5 * max(abs(lhs), abs(rhs), 1.0), (name, c, lhs, rhs)
        res = batch.solve(fixed_iters=iters, check_every=iters)
        assert np.all(res["itn"] == iters) and np.all(np.isfinite(res["score"]))
        assert np.all(res["score"] > 0) and np.all(res["score"] <= 1.0 + 1e-6)
        scores = res["score"].copy()
        x_first = batch.x(0)
        batch.close()
        # the same first candidate alone gives the same iterate and score (batching does not couple candidates)
        single = Batch(prob, L3, specs[:1])
        r1 = single.solve(fixed_iters=iters, check_every=iters)
        assert abs(float(r1["score"][0]) - float(scores[0])) <= 1e-6
        assert np.array_equal(single.x(0), x_first)
        single.close()
        prob.close()
        print(f"{name}: L3={L3} n={n3} cands={len(specs)} scores={scores}")


@pytest.mark.parametrize("N,interp,tilt", [(200, "linear", 0.0), (256, "nn", 4.0)], ids=["cfg1_200_trilinear", "cfg2_256_tilted"])
def test_explicit_rows_operator_at_full_size(N, interp, tilt):
    """Explicit GPU-built rows at BASELINE sizes (trilinear at the cfg1 shape, a tilted candidate at the cfg2 shape):
    the explicit operator must equal the exported CSR applied on the host (float32 round-off), its transpose kernel
    must be its adjoint, rows outside the explicit rows stay zero, a short solve gives a finite score."""
    from helicon_b200.engine import ExplicitBatch, Problem
    from helicon_b200.grid import build_tasks
    from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec

    img = _image(N)
    rng = np.random.default_rng(2)
    tasks, _ = build_tasks(N, N, 1.3, [-1.2], [4.75], csyms=(1,), reconstruct_length_rise=3)
    g = tasks[0].geom
    prob = Problem(img, g["s"], g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
    n3 = g["L3"] * prob.ndisk
    target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
    spec = CandidateSpec(-1.2, 4.75 / g["apix3d"], 1, target, target, False)
    batch = ExplicitBatch(prob, g["L3"], spec, tilt_degree=tilt, interpolation=interp)
    nd_pad, tot = batch.rows_padded(0)
    m = batch.m_rows + batch.m_sym_explicit
    A, b, pid = batch.data_csr(0)
    assert A.shape == (batch.m_rows, batch.n) and len(np.unique(pid)) <= g["D2"] * g["L2"]
    x = rng.standard_normal(batch.n).astype(np.float32)
    y = batch.apply_forward(0, x)
    ref = A.astype(np.float64) @ x.astype(np.float64)
    assert np.abs(y[:batch.m_rows] - ref).max() <= 2e-5 * np.abs(ref).max()
    assert not y[m:nd_pad].any()
    u = np.zeros(tot, np.float32)
    u[:m] = rng.standard_normal(m).astype(np.float32)
    u[nd_pad:] = rng.standard_normal(tot - nd_pad).astype(np.float32)
    gx = batch.apply_adjoint(0, u)
    lhs = float(np.dot(y.astype(np.float64), u.astype(np.float64)))
    rhs = float(np.dot(x.astype(np.float64), gx.astype(np.float64)))
    assert abs(lhs - rhs) <= 2e-5 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)
    res = batch.solve(fixed_iters=6, check_every=6)
    assert res[0]["itn"] == 6 and np.isfinite(res[0]["score"]) and 0 < res[0]["score"] <= 1.0 + 1e-6
    print(f"N={N} {interp} tilt={tilt}: rows {batch.m_rows}+{batch.m_sym_explicit} entries {batch.nnz} score {res[0]['score']:.5f}")
    batch.close(); prob.close()
