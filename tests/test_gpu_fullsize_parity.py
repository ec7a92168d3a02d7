"""Parity at the BASELINE config shapes, through the C ABI.

Part A -- cfg1 (200x200) and cfg2 (256x256): outputs of the UNMODIFIED reference's ``lsq_reconstruct``
(oracle/make_golden_fullsize.py: score, LSMR stopping iteration, TRF outer iterations, x) for two candidates each, on
the unbounded path (positive_constraint=0) and on the reference-default rule (positive_constraint=-1 -> bounded TRF).
North-star tolerances: score |d| <= 1e-5, same stopping iteration +-1; x is asserted against the reference's own
row-permutation floor (2e-4...8e-4 at the stopping point of the loosely converged LSMR; the bounded path's
float64 outer iterations are reproducible to ~2e-4).

Part B -- the kernel templates only the big shapes select (512x512 -> uint32 maps, k_fwd_data<uint32>; 384x384 with
rise >= 21 A -> L3P > 16 -> k_adj_pq / chunked tile adjoint): the oracle's fixed-iteration LSMR iterate (20 iterations,
scipy's executed precision map) with the floor of that iterate under a row permutation stored beside it -- once with
numpy's float32 BLAS norms (scipy as executed) and once with exactly rounded norms.
"""
import numpy as np
import pytest

import os

from tests.helpers import GOLDEN, load

pytestmark = pytest.mark.gpu

REF_CASES = ["full_cfg1_true_unb", "full_cfg1_b_unb", "full_cfg2_true_unb", "full_cfg2_b_unb",
             "full_cfg1_true_pos", "full_cfg1_b_pos", "full_cfg2_true_pos", "full_cfg2_b_pos"]
FIXED_CASES = ["fixed_512_u32", "fixed_512_c3", "fixed_384_l3p52", "fixed_384_l3p104"]


def _disk(D3):
    from oracle import denovo3d_oracle as O

    return O.cylindrical_mask(1, D3, D3, 0, D3 // 2 - 1)[0]


def _reference_band(name):
    """The reference's OWN reproducibility at this case (committed, measured with the unmodified reference here):
    full_floor.npz = the same equations in permuted row order (rows [TRF nit, d score, rel-L2 x] of 3 permutations,
    oracle/make_golden_fullsize_floor.py); full_band.npz = the same call under other BLAS configurations (8 OpenBLAS
    threads; the AVX2 'Haswell' kernels instead of AVX-512: rows [d itn, d TRF nit, d score, rel-L2 x],
    oracle/make_golden_fullsize.py merge_band).  Returns (d itn, d trf nit, d score, rel) maxima."""
    fl, bd = load("full_floor"), load("full_band")
    f, b = fl[name], bd[name]
    gold_nit = int(load(name)["trf_nit"])
    d_trf = max(int(np.abs(b[:, 1]).max()), int(np.abs(f[:, 0] - gold_nit).max()) if gold_nit else 0)
    return (int(np.abs(b[:, 0]).max()), d_trf, float(max(np.abs(b[:, 2]).max(), np.abs(f[:, 1]).max())),
            float(max(b[:, 3].max(), f[:, 2].max())))


@pytest.mark.parametrize("name", REF_CASES)
def test_full_size_solve_vs_reference(name):
    """North-star tolerances (score 1e-5, x 1e-4) wherever the reference itself is reproducible to that level; where
    its own band (row order / BLAS kernel, see _reference_band) is wider, twice that band.  The LSMR norms follow the
    reference's executed float32 BLAS accumulation (norm_mode=1, hb2_kernels.cuh:k_chain_sumsq)."""
    from helicon_b200 import solver_linear_regression as S

    d = load(name)
    apix, twist, rise, csym, pc, so, L3, D2, L2, D3 = d["args"]
    (rec, _, _), score, info = S.lsq_reconstruct(
        d["image"], 1.0, float(twist), float(rise / apix), int(csym), positive_constraint=int(pc),
        reconstruct_diameter_2d_pixel=int(D2), reconstruct_length_2d_pixel=int(L2), reconstruct_diameter_3d_pixel=int(D3),
        reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so), interpolation="nn", return_info=True)
    x = rec[:, _disk(int(D3))].ravel()
    ref = d["x"]
    rel = float(np.linalg.norm(x - ref) / np.linalg.norm(ref))
    dscore = abs(float(score) - float(d["score"]))
    r = info["res"]
    b_itn, b_trf, b_score, b_rel = _reference_band(name)
    print(f"{name}: itn gpu={int(r['itn'])} ref={int(d['itn'])} istop={int(r['istop'])}/{int(d['istop'])} "
          f"trf nit gpu={int(r['trf_nit'])} ref={int(d['trf_nit'])} score={float(score):.7f} ref={float(d['score']):.7f} "
          f"|dscore|={dscore:.2e} rel-L2(x)={rel:.2e} flags={int(r['flags'])} | reference's own band: d itn {b_itn}, "
          f"d trf {b_trf}, d score {b_score:.1e}, rel {b_rel:.1e}")
    assert int(r["istop"]) == int(d["istop"])
    assert abs(int(r["itn"]) - int(d["itn"])) <= max(1, b_itn)
    # a candidate whose bounded branch the reference itself does not reproduce (band d trf > 0: another row order or BLAS
    # kernel ends the TRF loop on another iteration) is held to twice that band, like its score and x
    assert abs(int(r["trf_nit"]) - int(d["trf_nit"])) <= 2 * b_trf
    assert dscore <= max(1e-5, 2 * b_score)
    assert rel <= max(1e-4, 2 * b_rel)


@pytest.mark.parametrize("name", FIXED_CASES)
@pytest.mark.parametrize("norm_mode", [1, 0], ids=["blas_norms", "exact_norms"])
def test_full_size_fixed_iterations_vs_oracle(name, norm_mode):
    """20 LSMR iterations at the shapes that select the other kernel templates.  norm_mode=1 (default) against the
    oracle with numpy's OpenBLAS float32 norms (scipy as executed); norm_mode=0 against the oracle with exactly rounded
    norms.  Both within max(1e-4, 4 x the oracle iterate's own row-permutation floor)."""
    from helicon_b200.engine import Batch, Problem
    from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec

    d = load(name)
    apix, twist, rise, csym, pc, so, L3, D2, L2, D3, iters, stride = d["args"]
    L3, D2, L2, D3, iters, stride = int(L3), int(D2), int(L2), int(D3), int(iters), int(stride)
    prob = Problem(d["image"], 1.0, D2, L2, D3, 0.0, D3 // 2 - 1)
    n3 = L3 * prob.ndisk
    assert n3 == int(d["n"])
    target = min(MAX_EQUATIONS, int(max(D2 * L2, n3) * int(so)))
    batch = Batch(prob, L3, [CandidateSpec(float(twist), float(rise / apix), int(csym), target, target, False)])
    try:
        pidx, kk, jj = batch.data_row_index(0)
        nd_pad, tot = batch.rows_padded(0)
        assert len(pidx) == int(d["m_data"]) and tot - nd_pad == int(d["m_sym"])
        b = batch.rhs_padded(0)[pidx]
        assert abs(float(b.astype(np.float64).sum()) - float(d["b_sum"])) <= 1e-6 * abs(float(d["b_sum"]))
        assert int((kk.astype(np.int64) * D2 + jj).sum()) == int(d["pid_sum"])
        res = batch.solve(fixed_iters=iters, check_every=iters, norm_mode=norm_mode)
        assert int(res[0]["itn"]) == iters
        xs = batch.x(0)[::stride]
        ref = d["x_sample_blas"] if norm_mode else d["x_sample"]
        other = d["x_sample"] if norm_mode else d["x_sample_blas"]
        rel = float(np.linalg.norm(xs - ref) / np.linalg.norm(ref))
        rel_other = float(np.linalg.norm(xs - other) / np.linalg.norm(other))
        floor = float(d["floor"])
        print(f"{name} norm_mode={norm_mode}: n={n3} L3={L3} rows={len(pidx)}+{tot - nd_pad} fixed {iters} it: "
              f"rel-L2(x sample)={rel:.2e} (oracle permutation floor {floor:.2e}; vs the oracle's OTHER norm variant "
              f"{rel_other:.2e}; the oracle's two variants differ by {float(d['blas_rel']):.2e})")
        # norm_mode=1: k_chain_sumsq reproduces the STRUCTURE of the BLAS accumulation (64 sequential float32 chains), not
        # its element order; where the vector's magnitude is strongly ordered (L3 = 104 slices) the two orders differ
        # by about the size of the BLAS effect itself -- the bound is then 1.5 x that effect
        # (and, where the BLAS effect is large -- 6e-2 at 512 x 512, csym 3 -- at least 90 % of it must be reproduced)
        br = float(d["blas_rel"])
        assert rel <= max(1e-4, 4 * floor, (0.1 * br if norm_mode == 1 else 0.0),
                          (1.5 * br) if norm_mode == 1 and br < 2e-3 else 0.0)
        if norm_mode == 0:  # the score of a 20-iteration iterate (1 - score ~ 4e-4, far from converged) is reported only loosely
            assert abs(float(res[0]["score"]) - float(d["score"])) <= 1e-4
    finally:
        batch.close()
        prob.close()


def test_bounded_path_statistics_vs_reference_band():
    """The reference-default (bounded TRF) path over a seeded sample of 16 cfg2 grid candidates
    (oracle/make_golden_bounded_band.py: the unmodified reference's builders + its scipy call, once in its own row order
    = run 0, once with the rows permuted = run 1).  The LSMR stage must stop where the reference stops (+-2); the final
    scores are compared as a DISTRIBUTION: the CUDA path may leave run 0 no more often / no further than the
    reference's own run 1 does (its TRF exit is a discontinuous decision, see the golden script's docstring)."""
    from helicon_b200.engine import Batch, Problem
    from helicon_b200.grid import derive_geometry
    from helicon_b200.planner import MAX_EQUATIONS, CandidateSpec
    import bench

    d = load("bounded_band_cfg2")
    cand, rows = d["cand"], d["rows"]          # rows[i, run] = [lsmr itn, trf nit, score]
    N, apix = 256, 1.3
    img = bench.synthetic_filament(n=N)
    g = derive_geometry(N, N, apix, 4.75, 4.75, N * apix, 0.0, N * apix, 3 * 4.75, apix, 0, -1)
    prob = Problem(img, 1.0, g["D2"], g["L2"], g["D3"], 0.0, g["D3"] // 2 - 1)
    n3 = g["L3"] * prob.ndisk
    target = min(MAX_EQUATIONS, int(max(g["D2"] * g["L2"], n3) * g["sym_oversample"]))
    batch = Batch(prob, g["L3"], [CandidateSpec(float(t), float(r) / apix, 1, target, target, True) for t, r in cand])
    try:
        res = batch.solve()
    finally:
        batch.close()
        prob.close()
    d_gpu = np.abs(res["score"].astype(np.float64) - rows[:, 0, 2])
    d_ref = np.abs(rows[:, 1, 2] - rows[:, 0, 2])
    ditn = np.abs(res["itn"] - rows[:, 0, 0])
    for i in range(len(cand)):
        print(f"twist {cand[i][0]:8.4f} rise {cand[i][1]:.4f}: itn gpu {int(res['itn'][i])} ref {int(rows[i, 0, 0])}/{int(rows[i, 1, 0])} "
              f"trf gpu {int(res['trf_nit'][i])} ref {int(rows[i, 0, 1])}/{int(rows[i, 1, 1])} |dscore| gpu {d_gpu[i]:.2e} "
              f"reference's own {d_ref[i]:.2e}")
    print(f"median |dscore|: gpu {np.median(d_gpu):.2e}, reference's own {np.median(d_ref):.2e}; "
          f"within 1e-5: gpu {int((d_gpu <= 1e-5).sum())}/16, reference {int((d_ref <= 1e-5).sum())}/16; "
          f"max: gpu {d_gpu.max():.2e}, reference {d_ref.max():.2e}")
    assert np.all(ditn <= 2)
    assert np.median(d_gpu) <= max(1e-5, 2 * np.median(d_ref))
    assert (d_gpu <= 1e-5).sum() >= (d_ref <= 1e-5).sum() - 3
    assert d_gpu.max() <= max(5e-2, 2 * d_ref.max())


LIN_CASES = [n for n in ("full_cfg1_b_lin_unb", "full_cfg2_b_lin_unb", "full_cfg1_b_lin_pos") if os.path.exists(os.path.join(GOLDEN, n + ".npz"))]


@pytest.mark.parametrize("name", LIN_CASES)
def test_full_size_trilinear_solve_vs_reference(name):
    """interpolation="linear" (the app's default) at the BASELINE shapes through the matrix-free trilinear operator
    (helicon_b200/bilinear.py) against the UNMODIFIED reference's lsq_reconstruct (oracle/make_golden_fullsize.py).
    Trilinear systems are ill-conditioned: the reference's own result moves by |dscore| 1.2e-4 / rel-L2 2.3e-2 when its
    equations are merely permuted (tests/golden/gen_solve_lin_48.npz, oracle/make_golden_band.py).  The bar: the north-star
    score tolerance 1e-5 (measured 2.1e-6 at 200 x 200, 9.5e-7 at 256 x 256), x inside twice that band (measured 5.4e-3 /
    2.1e-2), the same stopping reason and the stopping iteration within 3 (455 / 454, 451 / 449)."""
    from helicon_b200 import solver_linear_regression as S

    d = load(name)
    apix, twist, rise, csym, pc, so, L3, D2, L2, D3 = d["args"]
    (rec, _, _), score, info = S.lsq_reconstruct(
        d["image"], 1.0, float(twist), float(rise / apix), int(csym), positive_constraint=int(pc),
        reconstruct_diameter_2d_pixel=int(D2), reconstruct_length_2d_pixel=int(L2), reconstruct_diameter_3d_pixel=int(D3),
        reconstruct_length_3d_pixel=int(L3), sym_oversample=int(so), interpolation="linear", return_info=True)
    x = rec[:, _disk(int(D3))].ravel()
    ref = d["x"]
    rel = float(np.linalg.norm(x - ref) / np.linalg.norm(ref))
    dscore = abs(float(score) - float(d["score"]))
    r = info["res"]
    print(f"{name}: itn gpu={int(r['itn'])} ref={int(d['itn'])} istop={int(r['istop'])}/{int(d['istop'])} "
          f"score={float(score):.7f} ref={float(d['score']):.7f} |dscore|={dscore:.2e} rel-L2(x)={rel:.2e} "
          f"data rows={int(r['n_data_rows'])} (reference solve: {float(d['seconds']):.0f} s on one core)")
    assert int(r["istop"]) == int(d["istop"])
    assert abs(int(r["itn"]) - int(d["itn"])) <= 3
    if int(pc) != 0:
        # reference-default rule: scipy's bounded TRF branch on the trilinear system (float64 instantiations of the matrix-free
        # kernels).  The reference's own band for bounded trilinear solves is |dscore| 4.6e-2 / rel-L2 0.26 under a row
        # permutation (tests/golden/gen_solve_lin_48_pos.npz): the step kinds of the TRF loop are data-dependent
        print(f"    bounded branch: TRF iterations gpu={int(r['trf_nit'])} ref={int(d['trf_nit'])} flags={int(r['flags'])}")
        assert int(r["flags"]) & 4 and rec.min() >= 0.0
        assert abs(int(r["trf_nit"]) - int(d["trf_nit"])) <= 5
        assert dscore <= 4.6e-2 and rel <= 0.26
    else:
        assert dscore <= 1e-5 and rel <= 4.6e-2
