"""Drop-in for ``helicon.webApps.denovo3D.solver_linear_regression`` (the
reference's denovo3D solver, "SLR") with the per-candidate work on the GPU.

Same function names, argument meaning, return types and array layouts as the
reference; see INTEGRATION.md.  What is NOT implemented on the CUDA path raises
``NotImplementedError`` (there is no CPU fallback by design):
score metrics other than
"cosine", and solver models other than ``{"model": "lsq"}``.
"""

from __future__ import annotations

import functools
import logging

import numpy as np

from . import planner
from . import bilinear
from .engine import Batch, ExplicitBatch, Problem, build_trilinear_sym_rows
from .planner import MAX_EQUATIONS, CandidateSpec, positive_rule
from .planner import sorted_hsym_csym_pairs  # noqa: F401  (SLR:1749-1791, re-exported)
from .lsqr import solve_lsqr
from .regularized import MODELS, solve_model

logger = logging.getLogger(__name__)


def _cache_compatible(func):
    """The reference wraps its builders in ``helicon.cache`` (lib/cache.py:132-209),
    which adds ``clear_cache``/``get_cache_info``/``__wrapped__``.  Results here
    come from the GPU in milliseconds, so nothing is memoised on disk; the
    attributes exist for call-site compatibility."""

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        return func(*args, **kwargs)

    wrapper.clear_cache = lambda: None
    wrapper.get_cache_info = lambda: {"cache_dir": None, "cache_period": None, "function_name": func.__name__}
    return wrapper


def _unsupported(what):
    raise NotImplementedError(f"helicon_b200 (CUDA path): {what} is not implemented; there is no CPU fallback")


def _interp(interpolation):
    # SLR:1401: "linear", "linear10" and "linear11" all select the trilinear kernel, anything else nearest neighbour
    return "linear" if interpolation in ("linear", "linear10", "linear11") else "nn"


def _is_explicit(tilt_degree, psi_degree, dy_pixel, interpolation):
    """The matrix-free projector covers the grid-search case (tilt = psi = dy = 0, nearest neighbour); everything
    else runs on explicit GPU-built rows (engine.ExplicitBatch)."""
    return bool(tilt_degree != 0 or psi_degree != 0 or dy_pixel != 0 or _interp(interpolation) == "linear")


def back_project_2d_coords_to_3d_coords(
    image, scale2d_to_3d, reconstruct_diameter_2d_pixel=-1, reconstruct_length_2d_pixel=-1
):
    """SLR:1657-1746.  Host-only helper kept for API parity: the CUDA kernels
    never materialise these (L2, D2, D2) float64 tables (closed form in
    registers, see ``k_build_fmap``)."""
    from scipy.spatial.transform import Rotation as R

    ny, nx = image.shape
    D2 = ny if reconstruct_diameter_2d_pixel <= 0 else reconstruct_diameter_2d_pixel
    L2 = nx if reconstruct_length_2d_pixel <= 0 else reconstruct_length_2d_pixel
    D2, L2 = int(np.rint(D2)), int(np.rint(L2))
    depth = (np.arange(D2, dtype=np.int32) - D2 // 2).astype(np.float32)
    across = (np.arange(D2, dtype=np.int32) - D2 // 2).astype(np.float32)
    along = (np.arange(L2, dtype=np.int32) - L2 // 2).astype(np.float32)
    region = image[np.ix_(across.astype(np.int32) + ny // 2, along.astype(np.int32) + nx // 2)]
    Zg, Yg, Xg = np.meshgrid(depth, across, along, indexing="ij")
    pts = np.stack((Xg.ravel(), Yg.ravel(), Zg.ravel()), axis=1)
    pts = R.from_euler("y", 90, degrees=True).apply(pts, inverse=True)
    if scale2d_to_3d != 1.0:
        pts *= scale2d_to_3d
    tables = tuple(np.swapaxes(pts[:, c].reshape((D2, D2, L2)), 0, 2) for c in range(3))
    return tables, region


def _disk_mask(D3, rmin, rmax):
    """In-plane support of helicon.get_cylindrical_mask (lib/analysis.py:731-774): rmin^2 <= x^2 + y^2 < rmax^2."""
    j = np.arange(D3) - D3 // 2
    Y, X = np.meshgrid(j, j, indexing="ij")
    r2 = X * X + Y * Y
    m = r2 < rmax * rmax
    if 0 < rmin < rmax:
        m &= r2 >= rmin * rmin
    return m


def _make_problem(image, scale2d_to_3d, D2, L2, D3, D3_inner, rmax=None, interpolation="nn", device=0):
    rmin = D3_inner / 2
    if rmax is None:
        rmax = D3 // 2 - 1
    return Problem(image, scale2d_to_3d, D2, L2, D3, rmin, rmax, device=device, interpolation=interpolation)


@_cache_compatible
def build_A_data_matrix(
    image,
    scale2d_to_3d,
    twist_degree,
    rise_pixel,
    csym,
    tilt_degree,
    psi_degree,
    dy_pixel,
    reconstruct_diameter_2d_pixel,
    reconstruct_length_2d_pixel,
    reconstruct_diameter_3d_pixel,
    reconstruct_diameter_3d_inner_pixel,
    reconstruct_length_3d_pixel,
    min_projection_lines,
    interpolation,
    verbose=0,
    cpu=1,
):
    """SLR:1301-1654 -> (csr_matrix float32, b float32, b_pid int32).

    The sample->voxel maps, ray validity and right-hand side come from the CUDA
    kernels; only the COO->CSR packing happens on the host."""
    image = np.asarray(image)
    D2 = reconstruct_diameter_2d_pixel if reconstruct_diameter_2d_pixel > 0 else image.shape[0]
    L2 = reconstruct_length_2d_pixel if reconstruct_length_2d_pixel > 0 else image.shape[1]
    L3 = reconstruct_length_3d_pixel if reconstruct_length_3d_pixel > 0 else L2
    # the reference builds this mask on the (L3, D2, D2) grid with rmax from D3 (SLR:1382-1384)
    prob = _make_problem(image, scale2d_to_3d, D2, L2, D2, reconstruct_diameter_3d_inner_pixel,
                         rmax=reconstruct_diameter_3d_pixel // 2 - 1)
    spec = CandidateSpec(twist_degree, rise_pixel, csym, min_projection_lines, -1, False)
    if _is_explicit(tilt_degree, psi_degree, dy_pixel, interpolation):
        batch = ExplicitBatch(prob, L3, spec, tilt_degree, psi_degree, dy_pixel, _interp(interpolation))
    else:
        batch = Batch(prob, L3, [spec])
    try:
        A, b, pid = batch.data_csr(0)
        A.sum_duplicates()
        return A, b, pid
    finally:
        batch.close()
        prob.close()


@_cache_compatible
def build_A_helical_sym_matrix(
    nz, ny, nx, twist_degree, rise_pixel, csym, rmin, rmax, min_sym_pairs, interpolation, verbose=0
):
    """SLR:844-1298 -> (csr_matrix float32 | None, zeros float32 | None)."""
    if ny != nx:
        _unsupported("a non-square symmetry grid (ny != nx)")
    prob = Problem(np.zeros((ny, nx), dtype=np.float32), 1.0, ny, nx, ny, rmin, rmax if rmax >= 0 else ny // 2 - 1)
    if interpolation in ("linear", "linear01", "linear11"):  # SLR:907 (this list differs from the data builder's)
        try:
            return build_trilinear_sym_rows(prob, nz, twist_degree, rise_pixel, csym, min_sym_pairs)
        finally:
            prob.close()
    spec = CandidateSpec(twist_degree, rise_pixel, csym, 1, min_sym_pairs, False)
    batch = Batch(prob, nz, [spec])
    try:
        return batch.sym_csr(0)
    finally:
        batch.close()
        prob.close()


def lsq_reconstruct(
    projection_image,
    scale2d_to_3d,
    twist_degree,
    rise_pixel,
    csym=1,
    tilt_degree=0,
    psi_degree=0,
    dy_pixel=0,
    thresh_fraction=-1,
    positive_constraint=-1,
    reconstruct_diameter_3d_inner_pixel=0,
    reconstruct_diameter_2d_pixel=-1,
    reconstruct_diameter_3d_pixel=-1,
    reconstruct_length_2d_pixel=-1,
    reconstruct_length_3d_pixel=-1,
    sym_oversample=1,
    interpolation="nn",
    fsc_test=0,
    score_metric="cosine",
    target_apix2d=5.0,
    verbose=0,
    algorithm=dict(model="lsq"),
    refine_tilt_psi_dy_range=None,
    cpu=1,
    device=0,
    return_info=False,
):
    """SLR:31-547 -> ((rec3d, None, None), score) with rec3d float32 (L3, D3, D3).

    One candidate through the batched GPU path (a batch of one).  ``device`` and
    ``return_info`` are additive keyword arguments."""
    explicit = _is_explicit(tilt_degree, psi_degree, dy_pixel, interpolation)
    if interpolation in ("linear10", "linear01"):
        _unsupported(f"interpolation={interpolation!r} (the reference's data and symmetry builders disagree on it, "
                     "SLR:907 vs 1401)")
    model = algorithm.get("model", "lsq")
    if model not in ("lsq", "_reproject") + MODELS:
        _unsupported(f"algorithm model {model!r} ('lsq', SLR:243-270, and 'elasticnet' / 'lasso' / 'ridge', SLR:293-322, are "
                     "implemented; 'lreg' / 'ard' densify the matrix)")
    if score_metric not in ("cosine", "ssim", "ms_ssim", "mutual_information", "composite"):
        raise ValueError(f"unknown score_metric {score_metric!r}")
    if score_metric != "cosine" and fsc_test:
        # the reference itself cannot do this: it scatters a half set's prediction with the FULL set's pixel ids
        # (SLR:507-519, `pred_2d.ravel()[b_data_pid] = pred`) and numpy raises on the length mismatch
        _unsupported(f"score_metric {score_metric!r} together with fsc_test (fails in the reference as well, SLR:507-519)")
    want_refine = refine_tilt_psi_dy_range is not None and any(
        refine_tilt_psi_dy_range.get(k, 0) > 0 for k in ("tilt", "psi", "dy"))
    if want_refine:
        # SLR:372-437: solve at the given orientation, then the local Gauss-Newton refinement.  For model "lsq" the
        # reference's solve_equations returns score=None (SLR:270), so its test `score is None or score_refined > score`
        # (SLR:424) is always true: the refined solution, its score and _refined_params are ALWAYS adopted.
        kw = dict(csym=csym, tilt_degree=tilt_degree, psi_degree=psi_degree, dy_pixel=dy_pixel,
                  thresh_fraction=thresh_fraction, positive_constraint=positive_constraint,
                  reconstruct_diameter_3d_inner_pixel=reconstruct_diameter_3d_inner_pixel,
                  reconstruct_diameter_2d_pixel=reconstruct_diameter_2d_pixel,
                  reconstruct_diameter_3d_pixel=reconstruct_diameter_3d_pixel,
                  reconstruct_length_2d_pixel=reconstruct_length_2d_pixel,
                  reconstruct_length_3d_pixel=reconstruct_length_3d_pixel, sym_oversample=sym_oversample,
                  interpolation=interpolation, fsc_test=fsc_test, score_metric=score_metric, target_apix2d=target_apix2d,
                  verbose=verbose, algorithm=algorithm, refine_tilt_psi_dy_range=None, cpu=cpu, device=device)
        (rec3d, half1, half2), score, info0 = lsq_reconstruct(projection_image, scale2d_to_3d, twist_degree, rise_pixel,
                                                               return_info=True, **kw)
        r = refine_tilt_psi_dy_range
        tilt_o, psi_o, dy_o, x_ref, score_ref = refine_tilt_psi_dy(
            projection_image, scale2d_to_3d, twist_degree, rise_pixel, csym, reconstruct_diameter_2d_pixel,
            reconstruct_length_2d_pixel, reconstruct_diameter_3d_pixel,
            reconstruct_diameter_3d_inner_pixel, reconstruct_length_3d_pixel, sym_oversample, interpolation, None,
            tilt_0=0.0, psi_0=0.0, dy_0=0.0, delta_tilt=r.get("delta_tilt", 0.5), delta_psi=r.get("delta_psi", 1.0),
            delta_dy=r.get("delta_dy", 0.2), max_iter=r.get("max_iter", 5),
            bounds_tilt=(-r.get("tilt", 30.0), r.get("tilt", 30.0)), bounds_psi=(-r.get("psi", 45.0), r.get("psi", 45.0)),
            bounds_dy=(-r.get("dy", 5.0), r.get("dy", 5.0)), positive_constraint=positive_constraint,
            algorithm=algorithm, verbose=verbose, cpu=cpu, device=device)
        if score_ref is not None:
            D3r, L3r = reconstruct_diameter_3d_pixel, reconstruct_length_3d_pixel
            m2 = _disk_mask(D3r, reconstruct_diameter_3d_inner_pixel / 2, D3r // 2 - 1)
            rec3d = np.zeros((L3r, D3r, D3r), dtype=np.float32)
            rec3d[:, m2] = np.asarray(x_ref, dtype=np.float32).reshape(L3r, -1)
            score = score_ref
            if fsc_test:
                # SLR:441-528: the half sets were solved at the GIVEN orientation; every score is then recomputed, the
                # full set's as cosine(A_data(given orientation) @ x_refined, b_data), and combined 1/2, 1/4, 1/4
                kw0 = dict(kw, fsc_test=0, algorithm=dict(model="_reproject", x=np.asarray(x_ref, dtype=np.float32)))
                _, s0 = lsq_reconstruct(projection_image, scale2d_to_3d, twist_degree, rise_pixel, **kw0)
                s = info0["scores"]
                score = np.float32(s0) / 2 + (np.float32(s[1]) + np.float32(s[2])) / 4
            if not hasattr(lsq_reconstruct, "_refined_params"):  # SLR:431-435 (consumed by pipeline.py:429-434)
                lsq_reconstruct._refined_params = {}
            lsq_reconstruct._refined_params.update(tilt=tilt_o, psi=psi_o, dy=dy_o)
        if return_info:
            return (rec3d, half1, half2), score, dict(refined=(tilt_o, psi_o, dy_o, score_ref))
        return (rec3d, half1, half2), score
    image = np.asarray(projection_image)
    D3, L3 = reconstruct_diameter_3d_pixel, reconstruct_length_3d_pixel
    D2 = reconstruct_diameter_2d_pixel if reconstruct_diameter_2d_pixel > 0 else image.shape[0]
    L2 = reconstruct_length_2d_pixel if reconstruct_length_2d_pixel > 0 else image.shape[1]
    if D3 <= 0 or L3 <= 0:
        raise ValueError("reconstruct_diameter_3d_pixel and reconstruct_length_3d_pixel must be given")
    prob = _make_problem(image, scale2d_to_3d, D2, L2, D3, reconstruct_diameter_3d_inner_pixel,
                         interpolation=_interp(interpolation), device=device)
    try:
        n3 = L3 * prob.ndisk
        # SLR:130, 148-150, 168-170: the reference multiplies the RAW arguments (-1 * -1 = 1 with the defaults)
        n2 = int(reconstruct_diameter_2d_pixel) * int(reconstruct_length_2d_pixel)
        target = min(MAX_EQUATIONS, int(max(n2, n3) * sym_oversample))
        positive = positive_rule(positive_constraint, rise_pixel, twist_degree, L3)
        spec = CandidateSpec(twist_degree, rise_pixel, csym, target, target, positive)
        nsets = 3 if fsc_test >= 1 else 1
        # trilinear interpolation at tilt = psi = dy = 0: the matrix-free factorisation (bilinear.py) through the same code
        # path as nearest neighbour; the sklearn-model branch stays on the explicit rows
        if (explicit and _interp(interpolation) == "linear" and tilt_degree == 0 and psi_degree == 0 and dy_pixel == 0
                and model == "lsq" and bilinear.supported(prob, L3)):
            explicit = False
            make_batch = lambda specs: bilinear.BilinearBatch(prob, L3, specs)
        else:
            make_batch = lambda specs: Batch(prob, L3, specs)
        # fsc_test: the full set and the two half sets are three candidates of one batch that differ only in
        # which data rows they keep (SLR:441-482); the symmetry rows are shared by construction.
        if explicit:
            # explicit GPU-built data rows (general orientation / trilinear): one candidate per batch; the half sets of
            # fsc_test are two more batches whose builder drops the other half's rows
            xs, scs, infos = [], [], []
            masks = [None]
            for hs in range(nsets):
                batch = ExplicitBatch(prob, L3, spec, tilt_degree, psi_degree, dy_pixel, _interp(interpolation),
                                      pixel_mask=masks[hs])
                try:
                    if hs == 0 and nsets == 3:
                        _, kk, jj = batch.data_row_index(0)
                        set1 = split_pixel_ids((kk * D2 + jj).astype(np.int32), fsc_test)
                        m1 = np.zeros(L2 * D2, dtype=np.uint8)
                        m1[set1] = 1
                        masks += [m1, 1 - m1]
                    if model != "lsq":  # SLR:293-342 on the explicit rows (the builder already dropped a half set's rows)
                        minfo = {}
                        w = algorithm["x"] if model == "_reproject" else solve_model(batch, 0, algorithm, positive, info=minfo)
                        _, kk, jj = batch.data_row_index(0)
                        sc = _score_of(score_metric, batch.predict(w), batch.data_b(), kk * D2 + jj, image, D2, L2,
                                       thresh_fraction)
                        xs.append(_embed(w, L3, D3, reconstruct_diameter_3d_inner_pixel)); scs.append(np.float32(sc))
                        infos.append((dict(model=minfo), batch.timing()))
                        continue
                    r = batch.solve(clip_pred=int(thresh_fraction >= 0))
                    xs.append(batch.rec3d(0))
                    if score_metric != "cosine":  # after rec3d: the operator calls below reuse the solver's work vectors
                        x_sol = batch.x(0)
                        _, kk, jj = batch.data_row_index(0)
                        r[0]["score"] = _metric_2d(score_metric, batch.predict(x_sol), batch.data_b(),
                                                   (kk * D2 + jj), image, D2, L2, thresh_fraction)
                    scs.append(np.float32(r[0]["score"]))
                    infos.append((r[0].copy(), batch.timing()))
                finally:
                    batch.close()
            rec3d = xs[0]
            half1, half2 = (xs[1], xs[2]) if nsets == 3 else (None, None)
            score = scs[0] / 2 + (scs[1] + scs[2]) / 4 if nsets == 3 else scs[0]
            info = dict(res=infos[0][0], all_res=[i[0] for i in infos] if model != "lsq" else np.array([i[0] for i in infos]),
                        timing=infos[0][1], scores=list(scs))
            if return_info:
                return (rec3d, half1, half2), score, info
            return (rec3d, half1, half2), score
        batch = make_batch([spec] * nsets)
        half1 = half2 = None
        try:
            if nsets == 3:
                _, kk, jj = batch.data_row_index(0)
                set1 = split_pixel_ids((kk * D2 + jj).astype(np.int32), fsc_test)
                m1 = np.zeros(L2 * D2, dtype=np.uint8)
                m1[set1] = 1
                batch.set_pixel_masks(np.stack([m1, 1 - m1]), [-1, 0, 1])
            if model != "lsq":
                # SLR:293-342: ElasticNet / Lasso / Ridge on the same equations, minimised with the matrix-free operator
                # (regularized.py); the half sets of fsc_test are the masked candidates 1 and 2 of the batch
                keeps = [None] + ([m1, 1 - m1] if nsets == 3 else [])
                vols, scs, minfos = [], [], []
                nd_pad = batch.rows_padded(0)[0]
                for c in range(nsets):
                    minfo = {}
                    w = (algorithm["x"] if model == "_reproject"
                         else solve_model(batch, c, algorithm, positive, row_keep=keeps[c], info=minfo))
                    pidx, kk, jj = batch.data_row_index(c)
                    if keeps[c] is not None:
                        sel = keeps[c][kk * D2 + jj].astype(bool)
                        pidx, kk, jj = pidx[sel], kk[sel], jj[sel]
                    pred = batch.apply_forward(c, w)[:nd_pad][pidx]
                    scs.append(np.float32(_score_of(score_metric, pred, batch.rhs_padded(c)[pidx], kk * D2 + jj, image, D2,
                                                    L2, thresh_fraction)))
                    vols.append(_embed(w, L3, D3, reconstruct_diameter_3d_inner_pixel))
                    minfos.append(minfo)
                rec3d = vols[0]
                if nsets == 3:
                    half1, half2 = vols[1], vols[2]
                    score = scs[0] / 2 + (scs[1] + scs[2]) / 4  # SLR:527-528
                else:
                    score = scs[0]
                if return_info:
                    return (rec3d, half1, half2), score, dict(model=minfos[0], all_models=minfos, timing=batch.timing(),
                                                              scores=list(scs))
                return (rec3d, half1, half2), score
            res = batch.solve(clip_pred=int(thresh_fraction >= 0))
            rec3d = batch.rec3d(0)
            if score_metric != "cosine":
                pidx, kk, jj = batch.data_row_index(0)
                pred = batch.apply_forward(0, batch.x(0))[: batch.rows_padded(0)[0]][pidx]
                b_ref = batch.rhs_padded(0)[pidx]
                res[0]["score"] = _metric_2d(score_metric, pred, b_ref, (kk * D2 + jj), image, D2, L2, thresh_fraction)
            if nsets == 3:
                half1, half2 = batch.rec3d(1), batch.rec3d(2)
                s = [np.float32(r["score"]) for r in res]  # SLR:527-528
                score = s[0] / 2 + (s[1] + s[2]) / 4
            else:
                score = np.float32(res[0]["score"])
            info = dict(res=res[0].copy(), all_res=res.copy(), timing=batch.timing(),
                        scores=[np.float32(r["score"]) for r in res])
        finally:
            batch.close()
    finally:
        prob.close()
    if return_info:
        return (rec3d, half1, half2), score, info
    return (rec3d, half1, half2), score


def _embed(x, L3, D3, D3_inner):
    """SLR:536-540: the unknowns scattered into the (L3, D3, D3) volume through the cylinder mask."""
    vol = np.zeros((L3, D3, D3), dtype=np.float32)
    vol[:, _disk_mask(D3, D3_inner / 2, D3 // 2 - 1)] = np.asarray(x, dtype=np.float32).reshape(L3, -1)
    return vol


def _score_of(score_metric, pred, b_data, pid, image, D2, L2, thresh_fraction):
    """SLR:484-525 for a prediction already on the host (model branch: the reprojection is one operator apply)."""
    if score_metric != "cosine":
        return _metric_2d(score_metric, pred, b_data, pid, image, D2, L2, thresh_fraction)
    pred = np.asarray(pred, dtype=np.float32)
    if thresh_fraction >= 0:
        pred = np.clip(pred, 0, None)
    return planner_cosine(pred, np.asarray(b_data, dtype=np.float32))


def _metric_2d(score_metric, pred, b_data, pid, image, D2, L2, thresh_fraction):
    """SLR:484-525 for the 2-D metrics: the reprojection scattered back to the (L2, D2) image by pixel id (a later row
    of the same pixel overwrites an earlier one, as numpy's fancy assignment does) against the transposed input region."""
    from . import imageprep as M

    pred = np.asarray(pred, dtype=np.float32)
    if thresh_fraction >= 0:
        pred = np.clip(pred, 0, None)
    ny, nx = image.shape
    region = image[ny // 2 - D2 // 2: ny // 2 + D2 // 2, nx // 2 - L2 // 2: nx // 2 + L2 // 2]
    pred_2d = np.zeros((L2, D2), dtype=np.float32)
    pred_2d.ravel()[np.asarray(pid, dtype=np.int64)] = pred
    ref_2d = region.T
    if score_metric == "ssim":
        return M.ssim_score(pred_2d, ref_2d)
    if score_metric == "ms_ssim":
        return M.ms_ssim_score(pred_2d, ref_2d)
    if score_metric == "mutual_information":
        return M.mutual_information_score(pred_2d, ref_2d)
    parts = [planner_cosine(pred, b_data), M.ssim_score(pred_2d, ref_2d), M.ms_ssim_score(pred_2d, ref_2d),
             M.mutual_information_score(pred_2d, ref_2d)]
    return float(np.mean(parts))


def split_pixel_ids(b_id, mode):
    """Pixel ids of half set 1 exactly as ``split_A_b`` picks them (SLR:175-203): mode 1 random halves (the same
    ``np.random.shuffle`` call on the same ``list(set(b_id))``, so a seeded global RNG reproduces the reference's
    split), 2 even/odd, 3 left/right, else outer thirds vs centre."""
    b_id_unique = sorted(set(b_id))
    n = len(b_id_unique)
    if mode == 1:
        b_id_unique = list(set(b_id))
        np.random.shuffle(b_id_unique)
        set1 = b_id_unique[: n // 2]
    elif mode == 2:
        set1 = b_id_unique[::2]
    elif mode == 3:
        set1 = b_id_unique[: n // 2]
    else:
        set1 = b_id_unique[: n // 3] + b_id_unique[n * 2 // 3:]
    return np.asarray(set1, dtype=np.int64)


def refine_tilt_psi_dy(
    projection_image,
    scale2d_to_3d,
    twist_degree,
    rise_pixel,
    csym,
    reconstruct_diameter_2d_pixel,
    reconstruct_length_2d_pixel,
    reconstruct_diameter_3d_pixel,
    reconstruct_diameter_3d_inner_pixel,
    reconstruct_length_3d_pixel,
    sym_oversample,
    interpolation,
    x_init,
    tilt_0=0.0,
    psi_0=0.0,
    dy_0=0.0,
    delta_tilt=0.5,
    delta_psi=1.0,
    delta_dy=0.2,
    max_iter=5,
    tol_tilt=0.05,
    tol_psi=0.1,
    tol_dy=0.05,
    bounds_tilt=(-30.0, 30.0),
    bounds_psi=(-45.0, 45.0),
    bounds_dy=(-5.0, 5.0),
    positive_constraint=-1,
    algorithm=None,
    verbose=0,
    cpu=1,
    device=0,
    unbounded_solver="lsqr",
    solve_info=None,
):
    """SLR:550-841: Gauss-Newton refinement of (tilt, psi, dy) with a finite-difference Jacobian -> ``(tilt, psi, dy,
    x, score)``.  The loop (perturb, 3x3 normal equations, projected step, convergence test, re-solve) is the
    reference's, on the host with scalars; every matrix build, prediction ``A_data @ x`` and solve runs on the GPU
    (explicit rows, ``engine.ExplicitBatch``).

    Solver: as the reference -- unbounded systems by ``scipy.sparse.linalg.lsqr(atol=btol=1e-6)`` (SLR:711-714), here
    scipy's own recurrence over the CUDA operator (lsqr.py: both products on the GPU); ``lsq_linear`` at its default
    tolerance when the positive rule fires (SLR:704-709), here the batch's LSMR + TRF state machines (tol = 1e-10).
    ``unbounded_solver="lsmr"`` keeps round 1's device-resident LSMR at the same tolerances (same minimiser, iterates
    differ at the level of the tolerance).  ``solve_info`` (a list) collects one dict per unbounded solve.  ``x_init``
    is unused, as in the reference."""
    image = np.asarray(projection_image)
    n2 = int(reconstruct_diameter_2d_pixel) * int(reconstruct_length_2d_pixel)  # SLR:639: raw product (1 with the -1 defaults)
    D2 = int(reconstruct_diameter_2d_pixel) if reconstruct_diameter_2d_pixel > 0 else image.shape[0]
    L2 = int(reconstruct_length_2d_pixel) if reconstruct_length_2d_pixel > 0 else image.shape[1]
    D3, L3 = int(reconstruct_diameter_3d_pixel), int(reconstruct_length_3d_pixel)
    t = np.array([tilt_0, psi_0, dy_0], dtype=np.float64)
    deltas = np.array([delta_tilt, delta_psi, delta_dy], dtype=np.float64)
    lo = np.array([bounds_tilt[0], bounds_psi[0], bounds_dy[0]], dtype=np.float64)
    hi = np.array([bounds_tilt[1], bounds_psi[1], bounds_dy[1]], dtype=np.float64)
    # SLR:634-638: the row targets use the FULL grid size D3*D3*L3 here (not the mask count as lsq_reconstruct does)
    target = min(MAX_EQUATIONS, int(max(n2, D3 * D3 * L3) * sym_oversample))
    positive = positive_rule(positive_constraint, rise_pixel, twist_degree, L3)
    interp = _interp(interpolation)
    prob = _make_problem(image, scale2d_to_3d, D2, L2, D3, reconstruct_diameter_3d_inner_pixel, device=device)
    spec = CandidateSpec(twist_degree, rise_pixel, csym, target, target, positive)
    n = L3 * prob.ndisk

    def build(tv):
        return ExplicitBatch(prob, L3, spec, float(tv[0]), float(tv[1]), float(tv[2]), interp)

    def solve(batch):
        m = batch.rows_padded(0)[1]  # data + symmetry rows (padded count: an upper bound is enough for the limit)
        if positive:  # lsq_linear(A, b, bounds, max_iter=200): tol = 1e-10, lsmr_tol = 1e-2 * tol, lsmr_maxiter = min(m, n)
            batch.solve(atol=1e-12, btol=1e-12, max_iter=int(min(max(m, 1), n)), trf_tol=1e-10, trf_max_iter=200)
        elif unbounded_solver == "lsqr":  # lsqr(A, b, atol=1e-6, btol=1e-6): scipy's recurrence, GPU products
            info = {}
            x = solve_lsqr(batch, 0, atol=1e-6, btol=1e-6, info=info)
            if solve_info is not None:
                solve_info.append(info)
            return x
        else:         # device-resident LSMR at the same tolerances, iter_lim = 2 n
            batch.solve(atol=1e-6, btol=1e-6, max_iter=2 * n)
        return batch.x(0)

    try:
        base = build(t)
        try:
            b_data = base.data_b()
            x_cur = solve(base)
            p_0 = base.predict(x_cur)
        finally:
            base.close()
        n_base = len(b_data)
        for iteration in range(max_iter):
            J = np.zeros((n_base, 3), dtype=np.float64)
            for i in range(3):
                tp = t.copy()
                tp[i] = np.clip(tp[i] + deltas[i], lo[i], hi[i])
                pert = build(tp)
                try:
                    p_pert = pert.predict(x_cur)
                finally:
                    pert.close()
                actual = tp[i] - t[i]
                if abs(actual) > 1e-12:
                    nc = min(n_base, len(p_pert))  # SLR:776-779: row sets may differ in size
                    J[:nc, i] = (p_pert[:nc].astype(np.float64) - p_0[:nc]) / actual
            r_0 = p_0.astype(np.float64) - b_data
            G = J.T @ J
            g = J.T @ r_0
            cond = np.linalg.cond(G) if np.linalg.det(G) != 0 else float("inf")
            if cond > 1e10:
                G = G + 1e-6 * np.diag(np.diag(G))
            try:
                delta_t = np.linalg.solve(G, -g)
            except np.linalg.LinAlgError:
                logger.warning(f"  Refine iter {iteration}: singular system, stopping")
                break
            t_new = np.clip(t + delta_t, lo, hi)
            step = t_new - t
            converged = abs(step[0]) < tol_tilt and abs(step[1]) < tol_psi and abs(step[2]) < tol_dy
            t = t_new
            if converged:
                break
            cur = build(t)
            try:
                if cur.m_rows != n_base:
                    # the reference concatenates the NEW rows with the BASE right-hand side here (SLR:826) and fails on
                    # a size mismatch; keep the last consistent solution instead of raising from inside the loop
                    logger.warning("  Refine: the row set changed size at t=%s; stopping", t)
                    break
                x_cur = solve(cur)
                p_0 = cur.predict(x_cur)
            finally:
                cur.close()
    finally:
        prob.close()
    score = float(planner_cosine(p_0, b_data))
    return float(t[0]), float(t[1]), float(t[2]), x_cur, score


def planner_cosine(a, b):
    """helicon.cosine_similarity (lib/analysis.py:802-821) for the refinement's final score (host, two vectors)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    na, nb = np.linalg.norm(a), np.linalg.norm(b)
    if na == 0 or nb == 0:
        return 0.0
    return float(np.sum(a * b) / (na * nb))
