"""Drop-in for ``helicon.webApps.denovo3D.solver_linear_regression`` (the
reference's denovo3D solver, "SLR") with the per-candidate work on the GPU.

Same function names, argument meaning, return types and array layouts as the
reference; see INTEGRATION.md.  What is NOT implemented on the CUDA path raises
``NotImplementedError`` (there is no CPU fallback by design):
``refine_tilt_psi_dy``, score metrics other than
"cosine", and solver models other than ``{"model": "lsq"}``.
"""

from __future__ import annotations

import functools
import logging

import numpy as np

from . import planner
from .engine import Batch, ExplicitBatch, Problem, build_trilinear_sym_rows
from .planner import MAX_EQUATIONS, CandidateSpec, positive_rule
from .planner import sorted_hsym_csym_pairs  # noqa: F401  (SLR:1749-1791, re-exported)

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
    if algorithm.get("model", "lsq") != "lsq":
        _unsupported(f"algorithm model {algorithm.get('model')!r} (only 'lsq', SLR:243-270)")
    if score_metric != "cosine":
        _unsupported(f"score_metric {score_metric!r} (only 'cosine', SLR:500-525)")
    if refine_tilt_psi_dy_range is not None and any(
        refine_tilt_psi_dy_range.get(k, 0) > 0 for k in ("tilt", "psi", "dy")
    ):
        _unsupported("refine_tilt_psi_dy (SLR:550-841)")
    image = np.asarray(projection_image)
    D3, L3 = reconstruct_diameter_3d_pixel, reconstruct_length_3d_pixel
    D2 = reconstruct_diameter_2d_pixel if reconstruct_diameter_2d_pixel > 0 else image.shape[0]
    L2 = reconstruct_length_2d_pixel if reconstruct_length_2d_pixel > 0 else image.shape[1]
    if D3 <= 0 or L3 <= 0:
        raise ValueError("reconstruct_diameter_3d_pixel and reconstruct_length_3d_pixel must be given")
    prob = _make_problem(image, scale2d_to_3d, D2, L2, D3, reconstruct_diameter_3d_inner_pixel,
                         device=device)
    try:
        n3 = L3 * prob.ndisk
        target = min(MAX_EQUATIONS, int(max(D2 * L2, n3) * sym_oversample))  # SLR:148-150, 168-170
        positive = positive_rule(positive_constraint, rise_pixel, twist_degree, L3)
        spec = CandidateSpec(twist_degree, rise_pixel, csym, target, target, positive)
        nsets = 3 if fsc_test >= 1 else 1
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
                    r = batch.solve(clip_pred=int(thresh_fraction >= 0))
                    xs.append(batch.rec3d(0)); scs.append(np.float32(r[0]["score"]))
                    infos.append((r[0].copy(), batch.timing()))
                finally:
                    batch.close()
            rec3d = xs[0]
            half1, half2 = (xs[1], xs[2]) if nsets == 3 else (None, None)
            score = scs[0] / 2 + (scs[1] + scs[2]) / 4 if nsets == 3 else scs[0]
            info = dict(res=infos[0][0], all_res=np.array([i[0] for i in infos]), timing=infos[0][1])
            if return_info:
                return (rec3d, half1, half2), score, info
            return (rec3d, half1, half2), score
        batch = Batch(prob, L3, [spec] * nsets)
        half1 = half2 = None
        try:
            if nsets == 3:
                _, kk, jj = batch.data_row_index(0)
                set1 = split_pixel_ids((kk * D2 + jj).astype(np.int32), fsc_test)
                m1 = np.zeros(L2 * D2, dtype=np.uint8)
                m1[set1] = 1
                batch.set_pixel_masks(np.stack([m1, 1 - m1]), [-1, 0, 1])
            res = batch.solve(clip_pred=int(thresh_fraction >= 0))
            rec3d = batch.rec3d(0)
            if nsets == 3:
                half1, half2 = batch.rec3d(1), batch.rec3d(2)
                s = [np.float32(r["score"]) for r in res]  # SLR:527-528
                score = s[0] / 2 + (s[1] + s[2]) / 4
            else:
                score = np.float32(res[0]["score"])
            info = dict(res=res[0].copy(), all_res=res.copy(), timing=batch.timing())
        finally:
            batch.close()
    finally:
        prob.close()
    if return_info:
        return (rec3d, half1, half2), score, info
    return (rec3d, half1, half2), score


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


def refine_tilt_psi_dy(*args, **kwargs):
    """SLR:550-841: Gauss-Newton around ``lsqr(atol=btol=1e-6)`` / ``lsq_linear`` at its default tolerance.  Its
    building blocks (tilted rows, solves) run on the GPU through ``lsq_reconstruct(tilt_degree=..., ...)``; the
    refinement loop itself is not implemented."""
    _unsupported("refine_tilt_psi_dy (SLR:550-841)")
