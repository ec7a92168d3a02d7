"""Host-side candidate planner (exact, cheap): everything the reference decides
with Python scalars per candidate -- ordered symmetry copies, ordered symmetry
pairs, rotation-matrix entries, the image-column -> z-slice assignment and the
row-count early stop -- restated so that the CUDA kernels receive plain tables.

Reference: src/helicon/webApps/denovo3D/solver_linear_regression.py ("SLR").
Nothing here touches pixels or voxels; that is the GPU's job.
"""

from __future__ import annotations

import functools
import itertools
import math

import numpy as np

from . import _lib

MAX_EQUATIONS = 2**26  # SLR:131


@functools.lru_cache(maxsize=4096)
def halton_indices(n: int) -> tuple:
    """``qmc.Halton(d=1, scramble=False).integers(0, n, n=n)`` (SLR:1566-1571,
    1785-1790): floor(n * van-der-Corput_2(i)).  Not a permutation."""
    out = []
    for i in range(n):
        f, r, k = 0.5, 0.0, i
        while k:
            if k & 1:
                r += f
            k >>= 1
            f *= 0.5
        out.append(int(math.floor(r * n)))
    return tuple(out)


@functools.lru_cache(maxsize=4096)
def _copies(hsym_max: int, csym: int) -> tuple:
    hc = list(itertools.product(range(-hsym_max, hsym_max + 1), range(csym)))
    hc.sort(key=lambda x: (abs(x[0]), x[1]))
    return tuple(hc[i] for i in halton_indices(len(hc)))


def data_copies(rise_pixel, csym, L3, L2):
    """Ordered (h, c) copies of the data operator incl. duplicates (SLR:1559-1571)."""
    hsym_max = max(1, int(np.ceil(L3 + L2) / 2 / rise_pixel))
    return _copies(hsym_max, int(csym))


def sorted_hsym_csym_pairs(twist, rise, csym, nz):
    """SLR:1749-1791 (same return value)."""
    hsym_max = max(1, int(np.ceil(nz / (2 * rise))))
    hcsyms = itertools.product(range(-hsym_max, hsym_max + 1), range(csym))
    out = []
    for p in itertools.combinations(hcsyms, r=2):
        (h1, c1), (h2, c2) = p
        a1 = twist * h1 + c1 * 360 / csym
        a2 = twist * h2 + c2 * 360 / csym
        angle = round(abs((a2 - a1 + 180) % 360 - 180), 2)
        out.append((angle, abs(h1 + h2), abs(h1 - h2), abs(h1), abs(h2), p))
    out.sort(key=lambda x: x[:-1])
    return [out[i] for i in halton_indices(len(out))]


def positive_rule(positive_constraint, rise_pixel, twist_degree, L3):
    """SLR:352-355."""
    pitch_pixel = round(rise_pixel * 360 / abs(twist_degree))
    return bool(positive_constraint > 0 or (positive_constraint < 0 and pitch_pixel > round(L3 * 2)))


def z_rotation_entries(angles_deg):
    """(M00, M10) of ``Rotation.from_euler('z', angle, degrees=True).as_matrix()``
    for an array of angles: the same scipy call the reference makes (SLR:1225,
    1235, 1576, 1615), so the entries carry the same last bits."""
    from scipy.spatial.transform import Rotation as R

    angles_deg = np.atleast_1d(np.asarray(angles_deg, dtype=np.float64))
    M = R.from_euler("z", angles_deg.reshape(-1, 1), degrees=True).as_matrix()
    return np.ascontiguousarray(np.stack([M[:, 0, 0], M[:, 1, 0]], axis=1))


def column_slices(s, L2, L3, zshift):
    """z-slice of every image column k for one symmetry copy (SLR:1578-1581,
    1529 with tilt=psi=0): Z = s*(k - L2//2) - h*rise_pixel + L3//2, rint
    (half-to-even).  Returns (zi int array of length L2 with -1 where outside
    [0, L3), tie flag)."""
    k = np.arange(L2, dtype=np.float64) - (L2 // 2)
    z0 = k * s if s != 1.0 else k
    Z = (z0 - zshift) + (L3 // 2)
    zi = np.rint(Z).astype(np.int64)
    ok = (zi >= 0) & (zi <= L3 - 1)
    near = np.abs(np.abs(Z - np.floor(Z)) - 0.5) < 1e-9
    tie = bool(np.any(near & (Z > -1.0) & (Z < L3)))
    return np.where(ok, zi, -1), tie


class CandidateSpec:
    """One (twist, rise, csym) with the row targets of SLR:148-150, 168-170."""

    __slots__ = ("twist", "rise_pixel", "csym", "min_projection_lines", "min_sym_pairs", "positive")

    def __init__(self, twist, rise_pixel, csym, min_projection_lines, min_sym_pairs, positive):
        self.twist = float(twist)
        self.rise_pixel = float(rise_pixel)
        self.csym = int(csym)
        self.min_projection_lines = int(min_projection_lines)
        self.min_sym_pairs = int(min_sym_pairs)
        self.positive = bool(positive)


class BatchPlan:
    """Two-stage plan of a batch with uniform (geometry, L3).

    stage 1 (``__init__``): unique angles + per-copy column tables;
    stage 2 (``finalize(nvalid)``): row-count early stop (SLR:1647) once the GPU
    has reported the number of rays with data per angle, then flat tables.

    Everything per candidate is a handful of numpy calls over (n_h, L2) arrays
    (all symmetry copies at once); the arithmetic per element is the scalar
    sequence of ``column_slices``.
    """

    def __init__(self, s, D2, L2, L3, specs):
        self.s, self.D2, self.L2, self.L3 = float(s), int(D2), int(L2), int(L3)
        self.specs = list(specs)
        L2, L3 = self.L2, self.L3
        kk = np.arange(L2, dtype=np.float64) - (L2 // 2)
        z0 = kk * self.s if self.s != 1.0 else kk
        angle_index = {}
        angles = []
        self._cand = []  # per candidate: (copies, angle ids, index of each copy's h in hs, hs, ZI[n_h, L2])
        self.cand_tie_z = []
        mc = 1
        for sp in self.specs:
            copies = data_copies(sp.rise_pixel, sp.csym, L3, L2)
            hs = sorted({h for h, _ in copies})
            hpos = {h: i for i, h in enumerate(hs)}
            zshift = np.array([h * sp.rise_pixel for h in hs], dtype=np.float64)
            Z = (z0[None, :] - zshift[:, None]) + (L3 // 2)
            zi = np.rint(Z).astype(np.int64)
            ok = (zi >= 0) & (zi <= L3 - 1)
            near = np.abs(np.abs(Z - np.floor(Z)) - 0.5) < 1e-9
            self.cand_tie_z.append(bool(np.any(near & (Z > -1.0) & (Z < L3))))
            ZI = np.where(ok, zi, -1)
            if ok.any():
                flat = (np.arange(len(hs))[:, None] * L3 + zi)[ok]
                mc = max(mc, int(np.bincount(flat).max()))
            aid = np.empty(len(copies), dtype=np.int64)
            hidx = np.empty(len(copies), dtype=np.int64)
            tw, cs = sp.twist, sp.csym
            for i, (h, c) in enumerate(copies):
                angle = tw * h + 360 * c / cs
                a = angle_index.get(angle)
                if a is None:
                    a = len(angles)
                    angle_index[angle] = a
                    angles.append(angle)
                aid[i] = a
                hidx[i] = hpos[h]
            self._cand.append((copies, aid, hidx, hs, ZI))
        self.MC = mc
        self.angles = np.array(angles, dtype=np.float64)
        self.cos_sin = z_rotation_entries(self.angles)
        self.finalized = False

    def finalize(self, nvalid_rays):
        nvalid_rays = np.asarray(nvalid_rays, dtype=np.int64)
        L2, L3, MC = self.L2, self.L3, self.MC
        nc = len(self.specs)
        cands = np.zeros(nc, dtype=_lib.CANDIDATE_DTYPE)
        view_angle, colk = [], []
        self.cand_views = []  # per candidate: list of (angle_id, zi, h, c, n_rows_real)
        pair_meta = []        # (candidate, angle_i, angle_j, zshift_i, zshift_j) arrays, rotated in one scipy call below
        nviews = 0
        npairs = 0
        karange = np.arange(L2, dtype=np.int64)[None, :]
        for ci, sp in enumerate(self.specs):
            copies, aid, hidx, hs, ZI = self._cand[ci]
            valid = ZI >= 0
            ncols_h = valid.sum(axis=1)
            nrows = ncols_h[hidx] * nvalid_rays[aid]
            stop = len(copies)
            if sp.min_projection_lines > 0:
                over = np.nonzero(np.cumsum(nrows) > sp.min_projection_lines)[0]
                if len(over):
                    stop = int(over[0]) + 1
            sel = np.nonzero(nrows[:stop] > 0)[0]
            self.cand_views.append([(int(aid[i]), ZI[hidx[i]], copies[i][0], copies[i][1], int(nrows[i])) for i in sel])
            cands[ci]["view_begin"] = nviews
            cands[ci]["view_count"] = len(sel)
            if len(sel):
                # column table of every h: tab[z*MC + m] = m-th image column whose slice is z (columns ascending)
                prev = np.concatenate([np.full((len(hs), 1), -2, dtype=np.int64), ZI[:, :-1]], axis=1)
                run_start = np.maximum.accumulate(np.where(valid & (ZI != prev), karange, 0), axis=1)
                tab = np.full((len(hs), L3 * MC), -1, dtype=np.int32)
                hh, kq = np.nonzero(valid)
                tab[hh, ZI[hh, kq] * MC + (kq - run_start[hh, kq])] = kq
                colk.append(tab[hidx[sel]])
                view_angle.append(aid[sel])
                nviews += len(sel)
            # symmetry pairs (SLR:892, 1223-1243)
            cands[ci]["pair_begin"] = npairs
            plist = sorted_hsym_csym_pairs(sp.twist, sp.rise_pixel, sp.csym, L3) if sp.min_sym_pairs >= 0 else []
            if plist:
                pq = np.array([(p[-1][0][0], p[-1][0][1], p[-1][1][0], p[-1][1][1]) for p in plist], dtype=np.float64)
                ai = sp.twist * pq[:, 0] + pq[:, 1] * 360 / sp.csym
                aj = sp.twist * pq[:, 2] + pq[:, 3] * 360 / sp.csym
                pair_meta.append((ai, aj, sp.rise_pixel * pq[:, 0], sp.rise_pixel * pq[:, 2]))
                npairs += len(plist)
            cands[ci]["pair_count"] = len(plist)
            cands[ci]["min_sym_pairs"] = sp.min_sym_pairs
            cands[ci]["positive"] = int(sp.positive)
            cands[ci]["flags_in"] = _lib.HB2_FLAG_TIE_Z if self.cand_tie_z[ci] else 0
        self.cands = cands
        views = np.zeros(nviews, dtype=_lib.VIEW_DTYPE)
        if nviews:
            views["angle"] = np.concatenate(view_angle)
            views["col_begin"] = np.arange(nviews, dtype=np.int64) * (L3 * MC)
            self.colk = np.ascontiguousarray(np.concatenate(colk).reshape(-1), dtype=np.int32)
        else:
            self.colk = np.zeros(0, dtype=np.int32)
        self.views = views
        pairs = np.zeros(npairs, dtype=_lib.PAIR_DTYPE)
        if npairs:
            ai = np.concatenate([m[0] for m in pair_meta])
            aj = np.concatenate([m[1] for m in pair_meta])
            ei, ej = z_rotation_entries(ai), z_rotation_entries(aj)
            pairs["ci"], pairs["si"], pairs["zi"] = ei[:, 0], ei[:, 1], np.concatenate([m[2] for m in pair_meta])
            pairs["cj"], pairs["sj"], pairs["zj"] = ej[:, 0], ej[:, 1], np.concatenate([m[3] for m in pair_meta])
        self.pairs = pairs
        self.finalized = True
        return self
