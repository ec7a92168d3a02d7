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
    """

    def __init__(self, s, D2, L2, L3, specs):
        self.s, self.D2, self.L2, self.L3 = float(s), int(D2), int(L2), int(L3)
        self.specs = list(specs)
        angle_index = {}
        angles = []
        self.cand_copies = []  # per candidate: list of (angle_id, zi array, h, c)
        self.cand_tie_z = []
        mc = 1
        for sp in self.specs:
            copies = data_copies(sp.rise_pixel, sp.csym, self.L3, self.L2)
            lst = []
            tie_any = False
            col_cache = {}
            for h, c in copies:
                angle = sp.twist * h + 360 * c / sp.csym
                a = angle_index.get(angle)
                if a is None:
                    a = len(angles)
                    angle_index[angle] = a
                    angles.append(angle)
                if h not in col_cache:
                    zi, tie = column_slices(self.s, self.L2, self.L3, h * sp.rise_pixel)
                    cnt = np.bincount(zi[zi >= 0], minlength=self.L3) if np.any(zi >= 0) else np.zeros(self.L3, int)
                    col_cache[h] = (zi, tie, int(cnt.max()) if len(cnt) else 0)
                zi, tie, cmax = col_cache[h]
                tie_any |= tie
                mc = max(mc, cmax)
                lst.append((a, zi, h, c))
            self.cand_copies.append(lst)
            self.cand_tie_z.append(tie_any)
        self.MC = mc
        self.angles = np.array(angles, dtype=np.float64)
        self.cos_sin = z_rotation_entries(self.angles)
        self.finalized = False

    def finalize(self, nvalid_rays):
        nvalid_rays = np.asarray(nvalid_rays)
        L3, MC = self.L3, self.MC
        cands = np.zeros(len(self.specs), dtype=_lib.CANDIDATE_DTYPE)
        views, colk, pairs = [], [], []
        self.cand_views = []  # per candidate: list of (angle_id, zi, h, c, n_rows_real)
        for ci, sp in enumerate(self.specs):
            used = []
            n_b = 0
            for a, zi, h, c in self.cand_copies[ci]:
                ncols = int(np.count_nonzero(zi >= 0))
                nrows = ncols * int(nvalid_rays[a])
                n_b += nrows
                if nrows:
                    used.append((a, zi, h, c, nrows))
                if sp.min_projection_lines > 0 and n_b > sp.min_projection_lines:
                    break
            self.cand_views.append(used)
            cands[ci]["view_begin"] = len(views)
            cands[ci]["view_count"] = len(used)
            for a, zi, h, c, nrows in used:
                tab = np.full(L3 * MC, -1, dtype=np.int32)
                fill = np.zeros(L3, dtype=np.int64)
                for k in np.nonzero(zi >= 0)[0]:
                    z = int(zi[k])
                    tab[z * MC + fill[z]] = k
                    fill[z] += 1
                views.append((a, len(colk) * L3 * MC))
                colk.append(tab)
            # symmetry pairs (SLR:892, 1223-1243)
            cands[ci]["pair_begin"] = len(pairs)
            plist = sorted_hsym_csym_pairs(sp.twist, sp.rise_pixel, sp.csym, L3) if sp.min_sym_pairs >= 0 else []
            if plist:
                ai = np.array([sp.twist * p[-1][0][0] + p[-1][0][1] * 360 / sp.csym for p in plist])
                aj = np.array([sp.twist * p[-1][1][0] + p[-1][1][1] * 360 / sp.csym for p in plist])
                ei, ej = z_rotation_entries(ai), z_rotation_entries(aj)
                for t, p in enumerate(plist):
                    (hi, _), (hj, _) = p[-1]
                    pairs.append((ei[t, 0], ei[t, 1], sp.rise_pixel * hi, ej[t, 0], ej[t, 1], sp.rise_pixel * hj))
            cands[ci]["pair_count"] = len(plist)
            cands[ci]["min_sym_pairs"] = sp.min_sym_pairs
            cands[ci]["positive"] = int(sp.positive)
            cands[ci]["flags_in"] = _lib.HB2_FLAG_TIE_Z if self.cand_tie_z[ci] else 0
        self.cands = cands
        self.views = np.array(views, dtype=_lib.VIEW_DTYPE) if views else np.zeros(0, dtype=_lib.VIEW_DTYPE)
        self.colk = np.concatenate(colk).astype(np.int32) if colk else np.zeros(0, dtype=np.int32)
        self.pairs = np.array(pairs, dtype=_lib.PAIR_DTYPE) if pairs else np.zeros(0, dtype=_lib.PAIR_DTYPE)
        self.finalized = True
        return self
