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


def trilinear_pair_table(twist, rise_pixel, csym, nz):
    """Rows of ``hb2_batch_explicit_sym_rows``: for every pair of ``sorted_hsym_csym_pairs`` (SLR:892) the entries
    M00, M01, M10, M11, M22 of ``Rotation.from_euler('z', twist*h + 360*c/csym)`` (SLR:1225, 1235) and the z shift
    rise_pixel*h (SLR:1231, 1243) of both members -> float64 [n_pairs, 12]."""
    from scipy.spatial.transform import Rotation as R

    plist = sorted_hsym_csym_pairs(twist, rise_pixel, csym, nz)
    out = np.zeros((len(plist), 12), dtype=np.float64)
    for q, p in enumerate(plist):
        for m, (h, c) in enumerate(p[-1]):
            M = R.from_euler("z", twist * h + c * 360 / csym, degrees=True).as_matrix()
            out[q, 6 * m:6 * m + 6] = (M[0, 0], M[0, 1], M[1, 0], M[1, 1], M[2, 2], rise_pixel * h)
    return out


def z_rotation_m22(angles_deg):
    """M22 of the same scipy rotation matrices: 1.0 or 1 - 2**-53 depending on the angle."""
    from scipy.spatial.transform import Rotation as R

    angles_deg = np.atleast_1d(np.asarray(angles_deg, dtype=np.float64))
    if len(angles_deg) == 0:
        return np.zeros(0)
    return np.ascontiguousarray(R.from_euler("z", angles_deg.reshape(-1, 1), degrees=True).as_matrix()[:, 2, 2])


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


@functools.lru_cache(maxsize=16)
def reference_xz_tables(s, D2, L2):
    """x and z coordinate of sample (column k, depth i) exactly as the reference's coordinate tables hold them
    (SLR:1712-1719: ``R.from_euler('y', 90).apply(inverse=True)`` of the integer grid, then ``*= s``): the exact
    values -s*(i - D2//2) and s*(k - L2//2) plus last-bit noise that depends on (k, i) only (not on the ray j, and
    the y coordinate carries none; checked in tests/test_host_cpu.py).  The noise decides rounding ties.
    Returns two float64 arrays [L2, D2]."""
    from scipy.spatial.transform import Rotation as R

    depth = (np.arange(D2, dtype=np.int32) - D2 // 2).astype(np.float32)
    along = (np.arange(L2, dtype=np.int32) - L2 // 2).astype(np.float32)
    Zg, Yg, Xg = np.meshgrid(depth, np.zeros(1, np.float32), along, indexing="ij")
    pts = np.stack((Xg.ravel(), Yg.ravel(), Zg.ravel()), axis=1)
    pts = R.from_euler("y", 90, degrees=True).apply(pts, inverse=True)
    if s != 1.0:
        pts *= s
    tab = lambda c: np.ascontiguousarray(np.swapaxes(pts[:, c].reshape((D2, 1, L2)), 0, 2)[:, 0, :])
    return tab(0), tab(2)


def reference_z_table(s, D2, L2):
    return reference_xz_tables(s, D2, L2)[1]


class TieView:
    """A symmetry copy whose column -> slice rounding sits on a tie (h*rise_pixel half-integer): every sample picks
    its slice from the reference's noisy z table.  ``cols``: image columns with at least one sample inside [0, L3);
    ``zlo[t]``: lower slice of column t (may be -1); ``up[t, i]`` in {0, 1}: sample i of column t lands in zlo+up."""

    __slots__ = ("cols", "zlo", "up", "zt")

    def __init__(self, zt, L3):
        inside = (zt >= 0) & (zt < L3)
        self.cols = np.nonzero(inside.any(axis=1))[0]
        sub = zt[self.cols]
        self.zlo = sub.min(axis=1)
        up = sub - self.zlo[:, None]
        if up.size and up.max() > 1:
            raise AssertionError("a tie column spans more than two slices")
        self.up = up.astype(np.uint8)
        self.zt = zt


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
    stage 2 (``finalize(nvalid, angle_valid)``): row-count early stop (SLR:1647) once the GPU has reported the number
    of rays with data per angle, then flat tables.

    Everything per candidate is a handful of numpy calls over (n_h, L2) arrays (all symmetry copies at once); the
    arithmetic per element is the scalar sequence of ``column_slices``.  Copies whose column -> slice rounding is a
    tie become ``TieView``s (exact per-sample slices from the reference's z table) when ``exact_ties``.
    """

    def __init__(self, s, D2, L2, L3, specs, exact_ties=True):
        self.s, self.D2, self.L2, self.L3 = float(s), int(D2), int(L2), int(L3)
        self.specs = list(specs)
        L2, L3 = self.L2, self.L3
        kk = np.arange(L2, dtype=np.float64) - (L2 // 2)
        z0 = kk * self.s if self.s != 1.0 else kk
        angle_index = {}
        angles = []
        self._cand = []  # per candidate: (copies, angle ids, index of each copy's h in hs, hs, ZI[n_h, L2], ties {hidx: TieView})
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
            tie_h = np.any(near & (Z > -1.0) & (Z < L3), axis=1)
            exact = bool(exact_ties and tie_h.any() and L3 <= 16)
            self.cand_tie_z.append(bool(tie_h.any()) and not exact)  # approximate only when not handled exactly
            ZI = np.where(ok, zi, -1)
            reg = ~tie_h if exact else np.ones(len(hs), dtype=bool)
            okr = ok & reg[:, None]
            if okr.any():
                flat = (np.arange(len(hs))[:, None] * L3 + zi)[okr]
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
            # tie copies: the reference's z passes through Rotation.apply of the copy's z-rotation, whose M22 is
            # 1 or 1 - 2^-53 depending on the angle (quaternion normalisation) -- it decides the ties, so take it
            # from the same scipy call (SLR:1576-1577)
            ties = {}
            if exact:
                Zt = reference_z_table(self.s, self.D2, L2)
                tc = [i for i in range(len(copies)) if tie_h[hidx[i]]]
                ang = np.array([tw * copies[i][0] + 360 * copies[i][1] / cs for i in tc], dtype=np.float64)
                m22 = z_rotation_m22(ang)
                for i, m in zip(tc, m22):
                    key = (int(hidx[i]), int(copies[i][1]))
                    if key not in ties:
                        zt = np.rint(((Zt * m) - zshift[hidx[i]]) + (L3 // 2)).astype(np.int64)
                        ties[key] = TieView(zt, L3)
            self._cand.append((copies, aid, hidx, hs, ZI, ties))
        self.MC = mc
        self.angles = np.array(angles, dtype=np.float64)
        self.cos_sin = z_rotation_entries(self.angles)
        self.has_ties = any(c[5] for c in self._cand)
        self.exact_ties = bool(exact_ties)
        self._xy = [None] * len(self.specs)
        self.finalized = False

    def exact_map_requests(self, tie_counts):
        """In-plane tie views.  ``tie_counts[a]`` (from ``hb2_batch_begin``) > 0 marks the angles where a
        sample sits on a rounding boundary (twist*h = 30, 60, ... degrees); there the reference's voxel choice follows
        the noise of its x table, which differs per image column.  Every copy with such an angle is replaced by one
        single-column view per image column it uses, each with its own exact map (``hb2_batch_add_exact_maps``).
        Returns ``(cos_sin [nE, 2], x0rows [nE, D2])`` or None; the new maps get angle indices len(angles) + e."""
        tie_counts = np.asarray(tie_counts)
        nA = len(self.angles)
        if not self.exact_ties or not tie_counts[:nA].any():
            return None
        Xt = reference_xz_tables(self.s, self.D2, self.L2)[0]
        index, cs_rows, x0_rows = {}, [], []
        for ci in range(len(self.specs)):
            copies, aid, hidx, hs, ZI, ties = self._cand[ci]
            m = {}
            for i in np.nonzero(tie_counts[aid] > 0)[0]:
                if ties and (int(hidx[i]), int(copies[i][1])) in ties:
                    continue  # also a column->slice tie view: keeps that path (and the in-plane flag)
                lst = []
                for k in np.nonzero(ZI[hidx[i]] >= 0)[0]:
                    key = (int(aid[i]), int(k))
                    e = index.get(key)
                    if e is None:
                        e = index[key] = len(cs_rows)
                        cs_rows.append(self.cos_sin[aid[i]])
                        x0_rows.append(Xt[k])
                    lst.append((int(k), nA + e))
                m[int(i)] = lst
            self._xy[ci] = m or None
        if not cs_rows:
            return None
        return np.ascontiguousarray(cs_rows, dtype=np.float64), np.ascontiguousarray(x0_rows, dtype=np.float64)

    def finalize(self, nvalid_rays, angle_valid=None):
        """``angle_valid(a)`` -> bool [D2, D2] (sample (j, i) of angle a hits a voxel), needed only for tie views."""
        nvalid_rays = np.asarray(nvalid_rays, dtype=np.int64)
        L2, L3, MC, D2 = self.L2, self.L3, self.MC, self.D2
        ZMC = L3 * MC
        nc = len(self.specs)
        cands = np.zeros(nc, dtype=_lib.CANDIDATE_DTYPE)
        vrows = []   # per candidate: int64 array [n_slots, 5] = angle, tie index or -1, first tie column slot, dup_of, mult
        colk = []
        self._cv_src = []            # per candidate: what cand_views / cand_view_slots are built from (lazily)
        self._cv_explicit = []       # per candidate with tie views: explicit [(view tuple, slot)] list, else None
        self._cand_views = None
        self.cand_n_data_rows = np.zeros(nc, dtype=np.int64)
        pair_meta = []
        tie_zlo, tie_up, tie_rv = [], [], []
        nviews = 0
        npairs = 0
        karange = np.arange(L2, dtype=np.int64)[None, :]
        for ci, sp in enumerate(self.specs):
            copies, aid, hidx, hs, ZI, ties = self._cand[ci]
            valid = ZI >= 0
            ncols_h = valid.sum(axis=1)
            nrows = ncols_h[hidx] * nvalid_rays[aid]
            tie_rows = {}
            if ties:
                for i in range(len(copies)):
                    tv = ties.get((int(hidx[i]), int(copies[i][1])))
                    if tv is not None:
                        va = angle_valid(int(aid[i]))  # [j, i]
                        inside = (tv.zt[tv.cols] >= 0) & (tv.zt[tv.cols] < L3)  # [t, i]
                        rv = (va[None, :, :] & inside[:, None, :]).any(axis=2)  # [t, j]
                        tie_rows[i] = rv
                        nrows[i] = int(rv.sum())
            xy = self._xy[ci]
            if xy:
                for i, lst in xy.items():  # one single-column view per image column, each with its own exact map
                    nrows[i] = int(sum(nvalid_rays[a] for _, a in lst))
            stop = len(copies)
            if sp.min_projection_lines > 0:
                over = np.nonzero(np.cumsum(nrows) > sp.min_projection_lines)[0]
                if len(over):
                    stop = int(over[0]) + 1
            sel = np.nonzero(nrows[:stop] > 0)[0]
            vbeg = nviews
            slots = np.zeros(len(sel), dtype=np.int64)
            if len(sel):
                prev = np.concatenate([np.full((len(hs), 1), -2, dtype=np.int64), ZI[:, :-1]], axis=1)
                run_start = np.maximum.accumulate(np.where(valid & (ZI != prev), karange, 0), axis=1)
                tab = np.full((len(hs), ZMC), -1, dtype=np.int32)
                hh, kq = np.nonzero(valid)
                if ties:
                    tie_hs = {k[0] for k in ties}
                    keep = np.array([int(h) not in tie_hs for h in hh], dtype=bool)
                    hh, kq = hh[keep], kq[keep]
                tab[hh, ZI[hh, kq] * MC + (kq - run_start[hh, kq])] = kq
                explicit = None
                if not ties and not xy:
                    # all regular views: Halton duplicates (same (h, c)) are served by their first copy
                    key = hidx[sel] * sp.csym + np.array([copies[i][1] for i in sel], dtype=np.int64)
                    _, first_idx, inv = np.unique(key, return_index=True, return_inverse=True)
                    first_pos = first_idx[inv]
                    pos = np.arange(len(sel))
                    is_dup = first_pos != pos
                    vr = np.empty((len(sel), 5), dtype=np.int64)
                    vr[:, 0] = aid[sel]; vr[:, 1] = -1; vr[:, 2] = 0
                    vr[:, 3] = np.where(is_dup, first_pos, -1)
                    vr[:, 4] = np.where(is_dup, 0, np.bincount(first_pos, minlength=len(sel)))
                    vrows.append(vr)
                    colk.append(tab[hidx[sel]])
                    slots = vbeg + pos
                    nviews += len(sel)
                else:
                    first_of = {}
                    rows_c = []
                    explicit = []
                    for q_, i in enumerate(sel):
                        tv = ties.get((int(hidx[i]), int(copies[i][1]))) if ties else None
                        slots[q_] = nviews
                        if xy and int(i) in xy:
                            row = tab[hidx[i]]
                            for k, a_ex in xy[int(i)]:
                                if nvalid_rays[a_ex] == 0:
                                    continue
                                t = np.where(row == k, row, -1).astype(np.int32)
                                zi_k = np.where(np.arange(L2) == k, ZI[hidx[i]], -1)
                                explicit.append(((a_ex, zi_k, copies[i][0], copies[i][1], int(nvalid_rays[a_ex])), nviews))
                                colk.append(t[None, :])
                                rows_c.append([a_ex, -1, 0, -1, 1])
                                nviews += 1
                            continue
                        explicit.append(((int(aid[i]), ZI[hidx[i]] if tv is None else tv.zt, copies[i][0], copies[i][1],
                                         int(nrows[i])), nviews))
                        if tv is None:
                            colk.append(tab[hidx[i]][None, :])
                            rel = nviews - vbeg
                            prim = first_of.get(copies[i])
                            if prim is None:
                                first_of[copies[i]] = rel
                                rows_c.append([int(aid[i]), -1, 0, -1, 1])
                            else:  # Halton duplicate of an earlier copy: identical rows
                                rows_c[prim][4] += 1
                                rows_c.append([int(aid[i]), -1, 0, prim, 0])
                            nviews += 1
                        else:
                            ncol = len(tv.cols)
                            nslot = (ncol + ZMC - 1) // ZMC
                            tid = len(tie_zlo)
                            tie_zlo.append(tv.zlo); tie_up.append(tv.up); tie_rv.append(tie_rows[i])
                            for q in range(nslot):
                                t = np.full(ZMC, -1, dtype=np.int32)
                                cc = tv.cols[q * ZMC:(q + 1) * ZMC]
                                t[:len(cc)] = cc
                                colk.append(t[None, :])
                                rows_c.append([int(aid[i]), tid, q * ZMC, -1, 1])
                            nviews += nslot
                    vrows.append(np.array(rows_c, dtype=np.int64).reshape(-1, 5))
            self._cv_src.append((sel, slots, nrows[sel] if len(sel) else np.zeros(0, np.int64)))
            self._cv_explicit.append(explicit if len(sel) else None)
            self.cand_n_data_rows[ci] = int(nrows[sel].sum()) if len(sel) else 0
            cands[ci]["view_begin"] = vbeg
            cands[ci]["view_count"] = nviews - vbeg
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
            fl = _lib.HB2_FLAG_TIE_Z if self.cand_tie_z[ci] else 0
            if ties:
                fl |= _lib.HB2_FLAG_TIE_Z_EXACT
            cands[ci]["flags_in"] = fl
        view_rows = np.concatenate(vrows) if vrows else np.zeros((0, 5), dtype=np.int64)
        self.cands = cands
        views = np.zeros(nviews, dtype=_lib.VIEW_DTYPE)
        if nviews:
            vr = view_rows
            views["angle"] = vr[:, 0]
            views["tie"] = vr[:, 1]
            views["tie_slot0"] = vr[:, 2]
            views["dup_of"] = vr[:, 3]
            views["mult"] = vr[:, 4]
            views["col_begin"] = np.arange(nviews, dtype=np.int64) * ZMC
            self.colk = np.ascontiguousarray(np.concatenate(colk, axis=0).reshape(-1), dtype=np.int32)
        else:
            self.colk = np.zeros(0, dtype=np.int32)
        self.views = views
        # tie tables, padded to TS column slots per tie view
        self.n_tie = len(tie_zlo)
        if self.n_tie:
            TS = max(len(z) for z in tie_zlo)
            TS = (TS + ZMC - 1) // ZMC * ZMC
            self.tie_TS = TS
            self.tie_zlo = np.full((self.n_tie, TS), -100, dtype=np.int8)
            self.tie_up = np.zeros((self.n_tie, TS, D2), dtype=np.uint8)
            self.tie_rowvalid = np.zeros((self.n_tie, TS, D2), dtype=np.uint8)
            for t, (zl, up, rv) in enumerate(zip(tie_zlo, tie_up, tie_rv)):
                self.tie_zlo[t, :len(zl)] = zl
                self.tie_up[t, :len(zl)] = up
                self.tie_rowvalid[t, :len(zl)] = rv
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

    def _build_views(self):
        cvs, sls = [], []
        for ci in range(len(self.specs)):
            copies, aid, hidx, hs, ZI, ties = self._cand[ci]
            sel, slots, nrows = self._cv_src[ci]
            if self._cv_explicit[ci] is not None:
                cvs.append([e[0] for e in self._cv_explicit[ci]])
                sls.append([int(e[1]) for e in self._cv_explicit[ci]])
                continue
            cv = []
            for q, i in enumerate(sel):
                tv = ties.get((int(hidx[i]), int(copies[i][1]))) if ties else None
                cv.append((int(aid[i]), ZI[hidx[i]] if tv is None else tv.zt, copies[i][0], copies[i][1], int(nrows[q])))
            cvs.append(cv)
            sls.append([int(v) for v in slots])
        self._cand_views, self._cand_view_slots = cvs, sls

    @property
    def cand_views(self):
        """per candidate: list of (angle_id, zi (1-D, or 2-D [L2, D2] for a tie view), h, c, n_rows_real), in the
        reference's copy order incl. Halton duplicates (built on demand: exports and tests only)."""
        if self._cand_views is None:
            self._build_views()
        return self._cand_views

    @property
    def cand_view_slots(self):
        """per candidate: first view slot of every entry of ``cand_views`` (tie views take several slots; note that a
        duplicate copy has its OWN slot -- identical rows, stored twice)."""
        if self._cand_views is None:
            self._build_views()
        return self._cand_view_slots


def get_cylindrical_mask(nz, ny, nx, rmin=0, rmax=-1, return_xyz=False):
    """lib/analysis.py:731-774: boolean (nz, ny, nx) mask of the voxels with rmin^2 <= x^2 + y^2 < rmax^2 (rmax < 0:
    ny // 2 - 1); the order of its True entries is the reference's unknown order."""
    k = np.arange(0, nz, dtype=np.int32) - nz // 2
    j = np.arange(0, ny, dtype=np.int32) - ny // 2
    i = np.arange(0, nx, dtype=np.int32) - nx // 2
    Z, Y, X = np.meshgrid(k, j, i, indexing="ij")
    if rmax < 0:
        rmax = ny // 2 - 1
    r2 = X * X + Y * Y
    mask = r2 < rmax * rmax
    if 0 < rmin < rmax:
        mask &= r2 >= rmin * rmin
    return (mask, (Z, Y, X)) if return_xyz else mask
