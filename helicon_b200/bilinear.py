"""Matrix-free trilinear batches: ``interpolation="linear"`` in the grid-search case (tilt = psi = dy = 0,
scale2d_to_3d = 1), many candidates per batch (csrc/hb2_bilinear.cuh, include/helicon_b200.h).

The reference's trilinear data row of (symmetry copy, image column k, ray j) (SLR:1403-1510) factors into the in-plane
bilinear footprint of the ray -- per view angle, shared by the candidates of a batch -- times the two-slice blend of
column k.  This module is the host planner of that factorisation:

* one REGULAR map per distinct view angle;
* EXACT single-column maps where the reference's ``int()`` truncation follows the last-bit noise of its coordinate
  tables (SLR:1712-1719): views whose in-plane sample coordinates are integer-valued (angle 0 / 90 / 180 / 270 -- the
  h = 0 copy of every candidate) get one map per image column built from that column's table rows; copies with an
  integer ``h * rise_pixel`` get them for their two boundary columns (Z = -1 and Z = L3 - 1), where the slice-range
  test decides per sample;
* per view the column slots: slot t holds the column with ``int(Z) = t - 1`` and blend (1 - zf, zf); slot 0 the column
  with Z in (-1, 0), where ``int()`` truncates toward zero and the blend is (1 - Z, Z) on slices (0, 1);
* the reference's early stop over the copies (SLR:1640-1647) from the per-map ray counts the GPU reports.

Symmetry rows are the trilinear ones (SLR:910-1138), built per candidate on the GPU (16 entries per row).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, planner
from .engine import Batch, Problem, _stream_handle

BILMAP_DTYPE = np.dtype([("m00", "<f8"), ("m01", "<f8"), ("m10", "<f8"), ("m11", "<f8"), ("m22", "<f8"), ("zshift", "<f8"),
                         ("xrow", "<i4"), ("zrow", "<i4")], align=True)
assert BILMAP_DTYPE.itemsize == 56


def supported(problem: Problem, L3: int) -> bool:
    """The matrix-free factorisation applies (else: engine.ExplicitBatch)."""
    return problem.s == 1.0 and problem.D3 == problem.D2 and int(L3) <= 16


class _Plan:
    """What engine.Batch expects of a plan, for a bilinear batch."""

    def __init__(self):
        self.cands = None
        self.cand_n_data_rows = None
        self.MC = 1


class BilinearBatch(Batch):
    """Candidates (CandidateSpec) of one Problem with a common L3, trilinear interpolation, matrix-free."""

    def __init__(self, problem: Problem, L3: int, specs, stream=None):
        from scipy.spatial.transform import Rotation as R

        lib = _lib.require_gpu()
        if not supported(problem, L3):
            raise NotImplementedError("helicon_b200: matrix-free trilinear rows need scale2d_to_3d == 1, "
                                      "reconstruct_diameter_3d == 2d diameter and L3 <= 16 (use ExplicitBatch)")
        self.problem = problem
        self._stream_ref = stream
        self.L3 = L3 = int(L3)
        self.specs = list(specs)
        D2, L2 = problem.D2, problem.L2
        ZMP = (L3 + 3) // 4 * 4
        rpv = D2 * ZMP
        st = problem.stream if stream is None else _stream_handle(stream)
        dummy = np.array([[1.0, 0.0]], dtype=np.float64)
        nv1, tie1 = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
        self._h = C.c_void_p()
        _lib.check(lib.hb2_batch_begin(C.byref(self._h), problem._h, L3, 1, 1, _lib.ptr(dummy), _lib.ptr(nv1), _lib.ptr(tie1), st))
        self.nvalid, self.tie = nv1, tie1
        self._amap_cache = {}
        self.results = None
        self.n = L3 * problem.ndisk
        self.nc = len(self.specs)

        # ---- copies, angles, regular maps ------------------------------------------------------------------
        kk = np.arange(L2, dtype=np.float64) - (L2 // 2)
        angle_id = {}
        angles = []
        cand_copies = []
        for sp in self.specs:
            copies = planner.data_copies(sp.rise_pixel, sp.csym, L3, L2)
            ang = [sp.twist * h + 360 * c / sp.csym for h, c in copies]
            ids = np.empty(len(copies), dtype=np.int64)
            for i, a in enumerate(ang):
                q = angle_id.get(a)
                if q is None:
                    q = angle_id[a] = len(angles)
                    angles.append(a)
                ids[i] = q
            zsh = np.array([h * sp.rise_pixel for h, _ in copies], dtype=np.float64)
            cand_copies.append((copies, ids, zsh))
        nR = len(angles)
        M = R.from_euler("z", np.asarray(angles, dtype=np.float64).reshape(-1, 1), degrees=True).as_matrix()
        reg = np.zeros(nR, dtype=BILMAP_DTYPE)
        reg["m00"], reg["m01"], reg["m10"], reg["m11"], reg["m22"] = M[:, 0, 0], M[:, 0, 1], M[:, 1, 0], M[:, 1, 1], M[:, 2, 2]
        reg["xrow"] = -1
        reg["zrow"] = -1
        nvalid_reg = np.zeros(nR, dtype=np.int32)
        tie_reg = np.zeros(nR, dtype=np.int32)
        _lib.check(lib.hb2_batch_bilinear_maps(self._h, nR, _lib.ptr(reg), 0, None, None, 0, _lib.ptr(nvalid_reg), _lib.ptr(tie_reg), None))

        # ---- column slots per copy; exact maps ------------------------------------------------------------------
        exact_id = {}      # (angle id, k, zshift or None) -> map index
        exact_rows = []    # [(angle id, k, zshift or None)]
        tab_k = {}         # image column -> row of the coordinate tables handed to the library

        def exact_map(aid, k, zshift):
            key = (int(aid), int(k), zshift)
            q = exact_id.get(key)
            if q is None:
                q = exact_id[key] = nR + len(exact_rows)
                exact_rows.append(key)
                tab_k.setdefault(int(k), len(tab_k))
            return q

        # per candidate: list over copies of [(map, k, slot, a, b)] grouped into views
        cand_slots = []
        for (copies, ids, zsh) in cand_copies:
            m22 = reg["m22"][ids]
            Z = (m22[:, None] * kk[None, :] - zsh[:, None]) + (L3 // 2)   # same op order as the kernels (dmul, dsub, dadd)
            Zr = np.rint(Z)
            near = np.abs(Z - Zr) < 1e-9
            z_exact = np.any(near & (Z > -1.5) & (Z < L3 + 0.5), axis=1)
            per_copy = []
            for i in range(len(copies)):
                aid = int(ids[i])
                xy_exact = tie_reg[aid] > 0
                ent = []
                if z_exact[i]:
                    n = Zr[i].astype(np.int64)
                    for k in np.nonzero((n >= -1) & (n <= L3 - 1))[0]:
                        nk, Zk = int(n[k]), float(Z[i, k])
                        boundary = nk == -1 or nk == L3 - 1
                        if nk == -1:
                            slot, a, b = 0, 1.0 - Zk, Zk                   # zi = 0, zf = Z (about -1): (2, -1) on slices (0, 1)
                        elif nk == L3 - 1:
                            zf = Zk - (L3 - 2)                              # only samples with Z < L3 - 1 survive: zi = L3 - 2
                            slot, a, b = L3 - 1, 1.0 - zf, zf
                        else:
                            zf = Zk - nk
                            slot, a, b = nk + 1, 1.0 - zf, zf
                        if boundary or xy_exact:
                            mp = exact_map(aid, k, float(zsh[i]) if boundary else None)
                            ent.append((mp, int(k), slot, a, b, True))
                        else:
                            ent.append((aid, int(k), slot, a, b, False))
                else:
                    zi = np.trunc(Z[i])
                    ok = (Z[i] > -1.0) & (Z[i] < L3) & (zi + 1 <= L3 - 1)
                    for k in np.nonzero(ok)[0]:
                        Zk = float(Z[i, k])
                        z0 = int(zi[k])
                        zf = Zk - z0
                        slot = 0 if Zk < 0 else z0 + 1
                        if xy_exact:
                            ent.append((exact_map(aid, k, None), int(k), slot, 1.0 - zf, zf, True))
                        else:
                            ent.append((aid, int(k), slot, 1.0 - zf, zf, False))
                per_copy.append(ent)
            cand_slots.append(per_copy)

        # ---- all maps -----------------------------------------------------------------------------------------
        nE = len(exact_rows)
        maps = np.zeros(nR + nE, dtype=BILMAP_DTYPE)
        maps[:nR] = reg
        for q, (aid, k, zshift) in enumerate(exact_rows):
            maps[nR + q] = reg[aid]
            maps[nR + q]["xrow"] = tab_k[k]
            maps[nR + q]["zrow"] = tab_k[k] if zshift is not None else -1
            maps[nR + q]["zshift"] = 0.0 if zshift is None else zshift
        ntab = len(tab_k)
        xrows = zrows = None
        if ntab:
            Xt, Zt = planner.reference_xz_tables(problem.s, D2, L2)
            ks = np.array(sorted(tab_k, key=tab_k.get), dtype=np.int64)
            xrows = np.ascontiguousarray(Xt[ks], dtype=np.float64)
            zrows = np.ascontiguousarray(Zt[ks], dtype=np.float64)
        nvalid = np.zeros(nR + nE, dtype=np.int32)
        tie_all = np.zeros(nR + nE, dtype=np.int32)
        mhash = np.zeros(nR + nE, dtype=np.uint64)
        _lib.check(lib.hb2_batch_bilinear_maps(self._h, nR + nE, _lib.ptr(maps), ntab, _lib.ptr(xrows) if ntab else None,
                                               _lib.ptr(zrows) if ntab else None, 1, _lib.ptr(nvalid), _lib.ptr(tie_all),
                                               _lib.ptr(mhash)))
        self.maps, self.nvalid_maps, self.n_regular_maps = maps, nvalid, nR
        # exact maps with identical content (same angle by construction of the key below) share one representative
        canon = {}
        rep_of = np.arange(nR + nE)
        for q in range(nR, nR + nE):
            rep_of[q] = canon.setdefault((exact_rows[q - nR][0], int(mhash[q]), int(nvalid[q])), q)

        # ---- views per candidate with the reference's early stop (SLR:1640-1647) ----------------------------------------
        view_map, colk, ab, dup_of = [], [], [], []
        cand_nview = np.zeros(self.nc, dtype=np.int32)
        n_data_rows = np.zeros(self.nc, dtype=np.int64)
        self._row_src = []   # per candidate: [(copy index, [(view index in the candidate, map, k, slot)])] of the used copies
        for ci, sp in enumerate(self.specs):
            total = 0
            used = []
            nv = 0
            first_views = {}   # (h, c) -> first view of the first occurrence of the copy (Halton duplicates, SLR:1559-1571)
            copies_ci = cand_copies[ci][0]
            for i, ent in enumerate(cand_slots[ci]):
                rows_i = sum(int(nvalid[e[0]]) for e in ent)
                if rows_i > 0:
                    src = []
                    nv_copy0 = nv
                    regular = [e for e in ent if not e[5]]
                    if regular:
                        t_colk = np.full(ZMP, -1, dtype=np.int32)
                        t_ab = np.zeros((ZMP, 2), dtype=np.float64)
                        for (mp, k, slot, a, b, _) in regular:
                            if t_colk[slot] >= 0:
                                raise AssertionError("two columns of one copy in the same slice slot")
                            t_colk[slot] = k
                            t_ab[slot] = (a, b)
                            src.append((nv, mp, k, slot))
                        view_map.append(regular[0][0]); colk.append(t_colk); ab.append(t_ab)
                        nv += 1
                    groups = {}   # representative exact map -> the copy's columns that use it (one view per group)
                    for (mp, k, slot, a, b, ex) in ent:
                        if ex and nvalid[mp] > 0:
                            groups.setdefault(int(rep_of[mp]), []).append((k, slot, a, b))
                    for mp, cols in groups.items():
                        t_colk = np.full(ZMP, -1, dtype=np.int32)
                        t_ab = np.zeros((ZMP, 2), dtype=np.float64)
                        for (k, slot, a, b) in cols:
                            if t_colk[slot] >= 0:
                                raise AssertionError("two columns of one copy in the same slice slot")
                            t_colk[slot] = k
                            t_ab[slot] = (a, b)
                            src.append((nv, mp, k, slot))
                        view_map.append(mp); colk.append(t_colk); ab.append(t_ab)
                        nv += 1
                    # identical rows of a repeated copy: served by the first occurrence's views (same order, same count)
                    prim = first_views.setdefault(copies_ci[i], nv_copy0)
                    dup_of.extend([-1] * (nv - nv_copy0) if prim == nv_copy0 else [prim + q for q in range(nv - nv_copy0)])
                    used.append((i, src))
                total += rows_i
                if sp.min_projection_lines > 0 and total > sp.min_projection_lines:
                    break
            cand_nview[ci] = nv
            n_data_rows[ci] = total
            self._row_src.append(used)
        # ---- trilinear symmetry rows per candidate, then the pseudo views that hold them ---------------------------------
        self.m_sym = np.zeros(self.nc, dtype=np.int64)
        for ci, sp in enumerate(self.specs):
            ms = C.c_int64()
            if sp.min_sym_pairs >= 0:
                tab = planner.trilinear_pair_table(sp.twist, sp.rise_pixel, sp.csym, L3)
                _lib.check(lib.hb2_batch_bilinear_sym_rows(self._h, ci, len(tab), _lib.ptr(tab), int(sp.min_sym_pairs), C.byref(ms)))
            else:
                _lib.check(lib.hb2_batch_bilinear_sym_rows(self._h, ci, 0, None, -1, C.byref(ms)))
            self.m_sym[ci] = int(ms.value)
        # interleave: candidate c = its bilinear views, then ceil(m_sym / rpv) pseudo views
        vm_all, colk_all, ab_all, dup_all = [], [], [], []
        cands = np.zeros(self.nc, dtype=_lib.CANDIDATE_DTYPE)
        pos = 0
        vbeg = 0
        self._view0 = np.zeros(self.nc, dtype=np.int64)
        for ci, sp in enumerate(self.specs):
            nvd = int(cand_nview[ci])
            nps = int((self.m_sym[ci] + rpv - 1) // rpv)
            vm_all.extend(view_map[pos:pos + nvd]); colk_all.extend(colk[pos:pos + nvd]); ab_all.extend(ab[pos:pos + nvd])
            dup_all.extend(dup_of[pos:pos + nvd])
            for _ in range(nps):
                vm_all.append(-1); colk_all.append(np.full(ZMP, -1, dtype=np.int32)); ab_all.append(np.zeros((ZMP, 2)))
                dup_all.append(-1)
            pos += nvd
            cands[ci]["view_begin"], cands[ci]["view_count"] = vbeg, nvd + nps
            cands[ci]["pair_begin"], cands[ci]["pair_count"] = 0, 0
            cands[ci]["min_sym_pairs"] = -1
            cands[ci]["positive"] = int(sp.positive)
            cands[ci]["flags_in"] = 0
            self._view0[ci] = vbeg
            vbeg += nvd + nps
        nviews = vbeg
        if nviews == 0:
            raise _lib.HeliconB200Error("no candidate of the batch has projection data")
        vm_arr = np.ascontiguousarray(vm_all, dtype=np.int32)
        colk_arr = np.ascontiguousarray(np.stack(colk_all), dtype=np.int32)
        ab_arr = np.ascontiguousarray(np.stack(ab_all), dtype=np.float64)
        _lib.check(lib.hb2_batch_bilinear_views(self._h, nviews, _lib.ptr(vm_arr), _lib.ptr(colk_arr), _lib.ptr(ab_arr),
                                                self.nc, _lib.ptr(cand_nview)))
        views = np.zeros(nviews, dtype=_lib.VIEW_DTYPE)
        views["dup_of"] = np.asarray(dup_all, dtype=np.int32)   # relative to the candidate's first view
        views["mult"] = 1
        dcolk = np.full(L3, -1, dtype=np.int32)
        pairs = np.zeros(0, dtype=_lib.PAIR_DTYPE)
        self.plan = _Plan()
        self.plan.cands = cands
        self.plan.cand_n_data_rows = n_data_rows
        self.cand_nview = cand_nview
        self._ZMP, self._rpv = ZMP, rpv
        _lib.check(lib.hb2_batch_create(self._h, self.nc, _lib.ptr(cands), nviews, _lib.ptr(views), len(dcolk), _lib.ptr(dcolk),
                                        0, _lib.ptr(pairs)))

    # -- exports (tests) ---------------------------------------------------------------------------------------
    def ray_valid(self):
        out = np.empty((len(self.maps), self.problem.D2), dtype=np.uint8)
        _lib.check(_lib.load().hb2_batch_bilinear_ray_valid(self._h, _lib.ptr(out)))
        return out

    def data_row_index(self, c=0):
        """For every data row in the reference's order (copy, k, j): its index in the candidate's padded rows, k, j."""
        rv = self.ray_valid()
        D2, ZMP, rpv = self.problem.D2, self._ZMP, self._rpv
        idx, kk, jj = [], [], []
        for _, src in self._row_src[c]:
            for (v, mp, k, slot) in sorted(src, key=lambda e: e[2]):
                js = np.nonzero(rv[mp])[0]
                idx.append(v * rpv + js * ZMP + slot)
                kk.append(np.full(len(js), k))
                jj.append(js)
        if not idx:
            z = np.zeros(0, np.int64)
            return z, z, z
        return np.concatenate(idx), np.concatenate(kk), np.concatenate(jj)

    def sym_row_offset(self, c=0):
        """first padded row of the candidate's trilinear symmetry rows"""
        return int(self.cand_nview[c]) * self._rpv

    def sym_csr(self, c=0):
        from scipy.sparse import csr_matrix

        m = int(self.m_sym[c])
        if m == 0:
            return None, None
        cols = np.zeros(16 * m, dtype=np.int32)
        w = np.zeros(16 * m, dtype=np.float32)
        _lib.check(_lib.load().hb2_batch_bilinear_sym_export(self._h, int(c), _lib.ptr(cols), _lib.ptr(w)))
        A = csr_matrix((w, cols, np.arange(m + 1, dtype=np.int64) * 16), shape=(m, self.n), dtype=np.float32)
        A.sum_duplicates()
        return A, np.zeros(m, dtype=np.float32)

    def data_csr(self, c):
        raise NotImplementedError("a matrix-free trilinear batch has no explicit rows (engine.ExplicitBatch exports them)")
