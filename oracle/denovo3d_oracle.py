"""CPU oracle for the denovo3D solve+score hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement (numpy/scipy) of the reference's algorithm
(jianglab/helicon, ``src/helicon/webApps/denovo3D/solver_linear_regression.py``,
abbreviated SLR below, and ``src/helicon/lib/analysis.py``).  It exists to CHECK
the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` leg may import it; nothing under
``helicon_b200/`` does.

Parity status: the reference's own tests hold no numeric golden vectors for this
path, so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF, generated in the build container by ``oracle/make_golden.py``
(which imports ``/root/reference/src``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function here against them.

Third-party arithmetic that the reference calls and that is NOT in
``/root/reference`` (versions installed in this image define parity):
  * scipy 1.18.1  ``scipy.optimize.lsq_linear`` -> ``scipy.sparse.linalg.lsmr``
    and ``scipy.optimize._lsq.trf_linear`` (call site SLR:258-270).  The oracle
    calls the same scipy functions (``solve_lsq``); ``lsmr_mixed`` below
    additionally restates LSMR (Fong & Saunders 2011) with the exact precision
    map scipy executes for float32 CSR input so that the
    CUDA recurrences can be checked iteration by iteration.
  * ``scipy.spatial.transform.Rotation`` (SLR:1225-1238, 1393-1394, 1576-1577,
    1712-1717): called directly, so coordinates carry the same last-bit noise
    as the reference's.
  * ``scipy.stats.qmc.Halton`` (SLR:1566-1571, 1785-1790): restated in
    ``halton_indices`` (van der Corput base 2) and checked against scipy.
"""

from __future__ import annotations

import itertools
import math

import numpy as np
from scipy.sparse import csr_matrix, vstack

MAX_EQUATIONS = 2**26  # SLR:131


# ----------------------------------------------------------------------------
# helpers (lib/analysis.py)
# ----------------------------------------------------------------------------
def cylindrical_mask(nz, ny, nx, rmin=0, rmax=-1):
    """lib/analysis.py:731-774 ``get_cylindrical_mask`` (bool (nz,ny,nx))."""
    j = np.arange(0, ny, dtype=np.int32) - ny // 2
    i = np.arange(0, nx, dtype=np.int32) - nx // 2
    Y, X = np.meshgrid(j, i, indexing="ij")
    if rmax < 0:
        rmax = ny // 2 - 1
    m2 = X * X + Y * Y < rmax * rmax
    if 0 < rmin < rmax:
        m2 &= X * X + Y * Y >= rmin * rmin
    return np.broadcast_to(m2, (nz, ny, nx)).copy()


def cosine_similarity(a, b):
    """lib/analysis.py:802-821."""
    norm = np.linalg.norm(a) * np.linalg.norm(b)
    if norm == 0:
        return 0
    return np.sum(a * b) / norm


def halton_indices(n):
    """Restates ``qmc.Halton(d=1, scramble=False).integers(0, n, n=n)``
    (SLR:1566-1571): floor(n * vdc2(i)) for i = 0..n-1 (not a permutation:
    it has duplicates and omissions)."""
    out = np.empty(n, dtype=np.int64)
    for i in range(n):
        f, r, k = 0.5, 0.0, i
        while k:
            if k & 1:
                r += f
            k >>= 1
            f *= 0.5
        out[i] = int(math.floor(r * n))
    return out


def data_copies(rise_pixel, csym, L3, L2):
    """Ordered (h, c) symmetry copies of build_A_data_matrix (SLR:1559-1571)."""
    hsym_max = max(1, int(np.ceil(L3 + L2) / 2 / rise_pixel))
    hc = list(itertools.product(range(-hsym_max, hsym_max + 1), range(csym)))
    hc.sort(key=lambda x: (abs(x[0]), x[1]))
    return [hc[int(i)] for i in halton_indices(len(hc))]


def sorted_hsym_csym_pairs(twist, rise, csym, nz):
    """SLR:1749-1791."""
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
    return [out[int(i)] for i in halton_indices(len(out))]


# ----------------------------------------------------------------------------
# geometry (SLR:1657-1746)
# ----------------------------------------------------------------------------
def back_project_2d_coords_to_3d_coords(image, scale2d_to_3d, D2=-1, L2=-1):
    """SLR:1657-1746.  Returns ((X,Y,Z) each (L2,D2,D2) float64, pixel_vals (D2,L2))."""
    from scipy.spatial.transform import Rotation as R

    ny, nx = image.shape
    if D2 <= 0:
        D2 = ny
    if L2 <= 0:
        L2 = nx
    D2 = int(np.rint(D2))
    L2 = int(np.rint(L2))
    k = np.arange(0, D2, dtype=np.int32) - D2 // 2
    j = np.arange(0, D2, dtype=np.int32) - D2 // 2
    i = np.arange(0, L2, dtype=np.int32) - L2 // 2
    pix = image[np.ix_(j + ny // 2, i + nx // 2)]
    Z, Y, X = np.meshgrid(
        k.astype(np.float32), j.astype(np.float32), i.astype(np.float32), indexing="ij"
    )
    coords = np.vstack((X.ravel(), Y.ravel(), Z.ravel())).transpose()
    coords = R.from_euler("y", 90, degrees=True).apply(coords, inverse=True)
    if scale2d_to_3d != 1.0:
        coords *= scale2d_to_3d
    out = []
    for c in range(3):
        out.append(np.swapaxes(coords[:, c].reshape((D2, D2, L2)), 0, 2))
    return tuple(out), pix


def _disk_index(nz, ny, nx, rmin, rmax):
    mask = cylindrical_mask(nz, ny, nx, rmin, rmax)
    n_x = int(np.count_nonzero(mask))
    idx = np.zeros(mask.shape, dtype=np.int64) - 1
    idx[np.nonzero(mask)] = np.arange(n_x)
    return mask, idx, n_x


# ----------------------------------------------------------------------------
# forward model rows (SLR:1301-1654)
# ----------------------------------------------------------------------------
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
    return_blocks=False,
):
    """Literal (full 3-D coordinate table) restatement of SLR:1301-1654 with the
    numba triple loop (SLR:1403-1557) vectorised per symmetry copy.  Supports
    general tilt/psi/dy.  Row order, duplicate summation and the early stop are
    the reference's (cpu=1 path)."""
    from scipy.spatial.transform import Rotation as R

    (X0, Y0, Z0), pix = back_project_2d_coords_to_3d_coords(
        image,
        scale2d_to_3d,
        reconstruct_diameter_2d_pixel,
        reconstruct_length_2d_pixel,
    )
    rmin = reconstruct_diameter_3d_inner_pixel / 2
    rmax = reconstruct_diameter_3d_pixel // 2 - 1
    nz, ny, nx = X0.shape
    L3 = reconstruct_length_3d_pixel
    if L3 <= 0:
        L3 = nz
    mask, idx, n_x = _disk_index(L3, ny, nx, rmin, rmax)
    coords0 = np.vstack((X0.ravel(), Y0.ravel(), Z0.ravel())).transpose()
    coords0[:, 1] -= dy_pixel
    coords0 = R.from_euler("yx", (tilt_degree, psi_degree), degrees=True).apply(
        coords0, inverse=True
    )
    linear = interpolation in ["linear", "linear10", "linear11"]
    blocks, bs, pids, used = [], [], [], []
    n_b = 0
    kk, jj = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij")
    for hi, ci in data_copies(rise_pixel, csym, L3, nz):
        angle = twist_degree * hi + 360 * ci / csym
        coords = R.from_euler("z", angle, degrees=True).apply(coords0, inverse=True)
        coords[:, 2] -= hi * rise_pixel
        X = coords[:, 0].reshape((nz, ny, nx)) + nx // 2
        Y = coords[:, 1].reshape((nz, ny, nx)) + ny // 2
        Z = coords[:, 2].reshape((nz, ny, nx)) + L3 // 2
        if linear:
            A_blk, rows_kj = _rows_linear(Z, Y, X, mask, idx, n_x)
        else:
            A_blk, rows_kj = _rows_nn(Z, Y, X, mask, idx, n_x)
        nrow = len(rows_kj)
        n_b += nrow
        if nrow:
            k_r, j_r = rows_kj // ny, rows_kj % ny
            blocks.append(A_blk)
            bs.append(pix[j_r, k_r].astype(np.float32))
            pids.append((k_r * ny + j_r).astype(np.int32))
            used.append((hi, ci, nrow))
        if min_projection_lines > 0 and n_b > min_projection_lines:
            break
    A = vstack(blocks).tocsr()
    b = np.concatenate(bs, dtype=np.float32)
    b_pid = np.concatenate(pids)
    if return_blocks:
        return A, b, b_pid, used
    return A, b, b_pid


def build_A_data_matrix_fast(image, scale2d_to_3d, twist_degree, rise_pixel, csym, D2, L2, D3, D3_inner, L3,
                             min_projection_lines):
    """nn data rows for tilt=psi=dy=0 from per-copy 2-D tables: with no tilt every
    ray lies in one z-slice and the in-plane sample->voxel map is the same for all
    image columns of a copy.  Identical to
    ``build_A_data_matrix`` whenever no rounding decision sits within 1e-9 of a
    boundary; otherwise this function defers to the literal builder.
    Used for mid-size parity tests and as the CPU baseline's matrix builder (it is
    faster than the reference's own numba builder, so the baseline is not
    handicapped)."""
    from scipy.spatial.transform import Rotation as R

    s = float(scale2d_to_3d)
    ny_i, nx_i = image.shape
    pix = image[np.ix_(np.arange(D2) - D2 // 2 + ny_i // 2, np.arange(L2) - L2 // 2 + nx_i // 2)]
    rmin, rmax = D3_inner / 2, D3 // 2 - 1
    mask, idx, n_x = _disk_index(1, D2, D2, rmin, rmax)
    rank2d = idx[0]
    nd = n_x
    c0 = D2 // 2
    x0 = -(np.arange(D2, dtype=np.float64) - c0)
    y0 = np.arange(D2, dtype=np.float64) - c0
    kl = np.arange(L2, dtype=np.float64) - L2 // 2
    if s != 1.0:
        x0, y0, kl = x0 * s, y0 * s, kl * s
    Yg, Xg = np.meshgrid(y0, x0, indexing="ij")  # (j, i)
    pts = np.stack([Xg.ravel(), Yg.ravel(), np.zeros(D2 * D2)], axis=1)
    near = lambda V: np.abs(np.abs(V - np.floor(V)) - 0.5) < 1e-9
    ind_parts, cnt_parts, bs, pids = [], [], [], []
    r0 = 0
    for hi, ci in data_copies(rise_pixel, csym, L3, L2):
        angle = twist_degree * hi + 360 * ci / csym
        q = R.from_euler("z", angle, degrees=True).apply(pts, inverse=True)
        X = q[:, 0].reshape(D2, D2) + c0
        Y = q[:, 1].reshape(D2, D2) + c0
        Z = (kl - hi * rise_pixel) + L3 // 2
        inside = (X > -1) & (X < D2) & (Y > -1) & (Y < D2)
        if np.any((near(X) | near(Y)) & inside) or np.any(near(Z) & (Z > -1) & (Z < L3)):
            return build_A_data_matrix(image, s, twist_degree, rise_pixel, csym, 0, 0, 0, D2, L2, D3, D3_inner, L3,
                                       min_projection_lines, "nn")
        xi, yi, zi = np.rint(X).astype(np.int64), np.rint(Y).astype(np.int64), np.rint(Z).astype(np.int64)
        ok = (xi >= 0) & (xi <= D2 - 1) & (yi >= 0) & (yi <= D2 - 1)
        vox = np.full((D2, D2), -1, dtype=np.int64)
        vox[ok] = rank2d[yi[ok], xi[ok]]
        hit = vox >= 0
        js = np.nonzero(hit.any(axis=1))[0]
        ks = np.nonzero((zi >= 0) & (zi <= L3 - 1))[0]
        nrow = len(js) * len(ks)
        if nrow:
            sub = vox[js]
            h2 = sub >= 0
            per_ray = h2.sum(axis=1).astype(np.int64)
            v = sub[h2].astype(np.int32)  # ray-major, depth order: already CSR order within a column block
            for k in ks:
                ind_parts.append(v + np.int32(zi[k] * nd))
                cnt_parts.append(per_ray)
            bs.append(pix[np.ix_(js, ks)].T.ravel().astype(np.float32))
            pids.append((ks[:, None] * D2 + js[None, :]).ravel().astype(np.int32))
            r0 += nrow
        if min_projection_lines > 0 and r0 > min_projection_lines:
            break
    indices = np.concatenate(ind_parts)
    indptr = np.zeros(r0 + 1, dtype=np.int64)
    np.cumsum(np.concatenate(cnt_parts), out=indptr[1:])
    A = csr_matrix((np.ones(len(indices), dtype=np.float32), indices, indptr), shape=(r0, L3 * nd), dtype=np.float32)
    A.sum_duplicates()  # the reference's COO->CSR conversion sums duplicate samples of a ray
    return A, np.concatenate(bs), np.concatenate(pids)


def _rows_nn(Z, Y, X, mask, idx, n_x):
    """SLR:1514-1557 (half-to-even rounding like numba's round())."""
    nz, ny, nx = Z.shape
    mz, my, mx = mask.shape
    zi = np.rint(Z).astype(np.int64)
    yi = np.rint(Y).astype(np.int64)
    xi = np.rint(X).astype(np.int64)
    ok = (zi >= 0) & (zi <= mz - 1) & (yi >= 0) & (yi <= my - 1) & (xi >= 0) & (xi <= mx - 1)
    col = np.full(Z.shape, -1, dtype=np.int64)
    col[ok] = idx[zi[ok], yi[ok], xi[ok]]
    hit = col >= 0
    ray_has = hit.reshape(nz * ny, nx).any(axis=1)
    rows_kj = np.nonzero(ray_has)[0]
    rank = np.cumsum(ray_has) - 1
    ray_of_sample = np.repeat(np.arange(nz * ny), nx).reshape(Z.shape)
    r = rank[ray_of_sample[hit]]
    c = col[hit]
    A = csr_matrix(
        (np.ones(len(r), dtype=np.float32), (r, c)),
        shape=(len(rows_kj), n_x),
        dtype=np.float32,
    )
    return A, rows_kj


def _rows_linear(Z, Y, X, mask, idx, n_x):
    """SLR:1403-1510: int() truncation corner, all 8 corners in range and in
    mask, trilinear weights accumulated per row in float64 then stored f32."""
    nz, ny, nx = Z.shape
    mz, my, mx = mask.shape
    zi = np.trunc(Z).astype(np.int64)
    yi = np.trunc(Y).astype(np.int64)
    xi = np.trunc(X).astype(np.int64)
    ok = (
        (zi >= 0) & (zi + 1 <= mz - 1) & (yi >= 0) & (yi + 1 <= my - 1) & (xi >= 0) & (xi + 1 <= mx - 1)
    )
    z_, y_, x_ = zi[ok], yi[ok], xi[ok]
    allin = np.ones(len(z_), dtype=bool)
    for dz, dy, dx in itertools.product((0, 1), (0, 1), (0, 1)):
        allin &= mask[z_ + dz, y_ + dy, x_ + dx]
    okk = np.zeros(Z.shape, dtype=bool)
    okk[ok] = allin
    ray_has = okk.reshape(nz * ny, nx).any(axis=1)
    rows_kj = np.nonzero(ray_has)[0]
    rank = np.cumsum(ray_has) - 1
    ray_of_sample = np.repeat(np.arange(nz * ny), nx).reshape(Z.shape)
    r = rank[ray_of_sample[okk]]
    z_, y_, x_ = zi[okk], yi[okk], xi[okk]
    zf, yf, xf = Z[okk] - z_, Y[okk] - y_, X[okk] - x_
    rr, cc, dd = [], [], []
    for dz, dy, dx in itertools.product((0, 1), (0, 1), (0, 1)):
        w = (zf if dz else 1 - zf) * (yf if dy else 1 - yf) * (xf if dx else 1 - xf)
        rr.append(r)
        cc.append(idx[z_ + dz, y_ + dy, x_ + dx])
        dd.append(w)
    rr, cc, dd = np.concatenate(rr), np.concatenate(cc), np.concatenate(dd)
    # the reference accumulates per-row weights in a float64 dict, then stores f32
    A64 = csr_matrix((dd, (rr, cc)), shape=(len(rows_kj), n_x), dtype=np.float64)
    A64.sum_duplicates()
    return A64.astype(np.float32), rows_kj


# ----------------------------------------------------------------------------
# symmetry rows (SLR:844-1298)
# ----------------------------------------------------------------------------
def build_A_helical_sym_matrix(
    nz, ny, nx, twist_degree, rise_pixel, csym, rmin, rmax, min_sym_pairs, interpolation, verbose=0
):
    """SLR:844-1298.  nn: rows +1@a -1@b with first-seen-wins de-duplication of
    unordered (a,b) across pairs (list order) then voxels (mask order);
    linear: truncation corners, |dz|,|dy|,|dx|>=3, the corner-110 weight typo
    of SLR:1089/1125 reproduced."""
    from scipy.spatial.transform import Rotation as R

    pairs = sorted_hsym_csym_pairs(twist_degree, rise_pixel, csym, nz)
    mask, idx, n_x = _disk_index(nz, ny, nx, rmin, rmax)
    kz, jy, ix = np.nonzero(mask)
    xyz = np.stack([ix - nx // 2, jy - ny // 2, kz - nz // 2], axis=1).astype(np.float64)
    linear = interpolation in ["linear", "linear01", "linear11"]
    seen = set() if linear else {}
    blocks = []
    row_count = 0

    def image_of(h, c):
        t = R.from_euler("z", twist_degree * h + c * 360 / csym, degrees=True).apply(xyz, inverse=False)
        return t[:, 2] + nz // 2 + rise_pixel * h, t[:, 1] + ny // 2, t[:, 0] + nx // 2

    for p in pairs:
        (hi, ci), (hj, cj) = p[-1]
        Zi, Yi, Xi = image_of(hi, ci)
        Zj, Yj, Xj = image_of(hj, cj)
        if linear:
            blk, nrow = _hsym_rows_linear(Zi, Yi, Xi, Zj, Yj, Xj, mask, idx, n_x, seen)
        else:
            blk, nrow = _hsym_rows_nn(Zi, Yi, Xi, Zj, Yj, Xj, mask, idx, n_x, seen)
        row_count += nrow
        if nrow:
            blocks.append(blk)
        if row_count >= min_sym_pairs:
            break
    if blocks:
        return vstack(blocks).tocsr(), np.zeros(row_count, dtype=np.float32)
    return None, None


def _valid_index(Zr, Yr, Xr, mask, idx):
    mz, my, mx = mask.shape
    ok = (Zr >= 0) & (Zr <= mz - 1) & (Yr >= 0) & (Yr <= my - 1) & (Xr >= 0) & (Xr <= mx - 1)
    q = np.full(len(Zr), -1, dtype=np.int64)
    q[ok] = idx[Zr[ok], Yr[ok], Xr[ok]]
    return q


def _dedup_first_seen(a, b, n, seen):
    """First-seen-wins on unordered (a,b) (SLR:1197-1202): a candidate row is kept
    iff its unordered key was not produced by an earlier pair or an earlier voxel
    of this pair.  ``seen`` is a dict holding one sorted int64 array of keys
    (vectorised equivalent of the reference's Python set + sequential loop)."""
    lo, hi = np.minimum(a, b).astype(np.int64), np.maximum(a, b).astype(np.int64)
    key = lo * np.int64(n) + hi
    old = seen.get("keys")
    fresh = np.ones(len(key), dtype=bool) if old is None or len(old) == 0 else ~np.isin(key, old, assume_unique=False)
    uniq, first = np.unique(key, return_index=True)
    is_first = np.zeros(len(key), dtype=bool)
    is_first[first] = True
    keep = fresh & is_first
    newk = key[keep]
    seen["keys"] = newk if old is None else np.union1d(old, newk)
    return keep


def _hsym_rows_nn(Zi, Yi, Xi, Zj, Yj, Xj, mask, idx, n_x, seen):
    r_ = lambda v: np.rint(v).astype(np.int64)
    a = _valid_index(r_(Zi), r_(Yi), r_(Xi), mask, idx)
    b = _valid_index(r_(Zj), r_(Yj), r_(Xj), mask, idx)
    cand = np.nonzero((a >= 0) & (b >= 0))[0]
    keep = _dedup_first_seen(a[cand], b[cand], n_x, seen)
    sel = cand[keep]
    nrow = len(sel)
    if nrow == 0:
        return None, 0
    rows = np.repeat(np.arange(nrow), 2)
    cols = np.stack([a[sel], b[sel]], axis=1).ravel()
    vals = np.tile(np.array([1, -1], dtype=np.float32), nrow)
    return csr_matrix((vals, (rows, cols)), shape=(nrow, n_x), dtype=np.float32), nrow


def _hsym_rows_linear(Zi, Yi, Xi, Zj, Yj, Xj, mask, idx, n_x, seen):
    mz, my, mx = mask.shape
    t_ = lambda v: np.trunc(v).astype(np.int64)
    r_ = lambda v: np.rint(v).astype(np.int64)

    def corners_ok(Z, Y, X):
        z, y, x = t_(Z), t_(Y), t_(X)
        ok = (z >= 0) & (z + 1 <= mz - 1) & (y >= 0) & (y + 1 <= my - 1) & (x >= 0) & (x + 1 <= mx - 1)
        res = np.zeros(len(Z), dtype=bool)
        zz, yy, xx = z[ok], y[ok], x[ok]
        allin = np.ones(len(zz), dtype=bool)
        for dz, dy, dx in itertools.product((0, 1), (0, 1), (0, 1)):
            allin &= mask[zz + dz, yy + dy, xx + dx]
        res[ok] = allin
        return res, z, y, x

    oki, zi, yi, xi = corners_ok(Zi, Yi, Xi)
    okj, zj, yj, xj = corners_ok(Zj, Yj, Xj)
    far = ~((np.abs(zi - zj) < 3) | (np.abs(yi - yj) < 3) | (np.abs(xi - xj) < 3))
    cand = np.nonzero(oki & okj & far)[0]
    # dedup key from ROUNDED indices looked up WITHOUT bounds checks (SLR:1045-1058);
    # numba wraps negative indices like numpy, so emulate with modular indexing.
    ir = idx[r_(Zi[cand]) % mz, r_(Yi[cand]) % my, r_(Xi[cand]) % mx]
    jr = idx[r_(Zj[cand]) % mz, r_(Yj[cand]) % my, r_(Xj[cand]) % mx]
    n_indices = n_x
    keep = np.zeros(len(cand), dtype=bool)
    for t in range(len(cand)):
        key = int(ir[t]) * n_indices + int(jr[t])
        if key in seen:
            continue
        seen.add(key)
        seen.add(int(jr[t]) * n_indices + int(ir[t]))
        keep[t] = True
    sel = cand[keep]
    nrow = len(sel)
    if nrow == 0:
        return None, 0
    rr, cc, dd = [], [], []
    for sign, (Z, Y, X, z, y, x) in ((1.0, (Zi, Yi, Xi, zi, yi, xi)), (-1.0, (Zj, Yj, Xj, zj, yj, xj))):
        zf, yf, xf = Z[sel] - z[sel], Y[sel] - y[sel], X[sel] - x[sel]
        w = {
            (0, 0, 0): (1 - zf) * (1 - yf) * (1 - xf),
            (0, 0, 1): (1 - zf) * (1 - yf) * xf,
            (0, 1, 0): (1 - zf) * yf * (1 - xf),
            (0, 1, 1): (1 - zf) * yf * xf,
            (1, 0, 0): zf * (1 - yf) * (1 - xf),
            (1, 0, 1): zf * (1 - yf) * xf,
            (1, 1, 0): xf * yf * (1 - xf),  # reference typo (SLR:1089, 1125), kept
            (1, 1, 1): xf * yf * zf,
        }
        for (dz, dy, dx), wv in w.items():
            rr.append(np.arange(nrow))
            cc.append(idx[z[sel] + dz, y[sel] + dy, x[sel] + dx])
            dd.append((sign * wv).astype(np.float32))
    A = csr_matrix(
        (np.concatenate(dd), (np.concatenate(rr), np.concatenate(cc))), shape=(nrow, n_x), dtype=np.float32
    )
    return A, nrow


# ----------------------------------------------------------------------------
# solve + score (SLR:31-547)
# ----------------------------------------------------------------------------
def split_A_b(A, b, b_id, mode):
    """SLR:175-203: the two half sets of the data equations by image pixel id."""
    uniq = sorted(set(b_id))
    n = len(uniq)
    if mode == 1:  # random halves: the reference shuffles list(set(b_id)) with the global numpy RNG
        uniq = list(set(b_id))
        np.random.shuffle(uniq)
        first = uniq[: n // 2]
    elif mode == 2:
        first = uniq[::2]
    elif mode == 3:
        first = uniq[: n // 2]
    else:
        first = uniq[: n // 3] + uniq[n * 2 // 3:]
    sel = np.isin(b_id, first)
    return (A[sel], b[sel]), (A[~sel], b[~sel])


def positive_rule(positive_constraint, rise_pixel, twist_degree, L3):
    """SLR:352-355."""
    pitch_pixel = round(rise_pixel * 360 / abs(twist_degree))
    return positive_constraint > 0 or (positive_constraint < 0 and pitch_pixel > round(L3 * 2))


def solve_lsq(A_data, b_data, A_hsym, b_hsym, positive):
    """SLR:218-270 with algorithm={'model':'lsq'}: the reference's own scipy call."""
    from scipy.optimize import lsq_linear

    if not (A_hsym is None or b_hsym is None):
        A = vstack((A_data, A_hsym))
        b = np.concatenate((b_data, b_hsym))
    else:
        A, b = A_data, b_data
    if positive:
        lb, ub = 0.0, np.max(b_data)
    else:
        lb, ub = -np.inf, np.inf
    res = lsq_linear(A, b, bounds=(lb, ub), tol=1e-2, max_iter=200, lsmr_maxiter=1000, lsmr_tol="auto", verbose=0)
    return res.x.astype(np.float32), res


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
    return_details=False,
    fast=False,
    fsc_test=0,
):
    """SLR:31-547 for algorithm={'model':'lsq'}, score_metric='cosine', no tilt/psi/dy
    refinement.  ``fast=True`` uses ``build_A_data_matrix_fast``.  ``fsc_test >= 1`` adds the
    two half-set solves (SLR:441-482) and the combined score (SLR:527-528)."""
    D3, L3 = reconstruct_diameter_3d_pixel, reconstruct_length_3d_pixel
    rmin = reconstruct_diameter_3d_inner_pixel / 2
    rmax = D3 // 2 - 1
    mask = cylindrical_mask(L3, D3, D3, rmin, rmax)
    n3 = int(np.count_nonzero(mask))
    n2 = reconstruct_diameter_2d_pixel * reconstruct_length_2d_pixel
    target = min(MAX_EQUATIONS, int(max(n2, n3) * sym_oversample))
    if fast and interpolation == "nn" and tilt_degree == 0 and psi_degree == 0 and dy_pixel == 0:
        A_data, b_data, b_pid = build_A_data_matrix_fast(
            projection_image, scale2d_to_3d, twist_degree, rise_pixel, csym, reconstruct_diameter_2d_pixel,
            reconstruct_length_2d_pixel, D3, reconstruct_diameter_3d_inner_pixel, L3, target)
    else:
        A_data, b_data, b_pid = build_A_data_matrix(
            projection_image,
            scale2d_to_3d,
            twist_degree,
            rise_pixel,
            csym,
            tilt_degree,
            psi_degree,
            dy_pixel,
            reconstruct_diameter_2d_pixel,
            reconstruct_length_2d_pixel,
            D3,
            reconstruct_diameter_3d_inner_pixel,
            L3,
            target,
            interpolation,
        )
    A_hsym, b_hsym = build_A_helical_sym_matrix(
        L3, D3, D3, twist_degree, rise_pixel, csym, rmin, rmax, target, interpolation
    )
    positive = positive_rule(positive_constraint, rise_pixel, twist_degree, L3)
    x, res = solve_lsq(A_data, b_data, A_hsym, b_hsym, positive)
    pred = A_data.dot(x)
    if thresh_fraction >= 0:
        pred = np.clip(pred, 0, None)
    score = cosine_similarity(pred, b_data)
    rec3d = np.zeros((L3, D3, D3), dtype=np.float32)
    rec3d[mask] = x
    if fsc_test >= 1:
        halves, scores = [], [score]
        for A_h, b_h in split_A_b(A_data, b_data, b_pid, fsc_test):
            x_h, _ = solve_lsq(A_h, b_h, A_hsym, b_hsym, positive)
            pred = A_h.dot(x_h)
            if thresh_fraction >= 0:
                pred = np.clip(pred, 0, None)
            scores.append(cosine_similarity(pred, b_h))
            vol = np.zeros((L3, D3, D3), dtype=np.float32)
            vol[mask] = x_h
            halves.append(vol)
        return (rec3d, halves[0], halves[1]), scores[0] / 2 + (scores[1] + scores[2]) / 4
    if return_details:
        return (rec3d, None, None), score, dict(
            A_data=A_data, b_data=b_data, b_pid=b_pid, A_hsym=A_hsym, x=x, res=res, positive=positive
        )
    return (rec3d, None, None), score


# ----------------------------------------------------------------------------
# LSMR with scipy's executed precision map, for
# iteration-by-iteration checks of the CUDA recurrences.
# ----------------------------------------------------------------------------
def _sym_ortho32(a, b):
    """scipy lsqr.py:62-94 with float32 in/out, math.sqrt in float64."""
    f = np.float32
    if b == 0:
        return f(np.sign(a)), f(0), f(abs(a))
    elif a == 0:
        return f(0), f(np.sign(b)), f(abs(b))
    elif abs(b) > abs(a):
        tau = f(a / b)
        s = f(np.sign(b) / math.sqrt(f(1 + tau * tau)))
        c = f(s * tau)
        r = f(b / s)
    else:
        tau = f(b / a)
        c = f(np.sign(a) / math.sqrt(f(1 + tau * tau)))
        s = f(c * tau)
        r = f(a / c)
    return c, s, r


def lsmr_mixed(A, b, atol=1e-4, btol=1e-4, conlim=1e8, maxiter=1000, fixed_iters=None, trace=None, norm="blas"):
    """scipy ``lsmr`` (lsmr.py:197-480) for float32 CSR ``A`` and float32 ``b``,
    damp=0, x0=None: u,v,h float32; x,hbar float64; scalars float32.  Calls the
    real scipy on the same input give the same iterates up to summation order.
    ``fixed_iters`` runs exactly that many iterations ignoring stop tests.

    ``norm="blas"`` is what scipy executes: ``numpy.linalg.norm`` of a float32 vector = OpenBLAS ``sdot`` with float32
    accumulators, whose rounding grows with the vector length (measured here: -7.6e-6 relative at 6 M elements, -2.2e-5
    at 12 M) and depends on the CPU kernel OpenBLAS dispatches and on its thread count.  ``norm="exact"`` rounds the
    exactly accumulated (float64) norm to float32 once -- the value scipy's formula MEANS; the difference between the two
    variants is the reference's own host-dependent float32 BLAS noise, amplified by the Lanczos process."""
    f = np.float32
    if callable(norm):
        nrm = norm
    elif norm == "chain":  # the CUDA kernel's accumulation (oracle/blas_sdot.py:chain_sumsq), for CPU-side checks
        from oracle.blas_sdot import chain_sumsq

        nrm = lambda q: math.sqrt(float(chain_sumsq(q)))
    elif norm == "exact":
        nrm = lambda q: math.sqrt(float(np.dot(q.astype(np.float64), q.astype(np.float64))))
    else:
        nrm = np.linalg.norm
    A = A.tocsr()
    AT = A.T.tocsr()
    m, n = A.shape
    u = b.astype(f).copy()
    normb = f(nrm(u))
    x = np.zeros(n, np.float64)
    beta = f(normb)
    if beta > 0:
        u = f(1 / beta) * u
        v = AT.dot(u)
        alpha = f(nrm(v))
    else:
        v = np.zeros(n, f)
        alpha = f(0)
    if alpha > 0:
        v = f(1 / alpha) * v
    itn = 0
    zetabar = f(alpha * beta)
    alphabar = alpha
    rho = rhobar = cbar = f(1)
    sbar = f(0)
    h = v.copy()
    hbar = np.zeros(n, np.float64)
    betadd, betad = beta, f(0)
    rhodold, tautildeold, thetatilde, zeta, d = f(1), f(0), f(0), f(0), f(0)
    normA2 = f(alpha * alpha)
    maxrbar, minrbar = f(0), 1e100
    normA = math.sqrt(normA2)
    condA, normx = 1, 0.0
    istop = 0
    ctol = 1 / conlim if conlim > 0 else 0
    normr = beta
    normar = f(alpha * beta)
    if normar == 0 or normb == 0:
        return x, istop, itn, normr, normar, normA, condA, normx
    limit = maxiter if fixed_iters is None else fixed_iters
    with np.errstate(over="ignore"):
        while itn < limit:
            itn += 1
            u *= -alpha
            u += A.dot(v)
            beta = f(nrm(u))
            if beta > 0:
                u *= f(1 / beta)
                v *= -beta
                v += AT.dot(u)
                alpha = f(nrm(v))
                if alpha > 0:
                    v *= f(1 / alpha)
            chat, shat, alphahat = _sym_ortho32(alphabar, 0.0)
            rhoold = rho
            c, s, rho = _sym_ortho32(alphahat, beta)
            thetanew = f(s * alpha)
            alphabar = f(c * alpha)
            rhobarold, zetaold = rhobar, zeta
            thetabar = f(sbar * rho)
            rhotemp = f(cbar * rho)
            cbar, sbar, rhobar = _sym_ortho32(f(cbar * rho), thetanew)
            zeta = f(cbar * zetabar)
            zetabar = f(-sbar * zetabar)
            hbar *= np.float64(f(-(f(thetabar * rho) / f(rhoold * rhobarold))))
            hbar += h
            x += np.float64(f(zeta / f(rho * rhobar))) * hbar
            h *= f(-(thetanew / rho))
            h += v
            betaacute = f(chat * betadd)
            betacheck = f(-shat * betadd)
            betahat = f(c * betaacute)
            betadd = f(-s * betaacute)
            thetatildeold = thetatilde
            ctildeold, stildeold, rhotildeold = _sym_ortho32(rhodold, thetabar)
            thetatilde = f(stildeold * rhobar)
            rhodold = f(ctildeold * rhobar)
            betad = f(f(-stildeold * betad) + f(ctildeold * betahat))
            tautildeold = f(f(zetaold - f(thetatildeold * tautildeold)) / rhotildeold)
            taud = f(f(zeta - f(thetatilde * tautildeold)) / rhodold)
            d = f(d + f(betacheck * betacheck))
            dd = f(betad - taud)
            normr = math.sqrt(f(f(d + f(dd * dd)) + f(betadd * betadd)))
            normA2 = f(normA2 + f(beta * beta))
            normA = math.sqrt(normA2)
            normA2 = f(normA2 + f(alpha * alpha))
            maxrbar = max(maxrbar, rhobarold)
            if itn > 1:
                minrbar = min(minrbar, rhobarold)
            condA = max(maxrbar, rhotemp) / min(minrbar, rhotemp)
            normar = abs(zetabar)
            normx = np.linalg.norm(x)
            test1 = normr / normb
            test2 = normar / (normA * normr) if (normA * normr) != 0 else np.inf
            test3 = 1 / condA
            t1 = test1 / (1 + normA * normx / normb)
            rtol = btol + atol * normA * normx / normb
            if trace is not None:
                trace.append(dict(itn=itn, alpha=float(alpha), beta=float(beta), rho=float(rho), rhobar=float(rhobar),
                                  zeta=float(zeta), normr=float(normr), normA=float(normA), normx=float(normx),
                                  test1=float(test1), test2=float(test2)))
            if fixed_iters is not None:
                continue
            if itn >= maxiter:
                istop = 7
            if 1 + test3 <= 1:
                istop = 6
            if 1 + test2 <= 1:
                istop = 5
            if 1 + t1 <= 1:
                istop = 4
            if test3 <= ctol:
                istop = 3
            if test2 <= atol:
                istop = 2
            if test1 <= rtol:
                istop = 1
            if istop > 0:
                break
    return x, istop, itn, normr, normar, normA, condA, normx


# ----------------------------------------------------------------------------
# Bounded branch restated (scipy/optimize/_lsq/trf_linear.py:147-249 +
# common.py), float64, with a per-iteration trace; checked against the real
# scipy call in tests/test_oracle_golden.py.  Used to localise differences of
# the CUDA state machine (hb2_trf.cuh).
# ----------------------------------------------------------------------------
def trf_linear_restated(A, b, x_lsq, lb, ub, tol=1e-2, max_iter=200, lsmr_maxiter=1000, trace=None):
    from scipy.sparse.linalg import lsmr, LinearOperator

    EPS = np.finfo(float).eps
    A = A.tocsr()
    m, n = A.shape
    lbv, ubv = np.full(n, float(lb)), np.full(n, float(ub))

    def stb(x, s):
        steps = np.full(n, np.inf)
        nz = s != 0
        with np.errstate(over="ignore"):
            steps[nz] = np.maximum((lbv - x)[nz] / s[nz], (ubv - x)[nz] / s[nz])
        mn = steps.min()
        return mn, (steps == mn) * np.sign(s).astype(int)

    def minq(a, bq, lo, hi, c=0.0):
        t = [lo, hi]
        if a != 0:
            ext = -0.5 * bq / a
            if lo < ext < hi:
                t.append(ext)
        t = np.asarray(t)
        y = t * (a * t + bq) + c
        i = int(np.argmin(y))
        return t[i], y[i]

    d2 = ubv - lbv
    if np.all((x_lsq >= lbv) & (x_lsq <= ubv)):
        return x_lsq.copy(), 0, 3
    t = np.remainder(x_lsq - lbv, 2 * d2)
    x = lbv + np.minimum(t, 2 * d2 - t)
    ld, ud = x - lbv, ubv - x
    lth, uth = 0.1 * np.maximum(1, np.abs(lbv)), 0.1 * np.maximum(1, np.abs(ubv))
    la, ua = ld <= np.minimum(ud, lth), ud <= np.minimum(ld, uth)
    x = x.copy()
    x[la] = lbv[la] + lth[la]
    x[ua] = ubv[ua] - uth[ua]
    bad = (x < lbv) | (x > ubv)
    x[bad] = 0.5 * (lbv[bad] + ubv[bad])
    r = A.dot(x) - b
    g = A.T.dot(r)
    cost = 0.5 * np.dot(r, r)
    status = None
    it = -1
    for it in range(max_iter):
        v, dv = np.ones(n), np.zeros(n)
        mk = g < 0
        v[mk], dv[mk] = ubv[mk] - x[mk], -1
        mk = g > 0
        v[mk], dv[mk] = x[mk] - lbv[mk], 1
        g_norm = np.max(np.abs(g * v))
        if g_norm < tol:
            status = 1
        if status is not None:
            break
        diag_h = g * dv
        dr = diag_h**0.5
        d = v**0.5
        g_h = d * g
        eta = 1e-2 * min(0.5, g_norm)
        ltol = max(EPS, min(0.1, eta * g_norm))
        op = LinearOperator((m + n, n), matvec=lambda z: np.hstack((A.dot(np.ravel(z) * d), dr * np.ravel(z))),
                            rmatvec=lambda z: d * A.T.dot(z[:m]) + dr * z[m:], dtype=float)
        sol = lsmr(op, np.hstack((r, np.zeros(n))), maxiter=lsmr_maxiter, atol=ltol, btol=ltol)
        p_h = -sol[0]
        p = d * p_h
        p_dot_g = np.dot(p, g)
        if p_dot_g > 0:
            status = -1
        theta = 1 - min(0.005, g_norm)
        Ah = lambda z: A.dot(z * d)
        kind = "full"
        if np.all((x + p >= lbv) & (x + p <= ubv)):
            step = p
            pv = rv = agv = np.nan
        else:
            p_stride, hits = stb(x, p)
            r_h = p_h.copy()
            r_h[hits.astype(bool)] *= -1
            rr = d * r_h
            p = p * p_stride
            p_h = p_h * p_stride
            xb = x + p
            rsu, _ = stb(xb, rr)
            rsl = (1 - theta) * rsu
            rsu *= theta
            if rsu > 0:
                v1, u1 = Ah(r_h), Ah(p_h)
                a = 0.5 * (np.dot(v1, v1) + np.dot(r_h * diag_h, r_h))
                bq = np.dot(g_h, r_h) + np.dot(u1, v1) + np.dot(p_h * diag_h, r_h)
                c = 0.5 * np.dot(u1, u1) + np.dot(g_h, p_h) + 0.5 * np.dot(p_h * diag_h, p_h)
                rs, rv = minq(a, bq, rsl, rsu, c)
                r_h = p_h + r_h * rs
                rr = d * r_h
            else:
                rv = np.inf
            p_h = p_h * theta
            p = p * theta
            u1 = Ah(p_h)
            pv = 0.5 * (np.dot(u1, u1) + np.dot(p_h * diag_h, p_h)) + np.dot(p_h, g_h)
            ag_h = -g_h
            ag = d * ag_h
            agu, _ = stb(x, ag)
            agu *= theta
            v2 = Ah(ag_h)
            a = 0.5 * (np.dot(v2, v2) + np.dot(ag_h * diag_h, ag_h))
            bq = np.dot(g_h, ag_h)
            ags, agv = minq(a, bq, 0, agu)
            ag = ag * ags
            if pv < rv and pv < agv:
                step, kind = p, "p"
            elif rv < pv and rv < agv:
                step, kind = rr, "r"
            else:
                step, kind = ag, "ag"
        Js = A.dot(step)
        cost_change = -(0.5 * np.dot(Js, Js) + np.dot(step, g))
        if cost_change >= 0:
            xn = x + step
            xn[xn <= lbv] = np.nextafter(lbv, ubv)[xn <= lbv]
            xn[xn >= ubv] = np.nextafter(ubv, lbv)[xn >= ubv]
            x = xn
        if trace is not None:
            trace.append(dict(it=it, cost=cost, g_norm=g_norm, inner_itn=sol[2], kind=kind, p_value=pv, r_value=rv,
                              ag_value=agv, cost_change=cost_change, p_dot_g=p_dot_g, ltol=ltol))
        r = A.dot(x) - b
        g = A.T.dot(r)
        if cost_change < tol * cost:
            status = 2
        cost = 0.5 * np.dot(r, r)
    if status is None:
        status = 0
    return x, it + 1, status
