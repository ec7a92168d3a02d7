"""Drop-in for the two ``helicon`` functions ``pipeline.process_one_task`` calls after
the solve, on the GPU:

* ``apply_helical_symmetry`` -- ``helicon.apply_helical_symmetry``
  (lib/transforms.py:58-165), same arguments and result (bit-identical);
* ``symmetrize_and_project`` -- the same volume reduced on the device to the three
  sums the task keeps (pipeline.py:435-447), so that the (pitch-long) volume never
  crosses PCIe unless it is asked for.

The host only plans what the reference decides per output SLICE with Python/numpy
scalars (z window of the unit, the helical copies reaching each slice and their
interpolation weights, the 2x2 matrices); every per-voxel operation is a CUDA kernel
(``csrc/hb2_symm.cuh``).  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _plan(data, apix, twist_degree, rise_angstrom, csym, fraction, new_size, new_apix):
    """lib/transforms.py:69-112 without the (j, i) loops."""
    nz0, ny0, nx0 = data.shape
    if new_apix is None:
        new_apix = apix
    if new_size is None:
        new_size = data.shape
    new_size = tuple(int(v) for v in new_size)
    if new_size != tuple(data.shape):
        nz, ny, nx = max(nz0, new_size[0]), max(ny0, new_size[1]), max(nx0, new_size[2])
    else:
        nz, ny, nx = nz0, ny0, nx0
    hsym_max = max(1, int(nz * new_apix / rise_angstrom))
    hs = np.arange(-hsym_max, hsym_max + 1)
    profile_z = np.sum(np.sum(data, axis=-1), axis=-1)
    threshold = 0.01 * np.max(profile_z)
    nzi = np.where(profile_z > threshold)[0]
    if len(nzi) == 0:
        raise ValueError("apply_helical_symmetry: the input volume has no slice above 1 % of the maximum")
    z0, z1 = int(nzi[0]), int(nzi[-1])
    zmid = (z0 + z1) // 2 + (z0 + z1) % 2
    z0 = max(z0, zmid - int(nz0 * fraction + 0.5) // 2)
    z1 = min(z1, zmid + int(nz0 * fraction + 0.5) // 2)
    # output slices only (the reference crops the working grid afterwards, lib/transforms.py:157-163)
    if (nz, ny, nx) != new_size:
        oz, oy, ox = nz // 2 - new_size[0] // 2, ny // 2 - new_size[1] // 2, nx // 2 - new_size[2] // 2
        out = (2 * (new_size[0] // 2), 2 * (new_size[1] // 2), 2 * (new_size[2] // 2))
    else:
        oz = oy = ox = 0
        out = new_size
    kk = np.arange(oz, oz + out[0])
    K2 = ((kk[None, :] - nz // 2) * new_apix + hs[:, None] * rise_angstrom) / apix + nz0 // 2  # [h, k]
    ok = ~((K2 < z0) | (K2 >= z1))
    kq, hh = np.nonzero(ok.T)  # k-major, h ascending inside a slice (the reference's accumulation order per voxel)
    k2 = K2[hh, kq]
    fl = np.floor(k2)
    k_begin = np.zeros(out[0] + 1, dtype=np.int64)
    np.cumsum(np.bincount(kq, minlength=out[0]), out=k_begin[1:])
    rot = np.deg2rad(twist_degree * hs[:, None] + 360 * np.arange(csym)[None, :] / csym)
    mats = np.ascontiguousarray(np.stack([np.cos(rot), np.sin(rot), -np.sin(rot), np.cos(rot)], axis=-1).reshape(-1, 4))
    return dict(work=(nz, ny, nx), off=(oz, oy, ox), out=out, new_apix=float(new_apix), k_begin=k_begin,
                ent_h=np.ascontiguousarray(hh, dtype=np.int32), ent_floor=np.ascontiguousarray(fl, dtype=np.int32),
                ent_ceil=np.ascontiguousarray(np.ceil(k2), dtype=np.int32),
                ent_wk=np.ascontiguousarray(k2 - fl, dtype=np.float64), mats=mats)


def _run(data, apix, twist_degree, rise_angstrom, csym, fraction, new_size, new_apix, want_volume, nz_slab, device, stream):
    lib = _lib.require_gpu()
    data = np.ascontiguousarray(data, dtype=np.float32)
    if data.ndim != 3:
        raise ValueError("apply_helical_symmetry expects a 3-D volume")
    P = _plan(data, float(apix), float(twist_degree), float(rise_angstrom), int(csym), float(fraction), new_size, new_apix)
    nz1, ny1, nx1 = P["out"]
    zs0 = zs1 = 0
    if nz_slab:  # pipeline.py:443-446: nz_per_rise slices around the centre of the symmetrised volume
        zs0 = nz1 // 2 - int(nz_slab) // 2
        zs1 = zs0 + int(nz_slab)
        zs0, zs1 = max(0, zs0), min(nz1, zs1)
    prm = _lib.SymmParams(data.shape[0], data.shape[1], data.shape[2], *P["work"], *P["off"], nz1, ny1, nx1, float(apix),
                          P["new_apix"], int(csym), len(P["ent_h"]), zs0, max(zs0, zs1))
    vol = np.empty((nz1, ny1, nx1), dtype=np.float32) if want_volume else None
    xs = np.empty((nz1, ny1), dtype=np.float32)
    ys = np.empty((nz1, nx1), dtype=np.float32)
    zs = np.empty((ny1, nx1), dtype=np.float32)
    sh = None if stream is None else C.c_void_p(int(getattr(stream, "cuda_stream", stream)))
    _lib.check(lib.hb2_helical_symmetrize(
        _lib.ptr(data), C.byref(prm), _lib.ptr(P["k_begin"]), _lib.ptr(P["ent_h"]), _lib.ptr(P["ent_floor"]),
        _lib.ptr(P["ent_ceil"]), _lib.ptr(P["ent_wk"]), _lib.ptr(P["mats"]), len(P["mats"]),
        _lib.ptr(vol) if want_volume else None, _lib.ptr(xs), _lib.ptr(ys), _lib.ptr(zs), int(device), sh))
    return vol, xs, ys, zs


def apply_helical_symmetry(data, apix, twist_degree, rise_angstrom, csym=1, fraction=1.0, new_size=None, new_apix=None,
                           cpu=1, device=0, stream=None):
    """``helicon.apply_helical_symmetry`` (lib/transforms.py:58-165); ``cpu`` is accepted and ignored."""
    vol, _, _, _ = _run(data, apix, twist_degree, rise_angstrom, csym, fraction, new_size, new_apix, True, None, device, stream)
    return vol


def symmetrize_and_project(data, apix, twist_degree, rise_angstrom, csym, new_size, new_apix, rise, apix2d_orig,
                           return_volume=False, device=0, stream=None):
    """The display products of pipeline.py:405-458 for tilt = psi = dy = 0 (``transform_map`` is the identity):
    returns ``(rec3d_x_proj, rec3d_y_proj, rec3d_z_sections[, rec3d_xform])``."""
    nz_per_rise = max(1, int(np.ceil(rise / apix2d_orig)))
    vol, xs, ys, zs = _run(data, apix, twist_degree, rise_angstrom, csym, 1.0, new_size, new_apix, return_volume,
                           nz_per_rise, device, stream)
    x_proj = xs.T
    y_proj = ys.T.copy()
    y_max = y_proj.max()
    if y_max > 0:
        y_proj *= x_proj.max() / y_max
    z_sec = zs
    vmin, vmax = z_sec.min(), z_sec.max()
    if vmax > vmin:
        tmin, tmax = x_proj.min(), x_proj.max()
        z_sec = (z_sec - vmin) * (tmax - tmin) / (vmax - vmin) + tmin
    return (x_proj, y_proj, z_sec, vol) if return_volume else (x_proj, y_proj, z_sec)


def pad_to_size(data, shape):
    """lib/transforms.py:441-479: zero-pad a 2-D / 3-D array, centred, to ``shape`` (no cropping)."""
    assert data.ndim in [2, 3]
    if data.shape == tuple(shape):
        return data
    pads = []
    for n, m in zip(data.shape, shape):
        before = max(0, (m - n) // 2)
        pads.append((before, max(0, m - before - n)))
    return np.pad(data, pad_width=tuple(pads), mode="constant")


def transform_map(data, scale=1.0, rot=0, tilt=0, psi=0, dx=0, dy=0, dz=0):
    """lib/transforms.py:168-235: resample a volume under a ZYZ rotation + shift with cubic splines.

    Host side, like the reference: the arithmetic is scipy's (``Rotation.apply`` + ``ndimage.map_coordinates(order=3)``,
    third-party code the reference calls as well), applied to the display volume of a tilted / refined task only
    (pipeline.py:436-438; the identity for the untilted grid search, where nothing is computed).  Pinned to outputs of
    the reference in tests/golden/transform_map.npz."""
    if scale == 1 and rot == 0 and tilt == 0 and psi == 0 and dx == 0 and dy == 0 and dz == 0:
        return data
    from scipy.ndimage import map_coordinates
    from scipy.spatial.transform import Rotation as R

    nz, ny, nx = data.shape
    k = np.arange(0, nz, dtype=np.int32) - nz // 2
    j = np.arange(0, ny, dtype=np.int32) - ny // 2
    i = np.arange(0, nx, dtype=np.int32) - nx // 2
    Z, Y, X = np.meshgrid(k, j, i, indexing="ij")
    if scale != 1.0:
        Z, Y, X = Z * scale, Y * scale, X * scale
    xyz = R.from_euler("ZYZ", (rot, tilt, psi), degrees=True).apply(
        np.vstack((X.ravel(), Y.ravel(), Z.ravel())).transpose(), inverse=False)
    xyz[:, 0] += nx // 2 - dx
    xyz[:, 1] += ny // 2 - dy
    xyz[:, 2] += nz // 2 - dz
    return map_coordinates(data, xyz[:, [2, 1, 0]].T, order=3).reshape((nz, ny, nx))
