"""Drop-in for ``helicon.webApps.denovo3D.pipeline`` (the per-task wrapper the app's
grid search maps over, pipeline.py:85-497): ``process_one_task`` keeps the
reference's signature and return tuple; the solve + score runs through
``helicon_b200.solver_linear_regression.lsq_reconstruct`` and the post-solve display
products through ``helicon_b200.transforms`` -- both on the GPU.

Host side (once per task, not per voxel): image preparation and the integer geometry
(pipeline.py:180-349), restated line for line; the scikit-image calls of that preparation (rescale, closing, affine
warp, TV denoising) are restated with numpy / scipy.ndimage in ``helicon_b200.imageprep`` (parity unpinned: scikit-image
is not installed where this was built).  Still refused with ``NotImplementedError``: ``denoise`` "nl_mean" / "wavelet".
Tilted / refined tasks resample the display volume with ``transforms.transform_map``.
"""

from __future__ import annotations

import logging
import struct

import numpy as np

from . import transforms
from .solver_linear_regression import lsq_reconstruct

logger = logging.getLogger(__name__)

_MRC_MODES = {0: np.int8, 1: np.int16, 2: np.float32, 6: np.uint16, 12: np.float16}


def _read_mrc(path):
    """Minimal MRC2014 reader (little-endian, modes 0/1/2/6/12): the reference uses the ``mrcfile`` package
    (pipeline.py:37-43), which is not a dependency of this library."""
    with open(path, "rb") as f:
        hdr = f.read(1024)
        if len(hdr) < 1024:
            raise IOError(f"{path}: not an MRC file (short header)")
        nx, ny, nz, mode = struct.unpack("<4i", hdr[:16])
        mx = struct.unpack("<i", hdr[28:32])[0]
        cella_x = struct.unpack("<f", hdr[40:44])[0]
        nsymbt = struct.unpack("<i", hdr[92:96])[0]
        if mode not in _MRC_MODES or min(nx, ny, nz) <= 0:
            raise IOError(f"{path}: unsupported MRC mode {mode} or bad dimensions")
        f.seek(1024 + max(0, nsymbt))
        data = np.fromfile(f, dtype=np.dtype(_MRC_MODES[mode]).newbyteorder("<"), count=nx * ny * nz)
    data = data.reshape((nz, ny, nx)) if nz > 1 else data.reshape((ny, nx))
    apix = float(cella_x / mx) if mx > 0 else 0.0
    return data, apix


def get_images_from_file(imageFile):
    """pipeline.py:37-43: through ``mrcfile`` where it is installed (the reference's reader), else the MRC2014 reader of
    this module."""
    try:
        import mrcfile
    except ImportError:
        data, apix = _read_mrc(imageFile)
        return data, round(apix, 4)
    with mrcfile.open(imageFile) as mrc:
        apix = float(mrc.voxel_size.x)
        data = mrc.data
    return data, round(apix, 4)


def _image_reader():
    """``helicon.read_image_2d`` when this module is mounted inside the helicon package (the reference's
    tests patch that name, tests/test_denovo3D_pipeline.py:151), else the reader of this module."""
    import sys

    h = sys.modules.get("helicon")
    return getattr(h, "read_image_2d", None) or read_image_2d


def read_image_2d(imageFile, i):
    """lib/io_mrc.py:71-98: one 2-D slice of an MRC stack."""
    data, _ = _read_mrc(imageFile)
    return np.asarray(data[i] if data.ndim == 3 else data, dtype=np.float32)


def low_high_pass_filter(data, low_pass_fraction=0, high_pass_fraction=0):
    """lib/filters.py:314-372 (2-D case), host-side image preparation."""
    fft = np.fft.fft2(data)
    ny, nx = fft.shape
    Y, X = np.meshgrid(np.arange(ny, dtype=np.float32) - ny // 2, np.arange(nx, dtype=np.float32) - nx // 2, indexing="ij")
    Y /= ny // 2
    X /= nx // 2
    R2 = X**2 + Y**2
    if 0 < low_pass_fraction < 1:
        f2 = np.log(2) / (low_pass_fraction**2)
        fft *= np.fft.fftshift(np.exp(-f2 * R2))
    if 0 < high_pass_fraction < 1:
        f2 = np.log(2) / (high_pass_fraction**2)
        fft *= np.fft.fftshift(1.0 - np.exp(-f2 * R2))
    return np.real(np.fft.ifftn(fft))


def threshold_data(data, thresh_fraction=None, thresh_value=None):
    """lib/filters.py:283-311."""
    if thresh_fraction is not None and thresh_fraction >= 0:
        thresh = data.max() * thresh_fraction
    elif thresh_value is not None:
        thresh = thresh_value
    else:
        return data
    return np.clip(data, thresh, None) - thresh


def is_vertical(data):
    """webApps/denovo3D/utils.py:429-447."""
    return bool(np.max(np.sum(data, axis=0)) > np.max(np.sum(data, axis=1)))


def _unsupported(what):
    raise NotImplementedError(f"helicon_b200.pipeline: {what} is outside the CUDA hot path and not implemented")


def process_one_task(ti, ntasks, data, imageFile, imageIndex, twist, rise, rise_range, csym, tilt, tilt_range, psi,
                     psi_range, dy, dy_range, apix2d_orig, denoise, low_pass, transpose, horizontalize, target_apix3d,
                     target_apix2d, thresh_fraction, positive_constraint, tube_length, tube_diameter, tube_diameter_inner,
                     reconstruct_length, sym_oversample, interpolation, fsc_test, return_3d, score_metric, algorithm,
                     verbose, n_cpu=1):
    """pipeline.py:85-497, same arguments and return value
    ``(score, (x_proj, y_proj, z_sections, (rec3d, h1, h2) | None, D2, D3, L2, L3), (data_orig, imageFile, imageIndex,
    target_apix3d, target_apix2d, twist, rise, csym, tilt, psi, dy))`` or ``None`` for a blank image."""
    if data is None:
        data = _image_reader()(imageFile, imageIndex - 1)
    if not np.std(data):  # pipeline.py:183-187
        logger.warning(f"WARNING: the input image {imageFile}:{imageIndex} is a blank image")
        return None
    # ---- prepare_data (pipeline.py:146-178) ----------------------------------------------------------------------
    if low_pass > 2 * apix2d_orig:
        data = low_high_pass_filter(data, low_pass_fraction=2 * apix2d_orig / low_pass,
                                    high_pass_fraction=2.0 / np.max(data.shape))
    if denoise:
        if denoise == "tv":
            from .imageprep import denoise_tv_chambolle

            data = denoise_tv_chambolle(data)
        elif denoise in ("nl_mean", "wavelet"):
            _unsupported(f"denoise={denoise!r} (scikit-image non-local means / PyWavelets shrinkage; 'tv' is implemented)")
    if transpose > 0 or (transpose < 0 and is_vertical(data)):
        data = data.T
    if horizontalize:
        from .imageprep import auto_horizontalize

        data, theta_best, shift_best = auto_horizontalize(data, refine=True)
        logger.debug(f"Image {imageFile}-{imageIndex}: rotation={round(theta_best, 2)} deg shift={round(shift_best * apix2d_orig, 1)} A")
    data = np.ascontiguousarray(data, dtype=np.float32)
    ny, nx = data.shape
    ny_orig, nx_orig = ny, nx
    if tube_diameter < 0:  # pipeline.py:233-240
        from .imageprep import estimate_helix_rotation_center_diameter

        _, _, diameter = estimate_helix_rotation_center_diameter(data)
        tube_diameter = int(min(ny, diameter) * apix2d_orig * 2.5)
        logger.debug(f"Image {imageFile}-{imageIndex}: estimated tube diameter={tube_diameter} A")
    if tube_length < 0:  # pipeline.py:211-221
        if tube_diameter > ny * apix2d_orig / 2:
            tube_length = int(nx * apix2d_orig)
        else:
            tube_length = round(np.sqrt((nx * apix2d_orig) ** 2 / 4 - tube_diameter**2 / 4) * 2)
    # ---- geometry (pipeline.py:223-349) ---------------------------------------------------------------------------
    reconstruct_diameter = tube_diameter if 0 < tube_diameter < ny * apix2d_orig else ny * apix2d_orig
    reconstruct_diameter_inner = tube_diameter_inner if 0 < tube_diameter_inner < reconstruct_diameter else 0
    if reconstruct_length < rise:
        reconstruct_length = max(min(3 * np.max(rise_range), tube_length),
                                 round(np.tan(np.deg2rad(np.max(np.abs(tilt_range)))) * tube_diameter * 3))
    if target_apix2d < apix2d_orig:
        target_apix2d = apix2d_orig
    if target_apix2d != apix2d_orig:  # pipeline.py:268-275
        from .imageprep import down_scale

        data = np.ascontiguousarray(down_scale(data, target_apix=target_apix2d, apix_orig=apix2d_orig), dtype=np.float32)
    ny, nx = data.shape
    if thresh_fraction >= 0:  # pipeline.py:276-284 (modifies the caller's array in place, as the reference does)
        data_orig = data
        nr = min(ny // 2 - 1, int(np.ceil(reconstruct_diameter / 2 / target_apix2d) + 1))
        data -= np.median(data[(ny // 2 - nr, ny // 2 + nr), :])
        data = threshold_data(data, thresh_fraction=thresh_fraction)
        data /= np.max(data)
    else:
        data_orig = data
    if target_apix3d < 0:
        vol = reconstruct_length * (reconstruct_diameter**2 - reconstruct_diameter_inner**2) / 4 * np.pi
        target_apix3d = max(target_apix2d, round(np.power(vol / (nx * ny), 1 / 3) + 0.5))
    elif target_apix3d == 0:
        target_apix3d = target_apix2d
    D3 = int(round(reconstruct_diameter / target_apix3d))
    D3 += D3 % 2
    D3i = int(round(tube_diameter_inner / target_apix3d))
    D2 = int(round(reconstruct_diameter / target_apix2d))
    D2 += D2 % 2
    reconstruct_length_2d = tube_length if 0 < tube_length < nx * target_apix2d else nx * target_apix2d
    L2 = int(reconstruct_length_2d / target_apix2d)
    L2 += L2 % 2
    pitch = round(rise * 360 / abs(twist), 1)
    if reconstruct_length > 0:
        L3 = max(int(np.ceil(rise / target_apix3d)), int(np.ceil(reconstruct_length / target_apix3d)))
        L3 += L3 % 2
    else:
        L3 = int(L2 * target_apix2d / target_apix3d + 0.5)
        L3 += L3 % 2
    if sym_oversample <= 0:
        n_voxels = L3 * (D3**2 - D3i**2)
        ratio = 2**20 / n_voxels
        if ratio < 10:
            sym_oversample = max(1, int(round(ratio)))
        elif ratio < 100:
            sym_oversample = max(1, int(round(ratio / 10)) * 10)
        else:
            sym_oversample = max(1, int(round(ratio / 100)) * 100)
        if return_3d:
            sym_oversample *= 2
    # ---- solve + score (pipeline.py:351-404) on the GPU -------------------------------------------------------------
    refine_range = None
    if algorithm.get("model", "lsq") in ("lsq", "elasticnet", "lasso", "ridge"):
        r_dict = {}
        if tilt_range[1] > tilt_range[0]:
            r_dict["tilt"] = max(abs(tilt_range[0]), abs(tilt_range[1]))
        if psi_range > 0:
            r_dict["psi"] = psi_range
        if dy_range > 0:
            r_dict["dy"] = dy_range
        if r_dict:
            refine_range = r_dict
    (rec3d, rec3d_set_1, rec3d_set_2), score = lsq_reconstruct(
        projection_image=data, scale2d_to_3d=target_apix2d / target_apix3d, twist_degree=twist,
        rise_pixel=rise / target_apix3d, csym=csym, tilt_degree=tilt, psi_degree=psi, dy_pixel=dy / target_apix2d,
        thresh_fraction=thresh_fraction, positive_constraint=positive_constraint,
        reconstruct_diameter_3d_inner_pixel=D3i, reconstruct_diameter_2d_pixel=D2, reconstruct_diameter_3d_pixel=D3,
        reconstruct_length_2d_pixel=L2, reconstruct_length_3d_pixel=L3, sym_oversample=sym_oversample,
        interpolation=interpolation, fsc_test=fsc_test, score_metric=score_metric, target_apix2d=target_apix2d,
        verbose=verbose, algorithm=algorithm, refine_tilt_psi_dy_range=refine_range, cpu=n_cpu)
    # ---- display products (pipeline.py:405-458) on the GPU ----------------------------------------------------------
    twist_degree = twist if abs(twist) < 90 else 180 - abs(twist)
    if abs(twist_degree) > 1e-2:
        pitch_pixel = int(360 / abs(twist_degree) * rise / apix2d_orig + 0.5)
    else:
        pitch_pixel = int(np.ceil(2 * rise / apix2d_orig))
    new_length = max(nx_orig, int(pitch_pixel * 1.2))
    # refined tilt/psi/dy replace the given ones for the display (pipeline.py:428-436; consumed once)
    tilt_viz, psi_viz, dy_viz = tilt, psi, dy
    rp = getattr(lsq_reconstruct, "_refined_params", None)
    if rp:
        tilt_viz, psi_viz, dy_viz = rp.get("tilt", tilt), rp.get("psi", psi), rp.get("dy", dy)
        lsq_reconstruct._refined_params = {}
    if tilt_viz == 0 and psi_viz == 0 and dy_viz == 0:
        # transform_map is the identity: symmetrise and reduce on the device, the (pitch-long) volume never leaves it
        rec3d_x_proj, rec3d_y_proj, rec3d_z_sections = transforms.symmetrize_and_project(
            rec3d, target_apix3d, twist, rise, csym, (new_length, ny_orig, ny_orig), apix2d_orig, rise, apix2d_orig)
    else:
        rec3d_xform = transforms.apply_helical_symmetry(rec3d, target_apix3d, twist, rise, csym=csym,
                                                        new_size=(new_length, ny_orig, ny_orig), new_apix=apix2d_orig)
        rec3d_xform_2 = transforms.transform_map(rec3d_xform, scale=1.0, tilt=tilt_viz, psi=psi_viz, dy=dy_viz / apix2d_orig)
        rec3d_x_proj = np.sum(rec3d_xform_2, axis=2).T
        rec3d_y_proj = np.sum(rec3d_xform_2, axis=1).T
        y_max = rec3d_y_proj.max()
        if y_max > 0:
            rec3d_y_proj *= rec3d_x_proj.max() / y_max
        nz_per_rise = max(1, int(np.ceil(rise / apix2d_orig)))
        z0 = rec3d_xform.shape[0] // 2 - nz_per_rise // 2
        rec3d_z_sections = np.sum(rec3d_xform[z0:z0 + nz_per_rise, :, :], axis=0)
        vmin, vmax = rec3d_z_sections.min(), rec3d_z_sections.max()
        if vmax > vmin:
            t0, t1 = rec3d_x_proj.min(), rec3d_x_proj.max()
            rec3d_z_sections = (rec3d_z_sections - vmin) * (t1 - t0) / (vmax - vmin) + t0
    nz, ny, nx = rec3d.shape
    logger.info(f"Task {ti+1}/{ntasks}: {imageFile}-{imageIndex}:\\tpitch={round(pitch, 1)}A/twist={round(twist, 3)} "
                f"rise={round(rise, 3)}A csym={csym} => reconstruction size={nx}x{ny}x{nz}voxels "
                f"voxelsize={round(target_apix3d, 3)}A\\t=>\\tscore={round(float(score), 6)}")
    return_data = (rec3d_x_proj, rec3d_y_proj, rec3d_z_sections,
                   (rec3d, rec3d_set_1, rec3d_set_2) if return_3d else None, D2, D3, L2, L3)
    return (score, return_data,
            (data_orig, imageFile, imageIndex, target_apix3d, target_apix2d, twist, rise, csym, tilt, psi, dy))
