"""Host-side image preparation and 2-D score metrics of the denovo3D task wrapper.

These run once per TASK on one small 2-D image (not per voxel or per iteration) and are outside the CUDA hot path; the
reference implements them with scikit-image, which exists neither in the build container nor on the GPU box, so the
functions below RESTATE the published algorithms of the scikit-image 0.2x calls the reference makes with numpy /
scipy.ndimage.  **Parity unpinned**: the reference's own calls cannot be executed here (ImportError), so no golden
output exists; what is tested are the algorithms' defining properties (tests/test_host_cpu.py).

* ``down_scale``                -- lib/filters.py:375-412  (skimage.transform.rescale(order=3, anti_aliasing=True))
* ``ssim_score``                -- lib/analysis.py:487-514 (skimage.metrics.structural_similarity defaults)
* ``ms_ssim_score``             -- lib/analysis.py:517-582
* ``mutual_information_score``  -- lib/analysis.py:585-613 (skimage.metrics.normalized_mutual_information(bins=64) - 1)
"""

from __future__ import annotations

import numpy as np

from .transforms import pad_to_size


def _rescale(image, scale, order, anti_aliasing=True):
    """skimage.transform.rescale for a 2-D float image, mode="reflect", clip=True, preserve_range=False:
    output shape = round(scale * shape); Gaussian pre-filter sigma = max(0, (in/out - 1) / 2) per axis
    (resize(): anti_aliasing_sigma); ``scipy.ndimage.zoom(order, mode="reflect", grid_mode=True)`` (resize() since
    skimage 0.19); output clipped to the input's value range."""
    from scipy import ndimage as ndi

    image = np.asarray(image, dtype=np.float64)
    out_shape = tuple(int(v) for v in np.round(np.asarray(image.shape) * scale))
    factors = np.asarray(image.shape, dtype=np.float64) / np.asarray(out_shape, dtype=np.float64)
    filtered = image
    if anti_aliasing:
        sigma = np.maximum(0, (factors - 1) / 2)
        if np.any(sigma > 0):
            filtered = ndi.gaussian_filter(image, sigma, cval=0, mode="reflect")
    out = ndi.zoom(filtered, 1 / factors, order=order, mode="reflect", cval=0, grid_mode=True)
    if out.shape != out_shape:  # zoom rounds the same way; guard against a one-pixel disagreement
        out = pad_to_size(out[: out_shape[0], : out_shape[1]], out_shape)
    return np.clip(out, image.min(), image.max())


def down_scale(data, target_apix, apix_orig):
    """lib/filters.py:375-412: cubic, anti-aliased down-scaling to a larger pixel size; even output size (zero-padded)."""
    if target_apix == apix_orig or target_apix < apix_orig:
        return data
    out = _rescale(data, apix_orig / target_apix, order=3, anti_aliasing=True)
    ny, nx = out.shape
    return pad_to_size(out, shape=(ny + ny % 2, nx + nx % 2))


def _ssim(im1, im2, data_range, win_size=7, K1=0.01, K2=0.03):
    """skimage.metrics.structural_similarity with its defaults: 7 x 7 uniform window, sample covariance, mean of the
    SSIM map cropped by (win_size - 1) // 2 (Wang et al. 2004)."""
    from scipy.ndimage import uniform_filter

    im1 = np.asarray(im1, dtype=np.float64)
    im2 = np.asarray(im2, dtype=np.float64)
    if min(im1.shape) < win_size:
        raise ValueError("win_size exceeds image extent")
    NP = win_size**im1.ndim
    cov_norm = NP / (NP - 1)
    f = dict(size=win_size, mode="reflect")
    ux, uy = uniform_filter(im1, **f), uniform_filter(im2, **f)
    uxx, uyy, uxy = uniform_filter(im1 * im1, **f), uniform_filter(im2 * im2, **f), uniform_filter(im1 * im2, **f)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux**2 + uy**2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean(dtype=np.float64))


def ssim_score(img1, img2):
    """lib/analysis.py:487-514."""
    if img1.shape != img2.shape:
        raise ValueError(f"Image shapes must match: {img1.shape} vs {img2.shape}")
    try:
        data_range = max(img1.max() - img1.min(), img2.max() - img2.min())
        if data_range == 0:
            return 0.0
        return float(_ssim(img1, img2, data_range))
    except Exception:
        return 0.0


def ms_ssim_score(img1, img2):
    """lib/analysis.py:517-582: SSIM at up to five scales (each a rescale(0.5, anti_aliasing=True), bilinear), combined
    as a weighted geometric mean with the Wang et al. weights."""
    if img1.shape != img2.shape:
        raise ValueError(f"Image shapes must match: {img1.shape} vs {img2.shape}")
    try:
        data_range = max(img1.max() - img1.min(), img2.max() - img2.min())
        if data_range == 0:
            return 0.0
        all_weights = np.array([0.0448, 0.2856, 0.3001, 0.2363, 0.1333])
        vals = []
        for i in range(len(all_weights)):
            h, w = img1.shape
            if h < 8 or w < 8:
                break
            vals.append(max(_ssim(img1, img2, data_range), 0.0))
            if i < len(all_weights) - 1:
                img1 = _rescale(img1, 0.5, order=1)
                img2 = _rescale(img2, 0.5, order=1)
                data_range = max(img1.max() - img1.min(), img2.max() - img2.min())
                if data_range == 0:
                    break
        if not vals:
            return 0.0
        wts = all_weights[: len(vals)]
        wts = wts / wts.sum()
        out = 1.0
        for s, w in zip(vals, wts):
            out *= s**w
        return float(out)
    except Exception:
        return 0.0


def mutual_information_score(img1, img2):
    """lib/analysis.py:585-613: (H(X) + H(Y)) / H(X, Y) - 1 from a 64 x 64 joint histogram (Studholme et al. 1999)."""
    if img1.shape != img2.shape:
        raise ValueError(f"Image shapes must match: {img1.shape} vs {img2.shape}")
    try:
        hist, _ = np.histogramdd([np.reshape(img1, -1), np.reshape(img2, -1)], bins=64, density=True)

        def entropy(p):
            p = np.asarray(p, dtype=np.float64).ravel()
            p = p / p.sum()
            p = p[p > 0]
            return float(-(p * np.log(p)).sum())

        nmi = (entropy(hist.sum(axis=0)) + entropy(hist.sum(axis=1))) / entropy(hist)
        return float(nmi - 1.0)
    except Exception:
        return 0.0
