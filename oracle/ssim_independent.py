"""A SECOND, independent SSIM oracle — TEST INFRASTRUCTURE ONLY (imported by tests/ and nothing else).

Why it exists: scikit-image is absent from this image and cannot be installed, so ``oracle/ref_py.py`` restates
``skimage.metrics.structural_similarity`` (0.25.2, as called at /root/reference/src/dup/refine.py:52) on
``scipy.ndimage.uniform_filter``.  A restatement checked only against itself would let a shared misreading of
skimage (window, ``cov_norm``, crop, constants) through.  This module shares NO code with ``ref_py`` and does not
use ``uniform_filter``: it evaluates the published definition

    S(x) = (2 mu_a mu_b + C1)(2 cov_ab + C2) / ((mu_a^2 + mu_b^2 + C1)(var_a + var_b + C2))

    mu = mean over the 7x7 window centred on x, var/cov = SAMPLE (n-1 = 48) variance / covariance of the window,
    C1 = (0.01 L)^2, C2 = (0.03 L)^2 with L = data_range = 1.0 on images scaled to [0, 1],
    MSSIM = mean of S over the pixels whose window lies inside the image (3-pixel border dropped)

(Wang, Bovik, Sheikh, Simoncelli 2004, with skimage's defaults ``use_sample_covariance=True``, ``win_size=7``,
``gaussian_weights=False``) in float64, tap by tap, plus three closed forms that need no filter at all.
Agreement between this module, ``ref_py`` and the CUDA kernel within 1e-5 is asserted by
``tests/test_oracle_pinned.py`` (CPU) and ``tests/test_gpu_parity.py`` (GPU).
"""
from __future__ import annotations

import numpy as np

WIN = 7
NPIX = WIN * WIN


def mssim_bruteforce(a_u8: np.ndarray, b_u8: np.ndarray) -> float:
    """49 explicit taps per window, float64 throughout, two-pass (mean first, then centred moments)."""
    a = np.asarray(a_u8, dtype=np.float64) / 255.0
    b = np.asarray(b_u8, dtype=np.float64) / 255.0
    if a.shape != b.shape or a.ndim != 2:
        raise ValueError("two 2-D planes of equal shape")
    h, w = a.shape
    if h < WIN or w < WIN:
        raise ValueError("plane smaller than the window")
    oh, ow = h - WIN + 1, w - WIN + 1
    taps_a = [a[dy:dy + oh, dx:dx + ow] for dy in range(WIN) for dx in range(WIN)]
    taps_b = [b[dy:dy + oh, dx:dx + ow] for dy in range(WIN) for dx in range(WIN)]
    mu_a = sum(taps_a) / NPIX
    mu_b = sum(taps_b) / NPIX
    var_a = sum((t - mu_a) ** 2 for t in taps_a) / (NPIX - 1)
    var_b = sum((t - mu_b) ** 2 for t in taps_b) / (NPIX - 1)
    cov = sum((ta - mu_a) * (tb - mu_b) for ta, tb in zip(taps_a, taps_b)) / (NPIX - 1)
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    s = ((2.0 * mu_a * mu_b + c1) * (2.0 * cov + c2)) / ((mu_a ** 2 + mu_b ** 2 + c1) * (var_a + var_b + c2))
    return float(s.mean())


def closed_form_constants(level_a: int, level_b: int) -> float:
    """Two constant planes: every variance and covariance is 0, so S = (2ab + C1) / (a^2 + b^2 + C1) everywhere."""
    a, b = level_a / 255.0, level_b / 255.0
    c1 = 0.01 ** 2
    return (2.0 * a * b + c1) / (a * a + b * b + c1)


def checkerboard(h: int, w: int, p: int, q: int) -> np.ndarray:
    """1-pixel checkerboard: level p where (y + x) is even, q where it is odd."""
    yy, xx = np.mgrid[0:h, 0:w]
    return np.where((yy + xx) % 2 == 0, p, q).astype(np.uint8)


def closed_form_inverted_checkerboard(h: int, w: int, p: int, q: int) -> float:
    """A = checkerboard(p, q), B = 255 - A.  A 7x7 window holds 25 pixels of the level under its centre and 24 of the
    other, so per window (levels x = centre level, y = other, on the 0..1 scale):
        mu_a = (25x + 24y)/49,  mu_b = 1 - mu_a,  var_a = var_b = (x - y)^2 * 25/98,  cov = -var_a
    and MSSIM is the average of the two window kinds weighted by how many centres of each parity the interior has."""
    c1, c2 = 0.01 ** 2, 0.03 ** 2

    def s_of(x: float, y: float) -> float:
        mu_a = (25.0 * x + 24.0 * y) / 49.0
        mu_b = 1.0 - mu_a
        var = (x - y) ** 2 * 25.0 / 98.0
        return ((2.0 * mu_a * mu_b + c1) * (-2.0 * var + c2)) / ((mu_a ** 2 + mu_b ** 2 + c1) * (2.0 * var + c2))

    yy, xx = np.mgrid[3:h - 3, 3:w - 3]
    even = int(np.count_nonzero((yy + xx) % 2 == 0))
    odd = yy.size - even
    return (even * s_of(p / 255.0, q / 255.0) + odd * s_of(q / 255.0, p / 255.0)) / (even + odd)
