"""Drop-in for the reference's ``dup.refine`` (src/dup/refine.py:19-120) with SSIM on the GPU.

``refine_pair`` keeps the reference's contract: ``None`` when either image is unreadable
(:80-83); each metric in its own try/except (:89-98); ``is_duplicate = ssim >= thr or orb >= thr``
and the ``reason`` strings ``"ssim>=0.9, orb>=0.15"`` / ``"ssim unavailable, orb unavailable"`` /
``"below thresholds"`` (:100-108).  ``refine_pairs_batch`` is the batched sibling: host threads
decode and prepare the 'L' planes exactly like ``_compute_ssim`` (:44-51, Pillow
``ImageOps.fit(..., BICUBIC)``), all SSIMs of one plane geometry are computed by a single
``ke_ssim_batch`` launch, and the ORB cross-check (:55-68, OpenCV, stays on the host by scope) runs
in the same host threads.
"""
from __future__ import annotations

import logging
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from pathlib import Path
from typing import Sequence

import numpy as np
from PIL import Image, ImageOps

logger = logging.getLogger(__name__)


@dataclass(frozen=True)
class RefinementThresholds:
    ssim: float = 0.9
    orb: float = 0.15


@dataclass(frozen=True)
class RefinedMatch:
    file_id_a: int
    file_id_b: int
    ssim: float | None
    orb_ratio: float | None
    is_duplicate: bool
    reason: str


def _load(path):
    """The reference's own loader (utils.image_io.safe_load_image: EXIF transpose, size caps,
    alpha -> white RGB) when the reference tree is importable; a plain Pillow RGB load otherwise."""
    try:
        from utils.image_io import safe_load_image  # reference host code

        return safe_load_image(path)
    except ModuleNotFoundError:
        try:
            with Image.open(Path(path)) as im:
                return ImageOps.exif_transpose(im).convert("RGB")
        except Exception:
            return None


def _prepare_image(image: Image.Image, size: tuple[int, int]) -> np.ndarray:
    fitted = ImageOps.fit(image.convert("RGB"), size, Image.Resampling.BICUBIC)
    return np.asarray(fitted, dtype=np.float32) / 255.0


def _ssim_planes(img_a: Image.Image, img_b: Image.Image) -> tuple[np.ndarray, np.ndarray]:
    """The two uint8 'L' planes SSIM is evaluated on (reference :45-49)."""
    size = (min(img_a.width, img_b.width), min(img_a.height, img_b.height))
    if size[0] == 0 or size[1] == 0:
        size = (max(img_a.width, img_b.width), max(img_a.height, img_b.height))
    pa = ImageOps.fit(img_a.convert("L"), size, Image.Resampling.BICUBIC)
    pb = ImageOps.fit(img_b.convert("L"), size, Image.Resampling.BICUBIC)
    return np.asarray(pa, dtype=np.uint8), np.asarray(pb, dtype=np.uint8)


def _compute_ssim(img_a: Image.Image, img_b: Image.Image) -> float:
    """structural_similarity(a/255, b/255, data_range=1.0) on the GPU; raises ``ValueError`` like
    scikit-image when a side is smaller than the 7x7 window."""
    from .. import ops

    pa, pb = _ssim_planes(img_a, img_b)
    return float(ops.ssim_pairs(pa, pb)[0])


def _orb_descriptors(img_a: Image.Image, img_b: Image.Image):
    """ORB detect + describe of both images (reference :57-61; OpenCV on the host): ``(len(kpa), da, len(kpb), db)``
    with ``None`` descriptors where the reference returns 0.0 early."""
    import cv2

    ga, gb = np.asarray(img_a.convert("L")), np.asarray(img_b.convert("L"))
    orb = cv2.ORB_create()
    kpa, da = orb.detectAndCompute(ga, None)
    kpb, db = orb.detectAndCompute(gb, None)
    if da is None or db is None or not kpa or not kpb:
        return 0, None, 0, None
    return len(kpa), da, len(kpb), db


def _orb_ratio_from_matches(n_matches: int, n_kpa: int, n_kpb: int) -> float:
    return float(n_matches / min(n_kpa, n_kpb)) if n_matches and n_kpa and n_kpb else 0.0


def _compute_orb_ratio(img_a: Image.Image, img_b: Image.Image) -> float:
    """ORB + brute-force Hamming cross-check ratio (reference :55-68): detector / descriptor by OpenCV on the host, the
    cross-checked brute-force match (``cv2.BFMatcher(NORM_HAMMING, crossCheck=True)``) on the GPU."""
    from .. import ops

    n_kpa, da, n_kpb, db = _orb_descriptors(img_a, img_b)
    if da is None:
        return 0.0
    return _orb_ratio_from_matches(int(ops.orb_match_pairs([da], [db])[0]), n_kpa, n_kpb)


def _decide(file_id_a, file_id_b, ssim_value, orb_ratio, errors, cfg: RefinementThresholds) -> RefinedMatch:
    hits = []
    if ssim_value is not None and ssim_value >= cfg.ssim:
        hits.append(f"ssim>={cfg.ssim}")
    if orb_ratio is not None and orb_ratio >= cfg.orb:
        hits.append(f"orb>={cfg.orb}")
    reason = ", ".join(hits or errors) if hits or errors else "below thresholds"
    return RefinedMatch(file_id_a=file_id_a, file_id_b=file_id_b, ssim=ssim_value, orb_ratio=orb_ratio,
                        is_duplicate=bool(hits), reason=reason)


def refine_pair(file_id_a: int, file_id_b: int, path_a: str | Path, path_b: str | Path, *,
                thresholds: RefinementThresholds | None = None) -> RefinedMatch | None:
    """Compare two images with SSIM (GPU) and ORB (host)."""
    image_a, image_b = _load(path_a), _load(path_b)
    if image_a is None or image_b is None:
        return None
    cfg = thresholds or RefinementThresholds()
    ssim_value = orb_ratio = None
    errors: list[str] = []
    try:
        ssim_value = _compute_ssim(image_a, image_b)
    except Exception as exc:
        logger.warning("SSIM refinement failed for %s and %s: %s", path_a, path_b, exc)
        errors.append("ssim unavailable")
    try:
        orb_ratio = _compute_orb_ratio(image_a, image_b)
    except Exception as exc:
        logger.warning("ORB refinement failed for %s and %s: %s", path_a, path_b, exc)
        errors.append("orb unavailable")
    return _decide(file_id_a, file_id_b, ssim_value, orb_ratio, errors, cfg)


def refine_pairs_batch(pairs: Sequence[tuple[int, int, str | Path, str | Path]], *,
                       thresholds: RefinementThresholds | None = None, max_workers: int | None = None,
                       use_orb: bool = True) -> list[RefinedMatch | None]:
    """[(file_id_a, file_id_b, path_a, path_b)] -> [RefinedMatch | None], same per-pair semantics
    as ``refine_pair`` with every SSIM of one geometry in a single GPU launch."""
    cfg = thresholds or RefinementThresholds()
    # bounded rounds: the planes of at most KE_REFINE_CHUNK pairs (default 256) are resident between SSIM launches
    # (the reference holds one pair at a time; multi-megapixel planes of ALL pairs would not fit host memory)
    chunk = max(1, int(os.environ.get("KE_REFINE_CHUNK", "256")))
    out: list[RefinedMatch | None] = []
    for lo in range(0, len(pairs), chunk):
        out += _refine_chunk(pairs[lo:lo + chunk], cfg, max_workers, use_orb)
    return out


def _refine_chunk(pairs, cfg: RefinementThresholds, max_workers, use_orb) -> list[RefinedMatch | None]:
    from .. import ops

    workers = max_workers or max(1, (os.cpu_count() or 4) - 1)

    def prepare(pair):
        ida, idb, pa, pb = pair
        ia, ib = _load(pa), _load(pb)
        if ia is None or ib is None:
            return None
        rec = {"planes": None, "orb": None, "errors": []}
        try:
            rec["planes"] = _ssim_planes(ia, ib)
        except Exception as exc:
            logger.warning("SSIM refinement failed for %s and %s: %s", pa, pb, exc)
        if use_orb:
            try:
                rec["orb_desc"] = _orb_descriptors(ia, ib)  # matched below, all pairs of the round in one launch
            except Exception as exc:
                logger.warning("ORB refinement failed for %s and %s: %s", pa, pb, exc)
                rec["orb_failed"] = True
        return rec

    with ThreadPoolExecutor(max_workers=workers) as pool:
        prepared = list(pool.map(prepare, pairs))

    if use_orb:
        todo = [k for k, rec in enumerate(prepared) if rec is not None and "orb_desc" in rec]
        try:
            counts = ops.orb_match_pairs([prepared[k]["orb_desc"][1] for k in todo], [prepared[k]["orb_desc"][3] for k in todo])
            for k, m in zip(todo, counts.tolist()):
                n_kpa, _, n_kpb, _ = prepared[k]["orb_desc"]
                prepared[k]["orb"] = _orb_ratio_from_matches(int(m), n_kpa, n_kpb)
        except Exception as exc:
            logger.warning("ORB refinement failed for %d pair(s): %s", len(todo), exc)
            for k in todo:
                prepared[k]["orb_failed"] = True

    ssim_values: list[float | None] = [None] * len(pairs)
    groups: dict[tuple[int, int], list[int]] = {}
    for k, rec in enumerate(prepared):
        if rec is not None and rec["planes"] is not None:
            groups.setdefault(rec["planes"][0].shape, []).append(k)
    for shape, members in groups.items():
        try:
            vals = ops.ssim_pairs(np.stack([prepared[k]["planes"][0] for k in members]),
                                  np.stack([prepared[k]["planes"][1] for k in members]))
        except Exception as exc:  # e.g. a side < 7: skimage raises, the reference records "ssim unavailable" (:91-94)
            logger.warning("SSIM refinement failed for %d pair(s) of shape %s: %s", len(members), shape, exc)
            continue
        for k, v in zip(members, vals.tolist()):
            ssim_values[k] = float(v)

    out: list[RefinedMatch | None] = []
    for k, (pair, rec) in enumerate(zip(pairs, prepared)):
        if rec is None:
            out.append(None)
            continue
        errors = []
        if ssim_values[k] is None:
            errors.append("ssim unavailable")
        if use_orb and rec.get("orb_failed"):
            errors.append("orb unavailable")
        out.append(_decide(pair[0], pair[1], ssim_values[k], rec["orb"], errors, cfg))
    return out


__all__ = ["RefinementThresholds", "RefinedMatch", "refine_pair", "refine_pairs_batch"]
