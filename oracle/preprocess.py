"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of StreamMind's frame preprocessing.

Only tests/ may import this module, as the checker; the product path (streammind_b200/*) never imports it.

Path restated (paths relative to /root/reference):
  streammind/mm_utils.py:446-464   process_video, aspect_ratio == 'pad': Image.fromarray(frame) -> expand2square(image,
                                   tuple(int(x*255) for x in processor.image_mean)) -> processor.preprocess(...)['pixel_values']
  streammind/mm_utils.py:257-268   expand2square: paste the frame centred into a square of the background colour
  video_score_stream_demo.py:285-287  the streaming demo's per-frame call of the above; the result is .half()'ed for the tower

The arithmetic lives in two third-party dependencies that are NOT vendored under /root/reference:
  * transformers==4.44.2 (requirements.txt:355; pyproject pins 4.40.0) CLIPImageProcessor.preprocess of
    openai/clip-vit-large-patch14-336: convert_rgb -> resize(shortest_edge=336, resample=BICUBIC) through PIL
    (image_transforms.resize -> PIL.Image.resize, reducing_gap=None) -> center_crop(336) -> rescale: uint8 * (1/255) in
    float64, cast to float32 -> normalize: (x - mean) / std in float32 -> channels first.
    (transformers 5.5 in this image resolves CLIPImageProcessor to a torchvision backend with different resampling
    arithmetic; it is NOT the reference's path and is not used here.)
  * Pillow==9.4.0 (requirements.txt:218) src/libImaging/Resample.c, 8 bits per channel: precompute_coeffs (bicubic, a = -0.5,
    support 2 * max(scale, 1), coefficients normalised in double), normalize_coeffs_8bpc (fixed point, 22 bits, round half
    away from zero), ImagingResampleHorizontal_8bpc then ImagingResampleVertical_8bpc, each accumulating from 1 << 21 in
    int32 and clipping (ss >> 22) to [0, 255] -- the intermediate image is uint8.

Parity pin: resize_bicubic_u8 is checked bit-exactly against PIL.Image.resize of the Pillow in this image (12.2.0 -- the reference pins 9.4.0, which
cannot be installed here; the restated algorithm is the one documented for Resample.c's 8bpc path in both) in tests/test_preprocess_cpu.py, and against sha256 digests of Pillow's
outputs committed in tests/golden/preprocess_digests.json (generator: oracle/make_preprocess_golden.py).
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2            # Resample.c: PRECISION_BITS
OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def bicubic_filter(x: float) -> float:
    """Resample.c bicubic_filter, a = -0.5."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int) -> Tuple[int, np.ndarray, np.ndarray]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for box (0, in_size): (ksize, bounds[out,2], kk[out,ksize] int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """One 8bpc pass along axis 0 of a [n, m, c] uint8 array (vertical pass; the horizontal pass is its transpose)."""
    _, bounds, kk = precompute_coeffs(img.shape[0], out_size)
    src = img.astype(np.int64)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    for yy in range(out_size):
        ymin, ymax = int(bounds[yy, 0]), int(bounds[yy, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for y in range(ymax):
            acc += src[ymin + y] * int(kk[yy, y])
        out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bicubic_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), BICUBIC) of an [H, W, C] uint8 image: ImagingResample, horizontal pass first."""
    h, w = img.shape[:2]
    if (w, h) == (out_w, out_h):
        return img.copy()                            # Image.resize returns a copy; neither pass runs
    if w != out_w:
        img = _resample_axis0(img.transpose(1, 0, 2), out_w).transpose(1, 0, 2)
    if h != out_h:
        img = _resample_axis0(img, out_h)
    return np.ascontiguousarray(img)


def expand2square(img: np.ndarray, background: Sequence[int]) -> np.ndarray:
    """mm_utils.py:257-268 on an [H, W, 3] uint8 array."""
    h, w = img.shape[:2]
    if w == h:
        return img
    s = max(w, h)
    out = np.empty((s, s, 3), dtype=np.uint8)
    out[:] = np.asarray(background, dtype=np.uint8)
    if w > h:
        y0 = (w - h) // 2
        out[y0:y0 + h] = img
    else:
        x0 = (h - w) // 2
        out[:, x0:x0 + w] = img
    return out


def rescale_normalize(img_u8: np.ndarray, mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    """transformers 4.44.2 image_transforms.rescale + normalize + to_channel_dimension_format(FIRST): float32 [3, H, W]."""
    x = (img_u8 * (1 / 255)).astype(np.float32)      # uint8 * python float -> float64, then float32
    m = np.array(mean, dtype=np.float32)
    s = np.array(std, dtype=np.float32)
    x = (x - m) / s
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def preprocess_frames(frames: np.ndarray, size: int = 336, mean: Sequence[float] = OPENAI_CLIP_MEAN,
                      std: Sequence[float] = OPENAI_CLIP_STD) -> np.ndarray:
    """process_video(..., aspect_ratio='pad') on [n, H, W, 3] uint8 frames: float32 [n, 3, size, size] (the caller rounds to
    the tower's dtype, as the reference's .half() does)."""
    bg = tuple(int(x * 255) for x in mean)
    out = []
    for f in frames:
        sq = expand2square(np.asarray(f), bg)
        out.append(rescale_normalize(resize_bicubic_u8(sq, size, size), mean, std))
    return np.stack(out)
