"""Frame preprocessing oracle (oracle/preprocess.py) against the reference's dependencies: committed sha256 digests of
Pillow's output (tests/golden/preprocess_digests.json, oracle/make_preprocess_golden.py) and Pillow itself when importable."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import preprocess as P
from preprocess_cases import CASES, make_frame

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "preprocess_digests.json")))


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_pillow_digests(name):
    h, w, seed = CASES[name]
    frame = make_frame(h, w, seed)
    bg = tuple(int(x * 255) for x in P.OPENAI_CLIP_MEAN)
    assert bg == (122, 116, 104)
    u8 = P.resize_bicubic_u8(P.expand2square(frame, bg), 336, 336)
    assert hashlib.sha256(u8.tobytes()).hexdigest() == GOLD["cases"][name]["u8_sha256"]
    px = P.preprocess_frames(frame[None])[0]
    assert px.shape == (3, 336, 336) and px.dtype == np.float32
    assert hashlib.sha256(px.astype(np.float16).tobytes()).hexdigest() == GOLD["cases"][name]["f16_sha256"]


def test_oracle_matches_pillow_live():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(11)
    for h, w, ow, oh in [(61, 97, 336, 336), (97, 61, 40, 40), (200, 300, 336, 336), (700, 700, 336, 336), (20, 20, 336, 336)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(P.resize_bicubic_u8(img, ow, oh), ref), (h, w, ow, oh)


def test_coefficients_shape_and_normalisation():
    ksize, bounds, kk = P.precompute_coeffs(1920, 336)
    assert ksize == 25 and bounds.shape == (336, 2) and kk.shape == (336, 25)
    assert np.all(np.abs(kk.sum(axis=1) - (1 << P.PRECISION_BITS)) <= ksize)      # fixed-point rows sum to ~1.0
    ksize, bounds, kk = P.precompute_coeffs(336, 336)                              # same size: identity taps
    assert np.all(kk.max(axis=1) == 1 << P.PRECISION_BITS)
