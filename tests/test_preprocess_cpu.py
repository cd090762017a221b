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


def test_library_tap_tables_match_oracle(built_library):
    """The C++ table builder inside the library (host code, no GPU) against the oracle's precompute_coeffs for many
    geometries: down- and up-scaling, odd sizes, the identity."""
    import ctypes as C
    from streammind_b200 import lib as L
    lib = L.load()
    sizes = [(s, 336) for s in (20, 30, 61, 97, 335, 336, 337, 480, 500, 640, 720, 1080, 1280, 1920, 2160, 3840, 4001)]
    sizes += [(640, 112), (97, 40), (1000, 224), (224, 448)]
    for in_size, out_size in sizes:
        ksize, bounds, kk = P.precompute_coeffs(in_size, out_size)
        k = C.c_int(0)
        b = np.zeros((out_size, 2), dtype=np.int32)
        t = np.zeros((out_size, ksize), dtype=np.int32)
        rc = lib.sm_resample_table(in_size, out_size, C.byref(k), b.ctypes.data, t.ctypes.data, t.size)
        assert rc == 0 and k.value == ksize, (in_size, out_size, rc, k.value, ksize)
        assert np.array_equal(b, bounds) and np.array_equal(t, kk), (in_size, out_size)


def test_process_video_mirror_host_behaviour():
    """mm_utils.process_video: list / array inputs are stacked, the processor's mean / std are forwarded, anything but
    aspect_ratio='pad' is refused (the reference's other branch has no device path here)."""
    from streammind_b200 import mm_utils

    class FakeEngine:
        def preprocess_frames(self, frames, mean, std):
            self.seen = (np.asarray(frames).shape, tuple(mean), tuple(std))
            return "pixels"

    class Proc:
        image_mean = [0.5, 0.4, 0.3]
        image_std = [0.2, 0.2, 0.1]

    eng = FakeEngine()
    frames = [make_frame(20, 30, 1), make_frame(20, 30, 2)]
    assert mm_utils.process_video(frames, Proc(), "pad", engine=eng) == "pixels"
    assert eng.seen == ((2, 20, 30, 3), (0.5, 0.4, 0.3), (0.2, 0.2, 0.1))
    mm_utils.process_image(frames[0], None, "pad", engine=eng)
    assert eng.seen[0] == (1, 20, 30, 3) and eng.seen[1] == P.OPENAI_CLIP_MEAN and eng.seen[2] == P.OPENAI_CLIP_STD
    with pytest.raises(NotImplementedError):
        mm_utils.process_video(frames, Proc(), "resize", engine=eng)
