"""sm_preprocess_frames (device bicubic resample + normalise) against the oracle restatement of the reference's
PIL + CLIPImageProcessor arithmetic: bit-exact on the model-dtype outputs, and against the committed Pillow digests."""
import hashlib
import json
import os
import time

import numpy as np
import pytest
import torch

from oracle import preprocess as P
from parity_util import engine_config
from preprocess_cases import CASES, make_frame

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "preprocess_digests.json")))


@pytest.fixture(scope="module")
def engines(built_library):
    from streammind_b200.engine import Engine
    e = {dt: Engine(engine_config(dt, small=False, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
         for dt in (torch.float16, torch.bfloat16)}
    yield e
    for x in e.values():
        x.close()


@pytest.mark.parametrize("name", list(CASES))
def test_preprocess_bit_exact_fp16(engines, name):
    h, w, seed = CASES[name]
    frame = make_frame(h, w, seed)
    out = engines[torch.float16].preprocess_frames(frame[None])
    torch.cuda.synchronize()
    assert out.shape == (1, 3, 336, 336) and out.dtype == torch.float16
    got = out[0].cpu().numpy()
    ref = P.preprocess_frames(frame[None])[0].astype(np.float16)
    assert np.array_equal(got.view(np.uint16), ref.view(np.uint16)), np.abs(got.astype(np.float32) - ref.astype(np.float32)).max()
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == GOLD["cases"][name]["f16_sha256"]


def test_preprocess_batch_device_input_bf16(engines):
    """A batch of frames already on the device, bf16 tower: every frame equals the oracle rounded to bf16."""
    frames = np.stack([make_frame(240, 426, s) for s in range(5)])
    out = engines[torch.bfloat16].preprocess_frames(torch.from_numpy(frames).cuda())
    torch.cuda.synchronize()
    ref = torch.from_numpy(P.preprocess_frames(frames)).to(torch.bfloat16)
    assert torch.equal(out.cpu(), ref)


def test_process_video_mirror_and_timing(engines):
    """mm_utils.process_video mirror (list of HD frames, processor-like object), plus the device time per 1080p frame."""
    from streammind_b200 import mm_utils

    class Proc:
        image_mean = list(P.OPENAI_CLIP_MEAN)
        image_std = list(P.OPENAI_CLIP_STD)

    eng = engines[torch.float16]
    frames = [make_frame(1080, 1920, 20 + i) for i in range(4)]
    out = mm_utils.process_video(frames, Proc(), "pad", engine=eng)
    ref = torch.from_numpy(P.preprocess_frames(np.stack(frames))).half()
    assert torch.equal(out.cpu(), ref)
    dev = torch.from_numpy(np.stack(frames)).cuda()
    for _ in range(3):
        eng.preprocess_frames(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.preprocess_frames(dev)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (20 * 4)
    t0 = time.time(); P.preprocess_frames(np.stack(frames[:1])); cpu_ms = (time.time() - t0) * 1e3
    mb = (1080 * 1920 * 3 + 2 * 1920 * 336 * 3 + 3 * 336 * 336 * 2) / 1e6
    print(f"\npreprocess 1080p -> 336: {us:.1f} us per frame on the device ({mb / us * 1e3:.0f} GB/s of {mb:.1f} MB algorithmic traffic); "
          f"numpy oracle {cpu_ms:.0f} ms per frame")


@pytest.mark.parametrize("image", [70, 56])      # 70: byte-wise vertical pass (70 * 3 not a multiple of 4); 56: the 32-bit one
def test_preprocess_other_tower_sizes(built_library, image):
    from streammind_b200.engine import Engine
    eng = Engine(engine_config(torch.float16, vit_image=image, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
    frames = np.stack([make_frame(123, 211, 40), make_frame(123, 211, 41)])
    out = eng.preprocess_frames(frames)
    torch.cuda.synchronize()
    ref = torch.from_numpy(P.preprocess_frames(frames, size=image)).half()
    assert out.shape == (2, 3, image, image) and torch.equal(out.cpu(), ref)
    eng.close()
