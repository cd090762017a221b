"""Second parity witness for the vision tower: the reference's REAL GPU arithmetic -- hf CLIPVisionModel
run by PyTorch on the B200 in fp16 / bf16, the very thing CLIPVisionTower.forward executes -- against the CUDA tower, at the
full CLIP-ViT-L/14-336 size.  Two valid 16-bit implementations of 23 layers differ by rounding noise, so the bound is stated
against the noise floor measured in the same test: both must sit within 1.5 x floor of exact (fp32) arithmetic, and of each
other within 2 x floor (floor = distance of the hf 16-bit run from the hf fp32 run on the same weights)."""
import pytest
import torch

from parity_util import build_engine, engine_config, make_weights, rel_err
from streammind_b200 import hf_reference, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_cuda_tower_vs_hf_clip_on_the_gpu(built_library, dt):
    cfg = engine_config(dt, small=False, llm_layers=0, proj_d_model=0, gate_layers=0, max_frames=2, use_graphs=False)
    sd = make_weights(cfg)
    eng = build_engine(cfg, sd)
    frames = synth.make_frames(0, 0, 2, cfg.vit_image, dtype=dt)
    feats, _ = eng.vit_encode(frames.cuda())
    kw = dict(hidden=cfg.vit_hidden, ffn=cfg.vit_ffn, layers=cfg.vit_layers + 1, heads=cfg.vit_heads, image=cfg.vit_image, patch=cfg.vit_patch,
              eps=cfg.vit_eps)
    hf16 = hf_reference.clip_features(hf_reference.build_hf_clip(sd, dtype=dt, **kw), frames)
    hf32 = hf_reference.clip_features(hf_reference.build_hf_clip(sd, dtype=torch.float32, **kw), frames.float())
    floor = rel_err(hf16, hf32)[1]
    ours_exact, ours_hf = rel_err(feats, hf32)[1], rel_err(feats, hf16)[1]
    print(f"{dt}: hf 16-bit vs hf fp32 (noise floor) {floor:.2e}; CUDA tower vs hf fp32 {ours_exact:.2e}; CUDA tower vs hf 16-bit {ours_hf:.2e}")
    assert hf16.shape == feats.shape
    assert ours_exact < max(1e-3, 1.5 * floor), (ours_exact, floor)
    assert ours_hf < max(1e-3, 2.0 * floor), (ours_hf, floor)
    eng.close()
