"""Offline bulk encoder on the GPU: the feature files hold exactly what the CUDA vision tower returns for those
frames (bit for bit, whatever the batch they were encoded in), within the north_star tolerance of the oracle, and encoding with
`segment` writes the thinned tensor the reference gets by encoding everything and slicing afterwards."""
import pytest
import torch

from oracle import restate as R
from parity_util import build_engine, check_close, engine_config, f32, make_weights, oracle_configs
from streammind_b200 import offline, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def test_feature_files(tmp_path, dt):
    cfg = engine_config(dt, max_frames=4, gate_layers=0, llm_layers=0)
    sd = make_weights(cfg)
    eng = build_engine(cfg, sd)
    T = 13
    frames = synth.make_frames(0, 0, T, cfg.vit_image, dtype=dt)
    read = lambda ids: frames[ids].pin_memory()
    paths = offline.encode_video_to_files(eng, read, T, str(tmp_path / "features_video_encode_ddp" / "g"), "1", chunk_frames=5)
    full = torch.cat([torch.load(p, map_location="cpu") for p in paths], 1)
    assert full.shape == (1, T, cfg.num_patches, cfg.vit_hidden) and full.dtype == dt
    # the same tower, one frame per call: identical bits
    single = torch.cat([eng.vit_encode(frames[i:i + 1].cuda())[0] for i in range(T)], 0).cpu()
    assert torch.equal(full[0], single)
    with R.emulate(dt):
        ref = R.clip_vision_tower(f32(sd), oracle_configs(cfg).vit, frames.float())
    check_close("offline features", full[0], ref, 1e-3 if dt == torch.float16 else 8e-3)
    # segment: encode-then-thin == encode the thinned frames
    thin = [torch.load(offline.thin_feature_file(p, 3)) for p in paths]
    direct = [torch.load(p, map_location="cpu") for p in offline.encode_video_to_files(eng, read, T, str(tmp_path / "d"), "1", chunk_frames=5, segment=3)]
    for a, b in zip(thin, direct):
        assert torch.equal(a, b)
    # pooled files == mean over the patches as the projector takes it
    pooled = torch.cat([torch.load(p, map_location="cpu") for p in offline.encode_video_to_files(eng, read, T, str(tmp_path / "p"), "1", chunk_frames=5, pooled=True)], 1)
    assert torch.equal(pooled[0], eng.pool_features(full[0].cuda()).cpu())
    eng.close()
