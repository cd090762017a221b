"""Parity of the CUDA frame path (ViT -> pool -> projector step -> gate) against the oracle, through
the C ABI.  Small kernel-aligned dimensions (oracle in well under a second) and the full BASELINE
dimensions (CLIP-ViT-L/14-336, d_model 4096, 4-layer gate; oracle ~10 s per frame on 8 cores)."""
import numpy as np
import os
import pytest
import torch

from oracle import restate as R
from parity_util import (build_engine, engine_config, f32, make_weights, oracle_configs, rel_err)
from streammind_b200 import synth

pytestmark = pytest.mark.gpu

# north_star tolerance: 1e-3 relative (fp16); bf16 has 8x coarser rounding
TOL = {torch.float16: 1e-3, torch.bfloat16: 8e-3}


def _oracle_frames(sd32, oc, dt, frames_cpu):
    with R.emulate(dt):
        feats = R.clip_vision_tower(sd32, oc.vit, frames_cpu.float())
        st = R.MambaState.zeros(oc.mamba)
        toks, logits = [], []
        for t in range(feats.shape[0]):
            tok = R.projector_step(sd32, oc.mamba, R.pool_patches(feats[t]), st)
            toks.append(tok)
            logits.append(R.gate_logits_degenerate(sd32, oc.gate, tok))
    return feats, torch.stack(toks), torch.stack(logits)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("use_graphs", [False, True])
def test_small_frame_path(built_library, dt, use_graphs):
    cfg = engine_config(dt, llm_layers=0, max_frames=3, use_graphs=use_graphs)
    sd = make_weights(cfg)
    eng = build_engine(cfg, sd)
    oc = oracle_configs(cfg)
    frames = synth.make_frames(0, 0, 6, cfg.vit_image, dtype=dt)
    feats_o, toks_o, logits_o = _oracle_frames(f32(sd), oc, dt, frames)

    # (1) sub-model entry points one by one
    feats, pooled = eng.vit_encode(frames[:3].cuda())
    e = rel_err(feats, feats_o[:3])
    assert max(e) < TOL[dt], ("vit features", e)
    toks = eng.projector_step(pooled)
    e = rel_err(toks, toks_o[:3])
    assert max(e) < 2 * TOL[dt], ("projector tokens", e)
    lg = torch.stack([eng.gate_score(toks[i]) for i in range(3)])
    e = rel_err(lg, logits_o[:3])
    assert max(e) < 4 * TOL[dt], ("gate logits", e)

    # (2) the fused per-frame call, streaming: state carries over, chunks of 1, 2 and 3 frames
    eng.reset_stream()
    got_t, got_l = [], []
    for lo, hi in ((0, 1), (1, 3), (3, 6)):
        _, tk, lgd, lgh = eng.frame_step(frames[lo:hi].cuda(), want_feats=False)
        torch.cuda.synchronize()
        assert torch.equal(lgd.cpu(), lgh.clone())
        got_t.append(tk); got_l.append(lgd)
    e = rel_err(torch.cat(got_t), toks_o)
    assert max(e) < 2 * TOL[dt], ("frame_step tokens", e)
    e = rel_err(torch.cat(got_l), logits_o)
    assert max(e) < 4 * TOL[dt], ("frame_step logits", e)
    # same inputs, same state -> bit-identical outputs (deterministic reductions)
    eng.reset_stream()
    again = [eng.frame_step(frames[lo:hi].cuda())[2] for lo, hi in ((0, 1), (1, 3), (3, 6))]
    assert torch.equal(torch.cat(again), torch.cat(got_l))
    eng.close()


def test_full_size_frame_path_fp16(built_library):
    """BASELINE config 2 shapes: CLIP-ViT-L/14-336 + projector + gate, fp16, 2 frames vs the oracle,
    plus the golden fixture written from the reference's own modules (tests/golden/full_size.npz)."""
    dt = torch.float16
    cfg = engine_config(dt, small=False, llm_layers=0, max_frames=2, use_graphs=False)
    sd = make_weights(cfg)
    eng = build_engine(cfg, sd)
    oc = oracle_configs(cfg)
    frames = synth.make_frames(0, 0, 2, 336, dtype=dt)
    feats_o, toks_o, logits_o = _oracle_frames(f32(sd), oc, dt, frames)
    feats, toks, logits, _ = eng.frame_step(frames.cuda(), want_feats=True)
    torch.cuda.synchronize()
    e = rel_err(feats, feats_o)
    print("full-size ViT features rel err (max, l2):", e)
    assert max(e) < TOL[dt], ("vit features", e)
    e = rel_err(toks, toks_o)
    print("full-size projector tokens rel err:", e)
    assert max(e) < 2 * TOL[dt], ("projector tokens", e)
    e = rel_err(logits, logits_o)
    print("full-size gate logits rel err:", e, logits.cpu(), logits_o)
    assert max(e) < 4 * TOL[dt], ("gate logits", e)
    eng.close()


def test_full_size_golden_from_reference(built_library):
    """The fixture holds outputs of the REFERENCE's own Video_Mamba_seq / CLIPVisionTower at full size
    in fp32 (oracle/make_golden.py --full); the fp16 CUDA path must agree to fp16 accuracy."""
    path = os.path.join(os.path.dirname(__file__), "golden", "full_size.npz")
    z = np.load(path)
    dt = torch.float16
    seed = int(z["pg_seed"])
    cfg = engine_config(dt, small=False, llm_layers=0, max_frames=1, use_graphs=False)
    sd = {}
    sd.update(synth.make_vit_weights(seed, torch.float32))
    sd.update(synth.make_projector_gate_weights(seed, torch.float32))
    eng = build_engine(cfg, {k: v.to(dt) for k, v in sd.items()})
    # projector + gate on the fixture's feature stream
    T = int(z["pg_T"])
    g = torch.Generator().manual_seed(int(z["pg_feat_seed"]))
    feats = (torch.randn(1, T, 576, 1024, generator=g) * 1.5)[0].to(dt).cuda()
    toks = eng.projector_step(eng.pool_features(feats))
    logits = torch.stack([eng.gate_score(toks[i]) for i in range(T)])
    e = rel_err(toks[:, ::16], torch.from_numpy(z["pg_tokens"]))
    print("projector tokens vs reference fp32:", e)
    assert max(e) < 5e-3, e            # fp16 weights + activations vs fp32 reference
    e = rel_err(logits, torch.from_numpy(z["pg_logits"]))
    print("gate logits vs reference fp32:", e)
    assert max(e) < 2e-2, e
    # ViT on frame (stream 0, t 0)
    px = synth.make_frames(0, 0, 1, 336, dtype=torch.float32).to(dt).cuda()
    f, pooled = eng.vit_encode(px)
    e = rel_err(f[0, ::48, ::64], torch.from_numpy(z["vit_feats_sub"]))
    print("ViT features vs reference fp32:", e)
    assert max(e) < 5e-3, e
    eng.close()
