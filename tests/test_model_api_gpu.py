"""The reference-facing model API beyond the basic streaming call (tests/test_stream_api_gpu.py): frames encoded ahead
through the pipelined path (and under a running decode), ``forward`` with the device-resident KV cache,
``load_pretrained_model``, offline ``generate`` next to a live stream, and B streams on one engine."""
import json
import os

import pytest
import torch

from oracle import restate as R
from parity_util import check_close, engine_config, f32, make_weights, oracle_configs, rel_err
from streammind_b200 import synth
from streammind_b200.model import MultiStreamSession, StreamMindB200ForCausalLM

pytestmark = pytest.mark.gpu
GATE_TOL = {torch.float16: 4e-3, torch.bfloat16: 3.2e-2}


def _run_stream(model, frames, force, prompt0, turn_suffix, new, prefetch=0):
    """The demo loop: one frame per call, prompt grows after every fire.  prefetch > 0: keep that many frames submitted ahead."""
    prompt, outs, logits = list(prompt0), [], []
    submitted = 0
    for t in range(frames.shape[0]):
        if prefetch:
            hi = min(frames.shape[0], t + 1 + prefetch)
            if hi > submitted:
                model.prefetch_frames(frames[max(submitted, t):hi])
                submitted = hi
        out, pred = model.stream_generate_demo(torch.tensor([prompt]), images_or_videos=frames[t:t + 1], modal_list=["video"],
                                               do_sample=False, max_new_tokens=new, use_cache=True, force_pred=force[t])
        outs.append(out)
        logits.append(model.last_gate_logits.clone())
        if pred:
            prompt = prompt + out + turn_suffix
    return outs, torch.stack(logits)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_prefetched_frames_match_serial_and_overlap_decode(built_library, dt):
    """stream_generate_demo fed from prefetch_frames (pipelined towers in chunks of 8, batched gate, frames encoded while
    the previous fire decodes) returns the same decisions' logits (to rounding) and exactly the same ids as the serial path."""
    cfg = engine_config(dt, max_frames=1, use_graphs=True)
    sd = make_weights(cfg, llm=True)
    prompt0, turn_suffix = synth.make_prompt_ids(vocab=cfg.llm_vocab, n_sys=12, n_suffix=3)
    frames = synth.make_frames(5, 0, 20, cfg.vit_image, dtype=dt)
    force = [1 if t % 5 == 4 else 0 for t in range(20)]
    serial = StreamMindB200ForCausalLM(cfg, sd)
    outs_s, lg_s = _run_stream(serial, frames, force, prompt0, turn_suffix, new=6)
    serial.engine.close()
    piped = StreamMindB200ForCausalLM(cfg, sd)
    outs_p, lg_p = _run_stream(piped, frames.pin_memory(), force, prompt0, turn_suffix, new=6, prefetch=12)
    assert outs_p == outs_s, (outs_p, outs_s)
    check_close("gate logits, pipelined vs serial", lg_p, lg_s, GATE_TOL[dt])
    assert piped.interval_id_list == [5, 10, 15, 20]
    piped.engine.close()


def test_forward_logits_and_kv_token(built_library):
    dt = torch.bfloat16
    cfg = engine_config(dt, max_frames=2, use_graphs=False)
    sd = make_weights(cfg, llm=True)
    model = StreamMindB200ForCausalLM(cfg, sd)
    oc, sd32 = oracle_configs(cfg), f32(sd)
    frames = synth.make_frames(2, 0, 3, cfg.vit_image, dtype=dt)
    ids = [1, 17, 33, -201, 5, 9]
    out = model(input_ids=torch.tensor([ids]), images=[frames.cuda()], use_cache=True)
    with R.emulate(dt):
        feats = R.clip_vision_tower(sd32, oc.vit, frames.float())
        st = R.MambaState.zeros(oc.mamba)
        toks = torch.stack([R.projector_step(sd32, oc.mamba, R.pool_patches(feats[t]), st) for t in range(3)])
        emb = R.splice_prompt(sd32, ids, toks, [3])
        cache = R.KVCache()
        lg = R.mistral_forward(sd32, "", oc.llm, emb, cache)
        lg2 = R.mistral_forward(sd32, "", oc.llm, sd32["model.embed_tokens.weight"][torch.tensor([7, 8])], cache)
    assert out.logits.shape == (1, 1, cfg.llm_vocab)
    e = rel_err(out.logits[0, 0], lg)
    assert max(e) < 3e-2, e
    assert out.past_key_values.length == len(ids) - 1 + 3 == model.engine.kv_len
    out2 = model(input_ids=torch.tensor([[7, 8]]), past_key_values=out.past_key_values)
    assert model.engine.kv_len == len(ids) + 2 + 2
    e = rel_err(out2.logits[0, 0], lg2)
    assert max(e) < 3e-2, e
    out3 = model(input_ids=torch.tensor([[7, 8]]))                      # no past: a fresh sequence
    assert model.engine.kv_len == 2 and out3.past_key_values.length == 2
    with pytest.raises(ValueError):
        model(input_ids=torch.tensor([[1, -200, 5]]))                   # unexpanded <image> sentinel: rejected like nn.Embedding would
    model.engine.close()


def test_load_pretrained_model(built_library, tmp_path):
    from safetensors.torch import save_file
    from streammind_b200.builder import load_pretrained_model
    dt = torch.float16
    cfg = engine_config(dt, max_frames=1, use_graphs=False)
    sd = make_weights(cfg, llm=True)
    hf = dict(architectures=["Videollama2MistralForCausalLM"], hidden_size=cfg.llm_hidden, num_hidden_layers=cfg.llm_layers,
              num_attention_heads=cfg.llm_heads, num_key_value_heads=cfg.llm_kv_heads, head_dim=cfg.llm_head_dim,
              intermediate_size=cfg.llm_ffn, vocab_size=cfg.llm_vocab, rms_norm_eps=cfg.llm_eps, rope_theta=cfg.llm_rope_theta,
              torch_dtype="float16", mm_vision_select_layer=-2, max_sequence_length=512, llm_max_ctx=512,
              vision_config=dict(image_size=cfg.vit_image, patch_size=cfg.vit_patch, hidden_size=cfg.vit_hidden,
                                 num_hidden_layers=cfg.vit_layers + 1, num_attention_heads=cfg.vit_heads, intermediate_size=cfg.vit_ffn),
              gate_layers=cfg.gate_layers, gate_heads=cfg.gate_heads, gate_kv_heads=cfg.gate_kv_heads, gate_head_dim=cfg.gate_head_dim,
              gate_ffn=cfg.gate_ffn)
    (tmp_path / "config.json").write_text(json.dumps(hf))
    keys = sorted(sd)
    save_file({k: sd[k].contiguous() for k in keys[: len(keys) // 2]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    torch.save({k: sd[k] for k in keys[len(keys) // 2:]}, str(tmp_path / "pytorch_model-00002-of-00002.bin"))
    tokenizer, model, processor, ctx_len = load_pretrained_model(str(tmp_path), None, "videollama2-mistral-test")
    assert tokenizer is None and ctx_len == 512 and model.config.llm_layers == cfg.llm_layers and model.config.vit_layers == cfg.vit_layers
    ref = StreamMindB200ForCausalLM(cfg, sd)
    import numpy as np
    rgb = np.random.default_rng(0).integers(0, 256, size=(1, 90, 120, 3), dtype=np.uint8)
    px = processor.preprocess([rgb[0]])["pixel_values"]
    assert px.shape == (1, 3, cfg.vit_image, cfg.vit_image) and px.dtype == dt
    prompt0, _ = synth.make_prompt_ids(vocab=cfg.llm_vocab, n_sys=8, n_suffix=2)
    a = model.stream_generate_demo(torch.tensor([prompt0]), images_or_videos=px, do_sample=False, max_new_tokens=5, force_pred=1)
    b = ref.stream_generate_demo(torch.tensor([prompt0]), images_or_videos=px, do_sample=False, max_new_tokens=5, force_pred=1)
    assert a == b and len(a[0]) == 5
    assert torch.equal(model.last_gate_logits, ref.last_gate_logits)
    model.engine.close(); ref.engine.close()


def test_offline_generate_leaves_the_live_stream_alone(built_library):
    dt = torch.bfloat16
    cfg = engine_config(dt, max_frames=2, use_graphs=False, n_streams=2)
    sd = make_weights(cfg, llm=True)
    prompt0, turn_suffix = synth.make_prompt_ids(vocab=cfg.llm_vocab, n_sys=10, n_suffix=3)
    frames = synth.make_frames(7, 0, 6, cfg.vit_image, dtype=dt)
    force = [0, 1, 0, 0, 1, 1]
    plain = StreamMindB200ForCausalLM(cfg, sd)
    outs_ref, lg_ref = _run_stream(plain, frames, force, prompt0, turn_suffix, new=5)
    plain.engine.close()
    model = StreamMindB200ForCausalLM(cfg, sd)
    a, la = _run_stream(model, frames[:3], force[:3], prompt0, turn_suffix, new=5)
    off = model.generate(torch.tensor([[1, 9, -201, 4]]), images_or_videos=frames[4:6], do_sample=False, max_new_tokens=4)
    assert off.shape == (1, 4)
    prompt = list(prompt0) + a[1] + turn_suffix
    rest = []
    for t in range(3, 6):
        out, pred = model.stream_generate_demo(torch.tensor([prompt]), images_or_videos=frames[t:t + 1], do_sample=False,
                                               max_new_tokens=5, force_pred=force[t])
        rest.append(out)
        if pred:
            prompt = prompt + out + turn_suffix
    assert a + rest == outs_ref
    model.engine.close()


@pytest.mark.parametrize("dt", [torch.bfloat16])
def test_multi_stream_session_equals_independent_streams(built_library, dt):
    """Three streams on one engine: shared tower batch, shared projector / gate weight passes, firing streams decoded
    together -- same ids as three independent single-stream models, gate logits equal to rounding."""
    B, T, NEW = 3, 6, 5
    cfg = engine_config(dt, max_frames=1, use_graphs=False)
    sd = make_weights(cfg, llm=True)
    prompt0, turn_suffix = synth.make_prompt_ids(vocab=cfg.llm_vocab, n_sys=10, n_suffix=3)
    frames = [synth.make_frames(11 + s, 0, T, cfg.vit_image, dtype=dt) for s in range(B)]
    force = [[1 if (t + s) % 3 == 2 else 0 for t in range(T)] for s in range(B)]
    ref_outs, ref_lg = [], []
    for s in range(B):
        m = StreamMindB200ForCausalLM(cfg, sd)
        o, lg = _run_stream(m, frames[s], force[s], prompt0, turn_suffix, new=NEW)
        ref_outs.append(o); ref_lg.append(lg)
        m.engine.close()
    sess = MultiStreamSession(cfg, sd, n_streams=B)
    prompts = [list(prompt0) for _ in range(B)]
    for t in range(T):
        batch = torch.cat([frames[s][t:t + 1] for s in range(B)]).cuda()
        res = sess.stream_generate_demo_multi(prompts, batch, force_pred=[force[s][t] for s in range(B)], do_sample=False, max_new_tokens=NEW)
        for s, (out, pred) in enumerate(res):
            assert out == ref_outs[s][t], (s, t, out, ref_outs[s][t])
            check_close(f"stream {s} frame {t} gate logits", sess.streams[s].last_gate_logits, ref_lg[s][t], GATE_TOL[dt])
            if pred:
                prompts[s] = prompts[s] + out + turn_suffix
    sess.close()


def test_full_width_stream_generate_demo_ids_exact(built_library):
    """BASELINE widths end to end through the model API: CLIP-ViT-L/14-336 (23 layers) + projector + gate + a TWO-layer
    Mistral at full width (hidden 4096, FFN 14336, 32 q / 8 kv heads, vocab 32002), bf16, 4 frames, fires on frames 1 and 3
    (prefix re-use on the second fire), ids compared with the oracle's incremental twin."""
    dt = torch.bfloat16
    cfg = engine_config(dt, small=False, llm_layers=2, llm_max_ctx=1024, max_frames=1, use_graphs=True)
    sd = make_weights(cfg, llm=True)
    model = StreamMindB200ForCausalLM(cfg, sd)
    oc, sd32 = oracle_configs(cfg), f32(sd)
    prompt0, turn_suffix = synth.make_prompt_ids(vocab=32000, n_sys=20, n_suffix=4)
    frames = synth.make_frames(1, 0, 4, 336, dtype=dt)
    force, NEW = [0, 1, 0, 1], 8
    with R.emulate(dt):
        ora = R.IncrementalStream(sd32, oc)
        po, exp = list(prompt0), []
        for t in range(4):
            out, pred, lg, _ = ora.step(po, frames[t:t + 1].float(), NEW, force_pred=force[t])
            exp.append((out, ora.prefilled if pred else 0))
            if pred:
                po = po + out + turn_suffix
    prompt = list(prompt0)
    for t in range(4):
        out, pred = model.stream_generate_demo(torch.tensor([prompt]), images_or_videos=frames[t:t + 1], do_sample=False,
                                               max_new_tokens=NEW, force_pred=force[t])
        if pred:
            assert model.last_prefill_len == exp[t][1]
            assert out == exp[t][0], (t, out, exp[t][0])
            prompt = prompt + out + turn_suffix
    model.engine.close()
