"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference code.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

What it does
  1. installs oracle/shims.py and instantiates the reference's own classes
     (Videollama2MistralForCausalLM + CLIPVisionTower + Video_Mamba_seq + ClsNet) at small
     dimensions, fp32, seeded;
  2. drives ``model.stream_generate_demo`` frame by frame in token-id space with a fake tokenizer
     (the shape of streammind/eval/video_score_stream_demo.py:66-125,283-302), once with the model's
     own gate decision and once with the decision forced (the authors' ``# pred = 1`` switch,
     videollama2_arch.py:943);
  3. records inputs, all weights and every intermediate the B200 path must reproduce
     (ViT features, projector tokens, gate logits, decisions, generated ids);
  4. asserts that oracle/restate.py reproduces all of it before writing the fixture.

The only deviation from the reference's hard-coded sizes: ``MistralConfig()`` inside ClsNet
(multimodal_projector/builder.py:373) is given small defaults so the gate fits a fixture; the code
path is unchanged.  A full-size gate / projector / ViT fixture (weights regenerated from a seed,
outputs sub-sampled) is written by ``--full``.
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import restate as R  # noqa: E402
from oracle import weights as W  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class FakeTokenizer:
    """Token-id-space tokenizer: batch_decode returns the ids joined by spaces."""
    pad_token_id = 0
    eos_token_id = 2
    bos_token_id = 1

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(str(int(t)) for t in row) for row in ids]


def build_reference_model(tiny: W.TinyDims, seed: int):
    from oracle import shims
    shims.install()
    import transformers
    from transformers import CLIPVisionConfig, CLIPVisionModel, CLIPImageProcessor, MistralConfig
    import videollama2.model.multimodal_projector.builder as pb
    from videollama2.model.language_model.videollama2_mistral import (
        Videollama2MistralConfig, Videollama2MistralForCausalLM)

    tmp = tempfile.mkdtemp(prefix="tiny_clip_")          # name must contain 'clip' (encoder/builder.py:9)
    vcfg = CLIPVisionConfig(hidden_size=tiny.vit_hidden, intermediate_size=tiny.vit_ffn,
                            num_hidden_layers=tiny.vit_layers, num_attention_heads=tiny.vit_heads,
                            image_size=tiny.image_size, patch_size=tiny.patch_size,
                            hidden_act="quick_gelu", layer_norm_eps=1e-5)
    torch.manual_seed(seed)
    CLIPVisionModel(vcfg).save_pretrained(tmp)
    CLIPImageProcessor(size={"shortest_edge": tiny.image_size},
                       crop_size={"height": tiny.image_size, "width": tiny.image_size}).save_pretrained(tmp)

    real_cfg = MistralConfig

    def small_gate_cfg(*a, **k):
        return real_cfg(hidden_size=tiny.hidden, intermediate_size=tiny.gate_ffn,
                        num_attention_heads=tiny.gate_heads, num_key_value_heads=tiny.gate_kv_heads,
                        head_dim=tiny.hidden // tiny.gate_heads, max_position_embeddings=512)
    pb.MistralConfig = small_gate_cfg

    cfg = Videollama2MistralConfig(
        vocab_size=tiny.vocab, hidden_size=tiny.hidden, intermediate_size=tiny.llm_ffn,
        num_hidden_layers=tiny.llm_layers, num_attention_heads=tiny.llm_heads,
        num_key_value_heads=tiny.llm_kv_heads, head_dim=tiny.hidden // tiny.llm_heads,
        rms_norm_eps=1e-5, rope_theta=1e6, sliding_window=None, max_position_embeddings=4096,
        pad_token_id=0, bos_token_id=1, eos_token_id=2)
    cfg.mm_vision_tower = tmp
    cfg.mm_projector_type = "mamba"
    cfg.mm_hidden_size = tiny.vit_hidden
    cfg.mm_vision_select_layer = -2
    cfg.mm_vision_select_feature = "patch"
    torch.manual_seed(seed + 1)
    model = Videollama2MistralForCausalLM(cfg)
    model.get_vision_tower().load_model()
    pb.MistralConfig = real_cfg
    model.eval()
    # make the random model less degenerate: non-trivial norms / biases everywhere
    g = torch.Generator().manual_seed(seed + 2)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("A_log") or n.endswith(".D") or "dt_proj.bias" in n:
                continue
            if p.dim() == 1:
                if "norm" in n and n.endswith("weight"):
                    p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif "mm_projector" in n and "cls_net" not in n and p.dim() == 2:
                p.copy_(torch.randn(p.shape, generator=g) * (2.0 / p.shape[1]) ** 0.5)
            elif "cls_net" in n and p.dim() == 2:
                p.copy_(torch.randn(p.shape, generator=g) * (1.5 / p.shape[1]) ** 0.5)
            elif "vision_tower" in n and p.dim() == 2:
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / p.shape[1]) ** 0.5)
            elif p.dim() == 2 and "embed_tokens" not in n:
                p.copy_(torch.randn(p.shape, generator=g) * (1.5 / p.shape[1]) ** 0.5)
    return model


def run_reference_stream(model, tiny: W.TinyDims, frames: torch.Tensor, prompt0, force, max_new):
    """The demo loop (video_score_stream_demo.py:283-302,110-124) in token-id space."""
    tok = FakeTokenizer()
    model.frame_feature = None
    model.interval_id_list = []
    prompt = list(prompt0)
    rec = dict(feats=[], x_last=[], logits=[], pred=[], out=[], prompt_len=[])
    import videollama2.model.videollama2_arch as arch
    orig = model.encode_images_or_videos_score_cls_inference_allframe_demo

    grabbed = {}

    def spy(*a, **k):
        X, cls, ff, iid = orig(*a, **k)
        grabbed["X"], grabbed["cls"] = X.detach().clone(), cls.detach().clone()
        if force is not None:                       # authors' "# pred = 1" switch (arch.py:943)
            cls = torch.tensor([1.0, 0.0]) if force[iid - 1] == 0 else torch.tensor([0.0, 1.0])
        return X, cls, ff, iid
    model.encode_images_or_videos_score_cls_inference_allframe_demo = spy
    for t in range(frames.shape[0]):
        ids = torch.tensor(prompt, dtype=torch.long).unsqueeze(0)
        with torch.inference_mode():
            out, pred = model.stream_generate_demo(
                ids, attention_mask=torch.ones_like(ids), images_or_videos=frames[t:t + 1],
                modal_list=["video"], do_sample=False, max_new_tokens=max_new, use_cache=True,
                pad_token_id=tok.eos_token_id, tokenizer=tok, score_video=True)
        rec["feats"].append(model.frame_feature[0, -1].clone())
        rec["x_last"].append(grabbed["X"][0].clone())
        rec["logits"].append(grabbed["cls"].clone())
        rec["pred"].append(int(pred))
        rec["prompt_len"].append(len(prompt))
        if pred == 1:
            gen = [int(s) for s in out.split()] if out else []
            rec["out"].append(gen)
            # growth rule (video_score_stream_demo.py:123-124) in id space
            prompt = prompt + gen + list(tiny.turn_suffix_ids)
        else:
            rec["out"].append([])
    model.encode_images_or_videos_score_cls_inference_allframe_demo = orig
    return rec


def check_restatement(sd, cfgs, tiny, frames, prompt0, force, max_new, rec):
    """oracle/restate.py must reproduce the reference run before the fixture is written."""
    for cls in (R.ReferenceSemanticsStream, R.IncrementalStream):
        s = cls(sd, cfgs)
        prompt = list(prompt0)
        for t in range(frames.shape[0]):
            fp = None if force is None else force[t]
            out, pred, logits, x = s.step(prompt, frames[t:t + 1], max_new, stop_ids=(2,), force_pred=fp)
            f_ref = rec["feats"][t]
            f_me = R.clip_vision_tower(sd, cfgs.vit, frames[t:t + 1])[0]
            assert torch.allclose(f_me, f_ref, atol=2e-5, rtol=1e-5), (cls.__name__, "vit", t)
            assert torch.allclose(x[0], rec["x_last"][t], atol=5e-5, rtol=1e-4), (cls.__name__, "proj", t,
                (x[0] - rec["x_last"][t]).abs().max())
            assert torch.allclose(logits, rec["logits"][t], atol=5e-5, rtol=1e-4), (cls.__name__, "gate", t)
            assert pred == rec["pred"][t], (cls.__name__, "pred", t)
            if pred == 1:
                assert out == rec["out"][t], (cls.__name__, "ids", t, out, rec["out"][t])
                prompt = prompt + out + list(tiny.turn_suffix_ids)
        print(f"  restatement {cls.__name__}: OK")


def make_tiny(name: str, seed: int, n_frames: int, force, max_new: int):
    tiny = W.TinyDims()
    model = build_reference_model(tiny, seed)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cfgs = W.tiny_configs(tiny)
    g = torch.Generator().manual_seed(seed + 7)
    frames = torch.randn(n_frames, 3, tiny.image_size, tiny.image_size, generator=g)
    prompt0 = list(tiny.prompt_ids)
    rec = run_reference_stream(model, tiny, frames, prompt0, force, max_new)
    print(f"{name}: preds={rec['pred']} out_lens={[len(o) for o in rec['out']]}")
    check_restatement(sd, cfgs, tiny, frames, prompt0, force, max_new, rec)
    arrays = {f"w/{k}": v.numpy() for k, v in sd.items()
              if not k.endswith("embed_tokens.weight") or "cls_net" not in k}
    arrays["frames"] = frames.numpy()
    arrays["prompt0"] = np.array(prompt0, dtype=np.int64)
    arrays["force"] = np.array([-1] * n_frames if force is None else force, dtype=np.int64)
    arrays["max_new"] = np.array(max_new)
    arrays["feats"] = torch.stack(rec["feats"]).numpy()
    for t in range(n_frames):
        arrays[f"x/{t}"] = rec["x_last"][t].numpy()
        arrays[f"out/{t}"] = np.array(rec["out"][t], dtype=np.int64)
    arrays["logits"] = torch.stack(rec["logits"]).numpy()
    arrays["pred"] = np.array(rec["pred"], dtype=np.int64)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def make_full(seed: int = 1234):
    """Full-size gate + projector + 1-frame ViT through the reference's own modules, fp32, weights
    regenerated from ``oracle.weights.make_*`` (seeded); outputs stored sub-sampled."""
    from oracle import shims
    shims.install()
    import videollama2.model.multimodal_projector.builder as pb
    from videollama2.model.multimodal_encoder.clip_encoder import CLIPVisionTower
    from transformers import CLIPVisionConfig, CLIPVisionModel, CLIPImageProcessor
    import types
    torch.set_grad_enabled(False)
    out = {}
    # --- projector + gate at full size
    cfg = types.SimpleNamespace(mm_hidden_size=1024, hidden_size=4096, mm_projector_type="mamba")
    proj = pb.build_vision_projector(cfg).eval()
    sd = W.make_projector_gate_weights(seed, dtype=torch.float32)
    own = {k[len(R.PROJ_PREFIX):]: v for k, v in sd.items()}
    missing, unexpected = proj.load_state_dict(own, strict=False)
    assert not unexpected, unexpected
    assert all("embed_tokens" in m or "q_proj" in m or "k_proj" in m for m in missing), missing
    T = 6
    g = torch.Generator().manual_seed(seed + 11)
    feats = torch.randn(1, T, 576, 1024, generator=g) * 1.5
    xs, lgs = [], []
    for t in range(1, T + 1):
        x, lg = proj(feats[:, :t], cls_demo=True)
        xs.append(x[0, -1].clone()); lgs.append(lg.clone())
    xs, lgs = torch.stack(xs), torch.stack(lgs)
    mc, gc = R.MambaCfg(), R.gate_config()
    st = R.MambaState.zeros(mc)
    for t in range(T):
        tok = R.projector_step(sd, mc, R.pool_patches(feats[0, t]), st)
        lg = R.gate_logits_degenerate(sd, gc, tok)
        assert torch.allclose(tok, xs[t], atol=1e-4, rtol=1e-4), (t, (tok - xs[t]).abs().max())
        assert torch.allclose(lg, lgs[t], atol=1e-4, rtol=1e-4), (t, lg, lgs[t])
    print("  full-size projector+gate restatement: OK")
    out.update(pg_seed=np.array(seed), pg_feat_seed=np.array(seed + 11), pg_T=np.array(T),
               pg_tokens=xs.numpy()[:, ::16], pg_logits=lgs.numpy())
    # --- ViT-L/14-336 one frame
    tmp = tempfile.mkdtemp(prefix="full_clip_")
    vcfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                            num_attention_heads=16, image_size=336, patch_size=14,
                            hidden_act="quick_gelu", layer_norm_eps=1e-5)
    m = CLIPVisionModel(vcfg)
    vsd = W.make_vit_weights(seed, dtype=torch.float32)
    own = {k[len("model.vision_tower.vision_tower."):]: v for k, v in vsd.items()}
    res = m.load_state_dict(own, strict=False)
    assert not res.unexpected_keys and all("position_ids" in k for k in res.missing_keys), res
    m.save_pretrained(tmp)
    CLIPImageProcessor().save_pretrained(tmp)
    args = types.SimpleNamespace(mm_vision_select_layer=-2, mm_vision_select_feature="patch")
    tower = CLIPVisionTower(tmp, args).eval()
    px = W.make_frames(0, 0, 1, 336, dtype=torch.float32)
    f_ref = tower(px)
    f_me = R.clip_vision_tower(vsd, R.VitConfig(), px)
    err = (f_me - f_ref).abs().max().item()
    print(f"  full-size ViT restatement max|diff| = {err:.2e} (scale {f_ref.abs().max().item():.2f})")
    assert err < 2e-3 * f_ref.abs().max().item()
    out.update(vit_seed=np.array(seed), vit_feats_sub=f_ref[0, ::48, ::64].numpy(),
               vit_pooled=f_ref[0].mean(0).numpy())
    path = os.path.join(GOLDEN_DIR, "full_size.npz")
    np.savez_compressed(path, **out)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


class CharTokenizer:
    """Deterministic toy tokenizer for the host-logic fixture: BOS + one id per whitespace-separated
    word (hash-free: ids assigned in order of first appearance from a fixed vocabulary list)."""
    pad_token_id, bos_token_id, eos_token_id = 0, 1, 2

    def __init__(self):
        self.vocab = {"</s>": 2}

    def __call__(self, text):
        ids = [self.bos_token_id]
        for w in text.replace("\n", " \n ").split(" "):
            if w == "":
                continue
            ids.append(self.vocab.setdefault(w, len(self.vocab) + 3))
        return type("Enc", (), {"input_ids": ids})()

    def batch_decode(self, ids, skip_special_tokens=True):
        inv = {v: k for k, v in self.vocab.items()}
        return [" ".join(inv.get(int(t), "?") for t in row if not (skip_special_tokens and int(t) in (0, 1, 2)))
                for row in ids]


def make_host_logic():
    """Prompt template + tokenizer_MMODAL_token + KeywordsStoppingCriteria run through the REFERENCE's
    own code (conversation.py:78-98,383-393; mm_utils.py:567-647) on a toy tokenizer."""
    import json
    from oracle import shims
    shims.install()
    from videollama2.conversation import conv_templates
    from videollama2.mm_utils import tokenizer_MMODAL_token, KeywordsStoppingCriteria
    conv = conv_templates["mistral_instruct"].copy()
    conv.append_message(conv.roles[0], "<video>\n")
    conv.append_message(conv.roles[1], None)
    prompt0 = conv.get_prompt()
    tok = CharTokenizer()
    prompts = [prompt0]
    for out in ("a man opens the door", "he sits down"):
        prompts.append(prompts[-1] + " " + out + " </s>[INST] <video>\n [/INST]")
    ids = [tokenizer_MMODAL_token(p, tok, -201) for p in prompts]
    inp = torch.tensor([ids[0]])
    sc = KeywordsStoppingCriteria(["</s>"], tok, inp)
    stop_cases = []
    for tail in ([5, 6, 2], [5, 6, 7], [2]):
        o = torch.tensor([ids[0] + tail])
        stop_cases.append({"tail": tail, "stop": bool(sc(o, None))})
    path = os.path.join(GOLDEN_DIR, "host_logic.json")
    json.dump({"prompts": prompts, "ids": ids, "vocab": tok.vocab, "stop_cases": stop_cases,
               "keyword_ids": [k.tolist() for k in sc.keyword_ids]}, open(path, "w"), indent=1)
    print("  wrote", path)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--host", action="store_true")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    if a.host:
        make_host_logic()
    if a.tiny or not (a.full or a.host):
        make_tiny("tiny_model_gate", seed=11, n_frames=8, force=None, max_new=6)
        make_tiny("tiny_forced_gate", seed=23, n_frames=10, force=[0, 1, 0, 0, 1, 1, 0, 1, 0, 1], max_new=5)
    if a.full:
        make_full()
