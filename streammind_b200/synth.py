"""Synthetic (random-init, seeded) weights, frames and prompts of the architectures on the hot path.

There is no network and no checkpoint in this environment (BASELINE.json: "random-init
CLIP-ViT-L/14 + Mistral-7B", "synthetic 336x336 streaming frames"), so both the benchmark and the
parity tests draw weights here.  Keys are the reference model's own ``state_dict()`` keys
(``Videollama2MistralForCausalLM``: /root/reference/streammind/model/language_model/
videollama2_mistral.py:146 with the vision tower and projector attached by
/root/reference/streammind/model/videollama2_arch.py:29-34), so a real checkpoint's state_dict can be
handed to :class:`streammind_b200.engine.Engine` unchanged.

Initialisation (adjusted so activations are numerically non-degenerate):
linear weights N(0, gain/sqrt(fan_in)); LayerNorm/RMSNorm weights 1 + 0.1 N(0,1); biases
0.05 N(0,1); Mamba A_log / D / dt_proj exactly as its constructor does
(/root/reference/streammind/model/mamba_ssm/modules/mamba_simple.py:82-115).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch

VIT_PREFIX = "model.vision_tower.vision_tower.vision_model."
PROJ_PREFIX = "model.mm_projector."
GATE_PREFIX = "model.mm_projector.cls_net.cls_model."


def _gen(seed: int, device="cpu") -> torch.Generator:
    return torch.Generator(device=device).manual_seed(int(seed))


def _lin(g, out_f, in_f, dtype, device, gain=1.0):
    w = torch.randn(out_f, in_f, generator=g, device=device, dtype=torch.float32)
    return (w * (gain / math.sqrt(in_f))).to(dtype)


def _norm_w(g, n, dtype, device):
    return (1.0 + 0.1 * torch.randn(n, generator=g, device=device)).to(dtype)


def _bias(g, n, dtype, device, s=0.05):
    return (s * torch.randn(n, generator=g, device=device)).to(dtype)


def make_vit_weights(seed: int, dtype=torch.float16, device="cpu", hidden=1024, ffn=4096, layers=24,
                     heads=16, image_size=336, patch=14) -> Dict[str, torch.Tensor]:
    """CLIPVisionModel weights (hf CLIPVisionTransformer: embeddings, pre_layrnorm, encoder.layers.N,
    post_layernorm)."""
    g = _gen(seed * 7919 + 1, device)
    p = VIT_PREFIX
    n_pos = (image_size // patch) ** 2 + 1
    sd = {
        p + "embeddings.class_embedding": _bias(g, hidden, dtype, device, 0.5),
        p + "embeddings.patch_embedding.weight":
            (torch.randn(hidden, 3, patch, patch, generator=g, device=device) / math.sqrt(3 * patch * patch)).to(dtype),
        p + "embeddings.position_embedding.weight":
            (0.3 * torch.randn(n_pos, hidden, generator=g, device=device)).to(dtype),
        p + "pre_layrnorm.weight": _norm_w(g, hidden, dtype, device),
        p + "pre_layrnorm.bias": _bias(g, hidden, dtype, device),
        p + "post_layernorm.weight": _norm_w(g, hidden, dtype, device),
        p + "post_layernorm.bias": _bias(g, hidden, dtype, device),
    }
    for i in range(layers):
        lp = f"{p}encoder.layers.{i}."
        for nm in ("layer_norm1", "layer_norm2"):
            sd[lp + nm + ".weight"] = _norm_w(g, hidden, dtype, device)
            sd[lp + nm + ".bias"] = _bias(g, hidden, dtype, device)
        for nm, gain in (("q_proj", 1.6), ("k_proj", 1.6), ("v_proj", 1.0), ("out_proj", 0.7)):
            sd[lp + f"self_attn.{nm}.weight"] = _lin(g, hidden, hidden, dtype, device, gain)
            sd[lp + f"self_attn.{nm}.bias"] = _bias(g, hidden, dtype, device)
        sd[lp + "mlp.fc1.weight"] = _lin(g, ffn, hidden, dtype, device, 1.0)
        sd[lp + "mlp.fc1.bias"] = _bias(g, ffn, dtype, device)
        sd[lp + "mlp.fc2.weight"] = _lin(g, hidden, ffn, dtype, device, 0.7)
        sd[lp + "mlp.fc2.bias"] = _bias(g, hidden, dtype, device)
    return sd


def make_mistral_weights(seed: int, prefix: str, dtype, device="cpu", hidden=4096, ffn=14336, layers=32,
                         heads=32, kv_heads=8, head_dim=128, vocab=32002, with_embed=True,
                         with_qk=True) -> Dict[str, torch.Tensor]:
    """MistralForCausalLM weights under ``prefix`` ('' for the LLM, GATE_PREFIX for the gate)."""
    g = _gen(seed * 7919 + (2 if prefix == "" else 3), device)
    sd = {}
    if with_embed:
        sd[prefix + "model.embed_tokens.weight"] = torch.randn(vocab, hidden, generator=g, device=device).to(dtype)
    for i in range(layers):
        lp = f"{prefix}model.layers.{i}."
        sd[lp + "input_layernorm.weight"] = _norm_w(g, hidden, dtype, device)
        sd[lp + "post_attention_layernorm.weight"] = _norm_w(g, hidden, dtype, device)
        if with_qk:
            sd[lp + "self_attn.q_proj.weight"] = _lin(g, heads * head_dim, hidden, dtype, device, 1.5)
            sd[lp + "self_attn.k_proj.weight"] = _lin(g, kv_heads * head_dim, hidden, dtype, device, 1.5)
        sd[lp + "self_attn.v_proj.weight"] = _lin(g, kv_heads * head_dim, hidden, dtype, device, 1.0)
        sd[lp + "self_attn.o_proj.weight"] = _lin(g, hidden, heads * head_dim, dtype, device, 0.7)
        sd[lp + "mlp.gate_proj.weight"] = _lin(g, ffn, hidden, dtype, device, 1.0)
        sd[lp + "mlp.up_proj.weight"] = _lin(g, ffn, hidden, dtype, device, 1.0)
        sd[lp + "mlp.down_proj.weight"] = _lin(g, hidden, ffn, dtype, device, 1.0)
    sd[prefix + "model.norm.weight"] = _norm_w(g, hidden, dtype, device)
    sd[prefix + "lm_head.weight"] = _lin(g, vocab, hidden, dtype, device, 1.0)
    return sd


def make_projector_weights(seed: int, dtype=torch.float16, device="cpu", d_model=4096, mm_hidden=1024,
                           d_state=16, d_conv=4, expand=2) -> Dict[str, torch.Tensor]:
    """Video_Mamba_seq minus the gate: PreNet, VideoMamba(1 block), PostNet
    (/root/reference/streammind/model/multimodal_projector/builder.py:390-400)."""
    g = _gen(seed * 7919 + 4, device)
    p, mp = PROJ_PREFIX, PROJ_PREFIX + "mamba_model.ssms.0."
    d_inner, dt_rank = expand * d_model, math.ceil(d_model / 16)
    sd = {
        p + "pre_net.fc3.weight": _lin(g, d_model, mm_hidden, dtype, device, 1.4),
        p + "pre_net.fc3.bias": _bias(g, d_model, dtype, device),
        mp + "norm.weight": _norm_w(g, d_model, dtype, device),
        mp + "norm.bias": _bias(g, d_model, dtype, device),
        mp + "mixer.in_proj.weight": _lin(g, 2 * d_inner, d_model, dtype, device, 1.0),
        mp + "mixer.conv1d.weight":
            ((torch.rand(d_inner, 1, d_conv, generator=g, device=device) - 0.5) * (2.0 / math.sqrt(d_conv))).to(dtype),
        mp + "mixer.conv1d.bias":
            ((torch.rand(d_inner, generator=g, device=device) - 0.5) * (2.0 / math.sqrt(d_conv))).to(dtype),
        mp + "mixer.x_proj.weight": _lin(g, dt_rank + 2 * d_state, d_inner, dtype, device, 1.0),
        mp + "mixer.dt_proj.weight":
            ((torch.rand(d_inner, dt_rank, generator=g, device=device) * 2 - 1) * dt_rank ** -0.5).to(dtype),
        mp + "mixer.out_proj.weight": _lin(g, d_model, d_inner, dtype, device, 1.0),
        p + "mamba_model.norm_fn.weight": _norm_w(g, d_model, dtype, device),
        p + "mamba_model.norm_fn.bias": _bias(g, d_model, dtype, device),
        p + "post_net.fc3.weight": _lin(g, d_model, d_model, dtype, device, 1.4),
        p + "post_net.fc3.bias": _bias(g, d_model, dtype, device),
    }
    dt = torch.exp(torch.rand(d_inner, generator=g, device=device) * (math.log(0.1) - math.log(0.001))
                   + math.log(0.001)).clamp(min=1e-4)
    sd[mp + "mixer.dt_proj.bias"] = (dt + torch.log(-torch.expm1(-dt))).to(dtype)
    # A_log and D are kept fp32 by the reference constructor but ``model.half()`` /
    # from_pretrained(torch_dtype=...) casts every parameter, so they arrive in model dtype.
    sd[mp + "mixer.A_log"] = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32, device=device)
                                       ).repeat(d_inner, 1).to(dtype)
    sd[mp + "mixer.D"] = torch.ones(d_inner, device=device).to(dtype)
    return sd


def make_projector_gate_weights(seed: int, dtype=torch.float16, device="cpu", d_model=4096,
                                mm_hidden=1024, gate_ffn=14336, gate_heads=32, gate_kv_heads=8,
                                gate_head_dim=128, gate_layers=4, gate_with_qk=False) -> Dict[str, torch.Tensor]:
    """Projector + gate.  The gate's q_proj / k_proj never reach its output at L = 1 (SURVEY.md section
    8c (i)) and are only generated on request (the CPU reference arm executes them like the reference)."""
    sd = make_projector_weights(seed, dtype, device, d_model, mm_hidden)
    sd.update(make_mistral_weights(seed, GATE_PREFIX, dtype, device, hidden=d_model, ffn=gate_ffn,
                                   layers=gate_layers, heads=gate_heads, kv_heads=gate_kv_heads,
                                   head_dim=gate_head_dim, vocab=2, with_embed=False, with_qk=gate_with_qk))
    return sd


def make_frames(stream_id: int, t0: int, n: int, image_size: int = 336, dtype=torch.float16,
                device="cpu") -> torch.Tensor:
    """Frames t0..t0+n of stream ``stream_id``: N(0,1) per pixel (stands in for CLIP-normalised
    pixels), seeded per frame with stream_id*100003 + t."""
    out = torch.empty(n, 3, image_size, image_size, dtype=dtype)
    for i in range(n):
        g = _gen(stream_id * 100003 + t0 + i)
        out[i] = torch.randn(3, image_size, image_size, generator=g).to(dtype)
    return out.to(device)


def make_prompt_ids(vocab: int = 32000, n_sys: int = 60, n_suffix: int = 5, seed: int = 7) -> Tuple[list, list]:
    """Token-id-space stand-in for the LLAMA_2 template (/root/reference/streammind/conversation.py:78-98):
    ``[BOS] + n_sys ids + [<video>=-201] + n_suffix ids``; and the per-turn growth suffix
    ``[</s>=2] + 4 ids + [-201] + 5 ids`` (shape of video_score_stream_demo.py:124)."""
    g = _gen(seed)
    ids = torch.randint(3, vocab, (n_sys + n_suffix + 9,), generator=g).tolist()
    prompt = [1] + ids[:n_sys] + [-201] + ids[n_sys:n_sys + n_suffix]
    rest = ids[n_sys + n_suffix:]
    turn_suffix = [2] + rest[:4] + [-201] + rest[4:9]
    return prompt, turn_suffix
