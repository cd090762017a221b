"""Shared helpers for the parity tests: seeded weights at small-but-kernel-aligned and full sizes, the
oracle evaluated with rounding emulation, and error metrics.

Tolerance (north_star): "ViT features and gate logits within 1e-3 relative fp16".  We measure
relative error norm-wise -- max|a-b| / max|b| and ||a-b||_2 / ||b||_2 -- against the oracle computed
with exact accumulation and the reference's rounding points (oracle.restate.emulate)."""
from __future__ import annotations

import torch

from oracle import restate as R
from streammind_b200 import synth
from streammind_b200.engine import EngineConfig


def rel_err(a: torch.Tensor, b: torch.Tensor):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    d = (a - b).abs()
    return (d.max() / b.abs().max().clamp_min(1e-30)).item(), (d.norm() / b.norm().clamp_min(1e-30)).item()


def check_close(name, a, b, tol, max_factor=2.5):
    """Parity criterion: relative L2 error <= tol and max-norm error <= max_factor * tol.
    (One fp16 ulp at the largest-magnitude element is already 2^-10 ~ 1e-3 of max|b|, so the max-norm
    bound is 2.5 ulp-at-max; the L2 bound is the north_star figure.)"""
    emax, el2 = rel_err(a, b)
    print(f"{name}: rel err max-norm {emax:.2e}, L2 {el2:.2e} (tol {tol:.1e})")
    assert el2 < tol and emax < max_factor * tol, (name, emax, el2, tol)
    return emax, el2


SMALL = dict(vit_image=56, vit_patch=14, vit_hidden=128, vit_layers=2, vit_heads=2, vit_ffn=256,
             proj_d_model=256, gate_layers=2, gate_heads=2, gate_kv_heads=1, gate_head_dim=128, gate_ffn=512,
             llm_hidden=256, llm_layers=2, llm_heads=2, llm_kv_heads=1, llm_head_dim=128, llm_ffn=512,
             llm_vocab=1000, llm_max_ctx=512)


def engine_config(dtype, small=True, **over) -> EngineConfig:
    kw = dict(SMALL) if small else {}
    kw.update(over)
    return EngineConfig(dtype=dtype, **kw)


def make_weights(cfg: EngineConfig, seed=1234, vit=True, proj=True, gate=True, llm=False, device="cpu"):
    """state_dict (reference keys) in cfg.dtype for the enabled sub-models; ViT gets one extra layer
    (the reference computes and discards layer 24, clip_encoder.py:32 with select_layer=-2)."""
    sd = {}
    dt = cfg.dtype
    if vit and cfg.vit_layers > 0:
        sd.update(synth.make_vit_weights(seed, dt, device, hidden=cfg.vit_hidden, ffn=cfg.vit_ffn,
                                         layers=cfg.vit_layers + 1, heads=cfg.vit_heads,
                                         image_size=cfg.vit_image, patch=cfg.vit_patch))
    if proj and cfg.proj_d_model > 0:
        sd.update(synth.make_projector_weights(seed, dt, device, d_model=cfg.proj_d_model,
                                               mm_hidden=cfg.vit_hidden, d_state=cfg.proj_d_state,
                                               d_conv=cfg.proj_d_conv, expand=cfg.proj_expand))
    if gate and cfg.gate_layers > 0:
        sd.update(synth.make_mistral_weights(seed, synth.GATE_PREFIX, dt, device, hidden=cfg.proj_d_model,
                                             ffn=cfg.gate_ffn, layers=cfg.gate_layers, heads=cfg.gate_heads,
                                             kv_heads=cfg.gate_kv_heads, head_dim=cfg.gate_head_dim, vocab=2,
                                             with_embed=False, with_qk=False))
    if llm and cfg.llm_layers > 0:
        sd.update(synth.make_mistral_weights(seed, "", dt, device, hidden=cfg.llm_hidden, ffn=cfg.llm_ffn,
                                             layers=cfg.llm_layers, heads=cfg.llm_heads,
                                             kv_heads=cfg.llm_kv_heads, head_dim=cfg.llm_head_dim,
                                             vocab=cfg.llm_vocab))
    return sd


def oracle_configs(cfg: EngineConfig) -> R.StreamConfigs:
    return R.StreamConfigs(
        vit=R.VitConfig(image_size=cfg.vit_image, patch_size=cfg.vit_patch, hidden_size=cfg.vit_hidden,
                        num_layers=cfg.vit_layers + 1, num_heads=cfg.vit_heads, intermediate_size=cfg.vit_ffn,
                        layer_norm_eps=cfg.vit_eps, select_layer=-2),
        mamba=R.MambaCfg(d_model=cfg.proj_d_model, d_state=cfg.proj_d_state, d_conv=cfg.proj_d_conv,
                         expand=cfg.proj_expand, mm_hidden_size=cfg.vit_hidden, norm_eps=cfg.proj_eps),
        gate=R.gate_config(hidden_size=cfg.proj_d_model, num_layers=cfg.gate_layers, num_heads=cfg.gate_heads,
                           num_kv_heads=cfg.gate_kv_heads, head_dim=cfg.gate_head_dim,
                           intermediate_size=cfg.gate_ffn, rms_norm_eps=cfg.gate_eps),
        llm=R.MistralCfg(hidden_size=cfg.llm_hidden, num_layers=cfg.llm_layers, num_heads=cfg.llm_heads,
                         num_kv_heads=cfg.llm_kv_heads, head_dim=cfg.llm_head_dim,
                         intermediate_size=cfg.llm_ffn, vocab_size=cfg.llm_vocab, rms_norm_eps=cfg.llm_eps,
                         rope_theta=cfg.llm_rope_theta))


def f32(sd):
    """fp32 copies of model-dtype weights (exactly representable) for the oracle."""
    return {k: v.detach().float().cpu() for k, v in sd.items()}


def build_engine(cfg: EngineConfig, sd):
    from streammind_b200.engine import Engine
    eng = Engine(cfg)
    eng.load_state_dict(sd)
    eng.finalize()
    eng.reset_stream()
    return eng
