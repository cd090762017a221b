"""The reference's own GPU arithmetic for the vision tower, for comparison only: hf ``CLIPVisionModel`` (what
``CLIPVisionTower`` wraps, /root/reference/streammind/model/multimodal_encoder/clip_encoder.py:21-53) built from the same
state_dict and run by PyTorch on the GPU in the model dtype.  Used by tests/test_reference_gpu_arm.py (second parity witness:
real fp16 / bf16 CUDA arithmetic instead of the emulated oracle) and by bench.py's ``reference_gpu_vit`` key.  Not on the
product path: nothing in Engine / model.py imports this module."""
from __future__ import annotations

from typing import Dict

import torch

from .synth import VIT_PREFIX


def build_hf_clip(sd: Dict[str, torch.Tensor], hidden: int, ffn: int, layers: int, heads: int, image: int, patch: int, eps: float, dtype,
                  device="cuda"):
    """``layers`` = layers present in ``sd`` (the reference model has one more than the 23 whose output is used)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    cfg = CLIPVisionConfig(hidden_size=hidden, intermediate_size=ffn, num_hidden_layers=layers, num_attention_heads=heads, image_size=image,
                           patch_size=patch, hidden_act="quick_gelu", layer_norm_eps=eps, projection_dim=hidden)
    m = CLIPVisionModel(cfg)
    own = {k[len(VIT_PREFIX) - len("vision_model."):]: v for k, v in sd.items() if k.startswith(VIT_PREFIX)}
    missing, unexpected = m.load_state_dict(own, strict=False)
    missing = [k for k in missing if "position_ids" not in k]
    if missing or unexpected:
        raise RuntimeError(f"hf CLIPVisionModel: missing {missing[:4]} unexpected {unexpected[:4]}")
    return m.to(device=device, dtype=dtype).eval()


@torch.no_grad()
def clip_features(model, pixels: torch.Tensor, select_layer: int = -2) -> torch.Tensor:
    """``feature_select`` of the reference (clip_encoder.py:31-39): hidden_states[select_layer] with the CLS token dropped."""
    out = model(pixels.to(device=next(model.parameters()).device, dtype=next(model.parameters()).dtype), output_hidden_states=True)
    return out.hidden_states[select_layer][:, 1:]
