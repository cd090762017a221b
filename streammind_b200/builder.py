"""``load_pretrained_model`` of the reference (/root/reference/streammind/model/builder.py:30-210), for the inference path:
same arguments, same ``(tokenizer, model, processor, context_len)`` result, but the model is a
:class:`streammind_b200.model.StreamMindB200ForCausalLM` whose weights live in the CUDA library.

A checkpoint directory holds what ``save_pretrained`` of the reference model writes: ``config.json`` (hf
``Videollama2MistralConfig`` fields), the state_dict as ``*.safetensors`` and / or ``pytorch_model*.bin`` shards (the
reference's own keys, vision tower and projector included when they were trained), optionally the tokenizer files.
When the vision tower is not part of the LLM checkpoint (``mm_vision_tower`` names a separate CLIP directory, as in the
reference: multimodal_encoder/builder.py:9) its ``CLIPVisionModel`` state_dict is read from there and re-keyed under
``model.vision_tower.vision_tower.``.  LoRA / 8-bit / 4-bit loading (builder.py:43-53,60-140) is out of scope and raises.
"""
from __future__ import annotations

import glob
import json
import os
from typing import Dict, Optional, Tuple

import torch

from .engine import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, EngineConfig
from .model import StreamMindB200ForCausalLM

VIT_PREFIX = "model.vision_tower.vision_tower."


class FrameProcessor:
    """What the reference gets from ``vision_tower.image_processor`` (a ``CLIPImageProcessor``), backed by
    ``sm_preprocess_frames``: ``preprocess(images, return_tensors='pt')['pixel_values']`` and the attributes
    ``mm_utils.process_video`` reads (``image_mean``, ``image_std``, ``crop_size``, ``size``)."""

    def __init__(self, engine, image_mean=OPENAI_CLIP_MEAN, image_std=OPENAI_CLIP_STD):
        self.engine = engine
        self.image_mean, self.image_std = list(image_mean), list(image_std)
        side = engine.cfg.vit_image
        self.crop_size = {"height": side, "width": side}
        self.size = {"shortest_edge": side}

    def preprocess(self, images, return_tensors="pt", **kw):
        import numpy as np
        if not isinstance(images, (list, tuple)):
            images = [images]
        frames = np.stack([np.asarray(im) for im in images])
        return {"pixel_values": self.engine.preprocess_frames(frames, self.image_mean, self.image_std)}

    __call__ = preprocess


def _read_state_dict(path: str) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    for f in sorted(glob.glob(os.path.join(path, "*.safetensors"))):
        from safetensors.torch import load_file
        sd.update(load_file(f))
    for f in sorted(glob.glob(os.path.join(path, "pytorch_model*.bin"))) + sorted(glob.glob(os.path.join(path, "mm_projector.bin"))):
        sd.update(torch.load(f, map_location="cpu", weights_only=True))
    return sd


def engine_config_from_hf(cfg: dict, dtype: torch.dtype, **over) -> EngineConfig:
    """hf ``MistralConfig`` / ``Videollama2MistralConfig`` fields -> EngineConfig (defaults = Mistral-7B + CLIP-L/14-336)."""
    heads = cfg.get("num_attention_heads", 32)
    kw = dict(
        dtype=dtype,
        llm_hidden=cfg.get("hidden_size", 4096), llm_layers=cfg.get("num_hidden_layers", 32), llm_heads=heads,
        llm_kv_heads=cfg.get("num_key_value_heads", 8), llm_head_dim=cfg.get("head_dim") or cfg.get("hidden_size", 4096) // heads,
        llm_ffn=cfg.get("intermediate_size", 14336), llm_vocab=cfg.get("vocab_size", 32002),
        llm_eps=cfg.get("rms_norm_eps", 1e-5), llm_rope_theta=cfg.get("rope_theta", 1e6),
        proj_d_model=cfg.get("hidden_size", 4096),
    )
    vit = cfg.get("vision_config") or {}
    if vit:
        layers = vit.get("num_hidden_layers", 24)
        sel = cfg.get("mm_vision_select_layer", -2)
        kw.update(vit_image=vit.get("image_size", 336), vit_patch=vit.get("patch_size", 14), vit_hidden=vit.get("hidden_size", 1024),
                  vit_heads=vit.get("num_attention_heads", 16), vit_ffn=vit.get("intermediate_size", 4096),
                  vit_layers=sel if sel >= 0 else layers + 1 + sel, vit_eps=vit.get("layer_norm_eps", 1e-5))
    for k in ("gate_layers", "gate_heads", "gate_kv_heads", "gate_head_dim", "gate_ffn", "llm_max_ctx", "max_frames", "n_streams", "use_graphs"):
        if k in cfg:
            kw[k] = cfg[k]
    kw.update(over)
    return EngineConfig(**kw)


def load_pretrained_model(model_path, model_base, model_name, load_8bit=False, load_4bit=False, device_map="auto", device="cuda",
                          use_flash_attn=False, **kwargs) -> Tuple[Optional[object], StreamMindB200ForCausalLM, FrameProcessor, int]:
    if load_8bit or load_4bit:
        raise NotImplementedError("bitsandbytes 8-bit / 4-bit loading is not part of the B200 path (fp16 / bf16 weights only)")
    if "lora" in model_name.lower() or model_base is not None:
        raise NotImplementedError("LoRA / base+delta checkpoints: merge them with the reference's own tooling first")
    with open(os.path.join(model_path, "config.json")) as f:
        hf = json.load(f)
    dtype = kwargs.pop("torch_dtype", None) or {"bfloat16": torch.bfloat16}.get(hf.get("torch_dtype"), torch.float16)   # the reference loads fp16 (:54)
    dev_index = torch.cuda.current_device() if device == "cuda" else (torch.device(device).index or 0)
    sd = _read_state_dict(model_path)
    tower = hf.get("mm_vision_tower")
    if tower and not any(k.startswith(VIT_PREFIX) for k in sd):
        tower_dir = tower if os.path.isdir(tower) else os.path.join(model_path, tower)
        with open(os.path.join(tower_dir, "config.json")) as f:
            tcfg = json.load(f)
        hf.setdefault("vision_config", tcfg.get("vision_config", tcfg))
        sd.update({VIT_PREFIX + k: v for k, v in _read_state_dict(tower_dir).items()})
    cfg = engine_config_from_hf(hf, dtype, **{k: kwargs.pop(k) for k in list(kwargs) if k in EngineConfig.__dataclass_fields__})
    model = StreamMindB200ForCausalLM(cfg, sd, device=dev_index)
    tokenizer = None
    if any(os.path.exists(os.path.join(model_path, n)) for n in ("tokenizer.model", "tokenizer.json", "tokenizer_config.json")):
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(model_path, use_fast=False)
    processor = FrameProcessor(model.engine)
    context_len = hf.get("max_sequence_length", 2048)          # builder.py:205-208
    return tokenizer, model, processor, context_len
