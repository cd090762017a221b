"""TEST INFRASTRUCTURE ONLY -- dimensions of the committed golden fixtures plus re-exports of the
seeded synthetic weight/frame generators (streammind_b200/synth.py), so oracle code has one import."""
from __future__ import annotations

from dataclasses import dataclass

from streammind_b200.synth import (  # noqa: F401
    make_vit_weights, make_mistral_weights, make_projector_weights, make_projector_gate_weights,
    make_frames, make_prompt_ids, VIT_PREFIX, PROJ_PREFIX, GATE_PREFIX)


# ---------------------------------------------------------------------------------------------
# small dimensions used by the committed golden fixtures (tests/golden/tiny_*.npz)
# ---------------------------------------------------------------------------------------------
@dataclass
class TinyDims:
    image_size: int = 28
    patch_size: int = 14
    vit_hidden: int = 64
    vit_ffn: int = 128
    vit_layers: int = 3
    vit_heads: int = 4
    hidden: int = 64            # LLM hidden == projector d_model == gate hidden
    llm_ffn: int = 128
    llm_layers: int = 2
    llm_heads: int = 4
    llm_kv_heads: int = 2
    vocab: int = 96
    gate_ffn: int = 96
    gate_heads: int = 4
    gate_kv_heads: int = 2
    prompt_ids: tuple = (1, 17, 33, 5, 81, 44, -201, 9, 60)
    turn_suffix_ids: tuple = (2, 12, 70, -201, 31, 8)


def tiny_configs(t: TinyDims):
    """oracle.restate config objects for TinyDims."""
    from oracle import restate as R
    return R.StreamConfigs(
        vit=R.VitConfig(image_size=t.image_size, patch_size=t.patch_size, hidden_size=t.vit_hidden,
                        num_layers=t.vit_layers, num_heads=t.vit_heads, intermediate_size=t.vit_ffn),
        mamba=R.MambaCfg(d_model=t.hidden, mm_hidden_size=t.vit_hidden),
        gate=R.gate_config(hidden_size=t.hidden, num_heads=t.gate_heads, num_kv_heads=t.gate_kv_heads,
                           head_dim=t.hidden // t.gate_heads, intermediate_size=t.gate_ffn),
        llm=R.MistralCfg(hidden_size=t.hidden, num_layers=t.llm_layers, num_heads=t.llm_heads,
                         num_kv_heads=t.llm_kv_heads, head_dim=t.hidden // t.llm_heads,
                         intermediate_size=t.llm_ffn, vocab_size=t.vocab, rms_norm_eps=1e-5, rope_theta=1e6))
