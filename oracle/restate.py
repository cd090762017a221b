"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of StreamMind's per-frame hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker or the CPU baseline.  The product path
(streammind_b200/*) never imports it.

Every function restates, in plain torch CPU ops, one reference function on the path
 and cites the file:line it follows.  Paths are relative to /root/reference
unless prefixed ``hf:`` = site-packages/transformers (5.5.0 in this image; the reference pins
4.44.2 -- CLIP/Mistral arithmetic is unchanged between the two).

Parity pin: the reference has NO golden vectors / tests of its own, so this
restatement is pinned against outputs of the reference's own code executed in the build container
under oracle/shims.py -- see oracle/make_golden.py (generator, committed) and tests/golden/*.npz
(fixtures, committed), checked by tests/test_oracle_golden.py.

Weights are addressed by the reference model's own ``state_dict()`` keys, e.g.
``model.vision_tower.vision_tower.vision_model.encoder.layers.0.self_attn.q_proj.weight``.
All functions compute in the dtype of the tensors they are given (fp32 for the goldens; parity
tests feed fp16/bf16-rounded weights upcast to fp32, i.e. "exact arithmetic on the same weights").
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# ---------------------------------------------------------------------------------------------
# Rounding emulation.  On the GPU the reference runs in fp16 / bf16: every tensor an nn.Module
# returns is materialised in the model dtype (cuBLAS / cuDNN / SDPA accumulate in fp32 and round once
# on output).  With ``emulate(dtype)`` active the restatement keeps fp32 storage but rounds to ``dtype``
# at exactly those points, i.e. it computes "the reference's arithmetic with exact accumulation".
# Feed it weights / inputs that are already representable in ``dtype`` (w.to(dtype).float()).
# With no emulation (default) everything is plain fp32 -- that mode is pinned by tests/golden/.
# ---------------------------------------------------------------------------------------------
_EMULATE: Optional[torch.dtype] = None


class emulate:
    def __init__(self, dtype: Optional[torch.dtype]):
        self.dtype = dtype

    def __enter__(self):
        global _EMULATE
        self.prev, _EMULATE = _EMULATE, self.dtype
        return self

    def __exit__(self, *a):
        global _EMULATE
        _EMULATE = self.prev


def _r(x: torch.Tensor) -> torch.Tensor:
    return x if _EMULATE is None else x.to(_EMULATE).to(x.dtype)


def _linear(x, w, b=None):
    return _r(F.linear(x, w, b))


def _ln(x, n, w, b, eps):
    return _r(F.layer_norm(x, (n,), w, b, eps))


def _leaky(x):
    return _r(F.leaky_relu(x))

VIT_PREFIX = "model.vision_tower.vision_tower.vision_model."
PROJ_PREFIX = "model.mm_projector."
GATE_PREFIX = "model.mm_projector.cls_net.cls_model."
LLM_PREFIX = ""          # "model.layers.N...", "model.embed_tokens.weight", "lm_head.weight"
VIDEO_TOKEN_INDEX = -201  # streammind/constants.py:29  MMODAL_TOKEN_INDEX["VIDEO"]


# --------------------------------------------------------------------------------------------
# configs
# --------------------------------------------------------------------------------------------
@dataclass
class VitConfig:
    """CLIP-ViT-L/14-336 (scripts/custom/finetune_stage1.sh:36); hf: CLIPVisionConfig."""
    image_size: int = 336
    patch_size: int = 14
    hidden_size: int = 1024
    num_layers: int = 24
    num_heads: int = 16
    intermediate_size: int = 4096
    layer_norm_eps: float = 1e-5
    select_layer: int = -2          # mm_vision_select_layer (finetune_stage1.sh), clip_encoder.py:32

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def layers_used(self) -> int:
        """hidden_states has num_layers+1 entries; entry [select_layer] needs this many layers."""
        idx = self.select_layer if self.select_layer >= 0 else self.num_layers + 1 + self.select_layer
        return idx


@dataclass
class MistralCfg:
    """hf: MistralConfig.  Defaults = Mistral-7B-Instruct-v0.2 values."""
    hidden_size: int = 4096
    num_layers: int = 32
    num_heads: int = 32
    num_kv_heads: int = 8
    head_dim: int = 128
    intermediate_size: int = 14336
    vocab_size: int = 32002
    rms_norm_eps: float = 1e-5
    rope_theta: float = 1e6


def gate_config(hidden_size: int = 4096, **kw) -> MistralCfg:
    """ClsNet: ``MistralConfig()`` defaults with vocab_size=2, num_hidden_layers=4
    (streammind/model/multimodal_projector/builder.py:373-378).  HF defaults: rms_norm_eps=1e-6,
    rope_theta=10000 (dead at L=1)."""
    base = dict(hidden_size=hidden_size, num_layers=4, num_heads=32, num_kv_heads=8, head_dim=128,
                intermediate_size=14336, vocab_size=2, rms_norm_eps=1e-6, rope_theta=10000.0)
    base.update(kw)
    return MistralCfg(**base)


@dataclass
class MambaCfg:
    """Mamba-1 hyper-parameters (streammind/model/mamba_ssm/modules/mamba_simple.py:31-58)."""
    d_model: int = 4096
    d_state: int = 16
    d_conv: int = 4
    expand: int = 2
    mm_hidden_size: int = 1024      # PreNet input = CLIP hidden (builder.py:393)
    norm_eps: float = 1e-5

    @property
    def d_inner(self) -> int:
        return self.expand * self.d_model

    @property
    def dt_rank(self) -> int:
        return math.ceil(self.d_model / 16)


# --------------------------------------------------------------------------------------------
# a1: CLIP vision tower
# --------------------------------------------------------------------------------------------
def quick_gelu(x):
    """hf: activations.py QuickGELUActivation: x * sigmoid(1.702 x)."""
    return _r(x * _r(torch.sigmoid(_r(1.702 * x))))


def clip_vision_tower(sd: SD, cfg: VitConfig, pixels: torch.Tensor) -> torch.Tensor:
    """CLIPVisionTower.forward + feature_select('patch')
    (streammind/model/multimodal_encoder/clip_encoder.py:31-53) over hf: modeling_clip.py
    CLIPVisionEmbeddings (:138-218), CLIPEncoderLayer (:354-385), CLIPVisionTransformer (:667-696).

    pixels [B,3,H,W] -> hidden_states[select_layer][:, 1:]  = [B, num_patches, hidden].
    Layers after ``layers_used`` and post_layernorm do not influence the result and are skipped.
    """
    p = VIT_PREFIX
    B = pixels.shape[0]
    w = sd[p + "embeddings.patch_embedding.weight"]
    x = _r(F.conv2d(pixels.to(w.dtype), w, bias=None, stride=cfg.patch_size))   # [B,C,gh,gw]
    x = x.flatten(2).transpose(1, 2)                                              # [B,N,C]
    cls = sd[p + "embeddings.class_embedding"].expand(B, 1, -1)
    x = _r(torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"].unsqueeze(0))
    x = _ln(x, cfg.hidden_size, sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"], cfg.layer_norm_eps)
    H, D = cfg.num_heads, cfg.hidden_size // cfg.num_heads
    scale = D ** -0.5
    for i in range(cfg.layers_used):
        lp = f"{p}encoder.layers.{i}."
        r = x
        h = _ln(x, cfg.hidden_size, sd[lp + "layer_norm1.weight"], sd[lp + "layer_norm1.bias"], cfg.layer_norm_eps)
        q = _linear(h, sd[lp + "self_attn.q_proj.weight"], sd[lp + "self_attn.q_proj.bias"])
        k = _linear(h, sd[lp + "self_attn.k_proj.weight"], sd[lp + "self_attn.k_proj.bias"])
        v = _linear(h, sd[lp + "self_attn.v_proj.weight"], sd[lp + "self_attn.v_proj.bias"])
        S = q.shape[1]
        q = q.view(B, S, H, D).transpose(1, 2)
        k = k.view(B, S, H, D).transpose(1, 2)
        v = v.view(B, S, H, D).transpose(1, 2)
        # fused SDPA: scores and softmax in fp32, probabilities rounded before P@V, fp32 accumulate
        att = _r(torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1, dtype=torch.float32).to(q.dtype))
        o = _r((att @ v).transpose(1, 2).reshape(B, S, H * D))
        x = _r(r + _linear(o, sd[lp + "self_attn.out_proj.weight"], sd[lp + "self_attn.out_proj.bias"]))
        r = x
        h = _ln(x, cfg.hidden_size, sd[lp + "layer_norm2.weight"], sd[lp + "layer_norm2.bias"], cfg.layer_norm_eps)
        h = quick_gelu(_linear(h, sd[lp + "mlp.fc1.weight"], sd[lp + "mlp.fc1.bias"]))
        x = _r(r + _linear(h, sd[lp + "mlp.fc2.weight"], sd[lp + "mlp.fc2.bias"]))
    return x[:, 1:]


# --------------------------------------------------------------------------------------------
# a4-a6: projector (PreNet -> VideoMamba(1 x Mamba-1 block) -> PostNet)
# --------------------------------------------------------------------------------------------
@dataclass
class MambaState:
    """Mamba.step state (mamba_simple.py:208-253): conv window [d_inner, d_conv] and ssm state
    [d_inner, d_state] fp32."""
    conv: torch.Tensor
    ssm: torch.Tensor

    @staticmethod
    def zeros(cfg: MambaCfg, dtype=torch.float32) -> "MambaState":
        return MambaState(torch.zeros(cfg.d_inner, cfg.d_conv, dtype=dtype),
                          torch.zeros(cfg.d_inner, cfg.d_state, dtype=torch.float32))


def pool_patches(feats: torch.Tensor) -> torch.Tensor:
    """``torch.mean(x, dim=2)`` over the patch axis
    (streammind/model/multimodal_projector/builder.py:405).  [..., P, C] -> [..., C]."""
    return _r(feats.mean(dim=-2))


def projector_sequence(sd: SD, cfg: MambaCfg, feats: torch.Tensor) -> torch.Tensor:
    """Video_Mamba_seq.forward core, full-sequence form exactly as the reference runs it every frame
    (builder.py:403-414) -> VideoMamba.forward (ssm.py:69-100) -> Block.forward
    (mamba_ssm/modules/block.py:51-55) -> Mamba.forward non-fused branch
    (mamba_ssm/modules/mamba_simple.py:135-143,161-206) -> selective_scan_ref
    (mamba_ssm/ops/selective_scan_interface.py:91-157).

    feats [1,T,P,C] -> x [1,T,d_model]."""
    assert feats.shape[0] == 1
    toks = []
    st = MambaState.zeros(cfg, dtype=feats.dtype)
    pooled = pool_patches(feats[0])
    for t in range(pooled.shape[0]):
        toks.append(projector_step(sd, cfg, pooled[t], st))
    return torch.stack(toks, 0).unsqueeze(0)


def projector_step(sd: SD, cfg: MambaCfg, pooled: torch.Tensor, st: MambaState) -> torch.Tensor:
    """One frame through PreNet (builder.py:161-170) -> Block.norm (block.py:52-53) ->
    Mamba.step (mamba_simple.py:208-253, pure-torch branches :215-221,:238-246) ->
    + residual -> norm_fn (ssm.py:83-84) -> PostNet (builder.py:172-181).  Updates ``st`` in place.

    Equivalent to row t of projector_sequence() because conv1d(padding=d_conv-1)[..., :L] is causal
    and selective_scan is a left-to-right recurrence (SURVEY.md section 8c equivalence (ii)).
    pooled [C] -> tok [d_model]."""
    p = PROJ_PREFIX
    mp = p + "mamba_model.ssms.0."
    h0 = _leaky(_linear(pooled, sd[p + "pre_net.fc3.weight"], sd[p + "pre_net.fc3.bias"]))
    resid = h0
    hn = _ln(h0, cfg.d_model, sd[mp + "norm.weight"], sd[mp + "norm.bias"], cfg.norm_eps)
    xz = _linear(hn, sd[mp + "mixer.in_proj.weight"])
    x, z = xz[: cfg.d_inner], xz[cfg.d_inner:]
    # depth-wise causal conv as a rolling window (mamba_simple.py:215-221); rounding follows the
    # full-sequence path the reference actually runs: act(conv1d(x)) (mamba_simple.py:168-169)
    st.conv.copy_(torch.roll(st.conv, shifts=-1, dims=-1))
    st.conv[:, -1] = x
    cw = sd[mp + "mixer.conv1d.weight"].reshape(cfg.d_inner, cfg.d_conv)
    x = _r(torch.sum(st.conv * cw, dim=-1) + sd[mp + "mixer.conv1d.bias"])
    x = _r(F.silu(x))
    x_db = _linear(x, sd[mp + "mixer.x_proj.weight"])
    dt, Bm, Cm = torch.split(x_db, [cfg.dt_rank, cfg.d_state, cfg.d_state], dim=-1)
    dt = _linear(dt, sd[mp + "mixer.dt_proj.weight"])
    A = -torch.exp(sd[mp + "mixer.A_log"].float())
    # selective scan, one step, fp32 internally (selective_scan_interface.py:104-152)
    dtf = F.softplus(dt.float() + sd[mp + "mixer.dt_proj.bias"].float())
    dA = torch.exp(dtf[:, None] * A)
    dBx = dtf[:, None] * Bm.float()[None, :] * x.float()[:, None]
    st.ssm.copy_(st.ssm * dA + dBx)
    y = (st.ssm * Cm.float()[None, :]).sum(-1) + sd[mp + "mixer.D"].float() * x.float()
    y = _r((y * F.silu(z.float())).to(pooled.dtype))
    out = _linear(y, sd[mp + "mixer.out_proj.weight"])
    hid = _ln(_r(out + resid), cfg.d_model, sd[p + "mamba_model.norm_fn.weight"],
              sd[p + "mamba_model.norm_fn.bias"], cfg.norm_eps)
    return _linear(_leaky(hid), sd[p + "post_net.fc3.weight"], sd[p + "post_net.fc3.bias"])


# --------------------------------------------------------------------------------------------
# a10: Mistral decoder (used by the gate at L=1 and by the LLM)
# --------------------------------------------------------------------------------------------
def rms_norm(x, w, eps):
    """hf: modeling_mistral.py MistralRMSNorm (:182-199): fp32 variance, cast back, times weight."""
    dt = x.dtype
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return _r(w * _r(xf.to(dt)))


def rope_cos_sin(cfg: MistralCfg, positions: torch.Tensor):
    """hf: MistralRotaryEmbedding: inv_freq = theta^(-2i/d); fp32 cos/sin of cat(freqs, freqs)."""
    d = cfg.head_dim
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    fr = positions.float()[:, None] * inv[None, :]
    emb = torch.cat([fr, fr], dim=-1)
    return _r(emb.cos()), _r(emb.sin())      # cast to the model dtype before use (hf MistralRotaryEmbedding)


def apply_rope(x, cos, sin):
    """hf: modeling_mistral.py rotate_half / apply_rotary_pos_emb (:51-81). x [H,L,D]."""
    d = x.shape[-1]
    rot = torch.cat([-x[..., d // 2:], x[..., : d // 2]], dim=-1)
    return _r(_r(x * cos.to(x.dtype)[None]) + _r(rot * sin.to(x.dtype)[None]))


@dataclass
class KVCache:
    """Per-layer K/V after RoPE, [kv_heads, ctx, head_dim] each; grows by torch.cat."""
    k: List[Optional[torch.Tensor]] = field(default_factory=list)
    v: List[Optional[torch.Tensor]] = field(default_factory=list)

    def length(self) -> int:
        return 0 if not self.k or self.k[0] is None else self.k[0].shape[1]

    def truncate(self, n: int) -> None:
        for i in range(len(self.k)):
            if self.k[i] is not None:
                self.k[i] = self.k[i][:, :n] if n > 0 else None
                self.v[i] = self.v[i][:, :n] if n > 0 else None


def mistral_forward(sd: SD, prefix: str, cfg: MistralCfg, embeds: torch.Tensor,
                    cache: Optional[KVCache] = None, all_logits: bool = False) -> torch.Tensor:
    """hf: MistralForCausalLM.forward (:402-472) -> MistralModel (:328-399) -> MistralDecoderLayer
    (:202-241) -> MistralAttention (:122-179, GQA via repeat_kv :84-96, causal, scale head_dim^-0.5,
    softmax in fp32) + MistralMLP (:35-47).  No sliding window (SURVEY.md section 8d pins
    sliding_window=None).

    embeds [L, hidden] are appended at positions cache.length()..+L.  Returns fp32 logits of the
    last position [vocab] (or all positions [L, vocab])."""
    L = embeds.shape[0]
    if cache is None:
        cache = KVCache()
    if not cache.k:
        cache.k = [None] * cfg.num_layers
        cache.v = [None] * cfg.num_layers
    pos0 = cache.length()
    positions = torch.arange(pos0, pos0 + L)
    cos, sin = rope_cos_sin(cfg, positions)
    Hq, Hk, D = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    h = embeds
    for i in range(cfg.num_layers):
        lp = f"{prefix}model.layers.{i}."
        a = rms_norm(h, sd[lp + "input_layernorm.weight"], cfg.rms_norm_eps)
        q = _linear(a, sd[lp + "self_attn.q_proj.weight"]).view(L, Hq, D).transpose(0, 1)
        k = _linear(a, sd[lp + "self_attn.k_proj.weight"]).view(L, Hk, D).transpose(0, 1)
        v = _linear(a, sd[lp + "self_attn.v_proj.weight"]).view(L, Hk, D).transpose(0, 1)
        q = apply_rope(q, cos, sin)
        k = apply_rope(k, cos, sin)
        if cache.k[i] is not None:
            k = torch.cat([cache.k[i], k], dim=1)
            v = torch.cat([cache.v[i], v], dim=1)
        cache.k[i], cache.v[i] = k, v
        rep = Hq // Hk
        kk = k.repeat_interleave(rep, dim=0)
        vv = v.repeat_interleave(rep, dim=0)
        s = (q @ kk.transpose(-1, -2)) * (D ** -0.5)                 # [Hq, L, ctx]
        ctx = kk.shape[1]
        causal = torch.arange(ctx)[None, :] > positions[:, None]
        s = s.masked_fill(causal[None], float("-inf"))
        pr = _r(torch.softmax(s, dim=-1, dtype=torch.float32).to(q.dtype))
        o = _r((pr @ vv).transpose(0, 1).reshape(L, Hq * D))
        h = _r(h + _linear(o, sd[lp + "self_attn.o_proj.weight"]))
        a = rms_norm(h, sd[lp + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        m = _r(_r(F.silu(_linear(a, sd[lp + "mlp.gate_proj.weight"]))) * _linear(a, sd[lp + "mlp.up_proj.weight"]))
        h = _r(h + _linear(m, sd[lp + "mlp.down_proj.weight"]))
    h = rms_norm(h, sd[prefix + "model.norm.weight"], cfg.rms_norm_eps)
    if not all_logits:
        h = h[-1]
    return _linear(h, sd[prefix + "lm_head.weight"]).float()      # logits leave lm_head in model dtype


# --------------------------------------------------------------------------------------------
# a7: the event gate
# --------------------------------------------------------------------------------------------
def gate_logits(sd: SD, cfg: MistralCfg, tok: torch.Tensor) -> torch.Tensor:
    """Video_Mamba_seq.forward ``cls_demo`` branch (builder.py:547-562): the gate sees ONLY the last
    frame token as a length-1 sequence -> ClsNet.forward (:381-385) -> MistralForCausalLM_cls
    (:291-367, logits.float() at :330).  The 3-D bool mask ``input_embed.ne(0)`` (:555) is a no-op
    for L=1.  tok [hidden] -> logits [2] fp32 (index 0 = silence, 1 = respond,
    videollama2_arch.py:944-948).  Restated through the full Mistral forward (q/k/RoPE included)."""
    return mistral_forward(sd, GATE_PREFIX, cfg, tok[None, :])


def gate_logits_degenerate(sd: SD, cfg: MistralCfg, tok: torch.Tensor) -> torch.Tensor:
    """Same function with the dead work removed: at L=1 softmax over one key is 1, so attention
    output = o_proj(expand_gqa(v_proj(rms(h)))).  q_proj/k_proj/RoPE never reach the output
    (SURVEY.md section 8c equivalence (i)).  This is the arithmetic the CUDA gate kernel performs."""
    pfx = GATE_PREFIX
    rep = cfg.num_heads // cfg.num_kv_heads
    h = tok
    for i in range(cfg.num_layers):
        lp = f"{pfx}model.layers.{i}."
        a = rms_norm(h, sd[lp + "input_layernorm.weight"], cfg.rms_norm_eps)
        v = _linear(a, sd[lp + "self_attn.v_proj.weight"]).view(cfg.num_kv_heads, cfg.head_dim)
        o = v.repeat_interleave(rep, dim=0).reshape(-1)
        h = _r(h + _linear(o, sd[lp + "self_attn.o_proj.weight"]))
        a = rms_norm(h, sd[lp + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        m = _r(_r(F.silu(_linear(a, sd[lp + "mlp.gate_proj.weight"]))) * _linear(a, sd[lp + "mlp.up_proj.weight"]))
        h = _r(h + _linear(m, sd[lp + "mlp.down_proj.weight"]))
    h = rms_norm(h, sd[pfx + "model.norm.weight"], cfg.rms_norm_eps)
    return _linear(h, sd[pfx + "lm_head.weight"]).float()


def gate_decision(logits: torch.Tensor) -> int:
    """softmax -> argmax(dim=0).item() (streammind/model/videollama2_arch.py:938-941)."""
    return int(torch.softmax(logits, dim=0).argmax(dim=0).item())


# --------------------------------------------------------------------------------------------
# a8: splice frame tokens into the prompt
# --------------------------------------------------------------------------------------------
def splice_prompt(sd: SD, input_ids: List[int], frame_tokens: torch.Tensor,
                  interval_id_list: List[int]) -> torch.Tensor:
    """prepare_inputs_labels_for_multimodal_score_stream_inference_demo, fire branch
    (streammind/model/videollama2_arch.py:948-984): the i-th ``<video>`` sentinel (-201) is replaced
    by frame tokens [start_i, end_i) with end_i = interval_id_list[i], start_i = previous end (0 for
    the first); text ids go through embed_tokens.  Returns inputs_embeds [L', hidden]."""
    emb = sd["model.embed_tokens.weight"]
    starts = [0] + interval_id_list[:-1]
    out, chunk, vi = [], [], 0
    for tid in input_ids:
        if tid == VIDEO_TOKEN_INDEX:
            if chunk:
                out.append(emb[torch.tensor(chunk)])
                chunk = []
            out.append(frame_tokens[starts[vi]: interval_id_list[vi]].to(emb.dtype))
            vi += 1
        else:
            chunk.append(tid)
    if chunk:
        out.append(emb[torch.tensor(chunk)])
    return torch.cat(out, dim=0)


# --------------------------------------------------------------------------------------------
# a9: the per-frame streaming call
# --------------------------------------------------------------------------------------------
@dataclass
class StreamConfigs:
    vit: VitConfig
    mamba: MambaCfg
    gate: MistralCfg
    llm: MistralCfg


class ReferenceSemanticsStream:
    """stream_generate_demo exactly as the reference executes it
    (streammind/model/language_model/videollama2_mistral.py:385-439): state on the object
    (``frame_feature``, ``interval_id_list``, :159-162), every call re-runs the projector over ALL
    frames (videollama2_arch.py:190-198) and on a fire re-prefills the WHOLE dialogue with
    past_key_values=None (:413,426-431), then greedy-decodes.  O(T) work per frame; small cases only.
    """

    def __init__(self, sd: SD, cfgs: StreamConfigs):
        self.sd, self.c = sd, cfgs
        self.frame_feature: Optional[torch.Tensor] = None    # [1,T,P,C]
        self.interval_id_list: List[int] = []

    def step(self, input_ids: List[int], frames: torch.Tensor, max_new_tokens: int,
             stop_ids: Tuple[int, ...] = (), force_pred: Optional[int] = None):
        feats = clip_vision_tower(self.sd, self.c.vit, frames).unsqueeze(0)
        if self.frame_feature is not None:
            feats = torch.cat([self.frame_feature, feats], dim=1)
        self.frame_feature = feats
        T = feats.shape[1]
        x = projector_sequence(self.sd, self.c.mamba, feats)
        logits = gate_logits(self.sd, self.c.gate, x[0, -1])
        pred = gate_decision(logits) if force_pred is None else force_pred
        if pred == 0:
            return None, pred, logits, x
        self.interval_id_list.append(T)
        embeds = splice_prompt(self.sd, input_ids, x[0], self.interval_id_list)
        cache = KVCache()
        out = greedy_decode(self.sd, self.c.llm, embeds, cache, max_new_tokens, stop_ids)
        return out, pred, logits, x


def expand_dialogue(input_ids: List[int], interval_id_list: List[int]) -> List[Tuple[str, int]]:
    """Flatten a prompt with ``<video>`` sentinels into the item sequence the LLM actually sees:
    ('t', token_id) for text, ('f', frame_index) for each frame token of the i-th span
    [interval_id_list[i-1], interval_id_list[i])  (videollama2_arch.py:949-950,963)."""
    starts = [0] + interval_id_list[:-1]
    seq, vi = [], 0
    for tid in input_ids:
        if tid == VIDEO_TOKEN_INDEX:
            seq.extend(('f', j) for j in range(starts[vi], interval_id_list[vi]))
            vi += 1
        else:
            seq.append(('t', int(tid)))
    return seq


def embed_items(sd: SD, items: List[Tuple[str, int]], frame_tokens: torch.Tensor) -> torch.Tensor:
    """embed_tokens for text items, projector tokens for frame items (videollama2_arch.py:967-981)."""
    emb = sd["model.embed_tokens.weight"]
    rows = [emb[i] if kind == 't' else frame_tokens[i].to(emb.dtype) for kind, i in items]
    return torch.stack(rows, 0)


class IncrementalStream:
    """Semantic twin with persistent state -- the algorithm the B200 path implements
    (SURVEY.md section 7 step 0): Mamba state carried across frames (one projector_step per new
    frame) and ONE KV cache carried across fires.  On a fire the new dialogue is compared with the
    item sequence the cache already holds; the cache is cut back to the longest common prefix and
    only the remainder is prefilled (so a prompt whose re-tokenised text differs from the generated
    ids, SURVEY.md section 8a row a12, is still handled exactly).  Mathematically equal to
    ReferenceSemanticsStream; differs only in floating-point summation order."""

    def __init__(self, sd: SD, cfgs: StreamConfigs):
        self.sd, self.c = sd, cfgs
        self.state = MambaState.zeros(cfgs.mamba, dtype=sd[PROJ_PREFIX + "pre_net.fc3.weight"].dtype)
        self.tokens: List[torch.Tensor] = []
        self.interval_id_list: List[int] = []
        self.cache = KVCache()
        self.cached_items: List[Tuple[str, int]] = []
        self.prefilled = 0               # bookkeeping for tests: tokens prefilled by the last fire

    def step(self, input_ids: List[int], frames: torch.Tensor, max_new_tokens: int,
             stop_ids: Tuple[int, ...] = (), force_pred: Optional[int] = None):
        feats = clip_vision_tower(self.sd, self.c.vit, frames)
        for f in range(feats.shape[0]):
            self.tokens.append(projector_step(self.sd, self.c.mamba, pool_patches(feats[f]), self.state))
        x = torch.stack(self.tokens, 0)
        logits = gate_logits_degenerate(self.sd, self.c.gate, x[-1])
        pred = gate_decision(logits) if force_pred is None else force_pred
        if pred == 0:
            return None, pred, logits, x.unsqueeze(0)
        self.interval_id_list.append(len(self.tokens))
        items = expand_dialogue(input_ids, self.interval_id_list)
        lcp = 0
        while lcp < min(len(items), len(self.cached_items)) and items[lcp] == self.cached_items[lcp]:
            lcp += 1
        lcp = min(lcp, len(items) - 1)                  # always prefill at least one position
        self.cache.truncate(lcp)
        self.prefilled = len(items) - lcp
        embeds = embed_items(self.sd, items[lcp:], x)
        out = greedy_decode(self.sd, self.c.llm, embeds, self.cache, max_new_tokens, stop_ids)
        self.cached_items = items + [('t', t) for t in out[:-1]]   # last token was never fed back
        return out, pred, logits, x.unsqueeze(0)


def greedy_decode(sd: SD, cfg: MistralCfg, embeds: torch.Tensor, cache: KVCache,
                  max_new_tokens: int, stop_ids: Tuple[int, ...] = (),
                  return_logits: bool = False):
    """hf GenerationMixin.generate(do_sample=False) as driven from videollama2_mistral.py:426-431:
    prefill ``embeds`` on top of ``cache``, then argmax on fp32 logits, feed embed_tokens(token)
    back, until max_new_tokens or a stop id was produced (KeywordsStoppingCriteria's id rule,
    streammind/mm_utils.py:631-636; the stop token is part of the output, as in HF)."""
    emb = sd["model.embed_tokens.weight"]
    logits = mistral_forward(sd, LLM_PREFIX, cfg, embeds, cache)
    out, all_logits = [], []
    for _ in range(max_new_tokens):
        tok = int(torch.argmax(logits).item())
        out.append(tok)
        all_logits.append(logits)
        if tok in stop_ids or len(out) == max_new_tokens:
            break
        logits = mistral_forward(sd, LLM_PREFIX, cfg, emb[tok][None, :], cache)
    return (out, all_logits) if return_logits else out


# --------------------------------------------------------------------------------------------
# 8f-4: cognition sampling of a frame-token segment before the LLM
# --------------------------------------------------------------------------------------------
def exponential_sampling(tokens: torch.Tensor, percentage: float = 0.6):
    """videollama2_arch.py:595-601 (the shipped variant: linearly spaced indices): returns (kept rows, indices)."""
    n = tokens.size(0)
    num = 1 if int(percentage * n) == 0 else int(percentage * n)
    idx = torch.linspace(0, n - 1, num).int().tolist()
    return tokens[idx], idx


def similarity_sampling(tokens: torch.Tensor, percentage: float = 0.6):
    """videollama2_arch.py:603-611: the max(int(p n), 1) rows most cosine-similar to the last row, original order kept.
    torch's cosine_similarity (ATen: normalise both operands by clamp_min(norm, eps), multiply, sum) with every tensor
    rounded to the emulated dtype; sums in float64 so that no summation order matters.  The reference's argsort is
    unstable, i.e. tie order is unspecified there; here ties go to the lower index."""
    x = tokens.double()
    last = x[-1]
    eps = 1e-8
    nx = _r(x.pow(2).sum(-1).sqrt().float()).clamp_min(eps).double()
    nl = nx[-1]
    a = _r((x / nx[:, None]).float()).double()
    b = _r((last / nl).float()).double()
    sim = _r(_r((a * b[None]).float()).double().sum(-1).float())
    k = max(int(percentage * x.shape[0]), 1)
    order = sorted(range(x.shape[0]), key=lambda i: (-float(sim[i]), i))
    idx = sorted(order[:k])
    return tokens[idx], idx
