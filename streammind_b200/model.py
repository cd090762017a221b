"""Host-side mirror of the reference's model API for the per-frame path (hooks
B1 and B2): same class roles, method names, argument meaning and error behaviour as

  * ``CLIPVisionTower``                      /root/reference/streammind/model/multimodal_encoder/clip_encoder.py:7-84
  * ``Video_Mamba_seq`` (``mm_projector``)   /root/reference/streammind/model/multimodal_projector/builder.py:390-564
  * ``Videollama2MistralForCausalLM``        /root/reference/streammind/model/language_model/videollama2_mistral.py:146-449
  * ``infer`` of the streaming demo          /root/reference/streammind/eval/video_score_stream_demo.py:66-125

but every tensor op runs in the CUDA library behind ``Engine`` (no PyTorch arithmetic on the path).
What differs by design (DESIGN.md "incremental state"): the projector advances a persistent Mamba
state by the NEW frames only, and the LLM keeps ONE KV cache across fires, re-using the longest common
prefix of the dialogue (the reference re-runs the projector over all T frames and re-prefills the
whole dialogue with past_key_values=None on every fire: videollama2_arch.py:190-198,
videollama2_mistral.py:413,426-431).
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .constants import MMODAL_TOKEN_INDEX
from .engine import Engine, EngineConfig

VIDEO_TOKEN_INDEX = MMODAL_TOKEN_INDEX["VIDEO"]
Item = Tuple[str, int]        # ('t', token id) | ('f', frame index)


# --------------------------------------------------------------------------------------------------
# pure host logic (unit-tested on CPU)
# --------------------------------------------------------------------------------------------------
def expand_dialogue(input_ids: Sequence[int], interval_id_list: Sequence[int]) -> List[Item]:
    """The sequence the LLM sees for a prompt with ``<video>`` sentinels: the i-th sentinel stands for
    frame tokens [interval_id_list[i-1], interval_id_list[i]) (0 for i = 0); videollama2_arch.py:949-981."""
    n_sent = sum(1 for t in input_ids if t == VIDEO_TOKEN_INDEX)
    if n_sent > len(interval_id_list):
        raise ValueError(f"prompt has {n_sent} <video> sentinels but only {len(interval_id_list)} intervals were recorded")
    starts = [0] + list(interval_id_list[:-1])
    seq: List[Item] = []
    vi = 0
    for tid in input_ids:
        if tid == VIDEO_TOKEN_INDEX:
            seq.extend(("f", j) for j in range(starts[vi], interval_id_list[vi]))
            vi += 1
        else:
            seq.append(("t", int(tid)))
    return seq


class DialogueCache:
    """Which items the device KV cache currently holds, and how much of a new dialogue can be re-used."""

    def __init__(self):
        self.items: List[Item] = []

    def reset(self):
        self.items = []

    def plan(self, new_items: Sequence[Item]) -> int:
        """Length of the longest common prefix (always leaves at least one item to prefill, because the
        logits of the last position are needed to start decoding)."""
        n = 0
        lim = min(len(new_items), len(self.items))
        while n < lim and new_items[n] == self.items[n]:
            n += 1
        return min(n, len(new_items) - 1)

    def commit(self, new_items: Sequence[Item], generated: Sequence[int]):
        """After prefill + greedy decode: the cache holds the dialogue and every generated token except
        the last one (never fed back, as in HF generate)."""
        self.items = list(new_items) + [("t", int(t)) for t in generated[:-1]]


# --------------------------------------------------------------------------------------------------
# B2: component API
# --------------------------------------------------------------------------------------------------
class CLIPVisionTower:
    """``forward(images[B,3,H,W] | list) -> [B, num_patches, hidden]`` (patch features of
    hidden_states[select_layer], CLS dropped), output cast back to the input dtype."""

    def __init__(self, engine: Engine):
        self.engine = engine
        self.is_loaded = True
        self.select_layer = -2
        self.select_feature = "patch"

    @torch.no_grad()
    def forward(self, images):
        if isinstance(images, list):
            return [self.forward(im.unsqueeze(0)) for im in images]
        e = self.engine
        x = images.to(device=e.device, dtype=e.cfg.dtype)
        outs = []
        for i in range(0, x.shape[0], e.cfg.max_frames):
            outs.append(e.vit_encode(x[i:i + e.cfg.max_frames])[0])
        return torch.cat(outs, 0).to(images.dtype)

    __call__ = forward

    @property
    def dtype(self):
        return self.engine.cfg.dtype

    @property
    def device(self):
        return self.engine.device

    @property
    def hidden_size(self):
        return self.engine.cfg.vit_hidden

    @property
    def num_patches(self):
        return self.engine.cfg.num_patches

    @property
    def num_patches_per_side(self):
        return self.engine.cfg.vit_image // self.engine.cfg.vit_patch

    @property
    def dummy_feature(self):
        return torch.zeros(1, self.hidden_size, device=self.device, dtype=self.dtype)


class VideoMambaSeq:
    """``mm_projector(frames_features[1,T,P,C], cls_demo=True, frames_features_shape=...) ->
    (x[1,T,d_model], logits[2])`` and ``-> x`` when no flag is set (builder.py:403,562,564).

    The reference passes ALL T frames' features on every call; this object remembers how many it has
    already consumed and advances the Mamba state by the remainder only."""

    def __init__(self, engine: Engine):
        self.engine = engine
        self.tokens: Optional[torch.Tensor] = None      # [T, d_model] on device
        self.frames_seen = 0

    def reset(self):
        self.tokens, self.frames_seen = None, 0

    @torch.no_grad()
    def forward(self, x, cls_inference=False, cls_training=False, cls_demo=False, frames_features_shape=None,
                prompt_time_input_ids=None, prompt_time_lable=None):
        if cls_inference or cls_training:
            raise NotImplementedError("only the streaming (cls_demo) and plain projector calls are on the hot path")
        if x.dim() != 4 or x.shape[0] != 1:
            raise ValueError(f"expected frames_features of shape [1, T, P, C], got {tuple(x.shape)}")
        T = x.shape[1]
        if T < self.frames_seen:
            raise RuntimeError(f"feature history shrank ({T} < {self.frames_seen}); call reset() to start a new stream")
        e = self.engine
        if T > self.frames_seen:
            new = x[0, self.frames_seen:].to(device=e.device, dtype=e.cfg.dtype)
            toks = e.projector_step(e.pool_features(new))
            self.tokens = toks if self.tokens is None else torch.cat([self.tokens, toks], 0)
            self.frames_seen = T
        out = self.tokens.unsqueeze(0)
        if cls_demo:
            return out, e.gate_score(self.tokens[-1])
        return out

    __call__ = forward


# --------------------------------------------------------------------------------------------------
# B1: model API
# --------------------------------------------------------------------------------------------------
@dataclass
class KVToken:
    """Stand-in for hf ``past_key_values``: the cache itself stays on the device inside the handle."""
    owner: object
    length: int


@dataclass
class CausalLMOutput:
    """The fields of hf ``CausalLMOutputWithPast`` that generation consumes."""
    logits: torch.Tensor
    past_key_values: Optional[KVToken] = None
    loss: Optional[torch.Tensor] = None


class StreamMindB200ForCausalLM:
    """Drop-in for the calls the streaming demo / serve worker make on ``Videollama2MistralForCausalLM``."""

    def __init__(self, cfg: EngineConfig, state_dict: Optional[Dict[str, torch.Tensor]] = None, device: int = 0,
                 keep_frame_features: bool = False, engine: Optional[Engine] = None, stream_id: int = 0,
                 sample_per: float = 0.5, sample_type: str = "ssss"):
        self.config = cfg
        # cognition sampling of every <video> span in forward() (videollama2_mistral.py:166-167,204-205; arch.py:676-681):
        # "log" / "similarity" thin the span on the device, anything else (the reference's default "ssss") keeps it whole
        self.sample_per, self.sample_type = sample_per, sample_type
        self.engine = engine if engine is not None else Engine(cfg, device=device)     # several stream objects may share one engine
        if state_dict is not None:
            self.load_state_dict(state_dict)
        self.vision_tower = CLIPVisionTower(self.engine)
        self.mm_projector = VideoMambaSeq(self.engine)
        self.keep_frame_features = keep_frame_features
        # per-stream state, same attribute names as the reference (videollama2_mistral.py:159-162)
        self.frame_feature: Optional[torch.Tensor] = None
        self.interval_id_list: List[int] = []
        self._tokens: List[torch.Tensor] = []                # projector tokens, chunks of [t, d_model] in frame order
        self._num_frames = 0
        self._dialogue = DialogueCache()
        self.last_prefill_len = 0
        self._inflight = deque()                             # frames submitted ahead through the pipelined path (prefetch_frames)
        self.stream_id = stream_id                           # stream slot of the engine this object drives (multi-stream handles)

    # ---- loading -------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        self.engine.load_state_dict(sd)
        self.engine.finalize()
        for s in range(self.config.n_streams):
            self.engine.select_stream(s)
            self.engine.reset_stream()
        self.engine.select_stream(0)

    def get_vision_tower(self):
        return self.vision_tower

    def get_model(self):
        return self

    @property
    def device(self):
        return self.engine.device

    @property
    def dtype(self):
        return self.config.dtype

    def eval(self):
        return self

    def reset_stream(self):
        """Start a new video (the reference never resets; one stream per model instance)."""
        self._drain_inflight()
        self.engine.select_stream(self.stream_id)
        self.engine.reset_stream()
        self.mm_projector.reset()
        self.frame_feature, self.interval_id_list = None, []
        self._tokens, self._num_frames = [], 0
        self._dialogue.reset()

    def _token_rows(self, idx: Sequence[int]) -> torch.Tensor:
        toks = self._tokens[0] if len(self._tokens) == 1 else torch.cat(self._tokens, 0)
        self._tokens = [toks]
        return toks[torch.tensor(list(idx), device=toks.device)]

    def _pixels(self, frames):
        """Raw uint8 RGB frames [t, H, W, 3] (what the video decoder yields) are preprocessed on the device -- expand2square + CLIP
        preprocess of the reference's process_video (mm_utils.py:446-464), bit for bit -- instead of by PIL on the host; anything else
        is taken as already normalised pixels [t, 3, H, W]."""
        if isinstance(frames, torch.Tensor) and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3:
            return self.engine.preprocess_frames(frames)
        return frames

    # ---- per-frame path ------------------------------------------------------------------------
    def _encode_frames(self, frames: torch.Tensor):
        """encode_images_or_videos_score_cls_inference_allframe_demo (videollama2_arch.py:173-203),
        incremental: returns gate logits [2] fp32 (host) of the last frame."""
        e = self.engine
        frames = self._pixels(frames)
        if frames.dim() != 4:
            raise ValueError(f"expected frames [t, 3, H, W], got {tuple(frames.shape)}")
        if frames.dtype != e.cfg.dtype:
            frames = frames.to(e.cfg.dtype)
        last = None
        for i in range(0, frames.shape[0], e.cfg.max_frames):
            chunk = frames[i:i + e.cfg.max_frames]
            if not chunk.is_cuda and not chunk.is_pinned():
                chunk = chunk.to(e.device)
            feats, toks, _, lg_host = e.frame_step(chunk, want_feats=self.keep_frame_features)
            self._tokens.append(toks)
            if self.keep_frame_features:
                f = feats.unsqueeze(0)
                self.frame_feature = f if self.frame_feature is None else torch.cat([self.frame_feature, f], 1)
            torch.cuda.current_stream().synchronize()
            last = lg_host[chunk.shape[0] - 1].clone()
        self._num_frames += frames.shape[0]
        return last

    # ---- pipelined frames: encode ahead of the decisions (and under a running LLM decode) ----------------------------
    def prefetch_frames(self, frames: torch.Tensor):
        """Submit frames [t, 3, H, W] that LATER ``stream_generate_demo`` calls will be given, in this order, through the
        pipelined path (sm_frame_submit: towers of 8 consecutive frames run as one chunk, projector / gate batched, up to
        16 frames in flight on the library's own streams).  Nothing waits here.  The vision tower, projector and gate do
        not depend on the LLM, so frames submitted before a fire are encoded WHILE its prefill + greedy decode run on the
        caller's stream (north_star: the next frames' ViT encode overlaps the decode / KV appends); the reference does
        the two strictly one after the other (videollama2_mistral.py:410-431).  At most 15 frames may be ahead."""
        e = self.engine
        if e.cfg.max_frames != 1:
            raise RuntimeError("prefetch_frames needs a streaming engine (max_frames == 1)")
        frames = self._pixels(frames)
        if frames.dim() != 4:
            raise ValueError(f"expected frames [t, 3, H, W], got {tuple(frames.shape)}")
        if len(self._inflight) + frames.shape[0] > 15:
            raise RuntimeError("at most 15 frames may be submitted ahead of their stream_generate_demo calls")
        if frames.dtype != e.cfg.dtype:
            frames = frames.to(e.cfg.dtype)
        e.select_stream(self.stream_id)
        for i in range(frames.shape[0]):
            f = frames[i:i + 1]
            if not f.is_cuda and not f.is_pinned():
                f = f.to(e.device)
            tk, feats, toks, _, lg_host = e.frame_submit(f, want_feats=self.keep_frame_features, want_device_outputs=True)
            self._inflight.append((tk, feats, toks, lg_host))

    def _pop_prefetched(self):
        tk, feats, toks, lg_host = self._inflight.popleft()
        self.engine.frame_wait(tk, block=True, on_stream=True)       # decision on the host, tokens ordered on the current stream
        self._tokens.append(toks)
        if self.keep_frame_features:
            f = feats.unsqueeze(0)
            self.frame_feature = f if self.frame_feature is None else torch.cat([self.frame_feature, f], 1)
        self._num_frames += 1
        return lg_host[0].clone()

    def _drain_inflight(self):
        while self._inflight:
            self._pop_prefetched()

    @torch.no_grad()
    def stream_generate_demo(self, inputs: Optional[torch.Tensor] = None, images_or_videos: Optional[torch.Tensor] = None,
                             modal_list=None, **kwargs):
        """-> (text | None, pred).  Same keyword arguments as the reference (videollama2_mistral.py:385-439)."""
        kwargs.pop("position_ids", None)
        kwargs.pop("attention_mask", None)
        kwargs.pop("score_video", None)
        tokenizer = kwargs.pop("tokenizer", None)
        force_pred = kwargs.pop("force_pred", None)         # the authors' "# pred = 1" switch (arch.py:943)
        if "inputs_embeds" in kwargs:
            raise NotImplementedError("`inputs_embeds` is not supported")
        if kwargs.get("do_sample", False):
            raise NotImplementedError("only greedy decoding (do_sample=False) is implemented on the device")
        self.engine.select_stream(self.stream_id)
        if self._inflight:
            if images_or_videos is not None and images_or_videos.shape[0] != 1:
                raise ValueError("with prefetched frames every stream_generate_demo call consumes exactly one frame")
            logits = self._pop_prefetched()
        else:
            logits = self._encode_frames(images_or_videos)
        self.last_gate_logits = logits
        pred = int(torch.softmax(logits, dim=0).argmax(dim=0).item()) if force_pred is None else int(force_pred)
        if pred == 0:
            return None, pred
        self.interval_id_list.append(self._num_frames)
        ids = inputs[0].tolist() if isinstance(inputs, torch.Tensor) else list(inputs[0])
        out_ids = self._generate_from_dialogue(ids, kwargs)
        if tokenizer is None:
            return out_ids, pred
        text = tokenizer.batch_decode(torch.tensor([out_ids]), skip_special_tokens=True)[0].strip()
        return text, pred

    HOST_CHECK_CHUNK = 16      # tokens decoded per device call while a stopping criterion has to be evaluated on the host

    def _stop_ids(self, kwargs) -> List[int]:
        """Single-token stop ids run inside the device loop; criteria that need the host (multi-token keywords, any other
        callable) are returned by ``_host_criteria`` and evaluated between device calls."""
        stops: List[int] = []
        for sc in kwargs.get("stopping_criteria", None) or []:
            stops.extend(getattr(sc, "single_token_ids", []))
        eos = kwargs.get("eos_token_id", kwargs.get("pad_token_id", None))   # the demo passes pad_token_id=eos
        if eos is not None and eos not in stops:
            stops.append(int(eos))
        return stops

    @staticmethod
    def _host_criteria(kwargs) -> list:
        return [sc for sc in (kwargs.get("stopping_criteria", None) or [])
                if getattr(sc, "needs_host_check", not hasattr(sc, "single_token_ids"))]

    def _generate_from_dialogue(self, ids: Sequence[int], kwargs) -> List[int]:
        items = self._prefill_dialogue(ids)
        max_new, stops, host = int(kwargs.get("max_new_tokens", 1024)), self._stop_ids(kwargs), self._host_criteria(kwargs)
        out = self.engine.llm_decode(max_new, stops) if not host else self._decode_with_host_check(max_new, stops, host)
        self._dialogue.commit(items, out)
        return out

    def _decode_with_host_check(self, max_new: int, stops: List[int], host: list) -> List[int]:
        """Greedy decode with stopping criteria that only the host can evaluate (KeywordsStoppingCriteria with multi-token
        keywords, mm_utils.py:616-647): the device decodes HOST_CHECK_CHUNK tokens per call, the host applies the criteria token
        by token exactly as hf generate() does after every step -- on the NEW ids only, because the reference generates from
        ``inputs_embeds`` (videollama2_mistral.py:426-431) -- and on a hit the tokens decoded past it are dropped and the cache
        is rewound.  Between chunks the last token (never fed back by a decode call) goes through a one-position prefill,
        whose logits start the next chunk."""
        e = self.engine
        out: List[int] = []
        while len(out) < max_new:
            kv0 = e.kv_len
            part = e.llm_decode(min(self.HOST_CHECK_CHUNK, max_new - len(out)), stops)
            for j, t in enumerate(part):
                out.append(int(t))
                if int(t) in stops or any(bool(sc(torch.tensor([out]), None)) for sc in host):
                    e.kv_set_len(kv0 + j)          # the cache holds everything before the last kept token
                    return out
            if len(out) >= max_new or not part:
                break
            if e.kv_len + 1 >= e.cfg.llm_max_ctx:      # a full cache returns what fits
                break
            e.llm_prefill(e.embed_tokens(torch.tensor([out[-1]], dtype=torch.int32)))
        return out

    def _prefill_dialogue(self, ids: Sequence[int]) -> List[Item]:
        """Bring this stream's KV cache up to the end of the dialogue ``ids`` (longest-common-prefix re-use); returns the
        item sequence for ``DialogueCache.commit`` after decoding."""
        e = self.engine
        e.select_stream(self.stream_id)
        items = expand_dialogue(ids, self.interval_id_list)
        if len(items) > e.cfg.llm_max_ctx:
            # checked BEFORE the cache is touched, so a too-long dialogue costs nothing and the stream state stays valid.
            # There is no silent truncation policy: size EngineConfig.llm_max_ctx for the longest dialogue (the reference
            # relies on Mistral's 32k positions); decoding itself never fails on a full cache, it returns what fits.
            raise RuntimeError(f"dialogue of {len(items)} positions exceeds llm_max_ctx = {e.cfg.llm_max_ctx}")
        keep = self._dialogue.plan(items)
        if keep > e.kv_len:
            keep = e.kv_len
        e.kv_set_len(keep)
        todo = items[keep:]
        self.last_prefill_len = len(todo)
        text_pos = [i for i, (k, _) in enumerate(todo) if k == "t"]
        frame_pos = [i for i, (k, _) in enumerate(todo) if k == "f"]
        emb = torch.empty(len(todo), e.cfg.llm_hidden, dtype=e.cfg.dtype, device=e.device)
        if text_pos:
            tid = torch.tensor([todo[i][1] for i in text_pos], dtype=torch.int32)
            emb[torch.tensor(text_pos, device=e.device)] = e.embed_tokens(tid)
        if frame_pos:
            emb[torch.tensor(frame_pos, device=e.device)] = self._token_rows([todo[i][1] for i in frame_pos])
        e.llm_prefill(emb)
        return items

    @torch.no_grad()
    def generate(self, inputs: Optional[torch.Tensor] = None, images_or_videos: Optional[torch.Tensor] = None,
                 modal_list=None, **kwargs) -> torch.Tensor:
        """Offline call (videollama2_mistral.py:262-315): all frames at once, one ``<video>`` span, greedy
        decode; returns new token ids [1, n]."""
        if kwargs.get("do_sample", False):
            raise NotImplementedError("only greedy decoding (do_sample=False) is implemented on the device")
        # The reference's generate() re-encodes everything it is given and leaves the streaming attributes alone.  Here the
        # stream state lives in the engine, so an offline call runs on a scratch stream slot when the handle has one
        # (EngineConfig.n_streams > 1: the last slot) and the live stream is untouched; a single-slot handle has to reset
        # its only stream -- a live stream on it is lost (documented difference).
        saved = None
        if self.config.n_streams > 1 and self.stream_id != self.config.n_streams - 1:
            self._drain_inflight()
            saved = (self.stream_id, self.frame_feature, self.interval_id_list, self._tokens, self._num_frames, self._dialogue, self.mm_projector)
            self.stream_id = self.config.n_streams - 1
            self._dialogue, self.mm_projector = DialogueCache(), VideoMambaSeq(self.engine)
        try:
            self.reset_stream()
            if images_or_videos is not None:
                self._encode_frames(images_or_videos)
                self.interval_id_list = [self._num_frames]
            ids = inputs[0].tolist()
            out = torch.tensor([self._generate_from_dialogue(ids, kwargs)], dtype=torch.long)
        finally:
            if saved is not None:
                (self.stream_id, self.frame_feature, self.interval_id_list, self._tokens, self._num_frames, self._dialogue, self.mm_projector) = saved
                self.engine.select_stream(self.stream_id)
        return out

    @torch.no_grad()
    def forward(self, input_ids: Optional[torch.Tensor] = None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds: Optional[torch.Tensor] = None, labels=None, use_cache: Optional[bool] = None,
                output_attentions=None, output_hidden_states=None, images=None, return_dict=None, **kwargs):
        """``Videollama2MistralForCausalLM.forward`` for inference (videollama2_mistral.py:173-259 -> hf
        MistralForCausalLM.forward): appends ``inputs_embeds [1, L, hidden]`` -- or ``input_ids [1, L]`` whose ``<video>``
        sentinels are replaced by the projector tokens of ``images`` (one tensor of frames [t, 3, H, W] per sentinel) -- to
        the KV cache and returns the fp32 logits of the LAST position as ``CausalLMOutput.logits [1, 1, vocab]`` (what
        ``generate`` consumes) with ``past_key_values`` = an opaque token of the handle's cache.  Passing that token back
        continues the sequence; ``past_key_values=None`` starts a new one (hf semantics).  Training arguments
        (``labels``, ``timestamp`` ...) are outside the inference path and raise."""
        if labels is not None or kwargs.get("timestamp") is not None:
            raise NotImplementedError("the training / scoring forward is not on the inference hot path")
        e = self.engine
        e.select_stream(self.stream_id)
        if not isinstance(past_key_values, KVToken) or past_key_values.owner is not self or past_key_values.length != e.kv_len:
            if past_key_values is not None and not isinstance(past_key_values, KVToken):
                raise TypeError("past_key_values must be the token returned by an earlier forward() of this model (the KV cache lives on the device)")
            e.kv_set_len(0)
            self._dialogue.reset()
        if inputs_embeds is None:
            if input_ids is None:
                raise ValueError("forward() needs input_ids or inputs_embeds")
            ids = input_ids[0].tolist()
            frames = [] if images is None else (list(images) if isinstance(images, (list, tuple)) else [images])
            frames = [f[0] if isinstance(f, (list, tuple)) else f for f in frames]       # the reference passes (tensor, modal) pairs
            n_sent = sum(1 for t in ids if t == VIDEO_TOKEN_INDEX)
            if n_sent != len(frames):
                raise ValueError(f"{n_sent} <video> sentinels but {len(frames)} frame tensors")
            rows, chunk = [], []
            for t in ids:
                if t == VIDEO_TOKEN_INDEX:
                    if chunk:
                        rows.append(e.embed_tokens(torch.tensor(chunk)))
                        chunk = []
                    fr = frames.pop(0).to(device=e.device, dtype=e.cfg.dtype)
                    span = []
                    for i in range(0, fr.shape[0], e.cfg.max_frames):
                        _, pooled = e.vit_encode(fr[i:i + e.cfg.max_frames], want_feats=False)
                        span.append(e.projector_step(pooled))
                    span = span[0] if len(span) == 1 else torch.cat(span, 0)
                    if self.sample_type in ("log", "similarity"):
                        span, _ = e.cognition_sample(span.contiguous(), self.sample_per, self.sample_type)
                    rows.append(span)
                else:
                    chunk.append(int(t))
            if chunk:
                rows.append(e.embed_tokens(torch.tensor(chunk)))
            emb = torch.cat(rows, 0)
        else:
            emb = inputs_embeds[0].to(device=e.device, dtype=e.cfg.dtype)
        if e.kv_len + emb.shape[0] > e.cfg.llm_max_ctx:
            raise RuntimeError(f"{e.kv_len} + {emb.shape[0]} positions exceed llm_max_ctx = {e.cfg.llm_max_ctx}")
        logits = e.llm_prefill(emb.contiguous(), want_logits=True)
        return CausalLMOutput(logits=logits.view(1, 1, -1), past_key_values=KVToken(self, e.kv_len))

    __call__ = forward


class MultiStreamSession:
    """B video streams on ONE engine / GPU.  Every stream keeps the per-stream state the reference stores on
    its model object (frame tokens, ``interval_id_list``, dialogue / KV cache, Mamba state: videollama2_mistral.py:159-162) in
    its own slot of the handle; one call advances all of them by one frame:

      * one vision-tower batch for the B frames, every pass over the projector / gate weights shared (sm_frame_step_multi);
      * the streams whose gate fired are prefilled one by one (their dialogue suffixes differ in length) and then decoded
        TOGETHER: each pass over the 14.2 GB of LLM weights yields one token per firing stream (sm_llm_decode_multi).

    Results are those of B independent ``StreamMindB200ForCausalLM`` runs (same kernels, per-stream arithmetic independent
    of the batch: tests/test_multi_stream_gpu.py)."""

    def __init__(self, cfg: EngineConfig, state_dict: Optional[Dict[str, torch.Tensor]], n_streams: int, device: int = 0):
        import dataclasses
        cfg = dataclasses.replace(cfg, n_streams=n_streams, max_frames=max(cfg.max_frames, n_streams))
        self.config = cfg
        self.engine = Engine(cfg, device=device)
        self.streams = [StreamMindB200ForCausalLM(cfg, None, device=device, engine=self.engine, stream_id=s) for s in range(n_streams)]
        if state_dict is not None:
            self.streams[0].load_state_dict(state_dict)

    def reset(self):
        for m in self.streams:
            m.reset_stream()

    @torch.no_grad()
    def stream_generate_demo_multi(self, inputs: Sequence[Sequence[int]], frames: torch.Tensor, force_pred=None, **kwargs):
        """inputs[s]: prompt ids of stream s (with ``<video>`` sentinels); frames [B, 3, H, W]: the next frame of every stream.
        -> list of (new ids | None, pred) per stream."""
        e, B = self.engine, len(self.streams)
        if frames.shape[0] != B or len(inputs) != B:
            raise ValueError(f"expected one frame and one prompt per stream ({B})")
        if frames.dtype != e.cfg.dtype:
            frames = frames.to(e.cfg.dtype)
        if not frames.is_cuda and not frames.is_pinned():
            frames = frames.to(e.device)
        toks, _, lg_host = e.frame_step_multi(frames, first_stream=0)
        torch.cuda.current_stream().synchronize()
        results, firing = [None] * B, []
        for s, m in enumerate(self.streams):
            m._tokens.append(toks[s:s + 1])
            m._num_frames += 1
            m.last_gate_logits = lg_host[s].clone()
            pred = int(torch.softmax(m.last_gate_logits, dim=0).argmax(dim=0).item()) if force_pred is None else int(force_pred[s])
            if pred:
                m.interval_id_list.append(m._num_frames)
                firing.append(s)
            else:
                results[s] = (None, 0)
        max_new = int(kwargs.get("max_new_tokens", 1024))
        stops = self.streams[0]._stop_ids(kwargs)
        if self.streams[0]._host_criteria(kwargs):
            raise NotImplementedError("host-evaluated stopping criteria (multi-token keywords) are per stream: use "
                                      "StreamMindB200ForCausalLM.stream_generate_demo for such a stream")
        for lo in range(0, len(firing), 4):                     # the decode kernel takes up to 4 streams per pass
            group = firing[lo:lo + 4]
            items = [self.streams[s]._prefill_dialogue(list(inputs[s])) for s in group]
            outs = e.llm_decode_multi(group, [max_new] * len(group), stops)
            for s, it, out in zip(group, items, outs):
                self.streams[s]._dialogue.commit(it, out)
                results[s] = (out, 1)
        return results

    def close(self):
        self.engine.close()


def infer(model: StreamMindB200ForCausalLM, video: torch.Tensor, instruct: str, tokenizer, do_sample=False,
          version="mistral_instruct", score_video=None, prompt: Optional[str] = None, max_new_tokens: int = 1024):
    """The demo's per-frame call (eval/video_score_stream_demo.py:66-125): builds / extends the text
    prompt, tokenizes it with ``<video>`` sentinels, calls ``stream_generate_demo`` and applies the
    growth rule ``prompt += " " + outputs + " </s>[INST] <video>\\n [/INST]"`` (:123-124)."""
    from .constants import DEFAULT_MMODAL_TOKEN
    from .conversation import SeparatorStyle, conv_templates
    from .mm_utils import KeywordsStoppingCriteria, tokenizer_MMODAL_token
    conv = conv_templates["mistral_instruct"].copy()
    if prompt is None:
        conv.append_message(conv.roles[0], DEFAULT_MMODAL_TOKEN["VIDEO"] + "\n")
        conv.append_message(conv.roles[1], None)
        prompt = conv.get_prompt()
    input_ids = tokenizer_MMODAL_token(prompt, tokenizer, VIDEO_TOKEN_INDEX, return_tensors="pt").unsqueeze(0)
    stop_str = conv.sep if conv.sep_style in [SeparatorStyle.SINGLE] else conv.sep2
    stopping = KeywordsStoppingCriteria([stop_str], tokenizer, input_ids)
    outputs, pred = model.stream_generate_demo(
        input_ids, attention_mask=input_ids.ne(tokenizer.pad_token_id).long(), images_or_videos=video,
        modal_list=["video"], do_sample=do_sample, temperature=0.2 if do_sample else 0.0,
        max_new_tokens=max_new_tokens, use_cache=True, stopping_criteria=[stopping],
        pad_token_id=tokenizer.eos_token_id, score_video=score_video, tokenizer=tokenizer)
    if pred == 1:
        prompt += " " + outputs + " </s>[INST] <video>\n [/INST]"
    return outputs, prompt
