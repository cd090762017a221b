"""Tensor-in / tensor-out wrapper over the C ABI (PyTorch tensors for device memory and streams only).

``Engine`` owns one ``sm_handle`` = one GPU + one video stream's state.  Every method forwards to the
entry point of the same name in include/streammind_b200.h on ``torch.cuda.current_stream()`` and raises
``RuntimeError`` with ``sm_last_error`` on failure (the reference's error convention is Python
exceptions).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib as _lib

# openai/clip-vit-large-patch14-336 preprocessor_config.json (CLIPImageProcessor image_mean / image_std)
OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

_DT = {torch.float16: _lib.SM_DTYPE_F16, torch.bfloat16: _lib.SM_DTYPE_BF16}


@dataclass
class EngineConfig:
    """Dimensions of the four sub-models.  Defaults = the BASELINE.json configuration:
    CLIP-ViT-L/14-336 (select_layer -2 -> 23 layers executed), projector d_model 4096, gate =
    4-layer default MistralConfig, LLM = Mistral-7B (vocab 32000 + 2 special tokens,
    /root/reference/streammind/train_new_stream.py:857-858)."""
    dtype: torch.dtype = torch.float16
    max_frames: int = 1
    vit_image: int = 336
    vit_patch: int = 14
    vit_hidden: int = 1024
    vit_layers: int = 23
    vit_heads: int = 16
    vit_ffn: int = 4096
    vit_eps: float = 1e-5
    proj_d_model: int = 4096
    proj_d_state: int = 16
    proj_d_conv: int = 4
    proj_expand: int = 2
    proj_eps: float = 1e-5
    gate_layers: int = 4
    gate_heads: int = 32
    gate_kv_heads: int = 8
    gate_head_dim: int = 128
    gate_ffn: int = 14336
    gate_eps: float = 1e-6
    llm_hidden: int = 4096
    llm_layers: int = 32
    llm_heads: int = 32
    llm_kv_heads: int = 8
    llm_head_dim: int = 128
    llm_ffn: int = 14336
    llm_vocab: int = 32002
    llm_max_ctx: int = 8704
    llm_eps: float = 1e-5
    llm_rope_theta: float = 1e6
    use_graphs: bool = True
    n_streams: int = 1          # video streams per handle (multi-stream batching of the LLM decode)

    def to_c(self) -> _lib.SmConfig:
        c = _lib.SmConfig()
        for name, _ in _lib.SmConfig._fields_:
            v = getattr(self, name)
            if name == "dtype":
                v = _DT[v]
            setattr(c, name, int(v) if isinstance(v, bool) else v)
        return c

    @property
    def num_patches(self) -> int:
        return (self.vit_image // self.vit_patch) ** 2


class Engine:
    def __init__(self, cfg: EngineConfig, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("streammind_b200.Engine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device("cuda", device)
        self._h = C.c_void_p()
        cc = cfg.to_c()
        if self.lib.sm_create(C.byref(self._h), device, C.byref(cc)) != 0:
            raise RuntimeError(self.lib.sm_last_error(None).decode())
        self._pinned_logits = torch.empty(cfg.max_frames, 2, dtype=torch.float32).pin_memory()
        self._pinned_ring = torch.empty(16, cfg.max_frames, 2, dtype=torch.float32).pin_memory()   # frame_submit tickets (kTicketRing)
        # Buffers of the tickets in flight: the library reads `pixels` and writes the outputs on its internal streams after
        # frame_submit has returned, so they must outlive the call.  One entry per ring slot, released when the slot is
        # re-used (the library blocks on the old ticket first) -- the caller may drop or re-bind its tensors immediately.
        self._ring_refs = [None] * 16

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError(self.lib.sm_last_error(self._h).decode())

    @staticmethod
    def _stream() -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _t(self, t: torch.Tensor, name: str) -> torch.Tensor:
        if t.device != self.device:
            raise RuntimeError(f"{name}: expected a tensor on {self.device}, got {t.device}")
        if t.dtype != self.cfg.dtype:
            raise RuntimeError(f"{name}: expected dtype {self.cfg.dtype}, got {t.dtype}")
        return t.contiguous()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.sm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict_unknown: bool = False) -> int:
        """Upload weights by the reference model's own state_dict keys (host or device tensors)."""
        n = 0
        for k, v in sd.items():
            if v.dtype != self.cfg.dtype:
                v = v.to(self.cfg.dtype)
            v = v.contiguous()
            shape = (C.c_int64 * max(1, v.dim()))(*(list(v.shape) or [1]))
            on_host = 0 if v.is_cuda else 1
            rc = self.lib.sm_load_weight(self._h, k.encode(), C.c_void_p(v.data_ptr()), on_host,
                                         _DT[self.cfg.dtype], max(1, v.dim()), shape)
            if rc != 0:
                msg = self.lib.sm_last_error(self._h).decode()
                if "unknown key" in msg and not strict_unknown:
                    continue
                raise RuntimeError(msg)
            n += 1
        return n

    def finalize(self):
        self._check(self.lib.sm_finalize_weights(self._h))

    def reset_stream(self):
        """Zero the selected stream's Mamba state and KV length (ordered on the current CUDA stream)."""
        self._check(self.lib.sm_stream_reset(self._h, self._stream()))

    def select_stream(self, stream_id: int):
        """Multi-stream handles: the stream the single-stream calls act on (sm_stream_select)."""
        self._check(self.lib.sm_stream_select(self._h, int(stream_id)))

    # ------------------------------------------------------------------ frame preprocessing
    def preprocess_frames(self, frames, image_mean=OPENAI_CLIP_MEAN, image_std=OPENAI_CLIP_STD) -> torch.Tensor:
        """uint8 RGB frames [n, H, W, 3] (numpy / CPU tensor / CUDA tensor) -> pixels [n, 3, image, image] in the model
        dtype on the device: expand2square + CLIPImageProcessor.preprocess + .half() of the reference
        (mm_utils.py:446-464, 257-268), computed by sm_preprocess_frames."""
        import numpy as np
        t = torch.from_numpy(np.ascontiguousarray(frames)) if isinstance(frames, np.ndarray) else frames.contiguous()
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
            raise ValueError(f"frames must be uint8 [n, H, W, 3], got {t.dtype} {tuple(t.shape)}")
        if t.is_cuda and t.device != self.device:
            raise ValueError(f"frames are on {t.device}, the engine is on {self.device}")
        n, H, W, _ = t.shape
        img = self.cfg.vit_image
        out = torch.empty(n, 3, img, img, dtype=self.cfg.dtype, device=self.device)
        mean = (C.c_float * 3)(*[float(x) for x in image_mean])
        std = (C.c_float * 3)(*[float(x) for x in image_std])
        bg = (C.c_int * 3)(*[int(x * 255) for x in image_mean])      # the reference's background colour expression
        self._check(self.lib.sm_preprocess_frames(self._h, t.data_ptr(), n, H, W, 1 if t.is_cuda else 0, mean, std, bg,
                                                  out.data_ptr(), self._stream()))
        if not t.is_cuda and not t.is_pinned():
            torch.cuda.current_stream(self.device).synchronize()          # the pageable host buffer may go away with `t`
        return out

    # ------------------------------------------------------------------ sub-model calls
    def vit_encode(self, pixels: torch.Tensor, want_feats: bool = True) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        px = self._t(pixels, "pixels")
        B = px.shape[0]
        c = self.cfg
        feats = torch.empty(B, c.num_patches, c.vit_hidden, dtype=c.dtype, device=self.device) if want_feats else None
        pooled = torch.empty(B, c.vit_hidden, dtype=c.dtype, device=self.device)
        self._check(self.lib.sm_vit_encode(self._h, px.data_ptr(), B, feats.data_ptr() if want_feats else None,
                                           pooled.data_ptr(), self._stream()))
        return feats, pooled

    def pool_features(self, feats: torch.Tensor) -> torch.Tensor:
        f = self._t(feats, "feats")
        n = f.shape[0]
        pooled = torch.empty(n, self.cfg.vit_hidden, dtype=self.cfg.dtype, device=self.device)
        self._check(self.lib.sm_pool_features(self._h, f.data_ptr(), n, pooled.data_ptr(), self._stream()))
        return pooled

    def projector_step(self, pooled: torch.Tensor) -> torch.Tensor:
        p = self._t(pooled, "pooled")
        n = p.shape[0]
        toks = torch.empty(n, self.cfg.proj_d_model, dtype=self.cfg.dtype, device=self.device)
        self._check(self.lib.sm_projector_step(self._h, p.data_ptr(), n, toks.data_ptr(), self._stream()))
        return toks

    def gate_score(self, tok: torch.Tensor) -> torch.Tensor:
        t = self._t(tok, "tok")
        logits = torch.empty(2, dtype=torch.float32, device=self.device)
        self._check(self.lib.sm_gate_score(self._h, t.data_ptr(), logits.data_ptr(), self._stream()))
        return logits

    def frame_step(self, pixels: torch.Tensor, want_feats: bool = False, want_device_outputs: bool = True):
        """ViT -> projector -> gate for B frames.  ``pixels`` may be a pinned host tensor (the H2D copy
        is then part of the call).  Returns (feats|None, toks|None, logits_device|None, logits_pinned_host);
        the pinned host logits are valid after the current stream is synchronised."""
        c = self.cfg
        if pixels.dtype != c.dtype:
            raise RuntimeError(f"pixels: expected dtype {c.dtype}, got {pixels.dtype}")
        on_host = 0 if pixels.is_cuda else 1
        if on_host and not pixels.is_pinned():
            raise RuntimeError("pixels: host tensors must be pinned")
        px = pixels.contiguous()
        B = px.shape[0]
        feats = torch.empty(B, c.num_patches, c.vit_hidden, dtype=c.dtype, device=self.device) if want_feats else None
        toks = torch.empty(B, c.proj_d_model, dtype=c.dtype, device=self.device) if want_device_outputs else None
        logits = torch.empty(B, 2, dtype=torch.float32, device=self.device) if want_device_outputs else None
        self._check(self.lib.sm_frame_step(
            self._h, px.data_ptr(), on_host, B, feats.data_ptr() if feats is not None else None,
            toks.data_ptr() if toks is not None else None, logits.data_ptr() if logits is not None else None,
            self._pinned_logits.data_ptr(), self._stream()))
        return feats, toks, logits, self._pinned_logits[:B]

    def frame_step_multi(self, pixels: torch.Tensor, first_stream: int = 0):
        """One frame of each of ``pixels.shape[0]`` consecutive stream slots (sm_frame_step_multi): -> (toks [n, d_model],
        logits_device [n, 2], logits_pinned_host [n, 2]); the host logits are valid after the current stream is synchronised."""
        c = self.cfg
        if pixels.dtype != c.dtype:
            raise RuntimeError(f"pixels: expected dtype {c.dtype}, got {pixels.dtype}")
        on_host = 0 if pixels.is_cuda else 1
        if on_host and not pixels.is_pinned():
            raise RuntimeError("pixels: host tensors must be pinned")
        px = pixels.contiguous()
        n = px.shape[0]
        toks = torch.empty(n, c.proj_d_model, dtype=c.dtype, device=self.device)
        logits = torch.empty(n, 2, dtype=torch.float32, device=self.device)
        self._check(self.lib.sm_frame_step_multi(self._h, px.data_ptr(), on_host, n, int(first_stream), toks.data_ptr(),
                                                 logits.data_ptr(), self._pinned_logits.data_ptr(), self._stream()))
        return toks, logits, self._pinned_logits[:n]

    def frame_submit(self, pixels: torch.Tensor, want_feats: bool = False, want_device_outputs: bool = False):
        """Pipelined frame_step (sm_frame_submit): returns (ticket, feats|None, toks|None, logits_device|None,
        logits_pinned_host).  Nothing is ordered on the current stream: call ``frame_wait(ticket)`` before using
        any output (``block=True`` for the pinned host logits)."""
        c = self.cfg
        if pixels.dtype != c.dtype:
            raise RuntimeError(f"pixels: expected dtype {c.dtype}, got {pixels.dtype}")
        on_host = 0 if pixels.is_cuda else 1
        if on_host and not pixels.is_pinned():
            raise RuntimeError("pixels: host tensors must be pinned")
        px = pixels.contiguous()
        B = px.shape[0]
        feats = torch.empty(B, c.num_patches, c.vit_hidden, dtype=c.dtype, device=self.device) if want_feats else None
        toks = torch.empty(B, c.proj_d_model, dtype=c.dtype, device=self.device) if want_device_outputs else None
        logits = torch.empty(B, 2, dtype=torch.float32, device=self.device) if want_device_outputs else None
        tk = C.c_longlong(0)
        nxt = getattr(self, "_next_ticket", 0)
        host = self._pinned_ring[nxt % 16]
        self._check(self.lib.sm_frame_submit(
            self._h, px.data_ptr(), on_host, B, feats.data_ptr() if feats is not None else None,
            toks.data_ptr() if toks is not None else None, logits.data_ptr() if logits is not None else None,
            host.data_ptr(), self._stream(), C.byref(tk)))
        self._next_ticket = tk.value + 1
        self._ring_refs[tk.value % 16] = (px, feats, toks, logits)
        return tk.value, feats, toks, logits, host[:B]

    def frame_wait(self, ticket: int, block: bool = True, on_stream: bool = True):
        """Order the current stream after (on_stream) and/or block the host until (block) ticket's outputs."""
        self._check(self.lib.sm_frame_wait(self._h, ticket, self._stream() if on_stream else None, 1 if block else 0))

    def cognition_sample(self, toks: torch.Tensor, percentage: float, sample_type: str):
        """exponential_sampling ('log') / similarity_sampling ('similarity') of a frame-token segment [n, d]
        (videollama2_arch.py:595-611) -> (kept rows [k, d] in their original order, their indices [k] int32)."""
        t = self._t(toks, "toks")
        n, d = t.shape
        mode = {"log": 0, "similarity": 1}[sample_type]
        k = self.lib.sm_cognition_count(n, float(percentage), mode)
        out = torch.empty(k, d, dtype=t.dtype, device=self.device)
        idx = torch.empty(k, dtype=torch.int32, device=self.device)
        self._check(self.lib.sm_cognition_sample(self._h, t.data_ptr(), n, d, mode, float(percentage), out.data_ptr(), idx.data_ptr(), self._stream()))
        return out, idx

    def embed_tokens(self, ids: torch.Tensor) -> torch.Tensor:
        if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= self.cfg.llm_vocab):
            # the reference's nn.Embedding raises IndexError here (e.g. an unexpanded <image> = -200 sentinel)
            raise ValueError(f"token ids must lie in [0, {self.cfg.llm_vocab}), got [{int(ids.min())}, {int(ids.max())}]")
        ids32 = ids.to(device=self.device, dtype=torch.int32).contiguous()
        out = torch.empty(ids32.numel(), self.cfg.llm_hidden, dtype=self.cfg.dtype, device=self.device)
        self._check(self.lib.sm_embed_tokens(self._h, ids32.data_ptr(), ids32.numel(), out.data_ptr(), self._stream()))
        return out

    def llm_prefill(self, embeds: torch.Tensor, want_logits: bool = False) -> Optional[torch.Tensor]:
        e = self._t(embeds, "embeds")
        logits = torch.empty(self.cfg.llm_vocab, dtype=torch.float32, device=self.device) if want_logits else None
        self._check(self.lib.sm_llm_prefill(self._h, e.data_ptr(), e.shape[0],
                                            logits.data_ptr() if want_logits else None, self._stream()))
        return logits

    def llm_decode(self, max_new: int, stop_ids: Sequence[int] = ()) -> List[int]:
        out = (C.c_int32 * max_new)()
        n = C.c_int32(0)
        stops = (C.c_int32 * max(1, len(stop_ids)))(*stop_ids)
        self._check(self.lib.sm_llm_decode(self._h, max_new, stops, len(stop_ids), out, C.byref(n), self._stream()))
        return list(out[: n.value])

    def llm_decode_multi(self, stream_ids: Sequence[int], max_new: Sequence[int], stop_ids: Sequence[int] = ()) -> List[List[int]]:
        """Greedy-decode several prefilled streams of the handle together: one pass over the LLM weights per step
        serves every listed stream (sm_llm_decode_multi).  Returns the new ids per stream."""
        n = len(stream_ids)
        stride = max(int(m) for m in max_new)
        out = (C.c_int32 * (n * stride))()
        nout = (C.c_int32 * n)()
        sids = (C.c_int * n)(*[int(s) for s in stream_ids])
        mx = (C.c_int * n)(*[int(m) for m in max_new])
        stops = (C.c_int32 * max(1, len(stop_ids)))(*stop_ids)
        self._check(self.lib.sm_llm_decode_multi(self._h, n, sids, mx, stops, len(stop_ids), out, stride, nout, self._stream()))
        return [list(out[i * stride: i * stride + nout[i]]) for i in range(n)]

    def last_decode_logits(self, lane: int = 0) -> torch.Tensor:
        """fp32 logits [vocab] of the last decode step of `lane` (test hook, sm_debug_decode_logits)."""
        out = torch.empty(self.cfg.llm_vocab, dtype=torch.float32, device=self.device)
        self._check(self.lib.sm_debug_decode_logits(self._h, lane, out.data_ptr(), self._stream()))
        return out

    def decode_stats(self, reset: bool = True) -> Dict[str, float]:
        """Device time of the decode-step launches since the last reset (sm_decode_stats)."""
        ms, steps, toks, ctx = C.c_double(0), C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        self._check(self.lib.sm_decode_stats(self._h, C.byref(ms), C.byref(steps), C.byref(toks), C.byref(ctx), 1 if reset else 0))
        return {"ms": ms.value, "steps": steps.value, "tokens": toks.value, "ctx_sum": ctx.value}

    @property
    def kv_len(self) -> int:
        return self.lib.sm_kv_len(self._h)

    def kv_set_len(self, n: int):
        self._check(self.lib.sm_kv_set_len(self._h, n))

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.sm_launch_count(self._h, 1 if reset else 0))

    KERNEL_CLASSES = ["gemm_tc_kernel", "gemv_kernel", "attention_kernel", "layernorm_kernel", "im2col_kernel",
                      "vit_finalize_kernel", "mamba_scan_step_kernel", "gate_gemm_kernel"]

    def kernel_filter(self, classes=None):
        """Launch only the named kernel classes (None = all).  Measurement aid, see sm_debug_kernel_filter."""
        if classes is None:
            mask = 0xFFFFFFFF
        else:
            mask = 0
            for i in range(16):
                if self.lib.sm_profile_class_name(i).decode() in classes:
                    mask |= 1 << i
        self._check(self.lib.sm_debug_kernel_filter(self._h, mask))

    def profile(self, on: bool):
        self._check(self.lib.sm_profile_enable(self._h, 1 if on else 0))

    def profile_read(self) -> Dict[str, Tuple[float, int]]:
        """{kernel class: (accumulated ms, launches)} since the last read (synchronises the device)."""
        n = 16
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        k = self.lib.sm_profile_read(self._h, n, ms, cnt)
        return {self.lib.sm_profile_class_name(i).decode(): (ms[i], int(cnt[i])) for i in range(k) if cnt[i]}

    # ------------------------------------------------------------------ unit-test hooks
    def test_gemm(self, x, w, bias, epi: int, out: Optional[torch.Tensor] = None, force_swap=-1, force_bn=0):
        M, K = x.shape
        N = w.shape[0]
        if out is None:
            out = torch.empty(M, N, dtype=torch.float32 if epi == 3 else x.dtype, device=x.device)
        self._check(self.lib.sm_test_gemm(self._h, x.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None,
                                          out.data_ptr(), M, N, K, epi, force_swap, force_bn, self._stream()))
        return out

    def test_kv_attention(self, q: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, pos0: int, n_splits: int = 0) -> torch.Tensor:
        """q [P, Hq, 128] (rotated), kcache / vcache [Hk, max_ctx, 128] holding pos0 + P positions -> [P, Hq * 128].
        n_splits: -1 mma.sync kernel, 0 tcgen05 kernel with the planned split, > 0 forced split count."""
        P, Hq, D = q.shape
        Hk, max_ctx, _ = kcache.shape
        assert D == 128 and q.is_contiguous() and kcache.is_contiguous() and vcache.is_contiguous()
        out = torch.empty(P, Hq * D, dtype=q.dtype, device=q.device)
        self._check(self.lib.sm_test_kv_attention(self._h, q.data_ptr(), Hq * D, kcache.data_ptr(), vcache.data_ptr(), max_ctx,
                                                  out.data_ptr(), P, pos0, Hq, Hk, n_splits, self._stream()))
        return out

    def test_attention(self, qkv: torch.Tensor, B: int, S: int, H: int, D: int, mode: int = -1) -> torch.Tensor:
        """mode: -1 default kernel choice, 0 mma.sync kernel, 2 tcgen05 kernel (sm_debug_attention_mode)."""
        self._check(self.lib.sm_debug_attention_mode(self._h, mode))
        out = torch.empty(B * S, H * D, dtype=qkv.dtype, device=qkv.device)
        self._check(self.lib.sm_test_attention(self._h, qkv.data_ptr(), out.data_ptr(), B, S, H, D, self._stream()))
        return out
