"""The reference-facing Python API (streammind_b200/model.py) end to end on the GPU vs the oracle's
streaming twin: per-frame gate decisions, generated ids, prefix re-use of the persistent KV cache and
the B2 component hooks."""
import pytest
import torch

from oracle import restate as R
from parity_util import check_close, engine_config, f32, make_weights, oracle_configs
from streammind_b200 import synth
from streammind_b200.model import StreamMindB200ForCausalLM

pytestmark = pytest.mark.gpu


class IdTokenizer:
    pad_token_id, bos_token_id, eos_token_id = 0, 1, 2

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(str(int(t)) for t in row) for row in ids]


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_stream_generate_demo_vs_oracle(built_library, dt):
    cfg = engine_config(dt, max_frames=2, use_graphs=True)
    sd = make_weights(cfg, llm=True)
    model = StreamMindB200ForCausalLM(cfg, sd, keep_frame_features=True)
    oc = oracle_configs(cfg)
    sd32 = f32(sd)
    prompt0, turn_suffix = synth.make_prompt_ids(vocab=cfg.llm_vocab, n_sys=12, n_suffix=3)
    frames = synth.make_frames(3, 0, 8, cfg.vit_image, dtype=dt)
    force = [0, 1, 0, 1, 1, 0, 0, 1]
    NEW = 6
    tok = IdTokenizer()
    with R.emulate(dt):
        ora = R.IncrementalStream(sd32, oc)
        prompt_o = list(prompt0)
        exp = []
        for t in range(8):
            out, pred, lg, x = ora.step(prompt_o, frames[t:t + 1].float(), NEW, stop_ids=(2,), force_pred=force[t])
            exp.append((out, pred, lg.clone(), ora.prefilled if pred else 0))
            if pred:
                prompt_o = prompt_o + out + turn_suffix
    prompt = list(prompt0)
    noise_tol = {torch.float16: 4e-3, torch.bfloat16: 3e-2}[dt]
    for t in range(8):
        ids = torch.tensor([prompt])
        text, pred = model.stream_generate_demo(ids, images_or_videos=frames[t:t + 1], modal_list=["video"],
                                                attention_mask=torch.ones_like(ids), do_sample=False,
                                                max_new_tokens=NEW, use_cache=True, tokenizer=tok,
                                                pad_token_id=tok.eos_token_id, score_video=True, force_pred=force[t])
        out_o, pred_o, lg_o, prefilled_o = exp[t]
        assert pred == pred_o
        check_close(f"gate logits frame {t}", model.last_gate_logits, lg_o, 4 * {torch.float16: 1e-3, torch.bfloat16: 8e-3}[dt])
        if pred:
            got = [int(s) for s in text.split()]
            assert model.last_prefill_len == prefilled_o           # same prefix re-use as the oracle twin
            assert got == out_o, (t, got, out_o)                   # greedy ids identical on this seed
            prompt = prompt + got + turn_suffix
        else:
            assert text is None
    assert model.interval_id_list == [2, 4, 5, 8]
    assert model.frame_feature.shape == (1, 8, cfg.num_patches, cfg.vit_hidden)
    # state errors behave like exceptions, not silent fallbacks
    with pytest.raises(NotImplementedError):
        model.stream_generate_demo(torch.tensor([prompt]), images_or_videos=frames[:1], do_sample=True)
    model.engine.close()


class KeywordTokenizer(IdTokenizer):
    """a keyword string "a b c" tokenises to the ids [a, b, c] (what the reference's tokenizer(keyword).input_ids yields)"""

    class _Enc:
        def __init__(self, ids):
            self.input_ids = ids

    def __call__(self, text):
        return self._Enc([int(t) for t in text.split()])


def test_multi_token_stop_keyword_on_the_host(built_library):
    """KeywordsStoppingCriteria with a multi-token keyword (reference mm_utils.py:616-647) cannot run in the device loop: the
    model decodes in chunks, checks on the host after every token like hf generate(), cuts the output at the hit, rewinds the
    cache, and the stream continues exactly as if the device had stopped there."""
    from streammind_b200.mm_utils import KeywordsStoppingCriteria
    dt = torch.bfloat16
    cfg = engine_config(dt, max_frames=2, use_graphs=True)
    sd = make_weights(cfg, llm=True)
    model = StreamMindB200ForCausalLM(cfg, sd)
    model.HOST_CHECK_CHUNK = 4                         # several chunk boundaries inside one answer
    prompt0, turn_suffix = synth.make_prompt_ids(vocab=cfg.llm_vocab, n_sys=12, n_suffix=3)
    frames = synth.make_frames(5, 0, 3, cfg.vit_image, dtype=dt)
    tok = KeywordTokenizer()
    ids = torch.tensor([list(prompt0)])
    kw = dict(images_or_videos=frames[:1], modal_list=["video"], do_sample=False, use_cache=True, tokenizer=tok, force_pred=1)
    free, _ = model.stream_generate_demo(ids, max_new_tokens=14, **kw)
    free = [int(t) for t in free.split()]
    assert len(free) == 14
    # chunked decode without a hit reproduces the one-call decode (the chunk seams go through a one-position prefill)
    never = KeywordsStoppingCriteria([f"{cfg.llm_vocab + 5} {cfg.llm_vocab + 6}"], tok, ids)
    model.reset_stream()
    chunked, _ = model.stream_generate_demo(ids, max_new_tokens=14, stopping_criteria=[never], **kw)
    assert [int(t) for t in chunked.split()] == free
    # a two-token keyword whose first occurrence lies beyond the first chunk
    stop_at = next(j for j in range(5, 13) if not any(free[i:i + 2] == free[j:j + 2] for i in range(j)))
    crit = KeywordsStoppingCriteria([f"{free[stop_at]} {free[stop_at + 1]}"], tok, ids)
    assert crit.needs_host_check
    model.reset_stream()
    cut, _ = model.stream_generate_demo(ids, max_new_tokens=14, stopping_criteria=[crit], **kw)
    cut = [int(t) for t in cut.split()]
    assert cut == free[:stop_at + 2], (cut, free, stop_at)
    assert model.engine.kv_len == model.last_prefill_len + len(cut) - 1       # everything but the last kept token is cached
    # the stream goes on: next turn re-uses the cached prefix (only the suffix + the new frame token are prefilled)
    ids2 = torch.tensor([list(prompt0) + cut + turn_suffix])
    nxt, pred = model.stream_generate_demo(ids2, max_new_tokens=4, **{**kw, "images_or_videos": frames[1:2]})
    assert pred == 1 and len(nxt.split()) == 4
    assert model.last_prefill_len == 1 + len(turn_suffix)                     # last kept token + suffix (its <video> included)
    model.engine.close()


def test_component_hooks(built_library):
    dt = torch.float16
    cfg = engine_config(dt, max_frames=2, llm_layers=0, use_graphs=False)
    sd = make_weights(cfg)
    model = StreamMindB200ForCausalLM(cfg, sd)
    oc = oracle_configs(cfg)
    sd32 = f32(sd)
    frames = synth.make_frames(1, 0, 5, cfg.vit_image, dtype=dt)
    tower, proj = model.get_vision_tower(), model.mm_projector
    feats = tower(frames.cuda())                      # 5 frames through max_frames=2 chunks
    assert feats.shape == (5, tower.num_patches, tower.hidden_size) and feats.dtype == dt
    with R.emulate(dt):
        feats_o = R.clip_vision_tower(sd32, oc.vit, frames.float())
        st = R.MambaState.zeros(oc.mamba)
        toks_o = torch.stack([R.projector_step(sd32, oc.mamba, R.pool_patches(feats_o[t]), st) for t in range(5)])
        lg_o = R.gate_logits_degenerate(sd32, oc.gate, toks_o[-1])
    check_close("tower features", feats, feats_o, 1e-3)
    # the reference hands the WHOLE history to mm_projector on every call
    x3, lg3 = proj(feats[:3].unsqueeze(0), cls_demo=True, frames_features_shape=[])
    x5, lg5 = proj(feats.unsqueeze(0), cls_demo=True, frames_features_shape=[])
    assert x3.shape == (1, 3, cfg.proj_d_model) and x5.shape == (1, 5, cfg.proj_d_model)
    assert torch.equal(x5[0, :3], x3[0])
    check_close("projector tokens", x5[0], toks_o, 2e-3)
    check_close("gate logits", lg5, lg_o, 4e-3)
    assert torch.equal(proj(feats.unsqueeze(0)), x5)           # no flags -> tokens only
    with pytest.raises(RuntimeError):
        proj(feats[:2].unsqueeze(0), cls_demo=True)
    with pytest.raises(NotImplementedError):
        proj(feats.unsqueeze(0), cls_training=True)
    lst = tower([frames[0].cuda(), frames[1].cuda()])
    assert torch.equal(lst[0][0], feats[0]) and len(lst) == 2
    model.engine.close()
