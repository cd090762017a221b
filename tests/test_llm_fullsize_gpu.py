"""LLM prefill + decode at the BASELINE widths (hidden 4096, FFN 14336, 32 q / 8 kv heads x 128, vocab 32002) against
the oracle (hf MistralForCausalLM arithmetic with the reference's rounding points, oracle/restate.py), through the C ABI.

The model is a few layers deep so the CPU oracle finishes in seconds; every kernel runs at its BASELINE shape:
K = 4096 and K = 14336 weight streams, GQA group 4, vocab-32002 argmax, split-KV decode attention at ctx 4k / 8k.
Greedy ids: compared id by id; a mismatch is accepted only where the oracle's own top-1 margin is below the numerical
noise of the logits (random-init logits are nearly flat), and never before MIN_EXACT tokens."""
import pytest
import torch

from oracle import restate as R
from parity_util import build_engine, engine_config, f32, make_weights, oracle_configs, rel_err

pytestmark = pytest.mark.gpu

LOGIT_TOL = {torch.float16: 4e-3, torch.bfloat16: 3e-2}      # relative to max |logit| (same bound as the small-size test)
FULL = dict(small=False, vit_layers=0, proj_d_model=0, gate_layers=0)


def _setup(dt, layers, max_ctx=1024, **over):
    cfg = engine_config(dt, llm_layers=layers, llm_max_ctx=max_ctx, **FULL, **over)
    sd = make_weights(cfg, vit=False, proj=False, gate=False, llm=True)
    return cfg, sd, build_engine(cfg, sd), oracle_configs(cfg), f32(sd)


def _compare_ids(out, out_o, lg_o, noise, min_exact):
    agree = 0
    for i, (a, b) in enumerate(zip(out, out_o)):
        if a == b:
            agree += 1
            continue
        margin = (lg_o[i][b] - lg_o[i][a]).item()
        assert i >= min_exact, f"token {i} differs ({a} vs oracle {b}) before {min_exact} exact tokens"
        assert margin <= 2 * noise, f"token {i}: got {a}, oracle {b}, oracle margin {margin:.4f} > noise {noise:.4f}"
        break
    return agree


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_full_width_prefill_and_decode_two_layers(built_library, dt):
    cfg, sd, eng, oc, sd32 = _setup(dt, layers=2)
    g = torch.Generator().manual_seed(5)
    P, NEW = 40, 32
    ids = torch.randint(3, cfg.llm_vocab, (P,), generator=g)
    emb = eng.embed_tokens(ids.cuda())
    logits = eng.llm_prefill(emb, want_logits=True)
    out = eng.llm_decode(NEW)
    last = eng.last_decode_logits(0).cpu()
    assert len(out) == NEW and eng.kv_len == P + NEW - 1
    with R.emulate(dt):
        out_o, lg_o = R.greedy_decode(sd32, oc.llm, sd32["model.embed_tokens.weight"][ids], R.KVCache(), NEW, return_logits=True)
    e = rel_err(logits, lg_o[0])
    print(f"full-width prefill logits rel err {e}")
    assert max(e) < LOGIT_TOL[dt], e
    noise = LOGIT_TOL[dt] * lg_o[0].abs().max().item()
    agree = _compare_ids(out, out_o, lg_o, noise, min_exact=4)
    print(f"greedy ids: {agree}/{NEW} identical before the first (near-tie) divergence")
    if agree == NEW:        # the whole trajectory matched: the last step's logits must match the oracle's too
        e = rel_err(last, lg_o[-1])
        print(f"logits of decode step {NEW - 1}: rel err {e}")
        assert max(e) < LOGIT_TOL[dt], e
    # run-to-run determinism of the persistent kernel (fixed-order reductions)
    eng.kv_set_len(0)
    eng.llm_prefill(emb)
    assert eng.llm_decode(NEW) == out
    eng.close()


@pytest.mark.parametrize("dt", [torch.bfloat16])
def test_decode_attention_long_context(built_library, dt):
    """One full-width layer, context 4 096 and 8 192: the KV cache is filled by sm_llm_prefill, the oracle's cache
    analytically (with one layer K / V depend on the embeddings only, so no O(ctx^2) attention runs on the CPU); then
    single decode steps are compared: split-KV attention over 18 slices x 8 kv heads, fp32 softmax reference."""
    cfg, sd, eng, oc, sd32 = _setup(dt, layers=1, max_ctx=8704)
    g = torch.Generator().manual_seed(9)
    N = 8190
    ids = torch.randint(3, cfg.llm_vocab, (N,), generator=g)
    emb_w = sd32["model.embed_tokens.weight"]
    lp = "model.layers.0."
    with R.emulate(dt):
        x = emb_w[ids]
        a = R.rms_norm(x, sd32[lp + "input_layernorm.weight"], oc.llm.rms_norm_eps)
        Hk, D = oc.llm.num_kv_heads, oc.llm.head_dim
        k = R._linear(a, sd32[lp + "self_attn.k_proj.weight"]).view(N, Hk, D).transpose(0, 1)
        v = R._linear(a, sd32[lp + "self_attn.v_proj.weight"]).view(N, Hk, D).transpose(0, 1)
        cos, sin = R.rope_cos_sin(oc.llm, torch.arange(N))
        k = R.apply_rope(k, cos, sin)
    emb = eng.embed_tokens(ids.cuda())
    for ctx in (4095, 8190):
        # device: extend the cache to ctx - 1 positions, then prefill ONE more token to obtain start logits
        for lo in range(eng.kv_len, ctx - 1, 512):
            eng.llm_prefill(emb[lo:min(lo + 512, ctx - 1)])
        eng.llm_prefill(emb[ctx - 1:ctx])
        out = eng.llm_decode(3)                       # steps at context ctx + 1 and ctx + 2
        last = eng.last_decode_logits(0).cpu()
        with R.emulate(dt):
            cache = R.KVCache(k=[k[:, :ctx - 1].clone()], v=[v[:, :ctx - 1].clone()])
            out_o, lg_o = R.greedy_decode(sd32, oc.llm, emb_w[ids[ctx - 1:ctx]], cache, 3, return_logits=True)
        noise = LOGIT_TOL[dt] * lg_o[0].abs().max().item()
        agree = _compare_ids(out, out_o, lg_o, noise, min_exact=1)
        print(f"ctx {ctx}: ids {out} oracle {out_o}")
        if agree == 3:
            e = rel_err(last, lg_o[-1])
            print(f"ctx {ctx + 2}: decode logits rel err {e}")
            assert max(e) < LOGIT_TOL[dt], e
        eng.kv_set_len(ctx - 1)
    eng.close()


def test_multi_stream_decode_equals_single(built_library):
    """Multi-stream batching: B streams decoded together, one weight pass per step, give exactly the ids
    of B independent runs (per-stream KV cache, position and stop state; different prompt lengths and budgets)."""
    dt = torch.bfloat16
    cfg, sd, eng, oc, sd32 = _setup(dt, layers=2, n_streams=4)
    g = torch.Generator().manual_seed(21)
    prompts = [torch.randint(3, cfg.llm_vocab, (n,), generator=g) for n in (17, 40, 5, 29)]
    budgets = [24, 9, 16, 24]
    single = []
    for s, ids in enumerate(prompts):
        eng.select_stream(s)
        eng.llm_prefill(eng.embed_tokens(ids.cuda()))
        single.append(eng.llm_decode(budgets[s]))
        eng.kv_set_len(0)
    for s, ids in enumerate(prompts):
        eng.select_stream(s)
        eng.llm_prefill(eng.embed_tokens(ids.cuda()))
    multi = eng.llm_decode_multi([0, 1, 2, 3], budgets)
    assert multi == single, (multi, single)
    for s, ids in enumerate(prompts):
        eng.select_stream(s)
        assert eng.kv_len == len(ids) + budgets[s] - 1
    # a stop id ends one stream early while the others continue
    stop = single[1][3]
    for s, ids in enumerate(prompts):
        eng.select_stream(s)
        eng.kv_set_len(0)
        eng.llm_prefill(eng.embed_tokens(ids.cuda()))
    stopped = eng.llm_decode_multi([0, 1, 2, 3], budgets, stop_ids=[stop])
    for s in range(4):
        exp = single[s]
        if stop in exp:
            exp = exp[: exp.index(stop) + 1]
        assert stopped[s] == exp, (s, stopped[s], exp)
    # two of the streams as their own batch: lanes map to arbitrary stream ids
    for s in (3, 1):
        eng.select_stream(s)
        eng.kv_set_len(0)
        eng.llm_prefill(eng.embed_tokens(prompts[s].cuda()))
    pair = eng.llm_decode_multi([3, 1], [budgets[3], budgets[1]])
    assert pair == [single[3], single[1]]
    eng.close()
