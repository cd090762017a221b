"""sm_cognition_sample on the GPU (videollama2_arch.py:595-611): kept rows and indices are IDENTICAL to the
oracle evaluated in the model dtype (integer / index work: bit-exact), including exact ties, n = 1 and the 4096-wide
projector tokens; and forward() applies it to every <video> span like the reference's eval forward (arch.py:676-681)."""
import pytest
import torch

from oracle import restate as R
from oracle.make_cognition_golden import case_tokens
from parity_util import build_engine, engine_config, f32, make_weights, oracle_configs

pytestmark = pytest.mark.gpu

CASES = [(1, 1, 64, 0.6), (2, 2, 64, 0.6), (3, 7, 64, 0.5), (4, 16, 128, 0.5), (5, 33, 256, 0.6), (6, 100, 256, 0.3),
         (7, 257, 64, 0.5), (8, 1000, 32, 0.6), (9, 12, 64, 0.01), (10, 64, 4096, 0.5), (11, 1000, 4096, 0.3), (12, 4000, 256, 0.5)]


@pytest.fixture(scope="module", params=[torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
def eng(request):
    cfg = engine_config(request.param, vit_layers=0, gate_layers=0, llm_layers=0, proj_d_model=0)
    from streammind_b200.engine import Engine
    e = Engine(cfg)
    e.finalize()
    yield e
    e.close()


@pytest.mark.parametrize("seed,n,d,p", CASES)
@pytest.mark.parametrize("mode", ["log", "similarity"])
def test_sample_matches_oracle_bit_exact(eng, seed, n, d, p, mode):
    dt = eng.cfg.dtype
    x = case_tokens(seed, n, d).to(dt)
    with R.emulate(dt):
        rows_o, idx_o = (R.exponential_sampling if mode == "log" else R.similarity_sampling)(x.float(), p)
    rows, idx = eng.cognition_sample(x.cuda(), p, mode)
    assert idx.cpu().tolist() == idx_o
    assert torch.equal(rows.cpu().float(), rows_o)


def test_exact_ties_go_to_the_lower_index(eng):
    dt = eng.cfg.dtype
    base = case_tokens(21, 4, 64).to(dt)
    x = base[[0, 1, 0, 2, 1, 0, 3]].contiguous()            # duplicated rows -> identical similarities
    with R.emulate(dt):
        _, idx_o = R.similarity_sampling(x.float(), 0.5)
    _, idx = eng.cognition_sample(x.cuda(), 0.5, "similarity")
    assert idx.cpu().tolist() == idx_o


@pytest.mark.parametrize("sample_type", ["log", "similarity"])
def test_forward_applies_cognition_sampling(sample_type):
    from streammind_b200.model import StreamMindB200ForCausalLM, VIDEO_TOKEN_INDEX
    from streammind_b200 import synth
    dt = torch.float16
    cfg = engine_config(dt, max_frames=4)
    sd = make_weights(cfg, llm=True)
    m = StreamMindB200ForCausalLM(cfg, sd, sample_per=0.5, sample_type=sample_type)
    plain = StreamMindB200ForCausalLM(cfg, None, engine=m.engine)
    frames = synth.make_frames(0, 0, 8, cfg.vit_image, dtype=dt)
    ids = torch.tensor([[1, 7, 9, VIDEO_TOKEN_INDEX, 11, 12]])
    e = m.engine
    # the span the sampled forward must have used
    e.reset_stream()
    toks = torch.cat([e.projector_step(e.vit_encode(frames[i:i + 4].cuda(), want_feats=False)[1]) for i in (0, 4)], 0)
    with R.emulate(dt):
        _, idx_o = (R.exponential_sampling if sample_type == "log" else R.similarity_sampling)(toks.float().cpu(), 0.5)
    e.reset_stream()
    out = m.forward(input_ids=ids, images=[frames])
    assert out.past_key_values.length == 5 + len(idx_o) == e.kv_len
    # same logits as a plain forward over hand-built embeddings of the kept rows
    emb = torch.cat([e.embed_tokens(torch.tensor([1, 7, 9])), toks[idx_o], e.embed_tokens(torch.tensor([11, 12]))], 0)
    ref = plain.forward(inputs_embeds=emb.unsqueeze(0))
    assert torch.equal(out.logits, ref.logits)
    e.close()
