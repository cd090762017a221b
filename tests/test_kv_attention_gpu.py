"""LLM prefill attention (causal, GQA, head_dim 128, queries against a persistent KV cache) vs torch: fp32 softmax over the
visible keys, probabilities rounded to T before P @ V (hf MistralForCausalLM SDPA arithmetic with the reference's rounding
points).  Covers the shapes of the path: a fire's dialogue suffix (11-74 positions) against up to 8k cached positions with the
key range split over the machine, a first prompt (no cache), 512-position chunks, ragged tails, both GQA groupings in use
(Mistral 32 / 8 and the small test model 2 / 1) -- for the tcgen05 kernel (planned and forced splits) and the mma.sync one."""
import pytest
import torch

from parity_util import engine_config, rel_err

pytestmark = pytest.mark.gpu


def _reference(q, k, v, pos0, dt):
    P, Hq, D = q.shape
    Hk = k.shape[0]
    n = pos0 + P
    kk = k[:, :n].float().repeat_interleave(Hq // Hk, dim=0)           # [Hq, n, D]
    vv = v[:, :n].float().repeat_interleave(Hq // Hk, dim=0)
    s = torch.einsum("phd,hnd->hpn", q.float(), kk) * D ** -0.5
    vis = torch.arange(n, device=q.device)[None, :] <= (pos0 + torch.arange(P, device=q.device))[:, None]
    s = s.masked_fill(~vis[None], float("-inf"))
    p = torch.softmax(s, dim=-1).to(dt).float()
    return torch.einsum("hpn,hnd->phd", p, vv).reshape(P, Hq * D)


CASES = [  # P, pos0, Hq, Hk, max_ctx
    (11, 2048, 32, 8, 8704), (30, 8000, 32, 8, 8704), (74, 4001, 32, 8, 8704), (1, 777, 32, 8, 8704),
    (40, 0, 32, 8, 1024), (512, 0, 32, 8, 1024), (512, 512, 32, 8, 1024), (130, 61, 32, 8, 1024),
    (33, 95, 2, 1, 512), (200, 0, 2, 1, 512), (7, 300, 8, 8, 520),
]


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("P,pos0,Hq,Hk,max_ctx", CASES)
@pytest.mark.parametrize("splits", [0, 1, 3, -1])      # planned split, one split, forced 3 splits (some empty), mma.sync kernel
def test_kv_attention(built_library, dt, P, pos0, Hq, Hk, max_ctx, splits):
    from streammind_b200.engine import Engine
    if splits == 1 and (pos0 + P) * P > 2_000_000:
        pytest.skip("one split of a long key range with few rows: covered by the planned split")
    if splits > 1 and splits * P > 592:
        pytest.skip("forced split x positions beyond the partial buffer (the planner never asks for it)")
    eng = Engine(engine_config(dt, small=False, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
    g = torch.Generator(device="cuda").manual_seed(P * 31 + pos0)
    q = (torch.randn(P, Hq, 128, generator=g, device="cuda") * 1.2).to(dt)
    # the whole cache is filled: rows beyond pos0 + P hold stale (finite) values that the causal mask must hide
    k = (torch.randn(Hk, max_ctx, 128, generator=g, device="cuda") * 1.2).to(dt)
    v = torch.randn(Hk, max_ctx, 128, generator=g, device="cuda").to(dt)
    out = eng.test_kv_attention(q, k, v, pos0, splits)
    torch.cuda.synchronize()
    ref = _reference(q, k, v, pos0, dt)
    emax, el2 = rel_err(out, ref)
    tol = 3e-3 if dt == torch.float16 else 2e-2
    assert emax < tol and el2 < tol, (emax, el2)
    if splits >= 0:      # fixed-order merge: a rerun is bit-identical
        again = eng.test_kv_attention(q, k, v, pos0, splits)
        torch.cuda.synchronize()
        assert torch.equal(out, again)
    eng.close()
