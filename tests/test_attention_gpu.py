"""Fused attention kernel vs torch (fp32 softmax, probabilities rounded to T before P@V)."""
import pytest
import torch

from parity_util import engine_config, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,S,H,D", [(1, 577, 16, 64), (3, 17, 2, 64), (2, 64, 4, 64), (1, 130, 2, 128),
                                     (2, 128, 2, 64), (3, 577, 16, 64), (2, 257, 3, 64), (8, 577, 16, 64)])
@pytest.mark.parametrize("mode", [0, 2])     # 0 = mma.sync kernel, 2 = tcgen05 kernel (d = 64 only; d = 128 falls back)
def test_attention_noncausal(built_library, dt, B, S, H, D, mode):
    from streammind_b200.engine import Engine
    eng = Engine(engine_config(dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
    g = torch.Generator(device="cuda").manual_seed(S + H)
    C = H * D
    qkv = (torch.randn(B * S, 3 * C, generator=g, device="cuda") * 1.5).to(dt)
    out = eng.test_attention(qkv, B, S, H, D, mode)
    torch.cuda.synchronize()
    q, k, v = [t.view(B, S, H, D).transpose(1, 2).float() for t in qkv.view(B * S, 3, C).unbind(1)]
    p = torch.softmax(q @ k.transpose(-1, -2) * D ** -0.5, dim=-1).to(dt).float()
    ref = (p @ v).transpose(1, 2).reshape(B * S, C)
    emax, el2 = rel_err(out, ref)
    tol = 3e-3 if dt == torch.float16 else 2e-2
    assert emax < tol and el2 < tol, (emax, el2)
    eng.close()


@pytest.mark.parametrize("mode", [2])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_attention_tc_batch_invariant(built_library, dt, mode):
    """tcgen05 kernel: a frame's output does not depend on its batch neighbours (the last key tile of a frame reads the
    next frame's rows and masks them), and a rerun is bit-identical."""
    from streammind_b200.engine import Engine
    eng = Engine(engine_config(dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
    B, S, H, D = 8, 577, 16, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = (torch.randn(B * S, 3 * H * D, generator=g, device="cuda") * 3.0).to(dt)
    out8 = eng.test_attention(qkv, B, S, H, D, mode).clone()
    again = eng.test_attention(qkv, B, S, H, D, mode).clone()
    torch.cuda.synchronize()
    assert torch.equal(out8, again)
    for b in (0, 3, 7):
        o1 = eng.test_attention(qkv[b * S:(b + 1) * S].contiguous(), 1, S, H, D, mode)
        torch.cuda.synchronize()
        assert torch.equal(o1, out8[b * S:(b + 1) * S]), b
    eng.close()
