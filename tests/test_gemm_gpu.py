"""tcgen05 GEMM kernel vs torch fp32 matmul on the same fp16/bf16 inputs (bit-level layout check:
a wrong UMMA/TMA descriptor gives O(1) errors, a right one gives pure rounding error)."""
import pytest
import torch

from parity_util import engine_config, rel_err

pytestmark = pytest.mark.gpu

EPI_STORE, EPI_GELU, EPI_RESID, EPI_F32 = 0, 1, 2, 3


@pytest.fixture(scope="module")
def eng(built_library):
    from streammind_b200.engine import Engine
    cfgs = {}
    for dt in (torch.float16, torch.bfloat16):
        c = engine_config(dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0)
        cfgs[dt] = Engine(c)
    yield cfgs
    for e in cfgs.values():
        e.close()


def _ref(x, w, b):
    y = x.float() @ w.float().t()
    if b is not None:
        y = y + b.float()
    return y


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K,swap,bn", [
    (128, 128, 64, 0, 128), (128, 128, 64, 1, 128), (128, 256, 256, 0, 64), (577, 1024, 1024, 0, 128),
    (577, 3072, 1024, 0, 256), (577, 1024, 4096, 1, 160), (576, 1024, 640, 0, 128), (37, 512, 1024, 1, 48),
    (1, 4096, 4096, 1, 16), (1154, 4096, 1024, -1, 0), (64, 6144, 4096, -1, 0), (300, 1024, 1000, 0, 32),
    # swap = 2: 256 x 256 tile, two TMEM accumulators sharing each weight stage (GemmArgs::bm2)
    (577, 1024, 1024, 2, 256), (577, 3072, 1024, 2, 256), (1154, 4096, 1024, 2, 256), (300, 512, 384, 2, 256),
    (257, 256, 4096, 2, 256),
])
def test_gemm_store(eng, dt, M, N, K, swap, bn):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, generator=g, device="cuda").to(dt)
    w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).to(dt)
    b = torch.randn(N, generator=g, device="cuda").to(dt)
    out = eng[dt].test_gemm(x, w, b, EPI_STORE, force_swap=swap, force_bn=bn)
    torch.cuda.synchronize()
    ref = _ref(x, w, b)
    emax, el2 = rel_err(out, ref)
    tol = 2e-3 if dt == torch.float16 else 1.2e-2        # one rounding of the output to T
    assert emax < tol and el2 < tol, (emax, el2)


@pytest.mark.parametrize("swap", [0, 1, 2])
def test_gemm_epilogues(eng, swap):
    dt = torch.float16
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 200, 512, 384
    x = torch.randn(M, K, generator=g, device="cuda").to(dt)
    w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).to(dt)
    b = torch.randn(N, generator=g, device="cuda").to(dt)
    ref = _ref(x, w, b)
    # quick_gelu with the reference's rounding points
    h = ref.to(dt)
    gel = (h * torch.sigmoid(1.702 * h)).float()
    out = eng[dt].test_gemm(x, w, b, EPI_GELU, force_swap=swap)
    assert rel_err(out, gel)[0] < 3e-3
    # residual, in place
    res0 = torch.randn(M, N, generator=g, device="cuda").to(dt)
    out = eng[dt].test_gemm(x, w, b, EPI_RESID, out=res0.clone(), force_swap=swap)
    assert rel_err(out, res0.float() + ref.to(dt).float())[0] < 3e-3
    # fp32 store
    out = eng[dt].test_gemm(x, w, b, EPI_F32, force_swap=swap)
    assert out.dtype == torch.float32 and rel_err(out, ref)[0] < 1e-4
