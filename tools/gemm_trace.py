import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streammind_b200.engine import Engine, EngineConfig
dt = torch.float16
eng = Engine(EngineConfig(dtype=dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
buf = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
for name, M, N, K, bn in [("1cta_k64", 128, 128, 64, 128), ("1cta_k64", 128, 128, 64, 32), ("1cta_k1024", 128, 128, 1024, 128),
                          ("qkv", 577, 3072, 1024, 128), ("fc1", 577, 4096, 1024, 256), ("out", 577, 1024, 1024, 64), ("fc2", 577, 1024, 4096, 64)]:
    x = torch.randn(M, K, device="cuda").to(dt); w = torch.randn(N, K, device="cuda").to(dt)
    b = torch.randn(N, device="cuda").to(dt); out = torch.empty(M, N, device="cuda", dtype=dt)
    for _ in range(3): eng.test_gemm(x, w, b, 0, out=out, force_swap=0, force_bn=bn)
    torch.cuda.synchronize()
    eng.lib.sm_test_gemm_trace(eng._h, buf.data_ptr())
    eng.test_gemm(x, w, b, 0, out=out, force_swap=0, force_bn=bn)
    torch.cuda.synchronize()
    eng.lib.sm_test_gemm_trace(eng._h, None)
    ncta = ((M + 127) // 128) * ((N + bn - 1) // bn)
    t = buf[: ncta * 8].view(ncta, 8).cpu().double()
    t0 = t[:, 0].min()
    d = t - t0
    names = ["start", "setup_done", "epi_ready", "accum_done", "epi_loop_done", "epi_end", "exit"]
    print(f"--- {name} M={M} N={N} K={K} bn={bn} ctas={ncta}: kernel span {(t[:,6].max()-t0)/1e3:.2f} us")
    print("   mean ns since first CTA start: " + ", ".join(f"{n}={d[:, i].mean():.0f}" for i, n in enumerate(names)))
    print("   per-CTA phases (mean ns): setup=%.0f mainloop(wait accum)=%.0f epilogue=%.0f tail=%.0f  | cta start spread=%.0f" % (
        (t[:, 1] - t[:, 0]).mean(), (t[:, 3] - t[:, 1]).mean(), (t[:, 5] - t[:, 3]).mean(), (t[:, 6] - t[:, 5]).mean(), d[:, 0].max()))
