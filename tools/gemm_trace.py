import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streammind_b200.engine import Engine, EngineConfig
dt = torch.float16
eng = Engine(EngineConfig(dtype=dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
CASES = [("qkv", 577, 3072, 1024, 128, 0), ("fc1", 577, 4096, 1024, 256, 0), ("fc1", 577, 4096, 1024, 160, 1),
         ("fc1", 577, 4096, 1024, 96, 1), ("1cta", 128, 128, 1024, 96, 1), ("1cta", 128, 128, 1024, 128, 0)]
for name, M, N, K, bn, swap in CASES:
    x = torch.randn(M, K, device="cuda").to(dt); w = torch.randn(N, K, device="cuda").to(dt)
    b = torch.randn(N, device="cuda").to(dt); out = torch.empty(M, N, device="cuda", dtype=dt)
    for _ in range(3): eng.test_gemm(x, w, b, 0, out=out, force_swap=swap, force_bn=bn)
    torch.cuda.synchronize()
    buf.zero_(); buf[0] = 2 ** 62
    eng.lib.sm_test_gemm_trace(eng._h, buf.data_ptr())
    eng.test_gemm(x, w, b, 0, out=out, force_swap=swap, force_bn=bn)
    torch.cuda.synchronize()
    eng.lib.sm_test_gemm_trace(eng._h, None)
    t = buf[:8].cpu().double()
    print(f"--- {name} M={M} N={N} K={K} bn={bn} swap={swap}: span {(t[6]-t[0])/1e3:.2f} us | setup_end {(t[1]-t[0])/1e3:.2f} accum_done {(t[3]-t[0])/1e3:.2f} epi_loop_done {(t[4]-t[0])/1e3:.2f} epi_end {(t[5]-t[0])/1e3:.2f}")
