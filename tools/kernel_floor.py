"""True GPU time per launch of the GEMM for several shapes: N back-to-back launches captured in a CUDA
graph (no host launch overhead), replayed and timed with CUDA events."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streammind_b200.engine import Engine, EngineConfig

dt = torch.float16
eng = Engine(EngineConfig(dtype=dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
NREP = 40

def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(NREP):
                fn()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * NREP)

x1 = torch.zeros(8, device="cuda")
print(f"torch tiny add_ in graph: {graph_time(lambda: x1.add_(1)):.2f} us", flush=True)
shapes = [("1cta_k64", 128, 128, 64), ("1cta_k1024", 128, 128, 1024), ("1cta_k4096", 128, 128, 4096),
          ("120cta_k64", 577, 3072, 64), ("qkv", 577, 3072, 1024), ("out", 577, 1024, 1024),
          ("fc1", 577, 4096, 1024), ("fc2", 577, 1024, 4096)]
for name, M, N, K in shapes:
    x = torch.randn(M, K, device="cuda").to(dt); w = torch.randn(N, K, device="cuda").to(dt)
    b = torch.randn(N, device="cuda").to(dt); out = torch.empty(M, N, device="cuda", dtype=dt)
    for swap, bn in ((0, 32), (0, 64), (0, 128), (0, 256), (1, 96), (1, 160), (1, 208)):
        if bn > N: continue
        us = graph_time(lambda: eng.test_gemm(x, w, b, 0, out=out, force_swap=swap, force_bn=bn))
        print(f"{name:12s} M={M} N={N} K={K} swap={swap} bn={bn:3d}: {us:7.2f} us {2*M*N*K/us/1e6:7.1f} TF", flush=True)
