"""Graph-timed sweep of (swap, bn) for the ViT GEMM shapes; prints the best configs per shape."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streammind_b200.engine import Engine, EngineConfig
dt = torch.float16
eng = Engine(EngineConfig(dtype=dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
NREP = 20
def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(NREP): fn()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(3): g.replay()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * NREP)
res = {}
for B in (1, 2, 4, 8, 16):
    T = 577 * B
    for name, N, K, epi in (("patch", 1024, 640, 0), ("qkv", 3072, 1024, 0), ("out", 1024, 1024, 2), ("fc1", 4096, 1024, 1), ("fc2", 1024, 4096, 2)):
        M = 576 * B if name == "patch" else T
        x = torch.randn(M, K, device="cuda").to(dt); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(dt)
        b = torch.randn(N, device="cuda").to(dt); out = torch.zeros(M, N, device="cuda", dtype=dt)
        rows = []
        cands = [(0, bn) for bn in (32, 64, 128, 256)] + [(1, bn) for bn in range(16, 257, 16)]
        for swap, bn in cands:
            ctas = ((N + 127) // 128) * ((M + bn - 1) // bn) if swap else ((M + 127) // 128) * ((N + bn - 1) // bn)
            if ctas > 148 * 12 or (B > 2 and bn < 64): continue
            us = graph_time(lambda: eng.test_gemm(x, w, b, epi, out=out, force_swap=swap, force_bn=bn))
            rows.append((us, swap, bn, ctas))
        rows.sort()
        res[f"B{B}/{name}"] = rows[:4]
        print(f"B={B:2d} {name:5s} M={M} N={N} K={K}: " + "  ".join(f"[{'swap' if s else 'norm'} bn={bn} {us:.2f}us {c}cta]" for us, s, bn, c in rows[:4]), flush=True)
json.dump(res, open("gpurun_out/gemm_plan_sweep.json", "w"))
