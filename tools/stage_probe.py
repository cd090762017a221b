import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streammind_b200.engine import Engine, EngineConfig
dt = torch.float16
eng = Engine(EngineConfig(dtype=dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
NREP = 20
def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn(); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(NREP): fn()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5): g.replay()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * NREP)
for name, M, N, K in [("1cta_k4096", 128, 128, 4096), ("fc2", 577, 1024, 4096), ("fc2_8", 4616, 1024, 4096)]:
    x = torch.randn(M, K, device="cuda").to(dt); w = torch.randn(N, K, device="cuda").to(dt)
    b = torch.randn(N, device="cuda").to(dt); out = torch.empty(M, N, device="cuda", dtype=dt)
    for swap, bn in ((0, 32), (0, 128), (0, 256)):
        if bn > N: continue
        us = graph_time(lambda: eng.test_gemm(x, w, b, 0, out=out, force_swap=swap, force_bn=bn))
        print(f"stages={os.environ.get('SMB_GEMM_STAGES','max')} {name:12s} bn={bn:3d}: {us:7.2f} us  per-kblock {us*1e3/(K/64):6.1f} ns", flush=True)
