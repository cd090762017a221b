import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streammind_b200.engine import Engine, EngineConfig
dt = torch.float16
eng = Engine(EngineConfig(dtype=dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
NREP = 40
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
M, N, K = 577, 3072, 1024
x = torch.randn(M, K, device="cuda").to(dt); w = torch.randn(N, K, device="cuda").to(dt)
b = torch.randn(N, device="cuda").to(dt); out = torch.empty(M, N, device="cuda", dtype=dt)
small = torch.zeros(577 * 1024, device="cuda", dtype=dt)
qkv = torch.randn(577, 3072, device="cuda").to(dt)
gemm = lambda: eng.test_gemm(x, w, b, 0, out=out, force_swap=0, force_bn=128)
tiny = lambda: small.add_(1)
attn = lambda: eng.test_attention(qkv, 1, 577, 16, 64)
t_g = graph_time(gemm); t_t = graph_time(tiny); t_a = graph_time(attn)
t_gt = graph_time(lambda: (gemm(), tiny()))
t_ga = graph_time(lambda: (gemm(), attn()))
t_gat = graph_time(lambda: (gemm(), attn(), tiny()))
print(f"pdl={'off' if os.environ.get('SMB_NO_PDL') else 'on'} gemm {t_g:.2f}  tiny {t_t:.2f}  attn {t_a:.2f}  gemm+tiny {t_gt:.2f} (sum {t_g+t_t:.2f})  gemm+attn {t_ga:.2f} (sum {t_g+t_a:.2f})  gemm+attn+tiny {t_gat:.2f} (sum {t_g+t_a+t_t:.2f})")
