"""clock64 trace of one mid-grid CTA of the tcgen05 attention kernel (AttnTcArgs::dbg), batch 8 of the ViT shape.
Needs a library built with -DSMB_ATC_TRACE (the stamps are compiled out of the product build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streammind_b200.engine import Engine, EngineConfig
dt = torch.float16
eng = Engine(EngineConfig(dtype=dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
B, S, H, D = 8, 577, 16, 64
qkv = (torch.randn(B * S, 3 * H * D, device="cuda") * 1.5).to(dt)
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
for _ in range(3): eng.test_attention(qkv, B, S, H, D, 2)
torch.cuda.synchronize()
eng.lib.sm_test_gemm_trace(eng._h, buf.data_ptr())
eng.test_attention(qkv, B, S, H, D, 2)
torch.cuda.synchronize()
eng.lib.sm_test_gemm_trace(eng._h, None)
t = buf.cpu().view(8, 64)
t0 = int(t[7, 0])
G = 10
rel = lambda v: int(v) - t0
print("CTA start -> softmax end:", rel(t[7, 60]), "cycles")
print(" g | sm wait  s_ready  done | mma: enter  k_ok  issue_S | pv: enter  issue")
for g in range(G):
    pv = f"{rel(t[7, 1 + g]):7d} {rel(t[7, 24 + g]):7d}" if True else ""
    print(f"{g:2d} | {rel(t[0, g]):7d} {rel(t[1, g]):7d} {rel(t[2, g]):7d} | {rel(t[3, g]):7d} {rel(t[4, g]):7d} {rel(t[5, g]):7d} | {pv}")
print("producer k_empty passed (per box):", [rel(t[6, i]) for i in range(5)])
eng.close()
