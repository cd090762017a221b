"""Times the tcgen05 GEMM over tile configurations (CUDA events, warm L2, 50 iterations each)."""
import itertools, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streammind_b200.engine import Engine, EngineConfig

dt = torch.float16
eng = Engine(EngineConfig(dtype=dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
shapes = {"tiny": (128, 128, 64), "tinyK1024": (128, 128, 1024), "qkv": (577, 3072, 1024), "out": (577, 1024, 1024),
          "fc1": (577, 4096, 1024), "fc2": (577, 1024, 4096), "qkv8": (4616, 3072, 1024), "fc2_8": (4616, 1024, 4096),
          "fc1_8": (4616, 4096, 1024), "out8": (4616, 1024, 1024)}
res = {}
for name, (M, N, K) in shapes.items():
    x = torch.randn(M, K, device="cuda").to(dt)
    w = torch.randn(N, K, device="cuda").to(dt)
    b = torch.randn(N, device="cuda").to(dt)
    out = torch.empty(M, N, device="cuda", dtype=dt)
    cfgs = [(0, bn) for bn in (32, 64, 128, 256) if bn <= N] + [(1, bn) for bn in (16, 32, 64, 96, 128, 160, 192, 208, 256)]
    for swap, bn in cfgs:
        try:
            for _ in range(3):
                eng.test_gemm(x, w, b, 0, out=out, force_swap=swap, force_bn=bn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                eng.test_gemm(x, w, b, 0, out=out, force_swap=swap, force_bn=bn)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 50
            tf = 2 * M * N * K / us / 1e6
            res[f"{name}/swap{swap}/bn{bn}"] = (round(us, 2), round(tf, 1))
            print(f"{name:10s} M={M} N={N} K={K} swap={swap} bn={bn:3d}: {us:8.2f} us  {tf:7.1f} TFLOP/s", flush=True)
        except Exception as ex:
            print(name, swap, bn, "ERR", ex)
# launch-overhead reference: an empty-ish torch kernel
x = torch.zeros(1, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): x.add_(1)
e1.record(); torch.cuda.synchronize()
print("torch tiny kernel back-to-back:", e0.elapsed_time(e1) * 1e3 / 200, "us")
json.dump(res, open("gpurun_out/gemm_sweep.json", "w"), indent=1)
