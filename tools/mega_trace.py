"""Per-op timeline of the persistent vision-tower kernel for one streaming frame (B = CHUNK)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streammind_b200 import synth
from streammind_b200.engine import Engine, EngineConfig
dt = torch.float16
B = int(os.environ.get("CHUNK", "1"))
cfg = EngineConfig(dtype=dt, max_frames=B, llm_layers=0, use_graphs=False)
eng = Engine(cfg)
dev = torch.device("cuda", 0)
sd = {}
sd.update(synth.make_vit_weights(1234, dt, device=dev, layers=cfg.vit_layers))
sd.update(synth.make_projector_gate_weights(1234, dt, device=dev))
eng.load_state_dict(sd); eng.finalize(); del sd
frames = synth.make_frames(0, 0, 4 * B, 336, dtype=dt).to(dev)
for t in range(3):
    eng.frame_step(frames[t * B:(t + 1) * B])
torch.cuda.synchronize()
n = C.c_int(0); types = (C.c_int * 512)()
buf = torch.zeros(512 * 4, dtype=torch.int64, device="cuda")
eng.lib.sm_debug_mega_trace(eng._h, buf.data_ptr(), B, C.byref(n), types, 512)
eng.frame_step(frames[3 * B:4 * B])
torch.cuda.synchronize()
eng.lib.sm_debug_mega_trace(eng._h, None, B, None, None, 0)
t = buf.view(-1, 4).cpu().double()[: n.value] / 1e3
names = ["gemm", "splitk_ln", "embed_ln", "attn", "pool", "im2col"]
t0 = t[0, 0]
print("op type        start   work_done  arrived | wait_for_prev  work   fence+arrive")
acc = {}
prev = None
for i in range(n.value):
    nm = names[types[i]]
    st, wd, ar = t[i, 0] - t0, t[i, 1] - t0, t[i, 2] - t0
    wait = (st - prev) if prev is not None else 0.0
    prev = ar
    # layer-local index for gemm ops
    key = nm
    if nm == "gemm":
        j = (i - 3) % 7 if i >= 3 else -1
        key = {-1: "gemm_patch", 0: "gemm_qkv", 2: "gemm_out", 4: "gemm_fc1", 5: "gemm_fc2"}.get(j, "gemm?") if i != 1 else "gemm_patch"
    a = acc.setdefault(key, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += wait; a[2] += wd - st; a[3] += ar - wd
    if i < 12 or i > n.value - 4:
        print(f"{i:3d} {key:11s} {st:8.1f} {wd:8.1f} {ar:8.1f} | {wait:6.2f} {wd-st:7.2f} {ar-wd:6.2f}")
print("mean per kind (us): n, barrier latency (last arrive -> last start), work (max over CTAs), fence+arrive")
tot = 0.0
for k, a in acc.items():
    print(f"  {k:11s} {a[0]:3d} {a[1]/a[0]:7.2f} {a[2]/a[0]:7.2f} {a[3]/a[0]:7.2f}   total {sum(a[1:]):8.1f}")
    tot += sum(a[1:])
print(f"kernel span {t[n.value-1,2]-t0:.1f} us; sum {tot:.1f}")
