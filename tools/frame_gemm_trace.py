"""Timeline of the GEMM launches inside one real streaming frame (weights cold in HBM): per launch
[first CTA start, last CTA exit] from in-kernel globaltimer stamps, plus the gaps between launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streammind_b200 import synth
from streammind_b200.engine import Engine, EngineConfig
dt = torch.float16
cfg = EngineConfig(dtype=dt, max_frames=1, llm_layers=0, use_graphs=False)
eng = Engine(cfg)
dev = torch.device("cuda", 0)
sd = {}
sd.update(synth.make_vit_weights(1234, dt, device=dev, layers=cfg.vit_layers))
sd.update(synth.make_projector_gate_weights(1234, dt, device=dev))
eng.load_state_dict(sd); eng.finalize(); del sd
frames = synth.make_frames(0, 0, 4, 336, dtype=dt).to(dev)
for t in range(3):
    eng.frame_step(frames[t:t + 1])
torch.cuda.synchronize()
NL = 93
buf = torch.zeros(NL * 8 + 64, dtype=torch.int64, device="cuda")
buf.view(-1, 8)[:, 0] = 2 ** 62
torch.cuda._sleep(60_000_000)
eng.lib.sm_test_gemm_trace(eng._h, buf.data_ptr())
eng.frame_step(frames[3:4])
torch.cuda.synchronize()
eng.lib.sm_test_gemm_trace(eng._h, None)
t = buf[: NL * 8].view(NL, 8).cpu().double()
t0 = t[0, 0]
names = {0: "patch", 1: "qkv", 2: "out", 3: "fc1", 0.5: ""}
print("idx kind   start_us  dur_us  setup  mainloop  epilogue | gap_before_us")
prev_end = None
acc = {}
for i in range(NL):
    kind = "patch" if i == 0 else ["qkv", "out", "fc1", "fc2"][(i - 1) % 4]
    st, setup, accum, epi_end, ex = t[i, 0], t[i, 1], t[i, 3], t[i, 5], t[i, 6]
    dur = (ex - st) / 1e3
    gap = (st - prev_end) / 1e3 if prev_end is not None else 0.0
    prev_end = ex
    a = acc.setdefault(kind, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += dur; a[2] += (setup - st) / 1e3; a[3] += (accum - setup) / 1e3; a[4] += (epi_end - accum) / 1e3; a[5] += gap
    if i < 9 or i > 88:
        print(f"{i:3d} {kind:6s} {(st-t0)/1e3:8.1f} {dur:7.2f} {(setup-st)/1e3:6.2f} {(accum-setup)/1e3:8.2f} {(epi_end-accum)/1e3:8.2f} | {gap:6.2f}")
print("mean per kind: kind n dur setup mainloop epilogue gap_before")
for k, a in acc.items():
    n = a[0]
    print(f"  {k:6s} {n:3d} {a[1]/n:7.2f} {a[2]/n:6.2f} {a[3]/n:8.2f} {a[4]/n:8.2f} {a[5]/n:7.2f}")
print("frame span of GEMM launches: %.1f us; sum of GEMM durations %.1f us" % ((t[NL-1, 6] - t0) / 1e3, sum(a[1] for a in acc.values())))
