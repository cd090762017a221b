"""Per-kernel-class GPU time of one streaming frame, measured with in-stream CUDA events while the GPU is
held back by a spin kernel (so host launch latency cannot leak into the intervals)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streammind_b200 import synth
from streammind_b200.engine import Engine, EngineConfig
dt = torch.float16
B = int(os.environ.get("CHUNK", "1"))
cfg = EngineConfig(dtype=dt, max_frames=B, llm_layers=0, use_graphs=True)
eng = Engine(cfg)
dev = torch.device("cuda", 0)
sd = {}
sd.update(synth.make_vit_weights(1234, dt, device=dev, layers=cfg.vit_layers))
sd.update(synth.make_projector_gate_weights(1234, dt, device=dev))
eng.load_state_dict(sd); eng.finalize(); del sd
frames = synth.make_frames(0, 0, 8 * B, 336, dtype=dt).to(dev)
for t in range(4):
    eng.frame_step(frames[t * B:(t + 1) * B])
torch.cuda.synchronize()
# graph replay time per call
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(5):
    for t in range(8):
        eng.frame_step(frames[t * B:(t + 1) * B], want_device_outputs=False)
e1.record(); torch.cuda.synchronize()
graph_ms = e0.elapsed_time(e1) / 40
eng.profile(True)
tot = {}
NP = 6
for t in range(NP):
    torch.cuda._sleep(60_000_000)          # ~30 ms: the host enqueues the whole frame meanwhile
    _, pooled = eng.vit_encode(frames[t * B:(t + 1) * B], want_feats=False)
    for i in range(B):
        tok = eng.projector_step(pooled[i:i + 1])
        eng.gate_score(tok[0])
    torch.cuda.synchronize()
prof = eng.profile_read()
eng.profile(False)
s = sum(v[0] for v in prof.values()) / NP
print(f"chunk {B}: graph replay {graph_ms*1e3:.1f} us per call ({graph_ms*1e3/B:.1f} us/frame); sum of kernel intervals {s*1e3:.1f} us per call")
for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:26s} {ms/NP*1e3:9.1f} us/call  {n/NP:6.1f} launches  {ms/n*1e3:7.2f} us avg")
