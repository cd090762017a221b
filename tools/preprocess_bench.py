"""Device time of sm_preprocess_frames on 1080p frames (batches of 1, 4, 16) -- CUDA events, warm, inputs resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from streammind_b200.engine import Engine, EngineConfig
from preprocess_cases import make_frame
eng = Engine(EngineConfig(dtype=torch.float16, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=0))
mb = (1080 * 1920 * 3 + 2 * 1920 * 336 * 3 + 3 * 336 * 336 * 2) / 1e6
for n in (1, 4, 16):
    dev = torch.from_numpy(np.stack([make_frame(1080, 1920, i) for i in range(n)])).cuda()
    for _ in range(3): eng.preprocess_frames(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps): eng.preprocess_frames(dev)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * n)
    print(f"batch {n:2d}: {us:6.1f} us per 1080p frame, {mb / us * 1e3:6.0f} GB/s of {mb:.1f} MB algorithmic traffic per frame")
eng.close()
