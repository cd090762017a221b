"""Prefill probe: LLM-only engine (bf16, Mistral-7B widths), context `--ctx` in the cache, then a prefill of `--new` positions timed with
CUDA events (the cache is rewound between repetitions), with the per-class kernel profile of one call.
Run on the GPU box: python tools/prefill_probe.py --ctx 8000 --new 11"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streammind_b200 import synth
from streammind_b200.engine import Engine, EngineConfig

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--ctx", type=int, nargs="+", default=[2048])
ap.add_argument("--new", type=int, nargs="+", default=[11, 30])
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dt = torch.bfloat16
cfg = EngineConfig(dtype=dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=a.layers, llm_max_ctx=8704)
eng = Engine(cfg)
dev = torch.device("cuda", 0)
eng.load_state_dict(synth.make_mistral_weights(1234, "", dt, device=dev, layers=a.layers, vocab=cfg.llm_vocab))
eng.finalize()
g = torch.Generator().manual_seed(1)
ids = torch.randint(3, 32000, (8704,), generator=g)
emb = eng.embed_tokens(ids.cuda())
have = 0
for ctx in sorted(a.ctx):
    for lo in range(have, ctx, 512):
        eng.llm_prefill(emb[lo:min(ctx, lo + 512)])
    have = ctx
    for n in a.new:
        if ctx + n > 8704:
            continue
        eng.llm_prefill(emb[ctx:ctx + n]); eng.kv_set_len(ctx)        # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            eng.llm_prefill(emb[ctx:ctx + n]); eng.kv_set_len(ctx)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        eng.profile(True)
        eng.llm_prefill(emb[ctx:ctx + n]); eng.kv_set_len(ctx)
        torch.cuda.synchronize()
        prof = eng.profile_read()
        eng.profile(False)
        cls = ", ".join(f"{k} {v[0]:.3f} ms / {v[1]}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]) if v[1])
        print(f"ctx {ctx} + {n} positions: {ms:.3f} ms per prefill (weight pass at the HBM peak: {14.22e9 / 6451.5e9 * 1e3 * a.layers / 32:.2f} ms); serialised classes: {cls}", flush=True)
eng.close()
