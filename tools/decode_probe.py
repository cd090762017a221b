"""Decode-step probe: LLM-only engine (bf16, Mistral-7B widths, `--layers` deep), prefill `--ctx` tokens, decode `--new`
tokens; prints device ms per token (sm_decode_stats), achieved weight GB/s and the per-phase time of CTA 0
(sm_debug_decode_phases).  Run on the GPU box: python tools/decode_probe.py --layers 32 --ctx 2048"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streammind_b200 import synth
from streammind_b200.engine import Engine, EngineConfig

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--ctx", type=int, default=2048)
ap.add_argument("--new", type=int, default=64)
ap.add_argument("--streams", type=int, default=1)
ap.add_argument("--phases", action="store_true")
a = ap.parse_args()
dt = torch.bfloat16
cfg = EngineConfig(dtype=dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=a.layers, llm_max_ctx=8704, n_streams=a.streams)
eng = Engine(cfg)
dev = torch.device("cuda", 0)
eng.load_state_dict(synth.make_mistral_weights(1234, "", dt, device=dev, layers=a.layers, vocab=cfg.llm_vocab))
eng.finalize()
g = torch.Generator().manual_seed(1)
for s in range(a.streams):
    eng.select_stream(s)
    ids = torch.randint(3, 32000, (a.ctx,), generator=g)
    emb = eng.embed_tokens(ids.cuda())
    for lo in range(0, a.ctx, 512):
        eng.llm_prefill(emb[lo:lo + 512])
sids = list(range(a.streams))
eng.llm_decode_multi(sids, [8] * a.streams)            # warm-up
eng.decode_stats(reset=True)
buf = torch.zeros(8 * 160, dtype=torch.int64, device=dev)      # row 0: CTA 0, rows 1..148: every CTA, row 149: attention sub-phases of CTA 1 (probe build)      # row 0: CTA 0, rows 1..148: every CTA, row 149: attention sub-phases of CTA 1 (probe build)
if a.phases:
    eng.lib.sm_debug_decode_phases(eng._h, C.c_void_p(buf.data_ptr()))
eng.llm_decode_multi(sids, [a.new] * a.streams)
st = eng.decode_stats(reset=True)
per_layer_gb = 218.112e6 * 2 / 1e9
gb = a.layers * per_layer_gb + 131.08e6 * 2 / 1e9
ms_tok = st["ms"] / st["steps"]
kv_gb = st["ctx_sum"] / st["steps"] * a.layers * 4096 / 1e9
print(f"layers {a.layers} ctx {a.ctx} streams {a.streams}: {ms_tok:.4f} ms/step, {1e3 * a.streams / ms_tok:.1f} tok/s aggregate, "
      f"weights {gb / ms_tok * 1e3:.0f} GB/s (+KV {kv_gb / ms_tok * 1e3:.0f} GB/s), launches {eng.launch_count()}")
if a.phases:
    torch.cuda.synchronize()
    # categories 3, 5, 6, 7 are only filled by a -DSMB_DS_WAITPROBE build (tools/gpu/ds_probe_build.sh, SMB_LIB_PATH)
    names = ["prologue", "ring-compute", "epilogue", "producer-blocked", "attention", "producer-life", "wait-first-chunk", "wait-later-chunks"]
    v = buf.cpu().tolist()
    per = torch.tensor(v[8:8 + 8 * 148], dtype=torch.float64).view(148, 8) / st["steps"] / 1e3
    for c, n in ((0, "prologue"), (1, "ring"), (2, "epilogue"), (4, "attention"), (3, "producer-blocked"), (5, "producer-life"),
                 (6, "wait-first-chunk"), (7, "wait-later-chunks")):
        col = per[:, c]
        if float(col.max()) == 0.0:
            continue
        order = col.argsort()
        print(f"  per-CTA {n}: min {col.min():.0f} (cta {int(order[0])}), median {col.median():.0f}, max {col.max():.0f} (cta {int(order[-1])}); "
              f"lowest 5: {[int(x) for x in order[:5]]}, highest 5: {[int(x) for x in order[-5:]]}")
    sub = v[8 * 149:8 * 149 + 6]
    if any(sub):
        print("CTA 1 attention sub-phases (us per step): " + ", ".join(f"{n} {x / st['steps'] / 1e3:.1f}" for n, x in zip(
            ["wait q/k/v + rope", "merge warps + publish partial", "poll partials", "fold + publish", "key blocks (scores, softmax, P V)",
             "row sums + warp partials to smem"], sub)))
    sub = v[8 * 149:8 * 149 + 6]
    if any(sub):
        print("CTA 1 attention sub-phases (us per step): " + ", ".join(f"{n} {x / st['steps'] / 1e3:.1f}" for n, x in zip(
            ["wait q/k/v + rope", "merge warps + publish partial", "poll partials", "fold + publish", "key blocks (scores, softmax, P V)",
             "row sums + warp partials to smem"], sub)))
    tot = sum(v[:8])
    print("CTA 0 phases (us per step): " + ", ".join(f"{n} {x / st['steps'] / 1e3:.1f}" for n, x in zip(names, v)) + f", sum {tot / st['steps'] / 1e3:.1f}")
eng.close()
