#!/usr/bin/env python
"""Benchmark of the StreamMind per-frame hot path on B200 (contract: see the task statement / DESIGN.md section 6).

Default workload = BASELINE.json configs[2], the configuration the metric is quoted on: one synthetic 336x336 stream per GPU
through CLIP-ViT-L/14-336 encode -> Mamba projector step -> event gate -> (on a fire) Mistral-7B prefill of the new
dialogue suffix + greedy decode; bf16, random-init weights, 256 frames, a fire on every 16th frame (gate decision
overridden, logits still computed), 224 new tokens per fire, KV cache growing to ~4.1k positions.  Everything goes through
the reference-facing API (StreamMindB200ForCausalLM.stream_generate_demo, one frame per call); frames are submitted
ahead through the pipelined path, so the frames after a fire are encoded while it decodes.

  one STEP  = one 16-frame gate interval: 15 silent frames + 1 firing frame with its prefill and 224-token decode.  Steps
              walk the intervals of the stream in order (step k = interval k mod 16; the stream restarts after 16 steps =
              one whole configs[2] stream), so --steps 16 times exactly one configs[2] stream.
  value     : frames/s over all ranks, frames already resident in HBM when the timed region starts (CUDA events).
  e2e       : the same calls with the frames in PINNED HOST memory: H2D copy of every frame, D2H of every gate decision
              and of the generated ids inside the timed region.
  roofline  : the persistent decode kernel (one launch per token): algorithmic bytes per launch (14.221 GB of weights +
              128 KiB x context of KV) / the mean launch duration measured live with CUDA events on the launch stream
              (sm_decode_stats) against the measured HBM copy peak.
  cpu_baseline / --impl reference : the oracle port of the reference's CPU path on the host cores, bounded sample.

N > 1: one process per GPU (torchrun), one independent stream per rank, no data-path collective; NCCL only
for the barrier and the max over ranks of the device time.  Other workloads: --workload frames (configs[1]),
dense_decode (configs[4]), --streams B (multi-stream batching on one GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "streaming frames/sec (encode+gate+decode) @336px, Mistral-7B, 1/2/4/8 B200"   # BASELINE.json, verbatim

# algorithmic work
VIT_GFLOP_PER_FRAME = 366.0
GEMM_GFLOP_PER_FRAME = 334.65            # the tower's GEMM share
GATE_MB_PER_FRAME = 1577.1
PROJ_MB_PER_FRAME = 254.8
VIT_WEIGHT_MB = 578.8
DECODE_BYTES_PER_TOKEN = 14.221e9
KV_BYTES_PER_POSITION = 131072.0


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def _tracked_traffic(kernel: str):
    """dram bytes per launch of `kernel` from the tracked ncu summary (profiles/r02_ncu_traffic.json, written by
    tools/ncu_summary.py from a --set full capture of the same launches); None when no capture is tracked."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    e = d.get(kernel)
    return (e.get("dram_bytes_per_launch"), e.get("source")) if e else (None, None)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
# workload policies
# ------------------------------------------------------------------------------------------------------------------
POLICY = {
    "gated_decode": dict(n_frames=256, fire_every=16, max_new=224,
                         label="BASELINE configs[2]: 256-frame stream, fire every 16th frame, 224 greedy tokens per fire, KV -> ~4.1k"),
    "dense_decode": dict(n_frames=512, fire_every=1, max_new=5,
                         label="BASELINE configs[4]: gate-off dense mode, 512 frames, every frame fires, 5 greedy tokens, KV -> ~8.3k"),
}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------------------------
class CpuReference:
    """oracle/restate.py (the reference's algorithm, pinned against the reference's own outputs) at the BASELINE sizes on
    the host, fp32, all threads.  One sample step = 1 frame through ViT -> projector (re-run over all frames so far, as the
    reference does: videollama2_arch.py:190-198) -> gate, then one fire's LLM work in miniature: a 26-position prefill
    (the suffix a later fire of configs[2] adds: 10 template ids + 16 frame tokens) and 2 decode steps at that context.
    The LLM is Mistral-7B-shaped with ONE set of layer weights aliased 32 times (872 MB per layer, far beyond any cache, so
    timing-equivalent to 32 distinct layers while the host only holds 3 GB instead of 28 GB)."""

    def __init__(self, threads: int):
        import torch
        from oracle import restate as R
        from streammind_b200 import synth
        torch.set_num_threads(threads)
        self.torch, self.R, self.synth, self.threads = torch, R, synth, threads
        sd = {}
        sd.update(synth.make_vit_weights(1234, torch.float32))
        sd.update(synth.make_projector_gate_weights(1234, torch.float32, gate_with_qk=True))
        one = synth.make_mistral_weights(1234, "", torch.float32, layers=1, vocab=32002)
        for k, v in one.items():
            if k.startswith("model.layers.0."):
                for l in range(32):
                    sd[k.replace("model.layers.0.", f"model.layers.{l}.")] = v
            else:
                sd[k] = v
        self.sd = sd
        self.vit, self.mam, self.gate, self.llm = R.VitConfig(), R.MambaCfg(), R.gate_config(), R.MistralCfg()
        self.frames = synth.make_frames(0, 0, 4, 336, dtype=torch.float32)
        self.feats = None
        self.t = 0

    def sample_step(self):
        torch, R = self.torch, self.R
        with torch.no_grad():
            t0 = time.perf_counter()
            f = R.clip_vision_tower(self.sd, self.vit, self.frames[self.t % 4:self.t % 4 + 1]).unsqueeze(0)
            self.feats = f if self.feats is None or self.feats.shape[1] >= 4 else torch.cat([self.feats, f], dim=1)
            x = R.projector_sequence(self.sd, self.mam, self.feats)
            R.gate_decision(R.gate_logits(self.sd, self.gate, x[0, -1]))
            t1 = time.perf_counter()
            cache = R.KVCache()
            emb = torch.randn(26, self.llm.hidden_size)
            lg = R.mistral_forward(self.sd, "", self.llm, emb, cache)
            t2 = time.perf_counter()
            for _ in range(2):
                tok = int(lg.argmax())
                lg = R.mistral_forward(self.sd, "", self.llm, self.sd["model.embed_tokens.weight"][tok][None, :], cache)
            t3 = time.perf_counter()
        self.t += 1
        return dict(frame_s=t1 - t0, prefill26_s=t2 - t1, token_s=(t3 - t2) / 2, wall_s=t3 - t0)

    @staticmethod
    def interval_seconds(s, fire_every=16, max_new=224):
        """seconds the CPU path needs for one step of the workload (fire_every frames + one fire), from the sample's unit costs"""
        return fire_every * s["frame_s"] + s["prefill26_s"] + (max_new - 1) * s["token_s"]

    SAMPLE = ("per step: 1 frame (ViT-L/14-336 + projector re-run over the history + 4-layer gate) + a 26-position Mistral-7B "
              "prefill + 2 decode steps, fp32, oracle/restate.py on all host threads; LLM layer weights aliased 32x; the "
              "metric value is 16 frames / (16 t_frame + t_prefill + 223 t_token)")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pol = POLICY["gated_decode"]
    threads = os.cpu_count() or 1
    ref = CpuReference(threads)
    for _ in range(max(1, min(args.warmup, 2))):
        ref.sample_step()
    samples = [ref.sample_step() for _ in range(args.steps)]
    mean = {k: statistics.mean(s[k] for s in samples) for k in samples[0]}
    sec = CpuReference.interval_seconds(mean, pol["fire_every"], pol["max_new"])
    fps = pol["fire_every"] / sec
    emit({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": pol["label"] + " -- on the host cores, bounded sample per step",
                   "sample_wall_ms_per_step": 1e3 * mean["wall_s"], "unit_costs_s": mean},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": CpuReference.SAMPLE},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def reference_gpu_vit(dev, dt, frames_dev, n=24):
    """Secondary figure: the reference's own GPU path for the vision tower -- hf CLIPVisionModel (24 layers, hidden states kept,
    what CLIPVisionTower.forward runs: clip_encoder.py:41-53) executed by PyTorch on this GPU in the model dtype, one frame per
    call as in the streaming demo.  Timed with CUDA events after the main measurement; not part of `value`."""
    import torch
    try:
        from streammind_b200 import hf_reference, synth
        sd = synth.make_vit_weights(1234, dt, device=dev, layers=24)
        m = hf_reference.build_hf_clip(sd, 1024, 4096, 24, 16, 336, 14, 1e-5, dt, device=dev)
        del sd
        for i in range(5):
            hf_reference.clip_features(m, frames_dev[i:i + 1])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for i in range(n):
            hf_reference.clip_features(m, frames_dev[i % frames_dev.shape[0]:i % frames_dev.shape[0] + 1])
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        del m
        torch.cuda.empty_cache()
        return {"frames_per_s": 1e3 / ms, "ms_per_frame": ms, "frames_timed": n,
                "note": "hf CLIPVisionModel (transformers, PyTorch kernels) on the same GPU, B = 1 per call, model dtype, ViT encode only -- "
                        "compare with configs1_frames_stage.serial_b1_frames_per_s (which also runs the projector and the gate)"}
    except Exception as e:                                    # transformers missing / out of memory: the key says so
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


# ------------------------------------------------------------------------------------------------------------------
def _load_full_model(model_or_engine, cfg, dev, dt, seed=1234):
    """random-init weights of the full configuration, generated on the device part by part (17 GB in bf16)"""
    import torch
    from streammind_b200 import synth
    eng = getattr(model_or_engine, "engine", model_or_engine)
    for part in (synth.make_vit_weights(seed, dt, device=dev, layers=cfg.vit_layers),
                 synth.make_projector_gate_weights(seed, dt, device=dev)):
        eng.load_state_dict(part)
        del part
    if cfg.llm_layers > 0:
        full = synth.make_mistral_weights(seed, "", dt, device=dev, layers=cfg.llm_layers, vocab=cfg.llm_vocab)
        eng.load_state_dict(full)
        del full
    torch.cuda.empty_cache()
    eng.finalize()
    for s in range(cfg.n_streams):
        eng.select_stream(s)
        eng.reset_stream()
    eng.select_stream(0)


def run_stream_workload(args):
    """configs[2] / configs[4] through StreamMindB200ForCausalLM.stream_generate_demo."""
    import torch
    from streammind_b200 import dist_util, synth
    from streammind_b200.engine import Engine, EngineConfig
    from streammind_b200.model import StreamMindB200ForCausalLM
    rank, world, local = dist_util.env_rank_world()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a GPU; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_util.init("nccl", dev)
    pol = POLICY[args.workload]
    dt = torch.bfloat16
    n_frames, fire_every, max_new = pol["n_frames"], pol["fire_every"], pol["max_new"]
    intervals = n_frames // fire_every
    lookahead = 0 if args.no_prefetch else min(14, args.lookahead)     # + the current frame = 15 tickets in flight (ring of 16)
    cfg = EngineConfig(dtype=dt, max_frames=1, llm_max_ctx=8704, use_graphs=not args.no_graphs)
    model = StreamMindB200ForCausalLM(cfg, None, device=local)
    _load_full_model(model, cfg, dev, dt)
    eng = model.engine
    prompt0, turn_suffix = synth.make_prompt_ids(vocab=32000)
    frames_host = synth.make_frames(rank, 0, n_frames, 336, dtype=dt).pin_memory()
    frames_dev = frames_host.to(dev)

    class Walker:
        """walks the stream interval by interval (one step each), restarting after the last interval"""

        def __init__(self):
            self.k = 0
            self.tokens = self.fires = self.prefilled = 0
            self.restart()

        def restart(self):
            model.reset_stream()
            self.prompt = list(prompt0)
            self.submitted = 0
            self.pos = 0

        def step(self, frames):
            if self.pos >= n_frames:
                self.restart()
            for t in range(self.pos, self.pos + fire_every):
                if lookahead:
                    hi = min(n_frames, t + 1 + lookahead)
                    if hi > self.submitted:
                        model.prefetch_frames(frames[max(self.submitted, t):hi])
                        self.submitted = hi
                fire = 1 if (t % fire_every == fire_every - 1) else 0
                out, pred = model.stream_generate_demo(torch.tensor([self.prompt]), images_or_videos=frames[t:t + 1], modal_list=["video"],
                                                       do_sample=False, max_new_tokens=max_new, use_cache=True, force_pred=fire)
                if pred:
                    self.tokens += len(out); self.fires += 1; self.prefilled += model.last_prefill_len
                    self.prompt = self.prompt + out + turn_suffix
            self.pos += fire_every
            self.k += 1

    def timed(fn, steps):
        dist_util.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        marks = []
        for i in range(steps):
            fn()
            if i + 1 == min(steps, intervals):          # the first whole stream (or all steps when fewer)
                m = torch.cuda.Event(enable_timing=True); m.record(); marks.append(m)
        e1.record(); torch.cuda.synchronize()
        ms = dist_util.max_over_ranks(e0.elapsed_time(e1), device=dev)
        first = e0.elapsed_time(marks[0]) if marks else None
        dist_util.barrier()
        return ms, first

    w = Walker()
    for _ in range(args.warmup):
        w.step(frames_dev)
    w = Walker()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.launch_count(reset=True)
    eng.decode_stats(reset=True)
    ms, ms_first = timed(lambda: w.step(frames_dev), args.steps)
    launches = eng.launch_count(reset=True)
    dstat = eng.decode_stats(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    tokens, fires, prefilled, kv_end = w.tokens, w.fires, w.prefilled, eng.kv_len
    w2 = Walker()
    for _ in range(min(2, args.warmup)):
        w2.step(frames_host)
    w2 = Walker()
    ms_e2e, _ = timed(lambda: w2.step(frames_host), args.steps)
    # the same calls fed with RAW uint8 frames (what a video decoder yields; 360 x 640 here) from pinned host memory: the reference's
    # per-frame PIL work (expand2square + CLIP preprocess) runs on the device (sm_preprocess_frames) inside the timed region
    e2e_u8 = None
    if world == 1 and not args.no_frames_stage:
        g8 = torch.Generator().manual_seed(99)
        frames_u8 = torch.randint(0, 256, (n_frames, 360, 640, 3), dtype=torch.uint8, generator=g8).pin_memory()
        w3 = Walker()
        w3.step(frames_u8)
        w3 = Walker()
        steps_u8 = min(4, args.steps)
        ms_u8, _ = timed(lambda: w3.step(frames_u8), steps_u8)
        e2e_u8 = {"value": fire_every * steps_u8 / (ms_u8 / 1e3), "unit": "frames/s", "steps": steps_u8,
                  "h2d_bytes_per_step": fire_every * 360 * 640 * 3,
                  "note": "e2e fed with raw uint8 360 x 640 RGB frames: pad-to-square + bicubic resize + normalise on the device "
                          "(bit-exact with the reference's PIL path), then the same stream_generate_demo calls"}
        del frames_u8
    frames_per_step = fire_every
    total_frames = world * frames_per_step * args.steps
    value = total_frames / (ms / 1e3)
    e2e_value = total_frames / (ms_e2e / 1e3)

    if rank == 0:
        peaks = _peaks()
        # ---- roofline of the dominant kernel: decode_stream_kernel, one launch = one token
        dec_ms = dstat["ms"] / max(1, dstat["steps"])
        alg_bytes = DECODE_BYTES_PER_TOKEN + KV_BYTES_PER_POSITION * dstat["ctx_sum"] / max(1, dstat["steps"])
        achieved = alg_bytes / (dec_ms * 1e-3) / 1e9 if dec_ms else 0.0
        traffic, traffic_src = _tracked_traffic("decode_stream_kernel")
        roof = {"kernel": "decode_stream_kernel (one launch = one greedy token: 32 layers + lm_head + argmax)", "bound": "hbm",
                "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"], "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peaks["source"], "algorithmic_bytes_per_launch": alg_bytes,
                "avg_launch_ms": dec_ms, "launches_timed": dstat["steps"], "mean_context": dstat["ctx_sum"] / max(1, dstat["steps"]),
                "share_of_step": dstat["ms"] / ms,
                "note": "achieved = (14.221 GB of bf16 weights + 128 KiB x context of KV) / mean launch duration, CUDA events on the launch stream inside the timed region (sm_decode_stats)"}
        # ---- serial stream roofline: frames at max(tensor, HBM) + tokens + one weight pass per prefill call
        t_frame = max(VIT_GFLOP_PER_FRAME / peaks["tf_sustained"], (VIT_WEIGHT_MB / 8 + PROJ_MB_PER_FRAME / 4 + GATE_MB_PER_FRAME / 8) / peaks["hbm"]) * 1e-3
        t_tok = alg_bytes / (peaks["hbm"] * 1e9)
        roof_s = frames_per_step * args.steps * t_frame + dstat["tokens"] * t_tok + fires * DECODE_BYTES_PER_TOKEN / (peaks["hbm"] * 1e9)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            ref = CpuReference(threads)
            ref.sample_step()
            samples = [ref.sample_step() for _ in range(2)]
            mean = {k: statistics.mean(s[k] for s in samples) for k in samples[0]}
            cpu = {"value": fire_every / CpuReference.interval_seconds(mean, fire_every, max_new), "unit": "frames/s", "cores": threads,
                   "kind": "port", "sample": "2 sample steps -- " + CpuReference.SAMPLE, "unit_costs_s": mean}
        px_bytes = frames_per_step * 3 * 336 * 336 * 2
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": pol["label"] + "; CLIP-ViT-L/14-336 (23 layers) + Mamba projector + 4-layer gate + Mistral-7B (32 layers, "
                                   "vocab 32002), bf16, random-init, one frame per stream_generate_demo call; one step = one "
                                   f"{fire_every}-frame gate interval incl. its fire" + ("" if world == 1 else
                                   " -- per rank (BASELINE configs[3] policy: independent streams, one per GPU)"),
                       "frames_per_step": frames_per_step, "intervals_per_stream": intervals, "max_new_tokens": max_new,
                       "fires_timed": fires, "decoded_tokens_timed": tokens, "prefilled_positions_timed": prefilled, "kv_len_end": kv_end,
                       "frames_submitted_ahead": lookahead, "cuda_graphs": cfg.use_graphs,
                       "l2_policy": "inputs larger than L2: 14.2 GB of LLM weights are re-streamed per decoded token (L2 = 126 MB)",
                       "parallelism": f"{world} independent stream(s), one per GPU, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": px_bytes,
                    "d2h_bytes_per_step": frames_per_step * 8 + (max_new * 4 if fire_every > 1 else max_new * 4 * frames_per_step),
                    "ms_per_step": ms_e2e / args.steps,
                    "note": "pinned-host frames (H2D per frame), gate logits of every frame and the generated ids read back on the host"},
            "gpu_launches": launches, "clocks": clocks,
            "tokens_per_s": world * tokens / (ms / 1e3),
            "decode": {"tokens_per_s_in_decode_launches": 1e3 * dstat["tokens"] / dstat["ms"] if dstat["ms"] else None,
                       "ms_per_token": dec_ms, "launches_per_token": 1,
                       "note": "device time of the decode-step launches only (prefill and frame stage excluded)"},
            "roofline": roof,
            "stream_roofline": {"seconds_at_peak": roof_s, "frac": roof_s / (ms / 1e3),
                                "note": "serial per-stream bound: frames at max(tensor, HBM) + tokens at the HBM peak + one weight pass per prefill call"},
            "cpu_baseline": cpu,
        }
        if e2e_u8 is not None:
            line["e2e_uint8_frames"] = e2e_u8
        if ms_first is not None and args.steps >= intervals:
            line["whole_stream"] = {"frames": n_frames, "seconds": ms_first / 1e3, "frames_per_s": n_frames / (ms_first / 1e3),
                                    "note": f"the first {intervals} steps = exactly one configs stream (rank 0 clock)"}
        if (args.with_frames_stage or world == 1) and not args.no_frames_stage:
            line["configs1_frames_stage"] = frames_stage(eng, frames_dev[:64], peaks)
            line["reference_gpu_vit"] = reference_gpu_vit(dev, dt, frames_dev[:8])
        emit(line)
    eng.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()


def frames_stage(eng, frames_dev, peaks):
    """BASELINE configs[1] on the same handle (ViT encode + projector + gate only, no fires): pipelined (16 tickets in flight,
    tower chunks of 8) and serial B = 1 frames/s, and the tower GEMM class's share / achieved TFLOP/s (kernel-filter pass)."""
    import torch
    n = frames_dev.shape[0]

    def pipelined():
        tk = None
        for t in range(n):
            tk = eng.frame_submit(frames_dev[t:t + 1])[0]
        eng.frame_wait(tk, block=False)

    def serial():
        for t in range(n):
            eng.frame_step(frames_dev[t:t + 1], want_feats=False, want_device_outputs=False)

    def time_it(fn, reps=3):
        eng.reset_stream()
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms_p, ms_s = time_it(pipelined), time_it(serial)
    eng.kernel_filter(["gemm_tc_kernel"])
    ms_gemm = time_it(pipelined, reps=2)
    eng.kernel_filter(None)
    eng.reset_stream()
    tf = GEMM_GFLOP_PER_FRAME * n / ms_gemm
    return {"workload": f"BASELINE configs[1]: {n}-frame stream, ViT encode + projector + gate only, one frame per call, bf16 (this handle)",
            "pipelined_frames_per_s": n / (ms_p / 1e3), "serial_b1_frames_per_s": n / (ms_s / 1e3),
            "gemm_tc_kernel": {"bound": "tensor", "achieved": tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                               "frac": tf / peaks["tf_sustained"], "ms_per_frame": ms_gemm / n, "share_of_frame": ms_gemm / ms_p,
                               "note": "tower GEMM class alone inside the same captured graphs (sm_debug_kernel_filter), all tower batches in flight"}}


def run_frames_workload(args):
    """BASELINE configs[1]: 64-frame stream, ViT encode + projector + gate only, fp16."""
    import torch
    from streammind_b200 import dist_util, synth
    from streammind_b200.engine import Engine, EngineConfig
    rank, world, local = dist_util.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_util.init("nccl", dev)
    dt = torch.float16
    cfg = EngineConfig(dtype=dt, max_frames=1, llm_layers=0, use_graphs=not args.no_graphs)
    eng = Engine(cfg, device=local)
    _load_full_model(eng, cfg, dev, dt)
    n = 64
    frames_host = synth.make_frames(rank, 0, n, 336, dtype=dt).pin_memory()
    frames_dev = frames_host.to(dev)

    def step_device():
        tk = None
        for t in range(n):
            tk = eng.frame_submit(frames_dev[t:t + 1])[0]
        eng.frame_wait(tk, block=False)

    def step_e2e():
        inflight = []
        for t in range(n):
            inflight.append(eng.frame_submit(frames_host[t:t + 1]))
            if len(inflight) > 15:
                eng.frame_wait(inflight.pop(0)[0], block=True, on_stream=False)
        while inflight:
            eng.frame_wait(inflight.pop(0)[0], block=True, on_stream=len(inflight) == 0)

    def timed(fn, steps):
        dist_util.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = dist_util.max_over_ranks(e0.elapsed_time(e1), device=dev)
        dist_util.barrier()
        return ms
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.launch_count(reset=True)
    ms = timed(step_device, args.steps)
    launches = eng.launch_count(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    if rank == 0:
        peaks = _peaks()
        stage = frames_stage(eng, frames_dev, peaks)
        g = stage["gemm_tc_kernel"]
        traffic, traffic_src = _tracked_traffic("gemm_tc_kernel")
        emit({"metric": METRIC, "value": world * n * args.steps / (ms / 1e3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "f16", "data": "synthetic",
              "config": {"workload": "BASELINE configs[1]: 64-frame stream per GPU, CLIP-ViT-L/14-336 + projector + gate, fp16, one frame per "
                                     "call, NO LLM stage (encode+gate only: a sub-path of the metric)", "frames_per_step": n,
                         "l2_policy": "inputs larger than L2: 2.41 GB of weights per tower batch"},
              "e2e": {"value": world * n * args.steps / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": n * 3 * 336 * 336 * 2,
                      "d2h_bytes_per_step": n * 8}, "gpu_launches": launches, "clocks": clocks,
              "roofline": {"kernel": "gemm_tc_kernel", "bound": "tensor", "achieved": g["achieved"], "peak": g["peak"], "unit": "TFLOP/s",
                           "frac": g["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks["source"],
                           "share_of_step": g["share_of_frame"]},
              "serial_b1": {"value": stage["serial_b1_frames_per_s"], "unit": "frames/s"}, "cpu_baseline": None})
    eng.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()


def run_multi_stream(args):
    """B streams on ONE GPU through MultiStreamSession with the configs[2] policy on every stream (all fire
    on the same frames): aggregate frames/s and tokens/s.  N = 1 only."""
    import torch
    from streammind_b200 import synth
    from streammind_b200.engine import EngineConfig
    from streammind_b200.model import MultiStreamSession
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dt, B = torch.bfloat16, args.streams
    pol = POLICY["gated_decode"]
    fire_every, max_new = pol["fire_every"], pol["max_new"]
    sess = MultiStreamSession(EngineConfig(dtype=dt, max_frames=B, llm_max_ctx=8704, use_graphs=False), None, n_streams=B)
    _load_full_model(sess.engine, sess.config, dev, dt)
    prompt0, turn_suffix = synth.make_prompt_ids(vocab=32000)
    n_frames = fire_every * args.steps
    frames = torch.stack([synth.make_frames(s, 0, min(n_frames, 64), 336, dtype=dt) for s in range(B)], 1).to(dev)   # [T, B, 3, H, W]

    def run(n):
        sess.reset()
        prompts = [list(prompt0) for _ in range(B)]
        toks = 0
        for t in range(n):
            fire = 1 if (t % fire_every == fire_every - 1) else 0
            res = sess.stream_generate_demo_multi(prompts, frames[t % frames.shape[0]], force_pred=[fire] * B, do_sample=False, max_new_tokens=max_new)
            for s, (out, pred) in enumerate(res):
                if pred:
                    toks += len(out)
                    prompts[s] = prompts[s] + out + turn_suffix
        return toks
    run(fire_every)                                   # warm-up: one interval
    torch.cuda.synchronize()
    sess.engine.decode_stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    toks = run(n_frames)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    d = sess.engine.decode_stats(reset=True)
    peaks = _peaks()
    emit({"metric": "multi-stream batching on one GPU (SURVEY.md 8f-1): aggregate frames/s, configs[2] policy per stream",
          "value": B * n_frames / (ms / 1e3), "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": 1, "ms_per_step": ms / args.steps,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
          "config": {"workload": f"{B} streams on one engine, every stream: fire every {fire_every}th frame, {max_new} greedy tokens; "
                                 "shared tower batch, projector / gate / LLM-decode weight passes shared by the streams", "streams": B},
          "tokens_per_s": toks / (ms / 1e3),
          "decode": {"ms_per_step": d["ms"] / max(1, d["steps"]), "tokens_per_s_in_decode_launches": 1e3 * d["tokens"] / d["ms"],
                     "weights_GBps": DECODE_BYTES_PER_TOKEN * d["steps"] / d["ms"] / 1e6, "hbm_peak_GBps": peaks["hbm"]},
          "gpu_launches": sess.engine.launch_count()})
    sess.close()


_REAL_STDOUT = None


def _capture_stdout():
    """Libraries (NCCL's version banner, torchrun notices) print to stdout; the contract is ONE JSON line there.
    Route fd 1 to stderr for the run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    _capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16, help="timed steps; one step = one 16-frame gate interval (16 steps = one configs[2] stream)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gated_decode", choices=["gated_decode", "dense_decode", "frames"])
    ap.add_argument("--streams", type=int, default=1, help="> 1: multi-stream batching on one GPU (SURVEY.md 8f-1)")
    ap.add_argument("--lookahead", type=int, default=14, help="frames submitted ahead of the current one through the pipelined path (<= 14)")
    ap.add_argument("--no-prefetch", action="store_true", help="serial frame path (sm_frame_step per call), no overlap with the decode")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--with-frames-stage", action="store_true", help="also time the configs[1] sub-path on the same handle (extra keys); default at N = 1")
    ap.add_argument("--no-frames-stage", action="store_true", help="skip the configs[1] sub-path / reference-on-GPU extra keys")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.streams > 1:
        run_multi_stream(args)
    elif args.workload == "frames":
        run_frames_workload(args)
    else:
        run_stream_workload(args)


if __name__ == "__main__":
    main()
