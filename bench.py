#!/usr/bin/env python
"""Benchmark of the StreamMind per-frame hot path on B200 (contract: see the task statement).

Default workload (BASELINE.json configs[1]): one "step" = one 64-frame synthetic 336x336 stream through
CLIP-ViT-L/14-336 encode -> Mamba projector step -> event-gate score, fp16, streaming (one frame per
call, as the reference's demo loop does), random-init weights.  `--workload gated_decode` runs
BASELINE configs[2] (256 frames, fire every 16th, 224 greedy tokens, KV -> 4k, bf16).

  value : frames/s, frames already resident in HBM, device-timed (CUDA events), whole job over N GPUs
  e2e   : frames/s through the public per-frame call with frames in PINNED HOST memory: H2D copy of
          every frame and D2H read of every gate decision inside the timed region
  roofline / cpu_baseline : see DESIGN.md "Measurement"

N > 1: one process per GPU (torchrun), one independent stream per rank, no data-path collective
(SURVEY.md section 8e); max-over-ranks device time via an NCCL all-reduce of the elapsed time.
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

# BASELINE.json's metric, verbatim; config.workload names which of its configs a line measures (configs[1], the
# default, is the ViT encode + event-gate stream: the LLM decode stage is not triggered in it)
METRIC = "streaming frames/sec (encode+gate+decode) @336px, Mistral-7B, 1/2/4/8 B200"

# algorithmic work per frame (SURVEY.md section 8d / BASELINE.md section 3)
VIT_GFLOP_PER_FRAME = 366.0
GATE_MB_PER_FRAME = 1577.1
PROJ_MB_PER_FRAME = 254.8
DECODE_GB_PER_TOKEN = 14.221


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


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


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path, all host threads
# ------------------------------------------------------------------------------------------------
def cpu_reference_frames(n_frames: int, threads: int):
    """ViT -> projector (over ALL frames so far, as the reference re-runs it: videollama2_arch.py:190-198)
    -> gate, fp32, full-size random-init weights, on the host.  Returns seconds for n_frames."""
    import torch
    from oracle import restate as R
    from streammind_b200 import synth
    torch.set_num_threads(threads)
    sd = {}
    sd.update(synth.make_vit_weights(1234, torch.float32))
    sd.update(synth.make_projector_gate_weights(1234, torch.float32, gate_with_qk=True))
    vit, mam, gate = R.VitConfig(), R.MambaCfg(), R.gate_config()
    frames = synth.make_frames(0, 0, n_frames, 336, dtype=torch.float32)
    with torch.no_grad():
        R.clip_vision_tower(sd, vit, frames[:1])      # warm-up (thread pool, allocator)
        t0 = time.perf_counter()
        feats = None
        for t in range(n_frames):
            f = R.clip_vision_tower(sd, vit, frames[t:t + 1]).unsqueeze(0)
            feats = f if feats is None else torch.cat([feats, f], dim=1)
            x = R.projector_sequence(sd, mam, feats)
            R.gate_decision(R.gate_logits(sd, gate, x[0, -1]))
        dt = time.perf_counter() - t0
    return dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.cpu_frames
    times = []
    for _ in range(max(1, args.warmup // 3)):
        cpu_reference_frames(1, threads)
    for _ in range(max(1, min(args.steps, 3))):
        times.append(cpu_reference_frames(n, threads))
    sec = statistics.mean(times)
    fps = n / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * sec,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1] on host cores: {n}-frame sample of the 64-frame stream, "
                               "ViT-L/14-336 + projector (re-run over all frames, as the reference does) + gate"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{n} frames, fp32, oracle/restate.py (the reference's algorithm; the reference "
                                   "itself is Python that cannot travel to the GPU box)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from streammind_b200 import dist_util, synth
    from streammind_b200.engine import Engine, EngineConfig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a GPU; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_util.init("nccl", dev)
    dt = torch.float16
    n_frames, chunk = args.frames, args.chunk
    cfg = EngineConfig(dtype=dt, max_frames=max(chunk, 1), llm_layers=0, use_graphs=not args.no_graphs)
    eng = Engine(cfg, device=local)
    seed = 1234
    sd = {}
    sd.update(synth.make_vit_weights(seed, dt, device=dev, layers=cfg.vit_layers))
    sd.update(synth.make_projector_gate_weights(seed, dt, device=dev))
    eng.load_state_dict(sd)
    eng.finalize()
    del sd
    torch.cuda.empty_cache()

    frames_host = synth.make_frames(rank, 0, n_frames, 336, dtype=dt).pin_memory()
    frames_dev = frames_host.to(dev)

    pipelined = not args.no_pipeline

    def step_serial():
        for t in range(0, n_frames, chunk):
            eng.frame_step(frames_dev[t:t + chunk], want_feats=False, want_device_outputs=False)

    def step_device():
        if not pipelined:
            return step_serial()
        tk = None
        for t in range(0, n_frames, chunk):
            tk = eng.frame_submit(frames_dev[t:t + chunk])[0]
        eng.frame_wait(tk, block=False)             # the timing stream is ordered after the last gate decision

    preds = []

    def step_e2e():
        """Pinned-host frames in, every gate decision read back on the host.  Pipelined mode keeps `lookahead`
        frames in flight: frame t+lookahead is submitted before the host blocks on the decision of frame t."""
        preds.clear()
        if not pipelined:
            for t in range(0, n_frames, chunk):
                _, _, _, lg = eng.frame_step(frames_host[t:t + chunk], want_feats=False, want_device_outputs=False)
                torch.cuda.current_stream().synchronize()          # the host needs the decision to act on it
                for i in range(lg.shape[0]):
                    preds.append(int(lg[i, 1] > lg[i, 0]))
            return
        inflight = []
        def drain_one(on_stream):
            tk, _, _, _, lgh = inflight.pop(0)
            eng.frame_wait(tk, block=True, on_stream=on_stream)
            for i in range(lgh.shape[0]):
                preds.append(int(lgh[i, 1] > lgh[i, 0]))
        for t in range(0, n_frames, chunk):
            inflight.append(eng.frame_submit(frames_host[t:t + chunk]))
            if len(inflight) > args.lookahead:
                drain_one(False)
        while inflight:
            drain_one(len(inflight) == 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = dist_util.max_over_ranks(e0.elapsed_time(e1), device=dev)
        barrier()
        return ms

    eng.reset_stream()
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.launch_count(reset=True)
    ms = timed(step_device, args.steps)
    launches = eng.launch_count(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    total_frames = world * n_frames * args.steps
    value = total_frames / (ms / 1e3)
    e2e_value = total_frames / (ms_e2e / 1e3)

    if rank == 0:
        peaks = _peaks()
        # ---- per-kernel-class pass: the SAME step (same graphs, same tile plans, same number of frames in flight),
        # restricted to one kernel class at a time (sm_debug_kernel_filter) and timed with CUDA events on the
        # launch stream.  In the pipelined mode several towers run concurrently, so a class's ms/frame is its
        # share of the machine's time per frame (launch durations overlap); per-launch latency is reported from
        # the serial pass below.
        def class_ms_per_frame(classes, fn):
            eng.kernel_filter(classes)
            eng.reset_stream()
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            eng.launch_count(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                fn()
            e1.record()
            torch.cuda.synchronize()
            n = eng.launch_count(reset=True)
            eng.kernel_filter(None)
            return e0.elapsed_time(e1) / (2 * n_frames), n / (2 * n_frames)
        prof, prof_serial = {}, {}
        for cls in Engine.KERNEL_CLASSES:
            prof[cls] = class_ms_per_frame([cls], step_device)
        for cls in ("gemm_tc_kernel", "gemv_kernel"):
            prof_serial[cls] = class_ms_per_frame([cls], step_serial) if pipelined else prof[cls]
        eng.reset_stream()
        if pipelined:                       # rank-local timing (no collectives inside the rank-0 block)
            step_serial(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step_serial(); e1.record(); torch.cuda.synchronize()
            serial_ms = e0.elapsed_time(e1)
        else:
            serial_ms = ms / args.steps
        eng.reset_stream()
        tot_ms = sum(v[0] for v in prof.values())
        per_frame = {k: {"ms_per_frame": v[0], "share_of_class_sum": v[0] / tot_ms} for k, v in prof.items()}
        frame_ms = ms / args.steps / n_frames
        # frames sharing one pass over the weights: projector (GEMV, batches of <= 4), gate (GEMMs over the whole tower
        # batch when it has >= 5 frames, else GEMV batches of <= 4)
        tower_batch = int(os.environ.get("SMB_TOWER_BATCH", "8")) if (pipelined and chunk == 1) else chunk
        proj_batch = min(4, tower_batch)
        gate_as_gemm = tower_batch >= int(os.environ.get("SMB_GATE_GEMM", "5")) > 0
        gate_batch = tower_batch if gate_as_gemm else min(4, tower_batch)
        gemm, gemv = prof.get("gemm_tc_kernel", (0, 1)), prof.get("gemv_kernel", (0, 1))
        gate_gemm = prof.get("gate_gemm_kernel", (0.0, 0))
        gemm_gflop = VIT_GFLOP_PER_FRAME * (334.65 / 366.0)                     # GEMM share of the ViT flops
        # bytes actually streamed per frame by the weight-streaming kernels
        gate_mb_streamed = GATE_MB_PER_FRAME / gate_batch
        proj_mb_streamed = PROJ_MB_PER_FRAME / proj_batch
        gemv_mb_streamed = proj_mb_streamed + (0.0 if gate_as_gemm else gate_mb_streamed)
        gemm_tf = gemm_gflop / gemm[0] if gemm[0] else 0.0
        gemv_gbs = gemv_mb_streamed / gemv[0] if gemv[0] else 0.0
        gate_gbs = gate_mb_streamed / gate_gemm[0] if (gate_as_gemm and gate_gemm[0]) else None
        roof_gemv = {"kernel": "gemv_kernel", "bound": "hbm", "achieved": gemv_gbs, "peak": peaks["hbm"], "unit": "GB/s",
                     "frac": gemv_gbs / peaks["hbm"], "traffic": 234.99e6,
                     "traffic_note": "dram read bytes of the largest launch (gate|up of one gate layer, 2 x 14336 x 4096 x 2 B = "
                                     "234.9 MB algorithmic, shared by 4 frames) from ncu --set full, profiles/r01_ncu_pipelined.md: no re-reads",
                     "share_of_step": gemv[0] / frame_ms,
                     "avg_launch_us": 1e3 * gemv[0] * (proj_batch / 5.0 if gate_as_gemm else gate_batch / 22.0),
                     "launches_per_weight_pass": 5 if gate_as_gemm else 22,
                     "frames_per_weight_pass": proj_batch if gate_as_gemm else gate_batch,
                     "covers": "projector" if gate_as_gemm else "projector + gate",
                     "bytes_streamed_per_frame": gemv_mb_streamed * 1e6,
                     "algorithmic_bytes_per_frame_unbatched": (GATE_MB_PER_FRAME + PROJ_MB_PER_FRAME) * 1e6,
                     "serial_b1": {"ms_per_frame": prof_serial["gemv_kernel"][0],
                                   "achieved_GBps": (GATE_MB_PER_FRAME + PROJ_MB_PER_FRAME) / prof_serial["gemv_kernel"][0]}}
        roof_gemm = {"kernel": "gemm_tc_kernel", "bound": "tensor", "achieved": gemm_tf, "peak": peaks["tf_sustained"],
                     "unit": "TFLOP/s", "frac": gemm_tf / peaks["tf_sustained"], "traffic": 9.04e6,
                     "traffic_note": "mean dram bytes per launch of the 4 per-layer GEMMs with the pipelined tile plans (ncu --set "
                                     "full, profiles/r01_ncu_pipelined.md: qkv 7.56, out 4.55, fc1 9.67, fc2 14.39 MB) vs 6.3 MB of "
                                     "weights per launch on average: weights are read once, activations mostly from L2",
                     "share_of_step": gemm[0] / frame_ms,
                     "avg_launch_us": 1e3 * prof_serial["gemm_tc_kernel"][0] / 93.0,
                     "launches_per_frame": 93,
                     "algorithmic_flops_per_frame": gemm_gflop * 1e9,
                     "note": "achieved = GEMM flops per frame / the GEMM class's machine time per frame with all lanes in "
                             "flight; avg_launch_us is the latency of one launch in the serial B=1 pass",
                     "serial_b1": {"ms_per_frame": prof_serial["gemm_tc_kernel"][0],
                                   "achieved_TFLOPs": gemm_gflop / prof_serial["gemm_tc_kernel"][0]}}
        roof_gate = None
        if gate_as_gemm:
            roof_gate = {"kernel": "gemm_tc_kernel (batched gate: weights on the MMA lanes, weight streaming)", "bound": "hbm",
                         "achieved": gate_gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gate_gbs / peaks["hbm"],
                         "traffic": None, "share_of_step": gate_gemm[0] / frame_ms, "frames_per_weight_pass": gate_batch,
                         "bytes_streamed_per_frame": gate_mb_streamed * 1e6,
                         "algorithmic_bytes_per_frame_unbatched": GATE_MB_PER_FRAME * 1e6, "peak_source": peaks["source"],
                         "note": "includes the row kernels between the GEMMs (RMSNorm, GQA expand, SwiGLU, split-K sums)"}
        dominant, secondary = (roof_gemm, roof_gemv) if gemm[0] >= gemv[0] else (roof_gemv, roof_gemm)
        if roof_gate is not None and gate_gemm[0] > secondary["share_of_step"] * frame_ms:
            secondary = roof_gate
        dominant["peak_source"] = secondary["peak_source"] = peaks["source"]
        # frame-level roofline (with the gate weights shared by gate_batch frames)
        t_tensor = VIT_GFLOP_PER_FRAME / peaks["tf_sustained"]
        t_hbm = (578.8 / (tower_batch if (pipelined and chunk == 1) else chunk) + proj_mb_streamed + gate_mb_streamed) / peaks["hbm"]
        frame_roof_ms = max(t_tensor, t_hbm)

        # ---- CPU baseline (oracle port) on this box's host cores, bounded sample
        cpu = None
        if not args.no_cpu_baseline and world == 1:       # reported at N = 1 only (bounded sample on the host cores)
            threads = os.cpu_count() or 1
            sec = cpu_reference_frames(args.cpu_frames, threads)
            cpu = {"value": args.cpu_frames / sec, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_frames} frames of the same stream, fp32, oracle/restate.py on all host threads"}

        px_bytes = n_frames * 3 * 336 * 336 * 2
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: {n_frames}-frame 336x336 synthetic stream per GPU, CLIP-ViT-L/14-336 "
                                   f"(23 layers) + Mamba projector step + 4-layer Mistral gate, fp16, random-init, "
                                   f"{chunk} frame(s) per call",
                       "frames_per_step": n_frames, "chunk": chunk, "cuda_graphs": cfg.use_graphs,
                       "pipelined": pipelined, "frames_in_flight": (16 if pipelined else 1),
                       "tower_batch": (int(os.environ.get("SMB_TOWER_BATCH", "8")) if (pipelined and chunk == 1) else 1),
                       "gate_batch": gate_batch, "gate_as_gemm": gate_as_gemm,
                       "e2e_lookahead": (args.lookahead if pipelined else 0),
                       "l2_policy": "inputs larger than L2: 2.41 GB of weights are re-streamed per frame (L2 = 126 MB)",
                       "parallelism": f"{world} independent stream(s), one per GPU, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": px_bytes,
                    "d2h_bytes_per_step": n_frames * 8, "ms_per_step": ms_e2e / args.steps,
                    "note": "pinned-host frames, H2D per frame, gate logits read back and stream synchronised per call"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": dominant,
            "roofline_secondary": secondary,
            "roofline_gemv": roof_gemv,
            "frame_roofline": {"ms_per_frame_at_peak": frame_roof_ms, "tensor_ms": t_tensor, "hbm_ms": t_hbm,
                               "frac": frame_roof_ms / frame_ms,
                               "note": "per-frame bound max(tensor: 366 GFLOP of ViT, HBM: ViT weights once per tower batch, "
                                       "projector / gate weights once per batch) at the measured peaks"},
            "serial_b1": {"value": world * n_frames / (serial_ms / 1e3), "unit": "frames/s", "measured_on": "rank 0",
                          "note": "same stream through the serial sm_frame_step (one frame in flight, no batching)"},
            "kernel_breakdown": per_frame,
            "cpu_baseline": cpu,
            "gate_fire_rate": sum(preds) / max(1, len(preds)),
        }
        emit(line)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_decode_workload(args):
    """BASELINE configs[2] / [4]: ViT encode + gated Mistral-7B greedy decode through the reference-facing API
    (StreamMindB200ForCausalLM.stream_generate_demo), bf16, persistent KV cache.  Gate policy is overridden
    (fire schedule below, the authors' own "# pred = 1" switch) while gate logits are still computed."""
    import torch
    from streammind_b200 import dist_util, synth
    from streammind_b200.engine import EngineConfig
    from streammind_b200.model import StreamMindB200ForCausalLM
    rank, world, local = dist_util.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_util.init("nccl", dev)
    dt = torch.bfloat16
    if args.workload == "gated_decode":
        n_frames, fire_every, max_new, label = 256, 16, 224, "BASELINE configs[2]: 256-frame stream, fire every 16th frame, 224 greedy tokens per fire, KV -> ~4.1k"
    else:
        n_frames, fire_every, max_new, label = 512, 1, 5, "BASELINE configs[4]: gate-off dense mode, 512 frames, every frame fires, 5 greedy tokens, KV -> ~8.3k"
    if args.frames != 64:
        n_frames = args.frames
    cfg = EngineConfig(dtype=dt, max_frames=1, llm_max_ctx=8704, use_graphs=not args.no_graphs)
    model = StreamMindB200ForCausalLM(cfg, None, device=local)
    seed = 1234
    for part in (synth.make_vit_weights(seed, dt, device=dev, layers=cfg.vit_layers),
                 synth.make_projector_gate_weights(seed, dt, device=dev)):
        model.engine.load_state_dict(part)
        del part
    # the LLM is generated and uploaded layer by layer (14.5 GB)
    full = synth.make_mistral_weights(seed, "", dt, device=dev, vocab=cfg.llm_vocab)
    model.engine.load_state_dict(full)
    del full
    torch.cuda.empty_cache()
    model.engine.finalize()
    prompt0, turn_suffix = synth.make_prompt_ids(vocab=32000)
    frames_host = synth.make_frames(rank, 0, n_frames, 336, dtype=dt).pin_memory()
    frames_dev = frames_host.to(dev)
    stats = {}

    def one_stream(frames):
        model.reset_stream()
        prompt = list(prompt0)
        toks = fires = 0
        t_dec = 0.0
        for t in range(n_frames):
            fire = 1 if (t % fire_every == fire_every - 1) else 0
            ids = torch.tensor([prompt])
            t0 = time.perf_counter()
            out, pred = model.stream_generate_demo(ids, images_or_videos=frames[t:t + 1], modal_list=["video"],
                                                   do_sample=False, max_new_tokens=max_new, use_cache=True,
                                                   force_pred=fire)
            if pred:
                t_dec += time.perf_counter() - t0
                toks += len(out); fires += 1
                prompt = prompt + out + turn_suffix
        stats.update(tokens=toks, fires=fires, kv_len=model.engine.kv_len, decode_wall_s=t_dec)

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
        one_stream(frames_dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    model.engine.launch_count(reset=True)
    ms = timed(lambda: one_stream(frames_dev), args.steps)
    launches = model.engine.launch_count(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    dec_tok_s = stats["tokens"] / stats["decode_wall_s"]
    ms_e2e = timed(lambda: one_stream(frames_host), args.steps)
    if rank == 0:
        peaks = _peaks()
        fps = world * n_frames * args.steps / (ms / 1e3)
        gbs = DECODE_GB_PER_TOKEN * dec_tok_s            # weights only; KV traffic comes on top
        # serial roofline of the stream
        t_frame = (578.8 + GATE_MB_PER_FRAME + PROJ_MB_PER_FRAME) / peaks["hbm"] * 1e-3
        t_tok = DECODE_GB_PER_TOKEN / peaks["hbm"]
        roof_s = n_frames * t_frame + stats["tokens"] * t_tok + stats["fires"] * t_tok
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": label + "; CLIP-ViT-L/14-336 + projector + gate + Mistral-7B (32 layers, vocab 32002), random-init, one frame per call",
                       "frames_per_step": n_frames, "fires": stats["fires"], "decoded_tokens": stats["tokens"],
                       "kv_len_end": stats["kv_len"], "cuda_graphs": cfg.use_graphs,
                       "l2_policy": "inputs larger than L2: 14.2 GB of LLM weights are re-streamed per decoded token",
                       "parallelism": f"{world} independent stream(s), one per GPU"},
            "e2e": {"value": world * n_frames * args.steps / (ms_e2e / 1e3), "unit": "frames/s",
                    "h2d_bytes_per_step": n_frames * 3 * 336 * 336 * 2, "d2h_bytes_per_step": n_frames * 8 + stats["tokens"] * 4,
                    "note": "pinned-host frames, H2D per frame, gate logits and generated ids read back"},
            "gpu_launches": launches, "clocks": clocks,
            "decode": {"tokens_per_s": dec_tok_s, "ms_per_token": 1e3 / dec_tok_s,
                       "note": "wall time of the fire calls (prefill of the new dialogue suffix + greedy decode) / tokens"},
            "roofline": {"kernel": "gemv_kernel (LLM decode)", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm"],
                         "unit": "GB/s", "frac": gbs / peaks["hbm"], "traffic": None, "peak_source": peaks["source"],
                         "algorithmic_bytes_per_token": DECODE_GB_PER_TOKEN * 1e9},
            "stream_roofline": {"seconds_at_peak": roof_s, "frac": roof_s / (ms / args.steps / 1e3)},
            "cpu_baseline": None,
        }
        emit(line)
    model.engine.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()


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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="frames", choices=["frames", "gated_decode", "dense_decode"])
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=1, help="frames per call (1 = streaming, as the reference's demo)")
    ap.add_argument("--cpu-frames", type=int, default=8, help="frames in the bounded CPU sample")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="serial sm_frame_step instead of sm_frame_submit/wait")
    ap.add_argument("--lookahead", type=int, default=15, help="e2e: frames submitted ahead of the decision being read (<= 15)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload != "frames":
        run_decode_workload(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
