#!/usr/bin/env python
"""Headline benchmark: audio-seconds/s of the wav2vec2-base forward (Wav2Vec2ForCTC.__call__) at
seq=246000, batch 32 per GPU, on N B200s of one node (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
  python bench.py --impl reference ...                            # CPU oracle port of the reference path
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N    # one rank per GPU (weak scaling, no collective
                                                                  #  on the data path; NCCL only for barrier/max)
Prints ONE JSON line on rank 0 (contract in the task statement; extra keys: roofline, cpu_baseline,
breakdown_ms, logits_max_abs_err*).
"""
import argparse
import json
import logging
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))

import torch  # noqa: E402

SAMPLE_RATE = 16000
CPU_SAMPLE_BATCH = 2


def _ncu_traffic():
    """dram bytes per launch of the two roofline kernels from the committed ncu --set full capture (B=32 x 246000 only)."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def flops_forward(cfg, L):
    """Algorithmic FLOPs (2*MAC) of one utterance, per component."""
    fr = cfg.conv_frames(L)
    T, d, ff = fr[-1], cfg.hidden_size, cfg.intermediate_size
    convs, cin = [], 1
    for t, c, k in zip(fr, cfg.filter_sizes, cfg.kernal_sizes):
        convs.append(2.0 * t * c * cin * k)
        cin = c
    per_layer = {"qkv": 2.0 * T * d * 3 * d, "out": 2.0 * T * d * d, "attn": 4.0 * T * T * d, "ffn": 4.0 * T * d * ff}
    other = {"proj": 2.0 * T * cin * d,
             "posconv": 2.0 * T * d * (d // cfg.num_conv_pos_embedding_groups) * cfg.num_conv_pos_embeddings,
             "lm_head": 2.0 * T * d * cfg.vocab_size}
    total = sum(convs) + cfg.num_layers * sum(per_layer.values()) + sum(other.values())
    return {"convs": convs, "per_layer": per_layer, "other": other, "total": total, "frames": fr}


def run_reference_train(args, cfg, params, x, cores, world):
    """CPU arm of --mode train: forward + CTC + autograd backward + Adam of the oracle port (torch CPU fp32, stage-2
    trainable set: everything but the conv extractor) on a bounded sample of the step's batch."""
    import numpy as np
    from oracle import w2v2_oracle as O
    names = [k for k in params if "/feature_extractor/" not in k]
    p = {k: (t.clone().requires_grad_(True) if k in names else t) for k, t in params.items()}
    opt = torch.optim.Adam([p[k] for k in names], lr=5e-5, eps=1e-7)
    B = x.shape[0]
    np.random.seed(0)
    labels = torch.from_numpy(np.random.randint(1, 30, size=(B, 24)))
    T = cfg.num_frames(args.seq)
    steps = max(1, min(args.steps, 2))

    def step():
        opt.zero_grad(set_to_none=True)
        lp = torch.log_softmax(O.wav2vec2_for_ctc(x, p, cfg), -1).transpose(0, 1)
        loss = torch.nn.functional.ctc_loss(lp, labels, torch.full((B,), T), torch.full((B,), 24), blank=0, reduction="sum") / B
        loss.backward()
        opt.step()
        return float(loss)
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = B * args.seq / SAMPLE_RATE / dt
    sample = f"oracle port (torch CPU fp32 autograd + Adam, dropout 0) on {B} x {args.seq} samples per step, {steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": "audio-sec/s", "value": val, "unit": "audio-sec/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"wav2vec2-base stage-2 CTC fine-tune step (forward + CTC + backward + all-reduce + Adam), "
                               f"batch={args.batch}/GPU, seq={args.seq}", "global_batch": args.batch * max(world, 1),
                   "seq_len": args.seq, "note": f"CPU arm: each step is a bounded sample of {B} utterances of that workload"},
        "cpu_baseline": {"value": val, "unit": "audio-sec/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "audio-sec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference path (oracle port; TensorFlow is not installable
    here), all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import w2v2_oracle as O
    from wav2vec2.config import Wav2Vec2Config
    cfg = Wav2Vec2Config()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = O.random_params(cfg, seed=0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(CPU_SAMPLE_BATCH, args.seq, generator=g)
    if args.mode == "train":
        return run_reference_train(args, cfg, params, x, cores, world)
    with torch.no_grad():
        for _ in range(min(args.warmup, 1)):
            O.wav2vec2_for_ctc(x, params, cfg)
        steps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(steps):
            O.wav2vec2_for_ctc(x, params, cfg)
        dt = (time.perf_counter() - t0) / steps
    val = CPU_SAMPLE_BATCH * args.seq / SAMPLE_RATE / dt
    sample = f"oracle port (torch CPU fp32) on {CPU_SAMPLE_BATCH} x {args.seq} samples per step, {steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": "audio-sec/s", "value": val, "unit": "audio-sec/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"wav2vec2-base inference (Wav2Vec2ForCTC forward), batch={args.batch}/GPU, seq={args.seq} -> "
                               f"{cfg.num_frames(args.seq)} frames", "global_batch": args.batch * max(world, 1),
                   "seq_len": args.seq, "note": f"CPU arm: each step is a bounded sample of {CPU_SAMPLE_BATCH} utterances of that workload"},
        "cpu_baseline": {"value": val, "unit": "audio-sec/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "audio-sec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_train(args, rank, world, local_rank):
    """--mode train: BASELINE.json configs[2] - the stage-2 CTC fine-tune step (src/main.py:234-250) of wav2vec2-base,
    `--batch` utterances of `--seq` samples per GPU, data parallel over the ranks with ONE NCCL all-reduce of the flat
    fp32 gradient buffer per step.  Forward in the model's precision, backward products single-pass bf16."""
    import numpy as np
    import torch.distributed as dist
    from wav2vec2 import CTCLoss, Wav2Vec2Config, Wav2Vec2ForCTC, ops
    from wav2vec2.training import Stage2Trainer
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, K = max(args.warmup, 3), args.steps
    B, L = args.batch, args.seq
    cfg = Wav2Vec2Config()                                  # the reference's training defaults: dropout 0.1, SpecAugment on
    model = Wav2Vec2ForCTC(cfg, input_shape=(B, L), precision=args.precision, device=dev).init_random(seed=0)
    trainer = Stage2Trainer(model, CTCLoss(cfg, (B, L), division_factor=B * world), learning_rate=5e-5)
    x_host = torch.randn(B, L, generator=torch.Generator().manual_seed(rank)).pin_memory()
    np.random.seed(rank)
    lab = np.zeros((B, 256), dtype=np.int32)                # main.py:51: labels padded to 256
    lab[:, :24] = np.random.randint(1, 30, size=(B, 24))    # tests/test_wav2vec2.py:41-42
    lab_host = torch.from_numpy(lab).pin_memory()
    x, labels = x_host.to(dev), lab_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n0 = ops.LAUNCHES
    losses = [trainer.step(x, labels).item()]
    launches_per_step = ops.LAUNCHES - n0
    for _ in range(W - 1):
        losses.append(trainer.step(x, labels).item())
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for _ in range(K):
            loss = trainer.step(x, labels)
        e1.record()
        barrier()
    losses.append(loss.item())
    ms = e0.elapsed_time(e1)
    # end to end: pinned host speech + labels -> device, step, loss back on the host, every step
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(K):
        x.copy_(x_host, non_blocking=True)
        labels.copy_(lab_host, non_blocking=True)
        loss_host = trainer.step(x, labels).item()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
        dist.destroy_process_group()
    if rank != 0:
        return
    audio_s = world * B * L / SAMPLE_RATE
    fl = flops_forward(cfg, L)
    enc_fl = fl["total"] - sum(fl["convs"])                 # the extractor is frozen: forward only
    step_flops = B * (sum(fl["convs"]) + 3.0 * enc_fl)      # backward ~ 2x forward for the trained part
    print(json.dumps({
        "metric": "audio-sec/s", "value": audio_s / (ms / K / 1e3), "unit": "audio-sec/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "bf16x3 forward / bf16 backward", "data": "synthetic",
        "config": {"workload": f"wav2vec2-base stage-2 CTC fine-tune step (forward + CTC + backward + all-reduce + Adam), "
                               f"batch={B}/GPU, seq={L}", "global_batch": B * world, "seq_len": L,
                   "parallelism": f"dp{world} (batch sharded; one NCCL all-reduce of {trainer.flat_g.numel()} fp32 gradients per step)",
                   "trainable_params": int(trainer.flat_w.numel()), "dropout": float(cfg.dropout),
                   "spec_augment": bool(cfg.apply_spec_augment)},
        "clocks": clk.summary(),
        "e2e": {"value": audio_s / (ms_e2e / K / 1e3), "unit": "audio-sec/s",
                "h2d_bytes_per_step": int(x_host.numel() * 4 + lab_host.numel() * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": launches_per_step * K, "model_tflops": step_flops / (ms / K / 1e3) / 1e12,
        "loss_first_last": [losses[0], losses[-1]],
    }))


def run_sweep(args, rank, world, local_rank):
    """--mode sweep: BASELINE.json configs[4] - wav2vec2-base forward over seq in {16000, 64000, 246000} x batch in {1, 8, 32, 128}
    per GPU on `world` GPUs (batch sharded, no collective; one CUDA-graph replay per step), next to the CPU oracle port timed once
    per sequence length on the host cores (rank 0, a 2-utterance sample).  One JSON line with the whole table."""
    import torch.distributed as dist
    from wav2vec2 import Wav2Vec2Config, Wav2Vec2ForCTC
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = Wav2Vec2Config()
    model = Wav2Vec2ForCTC(cfg, precision=args.precision, device=dev).init_random(seed=0)
    model.enable_cuda_graph(True)
    W, K = max(args.warmup, 3), args.steps
    cells, cpu = [], {}
    xw = torch.randn(8, 64000, device=dev)
    for _ in range(40):                     # bring the clocks up before the first (smallest) cell is timed
        model(xw)
    model._invalidate_graphs()
    del xw
    for L in (16000, 64000, 246000):
        for B in (1, 8, 32, 128):
            x = torch.randn(B, L, generator=torch.Generator().manual_seed(rank)).to(dev)
            for _ in range(W):
                model(x)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(K):
                model(x)
            b.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item() / K
            cells.append({"seq": L, "batch_per_gpu": B, "ms_per_step": round(ms, 4),
                          "audio_s_per_s": round(world * B * L / SAMPLE_RATE / (ms / 1e3), 1)})
            model._invalidate_graphs()      # one graph (and arena) per shape: drop it before the next cell
            del x
            torch.cuda.empty_cache()
    params = {k: v.detach().cpu() for k, v in model.variables.items()}
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and not args.no_cpu_baseline:
        # CPU column: after the GPU sweep, when the other ranks have left (they busy-wait in barriers while it runs otherwise)
        from oracle import w2v2_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            O.wav2vec2_for_ctc(torch.randn(1, 16000), params, cfg)        # thread-pool warm-up
            for L in (16000, 64000, 246000):
                xs = torch.randn(2, L, generator=torch.Generator().manual_seed(0))
                best = None
                for _ in range(2):
                    t0 = time.perf_counter()
                    O.wav2vec2_for_ctc(xs, params, cfg)
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                cpu[L] = 2 * L / SAMPLE_RATE / best
    if rank == 0:
        for c in cells:
            if c["seq"] in cpu:
                c["cpu_audio_s_per_s"] = round(cpu[c["seq"]], 1)
                c["vs_cpu"] = round(c["audio_s_per_s"] / cpu[c["seq"]], 1)
        print(json.dumps({"metric": "audio-sec/s", "unit": "audio-sec/s", "n_gpus": world, "steps": K, "warmup": W, "dtype": args.precision,
                          "data": "synthetic", "scaling": "weak", "higher_is_better": True,
                          "config": {"workload": "wav2vec2-base inference sweep seq x batch (BASELINE configs[4])",
                                     "parallelism": f"dp{world} (batch sharded, no collective)"},
                          "cpu_baseline": {"kind": "port", "cores": os.cpu_count() or 1, "unit": "audio-sec/s",
                                           "sample": "oracle port (torch CPU fp32) on 2 utterances per sequence length, best of 2",
                                           "value_by_seq": {str(k): round(v, 1) for k, v in cpu.items()}},
                          "cells": cells}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train", "sweep"],
                    help="infer: BASELINE configs[1] (default, the headline); train: configs[2], the stage-2 fine-tune step; sweep: configs[4], seq x batch")
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU (default 32 for infer, 8 for train)")
    ap.add_argument("--seq", type=int, default=246000)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3", "fp16", "fp16f8"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the bf16x3 / large / train sub-records of the default line")
    ap.add_argument("--no-graph", action="store_true", help="launch the ~100 kernels of a forward eagerly instead of replaying one CUDA graph")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.batch is None:
        args.batch = 8 if args.mode == "train" else 32
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.mode == "train":
        run_train(args, rank, world, local_rank)
        return
    if args.mode == "sweep":
        run_sweep(args, rank, world, local_rank)
        return

    import torch.distributed as dist
    from wav2vec2 import Wav2Vec2Config, Wav2Vec2ForCTC, ops
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    cfg = Wav2Vec2Config()
    B, L = args.batch, args.seq
    model = Wav2Vec2ForCTC(cfg, input_shape=(B, L), precision=args.precision, device=dev).init_random(seed=0)
    g = torch.Generator().manual_seed(rank)
    x_host = torch.randn(B, L, generator=g).pin_memory()
    x = x_host.to(dev)
    T = cfg.num_frames(L)
    out_host = torch.empty(B, T, cfg.vocab_size).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput
    def timed_forward(mdl, inp, steps, sample_clocks=False):
        """W warm-ups, then `steps` forwards between CUDA events (barrier + synchronize on both sides)."""
        n0 = ops.LAUNCHES
        mdl(inp)                               # eager: packs weights, sizes the arena, counts the kernels of one forward
        per_forward = ops.LAUNCHES - n0
        if not args.no_graph:
            mdl.enable_cuda_graph(True)        # public API: the same launch sequence replayed as one CUDA graph per step
        for _ in range(W):
            mdl(inp)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        a.record()
        for _ in range(steps):
            out = mdl(inp)
        b.record()
        barrier()
        if sampler:
            sampler.__exit__()
        return a.elapsed_time(b), out, per_forward, sampler

    ms, logits, launches_per_forward, clk = timed_forward(model, x, K, sample_clocks=True)
    launches = launches_per_forward * K        # kernels inside the timed region (eager launches or graph replays)
    # ---------------- end to end through the public API: pinned host input -> logits on the host, every step.
    # The input copy of step i+1 runs on a copy stream into the other of two device buffers while step i computes, the logits of
    # step i are read back on a third stream; every step still pays its own H2D copy and its own D2H read inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    xbuf = [torch.empty_like(x), torch.empty_like(x)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(n):
        main = torch.cuda.current_stream()
        for ev in consumed:
            ev.record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[0])
            xbuf[0].copy_(x_host, non_blocking=True)
            copied[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[nxt])      # the forward that last read this buffer is done
                    xbuf[nxt].copy_(x_host, non_blocking=True)
                    copied[nxt].record(copy_stream)
            main.wait_event(copied[cur])
            out = model(xbuf[cur])
            consumed[cur].record(main)
            # the logits leave on their own stream: the next forward does not wait behind the 3 MB device-to-host read
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(consumed[cur])
                out.record_stream(d2h_stream)
                out_host.copy_(out, non_blocking=True)
        main.wait_stream(d2h_stream)                          # the timed region ends when the last logits are on the host

    e2e_loop(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(K)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---------------- per-kernel-class device times (CUDA events on the launching stream)
    model.enable_cuda_graph(False)             # per-kernel events need the eager launch sequence
    breakdown_eager, gemm_ffn1 = profile_classes(model, x, cfg, steps=min(K, 5))
    base_launches = dict(_CLASS_LAUNCHES)

    # ---------------- the parity-green mode, timed on the same workload: precision="bf16x3" (the drop-in default)
    sub = {}
    other_modes = [m for m in ("bf16x3", "fp16f8", "fp16") if m != args.precision] if not args.headline_only else []
    for prec in other_modes:
        par = Wav2Vec2ForCTC(cfg, input_shape=(B, L), precision=prec, device=dev)
        par.set_variables(model.variables)
        ms3, logits3, _, _ = timed_forward(par, x, K)
        par.enable_cuda_graph(False)
        bd3, ffn1_3 = profile_classes(par, x, cfg, steps=min(K, 3))
        sub[prec] = (ms3, logits3[:CPU_SAMPLE_BATCH].float().cpu(), bd3, ffn1_3)
        del par, logits3
        torch.cuda.empty_cache()
    # ---------------- BASELINE configs[3]: wav2vec2-large (robust, 24 layers, d = 1024) inference, batch 16 x 246000
    if (B, L) == (32, 246000) and not args.headline_only:
        from wav2vec2 import RobustWav2Vec2Config
        lcfg = RobustWav2Vec2Config()
        large = Wav2Vec2ForCTC(lcfg, input_shape=(16, L), precision=args.precision, device=dev).init_random(seed=0)
        xl = x[:16].contiguous()
        am = torch.ones(16, L, dtype=torch.int32, device=dev)
        logging.getLogger("wav2vec2.modeling").setLevel(logging.ERROR)
        ms_l, _, _, _ = timed_forward(large, xl, K)
        large.enable_cuda_graph(False)
        bd_l, ffn1_l = profile_classes(large, xl, lcfg, steps=min(K, 3))
        sub["large"] = (ms_l, lcfg, bd_l, ffn1_l)
        del large, am
        torch.cuda.empty_cache()
    # ---------------- BASELINE configs[2]: the stage-2 CTC fine-tune step, 8 utterances per GPU, ONE NCCL all-reduce per step
    if (B, L) == (32, 246000) and not args.headline_only:
        sub["train"] = train_subrecord(cfg, dev, rank, world, L, W, min(K, 10), barrier)
    t = torch.tensor([sub["large"][0] if "large" in sub else 0.0] + [sub[m][0] for m in other_modes], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    msl_max, mode_ms = t.tolist()[0], dict(zip(other_modes, t.tolist()[1:]))
    ms3_max = mode_ms.get("bf16x3", 0.0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    audio_s = world * B * L / SAMPLE_RATE
    value = audio_s / (ms / K / 1e3)
    e2e = audio_s / (ms_e2e / K / 1e3)
    peaks = _peaks()
    traffic = _ncu_traffic()
    fl = flops_forward(cfg, L)
    ffn1_flops = 2.0 * B * T * cfg.hidden_size * cfg.intermediate_size
    achieved = ffn1_flops / (gemm_ffn1 * 1e-3) / 1e12 if gemm_ffn1 else None
    breakdown = breakdown_eager               # conv0 / FFN1 rooflines use the event-timed launches themselves
    result = {
        "metric": "audio-sec/s", "value": value, "unit": "audio-sec/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16", "bf16x3": "bf16x3 (split-bf16, fp32-equivalent operands)", "fp16": "fp16",
                  "fp16f8": "fp16f8 (fp16 products + e4m3 cross terms)"}[args.precision],
        "data": "synthetic",
        "config": {"workload": f"wav2vec2-base inference (Wav2Vec2ForCTC forward), batch={B}/GPU, seq={L} -> {T} frames",
                   "global_batch": B * world, "seq_len": L, "parallelism": f"dp{world} (batch sharded, no collective)",
                   "l2": "activations per step (> 3 GB) far exceed the 126 MB L2; no explicit flush needed",
                   "weights": "random init (seeded), no checkpoint offline",
                   "launch": "eager" if args.no_graph else "one CUDA graph replay per step (public enable_cuda_graph())"},
        "clocks": clk.summary(),
        "e2e": {"value": e2e, "unit": "audio-sec/s", "h2d_bytes_per_step": x_host.numel() * 4,
                "d2h_bytes_per_step": out_host.numel() * 4},
        "gpu_launches": launches,
        "model_tflops": fl["total"] * B * world / (ms / K / 1e3) / 1e12,
        "roofline": {"bound": "tensor", "kernel": f"gemm_bf16_tcgen05 FFN1 [{B * T}x{cfg.hidden_size}]x[{cfg.hidden_size}x{cfg.intermediate_size}]"
                                                   + (" + folded LayerNorm (mean / rstd applied in the epilogue)" if getattr(model, "_fold", False) else "")
                                                   + " + bias + GELU",
                     "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": (achieved / peaks["bf16_tflops"]) if achieved else None,
                     "traffic": traffic.get("ffn1_gemm_bytes_per_launch") if (B, L) == (32, 246000) else None,
                     "algorithmic_flops_per_launch": ffn1_flops,
                     "peak_source": peaks["source"] + " (burst cuBLAS bf16; sustained %.0f)" % peaks["bf16_tflops_sustained"]},
    }
    # per-class device time: CUDA events around every eager launch give the SHARES; they are scaled to the graph-replayed step
    # (eager launches lose the PDL overlap, so their raw sum exceeds ms_per_step - kept as breakdown_eager_ms)
    result["breakdown_ms"] = scale_breakdown(breakdown_eager, ms / K)
    result["breakdown_eager_ms"] = breakdown_eager
    result["roofline_classes"] = class_rooflines(result["breakdown_ms"], cfg, B, L, peaks, passes=1 if args.precision == "bf16" else 3,
                                                   launches=base_launches)
    if "bf16x3" in sub:
        _, _, bd3, ffn1_3 = sub["bf16x3"]
        ach3 = ffn1_flops / (ffn1_3 * 1e-3) / 1e12 if ffn1_3 else None
        result["value_bf16x3"] = audio_s / (ms3_max / K / 1e3)
        result["ms_per_step_bf16x3"] = ms3_max / K
        result["roofline_bf16x3"] = {"bound": "tensor", "kernel": "FFN1 GEMM, 3 MMAs per product (hi*hi + lo*hi + hi*lo)",
                                     "achieved": ach3, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                     "frac": (ach3 / peaks["bf16_tflops"]) if ach3 else None,
                                     "note": "ALGORITHMIC flops (2MKN counted once); the tensor pipe executes 3x that"}
        result["breakdown_ms_bf16x3"] = scale_breakdown(bd3, ms3_max / K)
    # every precision mode on the SAME workload: step time, throughput, cost relative to the headline mode, FFN1 roofline on
    # algorithmic flops (logits errors vs the CPU oracle are added by the cpu_baseline leg)
    result["modes"] = {args.precision: {"ms_per_step": ms / K, "value": value, "rel_step": 1.0}}
    for m in other_modes:
        _, _, bdm, f1 = sub[m]
        am = ffn1_flops / (f1 * 1e-3) / 1e12 if f1 else None
        result["modes"][m] = {"ms_per_step": mode_ms[m] / K, "value": audio_s / (mode_ms[m] / K / 1e3),
                              "rel_step": mode_ms[m] / ms, "ffn1_algorithmic_tflops": am,
                              "ffn1_frac_of_bf16_peak": (am / peaks["bf16_tflops"]) if am else None,
                              "breakdown_ms": scale_breakdown(bdm, mode_ms[m] / K)}
    if "large" in sub:
        _, lcfg, bd_l, ffn1_l = sub["large"]
        lfl = flops_forward(lcfg, L)
        lf1 = 2.0 * 16 * T * lcfg.hidden_size * lcfg.intermediate_size
        ach_l = lf1 / (ffn1_l * 1e-3) / 1e12 if ffn1_l else None
        result["large"] = {"workload": f"wav2vec2-large (robust: 24 layers, d=1024, layer-norm extractor, pre-norm, attention mask off) "
                                       f"inference, batch=16/GPU, seq={L}", "dtype": args.precision,
                           "ms_per_step": msl_max / K, "value": world * 16 * L / SAMPLE_RATE / (msl_max / K / 1e3),
                           "unit": "audio-sec/s", "model_tflops": lfl["total"] * 16 * world / (msl_max / K / 1e3) / 1e12,
                           "roofline": {"bound": "tensor", "kernel": f"FFN1 [{16 * T}x1024]x[1024x4096] + bias + GELU",
                                        "achieved": ach_l, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                        "frac": (ach_l / peaks["bf16_tflops"]) if ach_l else None},
                           "breakdown_ms": scale_breakdown(bd_l, msl_max / K)}
    if "train" in sub:
        result["train"] = sub["train"]
    # conv0 (the HBM-bound kernel): algorithmic bytes = 4*L + 2*512*T0 per utterance
    if breakdown.get("conv0"):
        by = B * (4.0 * L + 2.0 * 512 * fl["frames"][0])
        gbs = by / (breakdown["conv0"] * 1e-3) / 1e9
        with_stats = by / ((breakdown["conv0"] + breakdown.get("conv0 stats+fold", 0.0)) * 1e-3) / 1e9
        result["roofline_conv0"] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": by,
                                    "frac_incl_stats_pass": with_stats / peaks["hbm_gbs"],
                                    "traffic": traffic.get("conv0_bytes_per_launch") if (B, L) == (32, 246000) else None}
        # conv0 is 97 % stores.  MEASURED_PEAKS' HBM figure is a COPY (half reads, half writes); a pure-write stream is slower on
        # this part, so the write ceiling is measured live with the library's own fill of a buffer of conv0's output size
        try:
            wbuf = torch.empty(int(B * 2.0 * 512 * fl["frames"][0]) // 2, dtype=torch.bfloat16, device=dev)
            for _ in range(2):
                wbuf.zero_()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(5):
                wbuf.zero_()
            ev1.record()
            torch.cuda.synchronize()
            fill_gbs = wbuf.numel() * 2 / (ev0.elapsed_time(ev1) / 5 * 1e-3) / 1e9
            result["roofline_conv0"]["write_only_fill_gbs"] = fill_gbs
            result["roofline_conv0"]["frac_of_write_only_fill"] = gbs / fill_gbs
            del wbuf
        except Exception as e:          # the measurement is informative only
            result["roofline_conv0"]["write_only_fill_gbs"] = f"unavailable: {e}"
    if not args.no_cpu_baseline:
        result.update(cpu_baseline_and_error(model, cfg, x_host, logits, args, {m: sub[m][1] for m in other_modes}))
        for m in result.get("modes", {}):
            result["modes"][m]["logits_max_abs_err"] = result.get(f"logits_max_abs_err_{m}")
    print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


def scale_breakdown(eager, step_ms):
    tot = sum(eager.values())
    return {k: round(v * step_ms / tot, 4) for k, v in eager.items()} if tot > 0 else {}


_CLASS_LAUNCHES = {}


def class_rooflines(breakdown, cfg, B, L, peaks, passes, launches=None):
    """Achieved fraction of the measured peak per kernel class: algorithmic FLOPs (2*MAC, counted once in 3-pass mode) against
    the cuBLAS bf16 burst peak for the tensor-core classes, algorithmic bytes against the measured copy bandwidth for the
    HBM-bound ones (conv0: 4 L in + 2 x 512 x T0 out per utterance; LayerNorm: 4 d in + 2 d out (+ planes) per row and launch)."""
    fl = flops_forward(cfg, L)
    T, d, nl = fl["frames"][-1], cfg.hidden_size, cfg.num_layers
    planes = 2 if passes == 3 else 1
    tensor = {"conv1-6 (implicit GEMM)": sum(fl["convs"][1:]), "gemm ffn1": nl * fl["per_layer"]["ffn"] / 2,
              "gemm ffn2": nl * fl["per_layer"]["ffn"] / 2, "gemm qkv": nl * fl["per_layer"]["qkv"],
              "gemm out_proj": nl * fl["per_layer"]["out"], "attention": nl * fl["per_layer"]["attn"],
              "posconv": fl["other"]["posconv"], "gemm proj": fl["other"]["proj"], "gemm lm_head": fl["other"]["lm_head"]}
    # LayerNorm passes actually launched per forward (2 per layer + 2 without the fold; 3 with it: the others live in GEMM epilogues)
    ln_launches = (launches or {}).get("layernorm", 2 * nl + 2)
    hbm = {"conv0": 4.0 * L + 2.0 * planes * cfg.filter_sizes[0] * fl["frames"][0],
           "conv0 stats+fold": 4.0 * L,
           "layernorm": ln_launches * T * d * (4.0 + 2.0 * planes)}
    out = {}
    for k, ms in breakdown.items():
        if ms <= 0:
            continue
        if k in tensor:
            a = tensor[k] * B / (ms * 1e-3) / 1e12
            out[k] = {"bound": "tensor", "achieved": round(a, 1), "unit": "TFLOP/s", "frac": round(a / peaks["bf16_tflops"], 3)}
        elif k in hbm:
            a = hbm[k] * B / (ms * 1e-3) / 1e9
            out[k] = {"bound": "hbm", "achieved": round(a, 1), "unit": "GB/s", "frac": round(a / peaks["hbm_gbs"], 3)}
    return out


def train_subrecord(cfg_unused, dev, rank, world, L, W, K, barrier):
    """BASELINE configs[2] inside the default line: stage-2 CTC fine-tune step of wav2vec2-base (src/main.py:234-250) with the
    reference's training defaults (dropout 0.1, SpecAugment), 8 utterances of 246000 samples per GPU, data parallel with ONE
    NCCL all-reduce of the flat fp32 gradient buffer per step (world > 1).  Timed like the headline: W warm-ups, K steps between
    CUDA events, max over ranks; the all-reduce is also timed alone on the same buffer."""
    import numpy as np
    import torch.distributed as dist
    from wav2vec2 import CTCLoss, Wav2Vec2Config, Wav2Vec2ForCTC, ops
    from wav2vec2.training import Stage2Trainer
    Bt = 8
    cfg = Wav2Vec2Config()
    model = Wav2Vec2ForCTC(cfg, input_shape=(Bt, L), precision="bf16", device=dev).init_random(seed=0)
    trainer = Stage2Trainer(model, CTCLoss(cfg, (Bt, L), division_factor=Bt * world), learning_rate=5e-5)
    x = torch.randn(Bt, L, generator=torch.Generator().manual_seed(100 + rank)).to(dev)
    np.random.seed(rank)
    lab = np.zeros((Bt, 256), dtype=np.int32)               # main.py:51: labels padded to 256
    lab[:, :24] = np.random.randint(1, 30, size=(Bt, 24))
    labels = torch.from_numpy(lab).to(dev)
    n0 = ops.LAUNCHES
    # next_speech: the frozen extractor's forward of the NEXT batch overlaps this step's gradient all-reduce
    losses = [trainer.step(x, labels, next_speech=x).item()]
    per_step = ops.LAUNCHES - n0
    for _ in range(max(W, 3) - 1):
        losses.append(trainer.step(x, labels, next_speech=x).item())
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        loss = trainer.step(x, labels, next_speech=x)
    b.record()
    barrier()
    losses.append(loss.item())
    ms = a.elapsed_time(b)
    ar_ms = 0.0
    if world > 1:
        for _ in range(2):
            dist.all_reduce(trainer.flat_g)
        barrier()
        c, d_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c.record()
        for _ in range(5):
            dist.all_reduce(trainer.flat_g)
        d_.record()
        barrier()
        ar_ms = c.elapsed_time(d_) / 5
    t = torch.tensor([ms, ar_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_ms = t.tolist()
    nbytes = trainer.flat_g.numel() * trainer.flat_g.element_size()
    rec = {"workload": f"wav2vec2-base stage-2 CTC fine-tune step (forward + CTC + backward + all-reduce + Adam), batch={Bt}/GPU, "
                       f"seq={L}, dropout {cfg.dropout}, SpecAugment on", "global_batch": Bt * world, "steps": K,
           "dtype": "bf16", "ms_per_step": ms / K, "value": world * Bt * L / SAMPLE_RATE / (ms / K / 1e3), "unit": "audio-sec/s",
           "allreduce_ms": ar_ms, "allreduce_bytes": nbytes,
           "allreduce_busbw_gbs": (2.0 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9) if ar_ms > 0 else None,
           "gpu_launches": per_step * K, "trainable_params": int(trainer.flat_w.numel()),
           "loss_first_last": [losses[0], losses[-1]]}
    del trainer, model
    torch.cuda.empty_cache()
    return rec


def profile_classes(model, x, cfg, steps):
    """Device time per kernel class: every C-ABI launch is bracketed by CUDA events on the launching stream."""
    from wav2vec2 import ops
    records = []
    originals = {}

    def wrap(name, label_fn):
        fn = getattr(ops, name)
        originals[name] = fn

        def inner(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **kw)
            e.record()
            records.append((label_fn(*a, **kw), s, e))
            return r
        setattr(ops, name, inner)

    d, ff = cfg.hidden_size, cfg.intermediate_size

    def gemm_label(a, w, **kw):
        K, N = kw["K"], kw["N"]
        if "a_batch_stride" in kw:
            return "conv0" if K == 64 else "conv1-6 (implicit GEMM)"
        if K == d and N == ff:
            return "gemm ffn1"
        if K == ff and N == d:
            return "gemm ffn2"
        if N == 3 * d:
            return "gemm qkv"
        if K == d and N == d:
            return "gemm out_proj"
        if N == cfg.vocab_size:
            return "gemm lm_head"
        return "gemm proj"

    wrap("gemm", gemm_label)
    wrap("conv0", lambda *a, **kw: "conv0")
    wrap("conv0_im2col", lambda *a, **kw: "conv0")
    wrap("conv0_gn_gelu", lambda *a, **kw: "conv0")
    wrap("conv0_ln_gelu", lambda *a, **kw: "conv0")
    wrap("wave_stats", lambda *a, **kw: "conv0 stats+fold")
    wrap("conv0_fold", lambda *a, **kw: "conv0 stats+fold")
    wrap("ln_rows", lambda *a, **kw: "layernorm")
    wrap("attn_fwd", lambda *a, **kw: "attention")
    wrap("posconv", lambda *a, **kw: "posconv")
    try:
        for _ in range(steps):
            model(x)
        torch.cuda.synchronize()
    finally:
        for k, fn in originals.items():
            setattr(ops, k, fn)
    agg, cnt = {}, {}
    for label, s, e in records:
        agg[label] = agg.get(label, 0.0) + s.elapsed_time(e)
        cnt[label] = cnt.get(label, 0) + 1
    out = {k: v / steps for k, v in agg.items()}
    _CLASS_LAUNCHES.clear()
    _CLASS_LAUNCHES.update({k: c / steps for k, c in cnt.items()})      # launches per forward (class_rooflines: LayerNorm passes made)
    ffn1 = agg.get("gemm ffn1", 0.0) / max(cnt.get("gemm ffn1", 1), 1)
    return {k: round(v, 4) for k, v in sorted(out.items(), key=lambda kv: -kv[1])}, ffn1


def cpu_baseline_and_error(model, cfg, x_host, logits, args, mode_logits=None):
    """cpu_baseline leg: the oracle port timed on the host cores on a bounded sample of the same workload, and the
    logits error of the GPU path against it on those utterances (checker use of oracle/ only)."""
    from oracle import w2v2_oracle as O
    from wav2vec2 import Wav2Vec2ForCTC
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nb = min(CPU_SAMPLE_BATCH, x_host.shape[0])
    xs = x_host[:nb].clone()
    params = {k: v.detach().cpu() for k, v in model.variables.items()}
    with torch.no_grad():
        O.wav2vec2_for_ctc(xs[:1, : min(args.seq, 32000)], params, cfg)       # thread-pool warm-up
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            ref = O.wav2vec2_for_ctc(xs, params, cfg)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    val = nb * args.seq / SAMPLE_RATE / best
    err_fast = (logits[:nb].float().cpu() - ref).abs().max().item()
    out = {"cpu_baseline": {"value": val, "unit": "audio-sec/s", "cores": cores, "kind": "port",
                            "sample": f"oracle port (torch CPU fp32, stand-in for the reference's TF-2 CPU path) on "
                                      f"{nb} x {args.seq} samples, best of 3"},
           f"logits_max_abs_err_{args.precision}": err_fast, "logits_max_abs": ref.abs().max().item()}
    for m, lg in (mode_logits or {}).items():     # logits of the TIMED runs of the other precision modes (same batch, utterances 0..nb-1)
        out[f"logits_max_abs_err_{m}"] = (lg[:nb] - ref).abs().max().item()
    if args.precision == "bf16":
        if "bf16x3" not in (mode_logits or {}):
            par = Wav2Vec2ForCTC(cfg, input_shape=tuple(xs.shape), precision="bf16x3", device=model.device)
            par.set_variables(model.variables)
            out["logits_max_abs_err_bf16x3"] = (par(xs.to(model.device)).float().cpu() - ref).abs().max().item()
        out["argmax_agreement_bf16"] = (logits[:nb].argmax(-1).cpu() == ref.argmax(-1)).float().mean().item()
    return out


if __name__ == "__main__":
    main()
