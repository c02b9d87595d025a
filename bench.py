#!/usr/bin/env python
"""bench.py -- images/sec of the 1000-step DDPM UNet2D sampling loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is ONE pass of the hot path (image_generator.py:400-403: UNet2D forward + DDPMScheduler
step) over one batch of 64 synthetic images on every GPU; an image needs T = 1000 such steps, so
    images/sec = n_gpus * 64 / (1000 * seconds_per_step).
Steps are homogeneous (same kernels, same shapes for every t), the K timed steps are consecutive
steps of the real 1000-step schedule replayed from the captured CUDA graph.

Keys beyond the base contract: ``roofline`` (dominant kernel = the tcgen05 implicit-GEMM
convolution, timed live with CUDA-event pairs around every launch of one step), ``cpu_baseline``
(the fp32 PyTorch oracle on the host cores, bounded sample), ``e2e`` (the same step driven through
the public API with pinned-host input/output copies inside the timed region), ``clocks``.

``--impl reference`` times the reference's own CPU path.  The reference's arithmetic lives in
`diffusers`, which is not installed here, so the arm runs the oracle port of it
(oracle/) with all host threads at the reference's own batch size (B=1, image_generator.py:379).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_STEPS = 1000
GFLOP_PER_IMAGE_STEP = 75.277          # UNet forward, 2*MAC
METRIC = "images/sec, 1000-step DDPM UNet2D sampling"


def load_conv_traffic():
    """Average DRAM bytes per conv_tc2 launch from the newest committed ncu launch list (profiles/*_traffic.json,
    written by tools/summarize_ncu.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`)."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            k = json.load(open(path))["kernels"]
            if not any(n.startswith("attention_tc") for n in k):
                continue                                         # not a sampling-step capture (e.g. the classifier gradient pass)
            tot = sum(v["avg_dram_bytes_per_launch"] * v["launches"] for n, v in k.items()
                      if n.startswith("conv_tc") and v["avg_dram_bytes_per_launch"])
            cnt = sum(v["launches"] for n, v in k.items() if n.startswith("conv_tc") and v["avg_dram_bytes_per_launch"])
            if cnt:
                return tot / cnt, os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": sorted(reasons)}


def cpu_oracle_steps(batch: int, steps: int, warmup: int):
    """fp32 PyTorch oracle: UNet forward + scheduler step on the host cores."""
    import torch
    from oracle.ddpm import DDPMSchedulerOracle
    from oracle.unet2d import build_unet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_unet(0)
    sched = DDPMSchedulerOracle()
    sched.set_timesteps(T_STEPS)
    g = torch.Generator().manual_seed(42)
    x = torch.randn(batch, 3, 128, 128, generator=g)
    ts = sched.timesteps.tolist()
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            eps = model(x, ts[i % T_STEPS]).sample
            x = sched.step(eps, ts[i % T_STEPS], x, generator=g).prev_sample
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b_ref = 1                                   # the reference samples at B=1 (image_generator.py:379)
    sec, cores = cpu_oracle_steps(b_ref, args.steps, args.warmup)
    value = b_ref / (T_STEPS * sec)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1000-step DDPM UNet2D sampling (BASELINE configs[1]); reference CPU path",
                   "step": "UNet2D forward + DDPMScheduler.step", "batch_per_step": b_ref, "T": T_STEPS},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} denoising steps at B={b_ref} (oracle port of the diffusers path; "
                                   "diffusers itself is not installable offline)"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _dbg(msg):
    if os.environ.get("BENCH_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:.1f}s] {msg}", file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def run_ours(args):
    import torch
    from synt_isic_b200 import DDPMScheduler, SUPPORTED_CONFIG, UNet2DModel
    from synt_isic_b200.dist import init_from_env, max_over_ranks
    import torch.distributed as dist

    # NCCL prints its version banner on stdout when the first communicator is built: keep stdout for the JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        rank, world, local = init_from_env("nccl")
        assert world == args.gpus or world == 1, f"WORLD_SIZE={world} but --gpus {args.gpus}"
        dev = torch.device(f"cuda:{local}")
        torch.cuda.set_device(dev)
        if world > 1:
            dist.barrier(device_ids=[local])
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    B = args.batch
    peaks = load_peaks()

    torch.manual_seed(1000 + rank)
    model = UNet2DModel(precision="bf16", **SUPPORTED_CONFIG).to(dev)    # random-init weights of the reference architecture
    sched = DDPMScheduler(num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2", prediction_type="epsilon")
    sched.set_timesteps(T_STEPS)
    g = torch.Generator().manual_seed(42 + rank)
    x_host = torch.randn(B, 3, 128, 128, generator=g).pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    stream = torch.cuda.Stream(device=dev)
    K, W = args.steps, max(args.warmup, 3)

    def run_steps(x, begin, n):
        """n consecutive steps of the 1000-step schedule starting at index `begin` (wraps)."""
        done = 0
        while done < n:
            b = (begin + done) % T_STEPS
            e = min(T_STEPS, b + (n - done))
            model.sample(x, sched, seed=1234 + rank, image_offset=rank * B, step_begin=b, step_end=e,
                         micro_batch=args.micro_batch)
            done += e - b

    with torch.cuda.stream(stream):
        x = x_host.to(dev, non_blocking=True)
        _dbg("model on device")
        run_steps(x, 0, W)                                    # warm-up: pool sizing, graph capture, clocks
        _dbg("warm-up done")
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = model.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        run_steps(x, W, K)
        ev1.record(stream)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None
        launches = model.launch_count() - l0
        _dbg("timed region done")
        sec_step = max_over_ranks(ev0.elapsed_time(ev1) / 1e3 / K, dev, dist.group.WORLD if world > 1 else None)
        _dbg("max over ranks done")
        finite = bool(torch.isfinite(x).all().item())

        # ---- e2e: public API, pinned host input/output copies inside the timed region, every step
        Ke = min(K, 20)
        for i in range(2):
            x.copy_(x_host, non_blocking=True); run_steps(x, i, 1); out_host.copy_(x, non_blocking=True)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(Ke):
            x.copy_(x_host, non_blocking=True)
            run_steps(x, (W + i) % T_STEPS, 1)
            out_host.copy_(x, non_blocking=True)
        e1.record(stream)
        stream.synchronize()
        sec_e2e_serial = max_over_ranks(e0.elapsed_time(e1) / 1e3 / Ke, dev, dist.group.WORLD if world > 1 else None)

        # the same work with the copies on their own streams (double-buffered staging tensors): the H2D copy of step
        # i+1 and the D2H copy of step i-1 overlap the kernels of step i; every step still starts from pinned host
        # memory and ends in pinned host memory
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        stage_in = [torch.empty_like(x) for _ in range(2)]
        stage_out = [torch.empty_like(x) for _ in range(2)]
        outs = [torch.empty_like(x_host).pin_memory() for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]

        def pipelined(n):
            for i in range(n):
                b = i & 1
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_done[b])                       # staging buffer b was consumed by step i-2
                    stage_in[b].copy_(x_host, non_blocking=True)
                    ev_in[b].record(s_in)
                stream.wait_event(ev_in[b])
                stream.wait_event(ev_out[b])                          # result buffer b was read back (step i-2)
                x.copy_(stage_in[b], non_blocking=True)
                run_steps(x, (W + i) % T_STEPS, 1)
                stage_out[b].copy_(x, non_blocking=True)
                ev_done[b].record(stream)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[b])
                    outs[b].copy_(stage_out[b], non_blocking=True)
                    ev_out[b].record(s_out)
            stream.wait_event(ev_out[0]); stream.wait_event(ev_out[1])

        pipelined(4)
        stream.synchronize(); s_in.synchronize(); s_out.synchronize()
        if world > 1:
            dist.barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        pipelined(Ke)
        p1.record(stream)
        stream.synchronize(); s_in.synchronize(); s_out.synchronize()
        sec_e2e = max_over_ranks(p0.elapsed_time(p1) / 1e3 / Ke, dev, dist.group.WORLD if world > 1 else None)
        _dbg("e2e done")

        # ---- per-kernel profile of one step (CUDA-event pair around every launch)
        prof = model.profile_step(x, sched, micro_batch=args.micro_batch) if rank == 0 else None
        stream.synchronize()

    _dbg("profile done")
    value = world * B / (T_STEPS * sec_step)
    e2e_value = world * B / (T_STEPS * sec_e2e)
    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    conv = prof["conv_tcgen05"]
    traffic, traffic_src = load_conv_traffic()
    conv_tflops = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    peak = peaks["bf16_sustained"]                       # the kernel is timed inside a long step
    step_ms_profiled = sum(v["ms"] for v in prof.values())
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": sec_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "1000-step DDPM UNet2D sampling, batch 64 per GPU, random-init repo-default UNet2D "
                               "(BASELINE configs[1])",
                   "step": "UNet2D forward + DDPMScheduler.step over the batch (one CUDA-graph replay)",
                   "batch_per_gpu": B, "micro_batch": args.micro_batch or B, "T": T_STEPS,
                   "noise": "in-kernel Philox", "parallelism": f"independent sample batches x{world}, no data-path collective",
                   "l2": "per-step activation working set (GBs) exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * 128 * 128 * 4,
                "d2h_bytes_per_step": B * 3 * 128 * 128 * 4, "ms_per_step": sec_e2e * 1e3,
                "serial_value": world * B / (T_STEPS * sec_e2e_serial), "serial_ms_per_step": sec_e2e_serial * 1e3,
                "api": "UNet2DModel.sample (C ABI synt_unet_sample); every step copies x from pinned host memory and its result back "
                       "to pinned host memory; value: copies on their own streams (double-buffered, overlapping the neighbouring steps), "
                       "serial_value: copies and kernels on one stream"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_tc2_kernel / conv_tc_kernel (persistent tcgen05 implicit-GEMM convolutions)",
                     "achieved": conv_tflops, "peak": peak, "unit": "TFLOP/s", "frac": conv_tflops / peak,
                     "peak_source": peaks["source"] + ", sustained bf16", "traffic": traffic, "traffic_unit": "bytes/launch (DRAM read+write, ncu)",
                     "traffic_source": traffic_src,
                     "algorithmic": "2*M*N*K of the reference's convolutions incl. the 1x1 attention projections (69.7 GFLOP/image/step), summed over the conv launches of "
                                    "one step / their summed CUDA-event durations; the fused Upsample2D convs execute 4/9 of their share",
                     "launches_per_step": conv["launches"], "flops_per_step": conv["flops"], "ms_per_step": conv["ms"],
                     "whole_step_tflops": B * GFLOP_PER_IMAGE_STEP * 1e9 / sec_step / 1e12,
                     "whole_step_frac": B * GFLOP_PER_IMAGE_STEP * 1e9 / sec_step / 1e12 / peak},
        "step_breakdown_ms": {k: round(v["ms"], 4) for k, v in prof.items()},
        "step_breakdown_launches": {k: v["launches"] for k, v in prof.items()},
        "step_ms_profiled": step_ms_profiled,
        "finite": finite,
        "workspace_gb": model.workspace_bytes() / 1e9,
    }
    # ---- second half of the BASELINE metric: Time-SHAP sec/image (xai/XAI.py:1179-1234 over a
    #      T=1000-frame trajectory; ResNet18 logit evaluations through the drop-in classifier)
    try:
        from synt_isic_b200 import MelanomaClassifierAdaptive, xai
        clf = MelanomaClassifierAdaptive(num_classes=7, pretrained=False, precision="bf16").to(dev).eval()
        gt = torch.Generator().manual_seed(7)
        traj_host = torch.tanh(torch.randn(T_STEPS, 3, 128, 128, generator=gt)).pin_memory()
        with torch.cuda.stream(stream):
            traj = traj_host.to(dev, non_blocking=True)
            xai.compute_time_shap(clf, traj, list(range(T_STEPS)), 0)           # warm-up
            stream.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(stream)
            for _ in range(3):
                xai.compute_time_shap(clf, traj, list(range(T_STEPS)), 0)
            t1.record(stream)
            stream.synchronize()
            sec_dev = t0.elapsed_time(t1) / 3e3
            w0 = time.perf_counter()                                             # e2e: pinned host frames in, scores out
            for _ in range(3):
                xai.compute_time_shap(clf, traj_host, list(range(T_STEPS)), 0)      # streams the pinned frames in chunks
            stream.synchronize()
            sec_e2e_ts = (time.perf_counter() - w0) / 3
        line["time_shap"] = {"sec_per_image": sec_dev, "e2e_sec_per_image": sec_e2e_ts, "frames": T_STEPS, "unit": "s/image",
                             "resnet18_img_per_s": T_STEPS / sec_dev, "tflops": T_STEPS * 3.627e9 / sec_dev / 1e12,
                             "h2d_bytes": T_STEPS * 3 * 128 * 128 * 4, "gpu_launches": clf.launch_count()}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import xai as oxai
            from oracle.classifier import build_classifier
            oc = build_classifier()
            fr = [traj_host[i:i + 1] for i in range(12)]
            oxai.time_shap(oc, fr[:2], [0, 1], 0)
            c0 = time.perf_counter()
            oxai.time_shap(oc, fr, list(range(12)), 0)
            per = (time.perf_counter() - c0) / 12
            line["time_shap"]["cpu_baseline"] = {"value": per * T_STEPS, "unit": "s/image", "cores": os.cpu_count(), "kind": "port",
                                                 "sample": "12 frames at the reference's call pattern (2 forwards/frame, B=1), extrapolated to 1000"}
    except Exception as e:                                  # the headline line must survive a classifier-side failure
        line["time_shap"] = {"error": str(e)}
    if world == 1 and not args.no_cpu_baseline:
        b_cpu, k_cpu = 2, 3
        sec, cores = cpu_oracle_steps(b_cpu, k_cpu, 1)
        line["cpu_baseline"] = {"value": b_cpu / (T_STEPS * sec), "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"{k_cpu} denoising steps at B={b_cpu} of the same workload (fp32 oracle port, "
                                          f"{sec:.2f} s/step)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
