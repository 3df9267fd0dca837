#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native Synthetic-Audio-Detection inference path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): 4-s 32 kHz segments/sec through (log-mel + merged ResNet-18 ensemble + decisions).
A "step" = one pass of the hot path over one batch of synthetic segments:
    workload = BASELINE.json configs[3]: bf16 tensor-core fused mel+ensemble, batch 2048, 6 heads, per GPU.
`value`  : inputs already resident in HBM (2048 x 128000 fp32 = 1.05 GB > the 126 MB L2, so no L2 flush is needed).
`e2e`    : the same batch through the C-ABI host entry (sad_forward_host): pinned host PCM -> H2D -> compute ->
           D2H of logits/probs/labels, every step.
`roofline`: the dominant kernel (tcgen05 implicit-GEMM convolution, all 20 conv launches per chunk): algorithmic
           FLOPs / CUDA-event time measured inside the timed region, against MEASURED_PEAKS.json.
`cpu_baseline`: the CPU oracle (torch-fp32 restatement of the reference) timed on this box's host cores on a
           bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "segments_per_sec_mel_plus_ensemble"
UNIT = "segments/s"
SEGMENT_BYTES = 128000 * 4
CLIP_SEGMENTS = 32            # clips of 32 segments; per-clip decisions are gathered


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=2048, help="segments per GPU per step")
    ap.add_argument("--heads", type=int, default=6)
    ap.add_argument("--max-batch", type=int, default=128, help="segments per internal pass (workspace size)")
    ap.add_argument("--cpu-sample", type=int, default=8, help="segments in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ingest", action="store_true", help="skip the ingest-stage measurement")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_throughput(sd, n_heads, n_seg, threads):
    """Oracle port (oracle/restatement.py == the reference's torch-CPU arithmetic) on `n_seg` segments."""
    import torch
    from oracle import fixtures as FX
    from oracle import restatement as R
    torch.set_num_threads(threads)
    x = FX.synth_segments(n_seg, first=0)

    def once():
        t0 = time.perf_counter()
        img = R.waveform_to_image(x)                               # front end, per-segment semantics
        logits = R.ensemble_forward(img.unsqueeze(1).repeat(1, 3, 1, 1), sd)
        R.interpret(logits, 0.5)
        return time.perf_counter() - t0

    once()                                                         # warm-up
    dt = min(once(), once())
    return n_seg / dt


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  The reference is pure Python over
    torch/torchaudio/torchvision; /root/reference does not exist on the GPU box, so the arm times the oracle port
    (bit-identical to the reference functions in the build container, tests/test_oracle_golden.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from sad_b200 import synthetic as S
    threads = os.cpu_count() or 1
    sd = S.random_merged_state_dict(args.heads, seed=0)
    n_seg = args.cpu_sample
    torch.set_num_threads(threads)
    times = []
    from oracle import fixtures as FX
    from oracle import restatement as R
    x = FX.synth_segments(n_seg, first=0)

    def step():
        img = R.waveform_to_image(x)
        logits = R.ensemble_forward(img.unsqueeze(1).repeat(1, 3, 1, 1), sd)
        R.interpret(logits, 0.5)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = n_seg * steps / dt
    sample = f"{n_seg} segments x {args.heads} heads per step, torch fp32, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"fused mel+ensemble, {args.heads} heads, CPU sample of {n_seg} segments per step "
                               f"(workload of the native arm: batch {args.batch} per GPU)", "heads": args.heads},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_native(args):
    import torch
    import torch.distributed as dist
    from sad_b200 import synthetic as S
    from sad_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    H, B, K, W = args.heads, args.batch, args.steps, max(args.warmup, 3)

    eng = Engine(H, dev, max_batch=args.max_batch)
    sd = S.random_merged_state_dict(H, seed=0)
    eng.load_merged_state_dict(sd)

    # this rank's shard of the step's segments: whole clips, contiguous
    x = S.synth_pcm(B, first=rank * B, device=dev)
    n_clips = B // CLIP_SEGMENTS
    clip_id = (torch.arange(B, device=dev, dtype=torch.int32) // CLIP_SEGMENTS).contiguous()
    gathered = [torch.empty(n_clips, H + 2, device=dev) for _ in range(world)] if world > 1 else None

    def step():
        logits, probs, labels = eng.forward_pcm(x, 0.5)
        cp, cl = eng.clip_reduce(probs, clip_id, n_clips, 0.5)
        if world > 1:   # the path's only exchange: per-clip decisions to every rank
            dist.all_gather(gathered, torch.cat([cp, cl.float().unsqueeze(1)], dim=1))
        return labels, cl

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.profile_enable(True)
    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.profile_enable(True)                    # resets the per-kernel counters
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    prof_ms, prof_n = eng.profile_read()
    eng.profile_enable(False)
    launches = eng.launches - launches0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * K / (ms / 1e3)

    # ---- end to end through the host entry of the C ABI --------------------------------------------------
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(B, 128000, dtype=torch.float32, pin_memory=True)
        xh.copy_(x)
        for _ in range(2):
            eng.forward_host(xh, 0.5)
        barrier()
        e0.record()
        for _ in range(K):
            lo, pr, la = eng.forward_host(xh, 0.5)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * K / (float(t.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": world * B * SEGMENT_BYTES, "d2h_bytes_per_step": world * B * ((H + 1) * 8 + 4)}
        del xh

    # ---- ingest stage: 44.1 kHz stereo int16 -> mono 32 kHz float32, stream larger than L2 -----------
    ingest = None
    if rank == 0 and not args.no_ingest:
        sr_in, ch, secs = 44100, 2, 1500
        pcm16 = torch.randint(-20000, 20000, (sr_in * secs, ch), dtype=torch.int16, device=dev)
        for _ in range(3):
            y = eng.ingest(pcm16, sr_in)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            y = eng.ingest(pcm16, sr_in)
        e1.record()
        torch.cuda.synchronize()
        ing_ms = e0.elapsed_time(e1) / 10
        ing_bytes = pcm16.numel() * 2 + y.numel() * 4
        ingest = {"ms": ing_ms, "bytes": ing_bytes, "seconds_of_audio": secs, "out_samples": int(y.numel())}
        del pcm16, y

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    flops = dict(S.conv_flops_per_head_segment(folded_stem=True))
    for c2 in (6, 11, 16):                      # downsample branch (index c2+1) runs inside conv2 when fused
        if prof_ms[c2 + 1] == 0:
            flops[c2] += flops[c2 + 1]
            flops[c2 + 1] = 0.0
    for c1 in (1, 3):                           # layer1 BasicBlocks run as ONE launch (block_rows.cu), timed under conv2
        if prof_ms[c1] == 0:
            flops[c1 + 1] += flops[c1]
            flops[c1] = 0.0
    conv_ms = sum(prof_ms[i] for i in range(20))
    conv_launches = sum(prof_n[i] for i in range(20))
    conv_tflop = sum(flops[i] for i in range(20)) * 1e9 * H * B * K / 1e12
    achieved = conv_tflop / (conv_ms / 1e3) if conv_ms > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    traffic = None                              # DRAM bytes per conv launch from the committed ncu --set full capture
    tp = os.path.join(ROOT, "profiles", "r01_roofline_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        if tj.get("chunk") == args.max_batch and tj.get("heads") == H:
            traffic = tj["dram_bytes_per_launch_mean"]
    n_conv_launches = sum(1 for i in range(20) if prof_n[i] > 0)
    # activation bytes per (head, segment): 40 MB with one launch per conv; a fused layer1 block keeps its
    # intermediate and residual on chip: -3 x 2.1 MB per block
    alg_mb = 40.0 - sum(6.3 for c1 in (1, 3) if prof_ms[c1] == 0)
    per_layer = {str(i): round(flops[i] * 1e9 * H * B * K / 1e12 / (prof_ms[i] / 1e3), 1) if prof_ms[i] > 0 else None
                 for i in range(20)}
    fe_ms = prof_ms[eng.PROF_FRONTEND]
    fe_gbs = (B * K * 512000 / 1e9) / (fe_ms / 1e3) if fe_ms > 0 else 0.0
    fe_tflops = (B * K * 16.0e6 / 1e12) / (fe_ms / 1e3) if fe_ms > 0 else 0.0
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12          # SMs x FMA lanes x 2 x max SM clock
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[3]: bf16 tensor-core fused mel+ensemble, batch 2048, 6 heads, "
                               "per GPU; per-clip decisions (32-segment clips) reduced and gathered",
                   "batch_per_gpu": B, "heads": H, "outputs": H + 1, "internal_chunk": args.max_batch,
                   "l2": "inputs larger than L2 (1.05 GB PCM per step, activations ~%d MB per chunk); no flush"
                         % int(args.max_batch * H * 7),
                   "weights": "random-init resnet18 x heads (seed 0)", "parallelism": f"segment-sharded x{world}"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "kernel": "conv_umma_kernel<128,2,TR> + conv_umma2_kernel<256> + block_rows_kernel + stem_fused_kernel "
                               "(%d launches per chunk)" % n_conv_launches,
                     "algorithmic_bytes_per_launch": (H * args.max_batch * alg_mb * 1e6) / n_conv_launches,
                     "flops_model": "18.1278 GFLOP/head/segment (channel-folded stem, K=49)",
                     "peak_source": pk["source"] + " bf16_tflops_sustained",
                     "kernel_ms_per_step": conv_ms / K, "kernel_launches": int(conv_launches),
                     "share_of_step": conv_ms / ms if ms > 0 else None,
                     "per_conv_tflops": per_layer},
        "roofline_frontend": {"bound": "hbm", "achieved": fe_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                              "frac": fe_gbs / pk["hbm_gbs"], "bytes_model": "512000 B/segment (PCM in; log-mel "
                              "stays on the device for the fused path)", "kernel_ms_per_step": fe_ms / K,
                              # the front end is above the CUDA-core ridge: compute-side view
                              "compute": {"flops_model": "16 MFLOP fp32/segment (126 packed 2048-pt FFTs + window, "
                                          "power, mel, log)", "achieved_tflops": fe_tflops,
                                          "fp32_peak_tflops": fp32_peak, "frac": fe_tflops / fp32_peak}},
        "roofline_ingest": None if ingest is None else {
            "bound": "hbm", "achieved": ingest["bytes"] / 1e9 / (ingest["ms"] / 1e3), "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": ingest["bytes"] / 1e9 / (ingest["ms"] / 1e3) / pk["hbm_gbs"],
            "workload": "sad_ingest: %d s of 44.1 kHz stereo int16 -> mono 32 kHz fp32 (mix, 18-tap polyphase sinc, pad); "
                        "%.0f MB in + out per launch (> L2), 10 launches" % (ingest["seconds_of_audio"], ingest["bytes"] / 1e6),
            "ms_per_launch": ingest["ms"],
            "audio_seconds_per_second": ingest["seconds_of_audio"] / (ingest["ms"] / 1e3)},
        "other_ms_per_step": {"image": prof_ms[eng.PROF_IMAGE] / K, "head_merge": prof_ms[eng.PROF_HEAD] / K},
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v = cpu_oracle_throughput(sd, H, args.cpu_sample, threads)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{args.cpu_sample} segments x {H} heads, oracle port (torch fp32), best of 2"}
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
