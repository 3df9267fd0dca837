#!/usr/bin/env python3
"""Batch-folder driver (precedent: legacy/source/inference_script.py:428-451 --IsBatch): analyse every
WAV of a folder and write one JSON per clip, optionally sharded over the GPUs of a box.

    python synthetic-audio-detection_b200/batch_runner.py --merged-model M --folder F --out-dir O [--smooth]
    torchrun --nproc-per-node 8 synthetic-audio-detection_b200/batch_runner.py ...      (clips split across ranks)

Files are assigned to ranks in contiguous blocks of the sorted listing (the same rule as sharded.clip_partition);
each rank writes its own JSON files and rank 0 additionally writes summary.json with every clip's percentages and
clip-level label (gathered with one all_gather_object at the end)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import sad_b200  # noqa: F401
    from sad_b200 import inference_runner as IR
    from sad_b200.sharded import clip_partition
else:
    from . import inference_runner as IR
    from .sharded import clip_partition

AUDIO_EXT = (".wav",)


def list_audio(folder: str):
    return sorted(os.path.join(folder, f) for f in os.listdir(folder) if f.lower().endswith(AUDIO_EXT))


def clip_label(percentages: dict, class_names, threshold: float) -> str:
    """Rule IR:207-213 applied to the clip-mean probabilities (the reference emits no clip label)."""
    if not percentages:
        return ""
    p = np.array([percentages[n] / 100.0 for n in class_names], dtype=np.float32)
    n = len(class_names) - 1
    if p[-1] >= threshold and (p[:n] < threshold).all():
        return class_names[-1]
    return class_names[int(np.argmax(p[:n]))]


def main(argv=None):
    ap = argparse.ArgumentParser(description="Batch-folder multi-head inference (one JSON per clip).")
    ap.add_argument("--merged-model", required=True)
    ap.add_argument("--folder", required=True)
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--threshold", type=float, default=0.5)
    ap.add_argument("--overlap", type=float, default=0.0)
    ap.add_argument("--silence-threshold", type=float, default=1e-3)
    ap.add_argument("--smooth", action="store_true")
    args = ap.parse_args(argv)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)

    files = list_audio(args.folder)
    lo, hi = clip_partition(len(files), world)[rank]
    model, meta = IR.load_merged_model(args.merged_model, device)
    names = meta["class_names"]
    cfg = IR.AudioConfig(32000, 4.0, args.overlap, args.silence_threshold)
    os.makedirs(args.out_dir, exist_ok=True)
    mine = []
    for path in files[lo:hi]:
        wf, sr = IR.preprocess_waveform(path, cfg)
        res = IR.analyze_waveform(model, wf, sr, names, cfg, args.threshold, args.smooth, device)
        out = {"filename": path, "segments": res["segments"], "percentages": res["percentages"]}
        with open(os.path.join(args.out_dir, os.path.splitext(os.path.basename(path))[0] + ".json"), "w") as f:
            json.dump(out, f, indent=4)
        mine.append({"filename": path, "percentages": res["percentages"], "n_segments": len(res["segments"]),
                     "label": clip_label(res["percentages"], names, args.threshold)})
    allr = [mine]
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    if rank == 0:
        with open(os.path.join(args.out_dir, "summary.json"), "w") as f:
            json.dump([r for part in allr for r in part], f, indent=2)
        print(f"{len(files)} clips -> {args.out_dir}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
