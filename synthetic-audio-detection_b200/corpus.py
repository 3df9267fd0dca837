"""Corpus driver: a synthetic clip corpus (BASELINE.json configs[4]: 3.8 M four-second segments = 118 750 clips of
32 segments on average, README.md:77 of the reference) streamed through the fused path on 1..8 GPUs.

Every rank owns a contiguous block of WHOLE clips (sharded.clip_partition), generates the PCM of the chunk it is
about to process on its own GPU (synthetic.synth_pcm, counter based: the fp32 corpus would be 1.95 TB), reduces its
clips locally and takes part in ONE gather of per-clip results.  Because a segment's bytes depend only on its global
index and the kernels are batch independent, per-clip results are bit-identical for every GPU count."""
from __future__ import annotations

import hashlib
import time
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import sharded
from . import synthetic as S


def clip_lengths(n_clips: int, mean_segments: int = 32, seed: int = 118750) -> np.ndarray:
    """Ragged clip lengths (segments per clip) in [mean/2, 3*mean/2] whose sum is exactly n_clips * mean_segments."""
    rs = np.random.RandomState(seed)
    half = max(mean_segments // 2, 1)
    d = rs.randint(-half, half + 1, size=n_clips // 2)
    lens = np.full(n_clips, mean_segments, dtype=np.int64)
    lens[0:2 * (n_clips // 2):2] += d                      # pairs (+d, -d): the total stays n_clips * mean
    lens[1:2 * (n_clips // 2):2] -= d
    return lens


def run_corpus(engine, lengths: np.ndarray, chunk: int = 2048, threshold: float = 0.5, seed: int = S.BASE_SEED,
               group=None) -> Dict[str, object]:
    """Returns clip_probs [n_clips,N+1] / clip_labels [n_clips] (every rank), this rank's segment labels and timings
    measured on the device (CUDA events on the engine's stream; the gather is timed separately)."""
    dev = engine.device
    buf = torch.empty(chunk, S.SEGMENT, device=dev, dtype=torch.float32)

    def fetch(s_lo, s_hi):
        return S.synth_pcm(s_hi - s_lo, s_lo, dev, seed=seed, out=buf)

    def forward(pcm):
        _, probs, labels = engine.forward_pcm(pcm, threshold)
        return probs, labels

    t_gather = {}

    def clip_reduce(probs, cid, n):
        cp, cl = engine.clip_reduce(probs, cid.contiguous(), n, threshold)
        torch.cuda.synchronize(dev)
        t_gather["t0"] = time.perf_counter()
        return cp, cl

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    if dist.is_initialized():
        dist.barrier(group)
    e0.record()
    probs, labels, seg_labels = sharded.run_sharded(list(lengths), fetch, forward, clip_reduce, group, chunk=chunk)
    e1.record()
    torch.cuda.synchronize(dev)
    gather_s = time.perf_counter() - t_gather["t0"]
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
    return {"clip_probs": probs, "clip_labels": labels, "segment_labels": seg_labels, "ms": float(ms.item()),
            "gather_ms": 1e3 * gather_s}


def digest(clip_probs: torch.Tensor, clip_labels: torch.Tensor) -> str:
    """sha256 over the raw bytes of the per-clip results: equal digests <=> bit-identical results."""
    h = hashlib.sha256()
    h.update(clip_probs.detach().cpu().contiguous().numpy().tobytes())
    h.update(clip_labels.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()
