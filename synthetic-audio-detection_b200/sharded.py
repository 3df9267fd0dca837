"""Segment-sharded multi-GPU driver.  One process per GPU; whole clips are assigned to ranks in
contiguous blocks so the per-clip reduction is rank-local; weights are replicated; the ONLY exchange on the path is
one all_gather of per-clip results (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def clip_partition(n_clips: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) clip ranges per rank: clip c -> rank c // ceil(n_clips / world)."""
    per = -(-n_clips // world) if n_clips > 0 else 0
    return [(min(r * per, n_clips), min((r + 1) * per, n_clips)) for r in range(world)]


def segment_range(clip_lengths: Sequence[int], lo: int, hi: int) -> Tuple[int, int]:
    """Global segment index range [s_lo, s_hi) covered by clips [lo, hi)."""
    s_lo = int(sum(clip_lengths[:lo]))
    return s_lo, s_lo + int(sum(clip_lengths[lo:hi]))


def gather_clip_results(local_probs: torch.Tensor, local_labels: torch.Tensor, n_clips: int, group=None):
    """All ranks contribute their [n_local, N+1] clip probabilities and [n_local] labels; every rank returns the
    full [n_clips, N+1] / [n_clips] in clip order.  Ranks may own different numbers of clips (ragged tail)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_probs, local_labels
    rank = dist.get_rank(group)
    parts = clip_partition(n_clips, world)
    per = max(hi - lo for lo, hi in parts)
    n1 = local_probs.shape[1]
    buf = torch.zeros(per, n1 + 1, device=local_probs.device, dtype=torch.float32)
    n_local = parts[rank][1] - parts[rank][0]
    assert local_probs.shape[0] == n_local == local_labels.shape[0], "rank does not hold its clip block"
    buf[:n_local, :n1] = local_probs
    buf[:n_local, n1] = local_labels.to(torch.float32)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    probs = torch.cat([o[:hi - lo, :n1] for o, (lo, hi) in zip(out, parts)])
    labels = torch.cat([o[:hi - lo, n1] for o, (lo, hi) in zip(out, parts)]).to(torch.int32)
    return probs, labels


def run_sharded(clip_lengths: Sequence[int], fetch_segments: Callable[[int, int], torch.Tensor],
                forward: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                clip_reduce: Callable[[torch.Tensor, torch.Tensor, int], Tuple[torch.Tensor, torch.Tensor]],
                group=None, chunk: Optional[int] = None):
    """clip_lengths[c] = number of (non-silent) segments of clip c, known to every rank.
    fetch_segments(s_lo, s_hi) -> PCM [s_hi-s_lo,128000] of GLOBAL segments s_lo..s_hi-1 (this rank's shard, or one
    `chunk` of it: a shard of the 3.8 M-segment corpus is 243 GB as fp32 and is streamed, never held);
    forward(pcm) -> (probs [n,N+1], labels [n]); clip_reduce(probs, local_clip_id, n_local_clips) ->
    (clip_probs, clip_labels).
    Returns (clip_probs [n_clips,N+1], clip_labels [n_clips]) on every rank, plus this rank's segment labels."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_clips = len(clip_lengths)
    lo, hi = clip_partition(n_clips, world)[rank]
    s_lo, s_hi = segment_range(clip_lengths, lo, hi)
    step = (s_hi - s_lo) if not chunk else int(chunk)
    probs_parts, label_parts = [], []
    for a in range(s_lo, s_hi, max(step, 1)):
        p, l = forward(fetch_segments(a, min(a + step, s_hi)))
        probs_parts.append(p)
        label_parts.append(l)
    if probs_parts:
        probs = probs_parts[0] if len(probs_parts) == 1 else torch.cat(probs_parts)
        seg_labels = label_parts[0] if len(label_parts) == 1 else torch.cat(label_parts)
    else:                                   # a rank without clips still takes part in the gather
        probs, seg_labels = forward(fetch_segments(s_lo, s_lo))
    local_id = torch.repeat_interleave(torch.arange(hi - lo, dtype=torch.int32),
                                       torch.as_tensor(list(clip_lengths[lo:hi]), dtype=torch.int64))
    if probs.is_cuda:                       # page-locked source: the copy is stream-ordered and does not stall the host
        local_id = local_id.pin_memory().to(probs.device, non_blocking=True)
    cp, cl = clip_reduce(probs, local_id, hi - lo)
    all_p, all_l = gather_clip_results(cp, cl, n_clips, group)
    return all_p, all_l, seg_labels
