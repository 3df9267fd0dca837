"""Multi-rank host logic of the segment-sharded driver on CPU: world_size 2 over gloo.  The per-rank
compute is replaced by a deterministic stand-in so only partitioning + the single gather are exercised."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restatement as R
from sad_b200 import sharded as S


def test_clip_partition_properties():
    for n in (0, 1, 7, 64, 118750):
        for w in (1, 2, 4, 8):
            parts = S.clip_partition(n, w)
            assert len(parts) == w and parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(s for s in sizes if s or True) <= max(sizes)   # contiguous, ragged tail allowed
    assert S.segment_range([3, 1, 4, 1, 5], 1, 4) == (3, 9)


def _fake_probs(s_lo, s_hi, n1):
    g = torch.Generator().manual_seed(1234)
    allp = torch.rand(10000, n1, generator=g)
    return allp[s_lo:s_hi]


def _worker(rank, world, port, clip_lengths, n1, out_dir, chunk=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def fetch(s_lo, s_hi):
        return torch.arange(s_lo, s_hi).float().unsqueeze(1)            # stand-in "pcm": carries the segment index

    def forward(pcm):
        idx = pcm[:, 0].long()
        p = _fake_probs(int(idx[0]) if len(idx) else 0, int(idx[-1]) + 1 if len(idx) else 0, n1)
        labels = torch.tensor([R.decide_from_probs(r, 0.5) for r in p.numpy()], dtype=torch.int32)
        return p, labels

    def clip_reduce(probs, cid, n):
        cp, cl = R.clip_aggregate(probs.numpy(), cid.numpy(), n, 0.5)
        return torch.from_numpy(cp), torch.from_numpy(cl)

    allp, alll, seg = S.run_sharded(clip_lengths, fetch, forward, clip_reduce, chunk=chunk)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), p=allp.numpy(), l=alll.numpy(), nseg=len(seg))
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("chunk", [None, 2])
def test_world2_gloo_gathers_all_clips(tmp_path, chunk):
    """chunk=2: the shard is streamed in pieces that cut across clip boundaries (the corpus driver's mode)."""
    clip_lengths = [3, 1, 4, 1, 5, 9, 2]           # 7 clips -> rank 0 owns 4, rank 1 owns 3 (ragged)
    n1 = 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, clip_lengths, n1, str(tmp_path), chunk), nprocs=2, join=True)
    total = sum(clip_lengths)
    probs = _fake_probs(0, total, n1).numpy()
    cid = np.repeat(np.arange(len(clip_lengths)), clip_lengths)
    want_p, want_l = R.clip_aggregate(probs, cid, len(clip_lengths), 0.5)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    for r in (r0, r1):
        np.testing.assert_allclose(r["p"], want_p, rtol=0, atol=1e-7)
        np.testing.assert_array_equal(r["l"], want_l)
    assert int(r0["nseg"]) + int(r1["nseg"]) == total


def test_corpus_clip_lengths_are_ragged_prefix_stable_and_exact():
    """configs[4]: 118 750 ragged clips whose lengths sum to exactly 3.8 M segments; a subset corpus is a PREFIX of the
    full one (what lets the 1/16-subset runs be compared with the full run bit for bit)."""
    from sad_b200 import corpus as CO
    full = CO.clip_lengths(118750)
    assert int(full.sum()) == 3_800_000 and full.min() >= 16 and full.max() <= 48 and len(set(full.tolist())) > 10
    sub = CO.clip_lengths(7422)
    np.testing.assert_array_equal(sub, full[:7422])
    assert int(sub.sum()) == 7422 * 32
    odd = CO.clip_lengths(7)
    assert len(odd) == 7 and int(odd.sum()) == 7 * 32
