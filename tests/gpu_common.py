"""Shared helpers for the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import functools
import os

import numpy as np
import torch

from oracle import fixtures as FX

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@functools.lru_cache(maxsize=None)
def merged_sd(n_heads):
    return FX.merged_state_dict(n_heads)


_ENGINES = {}


def engine(n_heads, max_batch=8):
    """One engine per (heads, batch) per process, weights of the seeded fixture loaded through the C ABI."""
    from sad_b200.engine import Engine
    key = (n_heads, max_batch)
    if key not in _ENGINES:
        e = Engine(n_heads, torch.device("cuda", 0), max_batch=max_batch)
        e.load_merged_state_dict(merged_sd(n_heads))
        _ENGINES[key] = e
    return _ENGINES[key]


def segs(ids):
    return torch.cat([FX.synth_segments(1, first=int(i)) for i in ids])


def rel_db_err(a, b):
    """|a-b| / (1e-4 * max(|b|, 1)): <= 1 means inside the north-star log-mel tolerance."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / (1e-4 * np.maximum(np.abs(b), 1.0))
