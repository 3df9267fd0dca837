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
def merged_sd(n_heads, backbone="resnet18"):
    return FX.merged_state_dict(n_heads, backbone=backbone)


_ENGINES = {}


def engine(n_heads, max_batch=8, backbone="resnet18", dtype=None):
    """One engine per (heads, batch, backbone, dtype) per process, weights of the seeded fixture loaded through the C ABI.
    dtype None = the engine's default (bf16 for resnet18/34, fp16 for the Bottleneck nets)."""
    from sad_b200.engine import Engine
    key = (n_heads, max_batch, backbone, dtype)
    if key not in _ENGINES:
        e = Engine(n_heads, torch.device("cuda", 0), max_batch=max_batch, backbone=backbone, dtype=dtype)
        e.load_merged_state_dict(merged_sd(n_heads, backbone))
        _ENGINES[key] = e
    return _ENGINES[key]


def segs(ids):
    return torch.cat([FX.synth_segments(1, first=int(i)) for i in ids])


def rel_db_err(a, b):
    """|a-b| / (1e-4 * max(|b|, 1)): <= 1 means inside the north-star log-mel tolerance."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / (1e-4 * np.maximum(np.abs(b), 1.0))


def decision_margin(logits):
    """Smallest logit perturbation that can change the decision of rule IR:207-213: the distance of any logit to the
    threshold (sigmoid(0) = 0.5) and, because the synthetic label is an arg-max, the gap between the two largest
    synthetic logits."""
    z = np.asarray(logits, np.float64)
    m = np.abs(z).min(axis=1)
    if z.shape[1] > 2:
        syn = np.sort(z[:, :-1], axis=1)
        m = np.minimum(m, syn[:, -1] - syn[:, -2])
    return m
