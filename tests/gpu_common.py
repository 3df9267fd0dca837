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
    """Smallest max-norm logit perturbation that changes the decision of rule IR:207-213 at threshold 0.5 (sigmoid(0)).
    Real (z_real >= 0, every z_syn < 0): any logit reaching 0 flips it -> min |z|.  Otherwise the label is the arg-max
    of the synthetic logits: it changes when the two largest meet (each moves half their gap) or when the
    row becomes Real, which needs EVERY violated condition repaired -> the largest violation."""
    z = np.asarray(logits, np.float64)
    syn, real = z[:, :-1], z[:, -1]
    is_real = (real >= 0) & (syn < 0).all(axis=1)
    m_real = np.abs(z).min(axis=1)
    to_real = np.maximum(np.maximum(-real, 0), np.maximum(syn.max(axis=1), 0))
    if syn.shape[1] > 1:
        srt = np.sort(syn, axis=1)
        gap = 0.5 * (srt[:, -1] - srt[:, -2])
    else:
        gap = np.full(z.shape[0], np.inf)
    return np.where(is_real, m_real, np.minimum(gap, to_real))
