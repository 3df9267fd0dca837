"""Synthetic workload for benchmarks: noise/tone PCM segments generated ON the device and random-init merged
checkpoints of the named architecture (BinaryClassifier('resnet18') x N, layout of model_merger.py:154-159).
No datasets or trained checkpoints exist offline; bench.py says so in its `data` field."""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

SEGMENT = 128000
BASE_SEED = 20251018


def synth_pcm(n: int, first: int, device, seed: int = BASE_SEED) -> torch.Tensor:
    """[n,128000] fp32 in [-1,1] on `device`: a_n*N(0,1) + a_t*sin(2*pi*f*t + phi), clipped (SURVEY 8d recipe).
    Segment i's parameters depend only on (seed, first+i) through a counter-style generator per call."""
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1_000_003 + first)
    u = torch.rand(n, 5, generator=g, device=device)
    a_n = torch.exp(math.log(1e-3) + u[:, 0] * (math.log(0.2) - math.log(1e-3)))
    a_t = 0.5 * u[:, 1]
    f = torch.exp(math.log(50.0) + u[:, 2] * (math.log(11000.0) - math.log(50.0)))
    phi = 2 * math.pi * u[:, 3]
    kind = u[:, 4]
    a_t = torch.where(kind < 0.1, torch.zeros_like(a_t), a_t)
    a_n = torch.where((kind >= 0.1) & (kind < 0.2), torch.full_like(a_n, 1e-3), a_n)
    t = torch.arange(SEGMENT, device=device, dtype=torch.float32) / 32000.0
    out = torch.empty(n, SEGMENT, device=device, dtype=torch.float32)
    step = 256
    for i in range(0, n, step):
        j = min(n, i + step)
        x = torch.randn(j - i, SEGMENT, generator=g, device=device) * a_n[i:j, None]
        x += a_t[i:j, None] * torch.sin(2 * math.pi * f[i:j, None] * t[None, :] + phi[i:j, None])
        out[i:j] = x.clamp_(-1.0, 1.0)
    return out


def _conv_plan():
    plan = [("conv", "conv1", 64, 3, 7), ("bn", "bn1", 64)]
    for li, (cin, cout) in enumerate(((64, 64), (64, 128), (128, 256), (256, 512)), start=1):
        for b in range(2):
            p = f"layer{li}.{b}"
            c0 = cin if b == 0 else cout
            plan += [("conv", f"{p}.conv1", cout, c0, 3), ("bn", f"{p}.bn1", cout),
                     ("conv", f"{p}.conv2", cout, cout, 3), ("bn", f"{p}.bn2", cout)]
            if b == 0 and li > 1:
                plan += [("conv", f"{p}.downsample.0", cout, c0, 1), ("bn", f"{p}.downsample.1", cout)]
    return plan


def random_merged_state_dict(n_heads: int, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Random-init weights: kaiming-normal convs (fan_out), nn.Linear default ranges, BN gamma~U(.5,1.5),
    beta~N(0,.1), running_mean~N(0,.1), running_var~U(.5,1.5) so that folding is exercised."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def bn(key, c, g):
        sd[key + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[key + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[key + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[key + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[key + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)

    def lin(key, fin, fout, g):
        bound = 1.0 / math.sqrt(fin)
        sd[key + ".weight"] = (2 * torch.rand(fout, fin, generator=g) - 1) * bound
        sd[key + ".bias"] = (2 * torch.rand(fout, generator=g) - 1) * bound

    for h in range(n_heads):
        g = torch.Generator().manual_seed(seed * 1000 + h)
        p = f"sub_models.{h}."
        for item in _conv_plan():
            if item[0] == "conv":
                _, name, cout, cin, k = item
                sd[f"{p}base.{name}.weight"] = math.sqrt(2.0 / (cout * k * k)) * torch.randn(cout, cin, k, k, generator=g)
            else:
                bn(f"{p}base.{item[1]}", item[2], g)
        lin(p + "head.2", 512, 512, g)
        bn(p + "head.3", 512, g)
        lin(p + "head.6", 512, 256, g)
        bn(p + "head.7", 256, g)
        lin(p + "head.10", 256, 2, g)
    return sd


# FLOP model (SURVEY.md 8a / BASELINE.md section 2), per head per segment
def conv_flops_per_head_segment(folded_stem: bool = True):
    """List of (conv index in state_dict order, GFLOP).  Stem: 65536*64*K*2 with K=49 (channel-folded, what this
    build executes usefully) or K=147 (as the reference computes it)."""
    out = [(0, 2 * 65536 * 64 * (49 if folded_stem else 147) / 1e9)]
    idx = 1
    hw = {1: 128, 2: 64, 3: 32, 4: 16}
    for li, (cin, cout) in enumerate(((64, 64), (64, 128), (128, 256), (256, 512)), start=1):
        m = hw[li] * hw[li]
        for b in range(2):
            c0 = cin if b == 0 else cout
            out.append((idx, 2 * m * cout * c0 * 9 / 1e9)); idx += 1
            out.append((idx, 2 * m * cout * cout * 9 / 1e9)); idx += 1
            if b == 0 and li > 1:
                out.append((idx, 2 * m * cout * c0 / 1e9)); idx += 1
    return out
