"""CPU emulation of the DEVICE data path's rounding points (TEST INFRASTRUCTURE ONLY).

Not the reference: this models what the sm_100a kernels compute -- BN folded into
the convolution in fp32, folded weights rounded once to bf16, activations stored
as bf16 between layers, fp32 accumulation, fp32 pooling / MLP / merge -- so that
GPU tests can localise a broken layer (device vs emulation agree to ~1e-2 relative
per layer) while parity itself is always judged against ``restatement`` (fp32).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from . import restatement as R


def fold_bn(w: torch.Tensor, sd: Dict[str, torch.Tensor], bn: str):
    """conv weight [Cout,...] + eval-mode BN -> (w', b') with y = conv(x, w') + b'."""
    g = sd[bn + ".weight"].double()
    b = sd[bn + ".bias"].double()
    m = sd[bn + ".running_mean"].double()
    v = sd[bn + ".running_var"].double()
    s = g / torch.sqrt(v + R.BN_EPS)
    wf = (w.double() * s.view(-1, *([1] * (w.dim() - 1)))).float()
    bf = (b - m * s).float()
    return wf, bf


_QT = [torch.bfloat16]          # storage type being emulated: bfloat16 (default build) or float16 (the fp16 build)


def set_storage_dtype(dt: torch.dtype) -> None:
    _QT[0] = dt


def _q(x: torch.Tensor) -> torch.Tensor:
    return x.to(_QT[0]).to(torch.float32)


def backbone_bf16(img1: torch.Tensor, sd: Dict[str, torch.Tensor], p: str, taps: List = None):
    """img1: [B,1,512,512] (or a general [B,3,512,512]) fp32; single-channel image (the three reference channels are identical,
    IR:173, so conv1's weights are summed over Cin).  Returns [B,512,16,16] fp32-valued bf16."""
    w, b = fold_bn(sd[p + "conv1.weight"], sd, p + "bn1")
    w1 = _q(w.sum(dim=1, keepdim=True)) if img1.shape[1] == 1 else _q(w)   # 3-channel input: general stem
    x = F.conv2d(_q(img1), w1, b, stride=2, padding=3)
    x = _q(F.relu(x))
    x = F.max_pool2d(x, 3, 2, 1)
    if taps is not None:
        taps.append(("stem", x))
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        for blk in range(64):
            q = f"{p}layer{li}.{blk}"
            if (q + ".conv1.weight") not in sd:
                break
            s = stride if blk == 0 else 1
            if (q + ".conv3.weight") in sd:      # Bottleneck: the projection is accumulated into conv3's fp32 tile (one rounding)
                w, b = fold_bn(sd[q + ".conv1.weight"], sd, q + ".bn1")
                o = _q(F.relu(F.conv2d(x, _q(w), b)))
                w, b = fold_bn(sd[q + ".conv2.weight"], sd, q + ".bn2")
                o = _q(F.relu(F.conv2d(o, _q(w), b, stride=s, padding=1)))
                w, b = fold_bn(sd[q + ".conv3.weight"], sd, q + ".bn3")
                o = F.conv2d(o, _q(w), b)
                if (q + ".downsample.0.weight") in sd:
                    wd, bd = fold_bn(sd[q + ".downsample.0.weight"], sd, q + ".downsample.1")
                    o = o + F.conv2d(x, _q(wd), bd, stride=s)
                else:
                    o = o + x
                x = _q(F.relu(o))
                if taps is not None:
                    taps.append((f"layer{li}.{blk}.conv3", x))
                continue
            w, b = fold_bn(sd[q + ".conv1.weight"], sd, q + ".bn1")
            o = _q(F.relu(F.conv2d(x, _q(w), b, stride=s, padding=1)))
            if taps is not None:
                taps.append((f"layer{li}.{blk}.conv1", o))
            if (q + ".downsample.0.weight") in sd:
                wd, bd = fold_bn(sd[q + ".downsample.0.weight"], sd, q + ".downsample.1")
                idn = _q(F.conv2d(x, _q(wd), bd, stride=s))
                if taps is not None:
                    taps.append((f"layer{li}.{blk}.downsample", idn))
            else:
                idn = x
            w, b = fold_bn(sd[q + ".conv2.weight"], sd, q + ".bn2")
            x = _q(F.relu(F.conv2d(o, _q(w), b, stride=1, padding=1) + idn))
            if taps is not None:
                taps.append((f"layer{li}.{blk}.conv2", x))
    return x


def head_fp32(feats: torch.Tensor, sd: Dict[str, torch.Tensor], p: str) -> torch.Tensor:
    """avg-pool + folded (Linear,BN1d) x2 + Linear, all fp32."""
    v = feats.mean(dim=(2, 3))
    w, b = fold_bn(sd[p + "2.weight"], sd, p + "3")
    b = b + (sd[p + "2.bias"].double() * (sd[p + "3.weight"].double()
             / torch.sqrt(sd[p + "3.running_var"].double() + R.BN_EPS))).float()
    v = F.relu(F.linear(v, w, b))
    w, b = fold_bn(sd[p + "6.weight"], sd, p + "7")
    b = b + (sd[p + "6.bias"].double() * (sd[p + "7.weight"].double()
             / torch.sqrt(sd[p + "7.running_var"].double() + R.BN_EPS))).float()
    v = F.relu(F.linear(v, w, b))
    return F.linear(v, sd[p + "10.weight"], sd[p + "10.bias"])


def ensemble_bf16(img1: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """[B,1,512,512] -> merged [B,N+1] logits through the emulated device data path."""
    outs = []
    with torch.no_grad():
        for i in R.head_indices(sd):
            p = f"sub_models.{i}."
            outs.append(head_fp32(backbone_bf16(img1, sd, p + "base."), sd, p + "head."))
    return R.merge_logits(torch.stack(outs, dim=1))
