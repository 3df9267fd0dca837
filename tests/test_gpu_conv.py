"""GPU parity: every convolution layer of the trunk (K3, tcgen05 implicit GEMM) in isolation.

The device kernel consumes bf16 activations and bf16 folded weights with fp32 accumulation; the check feeds the
SAME bf16-rounded operands to torch's fp32 conv2d on the CPU, so the only differences are accumulation order and
the final bf16 rounding of the output (<= 2^-8 relative).
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import bf16_emulation as E
from oracle import fixtures as FX
from tests import gpu_common as G

pytestmark = pytest.mark.gpu

LAYERS = [l for l in FX._layer_plan() if l[0] == "conv"]      # 20 convs in state_dict order
BNS = [l for l in FX._layer_plan() if l[0] == "bn"]


def _geometry(idx):
    name = LAYERS[idx][1]
    cout, cin, k = LAYERS[idx][2], LAYERS[idx][3], LAYERS[idx][4]
    li = int(name.split(".")[0][-1])
    hout = {1: 128, 2: 64, 3: 32, 4: 16}[li]
    stride = 2 if (name.endswith(".0.conv1") and li > 1) or "downsample" in name else 1
    return name, cin, cout, k, stride, hout * stride, hout


@pytest.mark.parametrize("idx", list(range(1, 20)))
@pytest.mark.parametrize("head", [1])
def test_conv_layer(idx, head):
    name, cin, cout, k, stride, hin, hout = _geometry(idx)
    sd = G.merged_sd(2)
    p = f"sub_models.{head}.base."
    w, b = E.fold_bn(sd[p + name + ".weight"], sd, p + BNS[idx][1])
    wq = w.to(torch.bfloat16).float()
    B = 3
    g = torch.Generator().manual_seed(100 + idx)
    x = torch.randn(B, cin, hin, hin, generator=g).to(torch.bfloat16)
    use_res = name.endswith("conv2")
    res = torch.randn(B, cout, hout, hout, generator=g).to(torch.bfloat16) if use_res else None
    relu = "downsample" not in name
    want = F.conv2d(x.float(), wq, b, stride=stride, padding=k // 2)
    if use_res:
        want = want + res.float()
    if relu:
        want = F.relu(want)
    e = G.engine(2)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    r_nhwc = res.permute(0, 2, 3, 1).contiguous().cuda() if use_res else None
    got = e.debug_conv(head, idx, x_nhwc, r_nhwc, (B, hout, hout, cout), relu)
    torch.cuda.synchronize()
    got = got.float().cpu().permute(0, 3, 1, 2)
    err = (got - want).abs()
    tol = 2.0 ** -7 * want.abs() + 2e-2        # bf16 output rounding (2^-8 rel) with margin + accumulation slack
    bad = (err > tol).float().mean().item()
    print(f"{name}: max abs err {err.max():.4f} (|want| max {want.abs().max():.2f}), frac out of tol {bad:.2e}")
    assert bad == 0.0
