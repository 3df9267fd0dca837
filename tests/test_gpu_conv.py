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

LAYERS = [l for l in FX._layer_plan("resnet18") if l[0] == "conv"]      # 20 convs in state_dict order
BNS = [l for l in FX._layer_plan("resnet18") if l[0] == "bn"]


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


def test_fused_stem_and_pool():
    """K2 (conv 7x7/2 + BN + ReLU + maxpool 3x3/2 fused, 1-channel folded stem) vs torch fp32 conv/pool fed with the
    SAME bf16 image and bf16 folded weights the device uses."""
    from oracle import restatement as R
    n_heads = 2
    sd = G.merged_sd(n_heads)
    e = G.engine(n_heads)
    x = FX.synth_segments(3, first=700)
    got = e.debug_stem(x.cuda())
    img = e.debug_read(0, (3, 512, 512), torch.bfloat16)
    torch.cuda.synchronize()
    img = img.float().cpu()
    want_img = R.waveform_to_image(x)
    assert (img - want_img).abs().max() <= 2.0 ** -8 * want_img.abs().max() + 1e-3      # bf16 rounding of the image
    got = got.float().cpu().view(n_heads, 3, 128, 128, 64)
    for h in range(n_heads):
        p = f"sub_models.{h}.base."
        w, b = E.fold_bn(sd[p + "conv1.weight"], sd, p + "bn1")
        w1 = w.sum(dim=1, keepdim=True).to(torch.bfloat16).float()
        y = F.max_pool2d(F.relu(F.conv2d(img.unsqueeze(1), w1, b, stride=2, padding=3)), 3, 2, 1)
        want = y.permute(0, 2, 3, 1)
        err = (got[h] - want).abs()
        tol = 2.0 ** -7 * want.abs() + 2e-2
        print(f"stem head {h}: max abs err {err.max():.4f} (|want| max {want.abs().max():.2f}), "
              f"frac out of tol {(err > tol).float().mean():.2e}")
        assert (err > tol).float().mean() == 0.0


@pytest.mark.parametrize("first", [1, 3])
@pytest.mark.parametrize("B", [1, 5])
def test_fused_basic_block_equals_two_launches(first, B):
    """K3c (block_rows.cu): conv1 -> conv2 + identity on a CTA pair through peer shared memory must reproduce the
    two single-conv launches BIT FOR BIT (same bf16 rounding of the intermediate, same accumulation order), and both
    must agree with torch fp32 on the same bf16 operands."""
    head = 1
    e = G.engine(2)
    g = torch.Generator().manual_seed(500 + first + B)
    x = (torch.randn(B, 128, 128, 64, generator=g) * 0.7).to(torch.bfloat16).cuda()
    mid = e.debug_conv(head, first, x, None, (B, 128, 128, 64), True)
    two = e.debug_conv(head, first + 1, mid, x, (B, 128, 128, 64), True)
    one = e.debug_block(head, first, x)
    torch.cuda.synchronize()
    assert torch.equal(one.view(torch.int16), two.view(torch.int16)), \
        f"fused block differs from two launches: {(one.float() - two.float()).abs().max().item()}"
    # independent check of the fused result (image borders included) against torch fp32 on the CPU
    sd = G.merged_sd(2)
    p = f"sub_models.{head}.base."
    xc = x.float().cpu().permute(0, 3, 1, 2)
    w1, b1 = E.fold_bn(sd[p + LAYERS[first][1] + ".weight"], sd, p + BNS[first][1])
    w2, b2 = E.fold_bn(sd[p + LAYERS[first + 1][1] + ".weight"], sd, p + BNS[first + 1][1])
    m = F.relu(F.conv2d(xc, w1.to(torch.bfloat16).float(), b1, padding=1)).to(torch.bfloat16).float()
    want = F.relu(F.conv2d(m, w2.to(torch.bfloat16).float(), b2, padding=1) + xc)
    got = one.float().cpu().permute(0, 3, 1, 2)
    err = (got - want).abs()
    tol = 2.0 ** -7 * want.abs() + 3e-2        # one extra bf16 rounding (the intermediate) can flip by one ulp
    assert (err > tol).float().mean().item() < 1e-5, f"max err {err.max():.4f}"


def test_fused_block_switch_gives_identical_logits(monkeypatch):
    """Whole path with SAD_FUSE_BLOCK=0 and =1: identical logits, and the fused plan launches two kernels fewer."""
    from sad_b200.engine import Engine
    x = FX.synth_segments(5, first=40).cuda()
    outs, launches = [], []
    for flag in ("0", "1"):
        monkeypatch.setenv("SAD_FUSE_BLOCK", flag)
        eng = Engine(2, torch.device("cuda", 0), max_batch=4)
        eng.load_merged_state_dict(G.merged_sd(2))
        n0 = eng.launches
        outs.append(eng.forward_pcm(x, 0.5)[0].cpu())
        launches.append(eng.launches - n0)
        eng.close()
    assert torch.equal(outs[0], outs[1])
    assert launches[0] - launches[1] == 2 * 2                      # two chunks (4 + 1 segments) x two blocks


R50_LAYERS = [l for l in FX._layer_plan("resnet50") if l[0] == "conv"]
R50_BNS = [l for l in FX._layer_plan("resnet50") if l[0] == "bn"]


def _geometry50(idx):
    name, cout, cin, k = R50_LAYERS[idx][1], R50_LAYERS[idx][2], R50_LAYERS[idx][3], R50_LAYERS[idx][4]
    li = int(name.split(".")[0][-1])
    hout = {1: 128, 2: 64, 3: 32, 4: 16}[li]
    first = name.split(".")[1] == "0" and li > 1
    stride = 2 if first and (name.endswith("conv2") or "downsample" in name) else 1
    hin = hout * 2 if first and (name.endswith("conv1") or stride == 2) else hout
    if first and name.endswith("conv1"):
        hout = hin                                   # the 1x1 entry conv of a stage still runs at the input resolution
    return name, cin, cout, k, stride, hin, hout


# one of every Bottleneck conv shape: 1x1 64->64 / 64->256 / 256->64 @128, the layer2 entry block (1x1 256->128 @128,
# 3x3/2, 1x1 128->512), a layer3 1x1 1024->256, the layer4 3x3/2 and the widest 1x1 (512->2048, 2048->512), and the
# conv3 + identity of non-entry blocks at every width (7: 64->256 @128, 17: 128->512 @64, 30: 256->1024 @32, 52: 512->2048 @16)
@pytest.mark.parametrize("idx", [1, 2, 3, 4, 5, 7, 11, 12, 13, 14, 17, 28, 30, 44, 45, 47, 48, 52])
def test_bottleneck_conv_layer(idx):
    name, cin, cout, k, stride, hin, hout = _geometry50(idx)
    head = 1
    sd = G.merged_sd(2, "resnet50")
    p = f"sub_models.{head}.base."
    w, b = E.fold_bn(sd[p + name + ".weight"], sd, p + R50_BNS[idx][1])
    e = G.engine(2, backbone="resnet50")                    # Bottleneck nets default to the fp16 build (csrc/act.cuh)
    qt = e.act_dtype
    wq = w.to(qt).float()
    B = 2
    g = torch.Generator().manual_seed(700 + idx)
    x = torch.randn(B, cin, hin, hin, generator=g).to(qt)
    use_res = name.endswith("conv3") and name.split(".")[1] != "0"
    res = torch.randn(B, cout, hout, hout, generator=g).to(qt) if use_res else None
    relu = "downsample" not in name
    want = F.conv2d(x.float(), wq, b, stride=stride, padding=k // 2)
    if use_res:
        want = want + res.float()
    if relu:
        want = F.relu(want)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    r_nhwc = res.permute(0, 2, 3, 1).contiguous().cuda() if use_res else None
    got = e.debug_conv(head, idx, x_nhwc, r_nhwc, (B, hout, hout, cout), relu)
    torch.cuda.synchronize()
    got = got.float().cpu().permute(0, 3, 1, 2)
    err = (got - want).abs()
    tol = 2.0 ** -7 * want.abs() + 2e-2
    bad = (err > tol).float().mean().item()
    print(f"resnet50 {name} [{cin}->{cout} k{k} s{stride} @{hin}]: max abs err {err.max():.4f}, frac out of tol {bad:.2e}")
    assert bad == 0.0


# ---------------------------------------------------------------- layer1 on CTA pairs (conv_rows2.cu, SAD_ROWS2)
def _engine_rows2(monkeypatch, mode, n_heads=2, max_batch=8):
    """A private engine created with SAD_ROWS2 set (the switch is read at sad_create time)."""
    from sad_b200.engine import Engine
    monkeypatch.setenv("SAD_ROWS2", str(mode))
    e = Engine(n_heads, torch.device("cuda", 0), max_batch=max_batch)
    e.load_merged_state_dict(G.merged_sd(n_heads))
    return e


@pytest.mark.parametrize("idx", [1, 2, 3, 4])
def test_layer1_conv_on_cta_pairs(idx, monkeypatch):
    """The four layer1 convolutions through conv_rows2_kernel (cta_group::2, UMMA 256x192x16, aliased accumulator ring
    with two partial accumulators for 2 of every 6 rows, phantom rows at the strip borders), odd and even image counts."""
    name, cin, cout, k, stride, hin, hout = _geometry(idx)
    head = 1
    sd = G.merged_sd(2)
    p = f"sub_models.{head}.base."
    w, b = E.fold_bn(sd[p + name + ".weight"], sd, p + BNS[idx][1])
    wq = w.to(torch.bfloat16).float()
    e = _engine_rows2(monkeypatch, 1)
    for B in (1, 4):
        g = torch.Generator().manual_seed(900 + idx + B)
        x = torch.randn(B, cin, hin, hin, generator=g).to(torch.bfloat16)
        use_res = name.endswith("conv2")
        res = torch.randn(B, cout, hout, hout, generator=g).to(torch.bfloat16) if use_res else None
        want = F.conv2d(x.float(), wq, b, stride=1, padding=1)
        if use_res:
            want = want + res.float()
        want = F.relu(want)
        got = e.debug_conv(head, idx, x.permute(0, 2, 3, 1).contiguous().cuda(),
                           res.permute(0, 2, 3, 1).contiguous().cuda() if use_res else None, (B, hout, hout, cout), True)
        torch.cuda.synchronize()
        err = (got.float().cpu().permute(0, 3, 1, 2) - want).abs()
        bad = (err > 2.0 ** -7 * want.abs() + 2e-2).float().mean().item()
        # every output row, also the split ones (rows 4, 5 mod 6 of a strip) and the strip borders (rows 0, 63, 64, 127)
        print(f"conv_rows2 {name} B={B}: max abs err {err.max():.4f}, worst row {int(err.amax(dim=(0, 1, 3)).argmax())}")
        assert bad == 0.0
    e.close()


def test_whole_path_on_cta_pairs_matches_reference_and_is_batch_independent(monkeypatch):
    """SAD_ROWS2=2: both layer1 BasicBlocks as two conv_rows2 launches each instead of the fused block_rows kernel.
    Logits vs the reference golden (2e-2), close to the default engine (same arithmetic up to fp32 summation order of
    the split rows), and bit-identical for a segment whatever the batch around it."""
    g = G.golden("ensemble_n2.npz")
    x = G.segs(g["seg_ids"]).cuda()
    e2 = _engine_rows2(monkeypatch, 2)
    lo2, _, la2 = e2.forward_pcm(x, 0.5)
    d = (lo2.cpu().numpy() - g["merged_logits"])
    lo0, _, _ = G.engine(2).forward_pcm(x, 0.5)
    print(f"rows2: max |logit diff| vs reference {abs(d).max():.4f}; vs the fused-block engine {(lo2 - lo0).abs().max().item():.5f}")
    assert abs(d).max() <= 2e-2
    assert (lo2 - lo0).abs().max().item() <= 5e-3
    sub = x[[5, 2]].contiguous()
    lo_sub, _, _ = e2.forward_pcm(sub, 0.5)
    assert torch.equal(lo_sub[0], lo2[5]) and torch.equal(lo_sub[1], lo2[2])
    e2.close()


@pytest.mark.parametrize("idx", [2, 6, 9, 14, 19])
def test_conv_layer_fp16_build(idx):
    """One convolution per kernel family on the fp16 build of the library (csrc/act.cuh): layer1 rows (2), layer2 with the
    folded downsample (6) and with the identity as K blocks (9), the 2-CTA layers 3 and 4 (14, 19).  Operands rounded to
    fp16 on both sides, so only accumulation order and the fp16 rounding of the output differ."""
    name, cin, cout, k, stride, hin, hout = _geometry(idx)
    head = 1
    sd = G.merged_sd(2)
    p = f"sub_models.{head}.base."
    w, b = E.fold_bn(sd[p + name + ".weight"], sd, p + BNS[idx][1])
    wq = w.to(torch.float16).float()
    B = 2
    g = torch.Generator().manual_seed(300 + idx)
    x = torch.randn(B, cin, hin, hin, generator=g).to(torch.float16)
    use_res = name.endswith("conv2")
    res = torch.randn(B, cout, hout, hout, generator=g).to(torch.float16) if use_res else None
    want = F.conv2d(x.float(), wq, b, stride=stride, padding=k // 2)
    if use_res:
        want = want + res.float()
    want = F.relu(want)
    e = G.engine(2, dtype="fp16")
    got = e.debug_conv(head, idx, x.permute(0, 2, 3, 1).contiguous().cuda(),
                       res.permute(0, 2, 3, 1).contiguous().cuda() if use_res else None, (B, hout, hout, cout), True)
    torch.cuda.synchronize()
    assert got.dtype == torch.float16
    err = (got.float().cpu().permute(0, 3, 1, 2) - want).abs()
    tol = 2.0 ** -10 * want.abs() + 3e-3        # fp16 output rounding (2^-11 rel) with margin + accumulation slack
    bad = (err > tol).float().mean().item()
    print(f"fp16 {name}: max abs err {err.max():.5f} (|want| max {want.abs().max():.2f}), frac out of tol {bad:.2e}")
    assert bad == 0.0
