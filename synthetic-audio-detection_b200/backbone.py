"""Parameter container with timm's ResNet key layout (conv1, bn1, layer{1..4}.{0,1}.{conv1,bn1,conv2,bn2,
downsample.{0,1}}), so state_dicts written by the reference (model_merger.py:154-159) load unchanged.

The reference builds this trunk with ``timm.create_model(name, pretrained=True, num_classes=0)``
(inference_runner.py:35), which needs the network; here the modules are only parameter holders -- the arithmetic
runs in the sm_100a kernels -- so construction is offline and cheap.  resnet18 / resnet34 (BasicBlock) and
resnet50 / resnet101 / resnet152 (Bottleneck, 1x1 - 3x3 - 1x1 with the stride on the 3x3 as timm builds them) are all
wired to kernels.
"""
import torch
import torch.nn as nn

DEPTHS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3),
          "resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}
BOTTLENECK = ("resnet50", "resnet101", "resnet152")
SUPPORTED = tuple(DEPTHS)


class _BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.act1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.act2 = nn.ReLU(inplace=True)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, planes, stride):
        super().__init__()
        cout = planes * self.expansion
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.act1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.act2 = nn.ReLU(inplace=True)
        self.conv3 = nn.Conv2d(planes, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.act3 = nn.ReLU(inplace=True)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


class ResNetTrunk(nn.Module):
    """ResNet feature trunk (resnet18/34: BasicBlock, resnet50/101/152: Bottleneck); ``forward_features`` is served by
    the CUDA engine of the owning classifier."""

    def __init__(self, model_name: str = "resnet18"):
        super().__init__()
        if model_name not in SUPPORTED:
            raise NotImplementedError(
                f"backbone {model_name!r}: only {SUPPORTED} have sm_100a kernels in this build")
        self.model_name = model_name
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.act1 = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        cin = 64
        for li, (cout, depth) in enumerate(zip((64, 128, 256, 512), DEPTHS[model_name]), start=1):
            stride = 1 if li == 1 else 2
            if model_name in BOTTLENECK:
                blocks = [_Bottleneck(cin, cout, stride)] + [_Bottleneck(4 * cout, cout, 1) for _ in range(depth - 1)]
                cin = 4 * cout
            else:
                blocks = [_BasicBlock(cin, cout, stride)] + [_BasicBlock(cout, cout, 1) for _ in range(depth - 1)]
                cin = cout
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        self.num_features = cin
        for m in self.modules():                      # timm's init: kaiming-normal convs, unit BN
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self._features_fn = None                      # installed by BinaryClassifier

    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        if self._features_fn is None:
            raise RuntimeError("ResNetTrunk has no engine attached; use it through BinaryClassifier")
        return self._features_fn(x)

    def forward(self, x):
        return self.forward_features(x).mean((2, 3))
