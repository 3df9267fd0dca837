"""Stand-in for the ``timm`` dependency (oracle / test infrastructure only).

The reference does ``timm.create_model(name, pretrained=True, num_classes=0)``
(``modular/source/inference_runner.py:35``, ``modular/source/model_merger.py:24``)
and then only touches ``.num_features`` and ``.forward_features``.  ``timm`` is
not installed in this image and there is no network, so the oracle provides a
module with that surface built on ``torchvision.models.resnet*``, whose
parameter names (``conv1``, ``bn1``, ``layer{1..4}.{i}.{conv1,bn1,conv2,bn2,
downsample.{0,1}}``) are the names timm's ResNet uses, so merged checkpoints are
key-compatible.  Weights are whatever torch's default init gives; callers load or
overwrite them.
"""
import sys
import types

import torch.nn as nn
import torchvision

_NAMES = ("resnet18", "resnet34", "resnet50", "resnet101", "resnet152")


class _FeatureResNet(nn.Module):
    """torchvision ResNet trunk without the classifier; timm-like surface."""

    def __init__(self, name: str):
        super().__init__()
        if name not in _NAMES:
            raise RuntimeError(f"Unknown model ({name})")
        trunk = getattr(torchvision.models, name)(weights=None)
        self.conv1 = trunk.conv1
        self.bn1 = trunk.bn1
        self.act1 = nn.ReLU(inplace=True)
        self.maxpool = trunk.maxpool
        self.layer1 = trunk.layer1
        self.layer2 = trunk.layer2
        self.layer3 = trunk.layer3
        self.layer4 = trunk.layer4
        self.num_features = trunk.fc.in_features

    def forward_features(self, x):
        x = self.maxpool(self.act1(self.bn1(self.conv1(x))))
        x = self.layer1(x)
        x = self.layer2(x)
        x = self.layer3(x)
        return self.layer4(x)

    def forward(self, x):  # timm: forward_head == global pool + Identity fc
        return self.forward_features(x).mean((2, 3))


def _create_model(model_name, pretrained=False, num_classes=0, **_unused):
    return _FeatureResNet(model_name)


def install():
    """Put the shim into ``sys.modules['timm']`` unless a real timm is importable."""
    if "timm" in sys.modules:
        return sys.modules["timm"]
    try:  # pragma: no cover - timm is absent in this image
        import timm  # noqa: F401
        return sys.modules["timm"]
    except Exception:
        pass
    shim = types.ModuleType("timm")
    shim.create_model = _create_model
    shim.list_models = lambda pattern="": list(_NAMES)
    shim.__is_oracle_shim__ = True
    sys.modules["timm"] = shim
    return shim
