"""The C-ABI library loads without a GPU and exports every symbol include/sad_b200.h declares (no compute here)."""
import ctypes as C
import os
import re

import pytest

from oracle import fixtures as FX

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sad_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sad_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_header_symbols_are_exported_and_bound(dtype):
    """Both builds of the library (bf16 activations = libsad_b200.so, fp16 = libsad_b200_f16.so, csrc/act.cuh)."""
    from sad_b200 import _lib
    lib = _lib.load(dtype)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sad_b200.h but not exported by {_lib.LIB_PATHS[dtype]}"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.sad_act_dtype().decode() == dtype and dtype in lib.sad_version().decode()


def test_weight_manifest_matches_checkpoint_layout():
    """sad_weight_name/numel describe exactly the fp32 tensors of one BinaryClassifier state_dict (IR:28-51)."""
    from sad_b200 import _lib
    lib = _lib.load()
    sd = FX.merged_state_dict(1, calibrate=False)
    want = [(k[len("sub_models.0."):], v.numel()) for k, v in sd.items() if not k.endswith("num_batches_tracked")]
    got = [(lib.sad_weight_name(i).decode(), lib.sad_weight_numel(i)) for i in range(lib.sad_weight_count())]
    assert got == want
    assert lib.sad_weight_name(-1) is None and lib.sad_weight_numel(10 ** 6) == -1
    sd34 = FX.merged_state_dict(1, calibrate=False, backbone="resnet34")
    want34 = [(k[len("sub_models.0."):], v.numel()) for k, v in sd34.items() if not k.endswith("num_batches_tracked")]
    got34 = [(lib.sad_backbone_weight_name(b"resnet34", i).decode(), lib.sad_backbone_weight_numel(b"resnet34", i))
             for i in range(lib.sad_backbone_weight_count(b"resnet34"))]
    assert got34 == want34 and len(got34) == 36 * 5 + 14
    for name, n_convs in (("resnet50", 53), ("resnet101", 104), ("resnet152", 155)):        # Bottleneck nets
        sdb = FX.merged_state_dict(1, calibrate=False, backbone=name)
        wantb = [(k[len("sub_models.0."):], v.numel()) for k, v in sdb.items() if not k.endswith("num_batches_tracked")]
        gotb = [(lib.sad_backbone_weight_name(name.encode(), i).decode(), lib.sad_backbone_weight_numel(name.encode(), i))
                for i in range(lib.sad_backbone_weight_count(name.encode()))]
        assert gotb == wantb and len(gotb) == n_convs * 5 + 14, name
    assert lib.sad_backbone_weight_count(b"resnext50_32x4d") == -1


def test_slice_count_is_len_of_python_range():
    from sad_b200 import _lib
    lib = _lib.load()
    for n, w, h in [(128000, 128000, 128000), (127999, 128000, 128000), (640777, 128000, 128000),
                    (640777, 128000, 19200), (768000, 128000, 19200), (0, 128000, 19200), (128001, 128000, 1)]:
        assert lib.sad_slice_count(n, w, h) == len(range(0, n - w + 1, h)), (n, w, h)
    assert lib.sad_slice_count(10, 5, 0) == -1


def test_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sad_b200 import _lib
    from sad_b200.engine import Engine
    lib = _lib.load()
    ctx = C.c_void_p(0)
    assert lib.sad_create(C.byref(ctx), 0, 2, 8) == _lib.SAD_ENODEVICE
    assert lib.sad_create(C.byref(ctx), 0, 0, 8) == _lib.SAD_EINVAL
    with pytest.raises(_lib.SadError):
        Engine(2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "synthetic-audio-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "/root/reference" not in src or f.endswith(".py") and "reference lines" in src.lower() or \
                    "= /root/reference" in src, f"{f} reads the reference at run time"
