"""Host logic of the drop-in API (no GPU): checkpoint layout, loader errors, merger CSV rules, WAV ingest, configs.
Reference behaviour cited as IR:/MM: lines of /root/reference/modular/source/{inference_runner,model_merger}.py."""
import os
import struct

import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import reference_api
import sad_b200.inference_runner as IR
import sad_b200.model_merger as MM


def test_dataclass_defaults_match_reference():
    a, s = IR.AudioConfig(), IR.SpectrogramConfig()
    assert (a.sample_rate, a.window_size, a.overlap, a.silence_threshold) == (32000, 4.0, 0.85, 1e-4)     # IR:127-132
    assert (s.n_fft, s.hop_length, s.n_mels, s.f_min, s.f_max, s.top_db, s.norm) == (2048, 512, 128, 20, 12000, 80, "slaney")
    assert IR.window_and_hop(32000, a) == (128000, 19200)              # IR:181 float quirk
    assert IR.window_and_hop(32000, IR.AudioConfig(32000, 4.0, 0.0, 1e-3)) == (128000, 128000)


def test_state_dict_keys_match_reference_layout():
    m = IR.ModularMultiHeadClassifier([IR.BinaryClassifier(), IR.BinaryClassifier()])
    want = FX.merged_state_dict(2, calibrate=False)
    got = m.state_dict()
    assert list(got.keys()) == list(want.keys())
    assert all(got[k].shape == want[k].shape and got[k].dtype == want[k].dtype for k in want)


@pytest.mark.skipif(not reference_api.available(), reason="reference sources not on this machine")
def test_state_dict_keys_match_live_reference():
    RIR, _ = reference_api.load()
    ref = RIR.ModularMultiHeadClassifier([RIR.BinaryClassifier(), RIR.BinaryClassifier()]).state_dict()
    mine = IR.ModularMultiHeadClassifier([IR.BinaryClassifier(), IR.BinaryClassifier()]).state_dict()
    assert list(ref.keys()) == list(mine.keys())
    assert all(ref[k].shape == mine[k].shape for k in ref)


def test_load_merged_model_errors_and_sparse_indices(tmp_path, capsys):
    sd = FX.merged_state_dict(2, calibrate=False)
    p = tmp_path / "m.pth"
    torch.save({"state_dict": sd}, p)
    with pytest.raises(ValueError, match="does not contain metadata for class names"):     # IR:85-86
        IR.load_merged_model(str(p), torch.device("cpu"))
    torch.save({"state_dict": sd, "metadata": {"foo": 1}}, p)
    with pytest.raises(ValueError):
        IR.load_merged_model(str(p), torch.device("cpu"))
    torch.save({"metadata": {"class_names": ["a", "Real"]}}, p)
    with pytest.raises(KeyError):                                                           # IR:83
        IR.load_merged_model(str(p), torch.device("cpu"))
    # non-contiguous indices are sorted and packed densely (IR:89-98); missing keys keep fresh values (IR:105-110)
    sparse = {k.replace("sub_models.1.", "sub_models.7."): v for k, v in sd.items()}
    del sparse["sub_models.7.head.10.bias"]
    torch.save({"state_dict": sparse, "metadata": {"class_names": ["A", "B", "Real"]}}, p)
    model, meta = IR.load_merged_model(str(p), torch.device("cpu"))
    assert "Found 2 sub-model(s): [0, 7]" in capsys.readouterr().out
    assert len(model.sub_models) == 2 and meta["class_names"] == ["A", "B", "Real"]
    assert torch.equal(model.sub_models[1].base.conv1.weight, sd["sub_models.1.base.conv1.weight"])
    assert not model.training


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = IR.ModularMultiHeadClassifier([IR.BinaryClassifier()]).eval()
    with pytest.raises(Exception, match="no CPU fallback"):
        m(torch.zeros(1, 3, 512, 512))
    with pytest.raises(Exception, match="no CPU fallback|CUDA"):
        IR.waveform_to_spectrogram(torch.zeros(128000), 32000, IR.SpectrogramConfig())
    with pytest.raises(NotImplementedError):
        IR.waveform_to_spectrogram(torch.zeros(128000), 16000, IR.SpectrogramConfig())
    with pytest.raises(NotImplementedError):
        IR.BinaryClassifier("resnext50_32x4d")
    b50 = IR.BinaryClassifier("resnet50")                    # Bottleneck nets build their parameter holders offline
    assert b50.base.num_features == 2048 and b50.head[2].in_features == 2048
    assert b50.base.layer3[5].conv3.weight.shape == (1024, 256, 1, 1)
    assert list(b50.state_dict().keys())[:2] == ["base.conv1.weight", "base.bn1.weight"]


def test_interpret_rule():
    names = ["A", "B"]
    cases = [([-1.0, -2.0, 3.0], "Real"), ([-1.0, -2.0, -3.0], "A"), ([2.0, 3.0, 5.0], "B"), ([0.0, -1.0, 0.0], "A")]
    for z, want in cases:
        lab, s = IR.interpret_multihead_logits(torch.tensor(z), 0.5, names, "Real")
        assert lab == want and s.dtype == np.float32 and s.shape == (3,)
    assert IR.interpret_multihead_logits(torch.tensor([0.0, 9.0, -9.0]), 0.5, ["A"], "Real")[0] == "Synthetic_2"


def test_merger_cli_roundtrip(tmp_path, capsys):
    """MM:93-160: CSV -> merged checkpoint with class_names = synthetic... + [most common real]."""
    for i in range(3):
        g = torch.Generator().manual_seed(50 + i)
        sub = {k: v for k, v in FX.random_head_state(g).items()}
        torch.save({"state_dict": sub, "epoch": 3}, tmp_path / f"m{i}.pth")
    csv = tmp_path / "merge.csv"
    csv.write_text("model_filename,synthetic_class,real_class\nm0.pth,SynA,Real\nm1.pth,SynB,Human\nm2.pth,SynC,Real\n")
    out = tmp_path / "merged.pth"
    MM.main(["--submodels-folder", str(tmp_path), "--csv-file", str(csv), "--output-path", str(out)])
    txt = capsys.readouterr().out
    assert "Warning: Not all real_class values match" in txt and "Saved merged model" in txt
    ck = torch.load(out)
    assert ck["metadata"]["class_names"] == ["SynA", "SynB", "SynC", "Real"]
    assert len(ck["state_dict"]) == 3 * 136
    g = torch.Generator().manual_seed(51)
    assert torch.equal(ck["state_dict"]["sub_models.1.base.layer3.0.conv2.weight"],
                       FX.random_head_state(g)["base.layer3.0.conv2.weight"])
    model, meta = IR.load_merged_model(str(out), torch.device("cpu"))
    assert len(model.sub_models) == 3
    # trainer-style checkpoints without the "base." prefix only match head.* (MM:55 strict=False quirk)
    bare = {k.replace("base.", ""): v for k, v in FX.random_head_state(torch.Generator().manual_seed(1)).items()}
    torch.save({"state_dict": bare}, tmp_path / "bare.pth")
    sm = MM.load_sub_model(str(tmp_path / "bare.pth"), torch.device("cpu"))
    assert torch.equal(sm.head[2].weight, bare["head.2.weight"])
    assert not torch.equal(sm.base.conv1.weight, bare["conv1.weight"])


def _write_wav(path, x, sr, bits=16, channels=1):
    x = np.asarray(x, np.float32)
    if channels > 1:
        x = np.stack([x] * channels, axis=1).reshape(-1)
    if bits == 16:
        raw = (np.clip(x, -1, 1) * 32767).astype("<i2").tobytes(); tag = 1
    else:
        raw = x.astype("<f4").tobytes(); tag = 3
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, tag, channels, sr, sr * channels * bits // 8, channels * bits // 8, bits) + b"data" + struct.pack("<I", len(raw))
    open(path, "wb").write(hdr + raw)


def test_wav_container_parsing(tmp_path):
    """The host half of IR:144-145: the container is parsed here, the samples stay as they sit in the file."""
    rng = np.random.default_rng(0)
    x = np.clip(0.2 * rng.standard_normal(5000), -0.99, 0.99).astype(np.float32)
    _write_wav(tmp_path / "f32.wav", x, 32000, bits=32)
    pcm, sr = IR._read_wav_pcm(str(tmp_path / "f32.wav"))
    assert sr == 32000 and pcm.dtype == np.float32 and pcm.shape == (5000, 1) and np.array_equal(pcm[:, 0], x)
    _write_wav(tmp_path / "stereo16.wav", x, 44100, bits=16, channels=2)
    pcm, sr = IR._read_wav_pcm(str(tmp_path / "stereo16.wav"))
    assert sr == 44100 and pcm.dtype == np.int16 and pcm.shape == (5000, 2)
    np.testing.assert_array_equal(pcm[:, 0], (x * 32767).astype(np.int16))
    wf, sr = IR._read_wav(str(tmp_path / "stereo16.wav"))            # what torchaudio.load would return
    assert wf.shape == (2, 5000) and wf.dtype == torch.float32
    assert torch.equal(wf[1], torch.from_numpy(pcm[:, 1].astype(np.float32) / 32768.0))
    with pytest.raises(ValueError):
        (tmp_path / "bad.wav").write_bytes(b"not a wav file at all")
        IR._read_wav_pcm(str(tmp_path / "bad.wav"))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_preprocess_waveform_fails_loudly_without_gpu(tmp_path):
    """Mix / resample / pad run in sad_ingest on the device; there is no CPU fallback."""
    from sad_b200._lib import SadError
    _write_wav(tmp_path / "a.wav", np.zeros(100, np.float32), 16000)
    with pytest.raises(SadError, match="no CPU fallback"):
        IR.preprocess_waveform(str(tmp_path / "a.wav"), IR.AudioConfig())
    with pytest.raises(NotImplementedError):
        IR.preprocess_waveform(str(tmp_path / "a.wav"), IR.AudioConfig(sample_rate=16000))


def test_default_activation_dtype_per_backbone(monkeypatch):
    """bf16 -- the dtype BASELINE.json names -- for the BasicBlock nets, fp16 for the Bottleneck nets."""
    from sad_b200 import engine as ENG
    assert [ENG.default_dtype(b) for b in ("resnet18", "resnet34", "resnet50", "resnet101", "resnet152")] == \
        ["bf16", "bf16", "fp16", "fp16", "fp16"]
    from sad_b200 import _lib
    with pytest.raises(ValueError):
        _lib.load("fp8")
    assert _lib.load("bf16") is not _lib.load("fp16")
