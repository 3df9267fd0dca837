"""GPU parity of the drop-in Python API (sad_b200.inference_runner == reference modular/source/inference_runner.py):
module forward, load_merged_model, slice/spectrogram helpers and the CLI's JSON, against the CPU oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import restatement as R
from tests import gpu_common as G
from tests.test_host_api import _write_wav
import sad_b200.inference_runner as IR

pytestmark = pytest.mark.gpu
LOGIT_TOL = 2e-2


@pytest.fixture(scope="module")
def merged_ckpt(tmp_path_factory):
    p = tmp_path_factory.mktemp("ck") / "merged.pth"
    FX.save_merged_checkpoint(str(p), 2)
    return str(p)


def test_load_merged_model_and_module_forward(merged_ckpt, capsys):
    model, meta = IR.load_merged_model(merged_ckpt, torch.device("cuda"))
    out = capsys.readouterr().out
    assert "Found 2 sub-model(s): [0, 1]" in out and "dummy output shape: torch.Size([2, 3])" in out   # IR:99,122
    assert meta["class_names"] == FX.class_names(2)
    sd = G.merged_sd(2)
    x = FX.synth_segments(3, first=60)
    img3 = R.waveform_to_image(x).unsqueeze(1).repeat(1, 3, 1, 1).contiguous()
    want = R.ensemble_forward(img3, sd)
    got = model(img3.cuda())
    assert got.shape == (3, 3) and got.is_cuda
    assert (got.cpu() - want).abs().max() <= LOGIT_TOL
    # one sub-model on its own: [B,2] = [Real, Synthetic] (IR:49-51) and timm-style forward_features
    ph = R.per_head_logits(img3, sd)
    one = model.sub_models[1](img3.cuda())
    assert one.shape == (3, 2) and (one.cpu() - ph[:, 1]).abs().max() <= LOGIT_TOL
    feats = model.sub_models[0].base.forward_features(img3.cuda())
    want_f = R.backbone_features(img3, sd, "sub_models.0.base.")
    assert feats.shape == (3, 512, 16, 16)
    rel = (feats.cpu() - want_f).norm() / want_f.norm()
    assert rel < 0.05
    # parameters changed after the engine was built -> weights are re-uploaded
    with torch.no_grad():
        model.sub_models[0].head[10].bias.add_(1.0)
    got2 = model(img3.cuda())
    assert abs(float((got2 - got)[:, 0].mean()) - 1.0) < 1e-3


def test_helpers_match_reference_goldens():
    g = G.golden("frontend.npz")
    x = G.segs(g["seg_ids"][:2])
    cfg = IR.SpectrogramConfig()
    img = IR.waveform_to_spectrogram(x[0], 32000, cfg)                   # CPU tensor in -> CPU tensor out (IR:157-174)
    assert img.shape == (1, 3, 512, 512) and img.device.type == "cpu"
    assert torch.equal(img[0, 0], img[0, 1]) and torch.equal(img[0, 0], img[0, 2])
    np.testing.assert_allclose(img[0, 0].numpy(), g["image_full"][0], rtol=0, atol=5e-4)
    gs = G.golden("slicing.npz")
    from tests.test_oracle_golden import _clips
    for name, wf in _clips().items():
        chunks, stamps = IR.slice_waveform(wf, 32000, IR.AudioConfig(32000, 4.0, 0.0, 1e-3))
        np.testing.assert_array_equal(np.array(stamps, dtype=np.float64), gs[f"{name}.cli.stamps"])
        for c, t in zip(chunks, stamps):
            s = int(round(t * 32000))
            assert torch.equal(c, wf[s:s + 128000])


@pytest.mark.parametrize("smooth", [False, True])
def test_cli_json_matches_oracle_pipeline(merged_ckpt, tmp_path, smooth):
    """python inference_runner.py --merged-model M --audio A [--smooth]: JSON schema and values (IR:218-353)."""
    wf = FX.synth_clip(7 * 128000 + 999, seed=21, silent_spans=[(2 * 128000, 3 * 128000)])
    wav = tmp_path / "clip.wav"
    _write_wav(str(wav), wf.numpy(), 32000, bits=32)
    out = tmp_path / "res.json"
    argv = ["--merged-model", merged_ckpt, "--audio", str(wav), "--output-json", str(out), "--confidence-threshold", "0.3"]
    IR.main(argv + (["--smooth"] if smooth else []))
    res = json.loads(out.read_text())
    assert set(res) == {"filename", "segments", "percentages"} and res["filename"] == str(wav)
    # oracle pipeline on the same samples
    sd = G.merged_sd(2)
    names = FX.class_names(2)
    starts, kept = R.slice_waveform(wf, 32000, 4.0, 0.0, 1e-3)
    ks = [s for s, k in zip(starts, kept) if k]
    assert len(ks) == 6                                                   # 7 windows, one silent
    x = torch.stack([wf[s:s + 128000] for s in ks])
    img3 = R.waveform_to_image(x).unsqueeze(1).repeat(1, 3, 1, 1)
    logits = R.ensemble_forward(img3, sd)
    labels, probs = R.interpret(logits, 0.5)
    if smooth:
        probs, labels = R.smooth_probs(probs, 0.5)
    margin = G.decision_margin(logits.numpy())
    assert [s["start_sec"] for s in res["segments"]] == [s / 32000 for s in ks]
    assert all(s["end_sec"] == s["start_sec"] + 4.0 for s in res["segments"])
    if not smooth:
        for seg, l, m in zip(res["segments"], labels, margin):
            assert seg["label"] == R.label_name(int(l), 2, names[:-1], names[-1]) or m <= LOGIT_TOL
    want_pct = np.mean(probs, axis=0) * 100
    got_pct = np.array([res["percentages"][n] for n in names])
    np.testing.assert_allclose(got_pct, want_pct, rtol=0, atol=100 * LOGIT_TOL / 4)


def test_cli_all_silent_writes_empty_result(merged_ckpt, tmp_path):
    wav = tmp_path / "silent.wav"
    _write_wav(str(wav), np.zeros(3 * 128000, np.float32), 32000, bits=16)
    out = tmp_path / "res.json"
    IR.main(["--merged-model", merged_ckpt, "--audio", str(wav), "--output-json", str(out)])
    assert json.loads(out.read_text()) == {"filename": str(wav), "segments": [], "percentages": {}}   # IR:264-273


def test_batch_folder_driver(merged_ckpt, tmp_path):
    """a folder of WAVs -> one JSON per clip + summary.json, same numbers as the single-file CLI."""
    import sad_b200.batch_runner as BR
    folder, out = tmp_path / "wavs", tmp_path / "out"
    folder.mkdir()
    for i, n in enumerate((3 * 128000 + 5, 128000, 60000)):
        _write_wav(str(folder / f"c{i}.wav"), FX.synth_clip(n, seed=30 + i).numpy(), 32000, bits=32)
    _write_wav(str(folder / "zsilent.wav"), np.zeros(2 * 128000, np.float32), 32000, bits=16)
    (folder / "notes.txt").write_text("ignored")
    BR.main(["--merged-model", merged_ckpt, "--folder", str(folder), "--out-dir", str(out)])
    summary = json.loads((out / "summary.json").read_text())
    assert [os.path.basename(r["filename"]) for r in summary] == ["c0.wav", "c1.wav", "c2.wav", "zsilent.wav"]
    assert [r["n_segments"] for r in summary] == [3, 1, 1, 0] and summary[3]["label"] == ""
    single = tmp_path / "single.json"
    IR.main(["--merged-model", merged_ckpt, "--audio", str(folder / "c0.wav"), "--output-json", str(single)])
    assert json.loads(single.read_text()) == json.loads((out / "c0.json").read_text())
    assert summary[0]["label"] in FX.class_names(2)


def test_resnet34_backbone_through_the_python_api(tmp_path, capsys):
    """backbone_name / --model-name = resnet34 (IR:77, MM:101): same kernels, deeper trunk."""
    p = tmp_path / "m34.pth"
    sd = FX.save_merged_checkpoint(str(p), 2, backbone="resnet34")
    model, meta = IR.load_merged_model(str(p), torch.device("cuda"), backbone_name="resnet34")
    assert "dummy output shape: torch.Size([2, 3])" in capsys.readouterr().out
    x = FX.synth_segments(2, first=70)
    img3 = R.waveform_to_image(x).unsqueeze(1).repeat(1, 3, 1, 1).contiguous()
    want = R.ensemble_forward(img3, sd)
    got = model(img3.cuda())
    lo, _, _ = model.forward_pcm(x.cuda(), 0.5)
    print("resnet34: max |logit diff| images", (got.cpu() - want).abs().max().item(), "pcm", (lo.cpu() - want).abs().max().item())
    assert (got.cpu() - want).abs().max() <= LOGIT_TOL and (lo.cpu() - want).abs().max() <= LOGIT_TOL
    with pytest.raises(NotImplementedError):
        IR.BinaryClassifier("resnext50_32x4d")


def test_resnet50_backbone_through_the_python_api(tmp_path, capsys):
    """--model-name resnet50 (MM:101, IR:77): the Bottleneck family runs on the same kernels; BinaryClassifier's first
    Linear takes the trunk's 2048 features (IR:36-37)."""
    sd = G.merged_sd(2, "resnet50")
    p = tmp_path / "m50.pth"
    torch.save({"state_dict": sd, "metadata": {"class_names": FX.class_names(2)}}, str(p))
    model, meta = IR.load_merged_model(str(p), torch.device("cuda"), backbone_name="resnet50")
    assert "dummy output shape: torch.Size([2, 3])" in capsys.readouterr().out
    assert model.sub_models[0].base.num_features == 2048 and model.sub_models[0].head[2].in_features == 2048
    x = FX.synth_segments(2, first=70)
    img3 = R.waveform_to_image(x).unsqueeze(1).repeat(1, 3, 1, 1).contiguous()
    want = R.ensemble_forward(img3, sd)
    got = model(img3.cuda())
    lo, _, _ = model.forward_pcm(x.cuda(), 0.5)
    f = model.sub_models[1].base.forward_features(img3[:1].cuda())
    assert f.shape == (1, 2048, 16, 16)
    print("resnet50: max |logit diff| images", (got.cpu() - want).abs().max().item(), "pcm", (lo.cpu() - want).abs().max().item())
    tol50 = 6e-2                        # bf16 activations through 53 convs: see tests/test_gpu_ensemble.py (r50_n2)
    assert (got.cpu() - want).abs().max() <= tol50 and (lo.cpu() - want).abs().max() <= tol50
