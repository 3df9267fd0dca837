"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    python -m oracle.make_golden            (from the repo root; needs /root/reference)

Every array below is an output of the reference's own functions
(/root/reference/modular/source/inference_runner.py, imported through the timm shim) or
of the third-party transforms it calls, on seeded inputs from ``oracle.fixtures``.
The goldens pin ``oracle.restatement`` (tests/test_oracle_golden.py) and, through it,
the CUDA path.  Re-running must reproduce the committed files bit for bit on the same
torch/torchaudio/torchvision build (2.11.0 / 2.11.0 / 0.26.0, CPU).
"""
import os
import sys
import tempfile

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import fixtures as FX          # noqa: E402
from oracle import reference_api           # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_slicing(IR):
    """slice_waveform (IR:176-190) on ragged / silent / short clips, both overlap settings."""
    cases = {}
    sr = 32000
    clips = {
        "ragged": FX.synth_clip(5 * 128000 + 777, seed=11),
        "silent_mid": FX.synth_clip(6 * 128000, seed=12, silent_spans=[(128000, 3 * 128000)]),
        "exact_one": FX.synth_clip(128000, seed=13),
        "short_padded": None,   # filled below through the reference's own padding rule
        "all_silent": torch.zeros(3 * 128000),
        "quiet_edge": FX.synth_clip(2 * 128000, seed=14) * 0.0 + 9.9e-4,   # just under the CLI gate
    }
    short = FX.synth_clip(50000, seed=15)
    padded = torch.zeros(128000)
    padded[:50000] = short              # IR:150-154
    clips["short_padded"] = padded
    for name, wf in clips.items():
        for tag, cfg in (("cli", IR.AudioConfig(32000, 4.0, 0.0, 1e-3)),       # IR:258
                         ("dflt", IR.AudioConfig())):                          # IR:128-132
            chunks, stamps = IR.slice_waveform(wf, sr, cfg)
            starts = np.array([int(round(t * sr)) for t in stamps], dtype=np.int64)
            for c, s in zip(chunks, starts):
                assert torch.equal(c, wf[s:s + 128000])
            cases[f"{name}.{tag}.starts"] = starts
            cases[f"{name}.{tag}.stamps"] = np.array(stamps, dtype=np.float64)
            cases[f"{name}.{tag}.n_samples"] = np.array(wf.shape[0], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "slicing.npz"), **cases)
    print("slicing:", len(cases), "arrays")


def golden_frontend(IR, seg_ids):
    """waveform_to_spectrogram (IR:157-174) + its intermediate torchaudio outputs."""
    x = torch.cat([FX.synth_segments(1, first=i) for i in seg_ids])
    cfg = IR.SpectrogramConfig(2048, 512, 128, 20, 12000, 80, "slaney")       # IR:259
    mel = torchaudio.transforms.MelSpectrogram(sample_rate=32000, n_fft=2048, hop_length=512, n_mels=128,
                                               f_min=20, f_max=12000, norm="slaney")   # IR:158-166
    a2d = torchaudio.transforms.AmplitudeToDB(top_db=80)                                # IR:167
    dbs, mus, sds, imgs = [], [], [], []
    for i in range(x.shape[0]):
        spec = a2d(mel(x[i].unsqueeze(0)))                                              # IR:169-170
        dbs.append(spec[0].numpy().copy())
        mus.append(float(spec.mean()))
        sds.append(float(spec.std()))
        img = IR.waveform_to_spectrogram(x[i], 32000, cfg)                              # [1,3,512,512]
        assert torch.equal(img[0, 0], img[0, 1]) and torch.equal(img[0, 0], img[0, 2])
        imgs.append(img[0, 0].numpy().copy())
    imgs = np.stack(imgs)
    np.savez_compressed(
        os.path.join(OUT, "frontend.npz"),
        seg_ids=np.array(seg_ids, dtype=np.int64),
        pcm_checksum=np.array([float(x[i].double().sum()) for i in range(x.shape[0])]),
        logmel_db=np.stack(dbs).astype(np.float32),
        mu=np.array(mus, np.float32), sigma=np.array(sds, np.float32),
        image_full=imgs[:2],                       # two complete 512x512 channels
        image_sub=imgs[:, ::7, ::5].copy(),        # strided sample of every image
        image_sum=imgs.astype(np.float64).sum(axis=(1, 2)),
        fb_nnz=np.array(int((mel.mel_scale.fb != 0).sum())),
        fb_colsum=mel.mel_scale.fb.sum(0).numpy(),
        window_sum=np.array(float(mel.spectrogram.window.sum())),
    )
    print("frontend:", imgs.shape)
    return x


def golden_ensemble(IR, MM, n_heads, seg_ids, tag, backbone="resnet18"):
    """load_merged_model + ModularMultiHeadClassifier.forward + interpret_multihead_logits."""
    x = torch.cat([FX.synth_segments(1, first=i) for i in seg_ids])
    cfg = IR.SpectrogramConfig(2048, 512, 128, 20, 12000, 80, "slaney")
    imgs = torch.cat([IR.waveform_to_spectrogram(x[i], 32000, cfg) for i in range(x.shape[0])])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "merged.pth")
        FX.save_merged_checkpoint(path, n_heads, backbone=backbone)
        model, meta = IR.load_merged_model(path, torch.device("cpu"), backbone_name=backbone)      # IR:77-123
    names = meta["class_names"]
    with torch.no_grad():
        merged = model(imgs)                                                            # IR:62-73
        per_head = torch.stack([m(imgs) for m in model.sub_models], dim=1)              # [B,N,2]
    labels, probs = [], []
    for row in merged:
        lab, s = IR.interpret_multihead_logits(row, 0.5, names[:-1], names[-1])         # IR:194-214
        labels.append(lab)
        probs.append(s)
    probs = np.stack(probs)
    pct = np.mean(list(probs), axis=0) * 100                                            # IR:328-334
    np.savez_compressed(
        os.path.join(OUT, f"ensemble_{tag}.npz"),
        seg_ids=np.array(seg_ids, dtype=np.int64), n_heads=np.array(n_heads),
        merged_logits=merged.numpy(), per_head_logits=per_head.numpy(),
        probs=probs, labels=np.array(labels), class_names=np.array(names),
        percentages=pct.astype(np.float64),
    )
    print(f"ensemble_{tag}: logits\n", merged.numpy(), "\nlabels", labels)


def golden_ingest(IR):
    """preprocess_waveform (IR:144-155), UNMODIFIED, on seeded int16 PCM.  torchaudio.load cannot decode in this image
    (it needs torchcodec), so the loader -- and only the loader -- is replaced by what it returns for 16-bit PCM:
    float32 [channels, frames] = int16 / 32768 and the sample rate."""
    import torchaudio as TA
    cfg = IR.AudioConfig(32000, 4.0, 0.0, 1e-3)                                       # IR:258
    out = {}
    real_load = TA.load
    try:
        for name, sr, ch, frames, seed in FX.INGEST_CASES:
            pcm = FX.synth_pcm16(frames, ch, sr, seed)
            TA.load = lambda path, _p=pcm, _sr=sr: (torch.from_numpy(_p.astype(np.float32) / 32768.0).T.contiguous(), _sr)
            wf, sr_out = IR.preprocess_waveform(f"{name}.wav", cfg)
            assert sr_out == 32000 and wf.dtype == torch.float32
            y = wf.numpy()
            n_real = int(np.ceil(np.float32(32000 * frames / sr))) if sr != 32000 else frames
            out[f"{name}.length"] = np.array(y.shape[0], dtype=np.int64)
            out[f"{name}.n_real"] = np.array(n_real, dtype=np.int64)
            out[f"{name}.pcm_checksum"] = np.array(int(pcm.astype(np.int64).sum()), dtype=np.int64)
            out[f"{name}.sum"] = np.array(float(y.astype(np.float64).sum()))
            if n_real <= 16384:
                out[f"{name}.head"] = y[:n_real].copy()                                # everything that is not padding
            else:
                out[f"{name}.head"] = y[:8192].copy()
                out[f"{name}.tail"] = y[n_real - 8192:n_real].copy()
                out[f"{name}.strided"] = y[::97].copy()
    finally:
        TA.load = real_load
    np.savez_compressed(os.path.join(OUT, "ingest.npz"), **out)
    print("ingest:", {k: int(v) for k, v in out.items() if k.endswith(".length")})


def main():
    if not reference_api.available():
        raise SystemExit("the reference is not present; goldens can only be made in the build container")
    torch.manual_seed(0)
    IR, MM = reference_api.load()
    os.makedirs(OUT, exist_ok=True)
    golden_slicing(IR)
    golden_ingest(IR)
    # segments 0..5 (mixed), plus the first "pure tone" and "pure noise" draws of the stream
    golden_frontend(IR, [0, 1, 2, 3, 4, 5, 6, 13])
    golden_ensemble(IR, MM, 2, [0, 1, 2, 3, 4, 5, FX.CAL_FIRST, FX.CAL_FIRST + 1], "n2")
    golden_ensemble(IR, MM, 5, [0, 1, FX.CAL_FIRST, FX.CAL_FIRST + 1], "n5")
    golden_ensemble(IR, MM, 2, [0, 1, FX.CAL_FIRST], "r34_n2", backbone="resnet34")
    golden_ensemble(IR, MM, 2, [0, 1, FX.CAL_FIRST], "r50_n2", backbone="resnet50")        # SURVEY 8f4, Bottleneck


if __name__ == "__main__":
    main()
