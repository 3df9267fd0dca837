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


def golden_frontend_256(IR, n_seg=256, first=20000):
    """BASELINE.json configs[1] asks for a >= 256-segment parity subset of the front end.  Full log-mel tensors would be
    33 MB, so per segment the golden keeps (mean, unbiased std), the max, the float64 sum and a strided sample of the
    log-mel dB (every 7th mel band x every 5th frame + the last frame), all from the reference's own transforms
    (IR:158-170) -- plus a strided sample of the 512x512 image from IR.waveform_to_spectrogram (IR:157-174)."""
    cfg = IR.SpectrogramConfig(2048, 512, 128, 20, 12000, 80, "slaney")
    mel = torchaudio.transforms.MelSpectrogram(sample_rate=32000, n_fft=2048, hop_length=512, n_mels=128,
                                               f_min=20, f_max=12000, norm="slaney")
    a2d = torchaudio.transforms.AmplitudeToDB(top_db=80)
    frames = np.r_[np.arange(0, 251, 5), 250]
    db_s, mu, sd, mx, sm, img_s = [], [], [], [], [], []
    for i in range(n_seg):
        x = FX.synth_segments(1, first=first + i)[0]
        spec = a2d(mel(x.unsqueeze(0)))[0]
        db_s.append(spec[::7][:, frames].numpy().copy())
        mu.append(float(spec.mean())); sd.append(float(spec.std())); mx.append(float(spec.max()))
        sm.append(float(spec.double().sum()))
        img = IR.waveform_to_spectrogram(x, 32000, cfg)[0, 0]
        img_s.append(img[::37, ::41].numpy().copy())
    np.savez_compressed(os.path.join(OUT, "frontend_256.npz"), first=np.array(first, dtype=np.int64),
                        mel_stride=np.array(7), frames=frames.astype(np.int64),
                        logmel_sample=np.stack(db_s).astype(np.float32), mu=np.array(mu, np.float32),
                        sigma=np.array(sd, np.float32), db_max=np.array(mx, np.float32), db_sum=np.array(sm, np.float64),
                        image_sample=np.stack(img_s).astype(np.float32))
    print("frontend_256:", np.stack(db_s).shape, np.stack(img_s).shape)


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


def golden_decisions(IR, n_heads, n_seg, tag, batch=32):
    """Decision parity at scale (north star: identical Real/Synthetic decision on >= 99.9% of segments): the LIVE
    reference -- load_merged_model (IR:77-123), waveform_to_spectrogram per segment (IR:157-174), the merged model in
    mini-batches (IR:284-288), interpret_multihead_logits per row (IR:194-214) -- on `n_seg` HELD-OUT segments of the
    class-structured corpus with the v2 fixture (oracle/fixtures.py: read-outs fitted on other segments)."""
    import time
    cfg = IR.SpectrogramConfig(2048, 512, 128, 20, 12000, 80, "slaney")
    sd = FX.decision_state_dict(n_heads)
    names = FX.class_names(n_heads)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "merged.pth")
        torch.save({"state_dict": sd, "metadata": {"class_names": names}}, path)
        model, meta = IR.load_merged_model(path, torch.device("cpu"))
    names = meta["class_names"]
    logits, labels, classes = [], [], []
    t0 = time.time()
    for b0 in range(0, n_seg, batch):
        x, cls = FX.family_segments(min(batch, n_seg - b0), FX.DEC_HELD_FIRST + b0, n_classes=n_heads + 1)
        imgs = torch.cat([IR.waveform_to_spectrogram(x[i], 32000, cfg) for i in range(x.shape[0])])
        with torch.no_grad():
            merged = model(imgs)
        for row in merged:
            lab, _ = IR.interpret_multihead_logits(row, 0.5, names[:-1], names[-1])
            labels.append(names.index(lab))
        logits.append(merged.numpy().copy())
        classes.append(cls)
        if (b0 // batch) % 8 == 0:
            print(f"decisions_{tag}: {b0}/{n_seg}  {time.time() - t0:.0f} s", flush=True)
    logits = np.concatenate(logits)
    labels = np.array(labels, dtype=np.int16)            # index into class_names: n_heads == Real
    classes = np.concatenate(classes).astype(np.int16)
    np.savez_compressed(os.path.join(OUT, f"decisions_{tag}.npz"), first=np.array(FX.DEC_HELD_FIRST, dtype=np.int64),
                        n_heads=np.array(n_heads), merged_logits=logits.astype(np.float32), labels=labels,
                        classes=classes, class_names=np.array(names))
    m = np.abs(logits).min(axis=1)
    want = np.where(classes == 0, n_heads, classes - 1)
    print(f"decisions_{tag}: {n_seg} segments; label == source family on {(labels == want).mean():.4f}; "
          f"min |logit| p0.1 {np.percentile(m, 0.1):.4f} p1 {np.percentile(m, 1):.4f} median {np.median(m):.4f}")


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
    if len(sys.argv) > 1 and sys.argv[1] == "deep":                # the two deepest Bottleneck backbones
        golden_ensemble(IR, MM, 2, [0, FX.CAL_FIRST], "r101_n2", backbone="resnet101")
        golden_ensemble(IR, MM, 2, [0, FX.CAL_FIRST], "r152_n2", backbone="resnet152")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "frontend256":
        golden_frontend_256(IR)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "decisions":           # ~40 min of CPU: made separately
        for n_heads, n_seg in ((2, 4096), (5, 2048), (6, 2048)):
            if len(sys.argv) > 2 and str(n_heads) not in sys.argv[2:]:
                continue
            golden_decisions(IR, n_heads, n_seg, f"n{n_heads}")
        return
    golden_slicing(IR)
    golden_ingest(IR)
    # segments 0..5 (mixed), plus the first "pure tone" and "pure noise" draws of the stream
    golden_frontend(IR, [0, 1, 2, 3, 4, 5, 6, 13])
    golden_frontend_256(IR)
    golden_ensemble(IR, MM, 2, [0, 1, 2, 3, 4, 5, FX.CAL_FIRST, FX.CAL_FIRST + 1], "n2")
    golden_ensemble(IR, MM, 5, [0, 1, FX.CAL_FIRST, FX.CAL_FIRST + 1], "n5")
    golden_ensemble(IR, MM, 2, [0, 1, FX.CAL_FIRST], "r34_n2", backbone="resnet34")
    golden_ensemble(IR, MM, 2, [0, 1, FX.CAL_FIRST], "r50_n2", backbone="resnet50")        # SURVEY 8f4, Bottleneck


if __name__ == "__main__":
    main()
