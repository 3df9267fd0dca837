"""GPU parity of the ingest stage (sad_ingest): int16 / float32 interleaved PCM -> mono 32 kHz padded.

Tolerance: the mono mix is exact; the resampler sums the same float32 taps as torchaudio's conv1d in a different order
(and skips the taps below 1e-32 where the Hann argument is clamped): |err| <= 2e-6 for |x| <= 1 (about 20 taps, fp32)."""
import os
import struct

import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import restatement as R
from tests import gpu_common as G
from tests.test_oracle_golden import check_ingest_against_golden

pytestmark = pytest.mark.gpu
ATOL = 2e-6


def test_ingest_matches_reference_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "ingest.npz"))
    e = G.engine(2)
    for name, sr, ch, frames, seed in FX.INGEST_CASES:
        pcm = FX.synth_pcm16(frames, ch, sr, seed)
        y = e.ingest(torch.from_numpy(pcm).cuda(), sr).cpu().numpy()
        check_ingest_against_golden(g, name, frames, y, atol=0.0 if sr == 32000 else ATOL)
        # float32 input of the same samples gives the same result as int16 (the scaling by 1/32768 is exact)
        yf = e.ingest(torch.from_numpy(pcm.astype(np.float32) / 32768.0).cuda(), sr).cpu().numpy()
        np.testing.assert_array_equal(y, yf)


@pytest.mark.parametrize("sr,ch,frames", [(44100, 2, 1_000_003), (48000, 6, 300_001), (11025, 1, 70_001),
                                          (192000, 2, 800_000), (32000, 5, 200_000), (12345, 1, 30_000),
                                          (44100, 1, 1), (8000, 2, 0)])
def test_ingest_matches_oracle_on_long_and_odd_inputs(sr, ch, frames):
    pcm = FX.synth_pcm16(frames, ch, sr, seed=frames % 1000)
    want = R.ingest(pcm, sr).numpy()
    e = G.engine(2)
    got = e.ingest(torch.from_numpy(pcm).cuda().reshape(frames, ch), sr).cpu().numpy()
    assert got.shape == want.shape
    # 3+ channel means: ATen divides the fp32 sum, the kernel does the same; allow 1 ulp of the mix through the filter
    np.testing.assert_allclose(got, want, rtol=0, atol=ATOL)


def test_ingest_rejects_bad_arguments():
    from sad_b200 import _lib
    e = G.engine(2)
    x = torch.zeros(100, 2, dtype=torch.int16, device="cuda")
    with pytest.raises(_lib.SadError, match="not supported"):
        e.ingest(x, 31999)                                   # coprime with 32000: 31999 phases x 32000-frame strides
    with pytest.raises(ValueError):
        e.ingest(x.to(torch.int32), 16000)
    with pytest.raises(_lib.SadError):
        e.ingest(x.cpu(), 16000)


def _write_wav16(path, pcm, sr):
    raw = pcm.astype("<i2").tobytes()
    ch = pcm.shape[1]
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, ch, sr, sr * ch * 2, ch * 2, 16) + b"data" + struct.pack("<I", len(raw))
    open(path, "wb").write(hdr + raw)


def test_preprocess_waveform_and_cli_from_cd_audio(tmp_path):
    """IR:144-155 + the rest of the CLI on a 44.1 kHz stereo file: preprocess_waveform == the oracle's ingest, and the
    per-window labels equal the oracle's decisions on the oracle-ingested waveform."""
    import sad_b200.inference_runner as IR
    sr, frames = 44100, int(44100 * 9.3)
    pcm = FX.synth_pcm16(frames, 2, sr, seed=77)
    _write_wav16(tmp_path / "cd.wav", pcm, sr)
    wf, sr_out = IR.preprocess_waveform(str(tmp_path / "cd.wav"), IR.AudioConfig(32000, 4.0, 0.0, 1e-3))
    want = R.ingest(pcm, sr)
    assert sr_out == 32000 and wf.is_cuda and wf.shape == want.shape
    np.testing.assert_allclose(wf.cpu().numpy(), want.numpy(), rtol=0, atol=ATOL)
    # short mono 16 kHz float file: padded to one window
    x = (0.2 * np.random.default_rng(3).standard_normal(20000)).astype(np.float32)
    raw = x.astype("<f4").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 3, 1, 16000, 16000 * 4, 4, 32) + b"data" + struct.pack("<I", len(raw))
    (tmp_path / "short.wav").write_bytes(hdr + raw)
    wf2, _ = IR.preprocess_waveform(str(tmp_path / "short.wav"), IR.AudioConfig())
    want2 = R.ingest(x[:, None], 16000)
    assert wf2.shape[0] == 128000
    np.testing.assert_allclose(wf2.cpu().numpy(), want2.numpy(), rtol=0, atol=ATOL)

    # whole CLI on the 44.1 kHz file
    ck = tmp_path / "merged.pth"
    FX.save_merged_checkpoint(str(ck), 2)
    out = tmp_path / "res.json"
    IR.main(["--merged-model", str(ck), "--audio", str(tmp_path / "cd.wav"), "--output-json", str(out)])
    import json
    res = json.loads(out.read_text())
    starts, kept = R.slice_waveform(want, 32000, 4.0, 0.0, 1e-3)
    segs = torch.stack([want[s:s + 128000] for s, k in zip(starts, kept) if k])
    img3 = R.waveform_to_image(segs).unsqueeze(1).repeat(1, 3, 1, 1)                      # IR:173
    logits = R.ensemble_forward(img3, G.merged_sd(2))
    labels, probs = R.interpret(logits, 0.5)
    assert len(res["segments"]) == segs.shape[0] == 2
    names = FX.class_names(2)
    margin = G.decision_margin(logits.numpy())
    for i, seg in enumerate(res["segments"]):
        if margin[i] > 0.02:                                  # the bf16 trunk may flip a decision only inside its tolerance
            assert seg["label"] == R.label_name(int(labels[i]), 2, names[:-1], names[-1])


_PAIR_PROBE = r"""
import hashlib, sys
import numpy as np, torch
sys.path.insert(0, %r)
from oracle import fixtures as FX
from tests import gpu_common as G
e = G.engine(2)
for sr, ch, frames, fmt, odd in %r:
    pcm = FX.synth_pcm16(frames, ch, sr, seed=5)
    x = torch.from_numpy(pcm if fmt == "s16" else pcm.astype(np.float32) / 32768.0).cuda()
    if odd:                                                  # a stream that starts one element into an allocation
        buf = torch.empty(frames * ch + 1, dtype=x.dtype, device="cuda")
        buf[1:].copy_(x.reshape(-1))
        x = buf[1:].reshape(frames, ch)
    y = e.ingest(x, sr)
    print(sr, ch, frames, fmt, odd, y.numel(), hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest())
    if odd == 0 and frames < 200_000:
        # the C ABI states no alignment for `out` either: an output that starts 4 bytes into an allocation
        from sad_b200 import _lib
        from sad_b200.engine import _ptr, _stream
        buf = torch.full((y.numel() + 1,), float("nan"), device="cuda")
        _lib.check(e.ctx, e.lib.sad_ingest(e.ctx, _ptr(x), 0 if fmt == "s16" else 1, frames, ch, sr, _ptr(buf[1:]),
                                           _stream(e.device)), "sad_ingest")
        assert torch.equal(buf[1:], y), (sr, ch, frames, "unaligned out")
"""


def test_pair_kernel_is_bit_identical_to_the_one_phase_kernel():
    """The staged resampler (ingest.cu, `pair`: two outputs per thread with the taps in registers; one output per thread
    at 88.2 kHz; the few-phase form with the taps as kernel parameters at 96 / 48 / 16 kHz) adds only products with zero
    weights to the sums of the one-phase-per-thread kernel: the two must agree to the bit, for every kind of ratio the
    staged kernel takes, both sample formats, mono and stereo, unaligned streams (which keep the older kernel) and
    lengths that end inside an item.  SAD_INGEST_PAIR=0 selects the older kernel everywhere."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cases = [(44100, 2, 400_003, "s16", 0), (44100, 2, 100_001, "s16", 1), (44100, 1, 90_000, "f32", 1),
             (48000, 2, 300_007, "s16", 0), (48000, 3, 50_001, "f32", 0), (22050, 2, 120_001, "s16", 0),
             (11025, 1, 70_001, "s16", 0), (16000, 1, 66_001, "f32", 0), (8000, 2, 40_003, "s16", 0),
             (24000, 2, 99_999, "s16", 0), (37800, 2, 77_777, "s16", 0), (44100, 2, 3, "s16", 0),
             (96000, 2, 200_001, "s16", 0), (88200, 1, 150_001, "f32", 0), (44100, 2, 9_000_001, "s16", 0),
             (48000, 1, 5_000_003, "s16", 0), (16000, 2, 3_000_001, "f32", 0),   # the long ones use full-size items
             (44100, 2, 1, "s16", 0), (44100, 2, 882, "s16", 0), (44100, 1, 883, "s16", 0), (44100, 2, 1763, "f32", 0),
             (48000, 1, 17, "s16", 0), (48000, 2, 959, "f32", 0), (16000, 2, 639, "s16", 0), (8000, 1, 161, "s16", 0),
             (22050, 2, 176_401, "f32", 0), (24000, 1, 96_001, "s16", 0)]
    outs = []
    for flag in ("1", "0"):
        env = dict(os.environ, SAD_INGEST_PAIR=flag, PYTHONPATH=root)
        r = subprocess.run([sys.executable, "-c", _PAIR_PROBE % (root, cases)], env=env, capture_output=True, text=True,
                           timeout=600, cwd=root)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines())
    assert len(outs[0]) == len(cases)
    for a, b in zip(*outs):
        assert a == b, (a, b)
