"""Randomised check of sad_ingest against the oracle's restatement of the reference's preprocess_waveform
(torchaudio resampling on the CPU): random rates, lengths, channel counts, sample formats and stream alignments.

    python tools/soak_ingest.py [cases] [seed]       (needs a GPU; test infrastructure, not product code)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures as FX      # noqa: E402
from oracle import restatement as R    # noqa: E402
from sad_b200.engine import Engine     # noqa: E402

RATES = [8000, 11025, 12000, 16000, 22050, 24000, 25600, 37800, 44100, 47250, 48000, 64000, 88200, 96000, 192000, 32000]


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
    eng = Engine(2, max_batch=1)
    worst = 0.0
    for case in range(n_cases):
        sr = int(rng.choice(RATES))
        ch = int(rng.choice([1, 1, 2, 2, 2, 3]))
        kind = rng.integers(0, 4)
        frames = int({0: rng.integers(1, 2000), 1: rng.integers(2000, 200_000), 2: rng.integers(200_000, 1_500_000),
                      3: rng.integers(1_500_000, 4_000_000)}[int(kind)])
        fmt = "s16" if rng.integers(0, 2) else "f32"
        odd = int(rng.integers(0, 4) == 0)
        pcm = FX.synth_pcm16(frames, ch, sr, seed=int(rng.integers(0, 1000)))
        want = R.ingest(pcm, sr).numpy()
        x = torch.from_numpy(pcm if fmt == "s16" else pcm.astype(np.float32) / 32768.0).cuda()
        if odd:
            buf = torch.empty(frames * ch + 1, dtype=x.dtype, device="cuda")
            buf[1:].copy_(x.reshape(-1))
            x = buf[1:].reshape(frames, ch)
        got = eng.ingest(x, sr).cpu().numpy()
        assert got.shape == want.shape, (sr, ch, frames, got.shape, want.shape)
        err = float(np.abs(got - want).max())
        worst = max(worst, err)
        tol = 0.0 if sr == 32000 else 2e-6
        status = "ok" if err <= tol else "FAIL"
        print(f"{status} sr {sr} ch {ch} frames {frames} {fmt} odd {odd}: out {got.shape[0]} max err {err:.2e}")
        if err > tol:
            sys.exit(1)
    print(f"{n_cases} cases, worst |err| {worst:.2e}")


if __name__ == "__main__":
    main()
