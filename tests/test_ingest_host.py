"""Host arithmetic of the ingest stage (csrc/ingest_taps.h, built with g++: no GPU needed) against the oracle's
restatement of torchaudio's resampling kernel, which tests/test_oracle_golden.py pins to the reference bit for bit."""
import importlib.util
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import restatement as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker():
    spec = importlib.util.spec_from_file_location("sad_build", os.path.join(ROOT, "synthetic-audio-detection_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build_ingest_host_check()


@pytest.mark.parametrize("sr", [44100, 48000, 16000, 22050, 8000, 11025, 96000, 192000, 24000, 12345])
def test_tap_bands_are_torchaudios_kernel(checker, sr):
    out = subprocess.run([checker, "taps", str(sr)], capture_output=True, text=True, check=True).stdout.split("\n")
    orig, new, width, full, max_taps = map(int, out[0].split())
    k, w_ref, o_ref, n_ref = R.resample_kernel(sr, 32000)
    assert (orig, new, width, full) == (o_ref, n_ref, w_ref, k.shape[1]) and k.shape[0] == new
    covered = np.zeros_like(k, dtype=bool)
    for p in range(new):
        f = out[1 + p].split()
        first = int(f[0])
        taps = np.array([int(x, 16) for x in f[1:]], dtype=np.uint32).view(np.float32)
        assert taps.shape[0] == max_taps and 0 <= first and first + max_taps <= full
        ref = k[p, first:first + max_taps]
        # float64 sin/cos of libm vs ATen's vectorised ones can differ in the last double bit; after the cast to float32
        # the taps are identical except for an occasional 1-ulp flip
        np.testing.assert_allclose(taps, ref, rtol=2e-7, atol=1e-12, err_msg=f"phase {p}")
        assert (taps.view(np.uint32) != ref.view(np.uint32)).mean() < 0.02
        covered[p, first:first + max_taps] = True
    # everything the band leaves out is where the Hann argument is clamped: below 1e-30, invisible in fp32 sums
    assert np.abs(k[~covered]).max(initial=0.0) < 1e-30
    assert max_taps <= 2 * int(np.ceil(6 * max(orig / new, 1.0) / 0.99)) + 2


def test_output_length_is_torchaudios_float32_ceil(checker):
    frames = [0, 1, 2, 13230, 44099, 44100, 441000, 10_584_011, 10_584_013, 2_000_000_001]
    for sr in (44100, 48000, 16000, 22050, 8000, 96000, 32000):
        lines = subprocess.run([checker, "length", str(sr)] + [str(n) for n in frames], capture_output=True, text=True,
                               check=True).stdout.split("\n")
        g = int(np.gcd(sr, 32000))
        orig, new = sr // g, 32000 // g            # Python ints, as in torchaudio (a numpy scalar would make as_tensor float64)
        for n, line in zip(frames, lines):
            total, n_real = map(int, line.split())
            want = n if sr == 32000 else int(torch.ceil(torch.as_tensor(new * n / orig)).long())
            assert n_real == want == (n if sr == 32000 else R.resample_length(n, orig, new)), (sr, n)
            assert total == max(want, 128000)
