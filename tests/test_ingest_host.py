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


PAIR_GRID = [(192000, 4, 800_000, 148), (192000, 2, 100_001, 148), (48000, 8, 2_000_001, 148), (16000, 4, 1_000_003, 148),
             (44100, 4, 300_000, 148), (44100, 4, 9_000_001, 148), (44100, 2, 70_001, 148), (44100, 8, 1_000_003, 132),
             (48000, 2, 100_001, 148), (48000, 4, 6_000_000, 148), (22050, 8, 70_001, 148), (22050, 4, 4_000_001, 148),
             (16000, 8, 5_000_000, 148), (16000, 2, 639, 148), (8000, 2, 4_001, 148), (8000, 4, 2_000_000, 148),
             (24000, 4, 96_001, 148), (12000, 4, 50_000, 148), (25600, 4, 123_457, 148), (37800, 4, 77_777, 148),
             (47250, 4, 500_001, 148), (64000, 4, 300_001, 148), (88200, 2, 150_001, 148), (96000, 4, 500_000, 148),
             (96000, 8, 9_000_000, 148), (44100, 4, 1, 148), (44100, 4, 882, 148), (44100, 4, 883, 1)]


@pytest.mark.parametrize("sr,bytes_per_frame,frames,sms", PAIR_GRID)
def test_pair_kernel_geometry_reads_the_right_frames(checker, sr, bytes_per_frame, frames, sms):
    """The staged (`pair`) resampler of csrc/ingest.cu, replayed on the CPU with the geometry its host side picks
    (ingest_taps.h choose_pair_geometry): raw 16-byte chunks of an item -> per-round float sub-spans -> per-thread
    16-byte windows.  Every tap of every output must land on the input frame torchaudio's dense kernel pairs it with,
    frames outside the stream must read as zero, and every output below the un-padded length must be produced once."""
    r = subprocess.run([checker, "pair", str(sr), str(bytes_per_frame), str(frames), str(sms)], capture_output=True, text=True)
    lines = r.stdout.strip().split("\n")
    assert r.returncode == 0, r.stdout[-500:]
    geo = dict(zip(lines[0].split()[::2], map(int, lines[0].split()[1::2])))
    few_phase = {96000: 4, 64000: 4, 192000: 2, 48000: 8, 16000: 8}             # one or two phases: taps as kernel parameters
    if sr in few_phase:
        assert geo["uniform"] == 1 and geo["G"] == few_phase[sr] and geo["smem"] <= 52 * 1024      # four blocks per SM
        assert geo["threads"] == (128 if bytes_per_frame == 8 and sr in (48000, 96000) else 256)
    else:
        assert geo["uniform"] == 0 and geo["G"] == (1 if sr in (88200, 96000) else 2)
    assert geo["threads"] % 32 == 0 or sr == 25600                 # 5 phases: 320 threads anyway
    assert geo["smem"] <= 100 * 1024                                 # two blocks per SM
    n_real = -(-32000 * frames // sr)                                # ceil; the float32 rounding of torchaudio matters only
    assert lines[1].startswith("ok ") and abs(int(lines[1].split()[1]) - n_real) <= 1   # above 2^24 samples


@pytest.mark.parametrize("sr,bytes_per_frame,frames", [(48000, 4, 6_000_000), (96000, 4, 500_000), (16000, 2, 70_001)])
def test_few_phase_ratios_also_replay_on_the_register_tap_geometry(checker, sr, bytes_per_frame, frames):
    """SAD_INGEST_UNIFORM=0 sends the one- and two-phase ratios through the register-tap form of the kernel."""
    r = subprocess.run([checker, "pair", str(sr), str(bytes_per_frame), str(frames), "148", "0"], capture_output=True, text=True)
    assert r.returncode == 0 and " uniform 0" in r.stdout and "\nok " in r.stdout, r.stdout[-300:]


@pytest.mark.parametrize("sr", [11025, 12345])
def test_ratios_the_pair_kernel_leaves_to_the_one_phase_kernel(checker, sr):
    out = subprocess.run([checker, "pair", str(sr), "4", "100000", "148"], capture_output=True, text=True, check=True).stdout
    assert out.strip() in ("not a pair ratio", "unsupported")
