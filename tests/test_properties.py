"""Property tests (hypothesis) of the integer / rule logic on the path: no GPU, no goldens -- the oracle restatement, the
C-ABI helpers that need no device, and the host-side sharding arithmetic must agree on EVERY input, not just the
fixtures.  Framing and segment indexing have to be bit exact (BASELINE.json north star)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import restatement as R
from sad_b200 import _lib
from sad_b200 import sharded as S
from tests import gpu_common as G


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 5_000_000), st.integers(1, 300_000), st.integers(1, 300_000))
def test_slice_count_matches_python_range_and_the_oracle(n, window, hop):
    """sad_slice_count == len(range(0, n - window + 1, hop)) (IR:184) == the oracle's start list."""
    lib = _lib.load()
    want = len(range(0, n - window + 1, hop))
    assert lib.sad_slice_count(n, window, hop) == want
    assert len(R.slice_starts(n, window, hop)) == want


@settings(max_examples=200, deadline=None)
@given(st.floats(0.0, 0.99), st.sampled_from([8000, 16000, 22050, 32000, 44100, 48000]), st.floats(0.5, 8.0))
def test_window_and_hop_follow_the_reference_float_arithmetic(overlap, sr, seconds):
    """window = int(window_size * sr), hop = int((1 - overlap) * window) in Python float arithmetic (IR:180-181)."""
    window, hop = R.window_and_hop(sr, seconds, overlap)
    assert window == int(seconds * sr) and hop == int((1 - overlap) * window)


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 8).flatmap(lambda n: st.lists(st.floats(-0.5, 0.5, width=32), min_size=n + 1, max_size=n + 1)),
       st.sampled_from([0.3, 0.5, 0.7]))
def test_decision_rule_and_margin(logits, thr):
    """The oracle's rule (IR:207-213) on probabilities; at threshold 0.5 a perturbation smaller than decision_margin never
    changes the decision (what the GPU tests use to judge flips)."""
    z = np.array([logits], dtype=np.float32)
    s = 1.0 / (1.0 + np.exp(-z.astype(np.float64)))
    lab = R.decide_from_probs(s[0].astype(np.float32), np.float32(thr))
    n = z.shape[1] - 1
    if s[0, -1] >= thr and (s[0, :-1] < thr).all():
        assert lab == n
    else:
        assert lab == int(np.argmax(s[0, :-1].astype(np.float32)))
    if thr == 0.5:
        m = G.decision_margin(z)[0]
        if m > 1e-4:
            rs = np.random.RandomState(abs(hash(tuple(logits))) % (2 ** 31))
            for _ in range(8):
                zp = (z + rs.uniform(-0.9 * m, 0.9 * m, size=z.shape)).astype(np.float32)
                lab_p, _ = R.interpret(torch.from_numpy(zp), 0.5)
                lab_0, _ = R.interpret(torch.from_numpy(z), 0.5)
                assert lab_p[0] == lab_0[0]


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 5000), st.integers(1, 16))
def test_clip_partition_is_a_contiguous_cover(n_clips, world):
    parts = S.clip_partition(n_clips, world)
    assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n_clips
    assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(parts, parts[1:]))
    sizes = [hi - lo for lo, hi in parts]
    assert max(sizes) <= -(-n_clips // world) if n_clips else max(sizes) == 0


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(0, 40), min_size=1, max_size=60), st.integers(1, 8))
def test_segment_ranges_tile_the_corpus(lengths, world):
    parts = S.clip_partition(len(lengths), world)
    ranges = [S.segment_range(lengths, lo, hi) for lo, hi in parts]
    assert ranges[0][0] == 0 and ranges[-1][1] == sum(lengths)
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 3_000_000), st.sampled_from([8000, 11025, 16000, 22050, 24000, 44100, 48000, 96000, 192000, 32000]))
def test_ingest_length_matches_torchaudios_rule(n_frames, sr):
    """sad_ingest_length = max(ceil_f32(32000 * n / sr), 128000) -- torchaudio's float32 ceil (functional.py) + IR:150-154."""
    lib = _lib.load()
    g = int(np.gcd(sr, 32000))
    want = n_frames if sr == 32000 else R.resample_length(n_frames, sr // g, 32000 // g)
    assert lib.sad_ingest_length(n_frames, sr) == max(want, 128000)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 6), st.integers(0, 10_000))
def test_reflect_framing_indices(frame, seed):
    """The oracle's reflect index (what the kernels implement) against torch's own reflect padding on a ramp."""
    n = R.WINDOW_SAMPLES
    x = torch.arange(n, dtype=torch.float32)
    xp = torch.nn.functional.pad(x.view(1, 1, -1), (1024, 1024), mode="reflect").view(-1)
    rs = np.random.RandomState(seed)
    for j in rs.randint(0, n + 2048, size=64):
        assert int(xp[j]) == R.reflect_pad_index(int(j))
