"""GPU decision parity at scale: identical Real/Synthetic decision on >= 99.9% of segments (BASELINE.json north star).

The goldens (tests/golden/decisions_n{2,5,6}.npz) hold the LIVE reference's merged logits and labels
(oracle/make_golden.py: load_merged_model + waveform_to_spectrogram + interpret_multihead_logits, IR:194-214) on
4096 / 2048 / 2048 HELD-OUT segments of the class-structured corpus, with the v2 fixture whose read-outs were fitted
on other segments (oracle/fixtures.py).  The CUDA path runs the same PCM through the C ABI; the agreement assert is
strict -- no margin escape."""
import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from tests import gpu_common as G

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-2
MIN_AGREEMENT = 0.999


def _histogram(margin):
    edges = [0, 0.005, 0.01, 0.02, 0.04, 0.08, 0.16, 1e9]
    h, _ = np.histogram(margin, bins=edges)
    return ", ".join(f"<{e}: {c}" for e, c in zip([f"{v:g}" for v in edges[1:-1]] + ["inf"], h))


@pytest.mark.parametrize("n_heads,dtype", [(2, "bf16"), (5, "bf16"), (6, "bf16"), (2, "fp16")])
def test_decisions_match_the_reference_on_held_out_segments(n_heads, dtype):
    """bf16 is the named dtype; (2, fp16) is the A/B of the activation format on the same segments."""
    g = G.golden(f"decisions_n{n_heads}.npz")
    want_logits = g["merged_logits"]
    want_labels = g["labels"].astype(np.int64)
    n = want_logits.shape[0]
    assert int(g["n_heads"]) == n_heads and n >= 2048
    from sad_b200.engine import Engine
    e = Engine(n_heads, torch.device("cuda", 0), max_batch=64, dtype=dtype)
    e.load_merged_state_dict(FX.decision_state_dict(n_heads))
    logits, labels = [], []
    first = int(g["first"])
    for b0 in range(0, n, 256):
        x, cls = FX.family_segments(min(256, n - b0), first + b0, n_classes=n_heads + 1)
        np.testing.assert_array_equal(cls, g["classes"][b0:b0 + x.shape[0]])      # same corpus as the golden run
        lo, _, la = e.forward_pcm(x.cuda(), 0.5)
        logits.append(lo.cpu().numpy())
        labels.append(la.cpu().numpy().astype(np.int64))
    e.close()
    logits = np.concatenate(logits)
    labels = np.concatenate(labels)
    d = np.abs(logits - want_logits)
    agree = labels == want_labels
    margin = G.decision_margin(want_logits)
    binary = (labels == n_heads) == (want_labels == n_heads)                       # Real vs any synthetic
    print(f"N={n_heads} ({dtype}): {n} held-out segments; max |logit diff| {d.max():.4f} (p99 {np.percentile(d, 99):.4f}, mean "
          f"{d.mean():.4f}); decision agreement {agree.mean():.5f} ({int((~agree).sum())} differ), Real-vs-synthetic "
          f"agreement {binary.mean():.5f}; reference decision-margin histogram: {_histogram(margin)}; "
          f"margins of differing segments {np.round(margin[~agree], 4).tolist()}")
    assert d.max() <= LOGIT_TOL
    assert agree.mean() >= MIN_AGREEMENT
    assert binary.mean() >= MIN_AGREEMENT


def test_continuous_corpus_flips_stay_inside_the_logit_tolerance():
    """The SURVEY 8(d) continuous noise/tone corpus has no class structure: the v2 read-outs see out-of-family inputs and
    many reference logits sit within rounding of the threshold, so agreement there is reported, not gated at 99.9% --
    but every differing decision must lie inside the 2e-2 logit tolerance band of the reference's own logits."""
    from oracle import restatement as R
    n_heads, n = 2, 96
    sd = FX.decision_state_dict(n_heads)
    x = FX.synth_segments(n, first=9000)
    img3 = R.waveform_to_image(x).unsqueeze(1).repeat(1, 3, 1, 1)
    want = torch.cat([R.ensemble_forward(img3[i:i + 16], sd) for i in range(0, n, 16)]).numpy()
    want_lab, _ = R.interpret(torch.from_numpy(want), 0.5)
    from sad_b200.engine import Engine
    e = Engine(n_heads, torch.device("cuda", 0), max_batch=32)
    e.load_merged_state_dict(sd)
    lo, _, la = e.forward_pcm(x.cuda(), 0.5)
    e.close()
    d = np.abs(lo.cpu().numpy() - want)
    agree = la.cpu().numpy() == want_lab
    margin = G.decision_margin(want)
    print(f"continuous corpus: max |logit diff| {d.max():.4f}; agreement {agree.mean():.4f}; margins of differing "
          f"segments {np.round(margin[~agree], 4).tolist()}; margin histogram: {_histogram(margin)}")
    assert np.all(margin[~agree] <= LOGIT_TOL)


def test_headline_shape_against_the_reference():
    """BASELINE.json configs[3] as the bench runs it -- ONE batch of 2048 segments, 6 heads, internal chunk 148 -- against
    the live reference's logits and labels for all 2048 held-out segments (tests/golden/decisions_n6.npz), through the
    device entry and through the host entry (pinned PCM -> H2D -> compute -> D2H), which must agree bit for bit."""
    n_heads = 6
    g = G.golden(f"decisions_n{n_heads}.npz")
    n = g["merged_logits"].shape[0]
    assert n == 2048
    x = torch.cat([FX.family_segments(256, int(g["first"]) + b0, n_classes=n_heads + 1)[0] for b0 in range(0, n, 256)])
    from sad_b200.engine import Engine
    e = Engine(n_heads, torch.device("cuda", 0), max_batch=148)
    e.load_merged_state_dict(FX.decision_state_dict(n_heads))
    lo, pr, la = e.forward_pcm(x.cuda(), 0.5)
    lo_h, pr_h, la_h = e.forward_host(x.pin_memory(), 0.5)
    e.close()
    assert torch.equal(lo.cpu(), lo_h) and torch.equal(pr.cpu(), pr_h) and torch.equal(la.cpu(), la_h)
    d = np.abs(lo_h.numpy() - g["merged_logits"])
    agree = la_h.numpy().astype(np.int64) == g["labels"].astype(np.int64)
    print(f"headline shape (2048 x 6 heads, chunk 148): max |logit diff| {d.max():.4f}; decisions identical on "
          f"{int(agree.sum())} / {n}")
    assert d.max() <= LOGIT_TOL
    assert agree.mean() >= MIN_AGREEMENT
