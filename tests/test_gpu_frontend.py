"""GPU parity: front end (K1/K2a) through the C ABI vs the oracle and the reference's golden vectors.

Tolerance (BASELINE.json north_star): log-mel within 1e-4 relative in fp32, written as
|a-b| <= 1e-4 * max(|b|, 1) on dB values; framing / segment indexing bit exact.
"""
import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import restatement as R
from tests import gpu_common as G

pytestmark = pytest.mark.gpu


def test_logmel_vs_reference_golden():
    g = G.golden("frontend.npz")
    x = G.segs(g["seg_ids"])
    e = G.engine(2)
    db, ms = e.logmel(x.cuda())
    db = db.cpu().numpy()
    err = G.rel_db_err(db, g["logmel_db"])
    print("log-mel vs golden: max |diff| dB", np.abs(db - g["logmel_db"]).max(), "max tol-units", err.max(),
          "p99.99", np.quantile(err, 0.9999))
    assert err.max() <= 1.0
    np.testing.assert_allclose(ms[:, 0].cpu().numpy(), g["mu"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(ms[:, 1].cpu().numpy(), g["sigma"], rtol=0, atol=2e-5)


def test_image_vs_reference_golden():
    g = G.golden("frontend.npz")
    x = G.segs(g["seg_ids"])
    img = G.engine(2).image(x.cuda()).cpu().numpy()
    # normalised units: 1e-4 dB-relative tolerance divided by sigma (~10 dB) leaves ~1e-4 absolute; allow 5e-4
    np.testing.assert_allclose(img[:2], g["image_full"], rtol=0, atol=5e-4)
    np.testing.assert_allclose(img[:, ::7, ::5], g["image_sub"], rtol=0, atol=5e-4)


def test_logmel_vs_oracle_larger_batch_and_chunking():
    """32 more segments (incl. pure-tone / pure-noise draws), batch > max_batch so chunking is exercised."""
    x = FX.synth_segments(32, first=8)
    want = R.logmel_db(x).numpy()
    db, _ = G.engine(2).logmel(x.cuda())
    err = G.rel_db_err(db.cpu().numpy(), want)
    # the oracle itself sits up to ~1 tolerance-unit from exact arithmetic on near-floor cells of pure tones
    # (tests/test_oracle_golden.py::test_fp32_reference_noise_floor_vs_fp64), so compare against fp64 too
    want64 = R.logmel_db(x.double(), torch.float64).numpy()
    err64 = G.rel_db_err(db.cpu().numpy(), want64)
    ref64 = G.rel_db_err(want, want64)
    print("vs oracle fp32: max", err.max(), "p99.99", np.quantile(err, 0.9999), "| vs fp64: ours", err64.max(),
          "reference", ref64.max())
    assert np.quantile(err, 0.9999) <= 1.0
    assert err64.max() <= max(1.0, 2.0 * ref64.max())


def test_framing_with_impulses():
    """Integer framing / reflect padding: a unit impulse at sample p lights exactly the frames that contain the
    padded positions of p (and of its reflections), with the window value at the right offset."""
    pos = [0, 1, 5, 511, 512, 1023, 1024, 1025, 64000, 126975, 126976, 127487, 127998, 127999]
    x = torch.zeros(len(pos), R.WINDOW_SAMPLES)
    for i, p in enumerate(pos):
        x[i, p] = 1.0
    want = R.mel_power(x)                      # oracle, bit-identical to the reference framing
    e = G.engine(2)
    db, _ = e.logmel(x.cuda())
    db = db.cpu()
    want_db = R.amplitude_to_db(want)
    # frames with no energy are exactly at the floor in both; compare the support and the values
    sup_want = (want_db > want_db.amin(dim=(1, 2), keepdim=True) + 1e-3)
    sup_got = (db > db.amin(dim=(1, 2), keepdim=True) + 1e-3)
    assert torch.equal(sup_want.any(dim=1), sup_got.any(dim=1)), "different set of frames touched"
    err = G.rel_db_err(db.numpy(), want_db.numpy())
    assert err.max() <= 1.0


def test_slicing_gate_bit_exact():
    """slice_waveform (inference_runner.py:176-190) start indices / kept mask vs the reference goldens."""
    g = G.golden("slicing.npz")
    e = G.engine(2)
    from tests.test_oracle_golden import _clips
    for name, wf in _clips().items():
        for tag, overlap, thr in (("cli", 0.0, 1e-3), ("dflt", 0.85, 1e-4)):
            window, hop = R.window_and_hop(32000, 4.0, overlap)
            keep = e.slice_gate(wf.cuda(), window, hop, thr).cpu().numpy().astype(bool)
            starts = np.arange(keep.shape[0], dtype=np.int64) * hop
            np.testing.assert_array_equal(starts[keep], g[f"{name}.{tag}.starts"], err_msg=f"{name}.{tag}")
            if keep.any():
                st = torch.from_numpy(starts[keep]).cuda()
                got = e.gather_windows(wf.cuda(), st, window).cpu()
                assert torch.equal(got[0], wf[starts[keep][0]:starts[keep][0] + window])
                assert torch.equal(got[-1], wf[starts[keep][-1]:starts[keep][-1] + window])


def test_logmel_and_image_vs_reference_golden_256_segments():
    """BASELINE.json configs[1]: a >= 256-segment parity subset of the front end against the LIVE reference's transforms
    (tests/golden/frontend_256.npz: strided log-mel samples, mean / std / max / sum per segment, strided image samples)."""
    g = G.golden("frontend_256.npz")
    n, first, ms_ = g["mu"].shape[0], int(g["first"]), int(g["mel_stride"])
    assert n >= 256
    x = FX.synth_segments(n, first=first).cuda()
    e = G.engine(2, max_batch=64)
    db, ms = e.logmel(x)
    img = e.image(x)
    db = db.cpu().numpy()
    got = db[:, ::ms_][:, :, g["frames"]]
    err = G.rel_db_err(got, g["logmel_sample"])
    print(f"{n} segments vs the reference: log-mel max tol-units {err.max():.3f} p99.99 {np.quantile(err, 0.9999):.3f}; "
          f"|mu diff| {np.abs(ms[:, 0].cpu().numpy() - g['mu']).max():.2e} |sigma diff| "
          f"{np.abs(ms[:, 1].cpu().numpy() - g['sigma']).max():.2e}")
    assert err.max() <= 1.0
    np.testing.assert_allclose(ms[:, 0].cpu().numpy(), g["mu"], rtol=0, atol=3e-5)
    np.testing.assert_allclose(ms[:, 1].cpu().numpy(), g["sigma"], rtol=0, atol=3e-5)
    assert np.all(np.abs(db.max(axis=(1, 2)) - g["db_max"]) <= 1e-4 * np.maximum(np.abs(g["db_max"]), 1))
    np.testing.assert_allclose(db.astype(np.float64).sum(axis=(1, 2)), g["db_sum"], rtol=0, atol=32128 * 2e-4)
    np.testing.assert_allclose(img.cpu().numpy()[:, ::37, ::41], g["image_sample"], rtol=0, atol=5e-4)
