"""GPU edge cases of the C ABI: empty / single / ragged batches, call-order errors, invalid constants, non-default
streams, several contexts alive at once."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import fixtures as FX
from oracle import restatement as R
from tests import gpu_common as G
from sad_b200 import _lib
from sad_b200.engine import Engine

pytestmark = pytest.mark.gpu


def test_empty_single_and_ragged_batches():
    e = G.engine(2)
    x = FX.synth_segments(9, first=900).cuda()                       # max_batch = 8 -> chunks of 8 + 1
    lo9, pr9, la9 = e.forward_pcm(x, 0.5)
    lo0, pr0, la0 = e.forward_pcm(x[:0].contiguous(), 0.5)
    assert lo0.shape == (0, 3) and la0.shape == (0,)
    lo1, _, la1 = e.forward_pcm(x[8:9].contiguous(), 0.5)
    assert torch.equal(lo1[0], lo9[8]) and int(la1[0]) == int(la9[8])
    db0, ms0 = e.logmel(x[:0].contiguous())
    assert db0.shape == (0, 128, 251)
    lo_h, _, _ = e.forward_host(x[:0].cpu(), 0.5)
    assert lo_h.shape == (0, 3)
    cp, cl = e.clip_reduce(pr9, torch.zeros(9, dtype=torch.int32, device="cuda"), 1, 0.5)
    np.testing.assert_allclose(cp[0].cpu().numpy(), np.mean(list(pr9.cpu().numpy()), axis=0), rtol=0, atol=1e-6)


def test_threshold_is_honoured():
    e = G.engine(2)
    x = FX.synth_segments(6, first=910).cuda()
    for thr in (0.3, 0.5, 0.7):
        lo, pr, la = e.forward_pcm(x, thr)
        want = np.array([R.decide_from_probs(row, np.float32(thr)) for row in pr.cpu().numpy()], dtype=np.int32)
        np.testing.assert_array_equal(la.cpu().numpy(), want)


def test_call_order_and_argument_errors():
    lib = _lib.load()
    eng = Engine(2, torch.device("cuda", 0), max_batch=4)            # weights NOT loaded
    x = FX.synth_segments(1, first=1).cuda()
    with pytest.raises(_lib.SadError, match="SAD_ESTATE.*not loaded"):
        eng.forward_pcm(x, 0.5)
    with pytest.raises(_lib.SadError):
        eng.forward_pcm(x.cpu(), 0.5)                                # CPU tensor: no fallback
    with pytest.raises(ValueError):
        eng.forward_pcm(x[:, :1000].contiguous(), 0.5)
    with pytest.raises(KeyError):
        eng.load_head(0, {})
    sd = G.merged_sd(2)
    with pytest.raises(ValueError, match="2 sub-models"):
        Engine(3, torch.device("cuda", 0), max_batch=4).load_merged_state_dict(sd)
    ctx = C.c_void_p(0)
    assert lib.sad_create(C.byref(ctx), 99, 2, 4) == _lib.SAD_ENODEVICE
    assert lib.sad_create(C.byref(ctx), 0, 40, 4) == _lib.SAD_EINVAL
    assert lib.sad_create_ex(C.byref(ctx), 0, 2, 4, b"resnext50_32x4d") == _lib.SAD_EINVAL
    # a filterbank with weight above FFT bin 768 is rejected (the kernel only forms power for bins <= 768)
    fb = torch.zeros(1025, 128)
    fb[900, 5] = 1.0
    w = torch.hann_window(2048)
    code = lib.sad_set_frontend_constants(eng.ctx, C.c_void_p(w.data_ptr()), C.c_void_p(fb.data_ptr()))
    assert code == _lib.SAD_EINVAL and b"768" in lib.sad_last_error(eng.ctx)
    eng.close()


def test_non_default_stream_and_two_contexts():
    e2, e5 = G.engine(2), G.engine(5)                                # two contexts alive on one device
    x = FX.synth_segments(5, first=920).cuda()
    ref2 = e2.forward_pcm(x, 0.5)[0]
    ref5 = e5.forward_pcm(x, 0.5)[0]
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        a = e2.forward_pcm(x, 0.5)[0]
        b = e5.forward_pcm(x, 0.5)[0]
        c = e2.forward_pcm(x, 0.5)[0]
    s.synchronize()
    assert torch.equal(a, ref2) and torch.equal(c, ref2) and torch.equal(b, ref5)
    assert torch.equal(ref2[:, :2], ref5[:, :2])                      # heads 0,1 are the same seeded sub-models


def test_profile_counters():
    e = G.engine(2)
    x = FX.synth_segments(3, first=930).cuda()
    e.profile_enable(True)
    n0 = e.launches
    e.forward_pcm(x, 0.5)
    ms, n = e.profile_read()
    e.profile_enable(False)
    # front end (1), image, stem, 2 fused layer1 blocks + 12 convs (downsample folded into conv2), head + merge
    assert e.launches - n0 == 1 + 1 + 1 + 14 + 2
    assert sum(1 for v in n[:40] if v) == 15 and n[e.PROF_FRONTEND] == 1 and n[e.PROF_HEAD] == 1
    assert all(v >= 0 for v in ms) and sum(ms) > 0


def test_mixed_streams_on_one_context_do_not_race():
    """One context = one workspace: a call on the caller's stream followed, without a sync, by the host entry (which runs
    on the context's internal streams) and by a call on another torch stream must all see their own inputs
    (the library orders each call after the previous one on that context with an event)."""
    e = G.engine(2)
    xa = FX.synth_segments(8, first=940).cuda()
    xb = FX.synth_segments(8, first=948)
    ref_a = e.forward_pcm(xa, 0.5)[0].clone()
    ref_b = e.forward_host(xb, 0.5)[0].clone()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    for _ in range(3):
        a = e.forward_pcm(xa, 0.5)[0]                     # current stream, not synchronised ...
        b = e.forward_host(xb, 0.5)[0]                    # ... internal streams
        with torch.cuda.stream(s):
            c = e.forward_pcm(xa, 0.5)[0]                 # ... a third stream
        d = e.forward_pcm(xb.cuda(), 0.5)[0]
        torch.cuda.synchronize()
        assert torch.equal(a, ref_a) and torch.equal(b, ref_b) and torch.equal(c, ref_a) and torch.equal(d.cpu(), ref_b)


def test_current_device_is_left_alone():
    """Every entry point restores the caller's current CUDA device (matters with several GPUs in one process; on a
    one-GPU box this only checks that nothing changes)."""
    before = torch.cuda.current_device()
    e = Engine(1, torch.device("cuda", torch.cuda.device_count() - 1), max_batch=2)
    x = torch.zeros(1, 128000, device=e.device)
    e.logmel(x)
    e.close()
    assert torch.cuda.current_device() == before


def test_nan_window_is_kept_and_many_windows_gather():
    """slice gate: `piece.abs().max() < thr` is False for a window holding a NaN (IR:186), so it is kept; gather: more
    kept windows than gridDim.y allows (65535) -- the reference has no such limit."""
    e = G.engine(2)
    wf = torch.zeros(3 * 1000, device="cuda")
    wf[1500] = float("nan")
    keep = e.slice_gate(wf, 1000, 1000, 1e-3)
    assert keep.cpu().tolist() == [0, 1, 0]
    n, window = 70000, 64
    src = torch.arange(n + window, device="cuda", dtype=torch.float32)
    starts = torch.arange(n, device="cuda", dtype=torch.int64)
    out = e.gather_windows(src, starts, window)
    want = src.unfold(0, window, 1)[:n]
    assert torch.equal(out, want)


def test_ingest_accepts_element_aligned_stereo_views():
    """A stereo stream that starts at an odd element is only element aligned (not frame aligned); the C ABI states no
    alignment requirement."""
    e = G.engine(2)
    for dtype in (torch.float32, torch.int16):
        base = (torch.randn(2 * 5000 + 1, device="cuda") * 8000).to(dtype) if dtype == torch.int16 else \
            torch.randn(2 * 5000 + 1, device="cuda") * 0.2
        view = base[1:].view(5000, 2)                              # storage offset 1 element
        got = e.ingest(view, 32000)
        want = e.ingest(view.clone(), 32000)                       # the clone is allocation aligned
        assert torch.equal(got, want)
