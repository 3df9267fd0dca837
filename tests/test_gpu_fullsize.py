"""Size-independent properties at BASELINE.json's full size (batch 2048, 6 heads) -- the oracle cannot run this many
segments in test time, so the CUDA path is checked against itself and against cheap exact invariants:
  * batch / chunk independence: a segment's logits do not depend on what else is in the batch (bit exact);
  * probs == sigmoid(logits); labels == rule IR:207-213 applied to the device probabilities;
  * clip probabilities == mean over the clip's segments, any clip partition sums to the same totals;
  * the host entry (H2D + compute + D2H, double-buffered chunks) returns exactly the device-entry result.
A sample of the batch is ALSO compared with the fp32 oracle (logit tolerance 2e-2)."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from sad_b200 import synthetic as S
from sad_b200.engine import Engine

pytestmark = pytest.mark.gpu

B, H = 2048, 6


@pytest.fixture(scope="module")
def setup():
    dev = torch.device("cuda", 0)
    eng = Engine(H, dev, max_batch=148)              # the bench's internal chunk: 13 chunks of 148 + one of 124
    sd = S.random_merged_state_dict(H, seed=3)
    eng.load_merged_state_dict(sd)
    x = S.synth_pcm(B, first=0, device=dev)
    lo, pr, la = eng.forward_pcm(x, 0.5)
    torch.cuda.synchronize()
    return eng, sd, x, lo, pr, la


def test_batch_and_chunk_independence(setup):
    eng, sd, x, lo, pr, la = setup
    g = torch.Generator().manual_seed(0)
    idx = torch.randperm(B, generator=g)[:37].to(x.device)          # ragged sub-batch, different chunk positions
    lo2, pr2, la2 = eng.forward_pcm(x[idx].contiguous(), 0.5)
    assert torch.equal(lo2, lo[idx]) and torch.equal(pr2, pr[idx]) and torch.equal(la2, la[idx])
    assert torch.isfinite(lo).all()


def test_chunk_size_independence(setup):
    """The same segments through a context with a different internal chunk (128, the round-1 value) give the same bits:
    what lets per-clip results be compared across runs, chunkings and GPU counts."""
    eng, sd, x, lo, pr, la = setup
    eng2 = Engine(H, x.device, max_batch=128)
    eng2.load_merged_state_dict(sd)
    lo2, _, la2 = eng2.forward_pcm(x[:300].contiguous(), 0.5)
    eng2.close()
    assert torch.equal(lo2, lo[:300]) and torch.equal(la2, la[:300])


def test_decision_rule_and_sigmoid_on_device_outputs(setup):
    eng, sd, x, lo, pr, la = setup
    lo_c, pr_c, la_c = lo.cpu(), pr.cpu().numpy(), la.cpu().numpy()
    np.testing.assert_allclose(pr_c, torch.sigmoid(lo_c).numpy(), rtol=0, atol=2e-7)
    want = np.array([R.decide_from_probs(row, np.float32(0.5)) for row in pr_c], dtype=np.int32)
    np.testing.assert_array_equal(la_c, want)
    assert float(lo_c.std(dim=0).min()) > 1e-4                         # outputs do depend on the input segment


def test_clip_reduce_and_host_entry(setup):
    eng, sd, x, lo, pr, la = setup
    clip = (torch.arange(B, dtype=torch.int32, device=x.device) // 32).contiguous()
    cp, cl = eng.clip_reduce(pr, clip, B // 32, 0.5)
    want = pr.view(B // 32, 32, H + 1)
    acc = torch.zeros(B // 32, H + 1, device=x.device)
    for j in range(32):                                                # sequential fp32 order, as numpy's mean(axis=0)
        acc = acc + want[:, j]
    assert torch.equal(cp, acc / 32)
    cp2, _ = eng.clip_reduce(pr, (clip // 2).contiguous(), B // 64, 0.5)
    np.testing.assert_allclose(cp2.cpu().numpy(), cp.view(B // 64, 2, H + 1).mean(1).cpu().numpy(), rtol=0, atol=1e-6)
    xh = x[:300].cpu().pin_memory()
    lo_h, pr_h, la_h = eng.forward_host(xh, 0.5)                      # 3 chunks: 148 + 148 + 4
    assert torch.equal(lo_h, lo[:300].cpu()) and torch.equal(la_h, la[:300].cpu())


def test_sample_against_fp32_oracle(setup):
    """Un-fitted random-init weights (bench fixture): logits are O(0.1); 2e-2 absolute still has to hold."""
    eng, sd, x, lo, pr, la = setup
    idx = [0, 511, 1024, 2047]
    xs = x[idx].cpu()
    img3 = R.waveform_to_image(xs).unsqueeze(1).repeat(1, 3, 1, 1)
    want = R.ensemble_forward(img3, sd)
    d = (lo[idx].cpu() - want).abs()
    print("full-size sample: max |logit diff|", d.max().item(), "logit scale", want.abs().max().item())
    assert d.max() <= 2e-2
