"""Throughput of the fused PCM -> decision path for any wired backbone (not the headline config):
   python tools/backbone_probe.py resnet50 [heads] [batch] [chunk]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sad_b200.engine import Engine
from sad_b200 import synthetic as S
from oracle import fixtures as FX            # weights only (random init, no calibration); tools/ is not product code

name = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 6
B = int(sys.argv[3]) if len(sys.argv) > 3 else 512
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 32
dev = torch.device("cuda", 0)
eng = Engine(H, dev, max_batch=chunk, backbone=name)
eng.load_merged_state_dict(FX.merged_state_dict(H, calibrate=False, backbone=name))
x = S.synth_pcm(B, 0, dev, seed=1)
for _ in range(2):
    eng.forward_pcm(x, 0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 3
for _ in range(K):
    eng.forward_pcm(x, 0.5)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
convs = [l for l in FX._layer_plan(name) if l[0] == "conv"]
print(f"{name}: {H} heads, batch {B} (chunks of {chunk}): {ms:.1f} ms/step = {B / (ms / 1e3):.0f} segments/s, {len(convs)} convs per head")

# per-conv profile (slots >= 39 are aggregated for deep nets)
eng.profile_enable(True)
eng.forward_pcm(x, 0.5)
torch.cuda.synchronize()
ms_k, n_k = eng.profile_read()
eng.profile_enable(False)
plan = [l for l in FX._layer_plan(name) if l[0] == "conv"]
hw = {1: 128, 2: 64, 3: 32, 4: 16}
rows = []
for i, (_, nm, cout, cin, k) in enumerate(plan):
    if i == 0 or i >= 39:
        continue
    li = int(nm.split(".")[0][-1])
    first = nm.split(".")[1] == "0" and li > 1
    hout = hw[li] * (2 if (first and nm.endswith("conv1") and name in FX.BOTTLENECK) else 1)
    fl = 2.0 * hout * hout * cout * cin * k * k * H * B / 1e12
    rows.append((i, nm, cin, cout, k, hout, ms_k[i], fl / (ms_k[i] / 1e3) if ms_k[i] > 0 else 0.0))
tot = sum(ms_k[:40])
print(f"conv time {tot:.1f} ms of {ms:.1f}; front end {ms_k[eng.PROF_FRONTEND]:.1f}, image {ms_k[eng.PROF_IMAGE]:.1f}, head {ms_k[eng.PROF_HEAD]:.1f}; slot 39+ {ms_k[39]:.1f}")
for r in rows:
    if r[6] > 0:
        print("  conv %2d %-22s %4d->%4d k%d @%3d  %6.2f ms  %6.0f TFLOP/s" % r)
