#!/usr/bin/env python3
"""Wider accuracy survey than the unit tests: N segments, CUDA path vs the fp32 oracle (tools/, not a test)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import fixtures as FX, restatement as R
from sad_b200.engine import Engine
n_heads = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_seg = int(sys.argv[2]) if len(sys.argv) > 2 else 256
sd = FX.merged_state_dict(n_heads)
eng = Engine(n_heads, torch.device("cuda", 0), max_batch=64)
eng.load_merged_state_dict(sd)
x = FX.synth_segments(n_seg, first=5000)
lo, pr, la = eng.forward_pcm(x.cuda(), 0.5)
t0 = time.time()
want = torch.cat([R.ensemble_forward(R.waveform_to_image(x[i:i + 16]).unsqueeze(1).repeat(1, 3, 1, 1), sd) for i in range(0, n_seg, 16)])
d = (lo.cpu() - want).abs().numpy()
lab, _ = R.interpret(want, 0.5)
agree = la.cpu().numpy() == lab
z = want.numpy(); srt = np.sort(z[:, :-1], axis=1)
margin = np.minimum(np.abs(z).min(axis=1), srt[:, -1] - srt[:, -2]) if z.shape[1] > 2 else np.abs(z).min(axis=1)
print(f"heads {n_heads}, {n_seg} held-out segments (oracle {time.time() - t0:.0f} s): |logit diff| max {d.max():.4f} p99 {np.quantile(d, 0.99):.4f} "
      f"mean {d.mean():.4f}; decisions identical {agree.mean() * 100:.2f}% ({(~agree).sum()} differ, their margins {np.round(margin[~agree], 4)}); "
      f"logit std {z.std():.3f}")
