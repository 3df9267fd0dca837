#!/usr/bin/env python3
"""Smallest program that shows every kernel of the fused path at the bench's shape: two chunks of 148 synthetic segments (one per SM)
through a 6-head engine (the first warms up).  Profiling target:

    python tools/one_chunk.py > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -s 20 -c 19 -o gpurun_out/r02_full python tools/one_chunk.py

Launch order per chunk: logmel, image, stem_fused, 2 x block_rows, 4 x conv_umma<128,2,TR>,
8 x conv_umma2<256>, head_mlp, merge_decide (19); one synth kernel precedes the first chunk."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch                                   # noqa: E402
from sad_b200 import synthetic as S            # noqa: E402
from sad_b200.engine import Engine             # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
heads = int(os.environ.get("HEADS", "6"))
eng = Engine(heads, dev, max_batch=148)
eng.load_merged_state_dict(S.random_merged_state_dict(heads, seed=0))
x = S.synth_pcm(296, 0, dev)
lo, pr, la = eng.forward_pcm(x, 0.5)
torch.cuda.synchronize()
print("launches", eng.launches, "labels", la[:8].tolist(), "logit[0]", [round(v, 4) for v in lo[0].tolist()])
eng.close()
