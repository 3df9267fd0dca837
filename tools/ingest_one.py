"""A few sad_ingest launches of one input rate, for ncu: python tools/ingest_one.py [sr_in] [seconds]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sad_b200.engine import Engine  # noqa: E402

sr = int(sys.argv[1]) if len(sys.argv) > 1 else 44100
secs = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
eng = Engine(2, max_batch=1)
pcm = torch.randint(-20000, 20000, (sr * secs, 2), dtype=torch.int16, device="cuda")
for _ in range(3):
    y = eng.ingest(pcm, sr)
torch.cuda.synchronize()
print(sr, y.shape)
