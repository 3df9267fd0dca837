#!/usr/bin/env python3
"""Race hunt: the whole fused path on the same input, many times, different batch shapes and streams -- every run must
reproduce the first one bit for bit (logits, probabilities, labels, log-mel, statistics).
   python tools/soak_path.py [rounds]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
from sad_b200 import synthetic as S            # noqa: E402
from sad_b200.engine import Engine             # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda", 0)
H = 6
eng = Engine(H, dev, max_batch=148)
eng.load_merged_state_dict(S.random_merged_state_dict(H, seed=0))
x = S.synth_pcm(700, 0, dev)                   # 4 chunks of 148 + one of 108
ref = eng.forward_pcm(x, 0.5)
db_ref, ms_ref = eng.logmel(x)
torch.cuda.synchronize()
bad = 0
side = torch.cuda.Stream()
for r in range(rounds):
    if r % 3 == 2:
        with torch.cuda.stream(side):
            out = eng.forward_pcm(x, 0.5)
            db, ms = eng.logmel(x)
        side.synchronize()
    else:
        out = eng.forward_pcm(x, 0.5)
        db, ms = eng.logmel(x)
    n = 1 + (r * 97) % 699                     # a ragged prefix through the same context in between
    sub = eng.forward_pcm(x[:n].contiguous(), 0.5)
    torch.cuda.synchronize()
    ok = all(torch.equal(a, b) for a, b in zip(out, ref)) and torch.equal(db, db_ref) and torch.equal(ms, ms_ref) \
        and all(torch.equal(a, b[:n]) for a, b in zip(sub, ref))
    bad += 0 if ok else 1
    if not ok:
        print("round", r, "MISMATCH")
print(f"soak: {rounds} rounds x (700 + prefix) segments, {bad} mismatches; launches {eng.launches}")
sys.exit(1 if bad else 0)
