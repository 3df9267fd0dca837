"""Soak the fused layer1 block (cross-CTA pipeline): many launches at full size, results must be bit-identical to the
two-launch path every time.  python tools/soak_fused.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sad_b200.engine import Engine
from sad_b200 import synthetic as S

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
H, B = 6, 128
sd = S.random_merged_state_dict(H, seed=0)
engs = {}
for flag in ("0", "1"):
    os.environ["SAD_FUSE_BLOCK"] = flag
    e = Engine(H, torch.device("cuda", 0), max_batch=B)
    e.load_merged_state_dict(sd)
    engs[flag] = e
x = S.synth_pcm(B, 0, torch.device("cuda", 0), seed=3)
ref = engs["0"].forward_pcm(x, 0.5)[0].clone()
bad = 0
for it in range(iters):
    got = engs["1"].forward_pcm(x, 0.5)[0]
    if not torch.equal(got, ref):
        bad += 1
        print("iteration", it, "differs: max", float((got - ref).abs().max()))
    if it % 3 == 0:                                       # vary the batch so unit counts / tails change
        nb = 1 + (it * 37) % B
        a = engs["1"].forward_pcm(x[:nb].contiguous(), 0.5)[0]
        b = engs["0"].forward_pcm(x[:nb].contiguous(), 0.5)[0]
        if not torch.equal(a, b):
            bad += 1
            print("iteration", it, "batch", nb, "differs")
torch.cuda.synchronize()
print(f"soak: {iters} iterations at {H} heads x {B} segments, mismatches: {bad}")
sys.exit(1 if bad else 0)
