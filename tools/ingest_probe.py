"""Run sad_ingest a few times on a long synthetic stream (for ncu / timing).  python tools/ingest_probe.py [sr] [channels]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sad_b200.engine import Engine
sr = int(sys.argv[1]) if len(sys.argv) > 1 else 44100
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = Engine(1, torch.device("cuda", 0), max_batch=1)
pcm = torch.randint(-20000, 20000, (sr * 1500, ch), dtype=torch.int16, device="cuda")
for _ in range(3):
    y = eng.ingest(pcm, sr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = eng.ingest(pcm, sr)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"sr {sr} ch {ch}: {ms:.3f} ms, {(pcm.numel() * 2 + y.numel() * 4) / 1e9 / (ms / 1e3):.0f} GB/s")
