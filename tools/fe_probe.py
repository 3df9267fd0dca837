import sys, torch
sys.path.insert(0, '/root/repo')
from sad_b200.engine import Engine
from sad_b200 import synthetic as S
dev = torch.device('cuda', 0)
eng = Engine(1, dev, max_batch=256)
x = S.synth_pcm(256, 0, dev)
for _ in range(3):
    db, ms = eng.logmel(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.logmel(x)
e1.record(); torch.cuda.synchronize()
print("logmel 256 segs: %.3f ms -> %.0f seg/s" % (e0.elapsed_time(e1) / 10, 2560 / (e0.elapsed_time(e1) / 1e3)))
