#!/usr/bin/env python3
"""Library-kernel bar: the reference's own torch model on cuDNN / cuBLAS on this GPU, next to the native
path, same 6-head ensemble.  Run on a GPU box:

    python tools/library_bar.py [--segments 2048] > gpurun_out/library_bar.json

(i) fp32 NCHW, cudnn.deterministic, batches of 128 -- exactly the reference's GPU path (inference_runner.py:240-241,
284-288); (ii) bf16 autocast + channels_last; (iii) the same with cudnn.benchmark=True (the library's best case)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--segments", type=int, default=2048)
    ap.add_argument("--heads", type=int, default=6)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    import torch
    import bench
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    r = bench.library_bar(a.heads, a.segments, dev, reps=a.reps)
    r["gpu"] = torch.cuda.get_device_name(dev)
    r["torch"] = torch.__version__
    r["cudnn"] = torch.backends.cudnn.version()
    print(json.dumps(r))


if __name__ == "__main__":
    main()
