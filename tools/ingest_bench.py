"""Throughput of sad_ingest per input rate (stereo int16, stream larger than L2): GB/s of input + output bytes.

    python tools/ingest_bench.py                # pair kernel where it applies
    SAD_INGEST_PAIR=0 python tools/ingest_bench.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sad_b200.engine import Engine  # noqa: E402


def main():
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    eng = Engine(2, max_batch=1)                            # ingest needs no weights
    rows = []
    for sr in (44100, 48000, 22050, 16000, 8000, 24000, 96000, 88200, 32000):
        frames = sr * 1500                                     # 25 min of audio
        pcm = torch.randint(-20000, 20000, (frames, 2), dtype=torch.int16, device="cuda")
        for _ in range(3):
            y = eng.ingest(pcm, sr)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            y = eng.ingest(pcm, sr)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        nbytes = pcm.numel() * 2 + y.numel() * 4
        rows.append({"sr_in": sr, "ms": round(ms, 4), "GBps": round(nbytes / 1e9 / (ms / 1e3), 1),
                     "frac_hbm": round(nbytes / 1e9 / (ms / 1e3) / peaks["hbm_gbs"], 3)})
        del pcm, y
    print(json.dumps({"pair": os.environ.get("SAD_INGEST_PAIR", "1"), "rows": rows}))


if __name__ == "__main__":
    main()
