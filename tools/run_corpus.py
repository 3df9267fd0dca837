#!/usr/bin/env python3
"""BASELINE.json configs[4]: the 3.8 M-segment synthetic corpus (118 750 ragged clips) sharded over N GPUs with one
per-clip gather.

    python tools/run_corpus.py --clips 118750                      (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/run_corpus.py --clips 118750 --out gpurun_out/corpus_n8.json

Prints (rank 0) one JSON line: segments/s over all ranks (device time, max over ranks), the gather time, and the sha256
of the per-clip probabilities + labels -- equal digests across N prove bit-identical per-clip decisions."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def compare(paths):
    """Bit-identity of per-clip probabilities and labels across runs (different GPU counts, subset vs full corpus)."""
    import numpy as np
    runs = []
    for p in paths:
        with open(p) as f:
            j = json.load(f)
        z = np.load(os.path.splitext(p)[0] + ".npz")
        runs.append((p, j, z["clip_probs"], z["clip_labels"]))
    runs.sort(key=lambda r: (r[1]["clips"], r[1]["n_gpus"]))
    base = runs[0]
    ok = True
    rows = []
    for p, j, probs, labels in runs:
        n = base[2].shape[0]
        same = bool(np.array_equal(probs[:n].view(np.uint32), base[2].view(np.uint32)) and np.array_equal(labels[:n], base[3]))
        ok &= same
        rows.append({"run": os.path.basename(p), "n_gpus": j["n_gpus"], "clips": j["clips"], "segments": j["segments"],
                     "segments_per_s": round(j["segments_per_s"], 1), "gather_ms": round(j["gather_ms"], 3),
                     "first_%d_clips_bit_identical_to_%s" % (n, os.path.basename(base[0])): same})
    one = [r for r in rows if r["clips"] == rows[0]["clips"]]
    t1 = next((r["segments_per_s"] for r in one if r["n_gpus"] == 1), None)
    for r in one:
        if t1:
            r["strong_scaling_efficiency"] = round(r["segments_per_s"] / (t1 * r["n_gpus"]), 4)
    print(json.dumps({"bit_identical_across_runs": ok, "runs": rows}, indent=1))
    return 0 if ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=118750)
    ap.add_argument("--mean-segments", type=int, default=32)
    ap.add_argument("--heads", type=int, default=6)
    ap.add_argument("--chunk", type=int, default=2072, help="segments generated + processed per pass (14 x 148)")
    ap.add_argument("--max-batch", type=int, default=148)
    ap.add_argument("--out", default=None, help="JSON summary; per-clip results go to the same path with .npz")
    ap.add_argument("--compare", nargs="+", default=None, help="JSON summaries of finished runs: check that per-clip "
                    "results are bit-identical (shorter runs against the prefix of longer ones)")
    args = ap.parse_args()
    if args.compare:
        return compare(args.compare)

    import numpy as np
    import torch
    import torch.distributed as dist
    from sad_b200 import corpus as CO
    from sad_b200 import synthetic as S
    from sad_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(args.heads, dev, max_batch=args.max_batch)
    eng.load_merged_state_dict(S.random_merged_state_dict(args.heads, seed=0))
    lengths = CO.clip_lengths(args.clips, args.mean_segments)
    warm = CO.clip_lengths(8 * world, args.mean_segments)          # warm-up: kernels, NCCL communicator
    CO.run_corpus(eng, warm, chunk=args.chunk)
    r = CO.run_corpus(eng, lengths, chunk=args.chunk)
    total = int(lengths.sum())
    labels = r["clip_labels"].cpu().numpy()
    out = {"config": "BASELINE.json configs[4]: synthetic corpus, %d clips (%d..%d segments, ragged), %d segments, %d heads"
                     % (args.clips, int(lengths.min()), int(lengths.max()), total, args.heads),
           "n_gpus": world, "segments": total, "clips": int(args.clips), "ms": r["ms"],
           "segments_per_s": total / (r["ms"] / 1e3), "gather_ms": r["gather_ms"],
           "digest": CO.digest(r["clip_probs"], r["clip_labels"]),
           "label_histogram": np.bincount(labels[labels >= 0], minlength=args.heads + 1).tolist(),
           "rank_segments": int(r["segment_labels"].shape[0])}
    if rank == 0:
        print(json.dumps(out))
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(out, f)
            np.savez_compressed(os.path.splitext(args.out)[0] + ".npz", clip_probs=r["clip_probs"].cpu().numpy(),
                                clip_labels=labels)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
