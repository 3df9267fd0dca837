"""Write tests/golden/decision_fixture.npz: BN statistics and fitted read-outs of the v2 "decision" fixture.

    python -m oracle.make_decision_fixture [first_head last_head]

TEST INFRASTRUCTURE.  Needs no reference: it only calibrates / fits seeded random-init heads on the seeded
class-structured corpus (oracle/fixtures.py).  ~3 min per head on 8 cores.  The decision goldens
(oracle/make_golden.py, live reference) are generated from the file this script writes.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import fixtures as FX          # noqa: E402


def main():
    lo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    hi = int(sys.argv[2]) if len(sys.argv) > 2 else FX.DEC_MAX_HEADS
    path = FX.decision_fixture_path()
    out = dict(np.load(path)) if os.path.exists(path) else {}
    for i in range(lo, hi):
        t0 = time.time()
        arrays = FX.fit_decision_head(i, log=lambda m: print(m, flush=True))
        for k, v in arrays.items():
            out[f"h{i}.{k}"] = v
        np.savez_compressed(path, **out)
        print(f"head {i}: {len(arrays)} arrays, {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
