#!/bin/bash
# front-end variants (different builds of the library) on ONE box: parity first, then the stand-alone line twice each
P=$PWD/synthetic-audio-detection_b200
for l in "$@"; do
  SAD_LIB=$P/$l timeout 300 python -m pytest tests/test_gpu_frontend.py -m gpu -q -x 2>&1 | tail -1
done
for r in 1 2; do for l in "$@"; do
  SAD_LIB=$P/$l python bench.py --frontend-only --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$l', round(d['value']), 'seg/s', round(d['ms_per_step'], 3), 'ms', round(d['roofline']['frac'], 4), d['clocks']['sm_mhz'])"
done; done
