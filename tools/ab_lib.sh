#!/bin/bash
# A/B differently built libraries on ONE box: tools/ab_lib.sh libA.so libB.so ... (each run twice, interleaved)
for i in 1 2; do for lib in "$@"; do
  SAD_LIB=$PWD/synthetic-audio-detection_b200/$lib python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); pc=d['roofline']['per_conv_tflops']
        f=lambda ks:[round(pc[str(k)]) for k in ks if pc[str(k)]]
        print('$lib', 'seg/s', round(d['value']), 'ms', round(d['ms_per_step'],1), 'convTF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'], 'stem', f([0]), 'L1', f([1,2,3,4]), 'L2', f([5,6,8,9]))
"; done; done
