#!/bin/bash
# A/B a switch on ONE box: tools/ab_bench.sh VAR v1 v2 [rounds]   (prints one compact line per run)
VAR=$1; A=$2; B=$3; N=${4:-2}
for i in $(seq $N); do for v in $A $B; do
  env $VAR=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); pc=d['roofline']['per_conv_tflops']
        f=lambda ks:[round(pc[str(k)]) for k in ks if pc[str(k)]]
        print('$VAR=$v', 'seg/s', round(d['value']), 'ms', round(d['ms_per_step'],1), 'convTF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'], 'L2', f([5,6,8,9]), 'L3', f([10,11,13,14]), 'L4', f([15,16,18,19]))
"; done; done
