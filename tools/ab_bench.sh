#!/bin/bash
# A/B a switch on ONE box: tools/ab_bench.sh VAR v1 v2 [rounds]   (prints one compact line per run)
# e.g. tools/ab_bench.sh SAD_PDL 0 1 2   or   tools/ab_bench.sh SAD_LIB $PWD/x/libsad_b200_prev.so $PWD/x/libsad_b200.so 2
VAR=$1; A=$2; B=$3; N=${4:-2}
for i in $(seq $N); do for v in $A $B; do
  env $VAR=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline --no-ingest 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); pc=d['roofline']['per_conv_tflops']
        f=lambda ks:[round(pc[str(k)]) for k in ks if pc[str(k)]]
        print('$VAR=' + '$v'.split('/')[-1], 'seg/s', round(d['value']), 'profiled', round(d['value_profiled']), 'ms', round(d['ms_per_step'],1), 'convTF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'], 'stem+L1', f([0,2,4]), 'L2', f([5,6,8,9]), 'L3', f([10,11,13,14]), 'L4', f([15,16,18,19]), 'fe/img/head ms', round(d['roofline_frontend']['kernel_ms_per_step'],2), [round(x,2) for x in d['other_ms_per_step'].values()])
"; done; done
