#!/bin/bash
# internal chunk size sweep on ONE box: tools/sweep_chunk.sh 128 148 256 296
for r in 1 2; do for mb in "$@"; do
  python bench.py --steps 3 --warmup 3 --max-batch $mb --no-cpu-baseline --no-e2e --no-library-baseline --no-ingest 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('max_batch $mb', round(d['value']), 'profiled', round(d['value_profiled']), 'convTF', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'],
      'fe/img/head', round(d['roofline_frontend']['kernel_ms_per_step'], 2), [round(x, 2) for x in d['other_ms_per_step'].values()])"
done; done
