#!/bin/bash
# is the un-profiled pass slower than the profiled one because of its position, or because of the events?
for o in up pu uup upu; do
  SAD_BENCH_ORDER=$o python bench.py --steps ${1:-5} --warmup 3 --no-cpu-baseline --no-e2e --no-library-baseline --no-ingest 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('order $o', 'value', round(d['value']), 'profiled', round(d['value_profiled']), 'repeat', [round(v) for v in (d['value_repeat'] or [])], 'clk', d['clocks']['sm_mhz'])"
done
