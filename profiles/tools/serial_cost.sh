#!/bin/bash
# upper bound of what a free serial section would buy: 8 forced iterations per level, with and without the solve/update
for dbg in "" 1; do
  if [ -n "$dbg" ]; then export ICT_DBG_SKIP_SERIAL=1; else unset ICT_DBG_SKIP_SERIAL; fi
  python bench.py --steps 3 --warmup 2 --seqs 8 --no-cpu --no-e2e --maxiter 8 --ratio 0 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('skip_serial=$dbg value %.4e kernel_ms %.3f iters %.1f' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['gn_iterations_per_track']))"
done
