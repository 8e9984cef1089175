#!/bin/bash
for sw in "" 1 "" 1; do
  if [ -n "$sw" ]; then export ICT_SERIAL_WARP_LAST=1; else unset ICT_SERIAL_WARP_LAST; fi
  python bench.py --steps 3 --warmup 2 --seqs 8 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('serial_warp_last=$sw value %.4e kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms_per_launch']))"
done
