#!/bin/bash
# A/B of k_track_fast variants on one box: rows per group (KT) x CTAs per SM; prints value per variant
for v in 0 1 2 3 4 5 0; do
  ICT_FAST_VARIANT=$v python bench.py --steps 3 --warmup 2 --seqs 8 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('variant $v value %.4e kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms_per_launch']))"
done
