#!/bin/bash
for nt in 256 128 256; do for v in 4 0; do
  ICT_NT=$nt ICT_FAST_VARIANT=$v python bench.py --steps 3 --warmup 2 --seqs 8 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('nt $nt variant $v value %.4e kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms_per_launch']))"
done; done
