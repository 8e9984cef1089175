#!/bin/bash
# parity tests, per-section cycle counts, then A/B of K2v2 settings against k_track_fast (ICT_FAST_V1=1) on one box
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_ab.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_ab.log
python profiles/tools/serial_vs_parallel_cycles.py
run() { python bench.py --steps 3 --warmup 2 --seqs 8 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$1 value %.4e kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms_per_launch']))"; }
run "v2 default"
for e in $V2_AB_ENVS; do env $e bash -c "$(declare -f run); run 'v2 $e'"; done
ICT_FAST_V1=1 run v1
