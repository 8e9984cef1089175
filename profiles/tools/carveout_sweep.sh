run() { python bench.py --steps 3 --warmup 2 --seqs 8 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$1 value %.4e kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms_per_launch']))"; }
ICT_V2_VARIANT=3 ICT_V2_CARVEOUT=72 run "variant3 carveout72"
ICT_V2_VARIANT=3 ICT_V2_CARVEOUT=86 run "variant3 carveout86"
ICT_V2_VARIANT=3 run "variant3 carveout100"
ICT_V2_VARIANT=0 run "variant0 carveout100"
