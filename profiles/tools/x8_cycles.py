"""Per-iteration SM cycles of the reference-order kernel for 8x8 patches (K2x8) from its own clock64() records
(trace fields 16-18): chain loop, the part of it spent waiting for the producers, serial section.
python x8_cycles.py [ntracks] [npts]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case, gpu_run
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 1
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 100
case = make_case(seed=21, w=640, h=480, psz=8, npts=NP, ntracks=NT)
g = gpu_run(ict, case, trace_cap=48, sum_order=1)
t0 = time.perf_counter(); g = gpu_run(ict, case, trace_cap=48, sum_order=1); dt = time.perf_counter() - t0
tr = g["trace"]
ok = tr[..., 0] >= 0
print("tracks %d points %d: records %d, iterations/track %.1f, whole run incl. upload %.2f ms" % (NT, NP, ok.sum(), ok.sum() / NT, dt * 1e3))
for k, name in ((16, "chain loop"), (17, "  of which waiting for the first round"), (19, "  of which loads + additions"), (18, "redux + solve + exp")):
    v = tr[..., k][ok]
    print("%-34s mean %.0f median %.0f" % (name, v.mean(), np.median(v)))
