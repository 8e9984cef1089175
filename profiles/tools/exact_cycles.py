"""Per-iteration SM cycles of the reference-order kernel from its own clock64() records (trace fields 22, 23):
chain loop and serial section, on NT C3-geometry tracks (all resident CTAs busy).  python exact_cycles.py [ntracks]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case, gpu_run
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
case = make_case(seed=21, w=1920, h=1080, psz=32, npts=4, ntracks=NT)
g = gpu_run(ict, case, trace_cap=48, sum_order=1)
tr = g["trace"]
ok = tr[..., 0] >= 0
print("records %d  iterations/track %.1f" % (ok.sum(), ok.sum() / NT))
print("chain loop cycles: mean %.0f  median %.0f  p10 %.0f p90 %.0f" % (tr[..., 23][ok].mean(), np.median(tr[..., 23][ok]),
      np.percentile(tr[..., 23][ok], 10), np.percentile(tr[..., 23][ok], 90)))
print("serial section cycles: mean %.0f  median %.0f" % (tr[..., 22][ok].mean(), np.median(tr[..., 22][ok])))
later = ok & (tr[..., 1] > 0)
print("serial: redux hand-off + solve: mean %.0f   pose update + exp: mean %.0f (iterations > 0)" % (tr[..., 17][ok].mean(), tr[..., 18][later].mean()))
for k in (16,):
    v = tr[..., k][ok]
    if v.any():
        print("field %d: mean %.0f median %.0f" % (k, v.mean(), np.median(v)))
first = ok & (tr[..., 1] == 0)
for k, name in ((19, "level: acquires + ref placement + window issue"), (20, "level: window wait + sampling + sd store"),
                (21, "level: Hessian"), (18, "level: factorisation + first placement")):
    v = tr[..., k][first]
    print("%-48s mean %.0f median %.0f" % (name, v.mean(), np.median(v)))
