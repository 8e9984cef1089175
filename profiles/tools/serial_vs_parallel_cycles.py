import sys
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case, gpu_run
case = make_case(seed=21, w=1920, h=1080, psz=32, npts=4, ntracks=4096)
g = gpu_run(ict, case, trace_cap=48, sum_order=int(os.environ.get('SUM_ORDER', '0')))
tr = g["trace"]; m = tr[...,0] >= 0
m0 = m & (tr[...,1] == 0)
print("level setup (Hessian sums + LU factor) cycles: median %.0f mean %.0f" % (np.median(tr[...,21][m0]), tr[...,21][m0].mean()))
print("records", m.sum(), "serial cycles mean %.0f median %.0f  parallel cycles mean %.0f median %.0f" % (tr[...,22][m].mean(), np.median(tr[...,22][m]), tr[...,23][m].mean(), np.median(tr[...,23][m])))
if tr[...,18][m].max() > 0:
    print("v2: barrier wait of the serial warp mean %.0f median %.0f; serial incl. placement mean %.0f median %.0f" % (tr[...,17][m].mean(), np.median(tr[...,17][m]), tr[...,18][m].mean(), np.median(tr[...,18][m])))
print("iterations per track %.1f" % (m.sum() / tr.shape[0]))
if tr[...,19][m0].max() > 0:
    print("v2: level placement + template gather cycles: median %.0f mean %.0f" % (np.median(tr[...,19][m0]), tr[...,19][m0].mean()))
