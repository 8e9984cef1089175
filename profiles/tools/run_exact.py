"""One launch of the reference-order kernel (sum_order 1) on 4096 C3-geometry tracks — target for ncu captures."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case, gpu_run
case = make_case(seed=21, w=1920, h=1080, psz=32, npts=4, ntracks=4096)
for rep in range(2):
    g = gpu_run(ict, case, trace_cap=0, sum_order=int(os.environ.get("SUM_ORDER", "1")))
print("iters/track %.1f  pixel-residuals %d" % (g["iters"].sum(axis=1).mean(), g["npixres"].sum()))
