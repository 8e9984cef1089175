"""One launch of the reference-order kernel on NT C3-geometry tracks (default 1184 = 4 waves of 2 CTAs x 148 SMs) —
target for ncu captures.  python run_exact_small.py [ntracks] [sum_order]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case, gpu_run
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
case = make_case(seed=21, w=1920, h=1080, psz=32, npts=4, ntracks=NT)
g = gpu_run(ict, case, trace_cap=0, sum_order=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
print("iters/track %.1f  pixel-residuals %d" % (g["iters"].sum(axis=1).mean(), g["npixres"].sum()))
