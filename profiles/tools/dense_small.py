"""A small dense reference-order TrackPose (320x240, one point per pixel) — target for ncu launch lists."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case, gpu_run
case = make_case(seed=61, w=320, h=240, lv_f=2, psz=1, dense_border=8, tilt=(0.15, -0.1))
g = gpu_run(ict, case, trace_cap=0, sum_order=int(sys.argv[1]) if len(sys.argv) > 1 else 1)
print(g["iters"].tolist(), g["npixres"].sum())
