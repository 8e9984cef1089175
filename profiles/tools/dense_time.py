"""Dense 1080p TrackPose (BASELINE configs[3]) timing in both summation orders.  python dense_time.py"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R)
import numpy as np
import invcompcamtrack_b200 as ict
from invcompcamtrack_b200 import synth
w, h = 1920, 1080
sc, A, B, p_gt = synth.make_pair(41, w, h, tilt=(0.05, -0.03))
pts = sc.dense_points(16); n = pts.size // 3
op = ict.make_optparam(lv_f=3, lv_l=0, psz=1, maxiter=10, normdp_ratio=0.01, donorm=0, dopatchnorm=0, maxpttrack=n)
fr = ict.Frames(2, w, h, 3, 1); fr.upload(0, np.stack([A, B]))
tr = ict.Tracker(op, sc.fc, sc.cc, sc.wh); tr.set_points(np.array([0, n], np.int64), pts.copy())
for order in (1, 0):
    tr.set_sum_order(order)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter(); r = tr.track_batch(fr, 0, 1, np.zeros((1, 6))); best = min(best, time.perf_counter() - t0)
    print("sum_order %d: %.3f ms per TrackPose, iters %s, %.3e pixel-residuals/s" % (order, 1e3 * best, r["iters"].tolist(), r["npixres"].sum() / best))
