"""Ground-truth error of the robustness modes (fast mode, psz 8): what each flag buys.  python robust_probe.py"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from invcompcamtrack_b200 import synth
from invcompcamtrack_b200.api import ROBUST_FULL_STEP, ROBUST_COMPOSE, ROBUST_FLOOR
from helpers import rot_angle_between


def run(sc, A, B, pts, npts, T, flags, maxiter=10, lv_f=3, ratio=0.01):
    op = ict.make_optparam(lv_f=lv_f, lv_l=0, psz=8, maxiter=maxiter, normdp_ratio=ratio, maxpttrack=npts)
    fr = ict.Frames(2, sc.w, sc.h, lv_f, 8); fr.upload(0, np.stack([A, B]))
    tr = ict.Tracker(op, sc.fc, sc.cc, sc.wh); tr.set_sum_order(0); tr.set_robust(flags)
    tr.set_points(np.arange(T + 1, dtype=np.int64) * npts, pts.copy())
    r = tr.track_batch(fr, 0, 1, np.zeros((T, 6)))
    tr.close(); fr.close()
    return r


def errs(r, p_gt):
    return (np.array([np.linalg.norm(p[:3] - p_gt[:3]) for p in r["p_out"]]),
            np.array([rot_angle_between(p, p_gt) for p in r["p_out"]]))


for scale in (1.0, 3.0, 5.0):
    sc, A, B, p_gt = synth.make_pair(77, 640, 480, motion_scale=scale)
    T, npts = 32, 100
    pts = np.concatenate([sc.points(500 + t, npts, 8, 3) for t in range(T)])
    for maxiter in (3, 10):
        for name, fl in (("reference", 0), ("full_step", ROBUST_FULL_STEP), ("compose", ROBUST_COMPOSE),
                         ("full_step+compose", ROBUST_FULL_STEP | ROBUST_COMPOSE)):
            r = run(sc, A, B, pts, npts, T, fl, maxiter=maxiter)
            et, er = errs(r, p_gt)
            print("motion x%.0f maxiter %2d %-18s median |t err| %.3e  median rot err %.3e  iterations/track %.1f"
                  % (scale, maxiter, name, np.median(et), np.median(er), r["iters"].sum(axis=1).mean()))
# integer reference coordinates >= 256: the reference's ceil(x + 1e-5f) places those template patches one pixel off
sc, A, B, p_gt = synth.make_pair(78, 640, 480)
T, npts = 16, 64
rng = np.random.default_rng(1)
pts = np.concatenate([sc.backproject(rng.integers(260, 600, npts).astype(np.float64), rng.integers(260, 440, npts).astype(np.float64))
                      for t in range(T)])
for name, fl in (("reference", 0), ("floor", ROBUST_FLOOR)):
    r = run(sc, A, B, pts, npts, T, fl, lv_f=0, maxiter=30, ratio=1e-3)
    et, er = errs(r, p_gt * 0 + p_gt)
    print("integer centres >= 256, level 0 only: %-10s median |t err| %.3e  median rot err %.3e" % (name, np.median(et), np.median(er)))
