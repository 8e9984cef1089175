"""BASELINE config 4: dense full-frame alignment (one point per pixel, psz 1) through the multi-CTA path; timing."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case
w, h = int(os.environ.get("W", 1920)), int(os.environ.get("H", 1080))
case = make_case(seed=41, w=w, h=h, psz=1, lv_f=3, dense_border=int(os.environ.get("BORDER", 16)), tilt=(0.05, -0.03))
c = case
op = ict.OptParam.from_buffer_copy(bytes(c["op"]))
fr = ict.Frames(2, c["w"], c["h"], c["lv_f"], c["psz"]); fr.upload(0, np.stack([c["A"], c["B"]]))
tr = ict.Tracker(op, c["sc"].fc, c["sc"].cc, c["sc"].wh)
tr.set_sum_order(int(os.environ.get("SUM_ORDER", "0")))
pts = c["pts"].copy(); tr.set_points(c["pt_off"], pts)
for rep in range(3):
    t0 = time.perf_counter()
    out = tr.track_batch(fr, 0, 1, np.zeros((1, 6)), trace_cap=0)
    dt = time.perf_counter() - t0
    npix = int(out["npixres"].sum())
    print("points %d  iters %s  pixel-residuals %d  %.2f ms  %.3e pixel-residuals/s  err_vs_gt %.2e" % (
        c["npts"], out["iters"][0].tolist(), npix, 1e3 * dt, npix / dt, np.abs(out["p_out"][0] - c["p_gt"]).max()))
