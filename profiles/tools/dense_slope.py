"""Dense 1080p alignment: cost per iteration launch and fixed cost per TrackPose, from runs with maxiter = 2 and 10 and
a stop ratio that never triggers (all iterations run; 4 levels)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case
res = {}
for mi in (2, 10):
    c = make_case(seed=41, w=1920, h=1080, psz=1, lv_f=3, dense_border=16, tilt=(0.05, -0.03), maxiter=mi, ratio=1e-30)
    op = ict.OptParam.from_buffer_copy(bytes(c["op"]))
    fr = ict.Frames(2, c["w"], c["h"], c["lv_f"], c["psz"]); fr.upload(0, np.stack([c["A"], c["B"]]))
    tr = ict.Tracker(op, c["sc"].fc, c["sc"].cc, c["sc"].wh)
    tr.set_points(c["pt_off"], c["pts"].copy())
    best = 1e9
    for rep in range(6):
        t0 = time.perf_counter()
        out = tr.track_batch(fr, 0, 1, np.zeros((1, 6)), trace_cap=0)
        best = min(best, time.perf_counter() - t0)
    res[mi] = best
    print("maxiter %d: iters %s  %.3f ms" % (mi, out["iters"][0].tolist(), 1e3 * best))
per = (res[10] - res[2]) / 32
print("per iteration launch %.2f us (%.0f GB/s at 44 B per point: 40 B streamed + 4 B texel), fixed %.0f us per TrackPose" % (
    1e6 * per, c["npts"] * 44 / per / 1e9, 1e6 * (res[2] - 8 * per)))
