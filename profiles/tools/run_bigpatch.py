"""One track with thousands of 8x8 patches (the MATLAB harness passes all visible model points): too large for a CTA,
so the multi-CTA path runs.  Timing."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case
n = int(os.environ.get("NPTS", 2000))
c = make_case(seed=43, w=1280, h=704, psz=8, npts=n, lv_f=3, donorm=int(os.environ.get("DONORM", 0)),
              dopatchnorm=int(os.environ.get("PATCHNORM", 0)))
op = ict.OptParam.from_buffer_copy(bytes(c["op"]))
fr = ict.Frames(2, c["w"], c["h"], c["lv_f"], c["psz"]); fr.upload(0, np.stack([c["A"], c["B"]]))
tr = ict.Tracker(op, c["sc"].fc, c["sc"].cc, c["sc"].wh)
tr.set_sum_order(int(os.environ.get("SUM_ORDER", "0")))
tr.set_points(c["pt_off"], c["pts"].copy())
for rep in range(3):
    t0 = time.perf_counter()
    out = tr.track_batch(fr, 0, 1, np.zeros((1, 6)))
    dt = time.perf_counter() - t0
    npix = int(out["npixres"].sum())
    print("points %d psz 8: iters %s  %.2f ms  %.3e pixel-residuals/s  err_vs_gt %.2e" % (
        n, out["iters"][0].tolist(), 1e3 * dt, npix / dt, np.abs(out["p_out"][0] - c["p_gt"]).max()))
