"""Batch timing of ict_track_batch for a patch size / points per track / track count in both summation orders,
with the oracle (1 thread, first 32 tracks) beside it.  python psz_time.py psz npts ntracks [w h]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case, oracle_run
from oracle import oracle as O
psz, npts, NT = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
w, h = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (1280, 704)
case = make_case(seed=21, w=w, h=h, psz=psz, npts=npts, ntracks=NT)
c = case
op = ict.OptParam.from_buffer_copy(bytes(c["op"]))
fr = ict.Frames(2, c["w"], c["h"], c["lv_f"], c["psz"])
fr.upload(0, np.stack([c["A"], c["B"]]))
tr = ict.Tracker(op, c["sc"].fc, c["sc"].cc, c["sc"].wh)
tr.set_points(c["pt_off"], c["pts"].copy())
p_in = np.zeros((NT, 6))
for order in (1, 0):
    tr.set_sum_order(order)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); out = tr.track_batch(fr, 0, 1, p_in); ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts)); npx = int(out["npixres"].sum())
    print("psz %d, %d points x %d tracks, sum_order %d: %8.3f ms  %.3e pixel-residuals/s  %.3e tracks/s" % (psz, npts, NT, order, dt * 1e3, npx / dt, NT / dt), flush=True)
    if order == 1: ref = out
sub = dict(case); n = min(32, NT); sub["T"] = n; sub["pt_off"] = case["pt_off"][:n + 1]; sub["pts"] = case["pts"][:3 * int(case["pt_off"][n])]
t0 = time.perf_counter(); o = oracle_run(O.OracleLib(), sub, trace_cap=0, nthreads=1); dt = time.perf_counter() - t0
print("oracle, 1 thread, %d tracks: %.3e pixel-residuals/s (incl. its pyramids); reference-order poses bit-identical: %s" %
      (n, int(o["npixres"].sum()) / dt, np.array_equal(o["p_out"], ref["p_out"][:n])))
