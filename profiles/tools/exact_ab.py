"""A/B of the reference-order kernels on C3-geometry tracks (4 points of 32x32 per track, 1080p pair):
K2r (default route of sum_order 1), K2x (tracker knob "no_k2r") and, for scale, K2v2 (sum_order 0).
Checks K2r against the oracle on the first NCHK tracks, then times ict_track_batch (host call, median of REPS).

    python profiles/tools/exact_ab.py [ntracks] [npts]
"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case, oracle_run, assert_bit_identical
from oracle import oracle as O

NT = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 4
NCHK, REPS = 64, 5
case = make_case(seed=21, w=1920, h=1080, psz=32, npts=NP, ntracks=NT)
c = case
op = ict.OptParam.from_buffer_copy(bytes(c["op"]))
fr = ict.Frames(2, c["w"], c["h"], c["lv_f"], c["psz"])
fr.upload(0, np.stack([c["A"], c["B"]]))
tr = ict.Tracker(op, c["sc"].fc, c["sc"].cc, c["sc"].wh)
tr.set_points(c["pt_off"], c["pts"].copy())
p_in = np.zeros((NT, 6))


def run(order, knob=None, trace_cap=0):
    tr.set_knob("no_k2r", 1 if knob == "no_k2r" else 0)
    tr.set_sum_order(order)
    ts = []
    out = None
    for _ in range(REPS if trace_cap == 0 else 1):
        t0 = time.perf_counter()
        out = tr.track_batch(fr, 0, 1, p_in, trace_cap=trace_cap)
        ts.append(time.perf_counter() - t0)
    return out, float(np.median(ts))


sub = dict(case)
sub["T"] = NCHK
sub["pt_off"] = case["pt_off"][:NCHK + 1]
sub["pts"] = case["pts"][:3 * NP * NCHK]
o = oracle_run(O.OracleLib(), sub, trace_cap=48)
g, _ = run(1, trace_cap=48)
gs = {k: (v[:NCHK] if v is not None else None) for k, v in g.items()}
assert_bit_identical(gs, o)
print("K2r bit-identical to the oracle on %d tracks (trace, iters, npixres, poses)" % NCHK, flush=True)
res = {}
for name, order, knob in (("K2r", 1, None), ("K2x", 1, "no_k2r"), ("K2v2", 0, None)):
    out, dt = run(order, knob)
    res[name] = out
    npx = int(out["npixres"].sum())
    print("%-22s %8.3f ms  %.3e pixel-residuals/s  %.3e tracks/s  iters/track %.1f" %
          (name, dt * 1e3, npx / dt, NT / dt, out["iters"].sum(axis=1).mean()), flush=True)
assert np.array_equal(res["K2r"]["p_out"], res["K2x"]["p_out"]) and np.array_equal(res["K2r"]["iters"], res["K2x"]["iters"])
print("K2r == K2x on all %d tracks" % NT)
