"""Small cases through every tracking kernel (K2v2, K2v8, K2x, K2x8, k_track_fast, k_track, dense fused + general
multi-CTA, NCC-free) — the target of compute-sanitizer runs."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import invcompcamtrack_b200 as ict
from helpers import make_case, gpu_run
cases = [("K2v8 / K2x8 / k_track<8>", dict(seed=1, w=320, h=240, psz=8, npts=37, lv_f=2, ntracks=3, dopatchnorm=1)),
         ("K2v2 / K2x / k_track<32>", dict(seed=2, w=640, h=384, psz=32, npts=4, lv_f=2, ntracks=3)),
         ("K2v2 512 threads", dict(seed=3, w=640, h=384, psz=32, npts=7, lv_f=2, ntracks=2)),
         ("k_track_fast<16>", dict(seed=4, w=320, h=240, psz=16, npts=9, lv_f=2, ntracks=2)),
         ("dense fused / general", dict(seed=5, w=320, h=240, lv_f=2, psz=1, dense_border=8, tilt=(0.1, -0.1)))]
for name, kw in cases:
    case = make_case(**kw)
    for so in (0, 1, 2):
        g = gpu_run(ict, case, trace_cap=8, sum_order=so)
        assert np.isfinite(g["p_out"]).all()
    print("ok", name, g["iters"][0].tolist())
