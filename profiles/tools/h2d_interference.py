"""Does a concurrent pinned H2D copy slow the tracking kernel?  Device-resident step timed alone and with copies in flight."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R)
import numpy as np, torch
import invcompcamtrack_b200 as ict
import bench
S, T, P, psz, w, h, lv_f = 8, 4096, 4, 32, 1920, 1080, 3
wl = bench.make_workload(0, S, T, P, psz, w, h, lv_f, 4)
op = ict.make_optparam(lv_f=lv_f, lv_l=0, psz=psz, maxiter=10, normdp_ratio=0.01, donorm=0, dopatchnorm=0, maxpttrack=P)
dev = torch.device("cuda", 0); NT = S * T
frames = ict.Frames(2 * S, w, h, lv_f, psz); tr = ict.Tracker(op, wl["fc"], wl["cc"], wl["wh"])
d_frames = torch.from_numpy(wl["frames"]).to(dev); d_pts = torch.from_numpy(wl["pts"]).to(dev); d_off = torch.from_numpy(wl["pt_off"]).to(dev)
d_ref = torch.from_numpy(wl["ref"]).to(dev); d_new = torch.from_numpy(wl["new"]).to(dev)
d_pin = torch.zeros(NT, 6, dtype=torch.float64, device=dev); d_pout = torch.zeros_like(d_pin)
st = torch.cuda.current_stream().cuda_stream
frames.build_dev(0, 2 * S, d_frames.data_ptr(), u8=True, stream=st)
tr.set_points_dev(NT, d_off.data_ptr(), d_pts.data_ptr(), NT * P, P, stream=st)
def track():
    tr.track_batch_dev(frames, d_ref.data_ptr(), d_new.data_ptr(), d_pin.data_ptr(), d_pout.data_ptr(), stream=st)
hbuf = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); dbuf = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
side = torch.cuda.Stream()
def timed(mode):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if mode == "h2d":
        with torch.cuda.stream(side):
            for _ in range(8): dbuf.copy_(hbuf, non_blocking=True)
    if mode == "d2d":
        with torch.cuda.stream(side):
            for _ in range(40): dbuf.copy_(dbuf.flip(0) if False else dbuf, non_blocking=True)
    e0.record(); track(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for _ in range(2): timed("none")
for mode in ("none", "h2d", "none", "h2d", "none"):
    print(mode, "%.3f ms" % timed(mode))
