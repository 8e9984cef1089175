"""Per-chunk timeline of the host-buffer (e2e) path of bench.py: CUDA events around every call of one step."""
import os, sys, time, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R)
import numpy as np, torch
import invcompcamtrack_b200 as ict
import bench

S, T, P, psz, w, h, lv_f = 32, 4096, 4, 32, 1920, 1080, 3
nchunk = int(os.environ.get("CHUNKS", "4"))
wl = bench.make_workload(0, S, T, P, psz, w, h, lv_f, 4)
op = ict.make_optparam(lv_f=lv_f, lv_l=0, psz=psz, maxiter=10, normdp_ratio=0.01, donorm=0, dopatchnorm=0, maxpttrack=P)
NT, L = S * T, lv_f + 1
frames = ict.Frames(2 * S, w, h, lv_f, psz)
h_frames = torch.from_numpy(wl["frames"]).pin_memory(); h_pts = torch.from_numpy(wl["pts"]).pin_memory()
h_pin = torch.zeros(NT, 6, dtype=torch.float64).pin_memory(); h_pout = torch.zeros(NT, 6, dtype=torch.float64).pin_memory()
h_iters = torch.zeros(NT, L, dtype=torch.int32).pin_memory(); h_npix = torch.zeros(NT, dtype=torch.int64).pin_memory()
h_ref = torch.from_numpy(wl["ref"]).pin_memory(); h_new = torch.from_numpy(wl["new"]).pin_memory()
lib = ict.lib(); v = C.c_void_p
bounds = [(c * S) // nchunk for c in range(nchunk + 1)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
if os.environ.get('STREAMS') == '1': streams[1] = streams[0]
trk = [ict.Tracker(op, wl["fc"], wl["cc"], wl["wh"]) for _ in range(nchunk)]
off = [torch.from_numpy(np.ascontiguousarray(wl["pt_off"][bounds[c] * T:bounds[c + 1] * T + 1] - wl["pt_off"][bounds[c] * T])).pin_memory() for c in range(nchunk)]

def step(ev=None):
    for c in range(nchunk):
        s_ = streams[c % len(streams)]; st = v(s_.cuda_stream)
        s0, s1 = bounds[c], bounds[c + 1]; t0, t1 = s0 * T, s1 * T
        def mark(name):
            if ev is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(s_); ev.append((c, name, e, time.perf_counter()))
        mark("start")
        lib.ict_frames_upload_u8_stream(frames.h_, 2 * s0, 2 * (s1 - s0), v(h_frames.data_ptr() + 2 * s0 * w * h), st); mark("frames")
        lib.ict_tracker_set_points_stream(trk[c].h_, t1 - t0, v(off[c].data_ptr()), v(h_pts.data_ptr() + 8 * 3 * int(wl["pt_off"][t0])), st); mark("points")
        lib.ict_track_batch_stream(trk[c].h_, frames.h_, v(h_ref.data_ptr() + 4 * t0), v(h_new.data_ptr() + 4 * t0), v(h_pin.data_ptr() + 48 * t0),
                                   v(h_pout.data_ptr() + 48 * t0), v(h_iters.data_ptr() + 4 * L * t0), v(h_npix.data_ptr() + 8 * t0), st); mark("track")

for _ in range(2): step()
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True); base.record(); 
for s_ in streams: s_.wait_event(base)
ev = []; h0 = time.perf_counter()
step(ev); step(ev)
torch.cuda.synchronize()
for c, name, e, ht in ev:
    print("chunk %d %-7s gpu %8.3f ms   host-issued %8.3f ms" % (c, name, base.elapsed_time(e), 1e3 * (ht - h0)))
