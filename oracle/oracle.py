"""ctypes bindings of the CPU ORACLE (test infrastructure, NOT the product).

Two libraries with the same call surface:
  * ``OracleLib``  -> oracle/libictrack_oracle.so, the plain-C restatement (oracle/ictrack_oracle.c);
  * ``RefLib``     -> oracle/_ref/libictrack_ref.so, the reference's own sources compiled against the stand-in
                      headers of oracle/shim/ (only buildable where /root/reference exists).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TRACE_FLOATS = 24
MAX_LEVELS = 8


class OptParam(C.Structure):
    """Bit-compatible with CTR::optparam (utilities.h:46-61) and ict_optparam (include/ictrack.h)."""
    _fields_ = [("maxpttrack", C.c_int), ("psz", C.c_int), ("pszd2", C.c_int), ("pszd2m3", C.c_int),
                ("novals", C.c_int), ("lv_f", C.c_int), ("lv_l", C.c_int), ("donorm", C.c_ubyte),
                ("dopatchnorm", C.c_ubyte), ("maxiter", C.c_int), ("normdp_ratio", C.c_float),
                ("verbosity", C.c_int)]


def make_optparam(lv_f=3, lv_l=0, psz=8, maxiter=10, normdp_ratio=0.01, donorm=0, dopatchnorm=0,
                  maxpttrack=100, verbosity=0):
    """Derived fields exactly as run_io_reprojection_test.cpp:112-127."""
    op = OptParam()
    op.lv_f, op.lv_l, op.psz = lv_f, lv_l, psz
    op.pszd2 = psz // 2
    op.pszd2m3 = psz + psz // 2 - 1
    op.novals = psz * psz
    op.maxiter, op.normdp_ratio = maxiter, normdp_ratio
    op.donorm, op.dopatchnorm = int(bool(donorm)), int(bool(dopatchnorm))
    if maxpttrack % 4:
        maxpttrack += 4 - maxpttrack % 4
    op.maxpttrack, op.verbosity = maxpttrack, verbosity
    return op


def pyramid_layout(w, h, lv_f, pad):
    off, sw, sh, tot = [], [], [], 0
    for l in range(lv_f + 1):
        sw.append((w >> l) + 2 * pad)
        sh.append((h >> l) + 2 * pad)
        off.append(tot)
        tot += sw[-1] * sh[-1]
    return tot, off, sw, sh


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def build(target="oracle"):
    """make -C oracle <target>; 'ref' needs /root/reference."""
    subprocess.run(["make", "-C", HERE, target], check=True, stdout=subprocess.DEVNULL)


class _Base:
    prefix = None

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self._keep = []

    def f(self, name):
        return getattr(self.lib, self.prefix + name)

    # -- small pieces ---------------------------------------------------------------------------------
    def camera_levels(self, noscales, fc, cc, wh, padding):
        out = np.zeros((noscales, 8), np.float32)
        fc = np.asarray(fc, np.float32); cc = np.asarray(cc, np.float32); wh = np.asarray(wh, np.int32)
        self.f("camera_levels")(C.c_int(noscales), _fp(fc), _fp(cc), wh.ctypes.data_as(C.POINTER(C.c_int)),
                                C.c_int(padding), _fp(out))
        return out

    def pyramid_build(self, img, lv_f, pad):
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape
        tot, off, sw, sh = pyramid_layout(w, h, lv_f, pad)
        I = np.zeros(tot, np.float32); dx = np.zeros(tot, np.float32); dy = np.zeros(tot, np.float32)
        self.f("pyramid_build")(_fp(img), C.c_int(w), C.c_int(h), C.c_int(lv_f), C.c_int(pad), _fp(I), _fp(dx), _fp(dy))
        return I, dx, dy

    def se3_exp(self, p, dtype=np.float32):
        p = np.ascontiguousarray(p, dtype); G = np.zeros(12, dtype)
        if dtype == np.float32:
            self.f("se3_exp_f")(_fp(G), _fp(p))
        else:
            self.f("se3_exp_d")(_dp(G), _dp(p))
        return G

    def se3_log(self, G, dtype=np.float32):
        G = np.ascontiguousarray(G, dtype); p = np.zeros(6, dtype)
        if dtype == np.float32:
            self.f("se3_log_f")(_fp(p), _fp(G))
        else:
            self.f("se3_log_d")(_dp(p), _dp(G))
        return p

    def getpatch(self, img_plane, mid, op, width):
        out = np.zeros(op.novals, np.float32)
        mid = np.asarray(mid, np.float32)
        self.f("getpatch")(_fp(img_plane), _fp(mid), _fp(out), C.byref(op), C.c_int(width))
        return out

    def getpatch_grad(self, I, dx, dy, mid, op, width):
        o = [np.zeros(op.novals, np.float32) for _ in range(3)]
        mid = np.asarray(mid, np.float32)
        self.f("getpatch_grad")(_fp(I), _fp(dx), _fp(dy), _fp(mid), _fp(o[0]), _fp(o[1]), _fp(o[2]), C.byref(op),
                                C.c_int(width))
        return o


class Odometer:
    """One CamClass+PoseClass+OdometerClass triple of either library."""

    def __init__(self, lib, op, fc, cc, wh):
        self.L = lib
        self.op = op                                   # borrowed by the C side: keep alive, mutations are seen
        self.fc = np.asarray(fc, np.float32); self.cc = np.asarray(cc, np.float32)
        self.wh = np.asarray(wh, np.int32)
        fn = lib.f("odom_create"); fn.restype = C.c_void_p
        self.h = C.c_void_p(fn(C.byref(op), _fp(self.fc), _fp(self.cc), self.wh.ctypes.data_as(C.POINTER(C.c_int))))
        self._planes = None

    def close(self):
        if self.h:
            self.L.f("odom_destroy")(self.h)
            self.h = None

    def set3dpoints(self, pts_soa):
        """pts_soa: float64 [3*n] X block, Y block, Z block; mutated in place when donorm."""
        assert pts_soa.dtype == np.float64 and pts_soa.flags.c_contiguous
        n = pts_soa.size // 3
        self.L.f("set3dpoints")(self.h, _dp(pts_soa), C.c_int(n))
        self.n = n

    def setpose(self, p_in, ref_planes, new_planes, lv_f, offs):
        """ref_planes = (I, dx, dy) flat plane sets, new_planes = (I, ...)."""
        p_in = np.ascontiguousarray(p_in, np.float64)
        PP = C.POINTER(C.c_float) * (lv_f + 1)

        def tab(a):
            return PP(*[C.cast(C.c_void_p(a.ctypes.data + 4 * offs[l]), C.POINTER(C.c_float)) for l in range(lv_f + 1)])
        self._planes = (ref_planes, new_planes, tab(ref_planes[0]), tab(ref_planes[1]), tab(ref_planes[2]),
                        tab(new_planes[0]))
        t = self._planes
        self.L.f("setpose")(self.h, _dp(p_in), t[2], t[3], t[4], t[5])

    def trackpose(self):
        p = np.zeros(6, np.float64)
        self.L.f("trackpose")(self.h, _dp(p))
        return p

    def get2dpoints(self):
        fn = self.L.f("get2dpoints"); fn.restype = C.POINTER(C.c_float)
        q = fn(self.h)
        M = self.op.maxpttrack
        a = np.ctypeslib.as_array(q, shape=(2 * M,)).copy()
        return np.concatenate([a[:self.n], a[M:M + self.n]])


class OracleLib(_Base):
    prefix = "ict_oracle_"

    def __init__(self, path=None, autobuild=True):
        path = path or os.path.join(HERE, "libictrack_oracle.so")
        if autobuild and not os.path.exists(path):
            build("oracle")
        super().__init__(path)
        self.lib.ict_oracle_sum.restype = C.c_float
        self.lib.ict_oracle_last_npixres.restype = C.c_int64
        self.lib.ict_oracle_last_iters.restype = C.POINTER(C.c_int)
        self.lib.ict_oracle_hessian.restype = C.POINTER(C.c_float)

    def set_sum_mode(self, m):
        self.lib.ict_oracle_set_sum_mode(C.c_int(m))

    def esum(self, a):
        a = np.ascontiguousarray(a, np.float32)
        return float(self.lib.ict_oracle_sum(_fp(a), C.c_int64(a.size)))

    def solve6(self, H, b):
        H = np.asfortranarray(H, np.float32); b = np.ascontiguousarray(b, np.float32); x = np.zeros(6, np.float32)
        self.lib.ict_oracle_solve6(H.ctypes.data_as(C.POINTER(C.c_float)), _fp(b), _fp(x))
        return x

    def track(self, od, trace_cap=0):
        """TrackPose with instrumentation: returns (p_out, iters[L], trace[n,16], npixres)."""
        L = od.op.lv_f - od.op.lv_l + 1
        trace = np.zeros((max(trace_cap, 1), TRACE_FLOATS), np.float32)
        self.lib.ict_oracle_odom_set_trace(od.h, _fp(trace) if trace_cap else None, C.c_int(trace_cap))
        p = od.trackpose()
        it = np.ctypeslib.as_array(self.lib.ict_oracle_last_iters(od.h), shape=(MAX_LEVELS,))[:L].copy()
        npx = int(self.lib.ict_oracle_last_npixres(od.h))
        self.lib.ict_oracle_odom_set_trace(od.h, None, C.c_int(0))
        return p, it, trace[trace[:, 0] >= 0] if trace_cap else trace[:0], npx

    def hessian(self, od):
        return np.ctypeslib.as_array(self.lib.ict_oracle_hessian(od.h), shape=(36,)).copy().reshape(6, 6)

    def track_batch(self, op, fc, cc, wh, planes_I, planes_dx, planes_dy, pt_off, pts, ref_frame, new_frame, p_in,
                    trace_cap=0, nthreads=0, want_pt2d=False):
        """planes_*: list of flat float32 plane sets (one per frame)."""
        fc = np.asarray(fc, np.float32); cc = np.asarray(cc, np.float32); wh = np.asarray(wh, np.int32)
        T = len(pt_off) - 1
        L = op.lv_f - op.lv_l + 1
        pt_off = np.ascontiguousarray(pt_off, np.int64); pts = np.ascontiguousarray(pts, np.float64)
        ref_frame = np.ascontiguousarray(ref_frame, np.int32); new_frame = np.ascontiguousarray(new_frame, np.int32)
        p_in = np.ascontiguousarray(p_in, np.float64)
        p_out = np.zeros((T, 6), np.float64); iters = np.zeros((T, L), np.int32)
        npix = np.zeros(T, np.int64)
        trace = np.zeros((T, max(trace_cap, 1), TRACE_FLOATS), np.float32)
        pt2d = np.zeros(2 * int(pt_off[-1]), np.float32)
        PP = C.POINTER(C.c_float) * len(planes_I)
        tI = PP(*[_fp(a) for a in planes_I]); tx = PP(*[_fp(a) for a in planes_dx]); ty = PP(*[_fp(a) for a in planes_dy])
        ip = C.POINTER(C.c_int)
        self.lib.ict_oracle_track_batch(
            C.byref(op), _fp(fc), _fp(cc), wh.ctypes.data_as(ip), C.c_int(len(planes_I)), tI, tx, ty, C.c_int(T),
            pt_off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(pts), ref_frame.ctypes.data_as(ip),
            new_frame.ctypes.data_as(ip), _dp(p_in), _dp(p_out), iters.ctypes.data_as(ip),
            _fp(trace) if trace_cap else None, C.c_int(trace_cap), npix.ctypes.data_as(C.POINTER(C.c_int64)),
            _fp(pt2d) if want_pt2d else None, C.c_int(nthreads))
        out = dict(p_out=p_out, iters=iters, npixres=npix, trace=trace if trace_cap else None)
        if want_pt2d:
            out["pt2d"] = pt2d
        return out

    def ncc_score(self, op, fc, cc, wh, img_b, img_r, img_f, nback, nfwd, pb, pr, pf):
        fc = np.asarray(fc, np.float32); cc = np.asarray(cc, np.float32); wh = np.asarray(wh, np.int32)
        n = pb.size // 2
        out = np.zeros(n, np.float32)
        self.lib.ict_oracle_ncc_score(C.byref(op), _fp(fc), _fp(cc), wh.ctypes.data_as(C.POINTER(C.c_int)), _fp(img_b),
                                      _fp(img_r), _fp(img_f), C.c_int(nback), C.c_int(nfwd), _fp(pb), _fp(pr),
                                      _fp(pf), C.c_int(n), _fp(out))
        return out


class RefLib(_Base):
    prefix = "ict_ref_"

    def __init__(self, path=None):
        path = path or os.path.join(HERE, "_ref", "libictrack_ref.so")
        super().__init__(path)

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libictrack_ref.so"))

    def track(self, od, trace_cap=64):
        """TrackPose of the unmodified reference; per-iteration (H, J^T r, delta_p) come from the shim's solve hook."""
        buf = np.zeros((max(trace_cap, 1), 48), np.float32)
        self.lib.ict_ref_set_solve_trace(_fp(buf), C.c_int(trace_cap))
        p = od.trackpose()
        n = int(self.lib.ict_ref_solve_trace_count())
        self.lib.ict_ref_set_solve_trace(None, C.c_int(0))
        return p, buf[:n]

    def track_batch(self, op, fc, cc, wh, planes_I, planes_dx, planes_dy, pt_off, pts, ref_frame, new_frame, p_in,
                    nthreads=1):
        fc = np.asarray(fc, np.float32); cc = np.asarray(cc, np.float32); wh = np.asarray(wh, np.int32)
        T = len(pt_off) - 1
        pt_off = np.ascontiguousarray(pt_off, np.int64); pts = np.ascontiguousarray(pts, np.float64)
        ref_frame = np.ascontiguousarray(ref_frame, np.int32); new_frame = np.ascontiguousarray(new_frame, np.int32)
        p_in = np.ascontiguousarray(p_in, np.float64)
        p_out = np.zeros((T, 6), np.float64)
        PP = C.POINTER(C.c_float) * len(planes_I)
        tI = PP(*[_fp(a) for a in planes_I]); tx = PP(*[_fp(a) for a in planes_dx]); ty = PP(*[_fp(a) for a in planes_dy])
        ip = C.POINTER(C.c_int)
        self.lib.ict_ref_track_batch(C.byref(op), _fp(fc), _fp(cc), wh.ctypes.data_as(ip), tI, tx, ty, C.c_int(T),
                                     pt_off.ctypes.data_as(C.POINTER(C.c_int64)), _dp(pts),
                                     ref_frame.ctypes.data_as(ip), new_frame.ctypes.data_as(ip), _dp(p_in), _dp(p_out),
                                     C.c_int(nthreads))
        return p_out
