"""Seeded synthetic scenes for the tracker: textured plane (or smooth depth), known 6-DoF
motion, uint8-quantised frames, points kept in bounds at every pyramid level.

The reference has no fixtures of its own; its MATLAB recipe run_io_test.m:18-41 (random cloud +
ground-truth camera) is the model.  Everything here is numpy on the host: it only makes INPUTS, it is not part of
the tracking path.
"""
import numpy as np


def se3_exp(p):
    """Closed-form exp of p = [t, w] in float64 (same parameterisation as utilities.h:84-145)."""
    p = np.asarray(p, np.float64)
    u, w = p[:3], p[3:]
    th = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th > 1e-8:
        a, b, c = np.sin(th) / th, (1 - np.cos(th)) / th ** 2, (th - np.sin(th)) / th ** 3
    else:
        a, b, c = 1.0, 0.5, 1.0 / 6
    R = np.eye(3) + a * W + b * W @ W
    V = np.eye(3) + b * W + c * W @ W
    G = np.eye(4)
    G[:3, :3] = R
    G[:3, 3] = V @ u
    return G


def se3_log(G):
    R, t = G[:3, :3], G[:3, 3]
    th = np.arccos(np.clip(0.5 * (np.trace(R) - 1), -1, 1))
    if th < 1e-10:
        w = np.zeros(3)
    else:
        W = (th / (2 * np.sin(th))) * (R - R.T)
        w = np.array([W[2, 1], W[0, 2], W[1, 0]])
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-5:
        Vi = np.eye(3) - 0.5 * W + (1.0 / 12) * W @ W
    else:
        Vi = np.eye(3) - 0.5 * W + ((1 - th / (2 * np.tan(th / 2))) / th ** 2) * W @ W
    return np.concatenate([Vi @ t, w])


def _blur(a, sigma):
    """Separable Gaussian blur with reflect borders (scipy when present, else numpy)."""
    try:
        from scipy.ndimage import gaussian_filter
        return gaussian_filter(a, sigma, mode="reflect", truncate=3.0)
    except ImportError:
        pass
    r = int(3 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    a = np.pad(a, ((r, r), (0, 0)), mode="reflect")
    a = sum(k[i] * a[i:a.shape[0] - 2 * r + i] for i in range(2 * r + 1))
    a = np.pad(a, ((0, 0), (r, r)), mode="reflect")
    a = sum(k[i] * a[:, i:a.shape[1] - 2 * r + i] for i in range(2 * r + 1))
    return a


def make_texture(seed, h, w):
    """Multi-octave blurred uniform noise, rescaled to [0,255], uint8 — has gradient content at every pyramid level."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w))
    for sigma, amp in ((2.0, 1.0), (4.0, 1.5), (8.0, 2.0), (16.0, 3.0), (32.0, 4.0)):
        n = _blur(rng.random((h, w)) - 0.5, sigma)
        acc += amp * n / n.std()
    acc -= acc.min()
    acc *= 255.0 / acc.max()
    return np.round(acc).astype(np.uint8)


def _bilinear(tex, x, y):
    h, w = tex.shape
    x = np.clip(x, 0, w - 1.001)
    y = np.clip(y, 0, h - 1.001)
    x0 = np.floor(x).astype(np.int64)
    y0 = np.floor(y).astype(np.int64)
    fx, fy = x - x0, y - y0
    t = tex.astype(np.float64)
    return ((1 - fx) * (1 - fy) * t[y0, x0] + fx * (1 - fy) * t[y0, x0 + 1] + (1 - fx) * fy * t[y0 + 1, x0] +
            fx * fy * t[y0 + 1, x0 + 1])


class Scene:
    """Textured plane n . X = depth, n = (tilt_x, tilt_y, 1), in the world frame (= camera frame of pose 0); the
    default tilt (0, 0) is the fronto-parallel plane Z = depth, a non-zero tilt gives every pixel its own depth."""

    def __init__(self, seed, w, h, depth=5.0, margin=None, tilt=(0.0, 0.0)):
        self.w, self.h, self.depth = w, h, depth
        self.n = np.array([tilt[0], tilt[1], 1.0])
        self.margin = margin if margin is not None else max(32, w // 8)
        self.tex = make_texture(1000 + seed, h + 2 * self.margin, w + 2 * self.margin)
        self.fc = np.array([w, w], np.float32)                       # fx = fy = W
        self.cc = np.array([w / 2, h / 2], np.float32)
        self.wh = np.array([w, h], np.int32)

    def render(self, p):
        """Frame seen from pose p (world -> camera: Xc = R Xw + t), uint8, by inverse warping through the plane."""
        G = se3_exp(p)
        R, t = G[:3, :3], G[:3, 3]
        fx, fy, cx, cy = float(self.fc[0]), float(self.fc[1]), float(self.cc[0]), float(self.cc[1])
        u, v = np.meshgrid(np.arange(self.w, dtype=np.float64), np.arange(self.h, dtype=np.float64))
        d = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], 0).reshape(3, -1)
        Rtd = R.T @ d
        Rtt = R.T @ t
        lam = (self.depth + self.n @ Rtt) / (self.n @ Rtd)
        Xw = Rtd * lam - Rtt[:, None]
        xa = Xw[0] / Xw[2] * fx + cx + self.margin
        ya = Xw[1] / Xw[2] * fy + cy + self.margin
        if np.allclose(p, 0):
            return self.tex[self.margin:self.margin + self.h, self.margin:self.margin + self.w].copy()
        img = _bilinear(self.tex, xa, ya).reshape(self.h, self.w)
        return np.clip(np.round(img), 0, 255).astype(np.uint8)

    def random_motion(self, seed, scale=1.0):
        """t ~ U(-0.02,0.02)^3 * depth, w ~ U(-0.01,0.01)^3 rad at 640 px width, scaled by 640/W
        so that the image flow stays <~ 13 px at level 0 whatever the resolution (a 4-level pyramid recovers it)."""
        rng = np.random.default_rng(2000 + seed)
        scale = scale * 640.0 / self.w
        t = rng.uniform(-0.02, 0.02, 3) * self.depth * scale
        w = rng.uniform(-0.01, 0.01, 3) * scale
        return np.concatenate([t, w])

    def points(self, seed, n, psz, lv_f, p_ref=None):
        """n world points on the plane whose projection under p_ref lies in
        [2*psz*2^lv_f, W - 2*psz*2^lv_f] x [.., H - ..] (clamped to keep a usable area).  Returns SoA float64 [3n]."""
        rng = np.random.default_rng(3000 + seed)
        m = min(2 * psz * (1 << lv_f), self.w // 4, self.h // 4)
        u = rng.uniform(m, self.w - m, n)
        v = rng.uniform(m, self.h - m, n)
        fx, fy, cx, cy = float(self.fc[0]), float(self.fc[1]), float(self.cc[0]), float(self.cc[1])
        return self.backproject(u, v, p_ref)

    def backproject(self, u, v, p_ref=None):
        """World points on the plane seen at pixels (u, v) of the camera at pose p_ref; SoA float64 [3n]."""
        fx, fy, cx, cy = float(self.fc[0]), float(self.fc[1]), float(self.cc[0]), float(self.cc[1])
        u = np.asarray(u, np.float64).reshape(-1)
        v = np.asarray(v, np.float64).reshape(-1)
        d = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], 0)
        if p_ref is None or np.allclose(p_ref, 0):
            Xw = d * (self.depth / (self.n @ d))
        else:
            G = se3_exp(p_ref)
            R, t = G[:3, :3], G[:3, 3]
            Rtd, Rtt = R.T @ d, R.T @ t
            lam = (self.depth + self.n @ Rtt) / (self.n @ Rtd)
            Xw = Rtd * lam - Rtt[:, None]
        return np.ascontiguousarray(Xw.reshape(-1), np.float64)

    def dense_points(self, border, step=1):
        """One point per pixel of the reference frame (pose 0) inside a border: the dense-alignment template with
        per-pixel depth (BASELINE config 4: psz = 1).  Returns SoA float64 [3n]."""
        u, v = np.meshgrid(np.arange(border, self.w - border, step, dtype=np.float64),
                           np.arange(border, self.h - border, step, dtype=np.float64))
        return self.backproject(u, v)


def make_pair(seed, w=640, h=480, depth=5.0, motion_scale=1.0, tilt=(0.0, 0.0)):
    """(scene, frame A at pose 0, frame B at pose p_gt, p_gt).  The tracker starts from p=0 and must recover p_gt."""
    sc = Scene(seed, w, h, depth, tilt=tilt)
    p_gt = sc.random_motion(seed, motion_scale)
    return sc, sc.render(np.zeros(6)), sc.render(p_gt), p_gt


def make_sequence(seed, nframes, w=640, h=480, depth=5.0, motion_scale=0.5):
    """Frames along a smooth random walk of poses; returns (scene, [frames], poses[nframes,6])."""
    sc = Scene(seed, w, h, depth)
    poses = np.zeros((nframes, 6))
    G = np.eye(4)
    frames = [sc.render(poses[0])]
    for k in range(1, nframes):
        G = se3_exp(sc.random_motion(seed * 1000 + k, motion_scale)) @ G
        poses[k] = se3_log(G)
        frames.append(sc.render(poses[k]))
    return sc, frames, poses


def sqrt_image(w=160, h=120):
    """The commented synthetic image of run_io_test.m:8-14: img(i,j) = round(sqrt(i^2+j^2)) with 1-based i,j."""
    j, i = np.meshgrid(np.arange(1, w + 1, dtype=np.float64), np.arange(1, h + 1, dtype=np.float64))
    return np.clip(np.round(np.sqrt(i * i + j * j)), 0, 255).astype(np.uint8)
