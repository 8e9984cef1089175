"""TEST INFRASTRUCTURE — numpy statement of the pose-hypothesis step.

func_ransac_fitcameras_odom.m:29-90 samples four 2D-3D correspondences, rejects degenerate samples, solves the pose
with the external ASPnP toolbox (not in the reference tree: parity unpinned for the solver itself — there is nothing to
restate or to run), reprojects all correspondences (:48-52) and keeps the hypotheses with at least four inliers
(:53-57).  This file states the algorithm the CUDA path (invcompcamtrack_b200/csrc/ict_hypotheses.cu) implements for
the solver slot — damped Gauss-Newton on the reprojection error of the four points in fp64, left-multiplied twists,
started from a common pose — plus the reference's degeneracy / inlier / rejection rules.  Only tests/ import it.
What pins it: the solved pose must reproject its four sample points (property), synthetic correspondences generated
from a known pose must be recovered (tests/test_hypotheses.py), and the CUDA path must agree with it to fp64 noise.
"""
import numpy as np


def _hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])


def se3_exp(p):
    """util_SE3_coeff_to_group<double>, utilities.h:84-145 (closed form, Taylor below sigma = 1e-4)."""
    p = np.asarray(p, np.float64)
    u, w = p[:3], p[3:]
    s = np.sqrt(w @ w)
    if s > 1e-4:
        a, b, c = np.sin(s) / s, (1 - np.cos(s)) / s ** 2, (s - np.sin(s)) / s ** 3
    else:
        s2 = s * s
        a = 1 - s2 / 6 * (1 - s2 / 20 * (1 - s2 / 42))
        b = .5 * (1 - s2 / 12 * (1 - s2 / 30 * (1 - s2 / 56)))
        c = (1 - s2 / 20 * (1 - s2 / 42 * (1 - s2 / 72))) / 6
    W = _hat(w)
    G = np.zeros((3, 4))
    G[:, :3] = np.eye(3) + a * W + b * W @ W
    G[:, 3] = (np.eye(3) + b * W + c * W @ W) @ u
    return G


def se3_log(G):
    """util_SE3_group_to_coeff<double>, utilities.h:149-241."""
    R, t = G[:, :3], G[:, 3]
    th = np.arccos(np.clip(0.5 * (np.trace(R) - 1), -1, 1))
    if th < 1e-10:
        w = np.zeros(3)
    else:
        Wm = (th / (2 * np.sin(th))) * (R - R.T)
        w = np.array([Wm[2, 1], Wm[0, 2], Wm[1, 0]])
    W = _hat(w)
    h = 1.0 / 12 if th < 1e-4 else (1 - th / (2 * np.tan(th / 2))) / th ** 2
    return np.concatenate([(np.eye(3) - 0.5 * W + h * W @ W) @ t, w])


def _cost(fc, cc, G, X, x):
    Xc = G[:, :3] @ X + G[:, 3:4]
    if not np.all(Xc[2] > 1e-12):
        return -1.0, None, Xc
    r = np.stack([Xc[0] / Xc[2] * fc[0] + cc[0] - x[0], Xc[1] / Xc[2] * fc[1] + cc[1] - x[1]], 1).reshape(-1)
    return float(r @ r), r, Xc


def degenerate(x4):
    """Three of the four image points (nearly) collinear; x4: [2, 4]."""
    for m in range(4):
        q = [n for n in range(4) if n != m]
        a, b = x4[:, q[1]] - x4[:, q[0]], x4[:, q[2]] - x4[:, q[0]]
        if not abs(a[0] * b[1] - a[1] * b[0]) > 1e-3 * np.sqrt(a @ a) * np.sqrt(b @ b):
            return True
    return False


def solve_sample(fc, cc, pt2d, pt3d, ids, p_init, inlthresh, maxiter=30):
    """Pose of one minimal sample: (ok, p[6])."""
    ids = list(ids)
    if len(set(ids)) < 4 or degenerate(pt2d[:, ids]):
        return False, np.zeros(6)
    X, x = pt3d[:, ids], pt2d[:, ids]
    G = se3_exp(p_init)
    cost, r, Xc = _cost(fc, cc, G, X, x)
    if cost < 0:
        return False, np.zeros(6)
    lam = 1e-4
    for _ in range(maxiter):
        J = np.zeros((8, 6))
        for m in range(4):
            xc, yc, zc = Xc[:, m]
            ju = np.array([fc[0] / zc, 0.0, -fc[0] * xc / zc ** 2])
            jv = np.array([0.0, fc[1] / zc, -fc[1] * yc / zc ** 2])
            for e, j in enumerate((ju, jv)):
                J[2 * m + e, :3] = j
                J[2 * m + e, 3:] = [-j[1] * zc + j[2] * yc, j[0] * zc - j[2] * xc, -j[0] * yc + j[1] * xc]
        H, g = J.T @ J, J.T @ r
        A = H + lam * np.diag(np.diag(H) + 1e-12)
        try:
            xi = np.linalg.solve(A, -g)
        except np.linalg.LinAlgError:
            return False, np.zeros(6)
        E = se3_exp(xi)
        Gn = np.zeros((3, 4))
        Gn[:, :3] = E[:, :3] @ G[:, :3]
        Gn[:, 3] = E[:, :3] @ G[:, 3] + E[:, 3]
        cn, rn, Xn = _cost(fc, cc, Gn, X, x)
        if 0 <= cn <= cost:
            done = cost - cn <= 1e-24 + 1e-16 * cost
            G, r, Xc, cost = Gn, rn, Xn, cn
            lam = max(lam * 0.1, 1e-12)
            if done:
                break
        else:
            lam *= 10
            if lam > 1e8:
                break
    if not cost <= 4 * inlthresh ** 2:
        return False, np.zeros(6)
    return True, se3_log(G)


def inliers(fc, cc, pt2d, pt3d, p, inlthresh):
    """func_ransac_fitcameras_odom.m:48-52: reprojection distance <= inlthresh."""
    G = se3_exp(p)
    Xc = G[:, :3] @ pt3d + G[:, 3:4]
    with np.errstate(divide="ignore", invalid="ignore"):
        du = Xc[0] / Xc[2] * fc[0] + cc[0] - pt2d[0]
        dv = Xc[1] / Xc[2] * fc[1] + cc[1] - pt2d[1]
    return (Xc[2] > 1e-12) & (np.sqrt(du * du + dv * dv) <= inlthresh)


def pose_hypotheses(fc, cc, pt2d, pt3d, sample_idx, p_init, inlthresh, maxiter=30):
    fc, cc = np.asarray(fc, np.float64), np.asarray(cc, np.float64)
    S, n = len(sample_idx), pt2d.shape[1]
    pose, status, ninl, mask = np.zeros((S, 6)), np.zeros(S, np.int32), np.zeros(S, np.int32), np.zeros((S, n), np.uint8)
    for s, ids in enumerate(sample_idx):
        ok, p = solve_sample(fc, cc, pt2d, pt3d, ids, p_init, inlthresh, maxiter)
        if ok:
            m = inliers(fc, cc, pt2d, pt3d, p, inlthresh)
            mask[s], ninl[s] = m, int(m.sum())
            ok = ninl[s] >= 4
            pose[s] = p
        status[s] = int(ok)
    return dict(pose=pose, status=status, ninl=ninl, mask=mask)
