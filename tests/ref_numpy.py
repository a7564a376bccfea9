"""Independent float64 restatement of the PCL-1.7 ICP loop with scipy.spatial.cKDTree + numpy.linalg.

Used to pin the C oracle (which stands in for the absent PCL) and to generate the
committed golden vectors in tests/golden/.  Shares no code with oracle/*.c.
"""
import numpy as np
from scipy.spatial import cKDTree


def euler_T(x):
    a, b, g = x[0], x[1], x[2]
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(g), -np.sin(g), 0], [np.sin(g), np.cos(g), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = x[3:]
    return T


def icp_numpy(src, tgt, tgt_normals, iters=10, estimator="plane", max_corr_dist=0.0, guess=None):
    P = np.asarray(src, np.float64)[:, :3]
    Q = np.asarray(tgt, np.float64)[:, :3]
    Nn = None if tgt_normals is None else np.asarray(tgt_normals, np.float64)
    tree = cKDTree(Q)
    T = np.eye(4) if guess is None else np.array(guess, np.float64)
    inl, fit = 0, 0.0
    for _ in range(iters):
        X = P @ T[:3, :3].T + T[:3, 3]
        d, j = tree.query(X, k=1)
        ok = np.ones(len(X), bool)
        if max_corr_dist > 0:
            ok &= d * d <= max_corr_dist ** 2
        if estimator == "plane":
            ok &= Nn[j, 3] != 0
        p, q = X[ok], Q[j[ok]]
        inl, fit = int(ok.sum()), float(np.mean(d[ok] ** 2)) if ok.any() else 0.0
        if inl < 3:
            return None
        if estimator == "plane":
            n = Nn[j[ok], :3]
            J = np.concatenate([np.cross(p, n), n], 1)
            r = np.einsum("ij,ij->i", n, q - p)
            x = np.linalg.solve(J.T @ J, J.T @ r)
            D = euler_T(x)
        else:
            pb, qb = p.mean(0), q.mean(0)
            H = (p - pb).T @ (q - qb)
            U, S, Vt = np.linalg.svd(H)
            Dg = np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ U.T))])
            R = Vt.T @ Dg @ U.T
            D = np.eye(4)
            D[:3, :3] = R
            D[:3, 3] = qb - R @ pb
        T = D @ T
    return dict(T=T, inliers=inl, fitness=fit)
