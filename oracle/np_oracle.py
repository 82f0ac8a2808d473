"""Second, independent restatement of the reference path in numpy + scipy.spatial.cKDTree.

TEST INFRASTRUCTURE ONLY (same rules as oracle/oracle.py).  Its job is to validate the C oracle
(oracle/o3d_oracle.c): the two were written separately, use different data structures (cKDTree vs
an own KD-tree, np.unique vs a hash map, LAPACK eigh / solve vs the analytic eigen-solver / LDLT)
and must agree to ~1e-10 before either is used as ground truth.
PARITY UNPINNED: neither has been diffed against a real Open3D v0.15.1 binary.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def within(kind, p, center=(0, 0, 0), rmin=0.0, rmax=np.inf, zmin=-np.inf, zmax=np.inf):
    """core/src/croppers.cpp:121-165"""
    d = p - np.asarray(center, dtype=np.float64)
    if kind == "None":
        return np.ones(len(p), dtype=bool)
    if kind == "MaxRadius":
        return np.linalg.norm(d, axis=1) <= rmax
    if kind == "MinRadius":
        return np.linalg.norm(d, axis=1) >= rmin
    if kind == "MinMaxRadius":
        r = np.linalg.norm(d, axis=1)
        return (r <= rmax) & (r >= rmin)
    if kind == "Cylinder":
        return (p[:, 2] >= zmin) & (p[:, 2] <= zmax) & (np.linalg.norm(d[:, :2], axis=1) <= rmax)
    raise ValueError(kind)


def voxel_down_sample(xyz, voxel):
    """[O3D] PointCloud::VoxelDownSample; returns (keys sorted lexicographically, means)."""
    vmin = xyz.min(axis=0) - 0.5 * voxel
    keys = np.floor((xyz - vmin) / voxel).astype(np.int64)
    uk, inv = np.unique(keys, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    sums = np.zeros((len(uk), 3))
    np.add.at(sums, inv, xyz)
    cnt = np.bincount(inv, minlength=len(uk)).astype(np.float64)
    return uk, sums / cnt[:, None], cnt.astype(np.int64)


def estimate_normals(xyz, knn, radius):
    """[O3D] EstimateNormals(Hybrid) + NormalizeNormals + OrientNormalsTowardsCameraLocation(0).
    Uses LAPACK eigh (not the analytic solver) on a covariance centred on the neighbourhood mean."""
    tree = cKDTree(xyz)
    k = min(knn, len(xyz))
    d, idx = tree.query(xyz, k=k)
    if k == 1:
        d = d[:, None]; idx = idx[:, None]
    normals = np.zeros_like(xyz)
    for i in range(len(xyz)):
        # squared distances recomputed in fp64 for the strict d2 < r2 cut
        dd = ((xyz[idx[i]] - xyz[i]) ** 2).sum(axis=1)
        sel = idx[i][dd < radius * radius]
        if len(sel) >= 3:
            q = xyz[sel]
            c = q - q.mean(axis=0)
            cov = c.T @ c / len(sel)
            w, v = np.linalg.eigh(cov)
            n = v[:, 0]
        else:
            n = np.array([0.0, 0.0, 1.0])
        nn = np.linalg.norm(n)
        if nn > 0:
            n = n / nn
        if np.dot(n, -xyz[i]) < 0:
            n = -n
        normals[i] = n
    return normals


def rot_zyx(a, b, g):
    ca, sa, cb, sb, cg, sg = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(g), np.sin(g)
    Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
    Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
    Rz = np.array([[cg, -sg, 0], [sg, cg, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def icp_p2plane(src, tgt, tgt_nrm, r, init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """[O3D] RegistrationICP with TransformationEstimationPointToPlane (L2)."""
    T = np.eye(4) if init is None else np.array(init, dtype=np.float64)
    tree = cKDTree(tgt)
    pcd = src.copy()
    if not np.allclose(T, np.eye(4), rtol=0, atol=1e-12):
        pcd = pcd @ T[:3, :3].T + T[:3, 3]

    def evaluate(p):
        d, j = tree.query(p, k=1)
        d2 = ((p - tgt[np.minimum(j, len(tgt) - 1)]) ** 2).sum(axis=1)
        ok = (j < len(tgt)) & (d2 < r * r)
        n = int(ok.sum())
        if n == 0:
            return ok, j, 0.0, 0.0
        return ok, j, n / len(p), float(np.sqrt(d2[ok].sum() / n))

    ok, j, fit, rmse = evaluate(pcd)
    iters = 0
    for i in range(max_iter):
        if ok.any():
            vs = pcd[ok]; vt = tgt[j[ok]]; nt = tgt_nrm[j[ok]]
            res = ((vs - vt) * nt).sum(axis=1)
            J = np.hstack([np.cross(vs, nt), nt])
            JTJ = J.T @ J
            JTr = J.T @ res
            x = np.linalg.solve(JTJ, -JTr)
            U = np.eye(4)
            U[:3, :3] = rot_zyx(x[0], x[1], x[2])
            U[:3, 3] = x[3:]
        else:
            U = np.eye(4)
        T = U @ T
        pcd = pcd @ U[:3, :3].T + U[:3, 3]
        bfit, brmse = fit, rmse
        ok, j, fit, rmse = evaluate(pcd)
        iters = i + 1
        if abs(bfit - fit) < rel_fitness and abs(brmse - rmse) < rel_rmse:
            break
    return T, fit, rmse, int(ok.sum()), iters


def transform(T, xyz, nrm=None):
    """core/src/helpers.cpp:273-305, including the near-identity duplication quirk."""
    ident = np.abs(T - np.eye(4)).max() < 1e-4
    h = np.hstack([xyz, np.ones((len(xyz), 1))]) @ T.T
    p = h[:, :3] / h[:, 3:4]
    n = None if nrm is None else nrm @ T[:3, :3].T
    if ident:
        p = np.vstack([xyz, p])
        n = None if nrm is None else np.vstack([nrm, n])
    return p, n


def voxelize_within_cropping_volume(voxel, inside_mask, xyz, nrm):
    """core/src/helpers.cpp:115-183; returns pass-through part and a dict key -> (point, normal)."""
    inv = 1.0 / voxel
    out = {}
    acc = {}
    for i in np.nonzero(inside_mask)[0]:
        k = tuple(np.floor(xyz[i] * inv).astype(np.int64))
        a = acc.setdefault(k, [np.zeros(3), np.zeros(3), 0])
        a[0] = a[0] + xyz[i]
        if not np.isnan(nrm[i]).any():
            a[1] = a[1] + nrm[i]
        a[2] += 1
    for k, (sp, sn, c) in acc.items():
        n = sn / c
        z = np.linalg.norm(n)
        out[k] = (sp / c, n / z if z > 0 else n)
    return xyz[~inside_mask], nrm[~inside_mask], out


def carve(map_xyz, map_nrm, inside_mask, scan_map_frame, sensor, voxel, max_len, trunc, min_dot):
    """Submap::carve -> getIdxsOfCarvedPoints (core/src/helpers.cpp:235-271) restated with numpy: all rays are marched
    together, one step of all rays at a time; voxels are looked up through a dict of integer keys."""
    inv = 1.0 / voxel
    vox = {}
    for i in np.nonzero(inside_mask)[0]:
        vox.setdefault(tuple(np.floor(map_xyz[i] * inv).astype(np.int64)), []).append(int(i))
    d = scan_map_frame - sensor
    length = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    u = d / length[:, None]
    mp = np.maximum(voxel, np.minimum(length - trunc, max_len))
    nn = np.sqrt((map_nrm * map_nrm).sum(axis=1))
    nhat = np.where(nn[:, None] > 0, map_nrm / np.where(nn > 0, nn, 1.0)[:, None], map_nrm)
    removed = np.zeros(len(map_xyz), dtype=bool)
    dist = 0.0
    alive = np.ones(len(d), dtype=bool)
    while True:
        alive &= dist < mp
        if not alive.any():
            break
        rays = np.nonzero(alive)[0]
        pos = dist * u[rays] + sensor
        keys = np.floor(pos * inv).astype(np.int64)
        for r, k in zip(rays, map(tuple, keys)):
            ids = vox.get(k)
            if ids is None:
                continue
            for j in ids:
                if abs(float(u[r] @ nhat[j])) > min_dot:
                    removed[j] = True
        dist += voxel
    return removed


def icp_p2point(src, tgt, r, init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """[O3D] RegistrationICP with TransformationEstimationPointToPoint = Eigen::umeyama(no scaling), LAPACK SVD."""
    T = np.eye(4) if init is None else np.array(init, dtype=np.float64)
    tree = cKDTree(tgt)
    pcd = src.copy()
    if not np.allclose(T, np.eye(4), rtol=0, atol=1e-12):
        pcd = pcd @ T[:3, :3].T + T[:3, 3]

    def evaluate(p):
        d, j = tree.query(p, k=1)
        d2 = ((p - tgt[np.minimum(j, len(tgt) - 1)]) ** 2).sum(axis=1)
        ok = (j < len(tgt)) & (d2 < r * r)
        n = int(ok.sum())
        if n == 0:
            return ok, j, 0.0, 0.0
        return ok, j, n / len(p), float(np.sqrt(d2[ok].sum() / n))

    ok, j, fit, rmse = evaluate(pcd)
    iters = 0
    for i in range(max_iter):
        U = np.eye(4)
        if ok.any():
            vs = pcd[ok]; vt = tgt[j[ok]]
            ms, mt = vs.mean(axis=0), vt.mean(axis=0)
            sigma = (vt - mt).T @ (vs - ms) / len(vs)
            Us, S, Vt = np.linalg.svd(sigma)
            D = np.eye(3)
            if np.linalg.det(Us) * np.linalg.det(Vt) < 0:
                D[2, 2] = -1.0
            U[:3, :3] = Us @ D @ Vt
            U[:3, 3] = mt - U[:3, :3] @ ms
        T = U @ T
        pcd = pcd @ U[:3, :3].T + U[:3, 3]
        bfit, brmse = fit, rmse
        ok, j, fit, rmse = evaluate(pcd)
        iters = i + 1
        if abs(bfit - fit) < rel_fitness and abs(brmse - rmse) < rel_rmse:
            break
    return T, fit, rmse, int(ok.sum()), iters


def gicp_covariances_from_normals(nrm, eps=1e-3):
    """[O3D] InitializePointCloudForGeneralizedICP, normals branch: Rx diag(eps,1,1) Rx^T with Rx = GetRotationFromE1ToX(n)."""
    out = np.empty((len(nrm), 3, 3))
    for i, x in enumerate(nrm):
        e1 = np.array([1.0, 0.0, 0.0])
        v = np.cross(e1, x); c = float(e1 @ x)
        if c < -0.99:
            Rx = np.eye(3)
        else:
            sv = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
            Rx = np.eye(3) + sv + (sv @ sv) * (1.0 / (1.0 + c))
        out[i] = Rx @ np.diag([eps, 1.0, 1.0]) @ Rx.T
    return out


def icp_gicp(src, src_nrm, tgt, tgt_nrm, r, init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6, eps=1e-3):
    """[O3D] RegistrationGeneralizedICP, written with the matrix square root W = (Ct + Cs)^-1/2 exactly as
    TransformationEstimationForGeneralizedICP::ComputeTransformation states it (rows of W [-skew(vs) | I], W d)."""
    T = np.eye(4) if init is None else np.array(init, dtype=np.float64)
    tree = cKDTree(tgt)
    pcd = src.copy()
    Cs = gicp_covariances_from_normals(src_nrm, eps); Ct = gicp_covariances_from_normals(tgt_nrm, eps)
    if not np.allclose(T, np.eye(4), rtol=0, atol=1e-12):
        pcd = pcd @ T[:3, :3].T + T[:3, 3]
        Cs = T[:3, :3] @ Cs @ T[:3, :3].T

    def evaluate(p):
        d, j = tree.query(p, k=1)
        d2 = ((p - tgt[np.minimum(j, len(tgt) - 1)]) ** 2).sum(axis=1)
        ok = (j < len(tgt)) & (d2 < r * r)
        n = int(ok.sum())
        if n == 0:
            return ok, j, 0.0, 0.0
        return ok, j, n / len(p), float(np.sqrt(d2[ok].sum() / n))

    ok, j, fit, rmse = evaluate(pcd)
    iters = 0
    for i in range(max_iter):
        U = np.eye(4)
        if ok.any():
            JTJ = np.zeros((6, 6)); JTr = np.zeros(6)
            for k in np.nonzero(ok)[0]:
                vs = pcd[k]; vt = tgt[j[k]]
                M = Ct[j[k]] + Cs[k]
                w, V = np.linalg.eigh(np.linalg.inv(M))
                W = V @ np.diag(np.sqrt(w)) @ V.T                      # symmetric positive definite square root
                A = np.hstack([-np.array([[0, -vs[2], vs[1]], [vs[2], 0, -vs[0]], [-vs[1], vs[0], 0]]), np.eye(3)])
                J = W @ A; res = W @ (vs - vt)
                JTJ += J.T @ J; JTr += J.T @ res
            x = np.linalg.solve(JTJ, -JTr)
            U[:3, :3] = rot_zyx(x[0], x[1], x[2]); U[:3, 3] = x[3:]
        T = U @ T
        pcd = pcd @ U[:3, :3].T + U[:3, 3]
        Cs = U[:3, :3] @ Cs @ U[:3, :3].T
        bfit, brmse = fit, rmse
        ok, j, fit, rmse = evaluate(pcd)
        iters = i + 1
        if abs(bfit - fit) < rel_fitness and abs(brmse - rmse) < rel_rmse:
            break
    return T, fit, rmse, int(ok.sum()), iters
