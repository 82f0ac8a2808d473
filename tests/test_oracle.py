"""CPU tests (-m "not gpu") that PIN THE ORACLE: oracle/o3d_oracle.c against the independent numpy/scipy restatement
(oracle/np_oracle.py), analytic known answers (SURVEY.md 8c tests 1-7) and the committed golden vectors.
The reference ships no tests or fixtures for this path, so this is all the pinning there is ("parity unpinned")."""
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle import np_oracle as NP
from oracle import oracle as O
from open3d_slam_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_config1_c_vs_numpy_and_golden():
    g = np.load(os.path.join(GOLD, "config1_icp.npz"))
    for tag, noise in (("clean", 0.0), ("noisy", 0.01)):
        src, tgt, nrm, T_true = synth.planar_cloud_config1(noise=noise)
        r = O.registration_icp_p2plane(src, tgt, nrm, 1.0, np.eye(4), max_iter=50, trace=True)
        T2, f2, e2, n2, i2 = NP.icp_p2plane(src, tgt, nrm, 1.0, max_iter=50)
        assert r.iters == i2 and r.n_corr == n2
        assert np.abs(r.T - T2).max() < 1e-10 and abs(r.inlier_rmse - e2) < 1e-12
        assert np.abs(r.T - g[f"{tag}_T"]).max() < 1e-12 and r.iters == int(g[f"{tag}_iters"])
        assert np.abs(r.trace - g[f"{tag}_trace"]).max() < 1e-9
    # noise-free: ICP recovers the displacement exactly (source = T_true * target  =>  result = T_true^-1)
    src, tgt, nrm, T_true = synth.planar_cloud_config1(noise=0.0)
    r = O.registration_icp_p2plane(src, tgt, nrm, 1.0, np.eye(4), max_iter=50)
    assert np.abs(r.T - np.linalg.inv(T_true)).max() < 1e-9


def test_one_step_closed_form():
    """SURVEY 8c test 1: exact correspondences, analytic normals -> JtJ, Jtr, x and T after one iteration by hand."""
    rng = np.random.default_rng(3)
    tgt = np.vstack([np.c_[rng.uniform(0, 5, (40, 2)), np.zeros(40)], np.c_[rng.uniform(0, 5, 40), np.zeros(40), rng.uniform(0, 5, 40)],
                     np.c_[np.zeros(40), rng.uniform(0, 5, (40, 2))]])
    nrm = np.vstack([np.tile([0, 0, 1.0], (40, 1)), np.tile([0, 1.0, 0], (40, 1)), np.tile([1.0, 0, 0], (40, 1))])
    T0 = synth.se3(0.002, -0.001, 0.003, (0.004, -0.003, 0.002))
    src = tgt @ T0[:3, :3].T + T0[:3, 3]
    fit, rmse, JTJ, JTr, corr = O.icp_evaluate_bruteforce(src, tgt, nrm, 10.0, np.eye(4))
    assert np.array_equal(corr, np.arange(len(tgt))) and fit == 1.0
    J = np.hstack([np.cross(src, nrm), nrm]); res = ((src - tgt) * nrm).sum(1)
    assert np.allclose(JTJ, J.T @ J, rtol=1e-13, atol=1e-12) and np.allclose(JTr, J.T @ res, rtol=1e-13, atol=1e-14)
    x = np.linalg.solve(J.T @ J, -(J.T @ res))
    assert np.allclose(O.ldlt6_solve(JTJ, -JTr), x, rtol=1e-10, atol=1e-14)
    r = O.registration_icp_p2plane(src, tgt, nrm, 10.0, np.eye(4), max_iter=1)
    U = np.eye(4); U[:3, :3] = NP.rot_zyx(*x[:3]); U[:3, 3] = x[3:]
    assert np.abs(r.T - U).max() < 1e-12 and np.abs(O.vec6_to_mat4(x) - U).max() < 1e-15


def test_ldlt_and_eigen_primitives():
    rng = np.random.default_rng(0)
    for _ in range(50):
        M = rng.normal(size=(6, 9)); A = M @ M.T; b = rng.normal(size=6)
        assert np.allclose(O.ldlt6_solve(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)
        C = rng.normal(size=(3, 7)); cov = C @ C.T / 7
        v = O.fast_eigen3x3(cov); w, V = np.linalg.eigh(cov)
        assert abs(abs(v @ V[:, 0]) - 1.0) < 1e-9
    assert np.array_equal(O.fast_eigen3x3(np.diag([3.0, 1.0, 2.0])), [0, 1, 0])
    assert np.array_equal(O.fast_eigen3x3(np.zeros((3, 3))), [0, 0, 0])
    assert np.array_equal(O.fast_eigen3x3(np.eye(3)), [0, 0, 1])      # identity covariance (< 3 neighbours) -> (0,0,1)


def test_kdtree_hybrid_vs_bruteforce():
    rng = np.random.default_rng(1)
    pts = rng.uniform(-5, 5, (3000, 3)); pts[100:110] = pts[100]     # exact duplicates: ties -> lower index
    q = rng.uniform(-6, 6, (200, 3)); q[:5] = pts[100]
    for (idx, d2), qi in zip(O.kdtree_search_hybrid(pts, q, 1.5, 12), q):
        dd = ((pts - qi) ** 2)
        dd = (dd[:, 0] + dd[:, 1]) + dd[:, 2]
        order = np.lexsort((np.arange(len(pts)), dd))[:12]
        order = order[dd[order] < 1.5 * 1.5]
        assert np.array_equal(idx, order) and np.array_equal(d2, dd[order])


def test_croppers_vs_numpy():
    rng = np.random.default_rng(2)
    p = rng.uniform(-20, 20, (5000, 3)); n = rng.normal(size=p.shape)
    for kind in ("MaxRadius", "MinRadius", "MinMaxRadius", "Cylinder"):
        for inv in (False, True):
            c = O.cropper(kind, 3.0, 12.0, -2.0, 4.0, (1.0, 2.0, -1.0), inv)
            x, nn = O.crop(c, p, n)
            m = NP.within(kind, p, (1.0, 2.0, -1.0), 3.0, 12.0, -2.0, 4.0) ^ inv
            assert np.array_equal(x, p[m]) and np.array_equal(nn, n[m])
    x, _ = O.crop(O.cropper("None"), p)
    assert np.array_equal(x, p)


def test_voxel_down_sample_vs_numpy_and_golden():
    g = np.load(os.path.join(GOLD, "scan_preprocess.npz"))
    raw = g["raw"].astype(np.float64)
    vx, _, keys = O.voxel_down_sample(raw, 0.1, return_keys=True)
    uk, means, cnt = NP.voxel_down_sample(raw, 0.1)
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    assert np.array_equal(keys[order], uk) and np.abs(vx[order] - means).max() < 1e-12
    assert np.array_equal(keys[order], g["voxel_keys"]) and np.array_equal(vx[order], g["voxel_means"])
    # floor semantics with negative coordinates and points exactly on voxel faces
    pts = np.array([[-0.25, 0.0, 0.0], [-0.2500001, 0, 0], [0.0, 0.25, -0.5], [0.0, 0.2499999, -0.5], [0.1, 0.1, 0.1], [0.1, 0.1, 0.1]])
    vx, _, keys = O.voxel_down_sample(pts, 0.25, return_keys=True)
    uk, means, cnt = NP.voxel_down_sample(pts, 0.25)
    assert len(vx) == len(uk) and sorted(map(tuple, keys)) == sorted(map(tuple, uk))
    # normals are averaged, NaN normals skipped but still counted
    nr = np.tile([0.0, 0.0, 1.0], (6, 1)); nr[4] = np.nan
    vx, vn = O.voxel_down_sample(pts, 10.0, nr)
    assert len(vx) == 1 and np.allclose(vn[0], [0, 0, 5.0 / 6.0])
    # voxel <= 0: o3d_slam::voxelize returns the cloud untouched
    same, _ = O.voxel_down_sample(pts, 0.0)
    assert np.array_equal(same, pts)


def test_normals_planes_and_degenerate():
    """SURVEY 8c test 4: planes at several orientations / ranges -> angle to the analytic normal < 1e-6 rad, sign rule."""
    rng = np.random.default_rng(4)
    for nvec, off in (([0, 0, 1.0], 3.0), ([1.0, 1.0, 0], 10.0), ([0.3, -0.5, 0.8], 25.0)):
        nvec = np.array(nvec) / np.linalg.norm(nvec)
        a = np.cross(nvec, [0.1, 0.2, 0.97]); a /= np.linalg.norm(a); b = np.cross(nvec, a)
        uv = rng.uniform(-2, 2, (400, 2))
        pts = off * nvec + uv[:, :1] * a + uv[:, 1:] * b
        got = O.estimate_normals(pts, 20, 3.0)
        ang = np.arccos(np.clip(np.abs(got @ nvec), -1, 1))
        assert ang.max() < 1e-6
        assert ((got * -pts).sum(1) >= 0).all()          # oriented towards the sensor at the origin
    pts = np.array([[1.0, 0, 0], [1.05, 0, 0], [50.0, 3, 1], [-20, 4, 2.0], [1.0, 0.05, 0.0]])
    got = O.estimate_normals(pts, 5, 0.5)
    assert np.array_equal(got[2], [0, 0, -1.0]) or np.array_equal(got[2], [0, 0, 1.0])   # < 3 neighbours -> +-(0,0,1)
    sub = O.estimate_normals(np.load(os.path.join(GOLD, "scan_preprocess.npz"))["voxel_means"][:1500], 20, 3.0)
    ref = NP.estimate_normals(np.load(os.path.join(GOLD, "scan_preprocess.npz"))["voxel_means"][:1500], 20, 3.0)
    assert np.abs((sub * ref).sum(1)).min() > 1 - 1e-8 and ((sub * ref).sum(1) > 0).all()


def test_random_down_sample_is_order_independent():
    rng = np.random.default_rng(5)
    p = rng.normal(size=(1000, 3)); n = rng.normal(size=(1000, 3))
    a, an = O.random_down_sample(p, 0.3, 7, n)
    assert len(a) == 300
    perm = rng.permutation(1000)
    b, bn = O.random_down_sample(p[perm], 0.3, 7, n[perm])
    assert sorted(map(tuple, a)) == sorted(map(tuple, b))            # same subset whatever the input order
    full, _ = O.random_down_sample(p, 1.0, 7)
    assert np.array_equal(full, p)
    c, _ = O.random_down_sample(p, 0.3, 8)
    assert sorted(map(tuple, a)) != sorted(map(tuple, c))            # the seed matters


def test_transform_quirk():
    rng = np.random.default_rng(6)
    p = rng.normal(size=(50, 3)); n = rng.normal(size=(50, 3))
    T = synth.se3(0.1, 0.2, 0.3, (1, 2, 3))
    x, xn = O.transform(T, p, n); rx, rn = NP.transform(T, p, n)
    assert x.shape == (50, 3) and np.allclose(x, rx, atol=1e-14) and np.allclose(xn, rn, atol=1e-14)
    x, xn = O.transform(np.eye(4), p, n)                             # near-identity: the cloud is emitted twice
    assert x.shape == (100, 3) and np.array_equal(x[:50], p) and np.array_equal(x[50:], p)


def test_fusion_hand_built():
    """SURVEY 8c test 5: 3-voxel map + 5-point scan, average-of-representatives, pass-through, normal re-normalisation."""
    v = 1.0
    map_x = np.array([[0.5, 0.5, 0.5], [1.5, 0.5, 0.5], [50.0, 0.5, 0.5]])          # third point is outside the cropper
    map_n = np.array([[0, 0, 1.0], [0, 1.0, 0], [1.0, 0, 0]])
    scan_x = np.array([[0.1, 0.1, 0.1], [0.9, 0.9, 0.9], [1.2, 0.2, 0.2], [2.5, 0.5, 0.5], [3.5, 0.5, 0.5]])
    scan_n = np.array([[0, 0, 1.0], [0, 0, 1.0], [np.nan, 0, 0], [0, 0, 2.0], [0, 3.0, 4.0]])
    T = synth.se3(0, 0, 0, (0.001, 0, 0))                                           # not identity (> 1e-4): no duplication
    c = O.cropper("MaxRadius", 0.0, 10.0)
    x, n, keys = O.submap_insert_scan(map_x, map_n, scan_x, scan_n, T, v, c, return_keys=True)
    d = {tuple(k): (p, q) for k, p, q in zip(keys, x, n)}
    assert tuple([-2 ** 31] * 3) in d and np.array_equal(d[tuple([-2 ** 31] * 3)][0], [50.0, 0.5, 0.5])   # pass-through
    p0, n0 = d[(0, 0, 0)]      # old representative counts ONCE: (0.5 + 0.101 + 0.901) / 3
    assert np.allclose(p0, [(0.5 + 0.101 + 0.901) / 3, (0.5 + 0.1 + 0.9) / 3, (0.5 + 0.1 + 0.9) / 3], atol=1e-15)
    assert np.allclose(n0, [0, 0, 1.0])
    p1, n1 = d[(1, 0, 0)]      # NaN normal skipped in the sum but counted in the divisor, then normalised
    assert np.allclose(p1, [(1.5 + 1.201) / 2, 0.35, 0.35]) and np.allclose(n1, [0, 1.0, 0])
    assert np.allclose(d[(2, 0, 0)][1], [0, 0, 1.0]) and np.allclose(d[(3, 0, 0)][1], [0, 0.6, 0.8])
    assert len(x) == 5
    # identity transform -> every scan point is inserted twice, the means do not move
    x2, n2, k2 = O.submap_insert_scan(map_x, map_n, scan_x, scan_n, np.eye(4), v, c, return_keys=True)
    d2 = {tuple(k): p for k, p in zip(k2, x2)}
    assert np.allclose(d2[(0, 0, 0)], [(0.5 + 2 * 0.1 + 2 * 0.9) / 5] * 3)
    # against the golden two-scan fusion and the numpy restatement
    g = np.load(os.path.join(GOLD, "fusion_two_scans.npz"))
    assert len(g["map_xyz"]) > 1000 and np.isfinite(g["map_xyz"]).all()
    inside = NP.within("MaxRadius", np.vstack([map_x, scan_x @ T[:3, :3].T + T[:3, 3]]), T[:3, 3], rmax=10.0)
    px, pn, vox = NP.voxelize_within_cropping_volume(v, inside, np.vstack([map_x, scan_x @ T[:3, :3].T + T[:3, 3]]), np.vstack([map_n, scan_n]))
    for k, (p, q) in vox.items():
        assert np.allclose(d[k][0], p, atol=1e-14) and np.allclose(d[k][1], q, atol=1e-14)


def test_dense_map_vs_bruteforce():
    rng = np.random.default_rng(7)
    pts = rng.uniform(-2, 2, (4000, 3))
    dm = O.DenseMap(0.25, 1 << 16); dm.insert(pts[:2500]); dm.insert(pts[2500:])
    x, n, keys = dm.to_cloud()
    k = np.floor(pts * 4.0).astype(np.int64)
    uk, inv = np.unique(k, axis=0, return_inverse=True); inv = inv.reshape(-1)
    sums = np.zeros((len(uk), 3)); np.add.at(sums, inv, pts); cnt = np.bincount(inv)
    o = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    assert np.array_equal(keys[o], uk) and np.abs(x[o] - sums / cnt[:, None]).max() < 1e-12


def test_icp_properties():
    """SURVEY 8c test 7: invariance under a common rigid motion, fitness in [0,1], rmse <= r, permutation independence."""
    src, tgt, nrm, _ = synth.planar_cloud_config1(n=800, noise=0.01)
    base = O.registration_icp_p2plane(src, tgt, nrm, 0.7, np.eye(4), max_iter=30)
    assert 0.0 <= base.fitness <= 1.0 and base.inlier_rmse <= 0.7
    G = synth.se3(0.3, -0.2, 1.0, (5.0, -3.0, 2.0))
    s2 = src @ G[:3, :3].T + G[:3, 3]; t2 = tgt @ G[:3, :3].T + G[:3, 3]; n2 = nrm @ G[:3, :3].T
    moved = O.registration_icp_p2plane(s2, t2, n2, 0.7, np.eye(4), max_iter=30)
    assert np.abs(moved.T - G @ base.T @ np.linalg.inv(G)).max() < 1e-6 and moved.n_corr == base.n_corr
    rng = np.random.default_rng(8); ps = rng.permutation(len(src)); pt = rng.permutation(len(tgt))
    perm = O.registration_icp_p2plane(src[ps], tgt[pt], nrm[pt], 0.7, np.eye(4), max_iter=30)
    assert np.abs(perm.T - base.T).max() < 1e-9 and perm.iters == base.iters
    with pytest.raises(RuntimeError):
        O.registration_icp_p2plane(src, tgt, nrm, 0.0)                    # [O3D] max_correspondence_distance <= 0


def test_single_plane_is_finite_not_a_parity_case():
    """A literally single-plane cloud leaves 3 DoF unobservable (SURVEY 8d config 1 note): only robustness is required."""
    rng = np.random.default_rng(9)
    tgt = np.c_[rng.uniform(0, 10, (500, 2)), np.zeros(500)]; nrm = np.tile([0, 0, 1.0], (500, 1))
    src = tgt + [0.0, 0.0, 0.05]
    r = O.registration_icp_p2plane(src, tgt, nrm, 1.0, np.eye(4), max_iter=5)
    assert r.n_corr == 500 and np.isfinite(r.fitness)


def test_process_scan_matches_golden():
    g = np.load(os.path.join(GOLD, "scan_preprocess.npz"))
    (mx, mn), (ax, an) = O.process_scan(g["raw"].astype(np.float64), O.cropper("MinMaxRadius", 2.0, 30.0), O.cropper("MinMaxRadius", 2.0, 25.0),
                                        0.1, 20, 3.0, 0.3, 5)
    assert mx.shape == g["merge_xyz"].shape and ax.shape == g["match_xyz"].shape
    d, j = cKDTree(g["merge_xyz"]).query(mx)
    assert d.max() == 0.0 and np.abs((mn * g["merge_nrm"][j]).sum(1)).min() > 1 - 1e-12
    with pytest.raises(RuntimeError):
        O.process_scan(np.array([[100.0, 0, 0]]), O.cropper("MinMaxRadius", 2.0, 30.0), O.cropper("MinMaxRadius", 2.0, 25.0), 0.1, 20, 3.0, 1.0, 0)


def test_space_carving_c_vs_numpy_and_hand_case():
    """C1: orc_carve against the numpy restatement and a hand-built case."""
    # hand case: sensor at the origin looking along +x at a wall point 5 m away; map points on the ray with normals
    # facing the sensor are carved, one with a perpendicular normal stays, one beyond (length - truncation) stays,
    # one outside the cropper stays although a ray crosses its voxel
    scan = np.array([[5.0, 0.02, 0.03], [0.03, 12.0, 0.02]])
    mx = np.array([[1.04, 0.02, 0.03], [2.03, 0.01, 0.02], [4.96, 0.02, 0.03], [3.05, 0.02, 0.01], [0.02, 11.03, 0.03]])
    mn = np.array([[-1.0, 0, 0], [0, 1.0, 0], [-1.0, 0, 0], [-0.9, 0.1, 0], [0, -1.0, 0]])
    crop = O.cropper("MaxRadius", 0.0, 8.0)
    rem = O.carve(mx, mn, scan, np.zeros(3), crop, 0.1, 20.0, 0.1, 0.5)
    assert rem.tolist() == [True, False, False, True, False]
    # random case vs numpy
    rng = np.random.default_rng(5)
    mx = rng.uniform(-4, 4, (3000, 3)); mn = rng.normal(size=(3000, 3)); mn[:20] = 0.0
    scan = rng.normal(size=(400, 3)); scan = scan / np.linalg.norm(scan, axis=1)[:, None] * rng.uniform(2, 9, (400, 1)) + [0.3, -0.2, 0.1]
    sensor = np.array([0.3, -0.2, 0.1])
    crop = O.cropper("MaxRadius", 0.0, 3.5, center=(0.5, 0.5, 0.0))
    inside = NP.within("MaxRadius", mx, (0.5, 0.5, 0.0), rmax=3.5)
    for voxel, trunc, mind in ((0.1, 0.1, 0.5), (0.3, 0.6, 0.2)):
        rem = O.carve(mx, mn, scan, sensor, crop, voxel, 6.0, trunc, mind)
        ref = NP.carve(mx, mn, inside, scan, sensor, voxel, 6.0, trunc, mind)
        assert rem.sum() > 10 and np.array_equal(rem, ref)
        assert not rem[~inside].any() and not rem[:20].any()


def test_point_to_point_icp_c_vs_numpy_and_closed_form():
    """R1' (SURVEY 8f rank 3): orc_svd3 against LAPACK, umeyama ICP against the numpy restatement and the exact answer."""
    rng = np.random.default_rng(11)
    for k in range(50):
        A = rng.normal(size=(3, 3)) * 10.0 ** rng.uniform(-3, 3)
        if k % 5 == 0:
            A[:, 2] = A[:, 0] * 0.5 - A[:, 1]          # rank 2
        if k % 10 == 0:
            A = np.outer(A[:, 0], A[0])                  # rank 1
        U, S, V = O.svd3(A)
        assert np.abs(U @ np.diag(S) @ V.T - A).max() < 1e-12 * max(1.0, np.abs(A).max())
        assert np.abs(U.T @ U - np.eye(3)).max() < 1e-12 and np.abs(V.T @ V - np.eye(3)).max() < 1e-12
        assert S[0] >= S[1] >= S[2] >= 0 and np.abs(S - np.linalg.svd(A, compute_uv=False)).max() < 1e-12 * max(1.0, S[0])
    # exact correspondences within r: one umeyama step recovers the rigid motion exactly
    tgt = rng.uniform(-5, 5, (400, 3))
    T0 = synth.se3(0.01, -0.02, 0.015, (0.02, -0.01, 0.03))
    src = (tgt - T0[:3, 3]) @ T0[:3, :3]                 # T0 * src = tgt
    r = O.registration_icp_p2point(src, tgt, 0.5, np.eye(4), max_iter=30)
    assert r.fitness == 1.0 and np.abs(r.T - T0).max() < 1e-12 and r.inlier_rmse < 1e-12
    # config 1 (noisy): C vs numpy
    src, tgt, nrm, T_true = synth.planar_cloud_config1(noise=0.01)
    for init in (np.eye(4), synth.se3(0.01, 0.0, -0.01, (0.02, 0.0, 0.01))):
        rc = O.registration_icp_p2point(src, tgt, 1.0, init, max_iter=50)
        T2, f2, e2, n2, i2 = NP.icp_p2point(src, tgt, 1.0, init, max_iter=50)
        assert rc.iters == i2 and rc.n_corr == n2
        assert np.abs(rc.T - T2).max() < 1e-10 and abs(rc.inlier_rmse - e2) < 1e-12


def test_overlap_and_information_matrix_vs_numpy():
    """L1: computeIndicesOfOverlappingPoints and [O3D] GetInformationMatrixFromPointClouds."""
    rng = np.random.default_rng(21)
    tgt = rng.uniform(-3, 3, (4000, 3)); src = rng.uniform(-1, 5, (3000, 3))
    T = synth.se3(0.02, -0.01, 0.3, (0.4, -0.2, 0.1))
    for voxel, m in ((0.5, 1), (0.8, 3)):
        fs, ft = O.overlap_flags(src, tgt, T, voxel, m)
        st = src @ T[:3, :3].T + T[:3, 3]
        ks = np.floor(st * (1.0 / voxel)).astype(np.int64); kt = np.floor(tgt * (1.0 / voxel)).astype(np.int64)
        from collections import Counter
        cs = Counter(map(tuple, ks)); ct = Counter(map(tuple, kt))
        rs = np.array([cs[tuple(k)] >= m and ct.get(tuple(k), 0) >= m for k in ks]); rt = np.array([ct[tuple(k)] >= m and cs.get(tuple(k), 0) >= m for k in kt])
        assert np.array_equal(fs, rs) and np.array_equal(ft, rt) and 0 < fs.sum() < len(src)
    G = O.information_matrix(src, tgt, 0.3, T)
    st = src @ T[:3, :3].T + T[:3, 3]
    d, j = cKDTree(tgt).query(st)
    ok = d < 0.3
    ref = np.zeros((6, 6))
    for x, y, z in tgt[j[ok]]:
        for r in (np.array([0, z, -y, 1, 0, 0.0]), np.array([-z, 0, x, 0, 1, 0.0]), np.array([y, -x, 0, 0, 0, 1.0])):
            ref += np.outer(r, r)
    assert ok.sum() > 50 and np.abs(G - ref).max() < 1e-9 * np.abs(ref).max()
    assert np.array_equal(G, G.T) and G[3, 3] == G[4, 4] == G[5, 5] == float(ok.sum())


def test_constant_velocity_deskew_vs_numpy():
    """D1: undistortInputPointCloud restated vs scipy Rotation (Rz Ry Rx) and the phase convention."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(4)
    pts = rng.uniform(-20, 20, (500, 3)); pts[0] = [3.0, 0.0, 1.0]; pts[1] = [-2.0, 0.0, 0.5]; pts[2] = [0.0, 0.0, 1.0]
    lv = np.array([5.0, -0.4, 0.1]); av = np.array([0.02, -0.05, 0.8])
    for cw in (True, False):
        out = O.undistort(pts, lv, av, 0.1, cw)
        ang = np.arctan2(pts[:, 1], pts[:, 0]); ang = np.where(ang < 0, ang + 2 * np.pi, ang)
        phase = np.where(ang == 0.0, 0.0, 1.0 - ang / (2 * np.pi) if cw else ang / (2 * np.pi))
        ref = np.empty_like(pts)
        for i, (p, ph) in enumerate(zip(pts, phase)):
            R = Rotation.from_euler("ZYX", (ph * 0.1 * av)[::-1]).as_matrix()      # Rz(yaw) Ry(pitch) Rx(roll)
            ref[i] = R @ p + ph * 0.1 * lv
        assert np.abs(out - ref).max() < 1e-12
        assert np.array_equal(out[0], pts[0])                   # angle 0 -> phase 0 -> untouched
    assert np.array_equal(O.undistort(pts, np.zeros(3), np.zeros(3)), pts)   # zero velocities: identity motion


def test_generalized_icp_c_vs_numpy():
    """R1'' (SURVEY 8f rank 3): orc_registration_gicp (J^T J = A^T M^-1 A) against the numpy restatement that forms
    W = (Ct + Cs)^-1/2 explicitly like [O3D] does, and the covariance-from-normal construction."""
    rng = np.random.default_rng(8)
    for k in range(20):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        if k == 0:
            n = np.array([-1.0, 0.0, 0.0])                      # the c < -0.99 branch: Rx = I
        if k == 1:
            n = np.array([-0.995, 0.0998749, 0.0])
        Cm = O.gicp_covariance_from_normal(n)
        assert np.abs(Cm - NP.gicp_covariances_from_normals(n[None])[0]).max() < 1e-15
        if n[0] >= -0.99:                                        # proper rotation: I - (1 - eps) n n^T
            assert np.abs(Cm - (np.eye(3) - (1 - 1e-3) * np.outer(n, n))).max() < 1e-12
    src, tgt, nrm, T_true = synth.planar_cloud_config1(n=600, noise=0.01)
    snrm = O.estimate_normals(src, 10, 2.0)
    for init in (np.eye(4), synth.se3(0.01, 0.0, -0.01, (0.02, 0.0, 0.01))):
        rc = O.registration_gicp(src, snrm, tgt, nrm, 1.0, init, max_iter=30)
        T2, f2, e2, n2, i2 = NP.icp_gicp(src, snrm, tgt, nrm, 1.0, init, max_iter=30)
        assert rc.iters == i2 and rc.n_corr == n2
        assert np.abs(rc.T - T2).max() < 1e-9 and abs(rc.inlier_rmse - e2) < 1e-10
        assert np.linalg.norm(rc.T[:3, 3] - np.linalg.inv(T_true)[:3, 3]) < 0.02


def test_dense_map_carving_vs_python_restatement():
    """C2: orc_dense_carve against a literal Python transcription of getKeysOfCarvedPoints / getVoxelsWithinPointNeighborhood."""
    import math
    rng = np.random.default_rng(13)
    voxel, radius, trunc, maxlen = 0.05, 0.1, 0.15, 3.0
    pts = rng.uniform(-1.5, 1.5, (6000, 3))
    dm = O.DenseMap(voxel, 1 << 16); dm.insert(pts)
    sensor = np.array([0.1, -0.05, 0.02])
    scan = rng.normal(size=(40, 3)); scan = scan / np.linalg.norm(scan, axis=1)[:, None] * rng.uniform(0.5, 2.5, (40, 1)) + sensor
    scan = np.vstack([scan, scan[:5] + 1e-4])                  # same voxel as an earlier point: dropped by the de-duplication
    present = {tuple(k) for k in dm.to_cloud()[2]}
    inv = 1.0 / voxel
    seen, rays = set(), []
    for p in scan:
        k = tuple(math.floor(v * inv) for v in p)
        if k not in seen:
            seen.add(k); rays.append(p)
    remove = set()
    step = 2.0 * radius
    for p in rays:
        d = p - sensor; length = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]); u = d / length
        mp = max(step, min(length - trunc, maxlen)); dist = 0.0
        while dist < mp:
            c = dist * u + sensor
            ck = tuple(math.floor(v / voxel) for v in c); added = False
            ox = -radius
            while ox <= radius:
                oy = -radius
                while oy <= radius:
                    oz = -radius
                    while oz <= radius:
                        t = c + np.array([ox, oy, oz]); k = tuple(math.floor(v / voxel) for v in t)
                        e = t - (np.array(k, dtype=np.float64) * voxel + voxel * 0.5)
                        if math.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) <= radius:
                            if k in present:
                                remove.add(k)
                            added = added or k == ck
                        oz += voxel
                    oy += voxel
                ox += voxel
            if not added and ck in present:
                remove.add(ck)
            dist += step
    n = dm.carve(scan, sensor, voxel, radius, trunc, maxlen)
    assert n == len(remove) and 0 < n < len(present)
    assert {tuple(k) for k in dm.to_cloud()[2]} == present - remove


def test_next_rows_match_golden():
    """The committed outputs of the "next" rows (tests/golden/next_rows.npz, written by make_golden.py) still come out."""
    g = np.load(os.path.join(GOLD, "next_rows.npz"))
    src, tgt, nrm, _ = synth.planar_cloud_config1(n=800, noise=0.01)
    snrm = O.estimate_normals(src, 10, 2.0)
    init = synth.se3(0.01, -0.02, 0.03, (0.05, 0.02, -0.01))
    r = O.registration_icp_p2point(src, tgt, 1.0, init, max_iter=50)
    assert r.iters == int(g["p2p_iters"]) and r.n_corr == int(g["p2p_ncorr"]) and np.abs(r.T - g["p2p_T"]).max() < 1e-12
    r = O.registration_gicp(src, snrm, tgt, nrm, 1.0, init, max_iter=30)
    assert r.iters == int(g["gicp_iters"]) and r.n_corr == int(g["gicp_ncorr"]) and np.abs(r.T - g["gicp_T"]).max() < 1e-12
    fs, ft = O.overlap_flags(src, tgt, init, 0.5, 2)
    assert np.array_equal(np.packbits(fs), g["overlap_src"]) and np.array_equal(np.packbits(ft), g["overlap_tgt"])
    assert np.abs(O.information_matrix(src, tgt, 0.3, init) - g["info"]).max() < 1e-9
    rng = np.random.default_rng(77)
    raw = rng.normal(size=(300, 3)); raw = raw / np.linalg.norm(raw, axis=1)[:, None] * rng.uniform(3, 9, (300, 1))
    Ts = synth.se3(t=(5.0, 5.0, 1.0))
    rem = O.carve(tgt, nrm, O.transform(Ts, raw)[0], Ts[:3, 3], O.cropper("MaxRadius", 0.0, 8.0, center=(5.0, 5.0, 0.0)), 0.25, 20.0, 0.1, 0.3)
    assert np.array_equal(np.packbits(rem), g["carved"]) and 0 < rem.sum() < len(tgt)
    assert np.abs(O.undistort(src[:50], np.array([5.0, -0.4, 0.1]), np.array([0.02, -0.05, 0.8]), 0.1, True) - g["deskew"]).max() < 1e-13
