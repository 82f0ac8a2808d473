"""-m gpu: the nearest-neighbour index + correspondence search tested DIRECTLY against brute force,
on inputs built to break a grid index: queries far outside the indexed box, exact distance ties (within a cell and across cell
borders), a search radius of ten cell edges, NaN points on both sides, duplicates, a single-point and an empty target.
Expected answers come from numpy with the oracle's own distance expression ((dx*dx + dy*dy) + dz*dz, strict d2 < r2, ties ->
lower index); indices must be IDENTICAL, squared distances bit-identical."""
import numpy as np
import pytest

from open3d_slam_b200 import engine as E
from open3d_slam_b200 import synth

pytestmark = pytest.mark.gpu


def brute(q, t, r):
    idx = np.full(len(q), -1, dtype=np.int64); d2o = np.full(len(q), -1.0)
    ok_t = ~np.isnan(t).any(axis=1)
    for i, p in enumerate(q):
        if np.isnan(p).any() or not ok_t.any():
            continue
        dx, dy, dz = p[0] - t[:, 0], p[1] - t[:, 1], p[2] - t[:, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        d2 = np.where(ok_t, d2, np.inf)
        j = int(np.argmin(d2))          # first minimum = lowest index among ties
        if d2[j] < r * r:
            idx[i] = j; d2o[i] = d2[j]
    return idx, d2o


def check(eng, q, t, r, T=None):
    gi, gd = E.nearestNeighbors(eng, eng.cloud(q), eng.cloud(t), r, T)
    if T is not None:
        x = ((T[0, 0] * q[:, 0] + T[0, 1] * q[:, 1]) + T[0, 2] * q[:, 2]) + T[0, 3]
        y = ((T[1, 0] * q[:, 0] + T[1, 1] * q[:, 1]) + T[1, 2] * q[:, 2]) + T[1, 3]
        z = ((T[2, 0] * q[:, 0] + T[2, 1] * q[:, 1]) + T[2, 2] * q[:, 2]) + T[2, 3]
        q = np.c_[x, y, z]
    ri, rd = brute(q, t, r)
    assert np.array_equal(gi, ri), np.flatnonzero(gi != ri)[:10]
    assert np.array_equal(gd, rd)
    return gi


def test_nn_random_and_far_outside_the_box(engine_factory):
    eng = engine_factory(E.MapperParameters())
    rng = np.random.default_rng(0)
    t = rng.uniform(-5, 5, (6000, 3))
    q = np.vstack([rng.uniform(-5, 5, (1500, 3)),                 # inside
                   rng.uniform(-5, 5, (300, 3)) + [40.0, 0, 0],   # far outside the indexed box: nothing within r
                   rng.uniform(-5.4, 5.4, (600, 3)),              # straddling the faces of the box
                   np.array([[5.3, 5.3, 5.3], [-5.2, 0.0, 9.0], [1e6, -1e6, 3.0]])])
    for r in (0.3, 1.0):
        gi = check(eng, q, t, r)
        assert (gi >= 0).sum() > 500 and (gi < 0).sum() >= 300


def test_nn_exact_ties_within_and_across_cells(engine_factory):
    """lattice targets, queries at cell centres / face centres / edge midpoints: 2, 4 or 8 targets at EXACTLY the same distance,
    sitting in different grid cells -- the lower index must win whatever the cell order"""
    eng = engine_factory(E.MapperParameters())
    g = np.arange(-8, 9) * 0.25                              # 0.25 = the cell edge for r = 1.0 (max_corr / 4): lattice points ON cell borders
    t = np.array([[x, y, z] for x in g for y in g for z in g[:5]])
    rng = np.random.default_rng(1)
    perm = rng.permutation(len(t)); t = t[perm]             # index order unrelated to position
    q = np.vstack([t[:400] + [0.125, 0.0, 0.0],             # midpoint of an x edge: 2-way tie
                   t[400:800] + [0.125, 0.125, 0.0],        # face centre: 4-way tie
                   t[800:1200] + [0.125, 0.125, 0.125],     # cell centre: 8-way tie
                   t[1200:1300]])                           # exact hits: d2 = 0
    for r in (1.0, 0.3, 0.2):
        check(eng, q, t, r)
    dup = np.vstack([t, t[:500]])                           # duplicated target points: identical distance, the first copy wins
    check(eng, q, dup, 1.0)


def test_nn_radius_of_ten_cells_and_strict_cut(engine_factory):
    p = E.MapperParameters()
    p.nnCellSize = 0.1                                      # r = 10 x cell
    eng = engine_factory(p)
    rng = np.random.default_rng(2)
    t = rng.uniform(-3, 3, (800, 3))                        # sparse: most neighbours are many cells away
    q = rng.uniform(-4, 4, (700, 3))
    check(eng, q, t, 1.0)
    # strict d2 < r2: a target at distance exactly r is NOT a correspondence
    t2 = np.array([[0.0, 0.0, 0.0]]); q2 = np.array([[1.0, 0.0, 0.0], [0.999999, 0.0, 0.0], [0.6, 0.8, 0.0]])
    gi, _ = E.nearestNeighbors(eng, eng.cloud(q2), eng.cloud(t2), 1.0)
    assert list(gi) == [-1, 0, -1]


def test_nn_nan_points_empty_and_single_target_and_transform(engine_factory):
    eng = engine_factory(E.MapperParameters())
    rng = np.random.default_rng(3)
    t = rng.uniform(-2, 2, (3000, 3)); t[::7] = np.nan      # NaN targets are never a neighbour
    q = rng.uniform(-2, 2, (900, 3)); q[::11, 1] = np.nan   # NaN queries have none
    gi = check(eng, q, t, 0.5)
    assert (gi[::11] == -1).all() and not np.isin(gi[gi >= 0] % 7, [0]).any()
    T = synth.se3(0.02, -0.03, 0.4, (0.3, -0.2, 0.1))
    check(eng, q, t, 0.5, T)                                # queries moved by T first (correspondence_set_ at a transformation)
    check(eng, q[:50], t[1:2], 5.0)                         # a single target point
    gi, gd = E.nearestNeighbors(eng, eng.cloud(q[:20]), eng.cloud(np.zeros((0, 3))), 1.0)
    assert (gi == -1).all() and (gd == -1.0).all()          # empty target
    big = rng.uniform(-6, 6, (95_000, 3))                   # more queries than one launch holds on chip: chunked
    tb = rng.uniform(-6, 6, (20_000, 3))
    check(eng, big[::1], tb, 0.4)


def test_registration_correspondence_set_matches_search(engine_factory):
    """correspondence_set_ of a converged registration = the search at the result's transformation; n_corr agrees."""
    src, tgt, nrm, _ = synth.planar_cloud_config1(noise=0.01)
    p = E.MapperParameters(); p.icp.maxCorrespondenceDistance = 1.0
    eng = engine_factory(p)
    reg = E.RegistrationIcpPointToPlane(eng, E.CloudRegistrationParameters(icp=p.icp))
    res = reg.registerClouds(eng.cloud(src), eng.cloud(tgt, nrm), np.eye(4))
    gi, gd = E.nearestNeighbors(eng, eng.cloud(src), eng.cloud(tgt), 1.0, res.transformation_)
    assert (gi >= 0).sum() == res.n_corr
    assert abs(np.sqrt(gd[gi >= 0].mean()) - res.inlier_rmse_) < 1e-9
