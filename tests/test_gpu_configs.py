"""-m gpu parity tests of BASELINE.json's configs 4 and 5 at the settings bench.py runs them with, of the device-side
gates of the mapper chain, and of the F4 VoxelMap container -- the CUDA path through the C ABI against the CPU oracle.

Tolerances (written where they are used): discrete outcomes (iteration / correspondence counts, carved sets, hand-over
decisions, voxel keys, overlap sets) must be IDENTICAL; transforms 1e-7 relative (north star: 1e-4) over 200 chained scans.
"""
import copy

import numpy as np
import pytest

from oracle import oracle as O
from open3d_slam_b200 import engine as E
from open3d_slam_b200 import dist as D
from open3d_slam_b200 import slam as S
from open3d_slam_b200 import synth
from open3d_slam_b200 import workloads as W

from oracle_backend import OracleBackend

pytestmark = pytest.mark.gpu


def rel_rot(Ta, Tb):
    return np.linalg.norm(Ta[:3, :3] - Tb[:3, :3]) / np.linalg.norm(Tb[:3, :3])


def rel_trans(Ta, Tb):
    return np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]) / max(np.linalg.norm(Tb[:3, 3]), 1.0)


def canon(xyz, voxel=0.1):
    """order of a voxelised cloud by voxel key: robust to the ~1e-9 differences chained poses carry"""
    k = np.floor(xyz / voxel).astype(np.int64)
    return xyz[np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0], k[:, 2], k[:, 1], k[:, 0]))]


# ----------------------------------------------------------------------------------------------------------------------
# F4: VoxelMap container
# ----------------------------------------------------------------------------------------------------------------------
def test_voxel_map_container_matches_reference_semantics(engine_factory):
    """o3d_slam::VoxelMap (Voxel.cpp:123-160): insertCloud / getIndicesInVoxel / hasVoxelContainingPoint / size / clear against a
    Python dict keyed by floor(p * (1 / voxel)), two layers, negative coordinates, points on voxel faces, the revisit check."""
    eng = engine_factory(E.MapperParameters())
    rng = np.random.default_rng(4)
    v = 0.25
    a = np.vstack([rng.uniform(-6, 6, (4000, 3)), np.array([[0.0, 0.0, 0.0], [0.25, -0.25, 0.5], [-0.25, 0.75, -1.0]])])   # incl. face points
    b = rng.uniform(-3, 9, (1500, 3))
    vm = E.VoxelMap(eng, v, 1 << 14)
    vm.insertCloud("map", eng.cloud(a))
    vm.insertCloud("scan", eng.cloud(b))
    inv = 1.0 / v
    ref = {"map": {}, "scan": {}}
    for name, pts in (("map", a), ("scan", b)):
        for i, k in enumerate(map(tuple, np.floor(pts * inv).astype(np.int64))):
            ref[name].setdefault(k, []).append(i)
    keys_all = set(ref["map"]) | set(ref["scan"])
    assert vm.size() == len(keys_all)
    q = np.vstack([a[::5] + rng.uniform(-0.01, 0.01, (len(a[::5]), 3)), rng.uniform(-12, 12, (800, 3))])
    qk = [tuple(k) for k in np.floor(q * inv).astype(np.int64)]
    qc = eng.cloud(q)
    for layer in ("map", "scan"):
        got = vm.getIndicesInVoxel(layer, qc)
        for g, k in zip(got, qk):
            assert list(g) == ref[layer].get(k, [])          # ascending = insertion order of insertCloud(layer, cloud)
    assert vm.getIndicesInVoxel("nolayer", qc)[0].size == 0
    flags, hits = vm.hasVoxelContainingPoint(qc)
    exp = np.array([k in keys_all for k in qk])
    assert np.array_equal(flags, exp) and hits == int(exp.sum()) and 0 < hits < len(q)
    T = synth.se3(0.02, -0.01, 0.3, (0.4, -0.2, 0.1))      # isSwitchingSubmapsConsistant: p = mapToRangeSensor * scan point
    x = ((T[0, 0] * q[:, 0] + T[0, 1] * q[:, 1]) + T[0, 2] * q[:, 2]) + T[0, 3]
    y = ((T[1, 0] * q[:, 0] + T[1, 1] * q[:, 1]) + T[1, 2] * q[:, 2]) + T[1, 3]
    z = ((T[2, 0] * q[:, 0] + T[2, 1] * q[:, 1]) + T[2, 2] * q[:, 2]) + T[2, 3]
    expT = np.array([tuple(k) in keys_all for k in np.floor(np.c_[x, y, z] * inv).astype(np.int64)])
    flagsT, hitsT = vm.hasVoxelContainingPoint(qc, T)
    assert np.array_equal(flagsT, expT) and hitsT == int(expT.sum())
    vm.clear()
    assert vm.size() == 0 and vm.hasVoxelContainingPoint(qc)[1] == 0
    vm.free()


# ----------------------------------------------------------------------------------------------------------------------
# the device-side gates of the mapper chain
# ----------------------------------------------------------------------------------------------------------------------
def _oracle_single_submap_loop(p, scans, deltas, min_move=0.0, carving=False, dense=False):
    be = OracleBackend(p, carving=carving, dense=dense)
    sm = be.new_submap()
    be.first_scan(sm, scans[0])
    last_ins = np.eye(4)
    out = []
    for k in range(1, len(scans)):
        if min_move > 0.0:   # Mapper.cpp:151-176 with the minimum-motion gate: registration always, fusion only after enough motion
            raw64 = scans[k].astype(np.float64)
            (mx, mn), (ax, an) = be._process(scans[k])
            px, pn = O.crop(be._cropper(p.scanProcessing.cropper, center=be.pose[:3, 3]), sm.xyz, sm.nrm)
            r = O.registration_icp_p2plane(ax, px, pn, p.icp.maxCorrespondenceDistance, be.pose @ deltas[k], max_iter=p.icp.maxNumIter)
            if not (r.fitness < p.minRefinementFitness):
                be.pose = np.array(r.T)
                motion = np.linalg.inv(last_ins) @ be.pose
                if not (np.linalg.norm(motion[:3, 3]) < min_move):
                    if carving:
                        be._carve(sm, raw64, be.pose)
                    be._insert(sm, mx, mn, be.pose)
                    last_ins = be.pose.copy()
            out.append(r)
        else:
            r, _ = be.step(sm, scans[k], deltas[k])
            out.append(r)
    return be, sm, out


@pytest.mark.parametrize("graph", [False, True])
def test_chain_minimum_motion_gate_and_carving_schedule(engine_factory, graph):
    """Mapper.cpp:170-176 and Submap.cpp:111 decided on the device: with minMovementBetweenMappingSteps = 0.8 m and 0.5 m between
    scans only every other scan is fused; carving (every 3rd insertion here) runs on the device's own insertion counter.  Same
    registrations, same map, same counters as the oracle loop; eager and graph-replayed chains agree."""
    p = E.MapperParameters(seed=3)
    p.minMovementBetweenMappingSteps = 0.8
    p.mapBuilder.carving.carveSpaceEveryNscans = 3
    eng = engine_factory(p)
    lp = W.ClosedLoop()
    n = 14
    scans = [lp.scan(k, seed=k) for k in range(n)]
    deltas = [lp.delta(k) for k in range(n)]
    mapper = E.Mapper(eng, 700_000)
    mapper.submap.setMapperOptions(minMovement=0.8, carving=p.mapBuilder.carving)
    mapper.addRangeMeasurement(eng.cloud(scans[0]), None)
    mapper.submap.setPose(np.eye(4))
    staging = mapper.enableGraph(65536) if graph else None
    got = []
    for k in range(1, n):
        c = eng.cloud(scans[k])
        if graph:
            mapper.stageCopy(c)
            slot = mapper.addRangeMeasurementAsync(staging, deltas[k])
        else:
            slot = mapper.addRangeMeasurementAsync(c, deltas[k], slot=k)
        got.append(mapper.fetchResult(slot))
    be, sm, ref = _oracle_single_submap_loop(p, scans, deltas, min_move=0.8, carving=True)
    for g, r in zip(got, ref):
        assert g.iters == r.iters and g.n_corr == r.n_corr
        assert np.abs(g.transformation_ - r.T).max() < 1e-8
    cnt = mapper.submap.mapperCounters()
    assert cnt["steps"] == n - 1 and cnt["accepted"] == n - 1
    assert cnt["inserted_map"] == sm.nScansInsertedMap and 1 + (n - 1) // 2 - 1 <= cnt["inserted_map"] <= 1 + (n - 1) // 2 + 1
    assert cnt["carve_runs"] == sm.carve_runs and cnt["carve_runs"] >= 1
    assert cnt["carved_points_total"] == sm.carved_total
    gx, gn = mapper.submap.getMapPointCloud()
    assert len(gx) == len(sm.xyz)
    assert np.abs(canon(gx) - canon(sm.xyz)).max() < 1e-8    # same voxels, same members (the chained poses agree to 1e-9)
    mapper.submap.free()


# ----------------------------------------------------------------------------------------------------------------------
# LidarOdometry: the scan-to-scan chain (config 1's real call site)
# ----------------------------------------------------------------------------------------------------------------------
def test_lidar_odometry_scan_to_scan_chain(engine_factory):
    """src/Odometry.cpp:25-79: preprocess (crop -> voxelize -> normals -> RandomDownSample), registerClouds(previous, current, I),
    cumulative *= result^-1 -- against the same chain over the oracle, 8 scans of the loop."""
    op = E.OdometryParameters(seed=5)
    op.scanMatcher.icp = E.IcpParameters(maxNumIter=50, maxCorrespondenceDistance=1.0, knn=20, maxDistanceKnn=3.0)
    p = E.MapperParameters(seed=5); p.icp = op.scanMatcher.icp
    eng = engine_factory(p)
    odo = E.LidarOdometry(eng, op)
    lp = W.ClosedLoop()
    crop = O.cropper("MinMaxRadius", 2.0, 30.0)
    prev = None; cum = np.eye(4)
    for k in range(8):
        raw = lp.scan(k, seed=40 + k)
        ok = odo.addRangeScan(eng.cloud(raw))
        cx, _ = O.crop(crop, raw.astype(np.float64))
        vx, _ = O.voxel_down_sample(cx, 0.1)
        vn = O.estimate_normals(vx, 20, 3.0)
        sx, sn = O.random_down_sample(vx, 0.3, 5, vn)
        if prev is not None:
            ref = O.registration_icp_p2plane(prev[0], sx, sn, 1.0, np.eye(4), max_iter=50)
            g = odo.lastResult
            assert ok and g.iters == ref.iters and g.n_corr == ref.n_corr
            assert rel_rot(g.transformation_, ref.T) < 1e-8 and rel_trans(g.transformation_, ref.T) < 1e-8     # north star: 1e-4
            cum = cum @ np.linalg.inv(ref.T)
        prev = (sx, sn)
    assert np.abs(odo.getOdomToRangeSensor() - cum).max() < 1e-7
    gt = np.linalg.inv(lp.pose(0)) @ lp.pose(7)
    assert np.linalg.norm(odo.getOdomToRangeSensor()[:3, 3] - gt[:3, 3]) < 0.1      # and it is odometry: 3.5 m travelled, < 10 cm off


# ----------------------------------------------------------------------------------------------------------------------
# config 5: the full mapper over a segment of the trajectory
# ----------------------------------------------------------------------------------------------------------------------
def test_config5_segment_full_mapper(engine_factory):
    """>= 200 scans of the closed lap through the full mapper: device chain replayed as a CUDA graph, float32 host scans in /
    RegistrationResult out in one C call (bench.py's path), ratio 0.3, carving every 10 insertions, dense map with carving,
    submap hand-overs (radius 10 m so that the 16 m loop needs them), buffered overlap scans, revisit check through the
    device VoxelMap, then the loop-closure refinement (overlap -> batched ICP -> information matrix) between the finished
    submaps -- against the same control flow over the CPU oracle."""
    N = 208
    p = E.MapperParameters(seed=3)
    lp = W.ClosedLoop()
    sp = S.SubmapParameters(radius=10.0)
    dev = S.DeviceBackend(copy.deepcopy(p), carving=True, dense=True, graph=True)
    ora = OracleBackend(copy.deepcopy(p), carving=True, dense=True)
    md, mo = S.SegmentMapper(dev, sp), S.SegmentMapper(ora, sp)
    worst_T, worst_fit = 0.0, 0.0
    for k in range(N):
        raw = lp.scan(k, seed=k)
        d = lp.delta(k)
        rd = md.addRangeMeasurement(raw, d)
        ro = mo.addRangeMeasurement(raw, d)
        if rd is None:
            assert ro is None
            continue
        assert rd.iters == ro.iters and rd.n_corr == ro.n_corr, (k, rd, ro)
        worst_fit = max(worst_fit, abs(rd.fitness_ - ro.fitness))
        worst_T = max(worst_T, rel_rot(rd.transformation_, ro.T), rel_trans(rd.transformation_, ro.T))
        assert md.submaps.activeSubmapIdx == mo.submaps.activeSubmapIdx, k
    assert worst_T < 1e-7 and worst_fit < 1e-10                                   # north star: 1e-4
    # identical host decisions: hand-overs and revisit checks (fitness of the check to 1e-12)
    ed, eo = md.submaps.events, mo.submaps.events
    assert [e[:1] + e[1:2] for e in ed] == [e[:1] + e[1:2] for e in eo] and len(ed) == len(eo)
    for a, b in zip(ed, eo):
        if a[0] == "revisit_check":
            assert abs(a[2] - b[2]) < 1e-12
        else:
            assert a == b
    assert len(md.submaps.submaps) == len(mo.submaps.submaps) >= 2
    assert any(e[0] == "active_submap_changed" for e in ed)
    # ground truth sanity: the chain follows the trajectory
    assert np.linalg.norm(md.mapToRangeSensor[:3, 3] - lp.map_frame_pose(N - 1)[:3, 3]) < 0.15
    carved = 0
    for sd, so in zip(md.submaps.submaps, mo.submaps.submaps):
        cd, co = dev.counters(sd.handle), ora.counters(so.handle)
        for key in ("inserted_map", "inserted_dense", "carve_runs", "carved_points_total", "dense_carve_runs", "carved_voxels_total"):
            assert cd[key] == co[key], (key, cd, co)
        carved += cd["carve_runs"]
        gx, gn = dev.map_cloud(sd.handle); rx, rn = ora.map_cloud(so.handle)
        assert len(gx) == len(rx)
        kg, kr = np.floor(gx / 0.1).astype(np.int64), np.floor(rx / 0.1).astype(np.int64)
        o1 = np.lexsort((gx[:, 2], gx[:, 1], gx[:, 0], kg[:, 2], kg[:, 1], kg[:, 0])); o2 = np.lexsort((rx[:, 2], rx[:, 1], rx[:, 0], kr[:, 2], kr[:, 1], kr[:, 0]))
        assert np.abs(gx[o1] - rx[o2]).max() < 1e-7          # (poses differ by ~1e-9 after 200 chained registrations)
        assert np.abs(gn[o1] - rn[o2]).max() < 1e-5
        dx, dk = dev.dense_map(sd.handle); ox, ok = ora.dense_map(so.handle)
        assert len(dx) == len(ox)
        q1 = np.lexsort((dk[:, 2], dk[:, 1], dk[:, 0])); q2 = np.lexsort((ok[:, 2], ok[:, 1], ok[:, 0]))
        assert np.array_equal(dk[q1], ok[q2]) and np.abs(dx[q1] - ox[q2]).max() < 1e-8
    assert carved >= 10
    # loop-closure refinement between the finished submap(s) and the last active one
    fin = md.submaps.finishedSubmapsIdxs
    assert fin == mo.submaps.finishedSubmapsIdxs and len(fin) >= 1
    src_i, tgt_i = fin[0], md.submaps.activeSubmapIdx if md.submaps.activeSubmapIdx != fin[0] else (fin[0] + 1) % len(md.submaps.submaps)
    inits = [synth.se3(0.01, -0.008, 0.015, (0.12, -0.08, 0.03)), np.eye(4)]
    lcp = S.LoopClosureParameters()
    sdc, tdc = dev.submap_as_cloud(md.submaps.submaps[src_i].handle), dev.submap_as_cloud(md.submaps.submaps[tgt_i].handle)
    soc, toc = ora.submap_as_cloud(mo.submaps.submaps[src_i].handle), ora.submap_as_cloud(mo.submaps.submaps[tgt_i].handle)
    gd = S.refineLoopClosures(dev, sdc, [tdc, tdc], inits, p.mapBuilder.mapVoxelSize, lcp)
    go = S.refineLoopClosures(ora, soc, [toc, toc], inits, p.mapBuilder.mapVoxelSize, lcp)
    for a, b in zip(gd, go):
        assert a["n_source_overlap"] == b["n_source_overlap"] > 1000 and a["n_target_overlap"] == b["n_target_overlap"] > 1000
        assert a["result"].iters == b["result"].iters and a["result"].n_corr == b["result"].n_corr
        assert np.abs(a["result"].transformation_ - b["result"].T).max() < 1e-7
        assert a["accepted"] == b["accepted"] and a["accepted"]
        assert np.abs(a["information"] - b["information"]).max() / np.abs(b["information"]).max() < 1e-8
    dev.close()


# ----------------------------------------------------------------------------------------------------------------------
# config 4: 512 pairs over shared 20 m-radius targets
# ----------------------------------------------------------------------------------------------------------------------
def test_config4_512_pairs_shared_targets(engine_factory):
    """The whole of config 4 on one GPU through dist.shard_range (world 1): 512 scan-submap pairs over 64 shared targets built by the
    engine's own S1 + F1 (>= 100 k points each), r = 0.3, <= 100 iterations, one b2s_register_batch call (every distinct target
    is indexed once).  Parity against the oracle on a seeded sample of 32 pairs; sanity of all 512 against ground truth."""
    p = E.MapperParameters(seed=3)
    eng = engine_factory(p)
    lp = W.ClosedLoop()
    c4 = W.Config4(lp, n_pairs=512, n_targets=64)
    icp = E.ScanToMapIcp(eng)
    mine = D.shard_range(c4.P, 1, 0)
    assert list(mine) == list(range(512))
    targets = {t: c4.build_target(E, eng, icp, p, t) for t in sorted({c4.target_of(i) for i in mine})}
    sizes = [len(c) for c in targets.values()]
    assert len(targets) == 64 and min(sizes) > 100_000
    sources = [c4.build_source(E, eng, icp, i) for i in mine]
    inits = [c4.init(i) for i in mine]
    reg = c4.registration(E, eng, p)
    res = reg.registerCloudsBatch(sources, [targets[c4.target_of(i)] for i in mine], inits)
    err = np.array([np.linalg.norm(r.transformation_[:3, 3] - c4.truth(i)[:3, 3]) for r, i in zip(res, mine)])
    fit = np.array([r.fitness_ for r in res])
    assert np.median(err) < 0.02 and (err < 0.1).mean() > 0.95 and np.median(fit) > 0.9
    sample = np.random.default_rng(2026).choice(512, 32, replace=False)
    host_targets = {}
    for i in sample:
        t = c4.target_of(int(i))
        if t not in host_targets:
            host_targets[t] = targets[t].download()
        sx, _sn = sources[int(i)].download()
        tx, tn = host_targets[t]
        ref = O.registration_icp_p2plane(sx, tx, tn, c4.R_ICP, inits[int(i)], max_iter=c4.MAX_ITER)
        g = res[int(i)]
        assert g.iters == ref.iters and g.n_corr == ref.n_corr, (i, g, ref)
        assert abs(g.fitness_ - ref.fitness) < 1e-12 and abs(g.inlier_rmse_ - ref.inlier_rmse) < 1e-9
        assert rel_rot(g.transformation_, ref.T) < 1e-8 and rel_trans(g.transformation_, ref.T) < 1e-8     # north star: 1e-4
    # a pair registered on its own gives the batch's answer (shared index == private index)
    i0 = int(sample[0])
    single = reg.registerClouds(sources[i0], targets[c4.target_of(i0)], inits[i0])
    assert single.iters == res[i0].iters and np.abs(single.transformation_ - res[i0].transformation_).max() < 1e-11
    for c in list(targets.values()) + sources:
        c.free()
