"""-m "not gpu": the host-side control flow of config 5 (open3d_slam_b200/slam.py) and the synthetic workloads, on the CPU.

slam.py holds the reference's Mapper / SubmapCollection decisions once, over a backend interface; here it runs over the oracle
backend only (tests/oracle_backend.py -- test infrastructure), which checks the control flow itself: first-scan insertion, the
fitness gate, submap hand-over with the buffered overlap scans, event log, ground truth.  The device backend runs the same code in
tests/test_gpu_configs.py and must produce the same events.
"""
import copy

import numpy as np

from open3d_slam_b200 import engine as E
from open3d_slam_b200 import slam as S
from open3d_slam_b200 import workloads as W
from open3d_slam_b200 import dist as D
from oracle_backend import OracleBackend


def test_closed_loop_workload_is_an_exact_lap():
    lp = W.ClosedLoop()
    assert lp.L >= 100 and abs(lp.step - 0.5) < 0.02
    for k in (0, 1, 17, lp.L - 1):
        assert np.array_equal(lp.pose(k), lp.pose(k + lp.L)) and np.array_equal(lp.pose(k), lp.pose(k + 3 * lp.L))
        assert np.array_equal(lp.delta(k + lp.L), lp.delta(k + 2 * lp.L))          # periodic for k >= 1
    assert np.array_equal(lp.delta(0), np.eye(4))
    # the odometry deltas carry a small error, but chaining the TRUE motion closes the loop exactly
    T = np.eye(4)
    for k in range(1, lp.L + 1):
        T = T @ (np.linalg.inv(lp.pose(k - 1)) @ lp.pose(k))
    assert np.abs(T - np.eye(4)).max() < 1e-9
    a, b = lp.scan(5, seed=1), lp.scan(5 + lp.L, seed=1)
    assert a.dtype == np.float32 and np.array_equal(a, b) and not np.array_equal(a, lp.scan(5, seed=2))
    assert np.abs(lp.map_frame_pose(0) - np.eye(4)).max() < 1e-12


def test_config4_layout_pairs_over_shared_targets():
    lp = W.ClosedLoop()
    c4 = W.Config4(lp, 512, 64)
    owners = [c4.target_of(i) for i in range(512)]
    assert sorted(set(owners)) == list(range(64)) and all(owners.count(t) == 8 for t in range(64))      # 8 pairs per target
    for t in (0, 13, 63):
        pos = c4.target_positions(t)
        assert len(pos) == 6 and all(b - a == 4 for a, b in zip(pos, pos[1:])) and pos[0] <= c4.center(t) <= pos[-1]
    for i in (0, 100, 511):
        assert abs(c4.source_position(i) - c4.center(c4.target_of(i))) <= 4
        d = np.linalg.inv(c4.truth(i)) @ c4.init(i)                                   # initial guess: within 0.5 m / 5 deg per axis of the truth
        assert np.abs(d[:3, 3]).max() <= 0.5 and np.degrees(np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1))) < 9.0
        assert np.array_equal(c4.init(i), c4.init(i))
    # sharding + ownership used by the multi-GPU run: every pair on exactly one rank, every target owned by exactly one rank
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += list(D.shard_range(512, world, r))
        assert seen == list(range(512))
        assert all(0 <= D.owner_of(t, world) < world for t in range(64))
        assert len({D.owner_of(t, world) for t in range(64)}) == min(world, 64)


def test_segment_mapper_control_flow_on_the_oracle_backend():
    """12 scans with a 2 m submap radius: the chain must follow the trajectory, hand the active submap over at least once and log it."""
    p = E.MapperParameters(seed=3)
    lp = W.ClosedLoop()
    ora = OracleBackend(copy.deepcopy(p), carving=True, dense=True)
    m = S.SegmentMapper(ora, S.SubmapParameters(radius=2.0))
    n_reg = 0
    for k in range(12):
        r = m.addRangeMeasurement(lp.scan(k, seed=k), lp.delta(k))
        if k == 0:
            assert r is None                                                    # Mapper.cpp:105-114: the first scan is only inserted
            continue
        n_reg += 1
        assert r.fitness > 0.9 and r.iters >= 1
        assert np.linalg.norm(m.mapToRangeSensor[:3, 3] - lp.map_frame_pose(k)[:3, 3]) < 0.1
    assert n_reg == 11
    ev = m.submaps.events
    assert any(e[0] == "active_submap_changed" for e in ev) and len(m.submaps.submaps) >= 2
    assert m.submaps.activeSubmapIdx == len(m.submaps.submaps) - 1 or m.submaps.activeSubmapIdx in range(len(m.submaps.submaps))
    assert len(m.submaps.finishedSubmapsIdxs) >= 1
    total = 0
    for s in m.submaps.submaps:
        c = ora.counters(s.handle)
        x, n = ora.map_cloud(s.handle)
        assert c["inserted_map"] >= 1 and len(x) > 1000 and len(x) == len(n)
        total += c["inserted_map"]
    assert total >= 12                                                          # every scan went into a map; overlap scans into two
