"""Generates the golden vectors under tests/golden/ with the CPU oracle (oracle/o3d_oracle.c).

PARITY UNPINNED: the reference has no tests / fixtures on this path and Open3D v0.15.1 cannot run here, so these vectors
pin the ORACLE (cross-checked against oracle/np_oracle.py and analytic answers in tests/test_oracle.py), not a run of the
reference binary.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from open3d_slam_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    # config 1: 2 000-pt three-plane cloud, noise 0 and 1 cm
    g = {}
    for tag, noise in (("clean", 0.0), ("noisy", 0.01)):
        src, tgt, nrm, T_true = synth.planar_cloud_config1(noise=noise)
        r = O.registration_icp_p2plane(src, tgt, nrm, 1.0, np.eye(4), max_iter=50, trace=True)
        g[f"{tag}_T"] = r.T; g[f"{tag}_fitness"] = r.fitness; g[f"{tag}_rmse"] = r.inlier_rmse
        g[f"{tag}_ncorr"] = r.n_corr; g[f"{tag}_iters"] = r.iters; g[f"{tag}_trace"] = r.trace
    np.savez_compressed(os.path.join(OUT, "config1_icp.npz"), **g)

    # pre-processing of one synthetic 64x1024 scan (reduced to 16 beams x 512 azimuths to keep the fixture small)
    scene = synth.Scene(); pose = synth.loop_trajectory(4)[1]
    raw = synth.lidar_scan(scene, pose, n_beams=16, n_az=512, seed=11)
    wide = O.cropper("MinMaxRadius", 2.0, 30.0); narrow = O.cropper("MinMaxRadius", 2.0, 25.0)
    (mx, mn), (ax, an) = O.process_scan(raw.astype(np.float64), wide, narrow, 0.1, 20, 3.0, 0.3, 5)
    vx, _, keys = O.voxel_down_sample(raw.astype(np.float64), 0.1, return_keys=True)
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    np.savez_compressed(os.path.join(OUT, "scan_preprocess.npz"), raw=raw, merge_xyz=mx, merge_nrm=mn, match_xyz=ax, match_nrm=an,
                        voxel_keys=keys[order], voxel_means=vx[order])

    # fusion: two scans inserted (first with identity -> duplication quirk)
    poses = synth.loop_trajectory(4)
    map_x = np.zeros((0, 3)); map_n = np.zeros((0, 3))
    for k in range(2):
        raw = synth.lidar_scan(scene, poses[k], n_beams=16, n_az=512, seed=20 + k).astype(np.float64)
        (mx, mn), _ = O.process_scan(raw, wide, wide, 0.1, 20, 3.0, 1.0, 0)
        T = np.eye(4) if k == 0 else np.linalg.inv(poses[0]) @ poses[1]
        c = O.cropper("MinMaxRadius", 2.0, 30.0, center=tuple(T[:3, 3]))   # mapBuilderCropper_->setPose(mapToRangeSensor), Submap.cpp:71
        map_x, map_n, keys = O.submap_insert_scan(map_x, map_n, mx, mn, T, 0.1, c, return_keys=True)
    order = np.lexsort((map_x[:, 2], map_x[:, 1], map_x[:, 0], keys[:, 2], keys[:, 1], keys[:, 0]))
    np.savez_compressed(os.path.join(OUT, "fusion_two_scans.npz"), map_xyz=map_x[order], map_nrm=map_n[order], keys=keys[order])
    # the "next" rows of SURVEY 8f on small inputs: point-to-point / generalized ICP on config 1, overlap + information
    # matrix, sparse-map carving, de-skew (inputs are regenerated from seeds by the test; only the outputs are stored)
    n = {}
    src, tgt, nrm, _ = synth.planar_cloud_config1(n=800, noise=0.01)
    snrm = O.estimate_normals(src, 10, 2.0)
    init = synth.se3(0.01, -0.02, 0.03, (0.05, 0.02, -0.01))
    r = O.registration_icp_p2point(src, tgt, 1.0, init, max_iter=50)
    n["p2p_T"] = r.T; n["p2p_iters"] = r.iters; n["p2p_ncorr"] = r.n_corr; n["p2p_rmse"] = r.inlier_rmse
    r = O.registration_gicp(src, snrm, tgt, nrm, 1.0, init, max_iter=30)
    n["gicp_T"] = r.T; n["gicp_iters"] = r.iters; n["gicp_ncorr"] = r.n_corr; n["gicp_rmse"] = r.inlier_rmse
    fs, ft = O.overlap_flags(src, tgt, init, 0.5, 2)
    n["overlap_src"] = np.packbits(fs); n["overlap_tgt"] = np.packbits(ft)
    n["info"] = O.information_matrix(src, tgt, 0.3, init)
    rng = np.random.default_rng(77)
    raw = rng.normal(size=(300, 3)); raw = raw / np.linalg.norm(raw, axis=1)[:, None] * rng.uniform(3, 9, (300, 1))   # sensor frame
    Ts = synth.se3(t=(5.0, 5.0, 1.0))
    n["carved"] = np.packbits(O.carve(tgt, nrm, O.transform(Ts, raw)[0], Ts[:3, 3], O.cropper("MaxRadius", 0.0, 8.0, center=(5.0, 5.0, 0.0)), 0.25,
                                      20.0, 0.1, 0.3))
    n["deskew"] = O.undistort(src[:50], np.array([5.0, -0.4, 0.1]), np.array([0.02, -0.05, 0.8]), 0.1, True)
    np.savez_compressed(os.path.join(OUT, "next_rows.npz"), **n)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
