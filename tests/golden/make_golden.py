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
        map_x, map_n, keys = O.submap_insert_scan(map_x, map_n, mx, mn, T, 0.1, wide, return_keys=True)
    order = np.lexsort((map_x[:, 2], map_x[:, 1], map_x[:, 0], keys[:, 2], keys[:, 1], keys[:, 0]))
    np.savez_compressed(os.path.join(OUT, "fusion_two_scans.npz"), map_xyz=map_x[order], map_nrm=map_n[order], keys=keys[order])
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
