"""Config 3 of BASELINE.json: voxel down-sample + normal estimation on a 1 M-point cloud, with the HBM roofline.

21 synthetic 64x1024 scans (returns only) concatenated in the map frame and cut to 1 048 576 points, voxel 0.1 m, knn 20, radius 3.0 m: the work of
prepareInitialMap + the initial-map voxelise (core/src/ScanToMapRegistration.cpp:81-84, core/src/Submap.cpp:47-52).
Times each stage with CUDA events on the launching stream (L2 flushed between repetitions), prints one JSON line and writes
it to gpurun_out/config3.json.  Parity of the same workload: tests/test_gpu_parity.py::test_config3_voxel_normals_1m.
usage: python tools/config3_microbench.py [--reps 10]
"""
import argparse, ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from open3d_slam_b200 import engine as E, synth, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--scans", type=int, default=21)
args = ap.parse_args()

VOXEL, KNN, RADIUS = 0.1, 20, 3.0
sc = synth.Scene(); poses = synth.loop_trajectory(600)
parts = []
for i in range(args.scans):
    T = poses[(i * 37) % 600]
    s = synth.lidar_scan(sc, T, seed=1000 + i).astype(np.float64)
    parts.append(s @ T[:3, :3].T + T[:3, 3])
xyz = np.ascontiguousarray(np.vstack(parts)[:1 << 20])   # scans lose their sky rays: 21 scans give > 2^20 returns
N = xyz.shape[0]

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
p = E.MapperParameters()
p.icp.knn = KNN; p.icp.maxDistanceKnn = RADIUS
eng = E.Engine(p, cuda_stream=stream.cuda_stream)
raw = eng.cloud(xyz)
vox = E.Cloud(eng)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = L.lib()


def timed(fn, reps):
    ts = []
    for r in range(reps + 3):
        with torch.cuda.stream(stream):
            flush.fill_(r & 0xFF)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
        stream.synchronize()
        if r >= 3:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


l0 = eng.launches
ms_vox, ms_vox_min = timed(lambda: L.check(lib.b2s_voxel_down_sample(eng._h, raw._c, C.c_double(VOXEL), vox._c)), args.reps)
launches_vox = (eng.launches - l0) // (args.reps + 3)
M = len(vox)
l0 = eng.launches
ms_nrm, ms_nrm_min = timed(lambda: L.check(lib.b2s_estimate_normals(eng._h, vox._c, C.c_int32(KNN), C.c_double(RADIUS))), args.reps)
launches_nrm = (eng.launches - l0) // (args.reps + 3)

peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
bytes_vox = 24.0 * N + 24.0 * M
bytes_nrm = 24.0 * M * (KNN + 2)
out = {"workload": "config3: voxel down-sample + normals, 2^20 returns of 21 scans of 64x1024 in the map frame", "N": N, "M": M, "voxel_size": VOXEL, "knn": KNN, "radius": RADIUS,
       "voxel": {"ms_median": ms_vox, "ms_min": ms_vox_min, "launches": int(launches_vox), "algorithmic_bytes": bytes_vox,
                 "achieved_gbs": bytes_vox / ms_vox / 1e6, "frac_of_hbm_peak": bytes_vox / ms_vox / 1e6 / peak, "mpoints_per_s": N / ms_vox / 1e3},
       "normals": {"ms_median": ms_nrm, "ms_min": ms_nrm_min, "launches": int(launches_nrm), "algorithmic_bytes": bytes_nrm,
                   "achieved_gbs": bytes_nrm / ms_nrm / 1e6, "frac_of_hbm_peak": bytes_nrm / ms_nrm / 1e6 / peak, "mpoints_per_s": M / ms_nrm / 1e3},
       "peak_gbs": peak}

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config3.json"), "w"), indent=1)
print(json.dumps(out))
