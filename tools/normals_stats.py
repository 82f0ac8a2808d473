"""Debug aid (GPU box): how the normal-estimation queries of one config-2 scan resolve (B2S_DEBUG_NORMALS counters)."""
import os, sys
os.environ["B2S_DEBUG_NORMALS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from open3d_slam_b200 import engine as E, synth
p = E.MapperParameters(seed=3)
eng = E.Engine(p)
sc = synth.Scene(); poses = synth.loop_trajectory(600)
icp = E.ScanToMapIcp(eng)
for k in (0, 100, 300):
    raw = synth.lidar_scan(sc, poses[k], seed=k).astype(np.float64)
    ps = icp.processForScanMatchingAndMerging(eng.cloud(raw))
    print("scan", k, "merge", len(ps.merge_), "match", len(ps.match_), flush=True)
