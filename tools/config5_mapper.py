"""Config 5 of BASELINE.json on its own: the full mapper (carving, dense map, submap hand-overs, loop-closure refinement) over the
closed synthetic lap.  Thin wrapper over open3d_slam_b200.benchmarks.run_config5 (the same function bench.py reports under "config5");
use it as the target of an ncu capture of the carving / dense-map kernels.
usage: python tools/config5_mapper.py [--scans 354]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open3d_slam_b200 import benchmarks as B

ap = argparse.ArgumentParser()
ap.add_argument("--scans", type=int, default=354)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
print(json.dumps(B.run_config5(dev, torch.cuda.Stream(device=dev), 1, 0, None, n_scans=args.scans)))
