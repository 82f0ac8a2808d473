"""Config 3 of BASELINE.json on its own: voxel down-sample + normal estimation on a 2^20-point cloud, with the HBM
roofline fractions.  Thin wrapper over open3d_slam_b200.benchmarks.run_config3 (the same function bench.py reports under "config3");
use it as the target of an ncu capture.  Parity of the workload: tests/test_gpu_parity.py::test_config3_voxel_normals_1m.
usage: python tools/config3_microbench.py [--reps 8] [--out gpurun_out/config3.json]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open3d_slam_b200 import benchmarks as B

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--out", default="gpurun_out/config3.json")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
res = B.run_config3(dev, torch.cuda.Stream(device=dev), reps=args.reps)
line = json.dumps(res)
print(line)
if args.out:
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    open(args.out, "w").write(line + "\n")
