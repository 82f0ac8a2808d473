"""Config 4 of BASELINE.json on its own: 512 scan-submap pairs over 64 shared 20 m-radius targets (loop-closure ICP, r = 0.3,
<= 100 iterations), strong scaling over the ranks with the targets built by their owner and broadcast over NCCL.  Thin wrapper over
open3d_slam_b200.benchmarks.run_config4 (the same function bench.py reports under "config4").
usage: python tools/config4_batch.py [--pairs 512] [--targets 64]
       python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/config4_batch.py
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from open3d_slam_b200 import benchmarks as B

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=512)
ap.add_argument("--targets", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--out", default="")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
res = B.run_config4(dev, torch.cuda.Stream(device=dev), world, rank, None, n_pairs=args.pairs, n_targets=args.targets, reps=args.reps)
if rank == 0:
    line = json.dumps(res)
    print(line)
    if args.out:
        open(args.out, "w").write(line + "\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
