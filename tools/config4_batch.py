"""Config 4 of BASELINE.json: 512 independent scan-submap registrations, sharded over the GPUs of one box.

Pair i: target = a 20 m-radius submap fused from six synthetic 64x1024 scans around loop position i (built with
this engine's own S1 + F1 at the ground-truth poses), source = the preprocessed scan from the middle of that stretch, initial guess = its true pose
displaced by a random SE(3) within (+-0.5 m, +-5 deg) (seed = pair id); r = 0.3, max_iter = 100
(core/src/PlaceRecognition.cpp:45-46,111).  Every pair has its own target, so every registration pays its own index build
(the reference rebuilds a KD-tree per RegistrationICP call).  Pairs are split by index across ranks with no data-path
collective; the 19 result scalars per pair are gathered at the end.  Timed with CUDA events around b2s_register_batch
(index builds + one batched ICP launch + result copy), max over ranks.  Parity of this workload against the oracle:
tests/test_gpu_parity.py::test_config4_scan_submap_pairs_batch.

usage: python tools/config4_batch.py [--pairs 512] [--reps 5]
       python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/config4_batch.py
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from open3d_slam_b200 import engine as E, synth, dist as D

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=512)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--scans-per-submap", type=int, default=6, help="scans (every 2nd loop position) fused into each target submap")
args = ap.parse_args()

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
mine = D.shard_range(args.pairs, world, rank)

p = E.MapperParameters(seed=3)
stream = torch.cuda.Stream(device=dev)
eng = E.Engine(p, device=local, cuda_stream=stream.cuda_stream)
icp = E.ScanToMapIcp(eng)
sc = synth.Scene(); poses = synth.loop_trajectory(600)
casts = {}


def scan(k):
    k %= 600
    if k not in casts:
        casts[k] = synth.lidar_cast(sc, poses[k])
    return synth.lidar_from_cast(casts[k], seed=k), poses[k]


targets, sources, inits, truth = [], [], [], []
import copy
p_full = copy.deepcopy(p); p_full.scanProcessing.downSamplingRatio = 1.0   # submaps keep every voxel of the scans they fuse
eng.set_parameters(p_full)
for i in mine:
    sm = E.Submap(eng, 600_000)
    for k in range(i, i + 2 * args.scans_per_submap, 2):
        raw, T = scan(k)
        ps = icp.processForScanMatchingAndMerging(eng.cloud(raw.astype(np.float64)))
        sm.insertScan(None, ps.merge_, T)
    xyz, nrm = sm.getMapPointCloud()
    targets.append(eng.cloud(xyz, nrm))
    sm.free()
eng.set_parameters(p)
for i in mine:
    raw, T = scan(i + args.scans_per_submap)
    sources.append(icp.processForScanMatchingAndMerging(eng.cloud(raw.astype(np.float64))).match_)
    rng = np.random.default_rng(i)
    inits.append(T @ synth.se3(*np.deg2rad(rng.uniform(-5, 5, 3)), rng.uniform(-0.5, 0.5, 3)))
    truth.append(T)

pc = E.CloudRegistrationParameters(icp=p.icp)
pc.icp.maxCorrespondenceDistance = 0.3
pc.icp.maxNumIter = 100
reg = E.RegistrationIcpPointToPlane(eng, pc)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = []
for r in range(args.reps + 2):
    with torch.cuda.stream(stream):
        flush.fill_(r)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        a.record(stream)
        res = reg.registerCloudsBatch(sources, targets, inits)
        b.record(stream)
    stream.synchronize()
    if r >= 2:
        times.append(a.elapsed_time(b))
eng.profile_enable(True)     # one more repetition with per-kernel-group CUDA events (not part of the timing above)
res = reg.registerCloudsBatch(sources, targets, inits)
prof = {k: round(v[0], 3) for k, v in eng.profile_read().items() if v[1] > 0}
eng.profile_enable(False)
ms = D.max_over_ranks(float(np.median(times)), world, dev)
tab = np.array([[*r.transformation_.ravel(), r.fitness_, r.inlier_rmse_, r.iters] for r in res]).reshape(len(res), 19)
err = np.array([np.linalg.norm(r.transformation_[:3, 3] - T[:3, 3]) for r, T in zip(res, truth)])
full = D.gather_results(np.c_[tab, err], args.pairs, world, rank, dev)
if rank == 0:
    n_src = float(np.mean([len(s) for s in sources])); n_tgt = float(np.mean([len(t) for t in targets]))
    out = {"workload": "config4: independent scan-submap pairs, r=0.3, max_iter=100", "pairs": args.pairs, "n_gpus": world,
           "ms_per_batch": ms, "kernel_group_ms_rank0": prof, "registrations_per_s": args.pairs / ms * 1e3, "mean_source_points": n_src, "mean_target_points": n_tgt,
           "mean_iters": float(full[:, 18].mean()), "min_fitness": float(full[:, 16].min()),
           "median_translation_error_m": float(np.median(full[:, 19])), "max_translation_error_m": float(full[:, 19].max()),
           "timing": "CUDA events around b2s_register_batch (per-pair index build + one batched ICP launch + D2H of results), median of %d, max over ranks; 256 MiB L2 flush before each" % args.reps}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"config4_n{world}.json"), "w"), indent=1)
    print(json.dumps(out))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
