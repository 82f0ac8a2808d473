#!/usr/bin/env python
"""bench.py -- scan-to-map ICP registrations/s on synthetic 64-beam x 1024-azimuth LiDAR clouds (BASELINE.json metric).

Workload (config[1] of BASELINE.json): the scan-to-map odometry loop with the reference's Lua defaults
(voxel 0.1 m, MinMaxRadius 2-30 m, knn 20 / 3 m, max corr. 1 m, <= 50 iterations, map voxel 0.1 m, PointToPlaneIcp,
downsampling ratio 0.3, fitness gate 0.7).  One "step" = every one of the `chains` independent odometry chains on
this GPU advances by ONE scan through the whole hot path:
    S1 crop+voxel+normals+select -> S2 map-patch crop + NN index + point-to-plane ICP -> fitness gate -> F1 map fusion.
Chains are independent trajectories (one b2s handle / CUDA stream each), the units that shard across GPUs
(weak scaling, no data-path collective).  Within a chain the scans stay strictly sequential.

  value : registrations/s with the raw scans already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e   : the same, every step uploading that step's float32 scans from pinned host memory and reading every
          chain's RegistrationResult back to the host, through the public mapper API
  roofline     : the dominant kernel group, timed live with CUDA events on its own stream (chain 0)
  cpu_baseline : the CPU oracle (oracle/, "port" of the reference's Open3D path) on a bounded sample, rank 0 only
  --impl reference : the same workload on the host cores through the oracle only (no GPU code on that path)
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from open3d_slam_b200 import synth  # noqa: E402

METRIC = "scan-to-map ICP registrations/sec (64x1024-pt clouds)"
UNIT = "registrations/s"
WORKLOAD = "config2: scan-to-map odometry loop, synthetic 64x1024 LiDAR, voxel 0.1 m, Lua defaults, PointToPlaneIcp"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def build_scans(n_scans, chains, noise=0.02):
    """Ray-cast the trajectory once, then one noise realisation per chain.  Returns poses, deltas and scans[c][k] (f32)."""
    scene = synth.Scene()
    poses = synth.loop_trajectory(n_scans)
    casts = [synth.lidar_cast(scene, P) for P in poses]
    scans = [[synth.lidar_from_cast(casts[k], noise, seed=1000 * c + k) for k in range(n_scans)] for c in range(chains)]
    rng = np.random.default_rng(12345)
    deltas = [np.eye(4)]
    for k in range(1, n_scans):
        pert = synth.se3(0.0, 0.0, rng.normal(0, 2e-3), rng.normal(0, 0.02, 3))   # stands in for the lidar odometry error
        deltas.append(np.linalg.inv(poses[k - 1]) @ poses[k] @ pert)
    return poses, deltas, scans


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restating the reference's own CPU path (never on the product path)
# ----------------------------------------------------------------------------------------------------------------------
class OracleChain:
    def __init__(self, ratio, seed):
        from oracle import oracle as O
        self.O = O
        self.wide = O.cropper("MinMaxRadius", 2.0, 30.0)
        self.narrow = O.cropper("MinMaxRadius", 2.0, 30.0)
        self.ratio, self.seed = ratio, seed
        self.map_x = np.zeros((0, 3)); self.map_n = np.zeros((0, 3)); self.pose = np.eye(4)
        self.first = True

    def step(self, raw32, delta):
        O = self.O
        (mx, mn), (ax, an) = O.process_scan(raw32.astype(np.float64), self.wide, self.narrow, 0.1, 20, 3.0, self.ratio, self.seed)
        if self.first:
            self.map_x, self.map_n = O.submap_insert_scan(self.map_x, self.map_n, mx, mn, np.eye(4), 0.1, self.wide)
            self.first = False
            return None
        c = O.cropper("MinMaxRadius", 2.0, 30.0, center=self.pose[:3, 3])
        px, pn = O.crop(c, self.map_x, self.map_n)
        res = O.registration_icp_p2plane(ax, px, pn, 1.0, self.pose @ delta, max_iter=50)
        if res.fitness >= 0.7:
            self.pose = res.T
            self.map_x, self.map_n = O.submap_insert_scan(self.map_x, self.map_n, mx, mn, self.pose, 0.1, self.wide)
        return res


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import oracle as O
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    chains = args.chains
    n_scans = args.warmup + args.steps + 1
    poses, deltas, scans = build_scans(n_scans, chains)
    per = max(1, cores // chains)
    os.environ["OMP_NUM_THREADS"] = str(per)
    O.lib()
    cs = [OracleChain(args.ratio, 3) for _ in range(chains)]
    pool = ThreadPoolExecutor(max_workers=chains)

    def do_step(k):
        list(pool.map(lambda c: cs[c].step(scans[c][k], deltas[k]), range(chains)))

    do_step(0)
    for k in range(1, 1 + args.warmup):
        do_step(k)
    t0 = time.perf_counter()
    for k in range(1 + args.warmup, n_scans):
        do_step(k)
    dt = time.perf_counter() - t0
    value = chains * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains_per_gpu": chains, "downsampling_ratio": args.ratio, "points_per_scan": int(len(scans[0][1]))},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{chains} chains x {args.steps} scans, {chains} chains in parallel x {per} OpenMP threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(scans0, deltas, ratio, n_sample):
    from oracle import oracle as O
    O.lib()
    ch = OracleChain(ratio, 3)
    ch.step(scans0[0], deltas[0])
    ch.step(scans0[1], deltas[1])  # warm the thread pool
    t0 = time.perf_counter()
    for k in range(2, 2 + n_sample):
        ch.step(scans0[k], deltas[k])
    dt = time.perf_counter() - t0
    return {"value": n_sample / dt, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
            "sample": f"1 chain x {n_sample} consecutive scans of the same workload (crop+voxel+normals+select, KD-tree rebuild + ICP, map fusion), "
                      f"{O.num_threads()} OpenMP threads"}


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def run_b2s_arm(args):
    import torch
    import torch.distributed as dist
    from open3d_slam_b200 import engine as E
    from open3d_slam_b200 import _lib as L

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the b2s engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    chains, K, W = args.chains, args.steps, args.warmup
    n_scans = W + K + 1
    poses, deltas, scans = build_scans(n_scans, chains)
    pts = int(np.mean([len(s) for s in scans[0]]))

    params = E.MapperParameters(seed=3)
    params.scanProcessing.downSamplingRatio = args.ratio
    params.nnCellSize = args.nn_cell
    main = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(chains)]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    use_graph = not args.no_graph

    def make_chains(n=None, graph=None):
        n = chains if n is None else n
        graph = use_graph if graph is None else graph
        engs = [E.Engine(params, device=local, cuda_stream=streams[c].cuda_stream) for c in range(n)]
        maps = [E.Mapper(e, 600_000) for e in engs]
        for c in range(n):   # first scan: pre-process and insert with identity (Mapper.cpp:105-114)
            maps[c].addRangeMeasurement(engs[c].cloud(scans[c][0]), None)
            maps[c].submap.setPose(np.eye(4))
            engs[c].synchronize()
        # CUDA-graph replay of the per-scan chain: every scan goes through a fixed-capacity staging cloud
        staging = [maps[c].enableGraph(65536) for c in range(n)] if graph else None
        return engs, maps, staging

    def timed_region(step_fn, k0, nsteps):
        """nsteps steps; each step is bracketed by events on the main stream, the chain streams fork/join around it and the
        L2 is flushed (256 MiB write) outside the brackets.  Returns total device ms."""
        total = 0.0
        evs = []
        for k in range(k0, k0 + nsteps):
            flush_buf.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(main)
            for s in streams:
                s.wait_event(a)
            step_fn(k)
            for s in streams:
                d = torch.cuda.Event(); d.record(s); main.wait_event(d)
            b.record(main)
            evs.append((a, b))
        torch.cuda.synchronize(dev)
        for a, b in evs:
            total += a.elapsed_time(b)
        return total

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: inputs resident in HBM ----------------
    engs, maps, staging = make_chains()
    dev_clouds = [[engs[c].cloud(scans[c][k]) for k in range(n_scans)] for c in range(chains)]
    for e in engs:
        e.synchronize()
    used_slot = {}

    # one host thread per chain: the ctypes calls release the GIL, so the ~80 kernel launches of the chains are issued
    # concurrently (each on its own stream) instead of one chain after the other
    from concurrent.futures import ThreadPoolExecutor
    # graph replay: a step is a handful of enqueue calls per chain, one host thread keeps all chains busy (measured: no
    # difference between 1 and 8 threads); eager launches (~45 per scan and chain) need the threads
    host_threads = args.host_threads if args.host_threads > 0 else (1 if use_graph else 8)
    pool = ThreadPoolExecutor(max_workers=min(chains, host_threads)) if host_threads > 1 else None

    def fan_out(fn):
        if pool is None:
            for c in range(chains):
                fn(c)
        else:
            list(pool.map(fn, range(chains)))

    def one_resident(c, k):
        if staging is not None:   # device->device copy of the resident scan into the graph's staging cloud (1.3 MB)
            maps[c].stageCopy(dev_clouds[c][k])
            used_slot[(c, k)] = maps[c].addRangeMeasurementAsync(staging[c], deltas[k])
        else:
            used_slot[(c, k)] = maps[c].addRangeMeasurementAsync(dev_clouds[c][k], deltas[k], slot=k % 256)

    def step_resident(k):
        fan_out(lambda c: one_resident(c, k))

    timed_region(step_resident, 1, W)
    l0 = sum(e.launches for e in engs)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ms_total = timed_region(step_resident, 1 + W, K)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(e.launches for e in engs) - l0
    ms_total = max_over_ranks(ms_total)
    value = world * chains * K / (ms_total * 1e-3)

    # results of the timed scans: iterations, source sizes, sanity against ground truth
    results = [[maps[c].fetchResult(used_slot[(c, k)]) for k in range(1 + W, n_scans)] for c in range(chains)]
    iters = np.array([[r.iters for r in rc] for rc in results]); fit = np.array([[r.fitness_ for r in rc] for rc in results])
    nsrc = np.array([[r.n_corr / max(r.fitness_, 1e-12) for r in rc] for rc in results])
    gt_last = np.linalg.inv(poses[0]) @ poses[n_scans - 1]
    pose_err = max(float(np.linalg.norm(maps[c].submap.getPose()[:3, 3] - gt_last[:3, 3])) for c in range(chains))
    map_pts = int(np.mean([m.submap.size() for m in maps]))

    # ---------------- per-kernel-group device times: one eager chain with CUDA events around every kernel group ----------------
    # (graph replay has no per-kernel events; this pass re-runs the same scans of chain 0 eagerly, on its own stream)
    for m in maps:
        m.submap.free()
    for e in engs:
        e.close()
    dev_clouds_keep = dev_clouds[0]
    p_engs, p_maps, _ = make_chains(1, graph=False)
    kp = min(K, 10)
    for k in range(1, 1 + W):
        p_maps[0].addRangeMeasurementAsync(dev_clouds_keep[k], deltas[k], slot=k % 256)
    p_engs[0].synchronize()
    p_engs[0].profile_enable(True)
    p_engs[0].profile_read()
    for k in range(1 + W, 1 + W + kp):
        flush_buf.zero_()
        torch.cuda.synchronize(dev)
        p_maps[0].addRangeMeasurementAsync(dev_clouds_keep[k], deltas[k], slot=k % 256)
    prof = p_engs[0].profile_read()
    p_engs[0].profile_enable(False)
    p_res = [p_maps[0].fetchResult(k % 256) for k in range(1 + W, 1 + W + kp)]
    p_iters = np.array([r.iters for r in p_res]); p_nsrc = np.array([r.n_corr / max(r.fitness_, 1e-12) for r in p_res])
    p_maps[0].submap.free()
    p_engs[0].close()

    # ---------------- roofline of the dominant kernel group (live CUDA events on its stream) ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    kinds = {k: v for k, v in prof.items() if v[1] > 0}
    dom = max(kinds, key=lambda k: kinds[k][0]) if kinds else "icp"
    # algorithmic bytes per launch (fp64 layout: 24 B per point / normal), DESIGN.md section "bytes"
    it0, n0 = p_iters, p_nsrc
    bytes_icp = float(np.mean(72.0 * n0 * (it0 + 1)))
    m_vox = float(np.mean(n0)) / max(args.ratio, 1e-9)              # points entering normal estimation (before the ratio down-sample)
    bytes_by_kind = {"icp": bytes_icp, "normals": 24.0 * m_vox * (20 + 2), "radix_sort": None, "nn_grid_build": None,
                     "voxel": 24.0 * pts + 24.0 * m_vox, "fuse": None, "select": None, "crop": None}
    dur_ms = kinds[dom][0] / kinds[dom][1] if kinds else float("nan")
    ab = bytes_by_kind.get(dom)
    if ab is None:   # fall back to the ICP kernel, whose unit of work is defined
        dom_for_roof = "icp"; ab = bytes_icp; dur_ms = prof["icp"][0] / max(prof["icp"][1], 1)
    else:
        dom_for_roof = dom
    achieved = ab / (dur_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None   # DRAM bytes per launch of that kernel from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        ent = json.load(open(tpath)).get(dom_for_roof)
        if ent:
            traffic, traffic_src = ent["bytes_per_launch"], ent["source"]
    roofline = {"bound": "hbm", "kernel": dom_for_roof, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "bytes_per_launch": ab, "avg_launch_ms": dur_ms,
                "note": "latency-bound: a single registration's working set is L2-resident and its iterations are sequential (SURVEY.md 8d)"}
    profile = {k: {"ms_per_scan": v[0] / kp, "launch_groups_per_scan": v[1] / kp} for k, v in prof.items()}

    # ---------------- e2e: host buffers in, results out, every step ----------------
    for lst in dev_clouds:
        for c_ in lst:
            c_.free()
    engs, maps, staging = make_chains()
    pinned = [[torch.from_numpy(scans[c][k]).pin_memory() for k in range(n_scans)] for c in range(chains)]
    res_sz = ctypes.sizeof(L.Result)
    res_pinned = torch.zeros((chains, n_scans, res_sz), dtype=torch.uint8).pin_memory()   # every step's RegistrationResult lands here
    h2d = sum(int(pinned[c][1 + W].numel()) * 4 for c in range(chains))
    d2h = chains * ctypes.sizeof(L.Result)

    def one_e2e(c, k):
        t = pinned[c][k]   # one C call enqueues: H2D of the float32 scan, the whole chain, D2H of the RegistrationResult
        maps[c].addRangeMeasurementHostAsync(t.data_ptr(), t.shape[0], deltas[k], res_pinned[c, k].data_ptr())

    def step_e2e(k):
        fan_out(lambda c: one_e2e(c, k))

    timed_region(step_e2e, 1, W)
    barrier()
    ms_e2e = timed_region(step_e2e, 1 + W, K)
    barrier()
    ms_e2e = max_over_ranks(ms_e2e)
    e2e_value = world * chains * K / (ms_e2e * 1e-3)
    e2e_res = [[L.Result.from_buffer_copy(res_pinned[c, k].numpy().tobytes()) for k in range(1 + W, 1 + W + K)] for c in range(chains)]
    e2e_fit = float(min(r.fitness for rc in e2e_res for r in rc))
    e2e_same = all(e2e_res[c][k].iters == int(iters[c][k]) for c in range(chains) for k in range(K))   # same scans as the resident leg

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_sample(scans[0], deltas, args.ratio, min(args.cpu_sample, n_scans - 2))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "chains_per_gpu": chains, "global_chains": world * chains, "downsampling_ratio": args.ratio,
                           "points_per_scan": pts, "map_points": map_pts, "mean_icp_iters": float(iters.mean()),
                           "mean_source_points": float(nsrc.mean()), "min_fitness": float(fit.min()), "final_pose_err_m": pose_err,
                           "l2": "256 MiB write between timed steps (outside the event brackets)", "parallelism": f"{world}x{chains} independent chains"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K,
                        "min_fitness": e2e_fit, "same_iteration_counts_as_resident_leg": bool(e2e_same)},
                "gpu_launches": int(launches),
                "roofline": roofline, "profile_chain0": profile, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2s", choices=["b2s", "reference"])
    ap.add_argument("--chains", type=int, default=8, help="independent odometry chains per GPU")
    ap.add_argument("--ratio", type=float, default=0.3, help="scan_processing.downsampling_ratio (Lua default 0.3)")
    ap.add_argument("--cpu-sample", type=int, default=12, help="scans in the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads issuing the chains' launches (0 = auto: 1 with graph replay, 8 eager)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying one CUDA graph per scan")
    ap.add_argument("--nn-cell", type=float, default=0.0, help="NN grid cell edge in metres (0 = max_corr_dist / 4)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b2s_arm(args)


if __name__ == "__main__":
    main()
