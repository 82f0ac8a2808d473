"""Throughput probe (GPU box): scans/s of single stages when many chains run them concurrently (one stream + host thread each).
Identifies which stage saturates the GPU.  usage: python tools/stage_probe.py [chains]"""
import ctypes as C, os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from open3d_slam_b200 import engine as E, synth, _lib as L

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = 20
dev = torch.device("cuda", 0)
streams = [torch.cuda.Stream(device=dev) for _ in range(chains)]
p = E.MapperParameters(seed=3)
sc = synth.Scene(); poses = synth.loop_trajectory(4)
raws = [synth.lidar_scan(sc, poses[k], seed=k) for k in range(3)]
engs = [E.Engine(p, cuda_stream=s.cuda_stream) for s in streams]
maps = [E.Mapper(e, 600_000) for e in engs]
raw_c = [e.cloud(raws[1]) for e in engs]
for m, e in zip(maps, engs):
    m.addRangeMeasurement(e.cloud(raws[0]), None)
    m.addRangeMeasurement(e.cloud(raws[1]), np.linalg.inv(poses[0]) @ poses[1])
merge = [E.Cloud(e) for e in engs]; match = [E.Cloud(e) for e in engs]
for c in range(chains):
    L.check(L.lib().b2s_process_scan(engs[c]._h, raw_c[c]._c, merge[c]._c, match[c]._c)); engs[c].synchronize()
pool = ThreadPoolExecutor(max_workers=min(8, chains))
T1 = np.ascontiguousarray(np.linalg.inv(poses[0]) @ poses[1])
pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


def run(name, fn, sync_each=False):
    def body(c):
        for _ in range(reps):
            fn(c)
    for w in range(2):
        list(pool.map(lambda c: fn(c), range(chains)))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    list(pool.map(body, range(chains)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:28s} chains {chains:3d}: {chains * reps / dt:9.1f} /s   {1e6 * dt / (chains * reps):8.1f} us exclusive each")


run("process_scan (S1)", lambda c: L.check(L.lib().b2s_process_scan(engs[c]._h, raw_c[c]._c, merge[c]._c, match[c]._c)))
t0c = [E.Cloud(e) for e in engs]
run("voxel_down_sample only", lambda c: L.check(L.lib().b2s_voxel_down_sample(engs[c]._h, raw_c[c]._c, C.c_double(0.1), t0c[c]._c)))
vx = [E.voxelize(e, rc, 0.1) for e, rc in zip(engs, raw_c)]
run("estimate_normals (all pts)", lambda c: L.check(L.lib().b2s_estimate_normals(engs[c]._h, vx[c]._c, C.c_int32(20), C.c_double(3.0))))
res = [L.Result() for _ in range(chains)]
run("register_to_submap (S2)", lambda c: L.check(L.lib().b2s_register_to_submap(engs[c]._h, match[c]._c, maps[c].submap._s, pd(T1), pd(T1), C.byref(res[c]))))
print("   (S2 call synchronises: result copy)  iters", res[0].iters)
run("submap_insert (F1)", lambda c: L.check(L.lib().b2s_submap_insert(engs[c]._h, maps[c].submap._s, merge[c]._c, pd(T1))))
