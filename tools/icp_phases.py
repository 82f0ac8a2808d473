"""Debug aid (GPU box): per-evaluation clock64 phases of the ICP kernel for a config-2 scan-to-map registration."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open3d_slam_b200 import engine as E, synth, _lib as L

p = E.MapperParameters(seed=3)
p.nnCellSize = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
eng = E.Engine(p)
sc = synth.Scene(); poses = synth.loop_trajectory(8)
m = E.Mapper(eng, 800_000)
buf = (C.c_longlong * 1024)()
for k in range(6):
    raw = synth.lidar_scan(sc, poses[k], seed=k)
    if k == 5:
        L.check(L.lib().b2s_debug_icp_clocks(eng._h, 1, None))
    m.addRangeMeasurement(eng.cloud(raw), np.eye(4) if k == 0 else np.linalg.inv(poses[k - 1]) @ poses[k])
L.check(L.lib().b2s_debug_icp_clocks(eng._h, 1, buf))
a = np.array(buf[:256]).reshape(64, 4)
bal = np.array(buf[256:512]).reshape(8, 8, 4)
cnt = np.array(buf[512:768]).reshape(32, 8)[:, :4]
r = m.lastResult
print("iters", r.iters, "n_corr", r.n_corr, "fitness", r.fitness_)
for e in range(r.iters + 1):
    t = a[e]
    nxt = a[e + 1][0] if e < r.iters else t[3]
    print(f"eval {e}: search {t[1]-t[0]:8d}  reduce+cluster {t[2]-t[1]:8d}  solve {t[3]-t[2]:8d} cycles")
for e in range(min(r.iters + 1, 8)):
    print(f"eval {e} per-CTA phase1 ns {bal[e,:,0].tolist()} phase2 ns {bal[e,:,1].tolist()} queue {bal[e,:,2].tolist()} pts {bal[e,:,3].tolist()}")
for e in range(r.iters + 1):
    c = cnt[e]
    print(f"eval {e}: phase-1 candidates {c[0]:8d} over {c[1]:6d} point evaluations ({c[0] / max(c[1], 1):6.1f} each), queued {c[2]:5d}, phase-2 candidates {c[3]:8d} ({c[3] / max(c[2], 1):7.1f} each)")
