"""Per-step DEVICE timeline of the bench workload inside one sh_run (no host syncs between steps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import shpkg
pkg = shpkg.load(); W = pkg.workloads
reps = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4,3,2").split(","))
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
flow = float(sys.argv[3]) if len(sys.argv) > 3 else 15.0
sync = int(sys.argv[4]) if len(sys.argv) > 4 else 0
shearv = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
cfg = W.tiled_packing(reps); cfg["v"] = cfg["v"] + np.array([flow, 0, 0])
if shearv:
    cfg = W.shear_box(cfg, shearv / float(cfg["box"][1][1] - cfg["box"][0][1]))
g = pkg.ShGpu(); W.apply(g, cfg); g.set_tuning("sync_rebuild", sync); g.set_tuning("step_trace", 1)
g.compute_forces(); g.run(40); g.reset_timers()
g.run(nsteps)
ms, fl = g.get_step_trace()
print("sync_rebuild", sync, "mean %.3f median %.3f" % (ms.mean(), np.median(ms)))
for k in range(nsteps):
    if fl[k] or ms[k] > 1.3 * np.median(ms):
        print("step %3d %.3f ms flags %d" % (k, ms[k], fl[k]))
print(g.get_cache_stats(), g.get_split_times(), g.get_timers(), g.get_counters(), g.dd_info())
