"""Short profiling target: the bench workload (96,000 particles, flow 15), warmed past two neighbor rebuilds, then
`nsteps` steps between cudaProfilerStart / cudaProfilerStop (run ncu with --profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import shpkg
pkg = shpkg.load(); W = pkg.workloads
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = tuple(int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "4,3,2").split(","))
lmax = int(sys.argv[3]) if len(sys.argv) > 3 else 30
if lmax == 30:
    cfg = W.tiled_packing(reps); cfg["v"] = cfg["v"] + np.array([15.0, 0, 0])
else:   # configs[4] style: l_max=50, 80x160 nodes, jittered FCC
    cfg = W.config3_packing(int(sys.argv[4]) if len(sys.argv) > 4 else 20000, lmax=lmax, grid=(80, 160), seed=30)
    cfg["v"] = cfg["v"] + np.array([15.0, 0, 0])
g = pkg.ShGpu(); W.apply(g, cfg); g.compute_forces(); g.run(40); g.reset_timers()
torch.cuda.synchronize()
torch.cuda.profiler.start()
g.run(nsteps)
torch.cuda.profiler.stop()
c = g.get_counters()
print("steps", nsteps, "run_ms_per_step", 1e3 * g.get_run_time()["last"] / nsteps, "builds", c["neighbor_builds"], g.get_split_times())
