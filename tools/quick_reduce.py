import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, shpkg
pkg = shpkg.load(); W = pkg.workloads
cfg = W.tiled_packing((4, 3, 2)); cfg["v"] = cfg["v"] + np.array([15.0, 0, 0])
for occ in (6, 4, 8):
    g = pkg.ShGpu(); W.apply(g, cfg); g.set_tuning("reduce_occ", occ); g.compute_forces(); g.run(40); g.reset_timers(); g.run(100)
    print("reduce_occ", occ, {k: round(1e3 * v / 100, 4) for k, v in g.get_split_times().items()}, "ms/step", 1e3 * g.get_run_time()["last"] / 100); g.close()
