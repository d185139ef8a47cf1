import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shpkg
pkg = shpkg.load(); W = pkg.workloads
cfg = W.tiled_packing((4, 3, 2)); g = pkg.ShGpu(); W.apply(g, cfg); g.compute_forces(); g.run(3); g.reset_timers(); g.run(3)
print(g.get_timers(), g.get_split_stats(), g.get_split_times(), g.get_counters())
