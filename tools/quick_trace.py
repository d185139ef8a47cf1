"""Per-step wall-clock trace of the bench workload (ad hoc): shows what rebuild steps cost."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import shpkg
pkg = shpkg.load(); W = pkg.workloads
reps = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4,3,2").split(","))
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 80
flow = float(sys.argv[3]) if len(sys.argv) > 3 else 15.0
cfg = W.tiled_packing(reps); cfg["v"] = cfg["v"] + np.array([flow, 0, 0])
g = pkg.ShGpu(); W.apply(g, cfg); g.compute_forces(); g.run(3); g.reset_timers()
prev = g.get_counters(); pc = g.get_cache_stats(); pt = g.get_timers()
for s in range(nsteps):
    t0 = time.perf_counter(); g.run(1); dt = time.perf_counter() - t0
    c = g.get_counters(); cs = g.get_cache_stats(); tm = g.get_timers()
    nb, cb = c["neighbor_builds"] - prev["neighbor_builds"], cs["cache_builds"] - pc["cache_builds"]
    if nb or cb or dt > 3e-3 or s < 3:
        print("step %3d wall %.3f ms dev %.3f ms nb %d cb %d neigh_ms %.3f cache_ms %.3f level %d slow %d" % (
            s, 1e3 * dt, 1e3 * g.get_run_time()["last"], nb, cb, 1e3 * (tm["seconds_neigh"] - pt["seconds_neigh"]),
            1e3 * (cs["seconds_cache"] - pc["seconds_cache"]), cs["level"], cs["slow_pairs"] - pc["slow_pairs"]), flush=True)
    prev, pc, pt = c, cs, tm
print(g.get_split_stats(), g.get_split_times())
