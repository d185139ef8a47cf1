"""Per-step device timeline of the decomposed bench workload (torchrun, one rank per GPU); rank 0 prints."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import shpkg
pkg = shpkg.load(); W = pkg.workloads; D = pkg.load_decomp()
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
reps = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4,3,2").split(","))
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
newton = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = W.tiled_packing(reps); cfg["v"] = cfg["v"] + np.array([15.0, 0, 0])
g = D.native_engine(pkg, cfg, local, tuning={"newton": newton, "step_trace": 1})
g.compute_forces(); g.run(60); g.reset_timers()
g.run(nsteps)
ms, fl = g.get_step_trace()
if rank == 0:
    print("newton", newton, "world", world, "mean %.3f median %.3f" % (ms.mean(), np.median(ms)), g.dd_info())
    for k in range(nsteps):
        if fl[k] or ms[k] > 1.3 * np.median(ms):
            print("step %3d %.3f ms flags %d" % (k, ms[k], fl[k]))
    print(g.get_cache_stats(), g.get_split_times(), g.get_timers(), g.get_counters())
dist.destroy_process_group()
