"""Ad-hoc timing of the split pipeline's launch knobs on the bench packing (not a bench line)."""
import sys, os, json, itertools, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shpkg
pkg = shpkg.load(); W = pkg.workloads
reps = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4,3,2").split(","))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cfg = W.tiled_packing(reps)
combos = [dict(), dict(cull_lpp=32), dict(cull_wpb=1), dict(cull_wpb=2), dict(cull_wpb=8), dict(cull_lpp=32, cull_wpb=1), dict(cache_level=0),
          dict(cube_n=48, cache_level=0), dict(cube_n=72, cache_level=0), dict(cube_n=96, cache_level=0), dict(cube_n=144, cache_level=0)]
if len(sys.argv) > 3:
    combos = [json.loads(a) for a in sys.argv[3:]]
for kn in combos:
    g = pkg.ShGpu()
    if "cube_n" in kn:
        g.set_tuning("cube_n", kn["cube_n"])
    t0 = time.time()
    W.apply(g, cfg)
    for k, v in kn.items():
        if k != "cube_n":
            g.set_tuning(k, v)
    g.compute_forces(); g.run(5); g.reset_timers()
    g.run(steps)
    c, t, st, cs = g.get_counters(), g.get_timers(), g.get_split_times(), g.get_cache_stats()
    rt = g.get_run_time()["last"]
    np_ = c["pair_evals"] / steps
    print(json.dumps(dict(knobs=kn, n=len(cfg["x"]), ms_step=1e3 * rt / steps, ms_pair=1e3 * t["seconds_pair"] / steps,
                          ms={k: round(1e3 * v / steps, 4) for k, v in st.items()}, cand_per_pair=c["nodes_transformed"] / c["pair_evals"],
                          eval_per_pair=c["nodes_evaluated"] / c["pair_evals"], inside_per_pair=c["nodes_inside"] / c["pair_evals"],
                          pairs=np_, split=g.get_split_stats(), cache=cs, setup_s=round(time.time() - t0, 1))), flush=True)
    g.close()
