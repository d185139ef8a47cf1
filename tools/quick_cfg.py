"""Ad-hoc timing of BASELINE configs 2 and 5 style workloads (not a test)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shpkg
pkg = shpkg.load(); W = pkg.workloads
which = sys.argv[1]
if which == "cfg2":
    cfg = W.config2_wall(10)
    g = pkg.ShGpu(); W.apply(g, cfg); g.compute_forces(); g.run(200); g.reset_timers()
    t0 = time.perf_counter(); g.run(2000); wall = time.perf_counter() - t0
    t = g.get_timers(); c = g.get_counters(); rt = g.get_run_time()
    print("cfg2 1000 particles wall: %.3f ms/step device, %.3f wall; pair %.3f ms/launch; builds %d; pairs/step %.0f; %.2f Mparticle-steps/s"
          % (1e3 * rt["last"] / 2000, 1e3 * wall / 2000, 1e3 * t["seconds_pair"] / max(1, t["pair_launches"]), c["neighbor_builds"],
             c["pair_evals"] / 2000, 1000 * 2000 / rt["last"] / 1e6), g.get_energy())
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    m = max(2, int(round((n / 4.0) ** (1.0 / 3.0))))
    cfg = W.packing((m, m, m), 50, (80, 160), nshapes=8, seed=50, nn_frac=1.85, name="cfg5")
    g = pkg.ShGpu(); W.apply(g, cfg); g.compute_forces(); g.run(2); g.reset_timers(); g.run(3)
    t = g.get_timers(); c = g.get_counters(); L = 50; T = (L + 1) * (L + 2) // 2
    fl = 24.0 * c["nodes_transformed"] + (7 * T + 14 * (L + 1) + 40) * c["nodes_evaluated"] + 30.0 * c["nodes_inside"] + 200.0 * c["pair_evals"]
    print("cfg5 n %d l50 80x160: pair %.2f ms/launch, %.2f Mpairs/s, %.2f TFLOP/s, eval/pair %.0f trans/pair %.0f" % (
        len(cfg["x"]), 1e3 * t["seconds_pair"] / t["pair_launches"], c["pair_evals"] / t["seconds_pair"] / 1e6, fl / t["seconds_pair"] / 1e12,
        c["nodes_evaluated"] / c["pair_evals"], c["nodes_transformed"] / c["pair_evals"]))
