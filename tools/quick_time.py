"""Ad-hoc timing helper (not a test): python tests/quick_time.py N threads variant [lmax nt nphi nn_frac]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shpkg
pkg = shpkg.load(); W = pkg.workloads
n = int(sys.argv[1]); tw = int(sys.argv[2]); var = int(sys.argv[3])
lmax = int(sys.argv[4]) if len(sys.argv) > 4 else 30
nt = int(sys.argv[5]) if len(sys.argv) > 5 else 48
nphi = int(sys.argv[6]) if len(sys.argv) > 6 else 96
nn = float(sys.argv[7]) if len(sys.argv) > 7 else 1.9
m = max(2, int(round((n / 4.0) ** (1.0 / 3.0))))
cfg = W.packing((m, m, m), lmax, (nt, nphi), nshapes=8, seed=30, nn_frac=nn)
g = pkg.ShGpu(); W.apply(g, cfg); g.set_pair_tuning(tw, 0, var)
g.compute_forces(); g.run(3); g.reset_timers(); g.run(5)
t = g.get_timers(); c = g.get_counters()
T = (lmax + 1) * (lmax + 2) // 2
fl = 24.0 * c["nodes_transformed"] + (7 * T + 14 * (lmax + 1) + 40) * c["nodes_evaluated"] + 30.0 * c["nodes_inside"] + 200.0 * c["pair_evals"]
ms = 1e3 * t["seconds_pair"] / t["pair_launches"]
print("n %d tw %d var %d L %d nn %.2f: pair %.3f ms  %.2f Mpairs/s  %.2f TFLOP/s  eval/pair %.1f trans/pair %.1f inside/pair %.1f" % (
    len(cfg["x"]), tw, var, lmax, nn, ms, c["pair_evals"] / t["seconds_pair"] / 1e6, fl / t["seconds_pair"] / 1e12,
    c["nodes_evaluated"] / c["pair_evals"], c["nodes_transformed"] / c["pair_evals"], c["nodes_inside"] / c["pair_evals"]))
