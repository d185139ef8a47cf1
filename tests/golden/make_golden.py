#!/usr/bin/env python
"""Writes the golden fixtures in this directory FROM THIS REPO'S CPU ORACLE (oracle/sh_oracle.c).

PARITY UNPINNED: the reference mount holds no source, test or vector (README only), so these are not reference outputs;
they pin the oracle against itself across rounds (a change of its arithmetic shows up as a diff here) and give the GPU
tests a committed target that does not need the oracle at run time.  Regenerate with `python tests/golden/make_golden.py`.
Cases: (1) BASELINE configs[0]: two SH ellipsoids (l_max=20, 32x64) at three separations / orientations, contact exponents
1 and 1.5; (2) a periodic 4-shape l_max=12 packing of 108 particles; (3) the same with the dissipative terms (A.5b);
(4) a Lees-Edwards sheared box after 40 steps.  Per case: the inputs needed to rebuild it (config name + seed) and the
per-pair V, F, torques (sorted by tag pair), per-atom forces, counters."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
import shpkg  # noqa: E402

W = shpkg.load().workloads


def cases():
    out = {}
    rng = np.random.default_rng(7)
    for t, ex in ((0, 1.0), (1, 1.5), (2, 1.0)):
        cfg = W.config1_two_particle(seed=100 + t, exponent=ex)
        sep = (1.35, 1.7, 1.95)[t]
        dirv = rng.normal(size=3); dirv /= np.linalg.norm(dirv)
        cfg["x"] = np.array([-0.5 * sep * dirv, 0.5 * sep * dirv])
        out["two_particle_%d" % t] = (cfg, 0)
    out["packing_l12"] = (W.packing((3, 3, 3), 12, (16, 32), nshapes=4, seed=11, periodic=True, name="g", nn_frac=1.8), 0)
    cfg = W.packing((3, 3, 3), 12, (16, 32), nshapes=4, seed=11, periodic=True, name="g", nn_frac=1.8, vel_sigma=0.8)
    cfg["angmom"] = np.random.default_rng(3).normal(0, 0.5, size=cfg["angmom"].shape)
    cfg["dissipation"] = (3.0, 2.0, 0.4)
    out["packing_l12_dissipative"] = (cfg, 0)
    cfg = W.shear_box(W.packing((4, 4, 3), 12, (16, 32), nshapes=2, seed=13, periodic=True, name="g", nn_frac=1.85, vel_sigma=0.3,
                                dt=4e-4, skin=0.04), 0.5)
    cfg["v"] = cfg["v"] + np.array([0.0, 6.0, 0.0])
    out["shear_box_40_steps"] = (cfg, 40)
    return out


def run_case(cfg, nsteps):
    o = O.Oracle(threads=4)
    W.apply(o, cfg)
    if nsteps:
        o.run(nsteps)
    else:
        o.compute_forces()
    p, a, c = o.get_pairs(), o.get_atoms(), o.get_counters()
    order = np.lexsort((p["tag_j"], p["tag_i"]))
    res = {"tag_i": p["tag_i"][order], "tag_j": p["tag_j"][order], "V": p["V"][order], "F": p["F"][order],
           "tau_i": p["tau_i"][order], "tau_j": p["tau_j"][order], "f": a["f"], "torque": a["torque"], "x": a["x"],
           "nodes_inside": np.int64(c["nodes_inside"]), "pair_evals": np.int64(c["pair_evals"])}
    o.close()
    return res


if __name__ == "__main__":
    for name, (cfg, nsteps) in cases().items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(cfg, nsteps))
        print("wrote", name)
