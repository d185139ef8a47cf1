#!/usr/bin/env python
"""Generate a mechanically relaxed periodic packing with the GPU code itself (cfg 3:
"poured/compressed ... by the oracle-validated GPU code itself, snapshot saved; then timed").

Protocol: dilute FCC-seeded periodic box, random orientations -> isotropic compression (box and
positions rescaled every few steps) under viscous damping until the contact energy per particle
reaches a small positive value (jammed, small overlaps) -> damped relaxation -> snapshot (.npz).
The snapshot is a periodic unit cell; workloads.tiled_packing() replicates it to any size.

  python tools/make_packing.py --cells 10 --lmax 30 --ntheta 48 --nphi 96 --out <file.npz>
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shpkg

pkg = shpkg.load()
W = pkg.workloads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=10)
    ap.add_argument("--lmax", type=int, default=30)
    ap.add_argument("--ntheta", type=int, default=48)
    ap.add_argument("--nphi", type=int, default=96)
    ap.add_argument("--seed", type=int, default=30)
    ap.add_argument("--e-target", type=float, default=0.1, help="contact energy per particle at which compression stops")
    ap.add_argument("--shrink", type=float, default=0.999)
    ap.add_argument("--every", type=int, default=40)
    ap.add_argument("--relax", type=int, default=3000)
    ap.add_argument("--relax-check", type=int, default=1000)
    ap.add_argument("--e-jam", type=float, default=0.02, help="relaxed contact energy per particle that counts as jammed")
    ap.add_argument("--z-min", type=float, default=4.0)
    ap.add_argument("--damp", type=float, default=3.0)
    ap.add_argument("--out", default="gpurun_out/packing.npz")
    args = ap.parse_args()

    m = args.cells
    cfg = W.packing((m, m, m), args.lmax, (args.ntheta, args.nphi), nshapes=8, seed=args.seed, nn_frac=2.4,
                    vel_sigma=0.3, k=1e3, dt=4e-4, skin=0.1)
    n = len(cfg["x"])
    sim = pkg.ShGpu()
    W.apply(sim, cfg)
    sim.set_damping(args.damp, args.damp)
    vols = np.array([sim.shape_props(s)["volume"] for s in range(len(cfg["shapes"]))])
    vtot = vols[cfg["shape_id"]].sum()
    box = np.array(cfg["box"][1], dtype=float)
    t0 = time.time()
    it = 0
    while True:
        sim.run(args.every)
        e = sim.get_energy()
        phi = vtot / np.prod(box)
        if it % 10 == 0:
            print("it %4d phi %.4f e_contact/N %.4g ke/N %.4g  (%.0f s)" % (it, phi, e["e_contact"] / n,
                  (e["ke_trans"] + e["ke_rot"]) / n, time.time() - t0), flush=True)
        if phi > 0.85:
            break
        if e["e_contact"] / n >= args.e_target:
            # candidate: relax with strong damping; accept only if the packing stays jammed
            sim.set_damping(4.0, 4.0)
            sim.run(args.relax_check)
            sim.set_damping(args.damp, args.damp)
            er = sim.get_energy()
            pr = sim.get_pairs()
            z = 2.0 * (pr["V"] > 0).sum() / n
            print("   relaxed at phi %.4f: e_contact/N %.4g z %.2f" % (phi, er["e_contact"] / n, z), flush=True)
            if er["e_contact"] / n >= args.e_jam and z >= args.z_min:
                break
        st = sim.get_atoms(("x",))
        x = st["x"] * args.shrink
        box = box * args.shrink
        sim.set_box(np.zeros(3), box, (1, 1, 1))
        sim.put_state(x=np.ascontiguousarray(x))
        it += 1
    sim.set_damping(3.0, 3.0)
    sim.run(args.relax)
    e = sim.get_energy()
    sim.reset_timers()
    sim.compute_forces()
    c = sim.get_counters()
    pr = sim.get_pairs()
    st = sim.get_atoms(("x", "quat", "v", "angmom"))
    x = st["x"] - box * np.floor(st["x"] / box)
    phi = vtot / np.prod(box)
    stats = dict(n=n, phi=phi, e_contact_per_particle=e["e_contact"] / n, ke_per_particle=(e["ke_trans"] + e["ke_rot"]) / n,
                 pairs_per_particle=len(pr["V"]) / n, contacts_per_particle=2.0 * (pr["V"] > 0).sum() / n,
                 eval_nodes_per_pair=c["nodes_evaluated"] / max(1, c["pair_evals"]),
                 inside_nodes_per_pair=c["nodes_inside"] / max(1, c["pair_evals"]),
                 mean_overlap_volume=float(pr["V"][pr["V"] > 0].mean()) if (pr["V"] > 0).any() else 0.0)
    print("final:", stats, "elapsed %.0f s" % (time.time() - t0))
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    np.savez_compressed(args.out, x=x, quat=st["quat"], shape_id=cfg["shape_id"], box=box, lmax=args.lmax,
                        grid=np.array([args.ntheta, args.nphi]), seed=args.seed, nshapes=8,
                        stats=np.array([repr(stats)]))
    print("saved", args.out)


if __name__ == "__main__":
    main()
