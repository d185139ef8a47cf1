#!/usr/bin/env python
"""bench.py — headline benchmark of the SPHERHARM contact hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle; see below)

Workload at N=1: BASELINE.json configs[2] — ~100k polydisperse-shape packing, 8 SH shape types,
l_max=30, 48x96 surface quadrature, periodic box (the configuration the north_star's single-GPU
target is quoted on).  A "step" is one full timestep of sh_run (integrate, neighbor decide/build,
pair kernel, gather, integrate).  `value` = contact-pair evaluations per second with the state
resident in HBM (device time, CUDA events on the library's stream, max over ranks).  `e2e` = the
same metric through the Pair::compute-style C-ABI offload with HOST buffers: every step pushes
x/quat from pinned host memory (sh_put_state), runs sh_compute_forces and reads f/torque back.

Reference arm: the reference (imaranresearch/LAMMPS-SPHERHARM) mount holds only a README, so
there is no reference binary or package to install or compile; `--impl reference` times this
repo's CPU oracle (oracle/, kind="port") with all host threads on a bounded sample of the same
workload.  PARITY/BASELINE UNPINNED — it is the builder's restatement, not LAMMPS+MPI.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import shpkg  # noqa: E402

CULL_DRAM_BYTES_PER_PAIR = 966.0    # pair_cull_kernel, ncu capture r01 (profiles/r01_cull_kernel_ncu_summary.txt): 553 MB / 572,376 pairs
EVAL_DRAM_BYTES_PER_RECORD = 33.2   # measured, see the roofline.traffic note below
METRIC = "contact_pair_evals_per_s"
UNIT = "pair-evals/s"


def f_eval(lmax):
    T = (lmax + 1) * (lmax + 2) // 2
    return 7 * T + 14 * (lmax + 1) + 40


def algorithmic_flops(c, lmax):
    """flops = 24/transformed node + F_eval(L)/evaluated node + 30/inside node + 200/pair."""
    return 24.0 * c["nodes_transformed"] + f_eval(lmax) * c["nodes_evaluated"] + 30.0 * c["nodes_inside"] + \
        200.0 * c["pair_evals"]


class ClockSampler:
    """One streaming `nvidia-smi -lms 50` process for the duration of the timed region."""
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.25)          # let the first samples land before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            for line in out.splitlines():
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 7:
                    try:
                        float(parts[0])
                        rows.append(parts)
                    except ValueError:
                        pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def make_workload(args, pkg, n_particles=None):
    W = pkg.workloads
    n = n_particles or args.particles
    if args.workload == "relaxed" and (args.lmax, args.ntheta, args.nphi) == (30, 48, 96):
        # relaxed jammed unit cell (4000 particles) tiled periodically to ~n particles
        t = max(1.0, n / 4000.0)
        r = max(1, int(round(t ** (1.0 / 3.0))))
        best = min(((abs(a * b * c - t), (a, b, c)) for a in (r - 1, r, r + 1) for b in (r - 1, r, r + 1)
                    for c in (r - 1, r, r + 1) if a >= 1 and b >= 1 and c >= 1 and a >= b >= c), key=lambda u: u[0])[1]
        return W.tiled_packing(best)
    return W.config3_packing(n, lmax=args.lmax, grid=(args.ntheta, args.nphi), seed=30)


def cpu_baseline_run(args, pkg, budget_s=20.0, steps=1, warmup=0):
    """Oracle (CPU port) with all host threads on a bounded sample: a smaller packing of the same kind."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    cores = os.cpu_count() or 1
    W = pkg.workloads
    # probe the pair rate on a tiny sample, then size the bounded sample to the time budget
    probe = make_workload(args, pkg, 256)
    o = O.Oracle(threads=cores)
    W.apply(o, probe)
    t0 = time.perf_counter()
    o.compute_forces()
    tp = time.perf_counter() - t0
    rate = o.get_counters()["pair_evals"] / max(tp, 1e-9)
    o.close()
    per_step = budget_s / max(1, steps + warmup)
    n_sample = int(min(args.particles, max(256, rate * per_step / 6.0)))
    cfg = make_workload(args, pkg, n_sample)
    o = O.Oracle(threads=cores)
    W.apply(o, cfg)
    o.compute_forces()       # setup (neighbor list + first force evaluation), untimed like the GPU arm
    for _ in range(warmup):
        o.run(1)
    c0 = o.get_counters()
    t0 = time.perf_counter()
    o.run(steps)
    dt = time.perf_counter() - t0
    c1 = o.get_counters()
    pairs = c1["pair_evals"] - c0["pair_evals"]
    n = len(cfg["x"])
    o.close()
    return dict(value=pairs / dt, unit=UNIT, cores=cores, kind="port",
                sample="%s: %d particles, %d steps, %d pair evals in %.2f s (oracle, OpenMP %d threads); "
                       "builder's CPU restatement, NOT the reference (LAMMPS+MPI unavailable)"
                       % (cfg["name"], n, steps, pairs, dt, cores),
                particle_steps_per_s=n * steps / dt, ms_per_step=1e3 * dt / steps, n_particles=n)


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    pkg = shpkg.load()
    res = cpu_baseline_run(args, pkg, budget_s=120.0, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, res["n_particles"]),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "particle_steps_per_s": res["particle_steps_per_s"], "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(args, n):
    kind = ("mechanically relaxed jammed packing (phi~0.71, compressed under damping by this code; 4000-particle "
            "periodic unit cell tiled)") if args.workload == "relaxed" else "jittered FCC-seeded packing (phi~0.56)"
    return {"workload": "BASELINE configs[2]: ~100k polydisperse-shape SH packing, 8 shape types, l_max=30, periodic; "
                        + kind, "n_particles": int(n), "lmax": args.lmax,
            "quadrature": "%dx%d" % (args.ntheta, args.nphi),
            "l2_policy": "compute-bound FP64 kernel; the per-step working set (atom SoA + pair list + per-pair/per-entry "
                         "result slots, ~150 MB at 100k particles) exceeds the 126 MB L2 and is rewritten every step; "
                         "shape tables (1.8 MB) are L2/shared-memory resident by design"}


def run_graft(args):
    rank, world, local = dist_env()
    pkg = shpkg.load()
    W = pkg.workloads
    import torch  # plumbing: device selection, barriers, max-over-ranks
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the SPHERHARM path has no CPU fallback")
    torch.cuda.set_device(local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- workload.  N = 1: the ~100k-particle packing on one GPU.  N > 1: ONE periodic packing of
    # ~N x 100k particles, spatially decomposed into N bricks (decomp.py): ghost forward exchange over
    # NCCL every step, migration on rebuild steps ("weak" scaling: fixed work per GPU).
    dd = None
    sim = pkg.ShGpu(device=local)
    if use_dist:
        cfg = make_workload(args, pkg, args.particles if args.strong else args.particles * world)
        D = pkg.load_decomp()
        dd = D.DomainDecomposition(sim, cfg, comm_device="cuda")
        n = dd.nlocal
        peak = sim.measure_fp64_peak() if rank == 0 else None
        dd.setup()
        dd.run(args.warmup)
    else:
        cfg = make_workload(args, pkg)
        n = len(cfg["x"])
        W.apply(sim, cfg)
        peak = sim.measure_fp64_peak() if rank == 0 else None
        sim.compute_forces()
        sim.run(args.warmup)
    n_global = len(cfg["x"])
    sim.reset_timers()

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()     # before the barrier: its start-up delay must not sit inside any rank's timed region
    barrier()
    t0 = time.perf_counter()
    nreb = 0
    if dd is not None:
        sim.mark_begin()
        nreb = dd.run(args.steps)
        dev_s = sim.mark_end()
    else:
        sim.run(args.steps)
        dev_s = sim.get_run_time()["last"]
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    cnt = sim.get_counters()
    tim = sim.get_timers()
    st_t, ev_nodes = sim.get_split_times(), sim.get_counter_raw(5)   # per-kernel device times of the timed region
    # a pair that straddles a brick boundary is evaluated by both ranks: count it once (half on each side)
    pairs_local = cnt["pair_evals"] - 0.5 * sim.get_ghost_pair_evals()
    t_all = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
    p_all = torch.tensor([float(pairs_local), float(n * args.steps)], dtype=torch.float64, device="cuda")
    if use_dist:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        dist.all_reduce(p_all, op=dist.ReduceOp.SUM)
    dev_s_max, wall_max = t_all.tolist()
    pairs_total, psteps_total = p_all.tolist()

    # ---- e2e: Pair::compute offload through the C-ABI with pinned HOST buffers, copies inside the timed region
    st = sim.get_atoms(("x", "quat"))           # owned + ghost atoms of this rank
    nall = len(st["x"])
    hx = torch.from_numpy(st["x"]).pin_memory()
    hq = torch.from_numpy(st["quat"]).pin_memory()
    hf = torch.empty((nall, 3), dtype=torch.float64).pin_memory()
    ht = torch.empty((nall, 3), dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 20))
    for _ in range(min(3, args.warmup)):
        sim.put_state(x=hx.data_ptr(), quat=hq.data_ptr()); sim.compute_forces(); sim.get_forces(hf.data_ptr(), ht.data_ptr())
    c0 = sim.get_counters()["pair_evals"]
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        hx[:, 0] += 1e-9  # the host code owns and mutates the positions between calls
        sim.put_state(x=hx.data_ptr(), quat=hq.data_ptr())
        sim.compute_forces()
        sim.get_forces(hf.data_ptr(), ht.data_ptr())
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_pairs = (sim.get_counters()["pair_evals"] - c0) * (pairs_local / max(1.0, float(cnt["pair_evals"])))
    e_all = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    ep_all = torch.tensor([float(e2e_pairs)], dtype=torch.float64, device="cuda")
    if use_dist:
        dist.all_reduce(e_all, op=dist.ReduceOp.MAX)
        dist.all_reduce(ep_all, op=dist.ReduceOp.SUM)

    if rank == 0:
        flops = algorithmic_flops(cnt, args.lmax)
        pair_s = tim["seconds_pair"]
        achieved = flops / max(pair_s, 1e-12) / 1e12
        peak_tf = peak["flops_per_s"] / 1e12
        # FP64 pipe slots: 12/transformed node + I_eval/evaluated node
        T = (args.lmax + 1) * (args.lmax + 2) // 2
        slots = 12.0 * cnt["nodes_transformed"] + (4 * T + 10 * (args.lmax + 1) + 30) * cnt["nodes_evaluated"] + 15.0 * cnt["nodes_inside"]
        # The pair phase is three kernels (cull / evaluate / reduce) plus the fused kernel on deep contacts.  The
        # roofline object describes the FP64-bound one, pair_eval_kernel: algorithmic flops = F_eval(L) per record it
        # evaluated, over its own CUDA-event time.  The whole phase (all four kernels) is reported beside it.
        ev_flops = f_eval(args.lmax) * float(ev_nodes)
        ev_s = max(st_t["eval"], 1e-12)
        phase_achieved = achieved
        achieved = ev_flops / ev_s / 1e12
        kshare = {k: v / max(pair_s, 1e-12) for k, v in st_t.items()}
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            hbm_src = "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        cull_gbs = CULL_DRAM_BYTES_PER_PAIR * cnt["pair_evals"] / max(st_t["cull"], 1e-12) / 1e9
        cull_info = {"bound": "latency / instruction issue (FP32 + integer + gathers from L2), not HBM and not FP64",
                     "hbm_gbs": cull_gbs, "hbm_peak_gbs": hbm_peak, "hbm_frac": cull_gbs / hbm_peak, "hbm_peak_source": hbm_src,
                     "avg_launch_ms": 1e3 * st_t["cull"] / max(1, tim["pair_launches"])}
        roofline = {"bound": "fp64", "kernel": "pair_eval_kernel", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved / peak_tf,
                    # dram__bytes_read.sum + dram__bytes_write.sum of pair_eval_kernel from one ncu --set full capture
                    # (profiles/r01_eval_kernel_ncu_summary.txt) expressed per record and scaled to this run's records
                    # per launch; algorithmic bytes are 33 per record (32 B record in, 1 B flag out)
                    "traffic": EVAL_DRAM_BYTES_PER_RECORD * float(ev_nodes) / max(1, tim["pair_launches"]),
                    "traffic_source": "ncu capture r01, %.1f B/record" % EVAL_DRAM_BYTES_PER_RECORD,
                    "eval_kernel_avg_launch_ms": 1e3 * ev_s / max(1, tim["pair_launches"]),
                    "eval_kernel_flops_per_launch": ev_flops / max(1, tim["pair_launches"]),
                    "pair_phase": {"achieved_tflops": phase_achieved, "frac": phase_achieved / peak_tf,
                                   "kernel_share_of_phase": kshare,
                                   "cull_kernel": cull_info,
                                   "note": "the phase is dominated by pair_cull_kernel, an instruction/latency-bound FP32+integer "
                                           "kernel (window scan, conservative FP32 pre-cull, exact FP64 test on the candidates)"},
                    "peak_source": "K0 DFMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 entry); "
                                   "nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2 TFLOP/s",
                    "frac_of_nominal": achieved / 37.2,
                    "pipe_slot_frac": slots * 2.0 / max(pair_s, 1e-12) / peak["flops_per_s"],
                    "avg_launch_ms": 1e3 * pair_s / max(1, tim["pair_launches"]),
                    "flops_per_launch": flops / max(1, tim["pair_launches"]),
                    "pair_kernel_share_of_step": pair_s / max(dev_s, 1e-12),
                    "evaluated_nodes_per_pair": cnt["nodes_evaluated"] / max(1, cnt["pair_evals"]),
                    "inside_nodes_per_pair": cnt["nodes_inside"] / max(1, cnt["pair_evals"])}
        cpu = cpu_baseline_run(args, pkg, budget_s=args.cpu_budget, steps=1, warmup=0) if (world == 1 and not args.no_cpu) else None
        line = {"metric": METRIC, "value": pairs_total / dev_s_max, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * dev_s_max / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(workload_config(args, n_global), parallelism=("single GPU" if world == 1 else
                               "spatial decomposition %s bricks, NCCL all_to_all ghost exchange every step, %d rebuilds"
                               % ("x".join(str(v) for v in dd.pgrid), nreb))),
                "particle_steps_per_s": psteps_total / dev_s_max,
                "wall_ms_per_step": 1e3 * wall_max / args.steps,
                "neighbor_builds": cnt["neighbor_builds"],
                "clocks": clocks,
                "e2e": {"value": ep_all.item() / e_all.item(), "unit": UNIT, "h2d_bytes_per_step": int(nall * 7 * 8),
                        "d2h_bytes_per_step": int(nall * 6 * 8), "steps": e2e_steps,
                        "path": "sh_put_state(x,quat pinned host) + sh_compute_forces + sh_get_forces(f,torque pinned host)"},
                "gpu_launches": int(cnt["kernel_launches"]),
                "roofline": roofline}
        if cpu:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    sim.close()
    if use_dist:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--particles", type=int, default=100000)
    ap.add_argument("--workload", default="relaxed", choices=["relaxed", "lattice"],
                    help="relaxed: committed jammed unit cell tiled periodically; lattice: jittered FCC sites")
    ap.add_argument("--lmax", type=int, default=30)
    ap.add_argument("--ntheta", type=int, default=48)
    ap.add_argument("--nphi", type=int, default=96)
    ap.add_argument("--strong", action="store_true", help="N>1: keep the TOTAL particle count at --particles (strong scaling)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "graft":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_graft(args)


if __name__ == "__main__":
    sys.exit(main())
