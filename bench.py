#!/usr/bin/env python
"""bench.py - BASELINE.json's metric for the N-body hot path: all-pairs interactions/s (and % of FP32 peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload plummer_1m_direct|uniform_64k_direct|plummer_1m_bh]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU code (oracle/_ref), same metric and config

A "step" is one AOctreeSearch::Tick (/root/reference/Source/NBody/OctreeSearch.cpp:21-34): one force evaluation over all
N^2 pairs + the kick-drift integration (+ the NCCL position all-gather when N GPUs share the i-rows). Default
workload = the configuration the metric is quoted on (BASELINE.json configs[2]): Plummer sphere, N = 1,048,576,
softened direct sum, which fits one B200; with --gpus N the SAME N is i-partitioned over the ranks (strong scaling).

Printed JSON (one line, rank 0): see the task contract. `value` = N^2 * K / (sum of per-step device times, max over
ranks) with bodies resident in HBM; `e2e` = the same through the public API with HOST buffers (FParticle AoS in
pinned memory copied in, Tick, FParticle AoS copied out, every step); `roofline` = the force kernel against the
FP32 FMA-pipe peak measured in the same process (the path is rsqrt/FMA bound, neither HBM nor tensor bound);
`cpu_baseline` = the reference's CPU code timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

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

FLOPS_PER_INTERACTION = 20.0          # GPU-Gems-3 convention fixed by BASELINE north_star
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.5: 148 SMs x 128 lanes x 2 flop x clocks.max.sm

WORKLOADS = {
    # name: (ic, N, method, eps, theta, dt)
    "plummer_1m_direct": ("plummer", 1 << 20, "direct", 0.01, 0.0, 1e-3),
    "uniform_64k_direct": ("uniform", 1 << 16, "direct", 0.01, 0.0, 1e-3),
    "plummer_4k_direct": ("plummer", 1 << 12, "direct", 0.01, 0.0, 1e-3),
    # Barnes-Hut: theta in the REFERENCE convention (half-width / distance); conventional opening angle = 2 * theta
    "plummer_1m_bh": ("plummer", 1 << 20, "bh", 0.01, 0.25, 1e-3),            # BASELINE configs[3]: theta_conv = 0.5
    "two_galaxies_16m_bh": ("two_galaxies", 1 << 24, "bh", 0.01, 0.35, 1e-3),  # BASELINE configs[4]: theta_conv = 0.7
    "two_galaxies_2m_bh": ("two_galaxies", 1 << 21, "bh", 0.01, 0.35, 1e-3),
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.th.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if r[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        # samples under load = the upper half by power draw (the sampler also sees the idle edges)
        if sm:
            order = np.argsort(pw)[len(pw) // 2:]
            med = float(np.median(np.asarray(sm)[order]))
        else:
            med = None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


def make_ic(name: str, n: int, seed: int = 1234):
    from parallelnbody_b200 import ic
    return ic.make(name, n, seed)


def dist_env():
    from parallelnbody_b200.launch import dist_env as f
    return f()


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, wl):
    """The reference's own CPU implementation of the path, on the host cores. Direct sum = its tree walk at Theta = 0
    (Octree::ComputeForces, OctreeSearch.h:99-108 - the only all-pairs evaluation the reference has); the compiled,
    unmodified reference lives in oracle/_ref/liboracle_ref.so (oracle/Makefile)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import oracle as O
    icname, n, method, eps, theta, dt = wl
    posm, vel = make_ic(icname, n)
    kind = "reference" if O.have_ref() else "port"
    threads = O.ref_max_threads() if kind == "reference" else O.max_threads()
    per_step = 4 * threads          # targets per step: ~1 s of host work per step at N = 1M
    times = []
    if kind == "reference":
        r = O.RefSim()
        r.SetParticles(O.to_aos(posm, vel))
        r.ComputeCubeSize()
        r.CreateOctree()             # tree build + the shipped Theta = 1.0 walk: set-up, not timed
        th = 0.0 if method == "direct" else theta
        for k in range(args.warmup + args.steps):
            i0 = (k * per_step) % max(n - per_step, 1)
            t = r.ComputeForces(th, i0, min(n, i0 + per_step), threads)
            if k >= args.warmup:
                times.append(t)
        sample = (f"reference tree walk at Theta={th:g} (OctreeSearch.h:99-108, eps=0 as shipped) for {per_step} targets "
                  f"x all {n} sources per step, {threads} OpenMP threads over targets; tree build excluded")
        r.close()
    else:
        for k in range(args.warmup + args.steps):
            i0 = (k * per_step) % max(n - per_step, 1)
            _, t = O.direct_f32(posm, eps=eps, i0=i0, i1=min(n, i0 + per_step), nthreads=threads, return_time=True)
            if k >= args.warmup:
                times.append(t)
        sample = f"restated fp32 direct loop for {per_step} targets x {n} sources per step, {threads} threads"
    total = float(sum(times))
    value = per_step * float(n) * len(times) / total
    line = {
        "impl": "reference", "metric": "all-pairs interactions/s", "value": value, "unit": "interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, wl, 1),
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, wl, world):
    icname, n, method, eps, theta, dt = wl
    return {"workload": f"{icname} N={n} {'softened direct sum' if method == 'direct' else 'Barnes-Hut'} + kick-drift",
            "name": args.workload, "N": n, "eps": eps, "dt": dt, "G": 1e4, "method": method,
            "theta_reference_convention": theta if method == "bh" else 0.0,
            "partition": ("1 GPU" if world == 1 else
                          f"i-rows over {world} GPUs, NCCL all-gather of float4 positions per step" if method == "direct" else
                          f"Morton domain split over {world} GPUs, body migration + locally-essential-tree exchange (NCCL all-to-all-v)"
                          if (getattr(args, "bh_exchange", -1) == 0 or (getattr(args, "bh_exchange", -1) < 0 and n > (1 << 23))) else
                          f"replicated tree, Morton-order slices over {world} GPUs, all-gather of positions + velocities"),
            "l2": "flushed between timed steps (256 MiB memset); sources (16 B/body) are re-read from L2 by design",
            "seed": 1234}


def cpu_baseline(wl, budget_s: float = 20.0):
    """Bounded sample of the reference's CPU code on this host (rank 0, N=1 only)."""
    from oracle import oracle as O
    icname, n, method, eps, theta, dt = wl
    posm, vel = make_ic(icname, n)
    if O.have_ref() and method == "direct":
        threads = O.ref_max_threads()
        r = O.RefSim()
        r.SetParticles(O.to_aos(posm, vel))
        r.ComputeCubeSize()
        r.CreateOctree()
        m = 2 * threads
        t = r.ComputeForces(0.0, 0, m, threads)
        # scale the sample to ~budget/2 seconds
        m2 = int(min(n, max(m, m * (0.5 * budget_s) / max(t, 1e-3))))
        t2 = r.ComputeForces(0.0, 0, m2, threads)
        r.close()
        _, tp = O.direct_f32(posm, eps=eps, i0=0, i1=min(n, 16 * threads), nthreads=threads, return_time=True)
        return {"value": m2 * float(n) / t2, "unit": "interactions/s", "cores": threads, "kind": "reference",
                "sample": f"reference Theta=0 tree walk (OctreeSearch.h:99-108), {m2} targets x {n} sources, {threads} threads, tree build excluded",
                "port_value": min(n, 16 * threads) * float(n) / tp,
                "port_sample": f"restated fp32 double loop (oracle/nbody_oracle.c), {min(n, 16 * threads)} targets x {n} sources, {threads} threads"}
    if O.have_ref():
        threads = 1
        r = O.RefSim()
        r.SetParticles(O.to_aos(posm, vel))
        r.PhDeltaTime = dt
        r.ComputeCubeSize()
        t0 = time.perf_counter()
        r.CreateOctree()
        t_build = time.perf_counter() - t0
        m = int(min(n, 20000))
        t = r.ComputeForces(theta, 0, m, O.ref_max_threads())
        r.close()
        return {"value": 1.0 / (t_build + t * n / m / 1.0), "unit": "steps/s", "cores": O.ref_max_threads(), "kind": "reference",
                "sample": f"reference build (single thread, incl. shipped Theta=1 walk) + Theta={theta} walk extrapolated from {m} targets"}
    threads = O.max_threads()
    m = 16 * threads
    _, tp = O.direct_f32(posm, eps=eps, i0=0, i1=min(n, m), nthreads=threads, return_time=True)
    return {"value": min(n, m) * float(n) / tp, "unit": "interactions/s", "cores": threads, "kind": "port",
            "sample": f"restated fp32 double loop, {min(n, m)} targets x {n} sources"}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import parallelnbody_b200 as P

    rank, local_rank, world = dist_env()
    icname, n, method, eps, theta, dt = wl
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    from parallelnbody_b200 import launch
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = launch.broadcast_unique_id(P.comm_unique_id, dist, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return launch.reduce_scalar(x, dist, "max", device="cuda") if world > 1 else x

    def sum_over_ranks(x: float) -> float:
        return launch.reduce_scalar(x, dist, "sum", device="cuda") if world > 1 else x

    posm, vel = make_ic(icname, n)
    meth = P.METHOD_DIRECT if method == "direct" else P.METHOD_BARNES_HUT
    sim = P.OctreeSearch(method=meth, G=1e4, eps=eps, theta=theta if method == "bh" else 0.0, PhDeltaTime=dt,
                         device=local_rank, rank=rank, world=world, nccl_unique_id=uid, bh_exchange=args.bh_exchange)
    sim.SetBodies(posm, vel)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    # ---- device-resident: value
    for _ in range(args.warmup):
        sim.Step(dt, 1)
    peak_tf, peak_mhz = P.measure_fp32_peak(local_rank)   # FFMA-chain burst peak, same process, same GPU
    l0 = sim.Stats()["kernel_launches"]
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    wall0 = time.perf_counter()
    ms_total = ms_force = ms_build = ms_integ = ms_comm = 0.0
    interactions = 0.0
    for _ in range(args.steps):
        flush_l2()
        sim.Step(dt, 1)
        st = sim.Stats()
        ms_total += st["ms_last_call"]; ms_force += st["ms_force"]; ms_build += st["ms_build"]
        ms_integ += st["ms_integrate"]; ms_comm += st["ms_comm"]
        interactions += st["interactions"]
    barrier()
    wall = time.perf_counter() - wall0
    clk = clocks.stop() if rank == 0 else None
    launches = sim.Stats()["kernel_launches"] - l0
    ms_total_max = max_over_ranks(ms_total)
    inter_all = sum_over_ranks(interactions)
    pairs_all = float(n) * float(n) * args.steps
    local_pairs = st["n_local"] * float(n)

    # ---- end to end through the public API with host buffers
    aos_in = torch.empty(n * 40, dtype=torch.uint8).pin_memory()
    aos_out = torch.empty(n * 40, dtype=torch.uint8).pin_memory()
    from parallelnbody_b200.api import to_particles
    aos_in.numpy().view(P.PARTICLE_DTYPE)[:] = to_particles(posm, vel)
    sim.SetParticlesRaw(aos_in.data_ptr(), n, 40); sim.Tick(); sim.GetParticlesRaw(aos_out.data_ptr(), n, 40)  # warm
    barrier()
    e0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(e2e_steps):
        t_it = time.perf_counter()
        sim.SetParticlesRaw(aos_in.data_ptr(), n, 40)     # H2D: this rank's FParticle records
        sim.Tick()                                        # OctreeSearch.cpp:21-34
        sim.GetParticlesRaw(aos_out.data_ptr(), n, 40)    # D2H: this rank's FParticle records
        if os.environ.get("NBODY_BENCH_TRACE"):
            print(f"[bench rank {rank}] e2e iteration {1e3 * (time.perf_counter() - t_it):.2f} ms", file=sys.stderr)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - e0)
    n_local = st["n_local"]

    if method == "direct":
        metric, unit = "all-pairs interactions/s", "interactions/s"
        value = pairs_all / (ms_total_max * 1e-3)
        e2e_value = float(n) * float(n) * e2e_steps / e2e_s
        ach = FLOPS_PER_INTERACTION * local_pairs * args.steps / (ms_force * 1e-3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum of one K1 launch from the committed ncu --set full capture of this
        # workload (profiles/r1_direct_kernel_ncu_summary.md): 40.2 MB read + 223.4 MB written, against 16.8 MB of sources
        # + 268 MB of j-split partials; irrelevant to the bound (0.6 GB/s-class traffic in a 413 ms kernel)
        traffic = 263.66e6 if (args.workload == "plummer_1m_direct" and world == 1) else None
        roof = {"bound": "fp32", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": traffic, "peak_kind": f"measured FFMA-chain burst on this GPU ({peak_mhz:.0f} MHz); MEASURED_PEAKS.json has no FP32 entry",
                "peak_nominal": FP32_NOMINAL_TFLOPS, "frac_nominal": ach / FP32_NOMINAL_TFLOPS,
                "kernel": "direct_packed_kernel", "flops_per_interaction": FLOPS_PER_INTERACTION,
                "ms_per_launch": ms_force / args.steps}
    else:
        metric, unit = "Barnes-Hut steps/s", "steps/s"
        value = args.steps / (ms_total_max * 1e-3)
        e2e_value = e2e_steps / e2e_s
        ach = FLOPS_PER_INTERACTION * interactions / (ms_force * 1e-3) / 1e12
        roof = {"bound": "fp32", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": None, "kernel": "bh_walk_group_kernel", "interactions_per_step": inter_all / args.steps,
                "ms_per_launch": ms_force / args.steps, "ms_build_per_step": ms_build / args.steps}

    lets = method == "bh" and (args.bh_exchange == 0 or (args.bh_exchange < 0 and n > (1 << 23)))
    if rank != 0:
        sim.close()
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, wl, world),
        "tflops_20flop": FLOPS_PER_INTERACTION * value / 1e12 if method == "direct" else None,
        "frac_fp32_peak_per_gpu": (FLOPS_PER_INTERACTION * value / 1e12 / world / peak_tf) if method == "direct" else None,
        "phases_ms_per_step": {"force": ms_force / args.steps, "build": ms_build / args.steps,
                               "integrate": ms_integ / args.steps, "comm": ms_comm / args.steps},
        "wall_s_timed_region": wall,
        "clocks": clk,
        # direct sum / domain split: every rank uploads and reads back its own share; replicated Barnes-Hut: every rank
        # uploads all bodies and reads back its share
        "e2e": {"value": e2e_value, "unit": unit,
                "h2d_bytes_per_step": int(n) * 40 * (world if (method == "bh" and world > 1 and not lets) else 1),
                "d2h_bytes_per_step": int(n) * 40,
                "steps": e2e_steps, "api": "OctreeSearch.Particles <- pinned FParticle AoS; Tick(); Particles -> pinned AoS"},
        "gpu_launches": int(launches),
        "roofline": roof,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wl)
    print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="plummer_1m_direct", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--bh-exchange", type=int, default=-1, choices=[-1, 0, 1],
                    help="multi-GPU Barnes-Hut: 0 = Morton domain split + LET exchange, 1 = replicated tree")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("bench.py: note: fewer than 3 warm-up steps", file=sys.stderr)
    wl = WORKLOADS[args.workload]
    rank, _, world = dist_env()
    if args.impl == "reference":
        run_reference(args, wl)
        return
    if world != args.gpus and not (world == 1 and args.gpus == 1):
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
            raise SystemExit(subprocess.call(cmd))
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    run_ours(args, wl)


if __name__ == "__main__":
    main()
