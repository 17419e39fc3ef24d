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
    "two_galaxies_4m_bh": ("two_galaxies", 1 << 22, "bh", 0.01, 0.35, 1e-3),
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


# ------------------------------------------------------------------------------------------ reference (CPU) legs
TARGETS_PER_THREAD = 16     # targets per step and host thread of the bounded Theta = 0 sample (dynamic,1 schedule: all busy)


class RefDirectSampler:
    """The reference's own all-pairs evaluation = its tree walk at Theta = 0 (Octree::ComputeForces, OctreeSearch.h:99-108;
    the compiled, unmodified reference in oracle/_ref/liboracle_ref.so), timed on a bounded sample of targets against ALL
    sources with every host thread this process may use. One step = TARGETS_PER_THREAD x threads targets. Falls back to the
    restated fp32 loop (oracle/nbody_oracle.c) where the reference build is absent. Shared by `--impl reference` and the
    in-line cpu_baseline so that the two report the same quantity."""

    def __init__(self, posm, vel, eps):
        from oracle import oracle as O
        self.O, self.posm, self.eps, self.n = O, posm, eps, posm.shape[0]
        self.kind = "reference" if O.have_ref() else "port"
        self.threads = O.host_threads()
        self.per_step = int(min(self.n, TARGETS_PER_THREAD * self.threads))
        self.k = 0
        self.t_setup = 0.0
        if self.kind == "reference":
            t0 = time.perf_counter()
            self.r = O.RefSim()
            self.r.SetParticles(O.to_aos(posm, vel))
            self.r.ComputeCubeSize()
            self.r.CreateOctree()            # tree build + monopoles + the shipped Theta = 1.0 walk: set-up, not timed
            self.t_setup = time.perf_counter() - t0

    def step(self) -> float:
        i0 = (self.k * self.per_step) % max(self.n - self.per_step, 1)
        self.k += 1
        if self.kind == "reference":
            return self.r.ComputeForces(0.0, i0, i0 + self.per_step, self.threads)
        _, t = self.O.direct_f32(self.posm, eps=self.eps, i0=i0, i1=i0 + self.per_step, nthreads=self.threads, return_time=True)
        return t

    def describe(self) -> str:
        what = ("reference tree walk at Theta=0 (OctreeSearch.h:99-108, eps=0 as shipped)" if self.kind == "reference"
                else "restated fp32 direct loop (oracle/nbody_oracle.c)")
        return (f"{what}: {self.per_step} targets x all {self.n} sources per step, {self.threads} busy OpenMP threads "
                f"(schedule dynamic,1 over targets); tree build ({self.t_setup:.1f} s, single thread) and integrator excluded")

    def close(self):
        if self.kind == "reference":
            self.r.close()


def ref_full_tick(n: int = 1 << 16, ticks: int = 3):
    """Second stated figure: the reference's complete Tick as shipped (OctreeSearch.cpp:21-34: cube size, octree build,
    monopoles, Theta = 1.0 walk, kick-drift) - single thread, which is how the actor runs inside UE."""
    from oracle import oracle as O
    if not O.have_ref():
        return None
    posm, vel = make_ic("plummer", n)
    r = O.RefSim()
    r.SetParticles(O.to_aos(posm, vel))
    r.PhDeltaTime = 1e-3
    r.Tick(1)
    t = r.Tick(ticks) / ticks
    r.close()
    return {"value": 1.0 / t, "unit": "steps/s", "ms_per_step": 1e3 * t, "cores": 1, "N": n,
            "what": "reference AOctreeSearch::Tick exactly as shipped (build + monopoles + Theta=1.0 walk + kick-drift), Plummer"}


def ref_bh_steps(wl, threads: int):
    """Barnes-Hut config on the reference's CPU code: one single-threaded build (as shipped) + its walk at the config's
    Theta on a sample of targets with all host threads, extrapolated to all N targets."""
    from oracle import oracle as O
    icname, n, method, eps, theta, dt = wl
    if not O.have_ref():
        return None
    posm, vel = make_ic(icname, n)
    r = O.RefSim()
    r.SetParticles(O.to_aos(posm, vel))
    r.PhDeltaTime = dt
    r.ComputeCubeSize()
    t0 = time.perf_counter()
    r.CreateOctree()
    t_build = time.perf_counter() - t0
    m = int(min(n, 2048 * threads))
    t = r.ComputeForces(theta, 0, m, threads)
    r.close()
    return {"value": 1.0 / (t_build + t * n / m), "unit": "steps/s", "cores": threads, "kind": "reference",
            "sample": f"reference CreateOctree (single thread, incl. its shipped Theta=1 walk: {t_build:.2f} s) + ComputeForces at "
                      f"Theta={theta} for {m} of {n} targets on {threads} threads ({t:.2f} s), extrapolated to all targets"}


def run_reference(args, wl):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores, same metric and config."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    icname, n, method, eps, theta, dt = wl
    posm, vel = make_ic(icname, n)
    smp = RefDirectSampler(posm, vel, eps)
    times = []
    for k in range(args.warmup + args.steps):
        t = smp.step()
        if k >= args.warmup:
            times.append(t)
    total = float(sum(times))
    value = smp.per_step * float(n) * len(times) / total
    line = {
        "impl": "reference", "metric": "all-pairs interactions/s", "value": value, "unit": "interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, wl, 1),
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": smp.threads, "busy_threads": smp.threads,
                         "kind": smp.kind, "sample": smp.describe()},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    smp.close()
    if not args.no_extras:
        line["full_tick"] = ref_full_tick()
        if n >= (1 << 20):
            line["bh"] = ref_bh_steps(WORKLOADS["plummer_1m_bh"], smp.threads)
    print(json.dumps(line), flush=True)


def workload_config(args, wl, world, name=None):
    icname, n, method, eps, theta, dt = wl
    return {"workload": f"{icname} N={n} {'softened direct sum' if method == 'direct' else 'Barnes-Hut'} + kick-drift",
            "name": name or args.workload, "N": n, "eps": eps, "dt": dt, "G": 1e4, "method": method,
            "theta_reference_convention": theta if method == "bh" else 0.0,
            "theta_conventional": 2 * theta if method == "bh" else 0.0,
            "partition": ("1 GPU" if world == 1 else
                          f"i-rows over {world} GPUs, NCCL all-gather of float4 positions per step" if method == "direct" else
                          f"Morton domain split over {world} GPUs, body migration + locally-essential-tree exchange (NCCL all-to-all-v)"
                          if (getattr(args, "bh_exchange", -1) == 0 or (getattr(args, "bh_exchange", -1) < 0 and n > (1 << 23))) else
                          f"replicated tree, Morton-order slices over {world} GPUs, all-gather of positions + velocities"),
            "l2": "flushed between timed steps (256 MiB memset); sources (16 B/body) are re-read from L2 by design",
            "seed": 1234}


def cpu_baseline(wl, posm, vel, budget_s: float = 15.0):
    """Bounded sample of the reference's CPU code on this host (rank 0, N=1 only)."""
    icname, n, method, eps, theta, dt = wl
    from oracle import oracle as O
    if method != "direct":
        return ref_bh_steps(wl, O.host_threads())
    smp = RefDirectSampler(posm, vel, eps)
    smp.step()
    times, t_all = [], 0.0
    while t_all < budget_s and len(times) < 64:
        times.append(smp.step())
        t_all += times[-1]
    out = {"value": smp.per_step * float(n) * len(times) / t_all, "unit": "interactions/s", "cores": smp.threads,
           "busy_threads": smp.threads, "kind": smp.kind, "sample": smp.describe() + f"; {len(times)} steps"}
    smp.close()
    m = int(min(n, 16 * smp.threads))
    _, tp = O.direct_f32(posm, eps=eps, i0=0, i1=m, nthreads=smp.threads, return_time=True)
    out["port_value"] = m * float(n) / tp
    out["port_sample"] = f"restated fp32 double loop (oracle/nbody_oracle.c), {m} targets x {n} sources, {smp.threads} threads"
    return out


def load_traffic(kernel: str, workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture of this
    workload (profiles/traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(workload, {}).get(kernel)
    except (OSError, ValueError):
        return None


# ------------------------------------------------------------------------------------------ our arm
def run_workload(args, name, ctx, steps, warmup, min_seconds=0.0, with_cpu_baseline=False, e2e=True):
    """One workload on the product path: device-resident timing, end-to-end timing, roofline. Returns the dict of results
    (rank 0) or None."""
    import torch
    import parallelnbody_b200 as P
    from parallelnbody_b200.api import to_particles
    wl = WORKLOADS[name]
    icname, n, method, eps, theta, dt = wl
    rank, local_rank, world, uid_fn, barrier, max_over_ranks, sum_over_ranks, flush_l2, peak_tf, peak_mhz = ctx
    posm, vel = make_ic(icname, n)
    meth = P.METHOD_DIRECT if method == "direct" else P.METHOD_BARNES_HUT
    sim = P.OctreeSearch(method=meth, G=1e4, eps=eps, theta=theta if method == "bh" else 0.0, PhDeltaTime=dt,
                         device=local_rank, rank=rank, world=world, nccl_unique_id=uid_fn(), bh_exchange=args.bh_exchange)
    sim.SetBodies(posm, vel)
    lets = method == "bh" and world > 1 and (args.bh_exchange == 0 or (args.bh_exchange < 0 and n > (1 << 23)))

    # ---- device-resident: value
    for _ in range(warmup):
        sim.Step(dt, 1)
    if min_seconds > 0:   # size the timed region from the warmed-up step time (the same count on every rank)
        est = max_over_ranks(sim.Stats()["ms_last_call"] * 1e-3)
        steps = int(max(steps, min(20000, np.ceil(min_seconds / max(est, 1e-5)))))
    l0 = sim.Stats()["kernel_launches"]
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    wall0 = time.perf_counter()
    ms = {"total": 0.0, "force": 0.0, "build": 0.0, "integrate": 0.0, "comm": 0.0}
    LET_KEYS = ("ms_let_migrate", "ms_let_plan", "ms_let_walk_local", "ms_let_import", "ms_let_walk_let")
    let_ms = {k: 0.0 for k in LET_KEYS}
    interactions = 0.0
    for _ in range(steps):
        flush_l2()
        sim.Step(dt, 1)
        st = sim.Stats()
        ms["total"] += st["ms_last_call"]; ms["force"] += st["ms_force"]; ms["build"] += st["ms_build"]
        ms["integrate"] += st["ms_integrate"]; ms["comm"] += st["ms_comm"]
        interactions += st["interactions"]
        for k in LET_KEYS:
            let_ms[k] += st[k]
    barrier()
    wall = time.perf_counter() - wall0
    clk = clocks.stop() if rank == 0 else None
    launches = sim.Stats()["kernel_launches"] - l0
    ms_total_max = max_over_ranks(ms["total"])
    inter_all = sum_over_ranks(interactions)
    local_pairs = st["n_local"] * float(n)
    phase_max = {k: max_over_ranks(v) / steps for k, v in ms.items() if k != "total"}
    n_local_max, n_local_min = max_over_ranks(float(st["n_local"])), -max_over_ranks(-float(st["n_local"]))
    let_max = max_over_ranks(float(st["let_points"]))
    let_phases = {k[7:]: max_over_ranks(let_ms[k] / steps) for k in LET_KEYS}
    let_phases_min = {k[7:]: -max_over_ranks(-let_ms[k] / steps) for k in LET_KEYS}

    # ---- end to end through the public API with HOST buffers (pinned FParticle arrays in, Tick, FParticle arrays out)
    if not e2e:
        stats = sim.Stats()
        sim.close()
        if rank != 0:
            return None
        return {"metric": "Barnes-Hut steps/s", "value": steps / (ms_total_max * 1e-3), "unit": "steps/s", "n_gpus": world, "steps": steps,
                "ms_per_step": ms_total_max / steps, "config": workload_config(args, wl, world, name),
                "phases_ms_per_step": {k: ms[k] / steps for k in ("force", "build", "integrate", "comm")}, "clocks": clk,
                "interactions_per_body": inter_all / steps / n, "sort_passes": stats["sort_passes"],
                "tree": {"nodes": stats["tree_nodes"], "depth": stats["tree_depth"], "walk_groups": stats["walk_groups"]}}
    aos_in = torch.empty(n * 40, dtype=torch.uint8).pin_memory()
    aos_out = torch.empty(n * 40, dtype=torch.uint8).pin_memory()
    aos_in.numpy().view(P.PARTICLE_DTYPE)[:] = to_particles(posm, vel)
    sim.SetParticlesRaw(aos_in.data_ptr(), n, 40); sim.Tick(); sim.GetParticlesRaw(aos_out.data_ptr(), n, 40)  # warm
    e2e_steps = max(1, min(steps, args.e2e_steps if method == "direct" else max(args.e2e_steps, 20)))
    barrier()
    e0 = time.perf_counter()
    e2e_parts = [0.0, 0.0, 0.0]   # host-clock time inside the three calls of an iteration (each returns synchronised)
    for _ in range(e2e_steps):
        t_it = time.perf_counter()
        sim.SetParticlesRaw(aos_in.data_ptr(), n, 40)     # H2D: this rank's FParticle records
        t_a = time.perf_counter()
        sim.Tick()                                        # OctreeSearch.cpp:21-34
        t_b = time.perf_counter()
        sim.GetParticlesRaw(aos_out.data_ptr(), n, 40)    # D2H: this rank's FParticle records
        t_c = time.perf_counter()
        e2e_parts[0] += t_a - t_it; e2e_parts[1] += t_b - t_a; e2e_parts[2] += t_c - t_b
        if os.environ.get("NBODY_BENCH_TRACE"):
            print(f"[bench rank {rank}] e2e iteration {1e3 * (time.perf_counter() - t_it):.2f} ms", file=sys.stderr)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - e0)
    e2e_parts = [max_over_ranks(x) * 1e3 / e2e_steps for x in e2e_parts]
    stats = sim.Stats()
    sim.close()

    if method == "direct":
        metric, unit = "all-pairs interactions/s", "interactions/s"
        value = float(n) * float(n) * steps / (ms_total_max * 1e-3)
        e2e_value = float(n) * float(n) * e2e_steps / e2e_s
        ach = FLOPS_PER_INTERACTION * local_pairs * steps / (ms["force"] * 1e-3) / 1e12
        roof = {"bound": "fp32", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": load_traffic("direct_packed_kernel", name) if world == 1 else None,
                "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                "algorithmic_bytes": 16.0 * n + 16.0 * st["n_local"] * stats["jsplit"],
                "peak_kind": f"measured FFMA-chain burst on this GPU ({peak_mhz:.0f} MHz); MEASURED_PEAKS.json has no FP32 entry",
                "peak_nominal": FP32_NOMINAL_TFLOPS, "frac_nominal": ach / FP32_NOMINAL_TFLOPS,
                "kernel": "direct_packed_kernel", "flops_per_interaction": FLOPS_PER_INTERACTION,
                "ms_per_launch": ms["force"] / steps, "jsplit": stats["jsplit"], "i_per_thread": stats["i_per_thread"],
                "equal_mass_kernel": bool(stats.get("equal_mass", 0))}
    else:
        metric, unit = "Barnes-Hut steps/s", "steps/s"
        value = steps / (ms_total_max * 1e-3)
        e2e_value = e2e_steps / e2e_s
        ach = FLOPS_PER_INTERACTION * interactions / (ms["force"] * 1e-3) / 1e12
        hbm = measured_hbm_gbs()
        nl = float(st["n_local"]) if lets else float(n)
        build_bytes = bh_build_bytes(nl, stats)
        build_gbs = build_bytes / max(ms["build"] / steps * 1e-3, 1e-9) / 1e9
        roof = {"bound": "fp32", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": load_traffic("bh_walk_group_kernel", name) if world == 1 else None,
                "kernel": "bh_walk_group_kernel", "interactions_per_step": inter_all / steps,
                "interactions_per_body": inter_all / steps / n, "flops_per_interaction": FLOPS_PER_INTERACTION,
                "ms_per_launch": ms["force"] / steps,
                "what": "raw: every (accepted cell or opened-leaf body) x group member the walk evaluates, 20 flop each; the "
                        "group criterion opens more cells than the reference's per-body rule needs (reference_rule below)",
                "build": {"bound": "hbm", "achieved": build_gbs, "peak": hbm, "unit": "GB/s", "frac": build_gbs / hbm,
                          "algorithmic_bytes": build_bytes, "ms": ms["build"] / steps,
                          "peak_kind": "MEASURED_PEAKS.json hbm_gbs" if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else "fallback 6550.7 GB/s (B200_PROFILING.md)",
                          "what": "cube size + Morton keys + radix sort passes + body gather + tree split + monopoles, algorithmic bytes (DESIGN.md section 4)"}}
    if method == "bh" and world == 1 and n <= (1 << 21):
        # the same system walked per body with the reference's own rule (OctreeSearch.h:103, mac = 1, outside any timed region):
        # how many interactions the reference would evaluate, hence the walk's rate in reference-rule-equivalent interactions
        with P.OctreeSearch(method=meth, G=1e4, eps=eps, theta=theta, PhDeltaTime=dt, device=local_rank, mac=1, leaf_size=1) as ref_rule:
            ref_rule.SetBodies(posm, vel)
            ref_rule.CreateOctree()
            per_body = ref_rule.Stats()["interactions"] / n
        roof["reference_rule"] = {"interactions_per_body": per_body, "useful_fraction": per_body / (inter_all / steps / n),
                                  "equivalent_interactions_per_s": per_body * n / (ms["force"] / steps * 1e-3)}
    if rank != 0:
        return None
    res = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_total_max / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, wl, world, name),
        "tflops_20flop": FLOPS_PER_INTERACTION * value / 1e12 if method == "direct" else None,
        "frac_fp32_peak_per_gpu": (FLOPS_PER_INTERACTION * value / 1e12 / world / peak_tf) if method == "direct" else None,
        "phases_ms_per_step": {k: ms[k] / steps for k in ("force", "build", "integrate", "comm")},
        "phases_ms_per_step_max_over_ranks": phase_max,
        "longest_phase": max(phase_max, key=phase_max.get),
        "wall_s_timed_region": wall,
        "clocks": clk,
        # direct sum / domain split: every rank uploads and reads back its own share; replicated Barnes-Hut: every rank
        # uploads all bodies and reads back its share
        "e2e": {"value": e2e_value, "unit": unit,
                "h2d_bytes_per_step": int(n) * 40 * (world if (method == "bh" and world > 1 and not lets) else 1),
                "d2h_bytes_per_step": int(n) * 40,
                "steps": e2e_steps, "api": "OctreeSearch.Particles <- pinned FParticle AoS; Tick(); Particles -> pinned AoS",
                "ms_per_call_max_over_ranks": {"set_particles": e2e_parts[0], "tick": e2e_parts[1], "get_particles": e2e_parts[2]}},
        "gpu_launches": int(launches),
        "roofline": roof,
    }
    if method == "bh":
        res["bodies_per_rank"] = {"min": n_local_min, "max": n_local_max}
        res["let_points_max"] = let_max
        if lets:
            res["domain_split_phases_ms_per_step_max_over_ranks"] = let_phases
            res["domain_split_phases_ms_per_step_min_over_ranks"] = let_phases_min
            res["longest_phase"] = "domain split: " + max(let_phases, key=let_phases.get)
        res["tree"] = {"nodes": stats["tree_nodes"], "depth": stats["tree_depth"], "walk_groups": stats["walk_groups"]}
    if with_cpu_baseline:
        res["cpu_baseline"] = cpu_baseline(wl, posm, vel)
    return res


def measured_hbm_gbs() -> float:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        return 6550.7


def bh_build_bytes(n: float, stats: dict) -> float:
    """Algorithmic bytes of one Barnes-Hut build over n bodies: cube size 16n; keys 16n r + 8n w; each
    radix pass 12n r + 12n w (+ one 8n histogram read of the keys); body gather 4n + 36n r + 36n w; tree split: keys 8n r +
    40 B per node; monopoles 16n r + 48 B per node."""
    passes = float(stats.get("sort_passes", 0)) or 8.0
    nodes = float(stats["tree_nodes"])
    return n * (16 + 24 + 8 + passes * 24 + 76 + 8 + 16) + nodes * (40 + 48)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import parallelnbody_b200 as P

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    from parallelnbody_b200 import launch
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def uid_fn():     # one fresh NCCL id per handle
        return launch.broadcast_unique_id(P.comm_unique_id, dist, device="cuda") if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return launch.reduce_scalar(x, dist, "max", device="cuda") if world > 1 else x

    def sum_over_ranks(x: float) -> float:
        return launch.reduce_scalar(x, dist, "sum", device="cuda") if world > 1 else x

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    peak_tf, peak_mhz = P.measure_fp32_peak(local_rank)   # FFMA-chain burst peak, same process, same GPU
    ctx = (rank, local_rank, world, uid_fn, barrier, max_over_ranks, sum_over_ranks, flush_l2, peak_tf, peak_mhz)
    main_wl = WORKLOADS[args.workload]
    line = run_workload(args, args.workload, ctx, args.steps, args.warmup,
                        min_seconds=2.0 if main_wl[2] == "bh" else 0.0,
                        with_cpu_baseline=(world == 1 and not args.no_cpu_baseline))
    # BASELINE.json's metric has a second half - "BH steps/s": the default run measures it too and attaches it as `bh`
    if args.workload == "plummer_1m_direct" and not args.no_bh:
        bh_name = "plummer_1m_bh" if world == 1 else "two_galaxies_16m_bh"
        bh = run_workload(args, bh_name, ctx, args.steps, max(args.warmup, 3), min_seconds=2.0,
                          with_cpu_baseline=(world == 1 and not args.no_cpu_baseline))
        if rank == 0:
            line["bh"] = bh
        if world == 1 and not args.no_bh_baseline:
            # the multi-GPU Barnes-Hut workload (BASELINE configs[4]) on ONE GPU, device-resident: the denominator of its
            # strong-scaling efficiency at N = 2 / 4 / 8
            b16 = run_workload(args, "two_galaxies_16m_bh", ctx, 20, 3, min_seconds=1.0, with_cpu_baseline=False, e2e=False)
            if rank == 0:
                line["bh_multi_gpu_workload_on_1_gpu"] = b16
    if rank == 0:
        print(json.dumps(line), flush=True)
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
    ap.add_argument("--no-bh", action="store_true", help="skip the Barnes-Hut object of the default run")
    ap.add_argument("--no-bh-baseline", action="store_true", help="1 GPU: skip the 16M-body Barnes-Hut scaling baseline")
    ap.add_argument("--no-extras", action="store_true", help="reference arm: skip the full-Tick and Barnes-Hut figures")
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
    run_ours(args)


if __name__ == "__main__":
    main()
