#!/usr/bin/env python
"""Benchmark of the COVID-ABS per-iteration agent step (BASELINE.json metric: agent-updates/s at 10M agents).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--agents A]

One "step" = one Simulation.execute() + the statistics row batch_experiment collects
(covid_abs/abs.py:232-314) over the whole population.  Workload at N=1: BASELINE.json configs[2]
("Simulation 10M agents, Conditional Lockdown + masks, synthetic uniform placement, 1 B200", SURVEY.md 8d
config 3).  The working set (2 x 24 B x 1e7 agents + the cell table) is far larger than the 126 MB L2, so
nothing is cache-resident between steps ("inputs larger than L2").

`--impl reference` times the CPU restatement of the same step (oracle/, numpy) on the box's host cores.
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

B_ALG = 147.0  # algorithmic bytes per agent-update of the whole step, the roofline constant
# algorithmic bytes per agent of each kernel of THIS implementation (DESIGN.md "Kernels"): columns the kernel
# must read/write once, c = cells per agent (1.25 at config 3)
KERNEL_BYTES = {
    "k_move_count<true>": lambda c: 16 + 8 + 8,               # R id,x,y,attr; W displacement,rank; histogram RMW
    "k_scan_cells": lambda c: 8 * c,                          # R + W the cell table
    "k_update_scatter<true>": lambda c: 32 + 4 + 24 + 4 * c,  # R record,wealth,aux; R cell_start; W record,wealth; zero old table
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def workload_kwargs(n_agents):
    """density 0.2 agents/unit^2 (the reference default 20/100), 1 % infected, 1 % immune,
    amplitudes 5, conditional lockdown (run_batch.py:628-643), masks (run_batch.py:178-187: distance .05, rate .1)."""
    side = int(round((n_agents / 0.2) ** 0.5))
    return dict(population_size=n_agents, length=side, height=side, initial_infected_perc=0.01,
                initial_immune_perc=0.01, contagion_distance=0.05, contagion_rate=0.1,
                total_wealth=1e4 * n_agents / 20, critical_limit=0.6)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '', 1).isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '', 1).isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(sample_agents, steps, kw_full):
    """The CPU restatement (oracle/sync_oracle.py, numpy, 1 thread) on a bounded sample of the same workload:
    same density, rates and triggers, `sample_agents` agents."""
    from oracle.sync_oracle import SyncSimulation, synthetic_population
    from oracle.seq_oracle import S, I, R
    side = int(round((sample_agents / 0.2) ** 0.5))
    pop, tw = synthetic_population(sample_agents, side, side, seed=1, infected=0.01, immune=0.01)
    sim = SyncSimulation(seed=1, population_size=sample_agents, length=side, height=side,
                         contagion_distance=kw_full["contagion_distance"], contagion_rate=kw_full["contagion_rate"],
                         total_wealth=tw, amplitudes={S: 5, R: 5, I: 5})
    sim.set_population(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    sim.execute()  # warm-up (imports, allocator)
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.execute()
    dt = time.perf_counter() - t0
    return sample_agents * steps / dt, dt


def _ref_worker(args):
    sample_agents, steps, kw = args
    os.environ["OMP_NUM_THREADS"] = "1"
    return cpu_baseline(sample_agents, steps, kw)


def run_reference(args, rank, world):
    """Reference arm: the CPU restatement on all host cores (one independent sample per process)."""
    if rank != 0:
        return
    import multiprocessing as mp
    kw = workload_kwargs(args.agents)
    cores = os.cpu_count() or 1
    sample = args.ref_sample
    steps_total = args.steps + args.warmup
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_ref_worker, [(sample, steps_total, kw)] * cores)
    wall = time.perf_counter() - t0
    value = sum(r[0] for r in res)   # aggregate agent-updates/s over the cores, each timing its own steps
    line = {
        "impl": "reference", "metric": "agent-updates/sec", "value": value, "unit": "agent-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * max(r[1] for r in res) / steps_total, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32/u8/f64", "data": "synthetic",
        "config": {"workload": "config3: Simulation, conditional lockdown + masks, synthetic uniform placement; "
                               "bounded sample of {} agents per core at the same density".format(sample),
                   "agents": args.agents, "sample_agents_per_core": sample},
        "cpu_baseline": {"value": value, "unit": "agent-updates/s", "cores": cores, "kind": "port",
                         "sample": "{} processes x {} agents x {} steps of oracle/sync_oracle.py (numpy + k-d tree), "
                                   "wall {:.1f} s".format(cores, sample, steps_total, wall)},
        "e2e": {"value": value, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def make_population(n, x_lo, x_hi, side, seed, id0, total_wealth_per_agent):
    """Synthetic uniform placement of n agents with x in [x_lo, x_hi] (one slab)."""
    rng = np.random.default_rng(seed)
    x = rng.integers(x_lo, x_hi + 1, size=n, dtype=np.int32)
    y = rng.integers(0, side + 1, size=n, dtype=np.int32)
    age = (rng.beta(2, 5, size=n) * 100).astype(np.uint8)
    stratum = rng.integers(0, 5, size=n, dtype=np.uint8)
    status = np.zeros(n, dtype=np.uint8)
    status[:int(n * 0.01)] = 1
    status[int(n * 0.01):int(n * 0.02)] = 2
    from covid19_agentbasedsimulation_b200.common import lorenz_curve
    wealth = np.zeros(n)
    for q in range(5):   # abs.py:134-139 with total_wealth proportional to the population
        sel = (stratum == q) & (age >= 18)
        wealth[sel] = lorenz_curve[q] * total_wealth_per_agent * n / max(1.0, float(sel.sum()))
    return {"x": x, "y": y, "status": status, "severity": np.ones(n, dtype=np.uint8),
            "infected_time": np.zeros(n, dtype=np.uint8), "age": age, "stratum": stratum, "wealth": wealth,
            "id": (np.arange(n, dtype=np.int64) + id0).astype(np.uint32)}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from covid19_agentbasedsimulation_b200 import build
    build.build_library()
    from covid19_agentbasedsimulation_b200.abs import Simulation, Trigger
    from covid19_agentbasedsimulation_b200.agents import Status
    from covid19_agentbasedsimulation_b200.slab import SlabSimulation, slab_bounds

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.agents                      # per GPU (weak scaling: the domain grows with the number of GPUs)
    kw = workload_kwargs(n * world)
    side = kw["length"]
    low = {Status.Susceptible: 1.5, Status.Recovered_Immune: 1.5, Status.Infected: 1.5}
    high = {Status.Susceptible: 5, Status.Recovered_Immune: 5, Status.Infected: 5}
    triggers = [Trigger('Infected', '>=', .2, 'amplitudes', low), Trigger('Infected', '<=', .05, 'amplitudes', high)]
    common = dict(seed=0, device=local_rank, triggers_simulation=triggers, max_cells=args.max_cells,
                  history_capacity=max(4096, args.steps + args.warmup + 64), **kw)
    if world == 1:
        sim = Simulation(**common)
        x_lo, x_hi = 0, side
    else:
        sim = SlabSimulation(distributed=True, slab_devices=[local_rank], transport=args.transport, **common)
        lo, hi = slab_bounds(side, world)[rank]
        x_lo, x_hi = max(lo, 0), min(hi - 1, side)
    # this rank's agents in pinned host memory (also the source of the end-to-end leg)
    cols = make_population(n, x_lo, x_hi, side, seed=rank, id0=rank * n, total_wealth_per_agent=1e4 / 20)
    if world == 1:
        del cols["id"]       # ids are the list positions 0..n-1 (abs.py:15): nothing to send
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in cols.items()}
    ptrs = {k: t.data_ptr() for k, t in pinned.items()}
    h2d_bytes = sum(t.numel() * t.element_size() for t in pinned.values())

    if world == 1:
        sim._ensure_handle(n)
        sim._push_params()
        h = sim._handle

        def upload():
            h.upload_pointers(n, ptrs)

        def step(k):
            h.step(k)
    else:
        sim.population_size = n * world
        sim._create([n])
        h = sim._handle

        def upload():
            h.upload_pointers(n, ptrs)
            sim._build(advance=False)

        def step(k):
            if args.transport == "p2p":
                h.slab_run(k)
            else:
                for _ in range(k):
                    sim._build(advance=True)

    def timed(fn):
        """Device time of fn() on the stream the library runs on, ms."""
        if world == 1 or args.transport == "p2p":   # the library's own stream
            h.event_record(0)
            fn()
            h.event_record(1)
            return h.event_elapsed_ms(0, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    upload()
    h.sync()

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps
    step(args.warmup)
    h.sync()
    launches0 = h.info()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ms = timed(lambda: step(args.steps))
    barrier()
    clocks = sampler.stop()
    launches = h.info()["kernel_launches"] - launches0
    ms = max_over_ranks(ms)
    value = world * n * args.steps / (ms * 1e-3)
    stats_last = h.stats()[0]

    # ---- per-kernel device times (events around every launch), same steady state
    h.profile_enable(True)
    step(args.steps)
    prof = h.profile_read()
    h.profile_enable(False)
    info = h.info()
    cells_per_agent = info["cells_x"] * info["cells_y"] / float(max(1, info["n"]))
    kernels = {}
    for name, (cnt, tot) in prof.items():
        avg = tot / cnt
        entry = {"launches_per_step": cnt / args.steps, "avg_ms": avg}
        if name in KERNEL_BYTES:
            entry["alg_bytes_per_agent"] = KERNEL_BYTES[name](cells_per_agent)
            entry["achieved_gbs"] = entry["alg_bytes_per_agent"] * n / (avg * 1e-3) / 1e9
        kernels[name] = entry
    step_ms_prof = sum(tot for _, tot in prof.values()) / args.steps
    top = max(kernels, key=lambda k: kernels[k]["avg_ms"] * kernels[k]["launches_per_step"])
    peak, peak_kind = peaks()
    roofline = {"bound": "hbm", "kernel": top, "achieved": kernels[top].get("achieved_gbs"), "peak": peak,
                "peak_source": peak_kind, "unit": "GB/s",
                "frac": (kernels[top].get("achieved_gbs") or 0.0) / peak, "traffic": None,
                "step_share": kernels[top]["avg_ms"] * kernels[top]["launches_per_step"] / step_ms_prof,
                "whole_step": {"alg_bytes_per_agent_update": B_ALG,
                               "achieved": value / world * B_ALG / 1e9,
                               "frac": value / world * B_ALG / 1e9 / peak},
                "kernels": kernels}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            roofline["traffic"] = json.load(f).get(top)

    # ---- end to end through the C ABI with HOST buffers: every step uploads the population from pinned
    #      host memory (set_population, abs.py:65-69), executes, and reads the statistics back (abs.py:291-314)
    e2e_steps = max(3, min(args.steps, args.e2e_steps))

    def e2e_once():
        upload()
        step(1)
        h.stats()

    e2e_once()
    barrier()
    e2e_ms = max_over_ranks(timed(lambda: [e2e_once() for _ in range(e2e_steps)]))
    barrier()
    e2e_value = world * n * e2e_steps / (e2e_ms * 1e-3)

    other = None
    if rank == 0 and world == 1 and not args.no_other_configs:
        # the GraphSimulation configs of BASELINE.json (configs[1], configs[4]-style), reported next to the headline
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import graph_bench
            other = {"config1_simulation_defaults": graph_bench.config1(),
                     "config2_graph_1k_lockdown": graph_bench.config2(),
                     "config5_style_ensemble": graph_bench.ensemble(2000, 32, 120)}
        except Exception as ex:  # the headline line must still be printed
            other = {"error": repr(ex)}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            base_value, base_dt = cpu_baseline(args.cpu_sample, 5, kw)
            cpu = {"value": base_value, "unit": "agent-updates/s", "cores": 1, "kind": "port",
                   "sample": "{} agents x 5 steps of oracle/sync_oracle.py (numpy, 1 thread) at the same density, "
                             "{:.1f} s".format(args.cpu_sample, base_dt)}
        line = {
            "metric": "agent-updates/sec", "value": value, "unit": "agent-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32/u8/f64", "data": "synthetic",
            "config": {"workload": ("config3: Simulation 1e7 agents, conditional lockdown + masks, synthetic uniform "
                                    "placement (BASELINE.json configs[2])" if world == 1 else
                                    "config4-style: one Simulation of {} agents slab-partitioned over {} GPUs, migrant + "
                                    "ghost exchange and statistics all-gather over NCCL every step, conditional lockdown "
                                    "+ masks (BASELINE.json configs[3] at {} agents/GPU)".format(n * world, world, n)),
                       "agents_per_gpu": n, "agents": n * world, "length": side, "height": side,
                       "contagion_distance": 0.05, "contagion_rate": 0.1,
                       "cache": "working set {:.0f} MB per GPU >> 126 MB L2 (inputs larger than L2)"
                       .format((56.0 * n + 8 * info["cells_x"] * info["cells_y"]) / 1e6),
                       "cell_edge": info["cell_edge"], "cells": info["cells_x"] * info["cells_y"],
                       "multi_gpu": ("x-slabs, transport=" + args.transport) if world > 1 else "single device"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "agent-updates/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 12 * 8 + 7 * 8 + 4, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "final_stats": {"Susceptible": float(stats_last[0]), "Infected": float(stats_last[1])},
            "other_configs": other,
        }
        print(json.dumps(line))
    if world > 1:
        if args.transport == "p2p":
            sim.close()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents", type=int, default=10_000_000, help="agents per GPU")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    ap.add_argument("--ref-sample", type=int, default=500_000)
    ap.add_argument("--transport", default="p2p", choices=["p2p", "collective"],
                    help="slab exchange at N > 1: NVLink peer stores + flags, or torch.distributed collectives")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the GraphSimulation side measurements")
    ap.add_argument("--max-cells", type=int, default=0, help="cap on cell-list cells (0 = library default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
