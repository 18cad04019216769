#!/usr/bin/env python
"""Benchmark of the COVID-ABS per-iteration agent step (BASELINE.json metric: agent-updates/s at 10M agents).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload simulation|ensemble]

One "step" = one Simulation.execute() + the statistics row batch_experiment collects (covid_abs/abs.py:232-314) over
the whole population.  Workload at N=1: BASELINE.json configs[2] ("Simulation 10M agents, Conditional Lockdown + masks,
synthetic uniform placement, 1 B200").  At N>1 the same population density on N x-slabs, 10^7
agents per GPU (weak scaling; BASELINE.json configs[3] style), plus, under `other_configs`, configs[3] literally
(10^8 agents over the N GPUs).  The working set (2 x 24 B x 1e7 agents + the cell table) is far larger than the
126 MB L2, so nothing is cache-resident between steps ("inputs larger than L2").

`--workload ensemble` is BASELINE.json configs[4]: 7 paper scenarios x seeds x 10^5 persons of GraphSimulation,
whole simulations sharded over the GPUs with no data-path communication.

`--impl reference` times the CPU side: the unmodified reference's own loop (baseline/_ref, single-threaded Python) at
the sizes it can run, and the CPU restatement of the same step (oracle/, numpy) on all host cores at a bounded sample
of the workload -- the latter is the ratio arm (the former cannot run 10^7 agents: 5e13 pair tests per step).
"""
import argparse
import hashlib
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
# must read/write once, c = cells per agent
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


def csrc_digest():
    """Hash of the kernel sources: profiles/traffic.json is only quoted for the build it was captured on."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "covid19_agentbasedsimulation_b200", "csrc")
    for name in sorted(os.listdir(d)):
        with open(os.path.join(d, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def workload_kwargs(n_agents, masked=True):
    """density 0.2 agents/unit^2 (the reference default 20/100), 1 % infected, 1 % immune,
    amplitudes 5, conditional lockdown (run_batch.py:628-643), masks (run_batch.py:178-187: distance .05, rate .1);
    masked=False is the variant the survey asks to report next to it (r = 1, rate .9)."""
    side = int(round((n_agents / 0.2) ** 0.5))
    return dict(population_size=n_agents, length=side, height=side, initial_infected_perc=0.01,
                initial_immune_perc=0.01, contagion_distance=0.05 if masked else 1.0,
                contagion_rate=0.1 if masked else 0.9, total_wealth=1e4 * n_agents / 20, critical_limit=0.6)


class ClockSampler(object):
    """SM clock and throttle reasons DURING the timed region: NVML polled from a thread every few milliseconds
    (nvidia-smi -lms cannot resolve a 60 ms region); nvidia-smi as the fallback."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop, self._thread, self._nvml = threading.Event(), None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._dev = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._dev, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
        self._thread = threading.Thread(target=self._poll if self._nvml else self._poll_smi, daemon=True)
        self._thread.start()

    def _poll(self):
        nv = self._nvml
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._dev, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._dev)
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def _poll_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, universal_newlines=True,
                                     timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                return

    def mark(self):
        """Index of the next sample: brackets the timed region inside a longer sampling window."""
        return len(self.samples)

    def stop(self, first=0, last=None):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=6)
        window = self.samples[first:last] or self.samples
        return {"sm_mhz": float(np.median(window)) if window else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(window),
                "source": "nvml" if self._nvml else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------------- CPU
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


def reference_path(budget):
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import reference_timing
        return reference_timing.reference_path(budget)
    except Exception as ex:
        return {"available": False, "why": repr(ex)}


def run_reference(args, rank, world):
    """Reference arm: (i) the unmodified reference's own loop at the sizes it can run (single-threaded Python),
    (ii) the CPU restatement of the step on all host cores, one bounded sample of the workload per core -- the
    line's `value`."""
    if rank != 0:
        return
    import multiprocessing as mp
    kw = workload_kwargs(args.agents)
    cores = os.cpu_count() or 1
    sample = args.ref_sample
    steps_total = min(args.steps + args.warmup, args.ref_max_steps)
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_ref_worker, [(sample, steps_total, kw)] * cores)
    wall = time.perf_counter() - t0
    value = sum(r[0] for r in res)   # aggregate agent-updates/s over the cores, each timing its own steps
    ref = reference_path("small")
    line = {
        "impl": "reference", "metric": "agent-updates/sec", "value": value, "unit": "agent-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * max(r[1] for r in res) / steps_total, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32/u8/f64", "data": "synthetic",
        "config": {"workload": "config3: Simulation, conditional lockdown + masks, synthetic uniform placement; "
                               "bounded sample of {} agents per core at the same density, {} steps per core".format(
                                   sample, steps_total),
                   "agents": args.agents, "sample_agents_per_core": sample},
        "cpu_baseline": {"value": value, "unit": "agent-updates/s", "cores": cores, "kind": "port",
                         "sample": "{} processes x {} agents x {} steps of oracle/sync_oracle.py (numpy + k-d tree), "
                                   "wall {:.1f} s; the unmodified reference (kind 'reference') is timed beside it in "
                                   "`reference_path` at the sizes its O(N^2) Python loop can run".format(
                                       cores, sample, steps_total, wall),
                         "reference_path": ref},
        "e2e": {"value": value, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------- GPU
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


def lockdown_triggers():
    from covid19_agentbasedsimulation_b200.abs import Trigger
    from covid19_agentbasedsimulation_b200.agents import Status
    low = {Status.Susceptible: 1.5, Status.Recovered_Immune: 1.5, Status.Infected: 1.5}
    high = {Status.Susceptible: 5, Status.Recovered_Immune: 5, Status.Infected: 5}
    return [Trigger('Infected', '>=', .2, 'amplitudes', low), Trigger('Infected', '<=', .05, 'amplitudes', high)]


def timed_on(h, fn):
    """Device time of fn() on the stream the library runs on, ms."""
    h.event_record(0)
    fn()
    h.event_record(1)
    return h.event_elapsed_ms(0, 1)


def variant_line(n, cols, masked, schedule, steps, device):
    """agent-updates/s of one more configuration of the same population (other_configs): returns mean and worst step."""
    from covid19_agentbasedsimulation_b200.abs import Simulation
    kw = workload_kwargs(n, masked=masked)
    sim = Simulation(seed=0, device=device, triggers_simulation=lockdown_triggers(), schedule=schedule,
                     history_capacity=steps + 64, **kw)
    sim.set_population_arrays(cols["x"], cols["y"], cols["status"], cols["age"], cols["stratum"], cols["wealth"])
    h = sim._handle
    h.step(4)
    h.sync()
    per_step = []
    if masked:      # steady state: one enqueue
        ms = timed_on(h, lambda: h.step(steps))
        per_step = None
    else:           # the epidemic sweeps through: the cost of a step follows it, so time every step
        ms = 0.0
        for _ in range(steps):
            t = timed_on(h, lambda: h.step(1))
            per_step.append(t)
            ms += t
    st = h.stats()[0]
    out = {"config": "Simulation {} agents, {} (distance {}, rate {}), schedule={}, {} steps".format(
               n, "masks" if masked else "no masks", kw["contagion_distance"], kw["contagion_rate"], schedule, steps),
           "ms_per_step": ms / steps, "agent_updates_per_s": n * steps / (ms * 1e-3),
           "final_infected": float(st[1]), "final_susceptible": float(st[0])}
    if per_step:
        out["worst_step_ms"] = max(per_step)
    sim._handle.close()
    return out


def slab_invariance(local_rank, world, transport, n_total=300000, steps=12):
    """A small population stepped on `world` slabs (one per rank, the transport of the headline run) and, on every rank,
    on ONE device: the statistics row of every step and every resident agent's record must be identical.  Returns
    (ok, mismatching statistics rows, mismatching agents)."""
    import torch
    import torch.distributed as dist
    from covid19_agentbasedsimulation_b200.abs import Simulation
    from covid19_agentbasedsimulation_b200.slab import SlabSimulation, slab_bounds
    rank = dist.get_rank()
    kw = workload_kwargs(n_total, masked=False)          # dense contacts: migrants, ghosts and infections all occur
    side = kw["length"]
    cols = make_population(n_total, 0, side, side, seed=77, id0=0, total_wealth_per_agent=1e4 / 20)
    one = Simulation(seed=3, device=local_rank, triggers_simulation=lockdown_triggers(), **kw)
    one.set_population_arrays(cols["x"], cols["y"], cols["status"], cols["age"], cols["stratum"], cols["wealth"])
    rows_one = one.run(steps)
    full = one._download()
    slabs = SlabSimulation(seed=3, distributed=True, slab_devices=[local_rank], transport=transport,
                           triggers_simulation=lockdown_triggers(), **kw)
    lo, hi = slab_bounds(side, world)[rank]
    mine = (cols["x"] >= max(lo, -2 ** 31)) & (cols["x"] < hi)
    part = {k: np.ascontiguousarray(v[mine]) for k, v in cols.items()}
    slabs.set_population_arrays(part["x"], part["y"], part["status"], part["age"], part["stratum"], part["wealth"],
                                ids=part["id"], population_size=n_total)
    rows_slab = slabs.run(steps)
    res = slabs._handle.download(slabs._handle.capacity)
    bad_rows = int(np.sum(np.any(rows_one != rows_slab, axis=1)))
    ids = res["id"].astype(np.int64)
    bad_agents = 0
    for k in ("x", "y", "status", "severity", "infected_time"):
        bad_agents += int(np.sum(full[k][ids] != res[k]))
    bad_agents += int(np.sum(full["wealth"][ids].view(np.int64) != res["wealth"].view(np.int64)))
    t = torch.tensor([bad_rows, bad_agents, len(ids)], device="cuda", dtype=torch.int64)
    dist.all_reduce(t)
    slabs.close()
    one._handle.close()
    resident_total = int(t[2].item())
    ok = int(t[0].item()) == 0 and int(t[1].item()) == 0 and resident_total == n_total
    return ok, int(t[0].item()), int(t[1].item()), resident_total


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from covid19_agentbasedsimulation_b200 import build
    build.build_library()
    from covid19_agentbasedsimulation_b200.abs import Simulation
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

    sampler = ClockSampler(local_rank)
    sampler.start()                      # before anything is timed: the first samples are already there at the barrier

    invariance = None
    if world > 1:
        ok, bad_rows, bad_agents, resident = slab_invariance(local_rank, world, args.transport)
        invariance = {"slab_invariance": ok, "mismatching_statistics_rows": bad_rows, "mismatching_agents": bad_agents,
                      "agents": resident, "invariance_steps": 12,
                      "what": "3e5 agents, unmasked (r = 1, rate .9) with the conditional-lockdown triggers, stepped on "
                              "{} slabs (one per rank, transport {}) and on one device by every rank: every "
                              "statistics row and every agent record (position, status, severity, infected_time, "
                              "wealth bits) compared".format(world, args.transport)}

    def build_case(n):
        """Simulation (N=1) or one slab of a SlabSimulation (N>1) with n agents on this rank, uploaded from pinned
        host memory.  Returns (sim, handle, upload, step, kw, side, pinned columns, h2d bytes)."""
        kw = workload_kwargs(n * world)
        side = kw["length"]
        common = dict(seed=0, device=local_rank, triggers_simulation=lockdown_triggers(), max_cells=args.max_cells,
                      history_capacity=max(4096, args.steps + args.warmup + 64), **kw)
        if world == 1:
            sim = Simulation(**common)
            x_lo, x_hi = 0, side
        else:
            sim = SlabSimulation(distributed=True, slab_devices=[local_rank], transport=args.transport, **common)
            lo, hi = slab_bounds(side, world)[rank]
            x_lo, x_hi = max(lo, 0), min(hi - 1, side)
        cols = make_population(n, x_lo, x_hi, side, seed=rank, id0=rank * n, total_wealth_per_agent=1e4 / 20)
        if world == 1:
            del cols["id"]       # ids are the list positions 0..n-1 (abs.py:15): nothing to send
        pinned = {k: torch.from_numpy(v).pin_memory() for k, v in cols.items()}
        ptrs = {k: t.data_ptr() for k, t in pinned.items()}
        h2d = sum(t.numel() * t.element_size() for t in pinned.values())
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
        return sim, h, upload, step, kw, side, cols, pinned, h2d

    def timed(h, fn):
        if world == 1 or args.transport == "p2p":   # the library's own stream
            return timed_on(h, fn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    n = args.agents                      # per GPU (weak scaling: the domain grows with the number of GPUs)
    sim, h, upload, step, kw, side, cols, pinned, h2d_bytes = build_case(n)
    upload()
    h.sync()

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps
    step(args.warmup)
    h.sync()
    launches0 = h.info()["kernel_launches"]
    barrier()
    s0 = sampler.mark()
    ms = timed(h, lambda: step(args.steps))
    s1 = sampler.mark()
    barrier()
    launches = h.info()["kernel_launches"] - launches0
    ms = max_over_ranks(ms)
    value = world * n * args.steps / (ms * 1e-3)
    stats_last = h.stats()[0]

    # ---- per-kernel device times (events around every launch), same steady state
    prof_steps = min(args.steps, 50)
    h.profile_enable(True)
    step(prof_steps)
    prof = h.profile_read()
    h.profile_enable(False)
    info = h.info()
    cells_per_agent = info["cells_x"] * info["cells_y"] / float(max(1, info["n"]))
    kernels = {}
    for name, (cnt, tot) in prof.items():
        avg = tot / cnt
        entry = {"launches_per_step": cnt / prof_steps, "avg_ms": avg}
        if name in KERNEL_BYTES:
            entry["alg_bytes_per_agent"] = KERNEL_BYTES[name](cells_per_agent)
            entry["achieved_gbs"] = entry["alg_bytes_per_agent"] * n / (avg * 1e-3) / 1e9
        kernels[name] = entry
    step_ms_prof = sum(tot for _, tot in prof.values()) / prof_steps
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
    if os.path.exists(traffic_file):     # DRAM bytes per launch of an ncu --set full capture: quoted only for the very
        with open(traffic_file) as f:    # kernel sources it was captured on, at this population size
            t = json.load(f)
        if t.get("csrc_digest") == csrc_digest() and t.get("agents") == n and top in t.get("dram_bytes_per_launch", {}):
            roofline["traffic"] = t["dram_bytes_per_launch"][top]
            roofline["traffic_source"] = t.get("source")

    # ---- end to end through the C ABI with HOST buffers: every step uploads the population from pinned
    #      host memory (set_population, abs.py:65-69), executes, and reads the statistics back (abs.py:291-314)
    e2e_steps = max(3, min(args.steps, args.e2e_steps))

    def e2e_once():
        upload()
        step(1)
        h.stats()

    e2e_once()
    barrier()
    e2e_ms = max_over_ranks(timed(h, lambda: [e2e_once() for _ in range(e2e_steps)]))
    barrier()
    e2e_value = world * n * e2e_steps / (e2e_ms * 1e-3)
    clocks = sampler.stop(s0, s1)

    # ---- the literal metric of SURVEY.md 8d through the PUBLIC classes on a resident population:
    #      for _ in range(K): sim.execute(); sim.get_statistics('all')
    e2e_python = None
    if world == 1:
        upload()
        sim.population_size = n
        sim.execute()
        sim.get_statistics('all')
        k_py = min(args.steps, 100)
        t0 = time.perf_counter()
        for _ in range(k_py):
            sim.execute()
            sim.get_statistics(kind='all')
        wall = time.perf_counter() - t0
        e2e_python = {"loop": "sim.execute(); sim.get_statistics('all') through covid19_agentbasedsimulation_b200.abs."
                              "Simulation, population resident, one host round trip (statistics) per iteration",
                      "steps": k_py, "ms_per_step": 1e3 * wall / k_py, "value": n * k_py / wall,
                      "unit": "agent-updates/s", "clock": "host wall clock around the loop",
                      "fraction_of_value": (n * k_py / wall) / value}

    # ---- the same call chain with the inputs ALREADY on the device (cabs_bind_device: torch tensors' data_ptr()):
    #      what a pipeline that produces populations on the GPU pays -- no PCIe copy, one pack pass + the first sort
    e2e_device = None
    if world == 1:
        dev_cols = {k: t.cuda(local_rank) for k, t in pinned.items()}
        dev_ptrs = {k: t.data_ptr() for k, t in dev_cols.items()}
        torch.cuda.synchronize()

        def dev_once():
            h.bind_device(n, dev_ptrs)
            step(1)
            h.stats()

        dev_once()
        dms = timed(h, lambda: [dev_once() for _ in range(e2e_steps)])
        e2e_device = {"what": "cabs_bind_device (device-resident columns, no host copy) + cabs_step(1) + cabs_stats, every step",
                      "steps": e2e_steps, "ms_per_step": dms / e2e_steps, "value": n * e2e_steps / (dms * 1e-3),
                      "unit": "agent-updates/s", "h2d_bytes_per_step": 0}
        del dev_cols

    other = None
    if not args.no_other_configs:
        other = {}
        try:
            if world == 1 and rank == 0:
                # the same population without masks (SURVEY.md 8d config 3 asks for both), and under the sequential
                # schedule (the reference's own contact order, exact), masked and unmasked
                other["config3_unmasked_synchronous"] = variant_line(n, cols, False, "synchronous", 45, local_rank)
                other["config3_masked_sequential"] = variant_line(n, cols, True, "sequential", 50, local_rank)
                other["config3_unmasked_sequential"] = variant_line(n, cols, False, "sequential", 45, local_rank)
                sys.path.insert(0, os.path.join(ROOT, "scripts"))
                import graph_bench
                other["config1_simulation_defaults"] = graph_bench.config1()
                other["config2_graph_1k_lockdown"] = graph_bench.config2()
                other["config5_ensemble_1gpu_sample"] = graph_bench.ensemble(args.ensemble_persons, 4, 48)
            if world > 1:
                # BASELINE.json configs[3] literally: 10^8 agents over the GPUs of the box (strong scaling in N)
                if args.transport == "p2p":
                    sim.close()
                del sim, h, pinned
                torch.cuda.empty_cache()
                n4 = args.config4_agents // world
                sim4, h4, upload4, step4, kw4, side4, cols4, pinned4, _ = build_case(n4)
                upload4()
                h4.sync()
                step4(args.warmup)
                h4.sync()
                barrier()
                k4 = min(args.steps, 100)
                ms4 = max_over_ranks(timed(h4, lambda: step4(k4)))
                barrier()
                v4 = world * n4 * k4 / (ms4 * 1e-3)
                other["config4_literal"] = {
                    "config": "Simulation {} agents slab-partitioned over {} GPUs ({} per GPU), {} steps, transport {}"
                              .format(n4 * world, world, n4, k4, args.transport),
                    "scaling": "strong", "ms_per_step": ms4 / k4, "agent_updates_per_s": v4,
                    "whole_step_frac": v4 / world * B_ALG / 1e9 / peak,
                    "final_infected": float(h4.stats()[0][1])}
                sim, h = sim4, h4
        except Exception as ex:  # the headline line must still be printed
            other["error"] = repr(ex)

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            base_value, base_dt = cpu_baseline(args.cpu_sample, 5, kw)
            cpu = {"value": base_value, "unit": "agent-updates/s", "cores": 1, "kind": "port",
                   "sample": "{} agents x 5 steps of oracle/sync_oracle.py (numpy, 1 thread) at the same density, "
                             "{:.1f} s".format(args.cpu_sample, base_dt),
                   "reference_path": reference_path("default")}
        line = {
            "metric": "agent-updates/sec", "value": value, "unit": "agent-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32/u8/f64", "data": "synthetic",
            "config": {"workload": ("config3: Simulation 1e7 agents, conditional lockdown + masks, synthetic uniform "
                                    "placement (BASELINE.json configs[2])" if world == 1 else
                                    "config4-style: one Simulation of {} agents slab-partitioned over {} GPUs, migrant + "
                                    "ghost exchange and statistics rows over NVLink every step, conditional lockdown "
                                    "+ masks (BASELINE.json configs[3] at {} agents/GPU; the literal 1e8-agent run is "
                                    "under other_configs.config4_literal)".format(n * world, world, n)),
                       "agents_per_gpu": n, "agents": n * world, "length": side, "height": side,
                       "contagion_distance": 0.05, "contagion_rate": 0.1, "schedule": "synchronous",
                       "cache": "working set {:.0f} MB per GPU >> 126 MB L2 (inputs larger than L2)"
                       .format((56.0 * n + 8 * info["cells_x"] * info["cells_y"]) / 1e6),
                       "cell_edge": info["cell_edge"], "cells": info["cells_x"] * info["cells_y"],
                       "multi_gpu": ("x-slabs, transport=" + args.transport) if world > 1 else "single device"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "agent-updates/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 12 * 8 + 7 * 8 + 4, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps},
            "e2e_python": e2e_python,
            "e2e_device_inputs": e2e_device,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "final_stats": {"Susceptible": float(stats_last[0]), "Infected": float(stats_last[1])},
            "other_configs": other,
        }
        if invariance is not None:
            line.update(invariance)
        print(json.dumps(line))
    if world > 1:
        if args.transport == "p2p":
            sim.close()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="simulation", choices=["simulation", "ensemble"])
    ap.add_argument("--agents", type=int, default=10_000_000, help="agents per GPU")
    ap.add_argument("--config4-agents", type=int, default=100_000_000, help="total agents of other_configs.config4_literal")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    ap.add_argument("--ref-sample", type=int, default=500_000)
    ap.add_argument("--ref-max-steps", type=int, default=25, help="steps per core of the reference arm's bounded sample")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "collective"],
                    help="slab exchange at N > 1: NVLink peer stores + flags, or torch.distributed collectives")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the side measurements")
    ap.add_argument("--max-cells", type=int, default=0, help="cap on cell-list cells (0 = library default)")
    ap.add_argument("--ensemble-persons", type=int, default=100_000)
    ap.add_argument("--ensemble-seeds", type=int, default=0,
                    help="seeds per scenario of --workload ensemble (0: 128 per GPU = 896 simulations per GPU, which at "
                         "--gpus 8 is BASELINE.json configs[4] literally: 7 x 1 024 x 1e5)")
    ap.add_argument("--ensemble-iterations", type=int, default=48)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "ensemble":
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import ensemble_bench
        ensemble_bench.run(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
