"""Slab decomposition on one GPU: G logical slabs in one process, exchanges by device copies (LocalComm).

The result must not depend on the number of slabs: draws are keyed on the global agent id, migrants carry
their complete record, ghosts push the cross-boundary contacts, statistics are summed over the slabs.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pop(n, side, seed, infected=0.02, immune=0.01):
    from oracle.sync_oracle import synthetic_population
    return synthetic_population(n, side, side, seed=seed, infected=infected, immune=immune)


def _run(cls, pop, tw, side, steps, seed, **kw):
    sim = cls(seed=seed, population_size=len(pop["x"]), length=side, height=side, total_wealth=tw, **kw)
    sim.set_population_arrays(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    rows = []
    for _ in range(steps):
        sim.execute()
        s = sim.get_statistics('all')
        rows.append([float(s[k]) for k in s])
    return sim, np.array(rows)


@pytest.mark.parametrize("slabs", [1, 2, 3, 5])
@pytest.mark.parametrize("dist,rate", [(1.0, 0.9), (2.5, 0.5), (0.05, 1.0)])
def test_partition_invariance(native_lib, slabs, dist, rate):
    from covid19_agentbasedsimulation_b200.abs import Simulation
    from covid19_agentbasedsimulation_b200.slab import SlabSimulation
    n, side = 30000, 300
    pop, tw = _pop(n, side, seed=3, infected=0.05)
    one, rows1 = _run(Simulation, pop, tw, side, 12, seed=11, contagion_distance=dist, contagion_rate=rate)
    many, rowsg = _run(SlabSimulation, pop, tw, side, 12, seed=11, contagion_distance=dist, contagion_rate=rate,
                       slabs=slabs)
    assert np.array_equal(rows1, rowsg)
    a, b = one._download(), many._download()
    for k in ("id", "x", "y", "status", "severity", "infected_time"):
        assert np.array_equal(a[k], b[k]), k
    assert a["wealth"].tobytes() == b["wealth"].tobytes()
    assert sum(many.resident_counts()) == n
    if slabs > 1:
        assert min(many.resident_counts()) > 0


def test_slabs_match_oracle_with_triggers(native_lib):
    from covid19_agentbasedsimulation_b200.abs import Trigger
    from covid19_agentbasedsimulation_b200.agents import Status
    from covid19_agentbasedsimulation_b200.slab import SlabSimulation
    from oracle.sync_oracle import SyncSimulation, Trigger as OTrigger
    from oracle.seq_oracle import S, I, R
    n, side = 20000, 200
    pop, tw = _pop(n, side, seed=5, infected=0.1)
    low = {Status.Susceptible: 1.5, Status.Recovered_Immune: 1.5, Status.Infected: 1.5}
    high = {Status.Susceptible: 5, Status.Recovered_Immune: 5, Status.Infected: 5}
    trig = [Trigger('Infected', '>=', .2, 'amplitudes', low), Trigger('Infected', '<=', .05, 'amplitudes', high)]
    gpu = SlabSimulation(seed=9, population_size=n, length=side, height=side, total_wealth=tw, slabs=4,
                         triggers_simulation=trig, critical_limit=0.02)
    gpu.set_population_arrays(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    orc = SyncSimulation(seed=9, population_size=n, length=side, height=side, total_wealth=tw, critical_limit=0.02,
                         triggers=[OTrigger('Infected', '>=', .2, 'amplitudes', {S: 1.5, R: 1.5, I: 1.5}),
                                   OTrigger('Infected', '<=', .05, 'amplitudes', {S: 5, R: 5, I: 5})])
    orc.set_population(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    table = gpu.run(25)
    for it in range(25):
        ostats = orc.execute()
        assert [float(v) for v in ostats.values()] == table[it].tolist(), it
    h = gpu._download()
    assert np.array_equal(h["x"], orc.x) and np.array_equal(h["status"], orc.status)
    assert h["wealth"].tobytes() == orc.wealth.tobytes()


def test_migrant_overflow_is_reported(native_lib):
    from covid19_agentbasedsimulation_b200 import _native
    from covid19_agentbasedsimulation_b200.slab import SlabSimulation
    n, side = 20000, 100
    pop, tw = _pop(n, side, seed=1)
    sim = SlabSimulation(seed=2, population_size=n, length=side, height=side, total_wealth=tw, slabs=2,
                         migrant_capacity=8)
    sim.set_population_arrays(pop["x"], pop["y"], pop["status"], pop["age"], pop["stratum"], pop["wealth"])
    sim.execute()
    with pytest.raises(_native.NativeError):
        sim.get_statistics('all')


@pytest.mark.parametrize("transport", ["collective", "p2p", "p2p-critical", "p2p-fused", "p2p-fused-critical"])
def test_two_process_nccl_slabs_match_single_device(native_lib, transport):
    """One slab per process (needs >= 2 GPUs; skipped on a one-GPU box): buffers exchanged by NCCL send/recv +
    all_gather ("collective") or written straight into the neighbour's memory over NVLink ("p2p")."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "dist_slab_check.py")]
    env = dict(os.environ)
    if "critical" in transport:         # the critical-limit flag flips mid-run: the emigrants updated in the exchange
        env["SLAB_CRIT"] = "0.005"      # kernels (or inside pass A) must see it exactly as the residents in pass B do
    if "fused" in transport:            # emigrants packed inside pass A, one contact launch (opt-in: measured slower)
        env["CABS_SLAB_FUSED"] = "1"
    transport = transport.split("-")[0]
    env["SLAB_TRANSPORT"] = transport
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=600,
                         env=env)
    assert res.returncode == 0 and "DIST_SLAB_CHECK {} OK".format(transport) in res.stdout, res.stdout[-3000:]
