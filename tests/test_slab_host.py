"""Host side of the slab decomposition (no GPU): slab geometry, ownership, and the two-process exchange pattern
of `DistComm` over gloo with CPU tensors standing in for the device buffers."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_cover_the_line_once():
    from covid19_agentbasedsimulation_b200.slab import slab_bounds, owner_of, INT32_MIN, INT32_MAX
    for length, g in [(10, 1), (10, 2), (7071, 8), (22360, 8), (5, 8), (100, 3)]:
        b = slab_bounds(length, g)
        assert len(b) == g and b[0][0] == INT32_MIN and b[-1][1] == INT32_MAX
        for k in range(g - 1):
            assert b[k][1] == b[k + 1][0] and b[k][0] < b[k][1]
        xs = np.arange(-50, length + 50)
        own = owner_of(xs, length, g)
        for x, r in zip(xs, own):
            assert b[r][0] <= x < b[r][1]


def test_exchange_buffer_sizes_match_header():
    import re
    from covid19_agentbasedsimulation_b200 import _native
    from covid19_agentbasedsimulation_b200.slab import exchange_bytes
    hdr = open(os.path.join(ROOT, "include", "covidabs.h")).read()
    get = lambda name: int(re.search(r"#define %s (\d+)" % name, hdr).group(1))  # noqa: E731
    assert get("CABS_SLAB_ROW") == _native.SLAB_ROW
    assert get("CABS_MIGRANT_BYTES") == _native.MIGRANT_BYTES
    assert get("CABS_GHOST_BYTES") == _native.GHOST_BYTES
    assert get("CABS_EXCHANGE_HEADER_BYTES") == _native.EXCHANGE_HEADER_BYTES
    assert exchange_bytes(10, 32) == 32 + 320


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from covid19_agentbasedsimulation_b200.slab import SlabBuffers, DistComm
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        b = SlabBuffers(world, migrant_capacity=4, ghost_capacity=3, device="cpu")
        for d in range(2):
            b.migrant_send[d].fill_(10 * rank + d + 1)
            b.ghost_send[d].fill_(100 + 10 * rank + d + 1)
        b.rows[rank] = torch.arange(16, dtype=torch.int64) + 1000 * rank
        comm = DistComm(b, rank, world)
        comm.exchange("migrants")
        comm.exchange("ghosts")
        comm.gather_rows()
        ok = True
        if rank > 0:      # from the left neighbour's "towards higher x" buffer
            ok &= bool((b.migrant_recv[0] == 10 * (rank - 1) + 2).all()) and bool((b.ghost_recv[0] == 100 + 10 * (rank - 1) + 2).all())
        else:
            ok &= bool((b.migrant_recv[0] == 0).all())
        if rank + 1 < world:  # from the right neighbour's "towards lower x" buffer
            ok &= bool((b.migrant_recv[1] == 10 * (rank + 1) + 1).all()) and bool((b.ghost_recv[1] == 100 + 10 * (rank + 1) + 1).all())
        else:
            ok &= bool((b.migrant_recv[1] == 0).all())
        for g in range(world):
            ok &= bool((b.rows[g] == torch.arange(16, dtype=torch.int64) + 1000 * g).all())
        out.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distcomm_neighbour_exchange_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
