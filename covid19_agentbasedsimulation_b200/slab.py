"""Slab decomposition of a `Simulation` over the GPUs of one box.

The reference is a single process (covid_abs/abs.py); this is the part of the B200 build that lets one population
span several devices.  The spatial domain is cut into x-slabs; each slab is one `cabs_sim` handle that owns the
lattice columns [x_lo, x_hi).  Per step (include/covidabs.h, cabs_slab_phase):

    phase 0  pass A; agents whose move leaves the slab finish their step and are packed      -> exchange migrants
    phase 1  immigrants counted, scan, pass B, immigrants scattered; Infected agents within
             reach of an edge are packed as ghosts                                            -> exchange ghosts
    phase 2  contacts pushed from the local Infected list and from the ghosts; partial
             statistics exported                                                              -> all-gather rows
    phase 3  global statistics, triggers, critical-limit flag, next grid

Draws are keyed on the global agent id, so the result does not depend on the number of slabs
(tests/test_gpu_slabs.py checks bit equality with the one-slab run).

Two transports move the buffers: `LocalComm` (several slabs in one process: device-to-device copies; used by the
partition-invariance tests on one GPU and by single-process multi-GPU runs) and `DistComm` (one slab per process,
`torch.distributed` send/recv + all_gather over NCCL/NVLink; gloo in the CPU tests).
"""
import numpy as np

from covid19_agentbasedsimulation_b200 import _native
from covid19_agentbasedsimulation_b200.abs import Simulation, STAT_KEYS

INT32_MIN, INT32_MAX = _native.INT32_MIN, _native.INT32_MAX


def slab_width(length, n_ranks):
    """Width of the interior slabs: ceil((length + 1) / n_ranks) lattice columns (positions 0..length)."""
    return max(1, -(-(int(length) + 1) // int(n_ranks)))


def slab_bounds(length, n_ranks):
    """[(x_lo, x_hi)] per rank; the outer sides are open because the reference does not confine agents (A.7-9)."""
    w = slab_width(length, n_ranks)
    edges = [k * w for k in range(1, n_ranks)]
    return list(zip([INT32_MIN] + edges, edges + [INT32_MAX]))


def owner_of(x, length, n_ranks):
    """Rank that owns lattice column(s) x."""
    w = slab_width(length, n_ranks)
    return np.clip(np.floor_divide(np.asarray(x, dtype=np.int64), w), 0, n_ranks - 1).astype(np.int64)


def exchange_bytes(capacity, record_bytes):
    return _native.EXCHANGE_HEADER_BYTES + int(capacity) * int(record_bytes)


class SlabBuffers(object):
    """Exchange buffers of one slab (torch tensors: device memory for the library, CPU memory in the gloo tests)."""

    def __init__(self, n_ranks, migrant_capacity, ghost_capacity, device):
        import torch
        self.migrant_capacity, self.ghost_capacity = int(migrant_capacity), int(ghost_capacity)
        mb = exchange_bytes(migrant_capacity, _native.MIGRANT_BYTES)
        gb = exchange_bytes(ghost_capacity, _native.GHOST_BYTES)
        z = lambda nbytes: torch.zeros(nbytes, dtype=torch.uint8, device=device)  # noqa: E731
        self.migrant_send = [z(mb), z(mb)]
        self.migrant_recv = [z(mb), z(mb)]
        self.ghost_send = [z(gb), z(gb)]
        self.ghost_recv = [z(gb), z(gb)]
        self.rows = torch.zeros((n_ranks, _native.SLAB_ROW), dtype=torch.int64, device=device)

    def native(self, rank, n_ranks, x_lo, x_hi):
        s = _native.Slab()
        s.rank, s.n_ranks, s.x_lo, s.x_hi = rank, n_ranks, x_lo, x_hi
        s.migrant_capacity, s.ghost_capacity = self.migrant_capacity, self.ghost_capacity
        for d in range(2):
            s.migrant_send[d] = self.migrant_send[d].data_ptr()
            s.migrant_recv[d] = self.migrant_recv[d].data_ptr()
            s.ghost_send[d] = self.ghost_send[d].data_ptr()
            s.ghost_recv[d] = self.ghost_recv[d].data_ptr()
        s.stats_rows = self.rows.data_ptr()
        return s


class LocalComm(object):
    """All slabs live in this process: neighbours exchange by tensor copies (peer copies if on different GPUs)."""

    def __init__(self, buffers):
        self.buffers = buffers

    def exchange(self, kind):
        b = self.buffers
        for r in range(len(b)):
            send, _ = (b[r].migrant_send, None) if kind == "migrants" else (b[r].ghost_send, None)
            if r > 0:
                dst = b[r - 1].migrant_recv if kind == "migrants" else b[r - 1].ghost_recv
                dst[1].copy_(send[0], non_blocking=True)
            if r + 1 < len(b):
                dst = b[r + 1].migrant_recv if kind == "migrants" else b[r + 1].ghost_recv
                dst[0].copy_(send[1], non_blocking=True)

    def gather_rows(self):
        b = self.buffers
        for r in range(len(b)):
            for g in range(len(b)):
                if g != r:
                    b[r].rows[g].copy_(b[g].rows[g], non_blocking=True)


class DistComm(object):
    """One slab per process: `torch.distributed` point-to-point with the two x-neighbours + an all-gather of the rows."""

    def __init__(self, buffers, rank, world):
        self.b, self.rank, self.world = buffers, int(rank), int(world)

    def exchange(self, kind):
        import torch.distributed as dist
        send = self.b.migrant_send if kind == "migrants" else self.b.ghost_send
        recv = self.b.migrant_recv if kind == "migrants" else self.b.ghost_recv
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, send[0], self.rank - 1))
            ops.append(dist.P2POp(dist.irecv, recv[0], self.rank - 1))
        if self.rank + 1 < self.world:
            ops.append(dist.P2POp(dist.isend, send[1], self.rank + 1))
            ops.append(dist.P2POp(dist.irecv, recv[1], self.rank + 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def gather_rows(self):
        import torch.distributed as dist
        mine = self.b.rows[self.rank].clone()
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(self.b.rows.view(-1), mine)
        else:   # gloo (CPU tests)
            dist.all_gather([self.b.rows[g] for g in range(self.world)], mine)


class SlabSimulation(Simulation):
    """`Simulation` whose population spans `slabs` x-slabs.

    slabs=G, distributed=False: G handles in this process (devices `slab_devices`, default all on `device`).
    distributed=True: this process owns slab `RANK` of `WORLD_SIZE` (torch.distributed must be initialised).
    Everything else is the `Simulation` API; statistics are global (identical on every rank).
    """

    def __init__(self, **kwargs):
        super(SlabSimulation, self).__init__(**kwargs)
        self.distributed = bool(kwargs.get("distributed", False))
        if self.distributed:
            import torch.distributed as dist
            self.n_slabs, self.my_ranks = dist.get_world_size(), [dist.get_rank()]
        else:
            self.n_slabs = int(kwargs.get("slabs", 1))
            self.my_ranks = list(range(self.n_slabs))
        self.slab_devices = kwargs.get("slab_devices", None) or [self.device] * len(self.my_ranks)
        self.migrant_capacity = kwargs.get("migrant_capacity", 0)
        self.ghost_capacity = kwargs.get("ghost_capacity", 0)
        self.capacity_factor = float(kwargs.get("capacity_factor", 1.25))
        self.transport = kwargs.get("transport", "collective")
        '''"collective": buffers move between the phases by torch.distributed send/recv + all_gather (or copies when
        the slabs share a process); "p2p" (distributed only): the kernels write into the neighbours' memory over
        NVLink peer mappings (CUDA IPC) and synchronise with flags -- no collective, one enqueue per step'''
        if self.schedule != "synchronous" and (self.n_slabs > 1 or self.distributed):
            raise ValueError("schedule='sequential' is defined on one device only: the in-step infection chain of the "
                             "reference is not local to a slab (SURVEY.md 8e)")
        if self.transport not in ("collective", "p2p"):
            raise ValueError("transport must be 'collective' or 'p2p'")
        if self.transport == "p2p" and not self.distributed:
            raise ValueError("transport='p2p' needs one slab per process (distributed=True)")
        self._handles, self._buffers, self._comm = [], [], None
        self._block, self._peer_blocks = None, []

    # ------------------------------------------------------------------ device side
    def _create(self, counts):
        import torch
        for h in self._handles:
            h.close()
        if self.seed is None:
            self.seed = int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))
        bounds = slab_bounds(self.length, self.n_slabs)
        biggest = max(1, int(max(counts)))
        mig = int(self.migrant_capacity) or max(4096, min(biggest // 16, 1 << 17))
        gho = int(self.ghost_capacity) or mig
        self._handles, self._buffers = [], []
        if self.distributed:
            # the exchange buffers have the same size on every rank (fixed-size messages): take the largest request
            import torch.distributed as dist
            sizes = [None] * self.n_slabs
            dist.all_gather_object(sizes, (int(mig), int(gho)))
            mig, gho = max(s[0] for s in sizes), max(s[1] for s in sizes)
        if self.transport == "p2p":
            self._create_p2p(counts[0], mig, gho, bounds)
            return
        for k, r in enumerate(self.my_ranks):
            dev = self.slab_devices[k]
            cap = int(max(1024, counts[k] * self.capacity_factor + 2 * mig))
            h = _native.Handle(capacity=cap, seed=self.seed, device=dev, max_cells=self.max_cells,
                               history_capacity=self.history_capacity)
            with torch.cuda.device(dev):
                h.set_stream(torch.cuda.current_stream().cuda_stream)
                b = SlabBuffers(self.n_slabs, mig, gho, torch.device("cuda", dev))
            self._handles.append(h)
            self._buffers.append(b)
        self._handle = self._handles[0]
        self._pushed = None
        self._pushed_triggers = None
        self._push_params()
        for k, r in enumerate(self.my_ranks):
            self._handles[k].set_slab(self._buffers[k].native(r, self.n_slabs, bounds[r][0], bounds[r][1]))
        if self.distributed:
            self._comm = DistComm(self._buffers[0], self.my_ranks[0], self.n_slabs)
        else:
            self._comm = LocalComm(self._buffers)

    def _create_p2p(self, count, mig, gho, bounds):
        """One slab per process; exchange blocks mapped into the neighbours through CUDA IPC."""
        import torch
        import torch.distributed as dist
        rank, dev = self.my_ranks[0], self.slab_devices[0]
        cap = int(max(1024, count * self.capacity_factor + 2 * mig))
        h = _native.Handle(capacity=cap, seed=self.seed, device=dev, max_cells=self.max_cells,
                           history_capacity=self.history_capacity)
        # the handle keeps its own stream: nothing of torch's is ordered against the step in this transport, and
        # pairs of steps are replayed as a captured CUDA graph (the legacy default stream cannot be captured)
        self._release_p2p()
        self._block = _native.device_alloc(dev, _native.slab_block_bytes(self.n_slabs, mig, gho))
        handles = [None] * self.n_slabs
        dist.all_gather_object(handles, _native.ipc_export(dev, self._block))
        d = _native.SlabDirect()
        d.rank, d.n_ranks, d.x_lo, d.x_hi = rank, self.n_slabs, bounds[rank][0], bounds[rank][1]
        d.migrant_capacity, d.ghost_capacity = mig, gho
        self._peer_blocks = []
        for r in range(self.n_slabs):
            if r == rank:
                d.blocks[r] = self._block
            else:
                p = _native.ipc_open(dev, handles[r])
                self._peer_blocks.append(p)
                d.blocks[r] = p
        self._handles, self._buffers = [h], []
        self._handle = h
        self._pushed = None
        self._pushed_triggers = None
        self._push_params()
        h.set_slab_direct(d)
        self._comm = None
        dist.barrier()

    def _release_p2p(self):
        dev = self.slab_devices[0]
        for p in self._peer_blocks:
            _native.ipc_close(dev, p)
        self._peer_blocks = []
        if self._block is not None:
            _native.device_free(dev, self._block)
            self._block = None

    def close(self):
        for h in self._handles:
            h.close()
        self._handles = []
        if self.transport == "p2p":
            import torch.distributed as dist
            dist.barrier()          # nobody unmaps a block a neighbour may still be writing to
            self._release_p2p()

    def _targets(self):
        return self._handles

    def _build(self, advance):
        if self.transport == "p2p":
            self._handles[0].slab_step(advance)
            return
        for phase in range(4):
            for h in self._handles:
                h.slab_phase(phase, advance)
            if phase == 0:
                self._comm.exchange("migrants")
            elif phase == 1:
                self._comm.exchange("ghosts")
            elif phase == 2:
                self._comm.gather_rows()
        if self.distributed:
            # collectives of consecutive builds are not queued behind each other: the conservative transport waits
            # for the build (the p2p transport is the asynchronous one)
            for h in self._handles:
                h.sync()

    def set_population_arrays(self, x, y, status, age, stratum, wealth, severity=None, infected_time=None, ids=None,
                              population_size=None):
        """Upload a population.  In-process: the global columns, split here by owner.  distributed=True: this rank's
        agents only, with their global `ids`, and the global `population_size`."""
        n = len(x)
        x = np.ascontiguousarray(x, dtype=np.int32)
        cols = {
            "x": x, "y": np.ascontiguousarray(y, dtype=np.int32),
            "status": np.ascontiguousarray(status, dtype=np.uint8),
            "severity": (np.full(n, 1, dtype=np.uint8) if severity is None
                         else np.ascontiguousarray(severity, dtype=np.uint8)),
            "infected_time": (np.zeros(n, dtype=np.uint8) if infected_time is None
                              else np.ascontiguousarray(infected_time, dtype=np.uint8)),
            "age": np.ascontiguousarray(np.minimum(age, 255), dtype=np.uint8),
            "stratum": np.ascontiguousarray(stratum, dtype=np.uint8),
            "wealth": np.ascontiguousarray(wealth, dtype=np.float64),
            "id": (np.arange(n, dtype=np.uint32) if ids is None else np.ascontiguousarray(ids, dtype=np.uint32)),
        }
        if self.distributed:
            if population_size is None:
                raise ValueError("distributed upload needs the global population_size")
            self.population_size = int(population_size)
            parts = [cols]
        else:
            self.population_size = n if population_size is None else int(population_size)
            owner = owner_of(x, self.length, self.n_slabs)
            parts = [{k: np.ascontiguousarray(v[owner == r]) for k, v in cols.items()} for r in self.my_ranks]
        counts = [len(p["x"]) for p in parts]
        if not self._handles or any(h.capacity < c for h, c in zip(self._handles, counts)):
            self._create(counts)
        else:
            self._push_params()
        for h, p in zip(self._handles, parts):
            h.upload(p)
        self._build(advance=False)
        self._invalidate_host()
        self.statistics = None

    # ------------------------------------------------------------------ step
    def execute(self):
        if not self._handles:
            raise RuntimeError("initialize() or set_population() first")
        if len(self.triggers_population) > 0:
            raise NotImplementedError("per-agent Python triggers cannot run on the device")
        self._push_params()
        self._build(advance=True)
        self._iteration += 1
        self._invalidate_host()
        for trigger in self._host_triggers():  # abs.py:267-271
            if trigger['condition'](self):
                attr = trigger['attribute']
                self.__dict__[attr] = trigger['action'](self.__dict__[attr])
        self.statistics = None

    def run(self, iterations):
        iterations = int(iterations)
        if self._host_triggers() or iterations > self.history_capacity:
            return super(SlabSimulation, self).run(iterations)
        self._push_params()
        first = self._handle.info()["step"]     # the handles' clock is not reset by a re-upload
        if self.transport == "p2p":
            self._handles[0].slab_run(iterations)     # pairs of steps replay a captured CUDA graph
        else:
            for _ in range(iterations):
                self._build(advance=True)
        self._iteration += iterations
        self._invalidate_host()
        self.statistics = None
        return self._handle.stats_history(first, iterations)

    def _download(self):
        """Resident agents of the local slabs, ordered by id (all agents when every slab is local)."""
        if self._host is None:
            parts = [h.download(h.capacity) for h in self._handles]
            cols = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
            order = np.argsort(cols["id"], kind="stable")
            self._host = {k: v[order] for k, v in cols.items()}
        return self._host

    def resident_counts(self):
        return [h.info()["n"] for h in self._handles]

    def get_contact_pairs(self):
        raise NotImplementedError("the contact-pair hook is per device; use a one-slab Simulation")

    def sync(self):
        for h in self._handles:
            h.sync()
