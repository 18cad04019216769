"""ctypes binding of libcovidabs.so (include/covidabs.h).

The CUDA library is the product path; there is deliberately no CPU fallback here: if the shared
library is missing or no GPU is visible, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcovidabs.so")

ABI_VERSION = 1
NUM_STATS = 12
MAX_TRIGGERS = 8
SYNCHRONOUS, SEQUENTIAL = 0, 1

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_DOMAIN, ERR_NOGPU = 0, -1, -2, -3, -4, -5

# every symbol include/covidabs.h declares (checked by tests/test_capi_symbols.py)
EXPORTED = ("cabs_create", "cabs_destroy", "cabs_last_error", "cabs_set_params", "cabs_get_params", "cabs_set_triggers",
            "cabs_upload", "cabs_bind_device", "cabs_snapshot", "cabs_upload_f64", "cabs_download_positions_f64", "cabs_set_rules", "cabs_set_step", "cabs_download",
            "cabs_step", "cabs_sync", "cabs_stats", "cabs_stats_ex", "cabs_stats_history",
            "cabs_contact_pairs", "cabs_cross_contact", "cabs_get_info", "cabs_event_record", "cabs_event_elapsed_ms",
            "cabs_profile_enable", "cabs_profile_read", "cabs_selftest_normals", "cabs_set_slab", "cabs_slab_phase", "cabs_set_stream",
            "cabs_slab_block_bytes", "cabs_set_slab_direct", "cabs_slab_step", "cabs_slab_run", "cabs_device_alloc", "cabs_device_free",
            "cabs_ipc_export", "cabs_ipc_open", "cabs_ipc_close",
            "cabs_graph_world_sizes", "cabs_graph_build_world", "cabs_graph_create", "cabs_graph_destroy", "cabs_graph_last_error", "cabs_graph_upload", "cabs_graph_run",
            "cabs_graph_sync", "cabs_graph_stats", "cabs_graph_stats_all", "cabs_graph_aggregate", "cabs_graph_moments", "cabs_graph_download",
            "cabs_graph_last_run_ms")

GRAPH_NUM_STATS, GRAPH_MAX_BUSINESS = 14, 16
GRAPH_MODE_NONE, GRAPH_MODE_LOCKDOWN, GRAPH_MODE_CONDITIONAL_LOCKDOWN, GRAPH_MODE_VERTICAL_ISOLATION = 0, 1, 2, 3

SLAB_ROW, MIGRANT_BYTES, GHOST_BYTES, EXCHANGE_HEADER_BYTES = 16, 32, 16, 32
INT32_MIN, INT32_MAX = -2 ** 31, 2 ** 31 - 1


class NativeError(RuntimeError):
    def __init__(self, code, message):
        super(NativeError, self).__init__("libcovidabs error {}: {}".format(code, message))
        self.code = code


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("capacity", C.c_uint64),
                ("max_cells", C.c_uint64), ("seed", C.c_uint64), ("history_capacity", C.c_uint32),
                ("flags", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("length", C.c_double), ("height", C.c_double), ("contagion_distance", C.c_double),
                ("contagion_rate", C.c_double), ("critical_limit", C.c_double), ("amplitudes", C.c_double * 4),
                ("minimum_income", C.c_double), ("minimum_expense", C.c_double), ("basic_income", C.c_double * 5),
                ("age_hospitalization_probs", C.c_double * 9), ("age_severe_probs", C.c_double * 9),
                ("age_death_probs", C.c_double * 9), ("population_size", C.c_uint64),
                ("recovering_time", C.c_int32), ("wealth_scale_bits", C.c_int32), ("schedule", C.c_int32),
                ("reserved", C.c_int32)]


class Trigger(C.Structure):
    _fields_ = [("metric", C.c_int32), ("op", C.c_int32), ("threshold", C.c_double), ("attribute", C.c_int32),
                ("reserved", C.c_int32), ("value", C.c_double * 4)]


FLAG_F64_POSITIONS = 1
RULE_MOVE_STAY, RULE_MOVE_POINT, RULE_ATTRIBUTE = 1, 2, 3
RATTR = {"status": 0, "infected_status": 1, "infected_time": 2, "age": 3, "social_stratum": 4, "wealth": 5}
RULE_CONDITION_WORDS, MAX_RULES = 250, 8


class Rule(C.Structure):
    _fields_ = [("kind", C.c_int32), ("attribute", C.c_int32), ("target_x", C.c_double), ("target_y", C.c_double),
                ("sigma", C.c_double), ("p", C.c_double), ("q", C.c_double),
                ("condition", C.c_uint32 * RULE_CONDITION_WORDS), ("reserved", C.c_uint32 * 2)]


class SoaHost(C.Structure):
    _fields_ = [("n", C.c_uint64), ("x", C.c_void_p), ("y", C.c_void_p), ("status", C.c_void_p),
                ("severity", C.c_void_p), ("infected_time", C.c_void_p), ("age", C.c_void_p),
                ("stratum", C.c_void_p), ("wealth", C.c_void_p), ("id", C.c_void_p)]


class Info(C.Structure):
    _fields_ = [("n", C.c_uint64), ("step", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("grid_origin_x", C.c_int32), ("grid_origin_y", C.c_int32), ("cell_edge", C.c_int32),
                ("cells_x", C.c_int32), ("cells_y", C.c_int32), ("device_error", C.c_uint32)]


class Slab(C.Structure):
    _fields_ = [("rank", C.c_int32), ("n_ranks", C.c_int32), ("x_lo", C.c_int32), ("x_hi", C.c_int32),
                ("migrant_capacity", C.c_uint32), ("ghost_capacity", C.c_uint32),
                ("migrant_send", C.c_void_p * 2), ("migrant_recv", C.c_void_p * 2),
                ("ghost_send", C.c_void_p * 2), ("ghost_recv", C.c_void_p * 2), ("stats_rows", C.c_void_p)]


MAX_RANKS = 16


class SlabDirect(C.Structure):
    _fields_ = [("rank", C.c_int32), ("n_ranks", C.c_int32), ("x_lo", C.c_int32), ("x_hi", C.c_int32),
                ("migrant_capacity", C.c_uint32), ("ghost_capacity", C.c_uint32), ("blocks", C.c_void_p * MAX_RANKS)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_uint64), ("total_ms", C.c_double)]


class GraphConfig(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("n_sims", C.c_uint32), ("max_persons", C.c_uint32),
                ("max_houses", C.c_uint32), ("max_business", C.c_uint32), ("history_capacity", C.c_uint32),
                ("cell_capacity", C.c_uint32)]


class GraphParams(C.Structure):
    _fields_ = [("amplitudes", C.c_double * 4), ("basic_income", C.c_double * 5), ("minimum_income", C.c_double),
                ("minimum_expense", C.c_double), ("total_wealth", C.c_double), ("critical_limit", C.c_double),
                ("contagion_distance", C.c_double), ("contagion_rate", C.c_double), ("business_distance", C.c_double),
                ("age_hospitalization_probs", C.c_double * 9), ("age_severe_probs", C.c_double * 9),
                ("age_death_probs", C.c_double * 9), ("population_size", C.c_uint64), ("seed", C.c_uint64),
                ("incubation_time", C.c_int32), ("contagion_time", C.c_int32), ("recovering_time", C.c_int32),
                ("mode", C.c_int32), ("sleep", C.c_int32), ("money_bits", C.c_int32)]


# (field, numpy dtype) of the array columns of cabs_graph_world, in declaration order
GRAPH_PERSON_COLS = (("px", np.int32), ("py", np.int32), ("status", np.uint8), ("severity", np.uint8),
                     ("infected_time", np.uint8), ("age", np.uint8), ("stratum", np.uint8), ("active", np.uint8),
                     ("isolated", np.uint8), ("house", np.int32), ("employer", np.int32), ("hire_seq", np.int32),
                     ("pwealth", np.float64))
GRAPH_HOUSE_COLS = (("hx", np.int32), ("hy", np.int32), ("hstratum", np.int32), ("hsize", np.int32),
                    ("hwealth", np.float64), ("hfixed", np.float64), ("hincomes", np.float64), ("hexpenses", np.float64))
GRAPH_BUSINESS_COLS = (("bx", np.int32), ("by", np.int32), ("bstratum", np.int32), ("bstocks", np.int32),
                       ("bsales", np.int32), ("bnum_employees", np.int32), ("bprice", np.float64),
                       ("bwealth", np.float64), ("bfixed", np.float64), ("bincomes", np.float64))
GRAPH_SCALARS = (("gov_x", C.c_int32), ("gov_y", C.c_int32), ("health_x", C.c_int32), ("health_y", C.c_int32),
                 ("gov_wealth", C.c_double), ("gov_price", C.c_double), ("gov_fixed", C.c_double),
                 ("health_wealth", C.c_double), ("health_fixed", C.c_double), ("health_expenses", C.c_double),
                 ("iteration", C.c_int32), ("reserved2", C.c_int32))


class GraphWorld(C.Structure):
    _fields_ = ([("n_persons", C.c_int32), ("n_houses", C.c_int32), ("n_business", C.c_int32), ("reserved", C.c_int32)] +
                [(k, C.c_void_p) for k, _ in GRAPH_PERSON_COLS] + [(k, C.c_void_p) for k, _ in GRAPH_HOUSE_COLS] +
                [(k, C.c_void_p) for k, _ in GRAPH_BUSINESS_COLS] + list(GRAPH_SCALARS))


class GraphWorldSpec(C.Structure):
    _fields_ = [("population_size", C.c_int32), ("length", C.c_int32), ("height", C.c_int32),
                ("total_business", C.c_int32), ("homemates_avg", C.c_int32), ("homemates_std", C.c_int32),
                ("initial_infected_perc", C.c_double), ("initial_immune_perc", C.c_double),
                ("homeless_rate", C.c_double), ("unemployment_rate", C.c_double), ("total_wealth", C.c_double),
                ("public_gdp_share", C.c_double), ("business_gdp_share", C.c_double), ("minimum_income", C.c_double),
                ("minimum_expense", C.c_double), ("basic_income", C.c_double * 5), ("lorenz_curve", C.c_double * 5),
                ("isolation_rate", C.c_double)]


_lib = None


def load():
    """Load libcovidabs.so (built by `__graft_entry__.build()` / `python -m covid19_agentbasedsimulation_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("{} is missing: build it with `python -m covid19_agentbasedsimulation_b200.build` "
                          "(there is no CPU fallback for the agent step)".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    if os.environ.get("CABS_AB_OLD_LIB"):   # A/B timing of a library built from an older tree (scripts/ab_variants.sh)
        class _Stub(object):
            argtypes = restype = None
        for name in EXPORTED:
            if not hasattr(lib, name):
                setattr(lib, name, _Stub())
    lib.cabs_create.argtypes = [C.POINTER(Config), C.POINTER(P)]
    lib.cabs_destroy.argtypes = [P]
    lib.cabs_destroy.restype = None
    lib.cabs_last_error.argtypes = [P]
    lib.cabs_last_error.restype = C.c_char_p
    lib.cabs_set_params.argtypes = [P, C.POINTER(Params)]
    lib.cabs_get_params.argtypes = [P, C.POINTER(Params)]
    lib.cabs_set_triggers.argtypes = [P, C.POINTER(Trigger), C.c_int]
    lib.cabs_upload.argtypes = [P, C.POINTER(SoaHost)]
    lib.cabs_download.argtypes = [P, C.POINTER(SoaHost)]
    lib.cabs_bind_device.argtypes = [P, C.POINTER(SoaHost)]
    lib.cabs_snapshot.argtypes = [P, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                  C.POINTER(C.c_uint64)]
    lib.cabs_set_step.argtypes = [P, C.c_uint64]
    lib.cabs_upload_f64.argtypes = [P, C.POINTER(SoaHost), C.c_void_p, C.c_void_p]
    lib.cabs_download_positions_f64.argtypes = [P, C.c_void_p, C.c_void_p, C.c_uint64]
    lib.cabs_set_rules.argtypes = [P, C.POINTER(Rule), C.c_int, C.c_int]
    lib.cabs_step.argtypes = [P, C.c_int]
    lib.cabs_sync.argtypes = [P]
    lib.cabs_stats.argtypes = [P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.cabs_stats_ex.argtypes = [P, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(Params), C.POINTER(C.c_uint64)]
    lib.cabs_stats_history.argtypes = [P, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]
    lib.cabs_contact_pairs.argtypes = [P, C.POINTER(C.c_int64), C.c_size_t, C.POINTER(C.c_size_t)]
    lib.cabs_get_info.argtypes = [P, C.POINTER(Info)]
    lib.cabs_event_record.argtypes = [P, C.c_int]
    lib.cabs_event_elapsed_ms.argtypes = [P, C.c_int, C.c_int, C.POINTER(C.c_double)]
    lib.cabs_profile_enable.argtypes = [P, C.c_int]
    lib.cabs_profile_read.argtypes = [P, C.POINTER(KernelTime), C.c_int, C.POINTER(C.c_int)]
    lib.cabs_selftest_normals.argtypes = [C.c_int, C.c_uint, C.POINTER(C.c_double)]
    lib.cabs_set_slab.argtypes = [P, C.POINTER(Slab)]
    lib.cabs_slab_phase.argtypes = [P, C.c_int, C.c_int]
    lib.cabs_set_stream.argtypes = [P, C.c_void_p]
    lib.cabs_slab_block_bytes.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    lib.cabs_slab_block_bytes.restype = C.c_size_t
    lib.cabs_set_slab_direct.argtypes = [P, C.POINTER(SlabDirect)]
    lib.cabs_slab_step.argtypes = [P, C.c_int]
    lib.cabs_slab_run.argtypes = [P, C.c_int]
    lib.cabs_device_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(P)]
    lib.cabs_device_free.argtypes = [C.c_int, P]
    lib.cabs_ipc_export.argtypes = [C.c_int, P, C.c_char * 64]
    lib.cabs_ipc_open.argtypes = [C.c_int, C.c_char * 64, C.POINTER(P)]
    lib.cabs_ipc_close.argtypes = [C.c_int, P]
    lib.cabs_graph_world_sizes.argtypes = [C.POINTER(GraphWorldSpec), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int32)]
    lib.cabs_graph_build_world.argtypes = [C.POINTER(GraphWorldSpec), C.c_uint64, C.POINTER(GraphWorld)]
    lib.cabs_graph_create.argtypes = [C.POINTER(GraphConfig), C.POINTER(P)]
    lib.cabs_cross_contact.argtypes = [P, P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_uint64, C.c_uint32]
    lib.cabs_graph_destroy.argtypes = [P]
    lib.cabs_graph_destroy.restype = None
    lib.cabs_graph_last_error.argtypes = [P]
    lib.cabs_graph_last_error.restype = C.c_char_p
    lib.cabs_graph_upload.argtypes = [P, C.c_uint32, C.POINTER(GraphWorld), C.POINTER(GraphParams)]
    lib.cabs_graph_run.argtypes = [P, C.c_int]
    lib.cabs_graph_sync.argtypes = [P]
    lib.cabs_graph_stats.argtypes = [P, C.c_uint32, C.c_int64, C.c_uint32, C.POINTER(C.c_double)]
    lib.cabs_graph_aggregate.argtypes = [P, C.c_int64, C.c_uint32, C.POINTER(C.c_double)]
    lib.cabs_graph_stats_all.argtypes = [P, C.c_int64, C.c_uint32, C.POINTER(C.c_double)]
    lib.cabs_graph_moments.argtypes = [P, C.c_uint32, C.c_uint32, C.c_int64, C.c_uint32, C.POINTER(C.c_double)]
    lib.cabs_graph_download.argtypes = [P, C.c_uint32, C.POINTER(GraphWorld)]
    lib.cabs_graph_last_run_ms.argtypes = [P, C.POINTER(C.c_double)]
    for name in EXPORTED:
        fn = getattr(lib, name)
        if name not in ("cabs_destroy", "cabs_last_error", "cabs_graph_destroy", "cabs_graph_last_error",
                        "cabs_slab_block_bytes"):
            fn.restype = C.c_int
    _lib = lib
    return lib


SOA_DTYPES = (("x", np.int32), ("y", np.int32), ("status", np.uint8), ("severity", np.uint8),
              ("infected_time", np.uint8), ("age", np.uint8), ("stratum", np.uint8), ("wealth", np.float64),
              ("id", np.uint32))


class Handle(object):
    """Owns one `cabs_sim*`."""

    def __init__(self, capacity, seed=0, device=0, max_cells=0, history_capacity=4096, f64=False):
        self._lib = load()
        self._h = C.c_void_p()
        self.f64 = bool(f64)
        cfg = Config(ABI_VERSION, int(device), int(capacity), int(max_cells), int(seed) & 0xFFFFFFFFFFFFFFFF,
                     int(history_capacity), FLAG_F64_POSITIONS if f64 else 0)
        rc = self._lib.cabs_create(C.byref(cfg), C.byref(self._h))
        if rc != OK:
            msg = self._lib.cabs_last_error(None)
            raise NativeError(rc, msg.decode() if msg else "cabs_create failed")
        self.capacity = int(capacity)
        self.steps = 0      # mirror of the handle's step counter (draw clock / index of the statistics ring)

    def _check(self, rc):
        if rc != OK:
            msg = self._lib.cabs_last_error(self._h)
            raise NativeError(rc, msg.decode() if msg else "?")

    def set_step(self, step):
        self._check(self._lib.cabs_set_step(self._h, int(step)))
        self.steps = int(step)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.cabs_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params):
        self._check(self._lib.cabs_set_params(self._h, C.byref(params)))

    def get_params(self):
        p = Params()
        self._check(self._lib.cabs_get_params(self._h, C.byref(p)))
        return p

    def set_triggers(self, triggers):
        arr = (Trigger * max(1, len(triggers)))(*triggers)
        self._check(self._lib.cabs_set_triggers(self._h, arr, len(triggers)))

    @staticmethod
    def _soa(cols, n):
        s = SoaHost()
        s.n = n
        for name, dt in SOA_DTYPES:
            a = cols.get(name)
            if a is None:
                setattr(s, name, None)
            else:
                assert a.dtype == dt and a.flags["C_CONTIGUOUS"] and a.shape == (n,), name
                setattr(s, name, a.ctypes.data)
        return s

    def upload(self, cols):
        """cols: dict of contiguous numpy arrays with the dtypes of SOA_DTYPES ('id' optional)."""
        n = len(cols["x"])
        s = self._soa(cols, n)
        self._check(self._lib.cabs_upload(self._h, C.byref(s)))

    def upload_f64(self, cols, xf, yf):
        """Like upload(), with binary64 positions (off-lattice handles); cols need no 'x' / 'y'."""
        n = len(cols["status"])
        xf = np.ascontiguousarray(xf, dtype=np.float64)
        yf = np.ascontiguousarray(yf, dtype=np.float64)
        assert xf.shape == (n,) and yf.shape == (n,)
        s = self._soa({k: v for k, v in cols.items() if k not in ("x", "y")}, n)
        self._check(self._lib.cabs_upload_f64(self._h, C.byref(s), xf.ctypes.data, yf.ctypes.data))

    def download_positions(self, n):
        xf, yf = np.empty(n, dtype=np.float64), np.empty(n, dtype=np.float64)
        self._check(self._lib.cabs_download_positions_f64(self._h, xf.ctypes.data, yf.ctypes.data, n))
        return xf, yf

    def set_rules(self, move_rules, attr_rules):
        rules = list(move_rules) + list(attr_rules)
        arr = (Rule * max(1, len(rules)))(*rules)
        self._check(self._lib.cabs_set_rules(self._h, arr, len(move_rules), len(attr_rules)))

    def upload_pointers(self, n, ptrs):
        """Upload from raw host addresses (e.g. pinned torch tensors): ptrs maps column name -> int address."""
        s = SoaHost()
        s.n = n
        for name, _ in SOA_DTYPES:
            setattr(s, name, ptrs.get(name))
        self._check(self._lib.cabs_upload(self._h, C.byref(s)))

    def bind_device(self, n, ptrs):
        """Ingest columns that already live in device memory (ptrs: column name -> device address, e.g. torch
        tensor.data_ptr()); no host copy is involved."""
        s = SoaHost()
        s.n = n
        for name, _ in SOA_DTYPES:
            setattr(s, name, ptrs.get(name))
        self._check(self._lib.cabs_bind_device(self._h, C.byref(s)))

    def snapshot(self, n, stride=1):
        """Position and status of every stride-th agent of the population list (the live-view feed)."""
        m = (int(n) + int(stride) - 1) // int(stride)
        out = {"x": np.empty(m, dtype=np.int32), "y": np.empty(m, dtype=np.int32),
               "status": np.empty(m, dtype=np.uint8), "severity": np.empty(m, dtype=np.uint8)}
        count = C.c_uint64(0)
        self._check(self._lib.cabs_snapshot(self._h, int(stride), out["x"].ctypes.data, out["y"].ctypes.data,
                                            out["status"].ctypes.data, out["severity"].ctypes.data, m, C.byref(count)))
        return {k: v[:count.value] for k, v in out.items()}

    def download(self, n):
        cols = {name: np.empty(n, dtype=dt) for name, dt in SOA_DTYPES}
        s = self._soa(cols, n)
        self._check(self._lib.cabs_download(self._h, C.byref(s)))
        if s.n != n:
            cols = {k: v[:s.n] for k, v in cols.items()}
        return cols

    def set_slab(self, slab):
        self._check(self._lib.cabs_set_slab(self._h, C.byref(slab)))

    def slab_phase(self, phase, advance=True):
        self._check(self._lib.cabs_slab_phase(self._h, int(phase), 1 if advance else 0))

    def set_slab_direct(self, slab):
        self._check(self._lib.cabs_set_slab_direct(self._h, C.byref(slab)))

    def slab_step(self, advance=True):
        self._check(self._lib.cabs_slab_step(self._h, 1 if advance else 0))

    def slab_run(self, nsteps):
        self._check(self._lib.cabs_slab_run(self._h, int(nsteps)))

    def set_stream(self, cuda_stream):
        self._check(self._lib.cabs_set_stream(self._h, C.c_void_p(int(cuda_stream) or None)))

    def step(self, nsteps=1):
        self._check(self._lib.cabs_step(self._h, int(nsteps)))
        self.steps += int(nsteps)

    def sync(self):
        self._check(self._lib.cabs_sync(self._h))

    def stats(self):
        v = (C.c_double * NUM_STATS)()
        c = (C.c_int64 * 7)()
        self._check(self._lib.cabs_stats(self._h, v, c))
        return np.array(v[:], dtype=np.float64), np.array(c[:], dtype=np.int64)

    def stats_ex(self):
        """(values[12], counts[7], Params as on the device now, trigger epoch) from one copy and one synchronisation."""
        v = (C.c_double * NUM_STATS)()
        c = (C.c_int64 * 7)()
        p = Params()
        e = C.c_uint64(0)
        self._check(self._lib.cabs_stats_ex(self._h, v, c, C.byref(p), C.byref(e)))
        return np.array(v[:], dtype=np.float64), np.array(c[:], dtype=np.int64), p, int(e.value)

    def stats_history(self, first, count):
        out = np.empty((int(count), NUM_STATS), dtype=np.float64)
        self._check(self._lib.cabs_stats_history(self._h, int(first), int(count),
                                                 out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def cross_contact(self, other, position, other_position, contagion_distance, seed, pair_code):
        """Contacts between this population placed at `position` and `other` at `other_position` (abs.py:352-366)."""
        self._check(self._lib.cabs_cross_contact(self._h, other._h, int(position[0]), int(position[1]),
                                                 int(other_position[0]), int(other_position[1]),
                                                 float(contagion_distance), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                 int(pair_code)))

    def contact_pairs(self, capacity=None):
        cap = int(capacity) if capacity else 1 << 16
        while True:
            buf = np.empty((cap, 2), dtype=np.int64)
            n = C.c_size_t(0)
            self._check(self._lib.cabs_contact_pairs(self._h, buf.ctypes.data_as(C.POINTER(C.c_int64)), cap,
                                                     C.byref(n)))
            if n.value <= cap:
                return buf[:n.value]
            cap = int(n.value)

    def info(self):
        i = Info()
        self._check(self._lib.cabs_get_info(self._h, C.byref(i)))
        return {k: getattr(i, k) for k, _ in Info._fields_}

    def event_record(self, slot):
        self._check(self._lib.cabs_event_record(self._h, int(slot)))

    def event_elapsed_ms(self, a, b):
        ms = C.c_double(0)
        self._check(self._lib.cabs_event_elapsed_ms(self._h, int(a), int(b), C.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        self._check(self._lib.cabs_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """{kernel name: (launches, total_ms)} since profiling was enabled / last read."""
        arr = (KernelTime * 32)()
        n = C.c_int(0)
        self._check(self._lib.cabs_profile_read(self._h, arr, 32, C.byref(n)))
        return {arr[i].name.decode(): (int(arr[i].launches), float(arr[i].total_ms)) for i in range(n.value)}


class GraphHandle(object):
    """Owns one `cabs_graph*`: `n_sims` GraphSimulation worlds of the same capacity on one device."""

    def __init__(self, n_sims, max_persons, max_houses, max_business, device=0, history_capacity=2048,
                 cell_capacity=0):
        self._lib = load()
        self._h = C.c_void_p()
        cfg = GraphConfig(ABI_VERSION, int(device), int(n_sims), int(max_persons), int(max_houses), int(max_business),
                          int(history_capacity), int(cell_capacity))
        rc = self._lib.cabs_graph_create(C.byref(cfg), C.byref(self._h))
        if rc != OK:
            msg = self._lib.cabs_graph_last_error(None)
            raise NativeError(rc, msg.decode() if msg else "cabs_graph_create failed")
        self.n_sims, self.max_persons, self.max_houses, self.max_business = (int(n_sims), int(max_persons),
                                                                             int(max_houses), int(max_business))
        self._sizes = {}

    def _check(self, rc):
        if rc != OK:
            msg = self._lib.cabs_graph_last_error(self._h)
            raise NativeError(rc, msg.decode() if msg else "?")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.cabs_graph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _world_struct(world, keep):
        """dict of arrays / scalars (keys of cabs_graph_world) -> GraphWorld; `keep` holds the converted arrays."""
        w = GraphWorld()
        w.n_persons, w.n_houses, w.n_business = len(world["px"]), len(world["hx"]), len(world["bx"])
        for cols, n in ((GRAPH_PERSON_COLS, w.n_persons), (GRAPH_HOUSE_COLS, w.n_houses),
                        (GRAPH_BUSINESS_COLS, w.n_business)):
            for name, dt in cols:
                a = world.get(name)
                a = np.zeros(n, dtype=dt) if a is None else np.ascontiguousarray(a, dtype=dt)
                assert a.shape == (n,), name
                keep.append(a)
                setattr(w, name, a.ctypes.data)
        for name, _ in GRAPH_SCALARS:
            if name != "reserved2":
                setattr(w, name, world.get(name, -1 if name == "iteration" else 0))
        return w

    def upload(self, sim, world, params):
        keep = []
        w = self._world_struct(world, keep)
        self._check(self._lib.cabs_graph_upload(self._h, int(sim), C.byref(w), C.byref(params)))
        self._sizes[int(sim)] = (w.n_persons, w.n_houses, w.n_business)

    def run(self, iterations):
        self._check(self._lib.cabs_graph_run(self._h, int(iterations)))

    def sync(self):
        self._check(self._lib.cabs_graph_sync(self._h))

    def last_run_ms(self):
        ms = C.c_double(0)
        self._check(self._lib.cabs_graph_last_run_ms(self._h, C.byref(ms)))
        return ms.value

    def stats(self, sim, first, count):
        out = np.empty((int(count), GRAPH_NUM_STATS), dtype=np.float64)
        self._check(self._lib.cabs_graph_stats(self._h, int(sim), int(first), int(count),
                                               out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def stats_all(self, first, count):
        """Rows of every simulation of the handle -> float64 [n_sims, count, 14] (one strided copy)."""
        out = np.empty((self.n_sims, int(count), GRAPH_NUM_STATS), dtype=np.float64)
        self._check(self._lib.cabs_graph_stats_all(self._h, int(first), int(count),
                                                   out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def aggregate(self, first, count):
        """Min, Avg, Std (ddof=0), Max over the simulations of the handle -> float64 [count, 14, 4] (device reduction)."""
        out = np.empty((int(count), GRAPH_NUM_STATS, 4), dtype=np.float64)
        self._check(self._lib.cabs_graph_aggregate(self._h, int(first), int(count),
                                                   out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def moments(self, sim_first, sim_count, first, count):
        """min, sum, sum of squares, max over simulations [sim_first, sim_first + sim_count) -> float64 [count, 14, 4]."""
        out = np.empty((int(count), GRAPH_NUM_STATS, 4), dtype=np.float64)
        self._check(self._lib.cabs_graph_moments(self._h, int(sim_first), int(sim_count), int(first), int(count),
                                                 out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def download(self, sim):
        n, nh, nb = self._sizes[int(sim)]
        world = {}
        for cols, m in ((GRAPH_PERSON_COLS, n), (GRAPH_HOUSE_COLS, nh), (GRAPH_BUSINESS_COLS, nb)):
            for name, dt in cols:
                world[name] = np.zeros(m, dtype=dt)
        keep = []
        w = self._world_struct(world, keep)
        for (name, _), a in zip(GRAPH_PERSON_COLS + GRAPH_HOUSE_COLS + GRAPH_BUSINESS_COLS, keep):
            world[name] = a
        self._check(self._lib.cabs_graph_download(self._h, int(sim), C.byref(w)))
        for name, _ in GRAPH_SCALARS:
            if name != "reserved2":
                world[name] = getattr(w, name)
        return world


def build_world_native(spec, seed):
    """cabs_graph_build_world -> dict of columns (keys of cabs_graph_world).  Host only; releases the GIL."""
    lib = load()
    n, nh, nb = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    _util_check(lib.cabs_graph_world_sizes(C.byref(spec), C.byref(n), C.byref(nh), C.byref(nb)), "cabs_graph_world_sizes")
    w = GraphWorld()
    world = {}
    for cols, m in ((GRAPH_PERSON_COLS, n.value), (GRAPH_HOUSE_COLS, nh.value), (GRAPH_BUSINESS_COLS, nb.value)):
        for name, dt in cols:
            world[name] = np.zeros(m, dtype=dt)
            setattr(w, name, world[name].ctypes.data)
    _util_check(lib.cabs_graph_build_world(C.byref(spec), int(seed) & 0xFFFFFFFFFFFFFFFF, C.byref(w)),
                "cabs_graph_build_world")
    for name, _ in GRAPH_HOUSE_COLS:
        world[name] = world[name][:w.n_houses]
    for name, _ in GRAPH_SCALARS:
        if name != "reserved2":
            world[name] = getattr(w, name)
    return world


# ---- device memory shared between the processes of one box (exchange blocks of the direct slab transport)
def _util_check(rc, what):
    if rc != OK:
        msg = load().cabs_last_error(None)
        raise NativeError(rc, "{}: {}".format(what, msg.decode() if msg else "?"))


def selftest_normals(device=0, sweep_log2=30):
    """Measured errors of the special-function move draw against its defining arithmetic (see covidabs.h)."""
    out = (C.c_double * 6)()
    _util_check(load().cabs_selftest_normals(int(device), int(sweep_log2), out), "cabs_selftest_normals")
    keys = ("radius_error_over_bound", "angle_error_over_bound", "radius_inputs_rejected", "accepted", "wrong", "draws")
    return dict(zip(keys, out[:]))


def slab_block_bytes(n_ranks, migrant_capacity, ghost_capacity):
    return int(load().cabs_slab_block_bytes(int(n_ranks), int(migrant_capacity), int(ghost_capacity)))


def device_alloc(device, nbytes):
    p = C.c_void_p()
    _util_check(load().cabs_device_alloc(int(device), int(nbytes), C.byref(p)), "cabs_device_alloc")
    return p.value


def device_free(device, ptr):
    _util_check(load().cabs_device_free(int(device), C.c_void_p(ptr)), "cabs_device_free")


def ipc_export(device, ptr):
    buf = (C.c_char * 64)()
    _util_check(load().cabs_ipc_export(int(device), C.c_void_p(ptr), buf), "cabs_ipc_export")
    return bytes(buf.raw)


def ipc_open(device, handle):
    buf = (C.c_char * 64).from_buffer_copy(handle)
    p = C.c_void_p()
    _util_check(load().cabs_ipc_open(int(device), buf, C.byref(p)), "cabs_ipc_open")
    return p.value


def ipc_close(device, ptr):
    _util_check(load().cabs_ipc_close(int(device), C.c_void_p(ptr)), "cabs_ipc_close")
