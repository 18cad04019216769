/*
 * covidabs.h -- C ABI of the B200-native COVID-ABS agent step.
 *
 * The reference (petroniocandido/COVID19_AgentBasedSimulation) is pure Python and has no FFI of
 * its own; the seam it offers is the duck type `batch_experiment` drives:
 *     sim = simulation_type(**kwargs)      covid_abs/experiments.py:185
 *     sim.initialize()                     covid_abs/experiments.py:186  -> abs.py:116
 *     sim.execute()                        covid_abs/experiments.py:193  -> abs.py:232
 *     sim.get_statistics(kind='all')       covid_abs/experiments.py:194  -> abs.py:291
 * Each entry point below cites the reference code it stands in for.  The Python host layer
 * (covid19_agentbasedsimulation_b200/abs.py) binds these with ctypes and keeps the covid_abs API.
 *
 * Conventions
 *  - plain pointers and sizes only; the caller owns every host buffer it passes in;
 *  - every call returns CABS_OK (0) or a negative error code and never throws across the ABI;
 *    `cabs_last_error` returns the message of the last failing call on that handle;
 *  - one handle = one CUDA device + one stream; handles are not thread-safe (the reference is
 *    single-threaded); `cabs_step` only enqueues work, `cabs_stats*`, `cabs_download`,
 *    `cabs_contact_pairs` and `cabs_sync` wait for it.
 */
#ifndef COVIDABS_H
#define COVIDABS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CABS_ABI_VERSION 1

/* error codes */
#define CABS_OK 0
#define CABS_ERR_INVALID (-1)   /* bad argument / handle state                          */
#define CABS_ERR_CUDA (-2)      /* CUDA runtime error (message has the CUDA string)    */
#define CABS_ERR_CAPACITY (-3)  /* population / output buffer larger than the capacity */
#define CABS_ERR_DOMAIN (-4)    /* coordinates outside the supported +-2^28 box        */
#define CABS_ERR_NOGPU (-5)     /* no CUDA device: there is no CPU fallback            */

/* covid_abs/agents.py:13-16 (declaration order = statistics key order) */
enum { CABS_SUSCEPTIBLE = 0, CABS_INFECTED = 1, CABS_RECOVERED_IMMUNE = 2, CABS_DEATH = 3 };
/* covid_abs/agents.py:23-26 */
enum { CABS_EXPOSED = 0, CABS_ASYMPTOMATIC = 1, CABS_HOSPITALIZATION = 2, CABS_SEVERE = 3 };

/* contact schedule */
enum {
  CABS_SYNCHRONOUS = 0, /* one neighbour scan: only agents infected before the contact phase transmit */
  CABS_SEQUENTIAL = 1   /* abs.py:261-265 order: in-step chains along increasing (i,j) rank; exact, by rounds
                           of relaxation inside one cooperative launch (no host round trip); one device only  */
};

/* statistics slots, in the key order of abs.py:300-312 */
enum {
  CABS_STAT_SUSCEPTIBLE = 0, CABS_STAT_INFECTED, CABS_STAT_RECOVERED_IMMUNE, CABS_STAT_DEATH,
  CABS_STAT_ASYMPTOMATIC, CABS_STAT_HOSPITALIZATION, CABS_STAT_SEVERE,
  CABS_STAT_Q1, CABS_STAT_Q2, CABS_STAT_Q3, CABS_STAT_Q4, CABS_STAT_Q5,
  CABS_NUM_STATS
};

typedef struct cabs_sim cabs_sim;

/* cabs_config.flags */
#define CABS_FLAG_F64_POSITIONS 1u /* off-lattice mode: positions are binary64 (see cabs_upload_f64, cabs_set_rules) */

/* creation-time configuration (sizes of the device allocations) */
typedef struct cabs_config {
  uint32_t abi_version;      /* CABS_ABI_VERSION                                                  */
  int32_t device;            /* CUDA device ordinal                                               */
  uint64_t capacity;         /* max agents resident on this device                                */
  uint64_t max_cells;        /* cap on cell-list cells (0 = auto: capacity/8 while the epidemic is sparse, 1.5*capacity when dense; explicit = fixed) */
  uint64_t seed;             /* Philox key; draws are keyed (seed, agent id, step)                */
  uint32_t history_capacity; /* per-step statistics records kept on the device (ring)             */
  uint32_t flags;            /* CABS_FLAG_*                                                       */
} cabs_config;

/* Simulation attributes a step reads: covid_abs/abs.py:17-49 and covid_abs/common.py:13-27.
 * May be changed between steps (that is what abs.py:267-271 `triggers_simulation` do). */
typedef struct cabs_params {
  double length, height;                 /* abs.py:19-22                                     */
  double contagion_distance;             /* abs.py:27  (sqrt(dx^2+dy^2) <= d, abs.py:10,258) */
  double contagion_rate;                 /* abs.py:29                                        */
  double critical_limit;                 /* abs.py:31                                        */
  double amplitudes[4];                  /* abs.py:33-36, by status code; Death ignored      */
  double minimum_income;                 /* abs.py:38 (unused by Simulation, kept for parity) */
  double minimum_expense;                /* abs.py:40                                        */
  double basic_income[5];                /* common.py:27                                     */
  double age_hospitalization_probs[9];   /* common.py:13                                     */
  double age_severe_probs[9];            /* common.py:14                                     */
  double age_death_probs[9];             /* common.py:15                                     */
  uint64_t population_size;              /* abs.py:17: denominator of the status fractions   */
  int32_t recovering_time;               /* 20 in abs.py:225                                 */
  int32_t wealth_scale_bits;             /* Q1..Q5 are summed as round(wealth*2^bits) int64  */
  int32_t schedule;                      /* CABS_SYNCHRONOUS / CABS_SEQUENTIAL               */
  int32_t reserved;
} cabs_params;

/* Declarative form of one abs.py:74-84 simulation trigger, evaluated on the device at the end of
 * every step against the statistics of the previous step (what the reference's memo holds when
 * batch_experiment drives it, abs.py:273,298 + experiments.py:193-194):
 *     if stats[metric] <op> threshold:  attribute = value                                         */
enum { CABS_OP_GE = 0, CABS_OP_LE = 1, CABS_OP_GT = 2, CABS_OP_LT = 3 };
enum { CABS_ATTR_AMPLITUDES = 0, CABS_ATTR_CONTAGION_RATE = 1, CABS_ATTR_CONTAGION_DISTANCE = 2,
       CABS_ATTR_CRITICAL_LIMIT = 3 };
typedef struct cabs_trigger {
  int32_t metric;    /* CABS_STAT_*                                   */
  int32_t op;        /* CABS_OP_*                                     */
  double threshold;
  int32_t attribute; /* CABS_ATTR_*                                   */
  int32_t reserved;
  double value[4];   /* amplitudes[4] by status code, else value[0]   */
} cabs_trigger;
#define CABS_MAX_TRIGGERS 8

/* The population as structure of arrays on the host: the fields of covid_abs/agents.py:44-64 that
 * the step reads.  `id` may be NULL on upload (ids 0..n-1 = list order of abs.py:15). */
typedef struct cabs_soa_host {
  uint64_t n;
  int32_t *x, *y;          /* agents.py:45-47 (integer lattice) */
  uint8_t *status;         /* agents.py:49  CABS_SUSCEPTIBLE..CABS_DEATH       */
  uint8_t *severity;       /* agents.py:52  CABS_EXPOSED..CABS_SEVERE          */
  uint8_t *infected_time;  /* agents.py:54                                     */
  uint8_t *age;            /* agents.py:56                                     */
  uint8_t *stratum;        /* agents.py:58  0..4                               */
  double *wealth;          /* agents.py:60                                     */
  uint32_t *id;            /* position in the reference's population list      */
} cabs_soa_host;

typedef struct cabs_info {
  uint64_t n;               /* agents resident                          */
  uint64_t step;            /* steps executed so far                    */
  uint64_t kernel_launches; /* CUDA kernels launched by this handle     */
  int32_t grid_origin_x, grid_origin_y, cell_edge, cells_x, cells_y;
  uint32_t device_error;    /* sticky device-side error bits            */
} cabs_info;

/* Replaces Simulation.__init__ (abs.py:14-49): allocates the device-side population. */
int cabs_create(const cabs_config *cfg, cabs_sim **out);
/* Releases everything the handle owns. */
void cabs_destroy(cabs_sim *sim);
/* Message of the last error on this handle (or of the last failed cabs_create when sim == NULL). */
const char *cabs_last_error(const cabs_sim *sim);

/* Sets the attributes of abs.py:17-49; callable between steps (abs.py:267-271). */
int cabs_set_params(cabs_sim *sim, const cabs_params *params);
/* The attributes as they are on the device now: what cabs_set_params sent, with the fields device-side triggers have
 * rewritten since (amplitudes, contagion_rate, contagion_distance, critical_limit; abs.py:267-271 makes such a change
 * permanent).  Waits for the enqueued steps. */
int cabs_get_params(cabs_sim *sim, cabs_params *out);
/* Installs up to CABS_MAX_TRIGGERS device-side simulation triggers (abs.py:74-84, 267-271). */
int cabs_set_triggers(cabs_sim *sim, const cabs_trigger *triggers, int count);

/* Replaces set_population (abs.py:65-69) / the list built by initialize (abs.py:116-139).  When `id` is given on a
 * one-device handle it must be a permutation of 0..n-1 (checked on the device; the error surfaces as
 * CABS_ERR_INVALID at the next synchronising call). */
int cabs_upload(cabs_sim *sim, const cabs_soa_host *pop);
/* cabs_upload for columns that ALREADY live in device memory of the handle's device (e.g. the data_ptr() of torch
 * tensors): the records are packed straight from them -- no host staging, no host-to-device copy.  `pop` holds device
 * pointers; the columns must be complete when the call is made (synchronise the stream that produced them, or run the
 * handle on that stream with cabs_set_stream).  The handle keeps no reference to them afterwards. */
int cabs_bind_device(cabs_sim *sim, const cabs_soa_host *pop);
/* The step counter is one word of every draw key (seed, agent id, step).  It starts at 0 on a fresh handle, advances
 * by one per executed step and is NOT reset by cabs_upload (re-uploading an edited population mid-run must not replay
 * the draws of step 0).  cabs_set_step restores the clock of a checkpoint: (cabs_download, cabs_get_info().step) ->
 * (cabs_upload, cabs_set_step) resumes a run bit for bit.  Statistics rows of earlier steps are dropped. */
int cabs_set_step(cabs_sim *sim, uint64_t step);
/* Replaces get_population (abs.py:57-63): writes the population back in id order. */
int cabs_download(cabs_sim *sim, cabs_soa_host *pop);

/* ---- off-lattice mode (handles created with CABS_FLAG_F64_POSITIONS; one device, no slabs) ----------------------
 * Population triggers (abs.py:86-95): an entry with attribute 'move' replaces the random walk of the agents its
 * condition selects by the position its action returns (abs.py:169-172; the shipped idiom is a point plus
 * np.random.normal(0, sigma), run_batch.py:446-450), after which x and y are binary64 and the contact test is the
 * reference's floating-point one (abs.py:9-10,258); any other entry rewrites an attribute after update()
 * (abs.py:244-247).  A Python lambda cannot run on the device, so rules are declarative:
 *   condition   any predicate of (status, severity, age, stratum), tabulated as 8 000 bits, bit index
 *               ((status * 4 + severity) * 100 + min(age, 99)) * 5 + stratum;
 *   move rule   new position = target + N(0, sigma) per axis (the move stream's two normals), target = the agent's
 *               own position (CABS_RULE_MOVE_STAY) or (target_x, target_y); the first matching rule wins and the agent
 *               earns no move income, like the early return of abs.py:172;
 *   attribute   status / severity / infected_time / age / stratum := (int)q, or wealth := p * wealth + q; applied in
 *               order after the step's update. */
enum { CABS_RULE_MOVE_STAY = 1, CABS_RULE_MOVE_POINT = 2, CABS_RULE_ATTRIBUTE = 3 };
enum { CABS_RATTR_STATUS = 0, CABS_RATTR_SEVERITY = 1, CABS_RATTR_INFECTED_TIME = 2, CABS_RATTR_AGE = 3,
       CABS_RATTR_STRATUM = 4, CABS_RATTR_WEALTH = 5 };
#define CABS_RULE_CONDITION_WORDS 250
#define CABS_MAX_RULES 8
typedef struct cabs_rule {
  int32_t kind;                 /* CABS_RULE_*                                   */
  int32_t attribute;            /* CABS_RATTR_* (attribute rules)                */
  double target_x, target_y, sigma;
  double p, q;
  uint32_t condition[CABS_RULE_CONDITION_WORDS];
  uint32_t reserved[2];
} cabs_rule;
/* rules[0 .. n_move) are move rules, rules[n_move .. n_move + n_attr) attribute rules, each group in trigger order. */
int cabs_set_rules(cabs_sim *sim, const cabs_rule *rules, int n_move, int n_attr);
/* cabs_upload with binary64 positions (x, y in the order of pop's columns; pop->x / pop->y may be NULL). */
int cabs_upload_f64(cabs_sim *sim, const cabs_soa_host *pop, const double *x, const double *y);
/* The exact positions in id order (cabs_download returns their floors in x, y). */
int cabs_download_positions_f64(cabs_sim *sim, double *x, double *y, uint64_t capacity);

/* Live-view feed: what the reference's animations read per frame (graphics.py:113-142,241-276: every agent's
 * position and status) for every stride-th agent of the population list, in list order -- count = ceil(n / stride)
 * records of 10 bytes cross the bus instead of the whole population.  Any of x, y, status, severity may be NULL.
 * (Off-lattice handles report the floors of the positions.) */
int cabs_snapshot(cabs_sim *sim, uint32_t stride, int32_t *x, int32_t *y, uint8_t *status, uint8_t *severity,
                  uint64_t capacity, uint64_t *count);

/* Replaces `nsteps` calls of Simulation.execute (abs.py:232-273): move, update, contacts,
 * triggers.  Asynchronous. */
int cabs_step(cabs_sim *sim, int nsteps);
/* Waits for the enqueued steps. */
int cabs_sync(cabs_sim *sim);

/* Replaces get_statistics(kind='all') (abs.py:291-314) for the current state.
 * `values`: CABS_NUM_STATS doubles in abs.py key order; `counts` (optional): the 7 raw counts. */
int cabs_stats(cabs_sim *sim, double *values, int64_t *counts);
/* cabs_stats plus, from the same single copy and synchronisation: the attributes as they are on the device now
 * (`params_now`, see cabs_get_params; may be NULL) and the number of times a device-side trigger has fired on this handle
 * (`trigger_epoch`; may be NULL) -- a host mirror refreshes its attributes only when that number moves. */
int cabs_stats_ex(cabs_sim *sim, double *values, int64_t *counts, cabs_params *params_now, uint64_t *trigger_epoch);
/* Statistics after each of steps [first_step, first_step+count) (the rows batch_experiment appends,
 * experiments.py:193-196); `values` holds count*CABS_NUM_STATS doubles. */
int cabs_stats_history(cabs_sim *sim, uint64_t first_step, uint64_t count, double *values);

/* Test hook for the pair loop of abs.py:253-259: all (i, j), i < j (ids), with
 * sqrt(dx^2+dy^2) <= contagion_distance on the current positions, in unspecified order.
 * Writes min(*n_pairs, capacity_pairs) pairs as ij[2k], ij[2k+1]; *n_pairs = total found. */
int cabs_contact_pairs(cabs_sim *sim, int64_t *ij, size_t capacity_pairs, size_t *n_pairs);

int cabs_get_info(cabs_sim *sim, cabs_info *info);

/* ---- slab decomposition over the GPUs of one box (no reference counterpart: the reference is one
 * process).  A handle owns the lattice columns [x_lo, x_hi); the caller owns the
 * exchange buffers (device memory) and moves them between neighbours (NCCL send/recv, or plain copies
 * when several slabs share a device) between the phases of a build:
 *     phase 0   pass A, emigrants packed            -> exchange migrants  (send[d] -> neighbour's recv[1-d])
 *     phase 1   immigrants counted, scan, pass B    -> exchange ghosts    (same pattern)
 *     phase 2   contacts (local + ghosts), export   -> all-gather stats_rows (row `rank` is this handle's)
 *     phase 3   close: global statistics, triggers, next grid
 * advance = 1: a simulation step (Simulation.execute, abs.py:232-273); 0: build the cell list of the current
 * positions (after cabs_upload).  Buffers: 32-byte header {uint32 count} + capacity records (migrants 32 B,
 * ghosts 16 B).  Index 0 = towards lower x, 1 = towards higher x.  With a slab configured, cabs_download
 * returns the resident agents in storage order with their ids, and cabs_contact_pairs only local pairs. */
typedef struct cabs_slab {
  int32_t rank, n_ranks;
  int32_t x_lo, x_hi;                   /* INT32_MIN / INT32_MAX on an open side */
  uint32_t migrant_capacity, ghost_capacity;
  void *migrant_send[2], *migrant_recv[2];
  void *ghost_send[2], *ghost_recv[2];
  void *stats_rows;                     /* n_ranks x CABS_SLAB_ROW int64 */
} cabs_slab;
#define CABS_SLAB_ROW 16
#define CABS_MIGRANT_BYTES 32
#define CABS_GHOST_BYTES 16
#define CABS_EXCHANGE_HEADER_BYTES 32
int cabs_set_slab(cabs_sim *sim, const cabs_slab *slab);
int cabs_slab_phase(cabs_sim *sim, int phase, int advance);

/* Direct exchange: instead of handing buffers to the caller between the phases, the kernels write migrants,
 * ghosts and statistics rows straight into the neighbours' memory over NVLink (peer mappings, e.g. CUDA IPC
 * between the processes of one box) and synchronise with flags, so a whole build is enqueued by ONE call and no
 * collective is launched.  Every rank owns one exchange block of cabs_slab_block_bytes() bytes of zeroed device
 * memory (cabs_device_alloc) and passes the addresses at which it sees every rank's block. */
#define CABS_MAX_RANKS 16
typedef struct cabs_slab_direct {
  int32_t rank, n_ranks;
  int32_t x_lo, x_hi;
  uint32_t migrant_capacity, ghost_capacity;
  void *blocks[CABS_MAX_RANKS];        /* [r] = rank r's exchange block as mapped in this process */
} cabs_slab_direct;
size_t cabs_slab_block_bytes(uint32_t n_ranks, uint32_t migrant_capacity, uint32_t ghost_capacity);
int cabs_set_slab_direct(cabs_sim *sim, const cabs_slab_direct *slab);
int cabs_slab_step(cabs_sim *sim, int advance);   /* the four phases back to back; asynchronous */
int cabs_slab_run(cabs_sim *sim, int nsteps);     /* nsteps steps; pairs of steps replay a captured CUDA graph */
/* Device memory that can be shared between processes, and its CUDA IPC handle (64 bytes). */
int cabs_device_alloc(int device, size_t bytes, void **ptr);
int cabs_device_free(int device, void *ptr);
int cabs_ipc_export(int device, void *ptr, unsigned char handle[64]);
int cabs_ipc_open(int device, const unsigned char handle[64], void **ptr);
int cabs_ipc_close(int device, void *ptr);
/* Run on the caller's CUDA stream (e.g. the one its collectives are ordered on; NULL = CUDA's default stream)
 * instead of the handle's own. */
int cabs_set_stream(cabs_sim *sim, void *cuda_stream);

/* Measurement support (no reference counterpart: the reference has no timing code).
 * CUDA events on the handle's own stream, so that a caller can bracket any sequence of calls above
 * (uploads, steps, statistics reads) with device-side timestamps. */
#define CABS_NUM_EVENTS 8
int cabs_event_record(cabs_sim *sim, int slot);
int cabs_event_elapsed_ms(cabs_sim *sim, int slot_begin, int slot_end, double *ms); /* waits for slot_end */

/* Per-kernel device times: while enabled, every kernel launch of cabs_step is bracketed by events.
 * cabs_profile_read waits, then returns up to `capacity` entries (kernel name, launches, total ms) and resets. */
typedef struct cabs_kernel_time {
  char name[32];
  uint64_t launches;
  double total_ms;
} cabs_kernel_time;
int cabs_profile_enable(cabs_sim *sim, int on);
int cabs_profile_read(cabs_sim *sim, cabs_kernel_time *out, int capacity, int *count);

/* Self-test of the move draw (no reference counterpart).  The displacement of abs.py:174-175, int(randn * amplitude),
 * is DEFINED by a Box-Muller transform in correctly rounded binary32 (csrc/abs_rng.cuh, reproduced by the CPU oracle);
 * the kernels evaluate it on the special-function unit and fall back to the defining arithmetic whenever the result is
 * within the proven error bound of an integer.  This call measures the actual errors on `device`, exhaustively over the
 * 2^23 radius and 2^25 angle inputs, and compares 2^sweep_log2 sampled draws with the defining path:
 *   out[0] worst radius error / bound, out[1] worst cos-sin error / bound (both must be < 1), out[2] radius inputs the
 *   fast path rejects outright, out[3] sampled draws accepted by the fast path, out[4] accepted draws whose integers
 *   differ from the defining path (must be 0), out[5] sampled draws. */
int cabs_selftest_normals(int device, unsigned sweep_log2, double out[6]);


/* Replaces the loop over one pair of populations in MultiPopulationSimulation.execute (covid_abs/abs.py:352-366):
 * `a` sits at (ax, ay) and `b` at (bx, by) of a common plane (the reference's `positions`); agents of the two within
 * `contagion_distance` meet, a Susceptible one next to an Infected one is infected when the draw keyed on
 * (seed, the two ids, step, pair_code) is <= the contagion_rate of ITS OWN population.  Synchronous rule: only agents
 * Infected before the call transmit.  Both handles on one device, at the same step; afterwards cabs_stats of either
 * handle describes its population including the new infections.  O(n_a * n_b) distance tests, like the reference. */
int cabs_cross_contact(cabs_sim *a, cabs_sim *b, int32_t ax, int32_t ay, int32_t bx, int32_t by,
                       double contagion_distance, uint64_t seed, uint32_t pair_code);

/* ======================================================================================================
 * GraphSimulation (covid_abs/network/graph_abs.py): the hourly routine model with houses, businesses,
 * government and healthcare.  One handle holds `n_sims` independent simulations of the same capacity (a
 * scenario x seed ensemble, or a single one); one thread block per simulation runs a whole batch of
 * iterations.  The world is built on the host (GraphSimulation.initialize, graph_abs.py:89-188) and
 * uploaded as structure of arrays; every array below is indexed like the reference's lists
 * (population, houses, business).
 * ====================================================================================================== */
typedef struct cabs_graph cabs_graph;

enum { CABS_GRAPH_MODE_NONE = 0,                 /* run_batch.py scenario1: routine only                  */
       CABS_GRAPH_MODE_LOCKDOWN = 1,             /* run_batch.py:16-19  on_person_move = lockdown         */
       CABS_GRAPH_MODE_CONDITIONAL_LOCKDOWN = 2, /* run_batch.py:22-26  lockdown while Infected > .05     */
       CABS_GRAPH_MODE_VERTICAL_ISOLATION = 3 }; /* run_batch.py:45-50  lockdown of the Inactive          */
/* partial isolation (run_batch.py:31-42) is the per-person `isolated` column with any mode */

#define CABS_GRAPH_NUM_STATS 14  /* Q1..Q5, Business, Government, S, I, R, D, Asymptomatic, Hospitalization, Severe
                                    = key order of graph_abs.py:305-322 and of the published CSVs */
#define CABS_GRAPH_MAX_BUSINESS 16

typedef struct cabs_graph_config {
  uint32_t abi_version;
  int32_t device;
  uint32_t n_sims;
  uint32_t max_persons, max_houses, max_business; /* per simulation */
  uint32_t history_capacity;                      /* statistics rows kept per simulation (ring)            */
  uint32_t cell_capacity;                         /* slots of the contact table per simulation: raised to
                                                   * the next power of two >= 2 x max_persons (0 = that)   */
} cabs_graph_config;

/* graph_abs.py:13-32 + abs.py:17-49 + common.py:13-27 */
typedef struct cabs_graph_params {
  double amplitudes[4];            /* by status code                                             */
  double basic_income[5];
  double minimum_income, minimum_expense, total_wealth, critical_limit;
  double contagion_distance, contagion_rate, business_distance;
  double age_hospitalization_probs[9], age_severe_probs[9], age_death_probs[9];
  uint64_t population_size;        /* denominator of the status fractions (graph_abs.py:315-322) */
  uint64_t seed;                   /* Philox key                                                  */
  int32_t incubation_time, contagion_time, recovering_time;
  int32_t mode;                    /* CABS_GRAPH_MODE_*                                           */
  int32_t sleep;                   /* run_batch.py:8-13: skip hours 1..7 of every day             */
  int32_t money_bits;              /* fractional bits of the fixed-point accounts                 */
} cabs_graph_params;

typedef struct cabs_graph_world {
  int32_t n_persons, n_houses, n_business, reserved;
  /* persons (network/agents.py:256-274) */
  int32_t *px, *py;
  uint8_t *status, *severity, *infected_time, *age, *stratum, *active, *isolated;
  int32_t *house, *employer, *hire_seq; /* -1 = none; hire_seq = position order in employer.employees */
  double *pwealth;
  /* houses (network/agents.py:171-184) */
  int32_t *hx, *hy, *hstratum, *hsize;
  double *hwealth, *hfixed, *hincomes, *hexpenses;
  /* businesses (network/agents.py:14-31) */
  int32_t *bx, *by, *bstratum, *bstocks, *bsales, *bnum_employees;
  double *bprice, *bwealth, *bfixed, *bincomes;
  /* government and healthcare (graph_abs.py:95-102) */
  int32_t gov_x, gov_y, health_x, health_y;
  double gov_wealth, gov_price, gov_fixed, health_wealth, health_fixed, health_expenses;
  int32_t iteration;               /* graph_abs.py:25: -1 before the first execute()              */
  int32_t reserved2;
} cabs_graph_world;

/* Host-side world builder: the construction of GraphSimulation.initialize (graph_abs.py:89-188 with
 * network/agents.py:36-42,186-194 -- every rule and quirk, in the same order over the persons) with counter-based draws
 * keyed by `seed`, in C++: a 10^5-person world in milliseconds instead of the ~1 s of the Python loop, re-entrant so that
 * the worlds of an ensemble are built concurrently.  Same distribution as the reference's worlds, not the same draws (the
 * Python mirror network/graph_abs.py::build_world keeps np.random's order for that).  Needs no GPU. */
typedef struct cabs_graph_world_spec {
  int32_t population_size, length, height, total_business, homemates_avg, homemates_std;
  double initial_infected_perc, initial_immune_perc, homeless_rate, unemployment_rate;
  double total_wealth, public_gdp_share, business_gdp_share, minimum_income, minimum_expense;
  double basic_income[5], lorenz_curve[5];   /* common.py:24-27 */
  double isolation_rate;                     /* run_batch.py:31-35 sample_isolated; < 0: nobody is isolated */
} cabs_graph_world_spec;
/* Array sizes a caller must allocate for cabs_graph_build_world (houses: an upper bound; n_houses is set on return). */
int cabs_graph_world_sizes(const cabs_graph_world_spec *spec, int32_t *n_persons, int32_t *max_houses,
                           int32_t *n_business);
int cabs_graph_build_world(const cabs_graph_world_spec *spec, uint64_t seed, cabs_graph_world *world);

int cabs_graph_create(const cabs_graph_config *cfg, cabs_graph **out);
void cabs_graph_destroy(cabs_graph *g);
const char *cabs_graph_last_error(const cabs_graph *g);
/* Replaces the lists GraphSimulation.initialize builds (graph_abs.py:89-188) for simulation `sim`. */
int cabs_graph_upload(cabs_graph *g, uint32_t sim, const cabs_graph_world *world, const cabs_graph_params *params);
/* Replaces `iterations` calls of GraphSimulation.execute (graph_abs.py:190-276) on every simulation. Asynchronous. */
int cabs_graph_run(cabs_graph *g, int iterations);
int cabs_graph_sync(cabs_graph *g);
/* Replaces get_statistics(kind='all') (graph_abs.py:302-324): rows after iterations
 * [first_iteration, first_iteration + count); first_iteration = -1 is the initial state. */
int cabs_graph_stats(cabs_graph *g, uint32_t sim, int64_t first_iteration, uint32_t count, double *rows);
/* The same rows for every simulation of the handle at once: rows[n_sims][count][CABS_GRAPH_NUM_STATS]. */
int cabs_graph_stats_all(cabs_graph *g, int64_t first_iteration, uint32_t count, double *rows);
/* Replaces the aggregation loop of batch_experiment (covid_abs/experiments.py:200-208) over the simulations of the
 * handle (its "experiments"): out[count][CABS_GRAPH_NUM_STATS][4] = Min, Avg, Std (ddof = 0), Max of every metric
 * after iterations [first_iteration, first_iteration + count), reduced on the device. */
int cabs_graph_aggregate(cabs_graph *g, int64_t first_iteration, uint32_t count, double *out);
/* The same reduction in mergeable form over simulations [sim_first, sim_first + sim_count) of the handle:
 * out[count][CABS_GRAPH_NUM_STATS][4] = min, sum, sum of squares, max.  A scenario x seed sweep sharded over several
 * GPUs reduces its part of every scenario on each device and merges the parts with min / + / + / max. */
int cabs_graph_moments(cabs_graph *g, uint32_t sim_first, uint32_t sim_count, int64_t first_iteration, uint32_t count,
                       double *out);
/* Current state of simulation `sim` (caller-allocated arrays of at least the uploaded sizes). */
int cabs_graph_download(cabs_graph *g, uint32_t sim, cabs_graph_world *world);
/* Device time of the last cabs_graph_run (CUDA events on the handle's stream); waits for it. */
int cabs_graph_last_run_ms(cabs_graph *g, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* COVIDABS_H */
