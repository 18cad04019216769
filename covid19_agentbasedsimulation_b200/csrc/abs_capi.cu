// C ABI of the B200-native COVID-ABS step (include/covidabs.h): host orchestration of the kernels
// in abs_kernels.cuh.  No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/covidabs.h"
#include "abs_kernels.cuh"

using namespace cabs;

// NVTX range around every entry point that enqueues or waits for device work (the timeline
// of a run -- uploads, steps, exchanges, statistics reads -- is readable in Nsight Systems; header-only, a no-op
// unless a tool is attached)
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

static thread_local std::string g_create_error;

struct cabs_sim {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  uint64_t capacity = 0;
  unsigned max_cells = 0;
  unsigned coarse_cells = 0;
  unsigned hist_cap = 0;
  // ping/pong population: 16-byte hot records {x, y, attr, id} + wealth; per-step displacement
  uint4 *hot[2] = {nullptr, nullptr};
  double *w[2] = {nullptr, nullptr};
  uint2 *aux = nullptr;         // {displacement, rank in destination cell}: pass A -> pass B
  // off-lattice mode (cabs_config.flags & CABS_FLAG_F64_POSITIONS): exact positions (ping/pong with hot) + the moved ones
  bool f64 = false;
  double2 *pos[2] = {nullptr, nullptr};
  double2 *npos = nullptr;
  DevRule *rules = nullptr;     // kMaxRules declarative per-agent rules: move rules first, then attribute rules
  int n_move_rules = 0, n_attr_rules = 0;
  double *sqrt_tab = nullptr;   // exact sqrt of small integers (income of abs.py:187)
  int cur = 0;
  unsigned *cells[2] = {nullptr, nullptr};  // histogram / cell_start (ping/pong), max_cells + 8 entries each
  int tab = 0;                  // table the next build uses (mirrors Ctl::cur_table)
  unsigned long long *tile_state = nullptr;  // look-back state of the scan
  unsigned *inf_list = nullptr; // slots of the Infected agents (sources of the contact pass)
  // slab decomposition
  bool slab_on = false;
  cabs_slab slab;
  bool direct = false;          // exchange through peer memory (cabs_set_slab_direct)
  unsigned char *blocks[CABS_MAX_RANKS] = {};
  unsigned char **dev_blocks = nullptr;
  unsigned build_seq = 0;
  bool fused_pack = false;      // direct mode: this build's pass A packed the emigrants itself (set by phase 0)
  bool slab_unfused = true;     // default: separate exchange kernels; CABS_SLAB_FUSED=1 folds them into pass A / one contact launch (measured slower)
  bool late_pending = false;    // direct mode: the global half of the last build's closing has not been enqueued
  unsigned late_offset = 0;     // its sequence number relative to Ctl::build_seq (0 once the local half has run)
  int late_record = 0;
  cudaGraphExec_t slab_graph = nullptr;   // two consecutive builds of the direct mode (see step_graph)
  uint64_t slab_graph_launches = 0;
  int slab_graph_cur = -1, slab_graph_tab = -1;
  unsigned slab_graph_parity = 0;
  unsigned *emig_list = nullptr, *imm_rank = nullptr;
  // sequential schedule
  unsigned long long *tinf = nullptr;
  unsigned *newinf_list = nullptr;
  unsigned *seq_round = nullptr;   // "labels lowered in round r" counters of the cooperative relaxation (3 slots)
  int seq_grid = 0;                // its grid: every block resident (0: no cooperative launch, host-driven rounds)
  cudaStream_t own_stream = nullptr;
  cudaStream_t side = nullptr;           // direct slab mode: the global half of the closing runs here, next to pass A
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool side_close = true;                // CABS_SLAB_SERIAL_CLOSE=1: the closing stays in the main stream (A/B timing)
  Ctl *ctl = nullptr;
  Ctl *host_ctl = nullptr;      // pinned mirror: statistics, error bits and trigger-rewritten attributes in ONE copy + ONE sync
  StatRecord *history = nullptr;
  uint8_t *stage = nullptr;     // byte-column staging for upload/download: 5 * capacity
  uint32_t *stage_id = nullptr;
  long long *pairs = nullptr;   // contact-pair output buffer (grown on demand)
  size_t pairs_cap = 0;
  unsigned long long *pair_count = nullptr;
  unsigned char *cross_mark = nullptr;   // cabs_cross_contact: one byte per agent (allocated on first use)
  cabs_params params;
  bool have_params = false;
  StaticParams sp;              // step-invariant parameters, passed by value to the agent passes (constant bank)
  uint64_t seed = 0;
  bool cells_dirty = true;      // cell_start does not describe the current positions / distance
  bool unsorted = false;        // the resident small-population kernel left the records out of cell order
  // two consecutive steps (one ping + one pong of the record / table copies) captured as a CUDA graph: a run of
  // steps is replayed without per-kernel launch gaps.  Valid while `cur == graph_cur`, `tab == graph_tab`.
  cudaGraphExec_t step_graph = nullptr;
  int graph_cur = -1, graph_tab = -1;
  bool graph_broken = false;    // the stream refused capture once: stay on plain launches
  bool use_pdl = false;         // chain the kernels of a step with programmatic dependent launch (CABS_PDL=1 enables)
  uint64_t n = 0, step = 0, launches = 0;
  uint64_t hist_first = 0;      // first step whose statistics row the ring can hold (moved by cabs_set_step)
  cudaEvent_t events[CABS_NUM_EVENTS] = {};
  // per-kernel profiling: events[i] is recorded before launch i of the current batch, names[i] its kernel
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<const char *> prof_names;
  size_t prof_used = 0;
  std::string err;
};

static void prof_mark(cabs_sim *sim, const char *name) {
  if (sim->prof_used == sim->prof_events.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    sim->prof_events.push_back(e);
    sim->prof_names.push_back(name);
  }
  sim->prof_names[sim->prof_used] = name;
  cudaEventRecord(sim->prof_events[sim->prof_used], sim->stream);
  sim->prof_used++;
}

#define CK(call)                                                                                \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      char buf_[512];                                                                           \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      sim->err = buf_;                                                                          \
      return CABS_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

static int flush_late_close(cabs_sim *sim);

static void drop_graphs(cabs_sim *sim) {
  if (sim->step_graph) { cudaGraphExecDestroy(sim->step_graph); sim->step_graph = nullptr; }
  if (sim->slab_graph) { cudaGraphExecDestroy(sim->slab_graph); sim->slab_graph = nullptr; }
}

static int fail(cabs_sim *sim, int code, const char *msg) {
  if (sim) sim->err = msg;
  else g_create_error = msg;
  return code;
}

static inline int agent_blocks(const cabs_sim *sim) { return sim->num_sms * 8; }

#define LAUNCH_AS(name, kernel, grid, block, ...)        \
  do {                                                  \
    if (sim->profiling) prof_mark(sim, name);           \
    kernel<<<(grid), (block), 0, sim->stream>>>(__VA_ARGS__); \
    sim->launches++;                                    \
  } while (0)
#define LAUNCH(kernel, grid, block, ...) LAUNCH_AS(#kernel, kernel, grid, block, __VA_ARGS__)
#define LAUNCH_SMEM_AS(name, kernel, grid, block, smem, ...)   \
  do {                                                        \
    if (sim->profiling) prof_mark(sim, name);                 \
    kernel<<<(grid), (block), (smem), sim->stream>>>(__VA_ARGS__); \
    sim->launches++;                                          \
  } while (0)
// The same with programmatic dependent launch on the preceding kernel of the stream (see pdl_wait in abs_kernels.cuh).
#define LAUNCH_PDL_AS(name, kernel, grid, block, ...)                          \
  do {                                                                        \
    if (sim->profiling) prof_mark(sim, name);                                 \
    cudaLaunchConfig_t cfg_ = {};                                             \
    cfg_.gridDim = dim3(grid);                                                \
    cfg_.blockDim = dim3(block);                                              \
    cfg_.stream = sim->stream;                                                \
    cudaLaunchAttribute attr_[1];                                             \
    attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;        \
    attr_[0].val.programmaticStreamSerializationAllowed = sim->use_pdl ? 1 : 0; \
    cfg_.attrs = attr_;                                                       \
    cfg_.numAttrs = 1;                                                        \
    cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                           \
    sim->launches++;                                                          \
  } while (0)
#define LAUNCH_PDLN(kernel, grid, block, ...) LAUNCH_PDL_AS(#kernel, kernel, grid, block, __VA_ARGS__)

extern "C" const char *cabs_last_error(const cabs_sim *sim) {
  return sim ? sim->err.c_str() : g_create_error.c_str();
}

extern "C" void cabs_destroy(cabs_sim *sim) {
  if (!sim) return;
  cudaSetDevice(sim->device);
  if (sim->stream) cudaStreamSynchronize(sim->stream);
  if (sim->own_stream) sim->stream = sim->own_stream;
  for (int b = 0; b < 2; ++b) {
    cudaFree(sim->hot[b]); cudaFree(sim->w[b]);
  }
  cudaFree(sim->aux);
  cudaFree(sim->pos[0]); cudaFree(sim->pos[1]); cudaFree(sim->npos); cudaFree(sim->rules);
  cudaFree(sim->sqrt_tab);
  cudaFree(sim->cells[0]); cudaFree(sim->cells[1]); cudaFree(sim->tile_state); cudaFree(sim->inf_list);
  cudaFree(sim->emig_list); cudaFree(sim->imm_rank); cudaFree(sim->tinf); cudaFree(sim->newinf_list);
  cudaFree(sim->seq_round);
  cudaFree(sim->dev_blocks);
  cudaFree(sim->ctl); cudaFree(sim->history);
  cudaFreeHost(sim->host_ctl);
  cudaFree(sim->stage); cudaFree(sim->stage_id); cudaFree(sim->pairs); cudaFree(sim->pair_count);
  cudaFree(sim->cross_mark);
  if (sim->step_graph) cudaGraphExecDestroy(sim->step_graph);
  if (sim->slab_graph) cudaGraphExecDestroy(sim->slab_graph);
  for (cudaEvent_t e : sim->events) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : sim->prof_events) cudaEventDestroy(e);
  if (sim->side) cudaStreamDestroy(sim->side);
  if (sim->ev_fork) cudaEventDestroy(sim->ev_fork);
  if (sim->ev_join) cudaEventDestroy(sim->ev_join);
  if (sim->stream) cudaStreamDestroy(sim->stream);
  delete sim;
}

static int create_impl(cabs_sim *sim, const cabs_config *cfg) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(sim, CABS_ERR_NOGPU, "no CUDA device visible: the agent step has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(sim, CABS_ERR_INVALID, "device ordinal out of range");
  sim->device = cfg->device;
  CK(cudaSetDevice(sim->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, sim->device));
  sim->num_sms = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&sim->stream, cudaStreamNonBlocking));
  sim->own_stream = sim->stream;
  CK(cudaStreamCreateWithFlags(&sim->side, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&sim->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&sim->ev_join, cudaEventDisableTiming));
  // pass B: shared memory for CABS_B_MINBLOCKS resident blocks (ring + partial sums + Infected list + 2 KB of static
  // and reserved space each) and the REST of the 256 KB to L1 -- its cell-start gather is the most expensive part of
  // the kernel (measured without it: 0.109 ms against 0.144 ms) and lives on L1 hits
  {
#ifdef CABS_B_CARVEOUT
    const int pct = CABS_B_CARVEOUT;
#else
    const size_t want = (size_t)CABS_B_MINBLOCKS * (kPassBSmem + 2048);
    const int pct = (int)std::min<size_t>(100, (want * 100 + prop.sharedMemPerMultiprocessor - 1) / prop.sharedMemPerMultiprocessor);
#endif
    CK(cudaFuncSetAttribute(k_update_scatter<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    CK(cudaFuncSetAttribute(k_update_scatter<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    CK(cudaFuncSetAttribute(k_update_scatter<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    CK(cudaFuncSetAttribute(k_update_scatter<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
  }
  CK(cudaFuncSetAttribute(k_update_scatter<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPassBSmem));
  CK(cudaFuncSetAttribute(k_update_scatter<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPassBSmem));
  CK(cudaFuncSetAttribute(k_update_scatter<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPassBSmem));
  CK(cudaFuncSetAttribute(k_update_scatter<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPassBSmem));
  memset(&sim->sp, 0, sizeof sim->sp);
  sim->seed = cfg->seed;
  {   // measured on B200 (gpurun_out/session7.log): 0.314 ms/step with the programmatic edges against 0.279 ms/step
      // without -- the parked blocks of the successor take registers and shared memory from the kernel that is still
      // running, which costs more than the launch gaps inside the captured graph were worth.  Opt-in only.
    const char *pdl = getenv("CABS_PDL");
    sim->use_pdl = pdl && pdl[0] == '1';
    // Folding the emigrants' update + pack into pass A and the ghost contacts into the local contact launch saves two
    // launches per step but doubles pass A's registers: 2 GPUs 0.391 ms/step against 0.297 ms.
    const char *serial = getenv("CABS_SLAB_SERIAL_CLOSE");
    sim->side_close = !(serial && serial[0] == '1');
    const char *fused = getenv("CABS_SLAB_FUSED");
    sim->slab_unfused = !(fused && fused[0] == '1');
  }
  sim->sp.k0 = (unsigned)(cfg->seed & 0xFFFFFFFFu);
  sim->sp.k1 = (unsigned)(cfg->seed >> 32);
  sim->sp.slab_lo = INT_MIN;
  sim->sp.slab_hi = INT_MAX;
  sim->capacity = cfg->capacity ? cfg->capacity : 1;
  if (sim->capacity >= (1ull << 31)) return fail(sim, CABS_ERR_CAPACITY, "capacity must be < 2^31 agents per device");
  // Two budgets for the cell table.  While few agents can transmit (or few can be infected) the contact pass is
  // cheap whatever the cell size, and a coarse grid (one cell per 8..32 agents) keeps the table small for the
  // scan and the scatter.  In a dense epidemic the contact pass visits every candidate of every source, so the
  // closing of a step switches to the fine budget (about one cell per agent).  Results do not depend on the grid.
  uint64_t mc = cfg->max_cells, coarse = 0;
  if (mc == 0) {
    mc = sim->capacity + sim->capacity / 2;
    if (mc > (1ull << 24)) mc = 1ull << 24;
    coarse = sim->capacity / 8;
  } else {
    coarse = mc;             // an explicit cap is used as is, in both regimes
  }
  if (mc < 4096 && cfg->max_cells == 0) mc = 4096;
  if (mc < 64) mc = 64;
  if (mc > (1ull << 30)) mc = 1ull << 30;
  if (coarse < 4096 && cfg->max_cells == 0) coarse = 4096;
  if (coarse < 64) coarse = 64;
  if (coarse > mc) coarse = mc;
  sim->max_cells = (unsigned)mc;
  sim->coarse_cells = (unsigned)coarse;
  sim->hist_cap = cfg->history_capacity ? cfg->history_capacity : 4096;
  const size_t cap = sim->capacity;
  const size_t capp = cap + kPad;  // the agent passes load whole tiles
  for (int b = 0; b < 2; ++b) {
    CK(cudaMalloc(&sim->hot[b], capp * sizeof(uint4)));
    CK(cudaMalloc(&sim->w[b], capp * sizeof(double)));
    CK(cudaMemsetAsync(sim->hot[b], 0, capp * sizeof(uint4), sim->stream));
    CK(cudaMemsetAsync(sim->w[b], 0, capp * sizeof(double), sim->stream));
  }
  CK(cudaMalloc(&sim->aux, capp * sizeof(uint2)));
  CK(cudaMemsetAsync(sim->aux, 0, capp * sizeof(uint2), sim->stream));
  sim->f64 = (cfg->flags & CABS_FLAG_F64_POSITIONS) != 0;
  if (sim->f64) {
    for (int b = 0; b < 2; ++b) {
      CK(cudaMalloc(&sim->pos[b], capp * sizeof(double2)));
      CK(cudaMemsetAsync(sim->pos[b], 0, capp * sizeof(double2), sim->stream));
    }
    CK(cudaMalloc(&sim->npos, capp * sizeof(double2)));
    CK(cudaMalloc(&sim->rules, kMaxRules * sizeof(DevRule)));
    CK(cudaMemsetAsync(sim->rules, 0, kMaxRules * sizeof(DevRule), sim->stream));
  }
  CK(cudaMalloc(&sim->inf_list, cap * sizeof(unsigned)));
  CK(cudaMalloc(&sim->sqrt_tab, kSqrtTable * sizeof(double)));
  for (int b = 0; b < 2; ++b) {
    CK(cudaMalloc(&sim->cells[b], ((size_t)sim->max_cells + 8) * sizeof(unsigned)));
    CK(cudaMemsetAsync(sim->cells[b], 0, ((size_t)sim->max_cells + 8) * sizeof(unsigned), sim->stream));
  }
  CK(cudaMalloc(&sim->tile_state, ((size_t)sim->max_cells / kScanTile + 2) * sizeof(unsigned long long)));
  CK(cudaMalloc(&sim->ctl, sizeof(Ctl)));
  CK(cudaHostAlloc(&sim->host_ctl, sizeof(Ctl), cudaHostAllocDefault));
  CK(cudaMalloc(&sim->history, (size_t)sim->hist_cap * sizeof(StatRecord)));
  CK(cudaMalloc(&sim->stage, cap * 5));
  CK(cudaMalloc(&sim->stage_id, cap * sizeof(uint32_t)));
  CK(cudaMalloc(&sim->pair_count, sizeof(unsigned long long)));
  CK(cudaMemsetAsync(sim->ctl, 0, sizeof(Ctl), sim->stream));
  CK(cudaMemsetAsync(sim->history, 0, (size_t)sim->hist_cap * sizeof(StatRecord), sim->stream));
  LAUNCH(k_init_ctl, 1, 256, sim->ctl, (unsigned)(cfg->seed & 0xFFFFFFFFu), (unsigned)(cfg->seed >> 32),
         sim->max_cells, sim->coarse_cells, sim->hist_cap, (unsigned)sim->capacity, sim->sqrt_tab);
  if (sim->f64) LAUNCH(k_set_rules, 1, 32, sim->ctl, 1, 0, 1, 0, 1, 0);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(sim->stream));
  return CABS_OK;
}

extern "C" int cabs_create(const cabs_config *cfg, cabs_sim **out) {
  if (!cfg || !out) return fail(nullptr, CABS_ERR_INVALID, "cabs_create: null argument");
  if (cfg->abi_version != CABS_ABI_VERSION) return fail(nullptr, CABS_ERR_INVALID, "cabs_create: ABI version mismatch");
  cabs_sim *sim = new (std::nothrow) cabs_sim();
  if (!sim) return fail(nullptr, CABS_ERR_INVALID, "cabs_create: out of host memory");
  const int rc = create_impl(sim, cfg);
  if (rc != CABS_OK) {
    g_create_error = sim->err;
    cabs_destroy(sim);
    *out = nullptr;
    return rc;
  }
  *out = sim;
  return CABS_OK;
}

extern "C" int cabs_set_params(cabs_sim *sim, const cabs_params *p) {
  NvtxRange nvtx_("cabs_set_params");
  if (!sim || !p) return fail(sim, CABS_ERR_INVALID, "cabs_set_params: null argument");
  if (p->wealth_scale_bits < 0 || p->wealth_scale_bits > 52)
    return fail(sim, CABS_ERR_INVALID, "wealth_scale_bits must be in [0, 52]");
  if (p->recovering_time < 0 || p->recovering_time > 250)
    return fail(sim, CABS_ERR_INVALID, "recovering_time must be in [0, 250] (infected_time is stored in 8 bits)");
  if (p->schedule != CABS_SYNCHRONOUS && p->schedule != CABS_SEQUENTIAL)
    return fail(sim, CABS_ERR_INVALID, "schedule must be CABS_SYNCHRONOUS or CABS_SEQUENTIAL");
  if (p->schedule == CABS_SEQUENTIAL && sim->slab_on)
    return fail(sim, CABS_ERR_INVALID, "the sequential schedule is defined on one device only (SURVEY.md 8e)");
  if (p->population_size == 0) return fail(sim, CABS_ERR_INVALID, "population_size must be > 0");
  for (int k = 0; k < 3; ++k)
    if (!(p->amplitudes[k] >= -5000.0 && p->amplitudes[k] <= 5000.0))
      return fail(sim, CABS_ERR_INVALID, "amplitudes must be within +-5000 (displacements are stored in 16 bits)");
  CK(cudaSetDevice(sim->device));
  {
    const int rc = flush_late_close(sim);   // the critical-limit flag is recomputed from the current statistics
    if (rc != CABS_OK) return rc;
  }
  ParamBlock b;
  for (int k = 0; k < 4; ++k) b.amp[k] = (float)p->amplitudes[k];
  b.amp[3] = 0.f;
  b.length = p->length; b.height = p->height; b.contagion_distance = p->contagion_distance;
  b.rate = p->contagion_rate; b.critical_limit = p->critical_limit; b.min_expense = p->minimum_expense;
  for (int k = 0; k < 5; ++k) b.bi[k] = p->basic_income[k];
  for (int k = 0; k < 9; ++k) {
    b.p_hosp[k] = p->age_hospitalization_probs[k];
    b.p_sev[k] = p->age_severe_probs[k];
    b.p_death[k] = p->age_death_probs[k];
  }
  b.pop_size = p->population_size;
  b.recovering_time = p->recovering_time; b.scale_bits = p->wealth_scale_bits; b.schedule = p->schedule;
  if (p->schedule == CABS_SEQUENTIAL && !sim->tinf) {
    CK(cudaMalloc(&sim->tinf, (sim->capacity + kPad) * sizeof(unsigned long long)));
    CK(cudaMalloc(&sim->newinf_list, (sim->capacity + kPad) * sizeof(unsigned)));
    CK(cudaMalloc(&sim->seq_round, 4 * sizeof(unsigned)));
    CK(cudaMemsetAsync(sim->seq_round, 0, 4 * sizeof(unsigned), sim->stream));
    // all blocks of the cooperative relaxation kernel must be resident at once
    int coop = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, sim->device);
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_contagion_sequential_all, 256, 0) == cudaSuccess)
      sim->seq_grid = per_sm * sim->num_sms;
  }
  {   // the same values load_params derives on the device, for the by-value copy the agent passes take
    StaticParams &sp = sim->sp;
    sp.min_expense = p->minimum_expense;
    sp.me32 = p->minimum_expense * 0x1p-32;
    for (int k = 0; k < 5; ++k) {
      sp.bi[k] = p->basic_income[k];
      sp.mebi[k] = p->minimum_expense * p->basic_income[k];
    }
    for (int k = 0; k < 9; ++k) {
      sp.thr_hosp[k] = prob_bound_lt(p->age_hospitalization_probs[k]);
      sp.thr_sev[k] = prob_bound_lt(p->age_severe_probs[k]);
      sp.thr_death[k] = prob_bound_lt(p->age_death_probs[k]);
    }
    sp.recovering_time = p->recovering_time;
    sp.scale_bits = p->wealth_scale_bits;
    sp.length = p->length;
    sp.height = p->height;
    sp.len_i = (int)fmin(fmax(ceil(p->length), -2147483647.0), 2147483647.0);
    sp.hei_i = (int)fmin(fmax(ceil(p->height), -2147483647.0), 2147483647.0);
    drop_graphs(sim);   // captured launches hold the old copy
  }
  LAUNCH(k_set_params, 1, 32, sim->ctl, b);
  if (sim->n > 0) LAUNCH(k_replan, 1, 32, sim->ctl, 0);  // amplitudes / distance may have changed the margin
  CK(cudaGetLastError());
  sim->params = *p;
  sim->have_params = true;
  sim->cells_dirty = true;
  return CABS_OK;
}

extern "C" int cabs_get_params(cabs_sim *sim, cabs_params *out) {
  if (!sim || !out) return fail(sim, CABS_ERR_INVALID, "cabs_get_params: null argument");
  if (!sim->have_params) return fail(sim, CABS_ERR_INVALID, "cabs_get_params: call cabs_set_params first");
  CK(cudaSetDevice(sim->device));
  {
    const int rc = flush_late_close(sim);
    if (rc != CABS_OK) return rc;
  }
  Ctl h;
  CK(cudaMemcpyAsync(&h, sim->ctl, sizeof h, cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  *out = sim->params;
  // what device-side triggers (abs.py:267-271) may have rewritten since the last cabs_set_params
  for (int k = 0; k < 4; ++k) out->amplitudes[k] = (double)h.amp[k];
  out->contagion_rate = h.rate;
  out->contagion_distance = h.contagion_distance;
  out->critical_limit = h.critical_limit;
  sim->params = *out;
  return CABS_OK;
}

extern "C" int cabs_set_triggers(cabs_sim *sim, const cabs_trigger *t, int count) {
  if (!sim || (count > 0 && !t)) return fail(sim, CABS_ERR_INVALID, "cabs_set_triggers: null argument");
  if (count < 0 || count > CABS_MAX_TRIGGERS) return fail(sim, CABS_ERR_INVALID, "too many triggers");
  CK(cudaSetDevice(sim->device));
  TriggerBlock b;
  memset(&b, 0, sizeof b);
  b.ntrig = count;
  for (int k = 0; k < count; ++k) {
    if (t[k].metric < 0 || t[k].metric >= CABS_NUM_STATS || t[k].op < 0 || t[k].op > 3 || t[k].attribute < 0 ||
        t[k].attribute > 3)
      return fail(sim, CABS_ERR_INVALID, "cabs_set_triggers: bad metric/op/attribute");
    b.trig[k].metric = t[k].metric; b.trig[k].op = t[k].op; b.trig[k].attribute = t[k].attribute;
    b.trig[k].threshold = t[k].threshold;
    for (int j = 0; j < 4; ++j) b.trig[k].value[j] = t[k].value[j];
    if (t[k].attribute == CABS_ATTR_AMPLITUDES)
      for (int j = 0; j < 3; ++j)
        if (!(t[k].value[j] >= -5000.0 && t[k].value[j] <= 5000.0))
          return fail(sim, CABS_ERR_INVALID, "trigger amplitudes must be within +-5000");
  }
  LAUNCH(k_set_triggers, 1, 32, sim->ctl, b);
  CK(cudaGetLastError());
  return CABS_OK;
}

// One build of the cell-sorted population: kAdvance = a real step, otherwise the cell list of the
// current positions (upload / contact-pair hook) without advancing the simulation.
template <bool kAdvance>
static void enqueue_build(cabs_sim *sim) {
  const int c = sim->cur, o = c ^ 1, t = sim->tab;
  const bool sequential = sim->have_params && sim->params.schedule == CABS_SEQUENTIAL;
  const int nb = agent_blocks(sim);
  SlabDirect no_direct;
  memset(&no_direct, 0, sizeof no_direct);
  const double2 *pos_o = sim->f64 ? sim->pos[o] : nullptr;   // exact positions the contact kernels test (off-lattice mode)
  if (sim->f64) {
    LAUNCH_AS(kAdvance ? "k_move_count_f64<true>" : "k_move_count_f64<false>", k_move_count_f64<kAdvance>, nb, 256,
              sim->ctl, sim->hot[c], sim->pos[c], sim->npos, sim->aux, sim->cells[t], sim->tile_state, sim->rules,
              sim->n_move_rules, sim->sp);
    LAUNCH(k_scan_cells, sim->num_sms * 3, kScanThreads, sim->ctl, sim->cells[t], sim->tile_state);
    LAUNCH_AS(kAdvance ? "k_update_scatter_f64<true>" : "k_update_scatter_f64<false>", k_update_scatter_f64<kAdvance>, nb,
              256, sim->ctl, sim->hot[c], sim->w[c], sim->aux, sim->npos, sim->cells[t], sim->cells[t ^ 1], sim->hot[o],
              sim->w[o], sim->pos[o], sim->inf_list, sim->sqrt_tab, (kAdvance && sequential) ? sim->tinf : nullptr,
              sim->rules, sim->n_move_rules, sim->n_attr_rules, sim->sp);
  } else {
  LAUNCH_AS(kAdvance ? "k_move_count<true>" : "k_move_count<false>", (k_move_count<kAdvance, false>), nb, 256, sim->ctl,
            sim->hot[c], sim->aux, sim->cells[t], sim->tile_state, sim->emig_list, sim->sp, nullptr, nullptr, nullptr,
            nullptr, SlabDirect(), 0);
  LAUNCH_PDL_AS("k_scan_cells", k_scan_cells, sim->num_sms * 3, kScanThreads, sim->ctl, sim->cells[t], sim->tile_state);
  }
  if (sim->f64) {
    // pass B ran above
  } else if (kAdvance && sequential) {
    LAUNCH_SMEM_AS("k_update_scatter<true>", (k_update_scatter<kAdvance, true>), nb, kAgentThreads, kPassBSmem, sim->ctl, sim->hot[c], sim->w[c],
              sim->aux, sim->cells[t], sim->cells[t ^ 1], sim->hot[o], sim->w[o], sim->inf_list, sim->sqrt_tab, nullptr,
              nullptr, sim->tinf, nullptr, nullptr, nullptr, no_direct, sim->sp);
  } else {
    LAUNCH_SMEM_AS(kAdvance ? "k_update_scatter<true>" : "k_update_scatter<false>", (k_update_scatter<kAdvance, false>),
                  nb, kAgentThreads, kPassBSmem, sim->ctl, (const uint4 *)sim->hot[c], (const double *)sim->w[c],
                  (const uint2 *)sim->aux, (const unsigned *)sim->cells[t], sim->cells[t ^ 1], sim->hot[o], sim->w[o],
                  sim->inf_list, (const double *)sim->sqrt_tab, (ExchangeHeader *)nullptr, (ExchangeHeader *)nullptr,
                  (unsigned long long *)nullptr, (const ExchangeHeader *)nullptr, (const ExchangeHeader *)nullptr,
                  (const unsigned *)nullptr, no_direct, sim->sp);
  }
  if (kAdvance && sequential && sim->seq_grid > 0) {
    // every round of the label-correcting relaxation inside one cooperative launch (grid-wide barriers between the
    // rounds); its block 0 closes the step
    if (sim->profiling) prof_mark(sim, "k_contagion_sequential_all");
    uint4 *hot_o = sim->hot[o];
    unsigned *cells_t = sim->cells[t];
    void *args[] = {&sim->ctl, &hot_o, &cells_t, &sim->tinf, &sim->inf_list, &sim->newinf_list, &sim->history,
                    &sim->seq_round, &pos_o};
    cudaLaunchCooperativeKernel((void *)k_contagion_sequential_all, dim3(sim->seq_grid), dim3(256), args, 0, sim->stream);
    sim->launches++;
  } else if (kAdvance && sequential) {
    // no cooperative launch on this device: rounds driven from the host (one round trip each)
    for (int round = 0;; ++round) {
      LAUNCH(k_contagion_sequential, nb, 256, sim->ctl, sim->hot[o], sim->cells[t], sim->tinf,
             round == 0 ? sim->inf_list : sim->newinf_list, round == 0 ? 1 : 0, sim->newinf_list, pos_o);
      unsigned changed = 0;
      if (cudaMemcpyAsync(&changed, &sim->ctl->seq_changed, sizeof changed, cudaMemcpyDeviceToHost, sim->stream) !=
              cudaSuccess || cudaStreamSynchronize(sim->stream) != cudaSuccess)
        break;
      if (!changed) break;
      cudaMemsetAsync(&sim->ctl->seq_changed, 0, sizeof(unsigned), sim->stream);
    }
    LAUNCH(k_finalize, 1, 32, sim->ctl, sim->history, 1);
  } else if (kAdvance) {
    SlabDirect none;
    memset(&none, 0, sizeof none);
    LAUNCH_PDL_AS("k_contagion", k_contagion, nb, 256, sim->ctl, sim->hot[o], (const unsigned *)sim->cells[t],
                  (const unsigned *)sim->inf_list, (const ExchangeHeader *)nullptr, (const ExchangeHeader *)nullptr,
                  sim->history, (long long *)nullptr, 1, none, (ExchangeHeader *)nullptr, pos_o);
  } else {
    LAUNCH(k_finalize, 1, 32, sim->ctl, sim->history, 0);
  }
  sim->cur = o;
  sim->tab = t ^ 1;
}

static int rebuild_static(cabs_sim *sim) {
  const int nb = agent_blocks(sim);
  LAUNCH(k_bbox, nb, 256, sim->ctl, sim->hot[sim->cur]);
  LAUNCH(k_replan, 1, 32, sim->ctl, 1);
  enqueue_build<false>(sim);
  CK(cudaGetLastError());
  sim->cells_dirty = false;
  sim->unsorted = false;
  return CABS_OK;
}


// ---- slab decomposition -------------------------------------------------------------------------
extern "C" int cabs_set_stream(cabs_sim *sim, void *cuda_stream) {
  if (!sim) return fail(sim, CABS_ERR_INVALID, "cabs_set_stream: null handle");
  CK(cudaSetDevice(sim->device));
  CK(cudaStreamSynchronize(sim->stream));
  sim->stream = (cudaStream_t)cuda_stream;  // NULL is CUDA's default stream
  return CABS_OK;
}

extern "C" int cabs_set_slab(cabs_sim *sim, const cabs_slab *slab) {
  if (!sim || !slab) return fail(sim, CABS_ERR_INVALID, "cabs_set_slab: null argument");
  if (slab->n_ranks < 1 || slab->rank < 0 || slab->rank >= slab->n_ranks || slab->x_lo >= slab->x_hi)
    return fail(sim, CABS_ERR_INVALID, "cabs_set_slab: bad rank / bounds");
  if (sim->f64) return fail(sim, CABS_ERR_INVALID, "cabs_set_slab: the off-lattice mode runs on one device");
  if (slab->migrant_capacity == 0 || slab->ghost_capacity == 0 || !slab->stats_rows)
    return fail(sim, CABS_ERR_INVALID, "cabs_set_slab: capacities and stats_rows are required");
  for (int d = 0; d < 2; ++d)
    if (!slab->migrant_send[d] || !slab->migrant_recv[d] || !slab->ghost_send[d] || !slab->ghost_recv[d])
      return fail(sim, CABS_ERR_INVALID, "cabs_set_slab: null exchange buffer");
  if (sim->have_params && sim->params.schedule == CABS_SEQUENTIAL)
    return fail(sim, CABS_ERR_INVALID, "the sequential schedule is defined on one device only (SURVEY.md 8e)");
  CK(cudaSetDevice(sim->device));
  CK(cudaStreamSynchronize(sim->stream));
  cudaFree(sim->emig_list); cudaFree(sim->imm_rank); cudaFree(sim->tinf); cudaFree(sim->newinf_list);
  cudaFree(sim->dev_blocks);
  sim->emig_list = sim->imm_rank = nullptr;
  cudaFree(sim->seq_round);
  sim->tinf = nullptr;
  sim->newinf_list = nullptr;
  sim->seq_round = nullptr;
  sim->seq_grid = 0;
  sim->dev_blocks = nullptr;
  CK(cudaMalloc(&sim->emig_list, 2 * (size_t)slab->migrant_capacity * sizeof(unsigned)));
  CK(cudaMalloc(&sim->imm_rank, 2 * (size_t)slab->migrant_capacity * sizeof(unsigned)));
  for (int d = 0; d < 2; ++d) {  // headers start at count 0; outer-side receive buffers are never written again
    CK(cudaMemsetAsync(slab->migrant_send[d], 0, sizeof(ExchangeHeader), sim->stream));
    CK(cudaMemsetAsync(slab->migrant_recv[d], 0, sizeof(ExchangeHeader), sim->stream));
    CK(cudaMemsetAsync(slab->ghost_send[d], 0, sizeof(ExchangeHeader), sim->stream));
    CK(cudaMemsetAsync(slab->ghost_recv[d], 0, sizeof(ExchangeHeader), sim->stream));
  }
  sim->slab = *slab;
  sim->slab_on = true;
  sim->direct = false;
  sim->sp.slab_lo = slab->x_lo;
  sim->sp.slab_hi = slab->x_hi;
  drop_graphs(sim);
  LAUNCH(k_set_slab, 1, 32, sim->ctl, slab->rank, slab->n_ranks, slab->x_lo, slab->x_hi, slab->migrant_capacity,
         slab->ghost_capacity);
  if (sim->n > 0) LAUNCH(k_replan, 1, 32, sim->ctl, 0);
  CK(cudaGetLastError());
  sim->cells_dirty = true;
  return CABS_OK;
}

// Direct mode: enqueue the global half of the last build's closing (waits on the device for every rank's row).
static int flush_late_close(cabs_sim *sim) {
  if (!sim->late_pending) return CABS_OK;
  SlabDirect D;
  memset(&D, 0, sizeof D);
  D.self = sim->blocks[sim->slab.rank];
  D.seq = sim->late_offset;
  D.enabled = 1;
  LAUNCH_PDLN(k_slab_close_global, 1, 32, sim->ctl, sim->history, sim->late_record, D);
  sim->late_pending = false;
  CK(cudaGetLastError());
  return CABS_OK;
}

template <bool kAdvance>
static int slab_phase(cabs_sim *sim, int phase) {
  const cabs_slab &sl = sim->slab;
  const int c = sim->cur, o = c ^ 1, t = sim->tab;
  const int nb = agent_blocks(sim), nbs = sim->num_sms;  // the exchange kernels see few records
  ExchangeHeader *ms[2], *mr[2], *gs[2], *gr[2];
  long long *rows;
  SlabDirect D;
  memset(&D, 0, sizeof D);
  if (sim->direct) {
    const BlockLayout L = block_layout(sl.n_ranks, sl.migrant_capacity, sl.ghost_capacity);
    unsigned char *self = sim->blocks[sl.rank];
    const unsigned seq = sim->build_seq + 1, par = seq & 1u;
    for (int d = 0; d < 2; ++d) {
      ms[d] = (ExchangeHeader *)(self + L.mig_send + d * L.mig_bytes);
      gs[d] = (ExchangeHeader *)(self + L.ghost_send + d * L.ghost_bytes);
      mr[d] = (ExchangeHeader *)(self + L.mig_recv + (d * 2 + par) * L.mig_bytes);
      gr[d] = (ExchangeHeader *)(self + L.ghost_recv + (d * 2 + par) * L.ghost_bytes);
    }
    rows = (long long *)(self + L.rows) + (size_t)par * kMaxRanks * kSlabRow;
    D.self = self;
    D.left = sl.rank > 0 ? sim->blocks[sl.rank - 1] : nullptr;
    D.right = sl.rank + 1 < sl.n_ranks ? sim->blocks[sl.rank + 1] : nullptr;
    D.all = sim->dev_blocks;
    D.seq = 1;   // relative to Ctl::build_seq: the build in flight
    D.enabled = 1;
  } else {
    for (int d = 0; d < 2; ++d) {
      ms[d] = (ExchangeHeader *)sl.migrant_send[d]; mr[d] = (ExchangeHeader *)sl.migrant_recv[d];
      gs[d] = (ExchangeHeader *)sl.ghost_send[d]; gr[d] = (ExchangeHeader *)sl.ghost_recv[d];
    }
    rows = (long long *)sl.stats_rows;
  }
  switch (phase) {
    case 0: {
      if (!kAdvance) {
        LAUNCH_PDLN(k_bbox, nb, 256, sim->ctl, sim->hot[c]);
        LAUNCH_PDLN(k_replan, 1, 32, sim->ctl, 1);
      }
      // Steady state of the direct mode (the build before was a step whose global closing is still pending, or nothing
      // is pending): pass A updates and sends the emigrants itself, and the global closing rides on
      // k_count_immigrants -- two launches fewer per step (opt-in, CABS_SLAB_FUSED=1).
      sim->fused_pack = kAdvance && sim->direct && !sim->slab_unfused &&
                        (!sim->late_pending || (sim->late_offset == 0 && sim->late_record == 1));
      if (sim->fused_pack) {
        LAUNCH_PDL_AS("k_move_count<true,pack>", (k_move_count<kAdvance, true>), nb, 256, sim->ctl, sim->hot[c], sim->aux,
                      sim->cells[t], sim->tile_state, sim->emig_list, sim->sp, (const double *)sim->w[c], ms[0], ms[1],
                      (const double *)sim->sqrt_tab, D, sim->late_pending ? 1 : 0);
        break;
      }
      // Pass A needs nothing global; pass B and the emigrants' update need the critical-limit flag the global half of
      // the previous step's closing sets.  That half (one thread waiting for every rank's statistics row) runs on a
      // second stream NEXT TO pass A -- forked and joined with events, so a captured graph carries it as a parallel
      // branch -- instead of between pass A and the pack: its wait for the slowest rank is off the critical path.
      const bool fork_close = sim->direct && sim->late_pending && sim->side_close && !sim->profiling && !sim->use_pdl;
      if (fork_close) {
        SlabDirect G;
        memset(&G, 0, sizeof G);
        G.self = sim->blocks[sl.rank];
        G.seq = sim->late_offset;
        G.enabled = 1;
        CK(cudaEventRecord(sim->ev_fork, sim->stream));
        CK(cudaStreamWaitEvent(sim->side, sim->ev_fork, 0));
        k_slab_close_global<<<1, 32, 0, sim->side>>>(sim->ctl, sim->history, sim->late_record, G);
        sim->launches++;
        CK(cudaEventRecord(sim->ev_join, sim->side));
        sim->late_pending = false;
      }
      LAUNCH_PDL_AS(kAdvance ? "k_move_count<true>" : "k_move_count<false>", (k_move_count<kAdvance, false>), nb, 256,
                    sim->ctl, sim->hot[c], sim->aux, sim->cells[t], sim->tile_state, sim->emig_list, sim->sp, nullptr,
                    nullptr, nullptr, nullptr, SlabDirect(), 0);
      if (fork_close) {
        CK(cudaStreamWaitEvent(sim->stream, sim->ev_join, 0));
      } else {
        const int rc = flush_late_close(sim);
        if (rc != CABS_OK) return rc;
      }
      LAUNCH_PDL_AS("k_pack_emigrants", k_pack_emigrants<kAdvance>, nbs, 256, sim->ctl, sim->hot[c], sim->w[c], sim->aux,
                sim->emig_list, ms[0], ms[1], sim->sqrt_tab, D);
      break;
    }
    case 1: {
      SlabDirect C;   // the rows the global closing reads: the build that was closed locally (sequence offset 0)
      memset(&C, 0, sizeof C);
      int close_first = 0;
      if (sim->fused_pack && sim->late_pending) {
        C.self = sim->blocks[sl.rank];
        C.enabled = 1;
        close_first = 1;
        sim->late_pending = false;
      }
      LAUNCH_PDLN(k_count_immigrants, nbs, 256, sim->ctl, mr[0], mr[1], sim->cells[t], sim->imm_rank, D, close_first, C,
                  sim->history);
      LAUNCH_PDLN(k_scan_cells, sim->num_sms * 3, kScanThreads, sim->ctl, sim->cells[t], sim->tile_state);
      LAUNCH_SMEM_AS(kAdvance ? "k_update_scatter<true>" : "k_update_scatter<false>", (k_update_scatter<kAdvance, true>), nb,
                kAgentThreads, kPassBSmem, sim->ctl, sim->hot[c], sim->w[c], sim->aux, sim->cells[t], sim->cells[t ^ 1],
                sim->hot[o], sim->w[o], sim->inf_list, sim->sqrt_tab, gs[0], gs[1], nullptr, mr[0], mr[1],
                sim->imm_rank, D, sim->sp);   // the immigrants take their slots inside pass B; its last block pushes the ghosts
      break;
    }
    case 2:
      if (kAdvance) {
        if (sim->direct && !sim->slab_unfused) {
          // one launch: every block pushes its share of the local sources first, which hides the wait for the ghosts
          LAUNCH_PDL_AS("k_contagion(local+ghosts)", k_contagion, nb, 256, sim->ctl, sim->hot[o], sim->cells[t],
                        sim->inf_list, gr[0], gr[1], sim->history, rows, 16 | 3, D, (ExchangeHeader *)sim->blocks[sl.rank],
                        nullptr);
        } else if (sim->direct) {
          // the local sources need nothing from the neighbours: their launch hides the wait for the ghosts
          SlabDirect none;
          memset(&none, 0, sizeof none);
          LAUNCH_PDLN(k_contagion, nb, 256, sim->ctl, sim->hot[o], sim->cells[t], sim->inf_list, nullptr, nullptr,
                 sim->history, rows, 0, none, nullptr, nullptr);
          LAUNCH_PDL_AS("k_contagion(ghosts)", k_contagion, nbs, 256, sim->ctl, sim->hot[o], sim->cells[t], sim->inf_list,
                    gr[0], gr[1], sim->history, rows, 8 | 3, D, (ExchangeHeader *)sim->blocks[sl.rank], nullptr);
        } else {
          LAUNCH_PDLN(k_contagion, nb, 256, sim->ctl, sim->hot[o], sim->cells[t], sim->inf_list, gr[0], gr[1], sim->history,
                 rows, 2, D, nullptr, nullptr);
        }
      } else {
        LAUNCH_PDLN(k_slab_export, 1, 32, sim->ctl, rows, D);
      }
      break;
    case 3:
      if (sim->direct) {
        // the local half ran inside the contact kernel (or runs here for a static build); the global half is
        // enqueued behind the next build's pass A, or by whoever asks for statistics first
        sim->late_pending = true;
        sim->late_offset = kAdvance ? 0u : 1u;   // a step has closed locally (build_seq advanced) when its global half runs
        sim->late_record = kAdvance ? 1 : 0;
        if (!kAdvance) {   // a static build closes in order: the global y-range first, then the next grid
          const int rc = flush_late_close(sim);
          if (rc != CABS_OK) return rc;
          LAUNCH_PDLN(k_slab_close_local, 1, 32, sim->ctl, sim->cells[t], 0, ms[0], ms[1], gs[0], gs[1]);
        }
      } else {
        LAUNCH_PDLN(k_slab_close, 1, 32, sim->ctl, (const long long *)rows, sim->cells[t], sim->history,
               kAdvance ? 1 : 0, ms[0], ms[1], gs[0], gs[1], D);
      }
      sim->build_seq++;
      sim->cur = o;
      sim->tab = t ^ 1;
      if (kAdvance) sim->step++;
      sim->cells_dirty = false;
      break;
    default:
      return fail(sim, CABS_ERR_INVALID, "cabs_slab_phase: phase must be 0..3");
  }
  if (sim->profiling) prof_mark(sim, nullptr);  // the exchange that follows is not a kernel of this library
  CK(cudaGetLastError());
  return CABS_OK;
}

extern "C" int cabs_slab_phase(cabs_sim *sim, int phase, int advance) {
  if (!sim) return fail(sim, CABS_ERR_INVALID, "cabs_slab_phase: null handle");
  if (!sim->slab_on || sim->direct) return fail(sim, CABS_ERR_INVALID, "cabs_slab_phase: call cabs_set_slab first");
  if (!sim->have_params) return fail(sim, CABS_ERR_INVALID, "cabs_slab_phase: call cabs_set_params first");
  CK(cudaSetDevice(sim->device));
  return advance ? slab_phase<true>(sim, phase) : slab_phase<false>(sim, phase);
}


// ---- direct exchange over peer memory ---------------------------------------------------------------
extern "C" size_t cabs_slab_block_bytes(uint32_t n_ranks, uint32_t migrant_capacity, uint32_t ghost_capacity) {
  return block_layout(n_ranks, migrant_capacity, ghost_capacity).total;
}

static thread_local std::string g_util_error;
static int util_fail(const char *what, cudaError_t e) {
  g_util_error = std::string(what) + ": " + cudaGetErrorString(e);
  g_create_error = g_util_error;
  return CABS_ERR_CUDA;
}

extern "C" int cabs_device_alloc(int device, size_t bytes, void **ptr) {
  if (!ptr || bytes == 0) return CABS_ERR_INVALID;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return util_fail("cudaSetDevice", e);
  if ((e = cudaMalloc(ptr, bytes)) != cudaSuccess) return util_fail("cudaMalloc", e);
  if ((e = cudaMemset(*ptr, 0, bytes)) != cudaSuccess) return util_fail("cudaMemset", e);
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return util_fail("cudaDeviceSynchronize", e);
  return CABS_OK;
}

extern "C" int cabs_device_free(int device, void *ptr) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return util_fail("cudaSetDevice", e);
  if ((e = cudaFree(ptr)) != cudaSuccess) return util_fail("cudaFree", e);
  return CABS_OK;
}

extern "C" int cabs_ipc_export(int device, void *ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return util_fail("cudaSetDevice", e);
  cudaIpcMemHandle_t h;
  if ((e = cudaIpcGetMemHandle(&h, ptr)) != cudaSuccess) return util_fail("cudaIpcGetMemHandle", e);
  memcpy(handle, &h, 64);
  return CABS_OK;
}

extern "C" int cabs_ipc_open(int device, const unsigned char handle[64], void **ptr) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return util_fail("cudaSetDevice", e);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  if ((e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess)
    return util_fail("cudaIpcOpenMemHandle", e);
  return CABS_OK;
}

extern "C" int cabs_ipc_close(int device, void *ptr) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return util_fail("cudaSetDevice", e);
  if ((e = cudaIpcCloseMemHandle(ptr)) != cudaSuccess) return util_fail("cudaIpcCloseMemHandle", e);
  return CABS_OK;
}

extern "C" int cabs_set_slab_direct(cabs_sim *sim, const cabs_slab_direct *d) {
  if (!sim || !d) return fail(sim, CABS_ERR_INVALID, "cabs_set_slab_direct: null argument");
  if (d->n_ranks < 1 || d->n_ranks > CABS_MAX_RANKS || d->rank < 0 || d->rank >= d->n_ranks || d->x_lo >= d->x_hi)
    return fail(sim, CABS_ERR_INVALID, "cabs_set_slab_direct: bad rank / bounds (at most 16 ranks)");
  if (d->migrant_capacity == 0 || d->ghost_capacity == 0)
    return fail(sim, CABS_ERR_INVALID, "cabs_set_slab_direct: capacities are required");
  if (sim->f64) return fail(sim, CABS_ERR_INVALID, "cabs_set_slab_direct: the off-lattice mode runs on one device");
  for (int r = 0; r < d->n_ranks; ++r)
    if (!d->blocks[r]) return fail(sim, CABS_ERR_INVALID, "cabs_set_slab_direct: null exchange block");
  if (sim->have_params && sim->params.schedule == CABS_SEQUENTIAL)
    return fail(sim, CABS_ERR_INVALID, "the sequential schedule is defined on one device only");
  CK(cudaSetDevice(sim->device));
  CK(cudaStreamSynchronize(sim->stream));
  cudaFree(sim->emig_list); cudaFree(sim->imm_rank); cudaFree(sim->dev_blocks);
  sim->emig_list = sim->imm_rank = nullptr;
  sim->dev_blocks = nullptr;
  CK(cudaMalloc(&sim->emig_list, 2 * (size_t)d->migrant_capacity * sizeof(unsigned)));
  CK(cudaMalloc(&sim->imm_rank, 2 * (size_t)d->migrant_capacity * sizeof(unsigned)));
  CK(cudaMalloc(&sim->dev_blocks, CABS_MAX_RANKS * sizeof(unsigned char *)));
  for (int r = 0; r < CABS_MAX_RANKS; ++r) sim->blocks[r] = r < d->n_ranks ? (unsigned char *)d->blocks[r] : nullptr;
  CK(cudaMemcpyAsync(sim->dev_blocks, sim->blocks, CABS_MAX_RANKS * sizeof(unsigned char *), cudaMemcpyHostToDevice,
                     sim->stream));
  memset(&sim->slab, 0, sizeof sim->slab);
  sim->slab.rank = d->rank; sim->slab.n_ranks = d->n_ranks; sim->slab.x_lo = d->x_lo; sim->slab.x_hi = d->x_hi;
  sim->slab.migrant_capacity = d->migrant_capacity; sim->slab.ghost_capacity = d->ghost_capacity;
  sim->slab_on = true;
  sim->direct = true;
  sim->build_seq = 0;
  sim->sp.slab_lo = d->x_lo;
  sim->sp.slab_hi = d->x_hi;
  drop_graphs(sim);
  LAUNCH(k_set_slab, 1, 32, sim->ctl, d->rank, d->n_ranks, d->x_lo, d->x_hi, d->migrant_capacity, d->ghost_capacity);
  if (sim->n > 0) LAUNCH(k_replan, 1, 32, sim->ctl, 0);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(sim->stream));
  sim->cells_dirty = true;
  return CABS_OK;
}

extern "C" int cabs_slab_step(cabs_sim *sim, int advance) {
  NvtxRange nvtx_("cabs_slab_step");
  if (!sim) return fail(sim, CABS_ERR_INVALID, "cabs_slab_step: null handle");
  if (!sim->slab_on || !sim->direct) return fail(sim, CABS_ERR_INVALID, "cabs_slab_step: call cabs_set_slab_direct first");
  if (!sim->have_params) return fail(sim, CABS_ERR_INVALID, "cabs_slab_step: call cabs_set_params first");
  CK(cudaSetDevice(sim->device));
  for (int phase = 0; phase < 4; ++phase) {
    const int rc = advance ? slab_phase<true>(sim, phase) : slab_phase<false>(sim, phase);
    if (rc != CABS_OK) return rc;
  }
  return CABS_OK;
}

// `nsteps` steps of the direct mode.  Pairs of steps replay a captured CUDA graph: every launch argument is the same
// for any two builds that start from the same ping/pong phase and flag parity (the sequence number itself lives in Ctl).
extern "C" int cabs_slab_run(cabs_sim *sim, int nsteps) {
  NvtxRange nvtx_("cabs_slab_run");
  if (!sim || nsteps < 0) return fail(sim, CABS_ERR_INVALID, "cabs_slab_run: bad argument");
  if (!sim->slab_on || !sim->direct) return fail(sim, CABS_ERR_INVALID, "cabs_slab_run: call cabs_set_slab_direct first");
  if (!sim->have_params) return fail(sim, CABS_ERR_INVALID, "cabs_slab_run: call cabs_set_params first");
  CK(cudaSetDevice(sim->device));
  int s = 0;
  // steady state only: the previous build was a step whose global closing is still pending
  if (!sim->profiling && nsteps >= 5) {
    for (; s < 1; ++s) {
      const int rc = cabs_slab_step(sim, 1);
      if (rc != CABS_OK) return rc;
    }
    const unsigned parity = (sim->build_seq + 1) & 1u;
    if (sim->slab_graph && (sim->slab_graph_cur != sim->cur || sim->slab_graph_tab != sim->tab ||
                            sim->slab_graph_parity != parity)) {
      cudaGraphExecDestroy(sim->slab_graph);
      sim->slab_graph = nullptr;
    }
    if (!sim->slab_graph && sim->late_pending && sim->late_record == 1 && sim->late_offset == 0) {
      cudaGraph_t graph = nullptr;
      const int cur0 = sim->cur, tab0 = sim->tab;
      const uint64_t launches0 = sim->launches, step0 = sim->step;
      const unsigned seq0 = sim->build_seq;
      CK(cudaStreamBeginCapture(sim->stream, cudaStreamCaptureModeThreadLocal));
      int rc = cabs_slab_step(sim, 1);
      if (rc == CABS_OK) rc = cabs_slab_step(sim, 1);
      cudaError_t ce = cudaStreamEndCapture(sim->stream, &graph);
      sim->slab_graph_launches = sim->launches - launches0;
      sim->cur = cur0; sim->tab = tab0; sim->launches = launches0; sim->step = step0; sim->build_seq = seq0;
      sim->late_pending = true; sim->late_offset = 0; sim->late_record = 1;
      if (rc != CABS_OK) return rc;
      if (ce != cudaSuccess) { sim->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce); return CABS_ERR_CUDA; }
      CK(cudaGraphInstantiate(&sim->slab_graph, graph, 0));
      CK(cudaGraphDestroy(graph));
      sim->slab_graph_cur = cur0; sim->slab_graph_tab = tab0; sim->slab_graph_parity = parity;
    }
    if (sim->slab_graph) {
      for (; s + 2 <= nsteps; s += 2) {
        CK(cudaGraphLaunch(sim->slab_graph, sim->stream));
        sim->launches += sim->slab_graph_launches;
        sim->step += 2;
        sim->build_seq += 2;
      }
    }
  }
  for (; s < nsteps; ++s) {
    const int rc = cabs_slab_step(sim, 1);
    if (rc != CABS_OK) return rc;
  }
  return CABS_OK;
}

// number of resident agents (changes on the device when agents migrate between slabs)
static int refresh_n(cabs_sim *sim) {
  if (!sim->slab_on) return CABS_OK;
  unsigned n = 0;
  CK(cudaMemcpyAsync(&n, &sim->ctl->n, sizeof n, cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  sim->n = n;
  return CABS_OK;
}

static int upload_impl(cabs_sim *sim, const cabs_soa_host *pop, const double *xf, const double *yf,
                       bool on_device = false);

extern "C" int cabs_upload(cabs_sim *sim, const cabs_soa_host *pop) {
  NvtxRange nvtx_("cabs_upload");
  return upload_impl(sim, pop, nullptr, nullptr);
}

static bool device_pointer(const cabs_sim *sim, const void *p) {
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return (a.type == cudaMemoryTypeDevice && a.device == sim->device) || a.type == cudaMemoryTypeManaged;
}

extern "C" int cabs_bind_device(cabs_sim *sim, const cabs_soa_host *pop) {
  NvtxRange nvtx_("cabs_bind_device");
  if (!sim || !pop) return fail(sim, CABS_ERR_INVALID, "cabs_bind_device: null argument");
  for (const void *p : {(const void *)pop->x, (const void *)pop->y, (const void *)pop->status, (const void *)pop->severity,
                        (const void *)pop->infected_time, (const void *)pop->age, (const void *)pop->stratum,
                        (const void *)pop->wealth, (const void *)pop->id})
    if (!device_pointer(sim, p)) return fail(sim, CABS_ERR_INVALID, "cabs_bind_device: every column must be device memory of the handle's device");
  return upload_impl(sim, pop, nullptr, nullptr, true);
}

extern "C" int cabs_upload_f64(cabs_sim *sim, const cabs_soa_host *pop, const double *x, const double *y) {
  NvtxRange nvtx_("cabs_upload_f64");
  if (sim && !sim->f64) return fail(sim, CABS_ERR_INVALID, "cabs_upload_f64: the handle was not created with CABS_FLAG_F64_POSITIONS");
  if (sim && pop && pop->n && (!x || !y)) return fail(sim, CABS_ERR_INVALID, "cabs_upload_f64: null position column");
  return upload_impl(sim, pop, x, y);
}

static int upload_impl(cabs_sim *sim, const cabs_soa_host *pop, const double *xf, const double *yf, bool on_device) {
  if (!sim || !pop) return fail(sim, CABS_ERR_INVALID, "cabs_upload: null argument");
  if (!sim->have_params) return fail(sim, CABS_ERR_INVALID, "cabs_upload: call cabs_set_params first");
  if (pop->n > sim->capacity) return fail(sim, CABS_ERR_CAPACITY, "cabs_upload: population larger than capacity");
  const size_t n = pop->n;
  if (n && ((!xf && (!pop->x || !pop->y)) || !pop->status || !pop->severity || !pop->infected_time || !pop->age ||
            !pop->stratum || !pop->wealth))
    return fail(sim, CABS_ERR_INVALID, "cabs_upload: null column");
  CK(cudaSetDevice(sim->device));
  const int c = sim->cur;
  const size_t cap = sim->capacity;
  int *sx = reinterpret_cast<int *>(sim->hot[c ^ 1]), *sy = sx + cap;  // the other copy is free: scratch
  if (n && on_device) {
    // the columns are device memory already (e.g. torch tensors): pack the records straight from them, no staging copy
    if (!pop->x || !pop->y) return fail(sim, CABS_ERR_INVALID, "cabs_bind_device: null column");
    CK(cudaMemcpyAsync(sim->w[c], pop->wealth, n * sizeof(double), cudaMemcpyDeviceToDevice, sim->stream));
    LAUNCH(k_pack, agent_blocks(sim), 256, (unsigned)n, (const int *)pop->x, (const int *)pop->y,
           (const uint8_t *)pop->status, (const uint8_t *)pop->severity, (const uint8_t *)pop->infected_time,
           (const uint8_t *)pop->age, (const uint8_t *)pop->stratum, (const uint32_t *)pop->id, sim->hot[c]);
  } else if (n) {
    if (pop->x && pop->y) {
      CK(cudaMemcpyAsync(sx, pop->x, n * sizeof(int), cudaMemcpyHostToDevice, sim->stream));
      CK(cudaMemcpyAsync(sy, pop->y, n * sizeof(int), cudaMemcpyHostToDevice, sim->stream));
    } else {   // binary64 positions follow (cabs_upload_f64): the lattice columns are derived from them on the device
      CK(cudaMemsetAsync(sx, 0, n * sizeof(int), sim->stream));
      CK(cudaMemsetAsync(sy, 0, n * sizeof(int), sim->stream));
    }
    CK(cudaMemcpyAsync(sim->w[c], pop->wealth, n * sizeof(double), cudaMemcpyHostToDevice, sim->stream));
    CK(cudaMemcpyAsync(sim->stage + 0 * cap, pop->status, n, cudaMemcpyHostToDevice, sim->stream));
    CK(cudaMemcpyAsync(sim->stage + 1 * cap, pop->severity, n, cudaMemcpyHostToDevice, sim->stream));
    CK(cudaMemcpyAsync(sim->stage + 2 * cap, pop->infected_time, n, cudaMemcpyHostToDevice, sim->stream));
    CK(cudaMemcpyAsync(sim->stage + 3 * cap, pop->age, n, cudaMemcpyHostToDevice, sim->stream));
    CK(cudaMemcpyAsync(sim->stage + 4 * cap, pop->stratum, n, cudaMemcpyHostToDevice, sim->stream));
    if (pop->id) CK(cudaMemcpyAsync(sim->stage_id, pop->id, n * sizeof(uint32_t), cudaMemcpyHostToDevice, sim->stream));
    LAUNCH(k_pack, agent_blocks(sim), 256, (unsigned)n, sx, sy, sim->stage + 0 * cap, sim->stage + 1 * cap,
           sim->stage + 2 * cap, sim->stage + 3 * cap, sim->stage + 4 * cap, pop->id ? sim->stage_id : nullptr,
           sim->hot[c]);
  }
  if (n && sim->f64) {   // exact positions: the caller's binary64 columns, or the lattice points themselves
    double *dxf = nullptr, *dyf = nullptr;
    if (xf) {            // staged through the moved-position array (free between builds)
      dxf = reinterpret_cast<double *>(sim->npos);
      dyf = dxf + cap;
      CK(cudaMemcpyAsync(dxf, xf, n * sizeof(double), cudaMemcpyHostToDevice, sim->stream));
      CK(cudaMemcpyAsync(dyf, yf, n * sizeof(double), cudaMemcpyHostToDevice, sim->stream));
    }
    LAUNCH(k_pack_pos, agent_blocks(sim), 256, (unsigned)n, dxf, dyf, sim->hot[c], sim->pos[c], sim->ctl);
  }
  // The step counter is part of the draw key (seed, agent, step) and is NOT reset here: a caller that downloads,
  // edits and re-uploads the population every iteration (set_population, abs.py:65-69) must not see the draws of
  // step 0 again.  A fresh handle starts at step 0; cabs_set_step restores the clock of a checkpoint.
  if (n && pop->id && !sim->slab_on) {   // non-slab handles scatter by id on download: ids must be a permutation of 0..n-1
    unsigned *mark = reinterpret_cast<unsigned *>(sim->aux);   // free until the next build
    CK(cudaMemsetAsync(mark, 0, n * sizeof(unsigned), sim->stream));
    LAUNCH(k_check_ids, agent_blocks(sim), 256, sim->ctl, sim->hot[c], (unsigned)n, mark);
  }
  const unsigned n32 = (unsigned)n;
  CK(cudaMemcpyAsync(&sim->ctl->n, &n32, sizeof(unsigned), cudaMemcpyHostToDevice, sim->stream));
  sim->n = n;
  if (sim->slab_on) {  // the caller drives the four phases of the first build (advance = 0)
    sim->cells_dirty = true;
    CK(cudaGetLastError());
    return CABS_OK;
  }
  if (sim->f64 || getenv("CABS_EAGER_SORT")) return rebuild_static(sim);
  // No sort here: the first step's pass B scatters into cell order from whatever order it reads.  One pass for the
  // statistics and the bounding box, then the closing that plans the first grid (k_upload_stats / k_close_upload).
  // (the accumulators are clean here: every closing, k_replan and k_init_ctl leave them reset)
  LAUNCH(k_upload_stats, agent_blocks(sim), 256, sim->ctl, sim->hot[c], sim->w[c]);
  LAUNCH(k_close_upload, 1, 32, sim->ctl);
  CK(cudaGetLastError());
  sim->cells_dirty = true;
  sim->unsorted = false;
  return CABS_OK;
}

extern "C" int cabs_set_step(cabs_sim *sim, uint64_t step) {
  if (!sim) return fail(sim, CABS_ERR_INVALID, "cabs_set_step: null handle");
  if (step >= (1ull << 32)) return fail(sim, CABS_ERR_INVALID, "cabs_set_step: the step counter is 32 bits wide");
  if (sim->slab_on) return fail(sim, CABS_ERR_INVALID, "cabs_set_step: not defined for slab handles");
  CK(cudaSetDevice(sim->device));
  const unsigned s32 = (unsigned)step;
  CK(cudaMemcpyAsync(&sim->ctl->step, &s32, sizeof(unsigned), cudaMemcpyHostToDevice, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  sim->step = step;
  sim->hist_first = step;   // statistics rows of earlier steps belong to another clock
  return CABS_OK;
}

extern "C" int cabs_download(cabs_sim *sim, cabs_soa_host *pop) {
  NvtxRange nvtx_("cabs_download");
  if (!sim || !pop) return fail(sim, CABS_ERR_INVALID, "cabs_download: null argument");
  CK(cudaSetDevice(sim->device));
  {
    const int rc = refresh_n(sim);
    if (rc != CABS_OK) return rc;
  }
  if (pop->n < sim->n) return fail(sim, CABS_ERR_CAPACITY, "cabs_download: host buffers smaller than the population");
  const size_t n = sim->n, cap = sim->capacity;
  pop->n = n;
  if (n == 0) return CABS_OK;
  const int c = sim->cur, o = c ^ 1;  // the other copy is free between steps: use it as scratch
  int *sx = reinterpret_cast<int *>(sim->hot[o]), *sy = sx + cap;
  LAUNCH(k_unpack, agent_blocks(sim), 256, (unsigned)n, sim->slab_on ? 1 : 0, sim->hot[c], sim->w[c], sx, sy, sim->stage + 0 * cap,
         sim->stage + 1 * cap, sim->stage + 2 * cap, sim->stage + 3 * cap, sim->stage + 4 * cap, sim->w[o],
         sim->stage_id);
  CK(cudaGetLastError());
  if (pop->x) CK(cudaMemcpyAsync(pop->x, sx, n * sizeof(int), cudaMemcpyDeviceToHost, sim->stream));
  if (pop->y) CK(cudaMemcpyAsync(pop->y, sy, n * sizeof(int), cudaMemcpyDeviceToHost, sim->stream));
  if (pop->wealth) CK(cudaMemcpyAsync(pop->wealth, sim->w[o], n * sizeof(double), cudaMemcpyDeviceToHost, sim->stream));
  if (pop->status) CK(cudaMemcpyAsync(pop->status, sim->stage + 0 * cap, n, cudaMemcpyDeviceToHost, sim->stream));
  if (pop->severity) CK(cudaMemcpyAsync(pop->severity, sim->stage + 1 * cap, n, cudaMemcpyDeviceToHost, sim->stream));
  if (pop->infected_time)
    CK(cudaMemcpyAsync(pop->infected_time, sim->stage + 2 * cap, n, cudaMemcpyDeviceToHost, sim->stream));
  if (pop->age) CK(cudaMemcpyAsync(pop->age, sim->stage + 3 * cap, n, cudaMemcpyDeviceToHost, sim->stream));
  if (pop->stratum) CK(cudaMemcpyAsync(pop->stratum, sim->stage + 4 * cap, n, cudaMemcpyDeviceToHost, sim->stream));
  if (pop->id) CK(cudaMemcpyAsync(pop->id, sim->stage_id, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return CABS_OK;
}

extern "C" int cabs_snapshot(cabs_sim *sim, uint32_t stride, int32_t *x, int32_t *y, uint8_t *status, uint8_t *severity,
                             uint64_t capacity, uint64_t *count) {
  NvtxRange nvtx_("cabs_snapshot");
  if (!sim || !count || stride == 0) return fail(sim, CABS_ERR_INVALID, "cabs_snapshot: bad argument");
  if (sim->slab_on) return fail(sim, CABS_ERR_INVALID, "cabs_snapshot: not defined for slab handles (ids are global there)");
  CK(cudaSetDevice(sim->device));
  const size_t n = sim->n, cap = sim->capacity, m = (n + stride - 1) / stride;
  *count = m;
  if (m == 0) return CABS_OK;
  if (capacity < m) return fail(sim, CABS_ERR_CAPACITY, "cabs_snapshot: host buffers smaller than the sample");
  int *sx = reinterpret_cast<int *>(sim->hot[sim->cur ^ 1]), *sy = sx + cap;   // the other copy is free between steps
  LAUNCH(k_snapshot, agent_blocks(sim), 256, (unsigned)n, stride, sim->hot[sim->cur], sx, sy, sim->stage, sim->stage + cap);
  CK(cudaGetLastError());
  if (x) CK(cudaMemcpyAsync(x, sx, m * sizeof(int), cudaMemcpyDeviceToHost, sim->stream));
  if (y) CK(cudaMemcpyAsync(y, sy, m * sizeof(int), cudaMemcpyDeviceToHost, sim->stream));
  if (status) CK(cudaMemcpyAsync(status, sim->stage, m, cudaMemcpyDeviceToHost, sim->stream));
  if (severity) CK(cudaMemcpyAsync(severity, sim->stage + cap, m, cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return CABS_OK;
}

extern "C" int cabs_download_positions_f64(cabs_sim *sim, double *x, double *y, uint64_t capacity) {
  NvtxRange nvtx_("cabs_download_positions_f64");
  if (!sim || !x || !y) return fail(sim, CABS_ERR_INVALID, "cabs_download_positions_f64: null argument");
  if (!sim->f64) return fail(sim, CABS_ERR_INVALID, "cabs_download_positions_f64: the handle holds lattice positions");
  if (capacity < sim->n) return fail(sim, CABS_ERR_CAPACITY, "cabs_download_positions_f64: host buffers smaller than the population");
  CK(cudaSetDevice(sim->device));
  const size_t n = sim->n, cap = sim->capacity;
  if (n == 0) return CABS_OK;
  double *dx = reinterpret_cast<double *>(sim->npos), *dy = dx + cap;   // free between builds
  LAUNCH(k_unpack_pos, agent_blocks(sim), 256, (unsigned)n, sim->hot[sim->cur], sim->pos[sim->cur], dx, dy);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(x, dx, n * sizeof(double), cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaMemcpyAsync(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return CABS_OK;
}

extern "C" int cabs_set_rules(cabs_sim *sim, const cabs_rule *rules, int n_move, int n_attr) {
  if (!sim || ((n_move + n_attr) > 0 && !rules)) return fail(sim, CABS_ERR_INVALID, "cabs_set_rules: null argument");
  if (!sim->f64) return fail(sim, CABS_ERR_INVALID, "cabs_set_rules: per-agent rules need a handle created with CABS_FLAG_F64_POSITIONS");
  if (n_move < 0 || n_attr < 0 || n_move + n_attr > kMaxRules) return fail(sim, CABS_ERR_INVALID, "cabs_set_rules: at most 8 rules");
  static_assert(sizeof(cabs_rule) == sizeof(DevRule), "cabs_rule and DevRule share one layout");
  int margin = 0, box[4] = {1, 0, 1, 0};
  for (int k = 0; k < n_move + n_attr; ++k) {
    const cabs_rule &r = rules[k];
    const bool move = k < n_move;
    if (move && r.kind != CABS_RULE_MOVE_STAY && r.kind != CABS_RULE_MOVE_POINT)
      return fail(sim, CABS_ERR_INVALID, "cabs_set_rules: the first n_move rules must be move rules");
    if (!move && (r.kind != CABS_RULE_ATTRIBUTE || r.attribute < 0 || r.attribute > CABS_RATTR_WEALTH))
      return fail(sim, CABS_ERR_INVALID, "cabs_set_rules: bad attribute rule");
    if (move) {
      if (!(r.sigma >= 0.0 && r.sigma < 1e6)) return fail(sim, CABS_ERR_INVALID, "cabs_set_rules: sigma out of range");
      const int m = (int)ceil(5.78 * r.sigma) + 2;   // the largest |normal| the draw map returns is 5.78
      if (r.kind == CABS_RULE_MOVE_STAY) {
        margin = margin > m ? margin : m;
      } else {
        if (!(fabs(r.target_x) < 2e8 && fabs(r.target_y) < 2e8)) return fail(sim, CABS_ERR_DOMAIN, "cabs_set_rules: target outside the supported box");
        const int x0 = (int)floor(r.target_x) - m, x1 = (int)floor(r.target_x) + m;
        const int y0 = (int)floor(r.target_y) - m, y1 = (int)floor(r.target_y) + m;
        if (box[0] > box[1]) { box[0] = x0; box[1] = x1; box[2] = y0; box[3] = y1; }
        else { box[0] = box[0] < x0 ? box[0] : x0; box[1] = box[1] > x1 ? box[1] : x1;
               box[2] = box[2] < y0 ? box[2] : y0; box[3] = box[3] > y1 ? box[3] : y1; }
      }
    }
  }
  CK(cudaSetDevice(sim->device));
  if (n_move + n_attr > 0)
    CK(cudaMemcpyAsync(sim->rules, rules, (size_t)(n_move + n_attr) * sizeof(DevRule), cudaMemcpyHostToDevice, sim->stream));
  LAUNCH(k_set_rules, 1, 32, sim->ctl, 1, margin, box[0], box[1], box[2], box[3]);
  if (sim->n > 0) LAUNCH(k_replan, 1, 32, sim->ctl, 0);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(sim->stream));   // `rules` is the caller's memory
  sim->n_move_rules = n_move;
  sim->n_attr_rules = n_attr;
  drop_graphs(sim);     // captured launches hold the old rule counts
  return CABS_OK;
}

extern "C" int cabs_step(cabs_sim *sim, int nsteps) {
  NvtxRange nvtx_("cabs_step");
  if (!sim || nsteps < 0) return fail(sim, CABS_ERR_INVALID, "cabs_step: bad argument");
  if (!sim->have_params) return fail(sim, CABS_ERR_INVALID, "cabs_step: call cabs_set_params first");
  if (sim->slab_on) return fail(sim, CABS_ERR_INVALID, "cabs_step: a slab handle is stepped with cabs_slab_phase");
  CK(cudaSetDevice(sim->device));
  if (sim->n > 0 && sim->n <= (uint64_t)kSmallMax && sim->params.schedule == CABS_SYNCHRONOUS && nsteps > 0 && !sim->f64) {
    // the whole run in one launch of one resident block (small populations are launch-latency bound otherwise)
    LAUNCH(k_small_run, 1, kSmallMax, sim->ctl, sim->hot[sim->cur], sim->w[sim->cur], sim->history, sim->sqrt_tab,
           nsteps);
    sim->step += nsteps;
    sim->cells_dirty = true;   // the records are no longer in cell order; the pair hook rebuilds on demand
    sim->unsorted = true;
    if (sim->profiling) prof_mark(sim, nullptr);
    CK(cudaGetLastError());
    return CABS_OK;
  }
  if (sim->unsorted && sim->n > 0) {   // steps of the tiled path start from a cell-sorted population
    const int rc = rebuild_static(sim);
    if (rc != CABS_OK) return rc;
  }
  int s = 0;
  const bool sequential = sim->params.schedule == CABS_SEQUENTIAL;
  const bool capturable = sim->stream != nullptr && sim->stream != cudaStreamLegacy && sim->stream != cudaStreamPerThread;
  if (!sim->profiling && (!sequential || sim->seq_grid > 0) && nsteps >= 4 && capturable && !sim->graph_broken) {
    // pairs of steps through the captured graph (the kernels take every changing quantity from Ctl, so the same
    // eight launches are valid for any pair of steps that starts from the same ping/pong phase)
    if (sim->step_graph && (sim->graph_cur != sim->cur || sim->graph_tab != sim->tab)) {
      cudaGraphExecDestroy(sim->step_graph);
      sim->step_graph = nullptr;
    }
    if (!sim->step_graph) {
      cudaGraph_t graph = nullptr;
      const int cur0 = sim->cur, tab0 = sim->tab;
      const uint64_t launches0 = sim->launches;
      cudaError_t ce = cudaStreamBeginCapture(sim->stream, cudaStreamCaptureModeThreadLocal);
      if (ce == cudaSuccess) {
        enqueue_build<true>(sim);
        enqueue_build<true>(sim);
        ce = cudaStreamEndCapture(sim->stream, &graph);
        sim->cur = cur0;              // capturing enqueued nothing: restore the host-side phase
        sim->tab = tab0;
        sim->launches = launches0;
        if (ce == cudaSuccess) ce = cudaGraphInstantiate(&sim->step_graph, graph, 0);
        if (graph) cudaGraphDestroy(graph);
      }
      if (ce != cudaSuccess) {        // a stream that cannot be captured: plain launches from now on
        if (getenv("CABS_DEBUG")) fprintf(stderr, "cabs_step: graph capture failed (%s), falling back to plain launches\n", cudaGetErrorString(ce));
        cudaGetLastError();
        sim->step_graph = nullptr;
        sim->graph_broken = true;
      } else {
        sim->graph_cur = cur0;
        sim->graph_tab = tab0;
      }
    }
    for (; sim->step_graph && s + 2 <= nsteps; s += 2) {
      CK(cudaGraphLaunch(sim->step_graph, sim->stream));
      sim->launches += 8;
      sim->step += 2;
    }
  }
  for (; s < nsteps; ++s) {
    enqueue_build<true>(sim);
    sim->step++;
  }
  if (sim->profiling) prof_mark(sim, nullptr);  // closing timestamp of the batch
  CK(cudaGetLastError());
  sim->cells_dirty = false;
  return CABS_OK;
}

static int report_device_error(cabs_sim *sim, unsigned err);

// The whole control block into the pinned mirror: one copy, one synchronisation; then the sticky error bits.
static int fetch_ctl(cabs_sim *sim) {
  {
    const int rc = flush_late_close(sim);
    if (rc != CABS_OK) return rc;
  }
  CK(cudaMemcpyAsync(sim->host_ctl, sim->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return report_device_error(sim, sim->host_ctl->err);
}

static int check_device_error(cabs_sim *sim) {
  {
    const int rc = flush_late_close(sim);
    if (rc != CABS_OK) return rc;
  }
  unsigned err = 0;
  CK(cudaMemcpyAsync(&err, &sim->ctl->err, sizeof(unsigned), cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  return report_device_error(sim, err);
}

static int report_device_error(cabs_sim *sim, unsigned err) {
  if (err & kErrBadIds) {
    const unsigned clear = ~kErrBadIds;   // the population was rejected, the handle stays usable for a correct upload
    LAUNCH(k_clear_error, 1, 32, sim->ctl, clear);
    return fail(sim, CABS_ERR_INVALID, "cabs_upload: ids must be a permutation of 0..n-1 (they key the draws and the download order)");
  }
  if (err & kErrDomain) return fail(sim, CABS_ERR_DOMAIN, "population bounding box exceeds the supported 2^28 lattice units");
  if (err & kErrOutOfGrid) return fail(sim, CABS_ERR_DOMAIN, "internal: an agent left the planned cell grid");
  if (err & kErrPeerTimeout)
    return fail(sim, CABS_ERR_CUDA, "slab exchange: a neighbour's message did not arrive within 30 s (peer lost or out of step)");
  if (err & kErrExchange)
    return fail(sim, CABS_ERR_CAPACITY, "slab exchange overflow: more migrants / ghosts / residents than the buffers hold");
  return CABS_OK;
}

extern "C" int cabs_sync(cabs_sim *sim) {
  if (!sim) return fail(sim, CABS_ERR_INVALID, "cabs_sync: null handle");
  CK(cudaSetDevice(sim->device));
  CK(cudaStreamSynchronize(sim->stream));
  return check_device_error(sim);
}

static void convert_record(const cabs_sim *sim, const StatRecord &r, double *values, int64_t *counts) {
  const double n = (double)sim->params.population_size;
  for (int k = 0; k < 7; ++k) {
    values[k] = (double)r.cnt[k] / n;  // abs.py:301-307: count / population_size
    if (counts) counts[k] = (int64_t)r.cnt[k];
  }
  const double scale = (double)(1ull << sim->params.wealth_scale_bits);
  for (int k = 0; k < 5; ++k) values[7 + k] = (double)r.q[k] / scale;  // abs.py:309-312
}

extern "C" int cabs_stats_ex(cabs_sim *sim, double *values, int64_t *counts, cabs_params *params_now,
                             uint64_t *trigger_epoch) {
  NvtxRange nvtx_("cabs_stats");
  if (!sim || !values) return fail(sim, CABS_ERR_INVALID, "cabs_stats: null argument");
  if (!sim->have_params) return fail(sim, CABS_ERR_INVALID, "cabs_stats: call cabs_set_params first");
  CK(cudaSetDevice(sim->device));
  const int rc = fetch_ctl(sim);   // synchronises
  if (rc != CABS_OK) return rc;
  const Ctl &h = *sim->host_ctl;
  convert_record(sim, h.cur, values, counts);
  if (params_now) {
    *params_now = sim->params;
    for (int k = 0; k < 4; ++k) params_now->amplitudes[k] = (double)h.amp[k];
    params_now->contagion_rate = h.rate;
    params_now->contagion_distance = h.contagion_distance;
    params_now->critical_limit = h.critical_limit;
    sim->params = *params_now;
  }
  if (trigger_epoch) *trigger_epoch = h.trig_epoch;
  return CABS_OK;
}

extern "C" int cabs_stats(cabs_sim *sim, double *values, int64_t *counts) {
  return cabs_stats_ex(sim, values, counts, nullptr, nullptr);
}

extern "C" int cabs_stats_history(cabs_sim *sim, uint64_t first_step, uint64_t count, double *values) {
  NvtxRange nvtx_("cabs_stats_history");
  if (!sim || (count && !values)) return fail(sim, CABS_ERR_INVALID, "cabs_stats_history: null argument");
  if (first_step + count > sim->step) return fail(sim, CABS_ERR_INVALID, "cabs_stats_history: steps not executed yet");
  if (count && first_step < sim->hist_first)
    return fail(sim, CABS_ERR_INVALID, "cabs_stats_history: steps before the last cabs_set_step are not on record");
  if (sim->step - first_step > sim->hist_cap)
    return fail(sim, CABS_ERR_CAPACITY, "cabs_stats_history: records already overwritten (history_capacity)");
  CK(cudaSetDevice(sim->device));
  {
    const int rc = flush_late_close(sim);
    if (rc != CABS_OK) return rc;
  }
  if (count == 0) return CABS_OK;
  std::vector<StatRecord> recs(count);
  uint64_t done = 0;
  while (done < count) {  // the ring may wrap
    const uint64_t slot = (first_step + done) % sim->hist_cap;
    const uint64_t run = std::min<uint64_t>(count - done, sim->hist_cap - slot);
    CK(cudaMemcpyAsync(recs.data() + done, sim->history + slot, run * sizeof(StatRecord), cudaMemcpyDeviceToHost,
                       sim->stream));
    done += run;
  }
  const int rc = check_device_error(sim);
  if (rc != CABS_OK) return rc;
  for (uint64_t k = 0; k < count; ++k) convert_record(sim, recs[k], values + k * CABS_NUM_STATS, nullptr);
  return CABS_OK;
}

// MultiPopulationSimulation.execute, the loop over one pair of populations (abs.py:352-366).
extern "C" int cabs_cross_contact(cabs_sim *a, cabs_sim *b, int32_t ax, int32_t ay, int32_t bx, int32_t by,
                                  double contagion_distance, uint64_t seed, uint32_t pair_code) {
  NvtxRange nvtx_("cabs_cross_contact");
  if (!a || !b || a == b) return fail(a, CABS_ERR_INVALID, "cabs_cross_contact: two different handles are needed");
  if (!a->have_params || !b->have_params) return fail(a, CABS_ERR_INVALID, "cabs_cross_contact: call cabs_set_params first");
  if (a->slab_on || b->slab_on) return fail(a, CABS_ERR_INVALID, "cabs_cross_contact: not defined for slab handles");
  if (a->f64 || b->f64) return fail(a, CABS_ERR_INVALID, "cabs_cross_contact: populations are placed on the integer lattice");
  if (a->device != b->device) return fail(a, CABS_ERR_INVALID, "cabs_cross_contact: both populations must live on one device");
  if (a->step != b->step) return fail(a, CABS_ERR_INVALID, "cabs_cross_contact: the populations are at different steps");
  if (a->n == 0 || b->n == 0) return CABS_OK;
  if (cudaSetDevice(a->device) != cudaSuccess) return fail(a, CABS_ERR_CUDA, "cabs_cross_contact: cudaSetDevice failed");
  for (cabs_sim *s : {a, b}) {
    if (!s->cross_mark) {
      if (cudaMalloc(&s->cross_mark, s->capacity + kPad) != cudaSuccess)
        return fail(a, CABS_ERR_CUDA, "cabs_cross_contact: out of device memory");
      if (cudaMemsetAsync(s->cross_mark, 0, s->capacity + kPad, s->stream) != cudaSuccess)
        return fail(a, CABS_ERR_CUDA, "cabs_cross_contact: memset failed");
    }
  }
  // everything runs on A's stream, after what B's stream holds; B's stream continues after it
  cudaEvent_t before = nullptr, after = nullptr;
  if (cudaEventCreateWithFlags(&before, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&after, cudaEventDisableTiming) != cudaSuccess)
    return fail(a, CABS_ERR_CUDA, "cabs_cross_contact: event creation failed");
  cudaEventRecord(before, b->stream);
  cudaStreamWaitEvent(a->stream, before, 0);
  {
    cabs_sim *sim = a;   // the launch macro counts on `sim`
    const unsigned na = (unsigned)a->n;
    LAUNCH(k_cross_contact, (na + 255u) / 256u, 256, a->ctl, a->hot[a->cur], b->ctl, b->hot[b->cur], ax - bx, ay - by,
           contagion_distance, (uint32_t)a->step, pair_code, (uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32),
           a->cross_mark, b->cross_mark);
    LAUNCH(k_cross_apply, a->num_sms * 4, 256, a->ctl, a->hot[a->cur], a->cross_mark);
    LAUNCH(k_cross_apply, a->num_sms * 4, 256, b->ctl, b->hot[b->cur], b->cross_mark);
  }
  cudaEventRecord(after, a->stream);
  cudaStreamWaitEvent(b->stream, after, 0);
  cudaEventDestroy(before);
  cudaEventDestroy(after);
  {
    cabs_sim *sim = a;
    CK(cudaGetLastError());
  }
  // the statistics record of each handle follows its population (MultiPopulationSimulation.get_statistics reads the
  // populations after the cross contacts, abs.py:389-410); the history rows of the steps stay as they were
  int rc = rebuild_static(a);
  if (rc == CABS_OK) rc = rebuild_static(b);
  return rc;
}

extern "C" int cabs_contact_pairs(cabs_sim *sim, int64_t *ij, size_t capacity_pairs, size_t *n_pairs) {
  NvtxRange nvtx_("cabs_contact_pairs");
  if (!sim || !n_pairs || (capacity_pairs && !ij)) return fail(sim, CABS_ERR_INVALID, "cabs_contact_pairs: null argument");
  CK(cudaSetDevice(sim->device));
  if (sim->cells_dirty) {
    if (sim->slab_on) return fail(sim, CABS_ERR_INVALID, "cabs_contact_pairs: run a slab build first");
    const int rc = rebuild_static(sim);
    if (rc != CABS_OK) return rc;
  }
  if (capacity_pairs > sim->pairs_cap) {
    cudaFree(sim->pairs);
    sim->pairs = nullptr;
    sim->pairs_cap = 0;
    CK(cudaMalloc(&sim->pairs, capacity_pairs * 2 * sizeof(long long)));
    sim->pairs_cap = capacity_pairs;
  }
  CK(cudaMemsetAsync(sim->pair_count, 0, sizeof(unsigned long long), sim->stream));
  const int c = sim->cur;
  LAUNCH(k_contact_pairs, agent_blocks(sim), 256, sim->ctl, sim->hot[c], sim->cells[sim->tab ^ 1], sim->pairs,
         (unsigned long long)capacity_pairs, sim->pair_count, sim->f64 ? sim->pos[c] : nullptr);
  CK(cudaGetLastError());
  unsigned long long total = 0;
  CK(cudaMemcpyAsync(&total, sim->pair_count, sizeof total, cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  const size_t ncopy = (size_t)std::min<unsigned long long>(total, capacity_pairs);
  if (ncopy) CK(cudaMemcpy(ij, sim->pairs, ncopy * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
  *n_pairs = (size_t)total;
  return check_device_error(sim);
}

extern "C" int cabs_get_info(cabs_sim *sim, cabs_info *info) {
  if (!sim || !info) return fail(sim, CABS_ERR_INVALID, "cabs_get_info: null argument");
  CK(cudaSetDevice(sim->device));
  Ctl h;
  CK(cudaMemcpyAsync(&h, sim->ctl, sizeof h, cudaMemcpyDeviceToHost, sim->stream));
  CK(cudaStreamSynchronize(sim->stream));
  if (sim->slab_on) sim->n = h.n;
  info->n = sim->n;
  info->step = sim->step;
  info->kernel_launches = sim->launches;
  info->grid_origin_x = h.built.gx0;
  info->grid_origin_y = h.built.gy0;
  info->cell_edge = 1 << h.built.eshift;
  info->cells_x = h.built.ncx;
  info->cells_y = h.built.ncy;
  info->device_error = h.err;
  return CABS_OK;
}

extern "C" int cabs_event_record(cabs_sim *sim, int slot) {
  if (!sim || slot < 0 || slot >= CABS_NUM_EVENTS) return fail(sim, CABS_ERR_INVALID, "cabs_event_record: bad slot");
  CK(cudaSetDevice(sim->device));
  if (!sim->events[slot]) CK(cudaEventCreate(&sim->events[slot]));
  CK(cudaEventRecord(sim->events[slot], sim->stream));
  return CABS_OK;
}

extern "C" int cabs_event_elapsed_ms(cabs_sim *sim, int a, int b, double *ms) {
  if (!sim || !ms || a < 0 || b < 0 || a >= CABS_NUM_EVENTS || b >= CABS_NUM_EVENTS || !sim->events[a] ||
      !sim->events[b])
    return fail(sim, CABS_ERR_INVALID, "cabs_event_elapsed_ms: bad slot");
  CK(cudaSetDevice(sim->device));
  CK(cudaEventSynchronize(sim->events[b]));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, sim->events[a], sim->events[b]));
  *ms = (double)f;
  return CABS_OK;
}

extern "C" int cabs_profile_enable(cabs_sim *sim, int on) {
  if (!sim) return fail(sim, CABS_ERR_INVALID, "cabs_profile_enable: null handle");
  CK(cudaSetDevice(sim->device));
  CK(cudaStreamSynchronize(sim->stream));
  sim->profiling = on != 0;
  sim->prof_used = 0;
  return CABS_OK;
}

extern "C" int cabs_profile_read(cabs_sim *sim, cabs_kernel_time *out, int capacity, int *count) {
  if (!sim || !out || !count) return fail(sim, CABS_ERR_INVALID, "cabs_profile_read: null argument");
  CK(cudaSetDevice(sim->device));
  CK(cudaStreamSynchronize(sim->stream));
  int n = 0;
  for (size_t i = 0; i + 1 < sim->prof_used; ++i) {
    const char *name = sim->prof_names[i];
    if (!name) continue;  // closing mark of a batch: the gap to the next batch is host time, not a kernel
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, sim->prof_events[i], sim->prof_events[i + 1]));
    int k = 0;
    for (; k < n; ++k)
      if (strncmp(out[k].name, name, sizeof(out[k].name) - 1) == 0) break;
    if (k == n) {
      if (n == capacity) continue;
      memset(&out[n], 0, sizeof(out[n]));
      strncpy(out[n].name, name, sizeof(out[n].name) - 1);
      ++n;
    }
    out[k].launches += 1;
    out[k].total_ms += (double)ms;
  }
  *count = n;
  sim->prof_used = 0;
  return CABS_OK;
}

extern "C" int cabs_selftest_normals(int device, unsigned sweep_log2, double out[6]) {
  if (!out || sweep_log2 > 40) return CABS_ERR_INVALID;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return util_fail("cudaSetDevice", e);
  unsigned long long *d = nullptr, h[6];
  if ((e = cudaMalloc(&d, sizeof h)) != cudaSuccess) return util_fail("cudaMalloc", e);
  cudaMemset(d, 0, sizeof h);
  k_selftest_normals<<<148 * 8, 256>>>(d, sweep_log2);
  e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return util_fail("k_selftest_normals", e);
  for (int k = 0; k < 2; ++k) {
    const unsigned bits = (unsigned)h[k];
    float f;
    memcpy(&f, &bits, sizeof f);
    out[k] = (double)f;
  }
  for (int k = 2; k < 6; ++k) out[k] = (double)h[k];
  return CABS_OK;
}
