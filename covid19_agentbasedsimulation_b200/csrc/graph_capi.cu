// C ABI of the GraphSimulation path (include/covidabs.h, cabs_graph_*): host orchestration of graph_kernels.cuh.
// No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <nvtx3/nvToolsExt.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "../../include/covidabs.h"
#include "graph_kernels.cuh"

using namespace cabs::graph;

// NVTX range around every entry point that enqueues or waits for device work (see abs_capi.cu)
struct NvtxRangeG {
  explicit NvtxRangeG(const char *name) { nvtxRangePushA(name); }
  ~NvtxRangeG() { nvtxRangePop(); }
};

static thread_local std::string g_graph_create_error;

struct cabs_graph {
  int device = 0;
  cudaStream_t stream = nullptr;
  cabs_graph_config cfg;
  Sim *sims = nullptr;            // device array of descriptors
  std::vector<Sim> host_sims;     // mirror (pointers + parameters); run state lives on the device
  std::vector<char> uploaded;
  std::vector<int> np, nh, nb;
  // device slabs, one per column, n_sims * max_* elements
  int *px = nullptr, *py = nullptr, *house = nullptr, *employer = nullptr, *hire_seq = nullptr;
  uint32_t *attr = nullptr;
  long long *pwealth = nullptr;
  unsigned long long *events = nullptr;
  HouseRec *hrec = nullptr;
  int *hstratum = nullptr, *hsize = nullptr;
  long long *hfixed = nullptr;
  int *bx = nullptr, *by = nullptr, *bstratum = nullptr, *bstocks = nullptr, *bsales = nullptr, *bnum = nullptr;
  double *bprice = nullptr;
  long long *bwealth = nullptr, *bfixed = nullptr, *bincomes = nullptr;
  unsigned *cell_count = nullptr, *req_idx = nullptr;
  unsigned long long *cell_keys = nullptr;
  int4 *cell_items = nullptr, *src_list = nullptr;
  double *history = nullptr, *initial_rows = nullptr;
  unsigned cell_cap = 0;
  size_t pstride = 0;            // persons per simulation in the device arrays: max_persons rounded up to 4 (16-byte loads)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  uint64_t launches = 0;
  int slots = 0;                 // blocks of k_graph_run the device holds at once (occupancy x SMs)
  int *queue = nullptr;          // task-queue schedule: [0] ticket counter, [1 + sim] hours done in the current launch
  std::string err;
};

#define GCK(call)                                                                               \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      char buf_[512];                                                                           \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      g->err = buf_;                                                                            \
      return CABS_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

static int gfail(cabs_graph *g, int code, const char *msg) {
  if (g) g->err = msg;
  else g_graph_create_error = msg;
  return code;
}

extern "C" const char *cabs_graph_last_error(const cabs_graph *g) {
  return g ? g->err.c_str() : g_graph_create_error.c_str();
}

extern "C" void cabs_graph_destroy(cabs_graph *g) {
  if (!g) return;
  cudaSetDevice(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  void *ptrs[] = {g->sims, g->px, g->py, g->house, g->employer, g->hire_seq, g->attr, g->pwealth, g->events, g->hrec,
                  g->hstratum, g->hsize, g->hfixed, g->bx, g->by, g->bstratum,
                  g->bstocks, g->bsales, g->bnum, g->bprice, g->bwealth, g->bfixed, g->bincomes, g->cell_count,
                  g->cell_items, g->src_list, g->cell_keys, g->req_idx, g->history, g->initial_rows, g->queue};
  for (void *p : ptrs) cudaFree(p);
  if (g->ev0) cudaEventDestroy(g->ev0);
  if (g->ev1) cudaEventDestroy(g->ev1);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
}

template <typename T>
static cudaError_t alloc(T **p, size_t count) {
  cudaError_t e = cudaMalloc(p, (count ? count : 1) * sizeof(T));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, (count ? count : 1) * sizeof(T));
  return e;
}

static int graph_create_impl(cabs_graph *g, const cabs_graph_config *cfg) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return gfail(g, CABS_ERR_NOGPU, "no CUDA device visible: the graph step has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return gfail(g, CABS_ERR_INVALID, "device ordinal out of range");
  if (cfg->n_sims == 0 || cfg->max_persons == 0 || cfg->max_business == 0 || cfg->max_business > CABS_GRAPH_MAX_BUSINESS)
    return gfail(g, CABS_ERR_INVALID, "cabs_graph_create: need n_sims > 0, max_persons > 0, 1 <= max_business <= 16");
  g->cfg = *cfg;
  if (g->cfg.history_capacity == 0) g->cfg.history_capacity = 2048;
  g->device = cfg->device;
  GCK(cudaSetDevice(g->device));
  GCK(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
  GCK(cudaEventCreate(&g->ev0));
  GCK(cudaEventCreate(&g->ev1));
  g->pstride = ((size_t)cfg->max_persons + 3) & ~(size_t)3;
  const size_t S = cfg->n_sims, P = S * g->pstride, H = S * (cfg->max_houses ? cfg->max_houses : 1),
               B = S * cfg->max_business;
  // slots of the contact table (graph_kernels.cuh): a power of two >= 2 x max_persons, so that the cells holding a
  // source (at most one per person) never load it beyond one half; cfg->cell_capacity can only raise it
  unsigned cc = 1024;
  while (cc < 2u * cfg->max_persons && cc < (1u << 30)) cc <<= 1;
  while (cc < cfg->cell_capacity && cc < (1u << 30)) cc <<= 1;
  g->cell_cap = cc;
  GCK(alloc(&g->sims, S));
  GCK(alloc(&g->px, P)); GCK(alloc(&g->py, P)); GCK(alloc(&g->house, P)); GCK(alloc(&g->employer, P));
  GCK(alloc(&g->hire_seq, P)); GCK(alloc(&g->attr, P)); GCK(alloc(&g->pwealth, P)); GCK(alloc(&g->events, P));
  GCK(alloc(&g->hrec, H)); GCK(alloc(&g->hstratum, H)); GCK(alloc(&g->hsize, H));
  GCK(alloc(&g->hfixed, H));
  GCK(alloc(&g->bx, B)); GCK(alloc(&g->by, B)); GCK(alloc(&g->bstratum, B)); GCK(alloc(&g->bstocks, B));
  GCK(alloc(&g->bsales, B)); GCK(alloc(&g->bnum, B)); GCK(alloc(&g->bprice, B)); GCK(alloc(&g->bwealth, B));
  GCK(alloc(&g->bfixed, B)); GCK(alloc(&g->bincomes, B));
  GCK(alloc(&g->cell_count, S * ((size_t)cc + 4))); GCK(alloc(&g->cell_keys, S * (size_t)cc)); GCK(alloc(&g->cell_items, P)); GCK(alloc(&g->src_list, P)); GCK(alloc(&g->req_idx, P));
  GCK(alloc(&g->history, S * (size_t)g->cfg.history_capacity * kStats));
  GCK(alloc(&g->initial_rows, S * kStats));
  g->host_sims.assign(S, Sim());
  g->uploaded.assign(S, 0);
  g->np.assign(S, 0); g->nh.assign(S, 0); g->nb.assign(S, 0);
  return CABS_OK;
}

extern "C" int cabs_graph_create(const cabs_graph_config *cfg, cabs_graph **out) {
  if (!cfg || !out) return gfail(nullptr, CABS_ERR_INVALID, "cabs_graph_create: null argument");
  if (cfg->abi_version != CABS_ABI_VERSION) return gfail(nullptr, CABS_ERR_INVALID, "cabs_graph_create: ABI version mismatch");
  cabs_graph *g = new (std::nothrow) cabs_graph();
  if (!g) return gfail(nullptr, CABS_ERR_INVALID, "cabs_graph_create: out of host memory");
  const int rc = graph_create_impl(g, cfg);
  if (rc != CABS_OK) {
    g_graph_create_error = g->err;
    cabs_graph_destroy(g);
    *out = nullptr;
    return rc;
  }
  *out = g;
  return CABS_OK;
}

// host copies of the device-side bounds (the CPU restatement of the tests states the same)
static long long bound_lt(double p) {
  const double t = ceil(p * 4294967296.0 - 0.5);
  return t <= 0.0 ? 0ll : t >= 4294967296.0 ? 4294967296ll : (long long)t;
}
static long long bound_le(double r) {
  const double t = floor(r * 4294967296.0 - 0.5);
  return !(t >= 0.0) ? -1ll : t >= 4294967295.0 ? 4294967295ll : (long long)t;
}
static int threshold_T(double r) {
  if (!(r >= 0.0)) return -1;
  if (r > 16000.0) r = 16000.0;
  long long n = (long long)floor(r * r);
  while (sqrt((double)(n + 1)) <= r) ++n;
  while (n >= 0 && sqrt((double)n) > r) --n;
  return (int)n;
}
static long long to_fx(double v, double scale) { return (long long)nearbyint(v * scale); }  // ties to even

template <typename T>
static cudaError_t up(T *dst, const T *src, size_t n, cudaStream_t s) {
  return n ? cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, s) : cudaSuccess;
}
static cudaError_t up_money(long long *dst, const double *src, size_t n, double scale, cudaStream_t s,
                            std::vector<std::vector<long long>> &keep) {
  keep.emplace_back(n);
  for (size_t k = 0; k < n; ++k) keep.back()[k] = src ? to_fx(src[k], scale) : 0;
  return n ? cudaMemcpyAsync(dst, keep.back().data(), n * sizeof(long long), cudaMemcpyHostToDevice, s) : cudaSuccess;
}

extern "C" int cabs_graph_upload(cabs_graph *g, uint32_t sim, const cabs_graph_world *w, const cabs_graph_params *p) {
  NvtxRangeG nvtx_("cabs_graph_upload");
  if (!g || !w || !p) return gfail(g, CABS_ERR_INVALID, "cabs_graph_upload: null argument");
  if (sim >= g->cfg.n_sims) return gfail(g, CABS_ERR_INVALID, "cabs_graph_upload: simulation index out of range");
  if (w->n_persons < 0 || (uint32_t)w->n_persons > g->cfg.max_persons || w->n_houses < 0 ||
      (uint32_t)w->n_houses > g->cfg.max_houses || w->n_business < 1 || (uint32_t)w->n_business > g->cfg.max_business)
    return gfail(g, CABS_ERR_CAPACITY, "cabs_graph_upload: world larger than the capacity of the handle");
  if (p->money_bits < 0 || p->money_bits > 40) return gfail(g, CABS_ERR_INVALID, "money_bits must be in [0, 40]");
  if (p->recovering_time < 0 || p->recovering_time > 250 || p->contagion_time > 250)
    return gfail(g, CABS_ERR_INVALID, "recovering_time / contagion_time must fit the 8-bit infected_time");
  if (p->population_size == 0) return gfail(g, CABS_ERR_INVALID, "population_size must be > 0");
  const size_t n = w->n_persons, nh = w->n_houses, nb = w->n_business;
  if (n && (!w->px || !w->py || !w->status || !w->severity || !w->infected_time || !w->age || !w->stratum || !w->active ||
            !w->isolated || !w->house || !w->employer || !w->hire_seq || !w->pwealth))
    return gfail(g, CABS_ERR_INVALID, "cabs_graph_upload: null person column");
  if (nh && (!w->hx || !w->hy || !w->hstratum || !w->hsize || !w->hwealth || !w->hfixed))
    return gfail(g, CABS_ERR_INVALID, "cabs_graph_upload: null house column");
  if (!w->bx || !w->by || !w->bstratum || !w->bstocks || !w->bnum_employees || !w->bprice || !w->bwealth || !w->bfixed)
    return gfail(g, CABS_ERR_INVALID, "cabs_graph_upload: null business column");
  GCK(cudaSetDevice(g->device));
  const double scale = ldexp(1.0, p->money_bits);
  const size_t op = (size_t)sim * g->pstride, oh = (size_t)sim * (g->cfg.max_houses ? g->cfg.max_houses : 1),
               ob = (size_t)sim * g->cfg.max_business;
  std::vector<std::vector<long long>> keep;  // staging that must outlive the async copies
  std::vector<uint32_t> attr(n);
  int max_seq = -1;
  for (size_t i = 0; i < n; ++i) {
    if (w->house[i] >= (int)nh || w->employer[i] >= (int)nb)
      return gfail(g, CABS_ERR_INVALID, "cabs_graph_upload: house / employer index out of range");
    attr[i] = (w->status[i] & 3u) | ((w->severity[i] & 3u) << 2) | ((w->stratum[i] & 7u) << 4) |
              ((uint32_t)w->infected_time[i] << 8) | ((uint32_t)w->age[i] << 16) | (w->active[i] ? kActive : 0u) |
              (w->isolated[i] ? kIsolated : 0u);
    if (w->hire_seq[i] > max_seq) max_seq = w->hire_seq[i];
  }
  GCK(up(g->px + op, w->px, n, g->stream)); GCK(up(g->py + op, w->py, n, g->stream));
  GCK(up(g->house + op, w->house, n, g->stream)); GCK(up(g->employer + op, w->employer, n, g->stream));
  GCK(up(g->hire_seq + op, w->hire_seq, n, g->stream)); GCK(up(g->attr + op, attr.data(), n, g->stream));
  GCK(up_money(g->pwealth + op, w->pwealth, n, scale, g->stream, keep));
  std::vector<HouseRec> hrec(nh);   // outlives the async copy: the call synchronises before returning
  for (size_t h = 0; h < nh; ++h) {
    hrec[h].x = w->hx[h]; hrec[h].y = w->hy[h];
    hrec[h].wealth = to_fx(w->hwealth[h], scale);
    hrec[h].expenses = w->hexpenses ? to_fx(w->hexpenses[h], scale) : 0;
    hrec[h].incomes = w->hincomes ? to_fx(w->hincomes[h], scale) : 0;
  }
  GCK(up(g->hrec + oh, hrec.data(), nh, g->stream));
  GCK(up(g->hstratum + oh, w->hstratum, nh, g->stream)); GCK(up(g->hsize + oh, w->hsize, nh, g->stream));
  GCK(up_money(g->hfixed + oh, w->hfixed, nh, scale, g->stream, keep));
  GCK(up(g->bx + ob, w->bx, nb, g->stream)); GCK(up(g->by + ob, w->by, nb, g->stream));
  GCK(up(g->bstratum + ob, w->bstratum, nb, g->stream)); GCK(up(g->bstocks + ob, w->bstocks, nb, g->stream));
  GCK(up(g->bnum + ob, w->bnum_employees, nb, g->stream)); GCK(up(g->bprice + ob, w->bprice, nb, g->stream));
  if (w->bsales) GCK(up(g->bsales + ob, w->bsales, nb, g->stream));
  else GCK(cudaMemsetAsync(g->bsales + ob, 0, nb * sizeof(int), g->stream));
  GCK(up_money(g->bwealth + ob, w->bwealth, nb, scale, g->stream, keep));
  GCK(up_money(g->bfixed + ob, w->bfixed, nb, scale, g->stream, keep));
  GCK(up_money(g->bincomes + ob, w->bincomes, nb, scale, g->stream, keep));

  Sim s;
  memset(&s, 0, sizeof s);
  s.n = (int)n; s.nh = (int)nh; s.nb = (int)nb;
  s.px = g->px + op; s.py = g->py + op; s.attr = g->attr + op; s.house = g->house + op; s.employer = g->employer + op;
  s.hire_seq = g->hire_seq + op; s.pwealth = g->pwealth + op; s.events = g->events + op;
  s.hrec = g->hrec + oh; s.hstratum = g->hstratum + oh; s.hsize = g->hsize + oh;
  s.hfixed = g->hfixed + oh;
  s.bx = g->bx + ob; s.by = g->by + ob; s.bstratum = g->bstratum + ob; s.bstocks = g->bstocks + ob;
  s.bsales = g->bsales + ob; s.bnum = g->bnum + ob; s.bprice = g->bprice + ob; s.bwealth = g->bwealth + ob;
  s.bfixed = g->bfixed + ob; s.bincomes = g->bincomes + ob;
  s.gov_x = w->gov_x; s.gov_y = w->gov_y; s.health_x = w->health_x; s.health_y = w->health_y;
  s.gov_wealth = to_fx(w->gov_wealth, scale); s.gov_fixed = to_fx(w->gov_fixed, scale);
  s.health_wealth = to_fx(w->health_wealth, scale); s.health_fixed = to_fx(w->health_fixed, scale);
  s.health_expenses = to_fx(w->health_expenses, scale);
  s.gov_price = w->gov_price;
  for (int k = 0; k < 4; ++k) s.amp[k] = p->amplitudes[k];
  s.amp[3] = 0.0;
  for (int k = 0; k < 5; ++k) s.bi[k] = p->basic_income[k];
  s.min_income = p->minimum_income; s.min_expense = p->minimum_expense; s.total_wealth = p->total_wealth;
  s.critical_limit = p->critical_limit; s.pop_size = (double)p->population_size; s.scale = scale;
  for (int k = 0; k < 9; ++k) {
    s.thr_hosp[k] = bound_lt(p->age_hospitalization_probs[k]);
    s.thr_sev[k] = bound_lt(p->age_severe_probs[k]);
    s.thr_death[k] = bound_lt(p->age_death_probs[k]);
  }
  s.thr_rate = bound_le(p->contagion_rate);
  s.T = threshold_T(p->contagion_distance);
  s.T_bus = threshold_T(p->business_distance);
  int reach = s.T > 0 ? (int)sqrt((double)s.T) : 0;
  while ((long long)reach * reach > s.T) --reach;
  while ((long long)(reach + 1) * (reach + 1) <= s.T) ++reach;
  s.reach = reach < 1 ? 1 : reach;
  s.incubation = p->incubation_time; s.contagion_time = p->contagion_time; s.recovering = p->recovering_time;
  s.mode = p->mode; s.sleep = p->sleep;
  s.k0 = (unsigned)(p->seed & 0xFFFFFFFFu); s.k1 = (unsigned)(p->seed >> 32);
  s.iteration = w->iteration;
  s.next_hire_seq = max_seq + 1;
  s.cell_count = g->cell_count + (size_t)sim * ((size_t)g->cell_cap + 4);   // 16-byte aligned slices
  s.cell_keys = g->cell_keys + (size_t)sim * (size_t)g->cell_cap;
  s.cell_items = g->cell_items + op; s.src_list = g->src_list + op; s.req_idx = g->req_idx + op;
  s.cell_cap = (int)g->cell_cap;
  s.history = g->history + (size_t)sim * g->cfg.history_capacity * kStats;
  s.hist_cap = (int)g->cfg.history_capacity;
  g->host_sims[sim] = s;
  GCK(cudaMemcpyAsync(g->sims + sim, &g->host_sims[sim], sizeof(Sim), cudaMemcpyHostToDevice, g->stream));
  GCK(cudaStreamSynchronize(g->stream));  // staging buffers go out of scope
  g->uploaded[sim] = 1;
  g->np[sim] = (int)n; g->nh[sim] = (int)nh; g->nb[sim] = (int)nb;
  // statistics of the initial state (memo of the first hour, row -1)
  k_graph_initial_stats<<<1, kThreads, 0, g->stream>>>(g->sims + sim, g->initial_rows + (size_t)sim * kStats);
  g->launches++;
  GCK(cudaGetLastError());
  return CABS_OK;
}

extern "C" int cabs_graph_run(cabs_graph *g, int iterations) {
  NvtxRangeG nvtx_("cabs_graph_run");
  if (!g || iterations < 0) return gfail(g, CABS_ERR_INVALID, "cabs_graph_run: bad argument");
  for (char u : g->uploaded)
    if (!u) return gfail(g, CABS_ERR_INVALID, "cabs_graph_run: every simulation of the handle must be uploaded first");
  GCK(cudaSetDevice(g->device));
  GCK(cudaEventRecord(g->ev0, g->stream));
  if (iterations > 0) {
    // simulations that fit the GPU at once keep one block each for the whole run; larger ensembles are scheduled hour
    // by hour from a ticket queue over a persistent grid (see k_graph_run)
    if (g->slots == 0) {
      int per_sm = 0, sms = 0;
      GCK(cudaFuncSetAttribute(k_graph_run<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGraphDynSmem));
      GCK(cudaFuncSetAttribute(k_graph_run<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGraphDynSmem));
      // one block of 1 024 threads per SM: 18.8 KB static + 64 KB dynamic (+ 1 KB reserved) -> the 100 KB carveout, 156 KB
      // of L1.  The carveout is left to the driver; L1 matters here (two 512-thread blocks: 164 KB carveout 2.63e10
      // agent-updates/s on the bench, 196 KB 2.43e10, 228 KB 1.77e10: profiles/r2_graph_ab_carveout_{1,2}.log)
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g->device);
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_graph_run<true>, kThreads, kGraphDynSmem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
      g->slots = per_sm * sms;
    }
    const char *force = getenv("CABS_GRAPH_QUEUE");   // "0" / "1": force a schedule (measurements, tests)
    const bool queue = force ? force[0] == '1' : (int)g->cfg.n_sims > g->slots;
    if (queue) {
      if (!g->queue) GCK(cudaMalloc(&g->queue, (1 + (size_t)g->cfg.n_sims) * sizeof(int)));
      GCK(cudaMemsetAsync(g->queue, 0, (1 + (size_t)g->cfg.n_sims) * sizeof(int), g->stream));
      const int grid = g->slots < (int)g->cfg.n_sims ? g->slots : (int)g->cfg.n_sims;
      k_graph_run<true><<<grid, kThreads, kGraphDynSmem, g->stream>>>(g->sims, iterations, (int)g->cfg.n_sims,
                                                          reinterpret_cast<unsigned *>(g->queue), g->queue + 1);
    } else {
      k_graph_run<false><<<g->cfg.n_sims, kThreads, kGraphDynSmem, g->stream>>>(g->sims, iterations, (int)g->cfg.n_sims, nullptr, nullptr);
    }
    g->launches++;
  }
  GCK(cudaEventRecord(g->ev1, g->stream));
  g->timed = true;
  GCK(cudaGetLastError());
  return CABS_OK;
}

extern "C" int cabs_graph_sync(cabs_graph *g) {
  if (!g) return gfail(g, CABS_ERR_INVALID, "cabs_graph_sync: null handle");
  GCK(cudaSetDevice(g->device));
  GCK(cudaStreamSynchronize(g->stream));
  return CABS_OK;
}

extern "C" int cabs_graph_last_run_ms(cabs_graph *g, double *ms) {
  if (!g || !ms) return gfail(g, CABS_ERR_INVALID, "cabs_graph_last_run_ms: null argument");
  if (!g->timed) return gfail(g, CABS_ERR_INVALID, "cabs_graph_last_run_ms: no run yet");
  GCK(cudaSetDevice(g->device));
  GCK(cudaEventSynchronize(g->ev1));
  float f = 0.f;
  GCK(cudaEventElapsedTime(&f, g->ev0, g->ev1));
  *ms = (double)f;
  return CABS_OK;
}

extern "C" int cabs_graph_stats(cabs_graph *g, uint32_t sim, int64_t first, uint32_t count, double *rows) {
  NvtxRangeG nvtx_("cabs_graph_stats");
  if (!g || (count && !rows)) return gfail(g, CABS_ERR_INVALID, "cabs_graph_stats: null argument");
  if (sim >= g->cfg.n_sims || !g->uploaded[sim]) return gfail(g, CABS_ERR_INVALID, "cabs_graph_stats: bad simulation index");
  GCK(cudaSetDevice(g->device));
  Sim cur;
  GCK(cudaMemcpyAsync(&cur, g->sims + sim, sizeof(Sim), cudaMemcpyDeviceToHost, g->stream));
  GCK(cudaStreamSynchronize(g->stream));
  const int64_t last = cur.iteration;  // rows exist for iterations -1 .. last
  if (first < -1 || first + (int64_t)count - 1 > last)
    return gfail(g, CABS_ERR_INVALID, "cabs_graph_stats: iterations not executed yet");
  if (last - first >= (int64_t)cur.hist_cap)
    return gfail(g, CABS_ERR_CAPACITY, "cabs_graph_stats: rows already overwritten (history_capacity)");
  // the rows of a run are contiguous in the ring except where it wraps: at most three copies
  uint32_t k = 0;
  if (count && first < 0) {
    GCK(cudaMemcpyAsync(rows, g->initial_rows + (size_t)sim * kStats, kStats * sizeof(double), cudaMemcpyDeviceToHost,
                        g->stream));
    k = 1;
  }
  while (k < count) {
    const int64_t slot = (first + k) % cur.hist_cap;
    const uint32_t run = (uint32_t)std::min<int64_t>(count - k, cur.hist_cap - slot);
    GCK(cudaMemcpyAsync(rows + (size_t)k * kStats, cur.history + (size_t)slot * kStats,
                        (size_t)run * kStats * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
    k += run;
  }
  GCK(cudaStreamSynchronize(g->stream));
  return CABS_OK;
}

// The rows of every simulation of the handle in one strided copy per contiguous piece of the ring:
// rows[sim][k][metric], k = first_iteration + 0 .. count-1.
extern "C" int cabs_graph_stats_all(cabs_graph *g, int64_t first, uint32_t count, double *rows) {
  NvtxRangeG nvtx_("cabs_graph_stats_all");
  if (!g || (count && !rows)) return gfail(g, CABS_ERR_INVALID, "cabs_graph_stats_all: null argument");
  for (uint32_t s = 0; s < g->cfg.n_sims; ++s)
    if (!g->uploaded[s]) return gfail(g, CABS_ERR_INVALID, "cabs_graph_stats_all: every simulation of the handle must be uploaded");
  if (count == 0) return CABS_OK;
  GCK(cudaSetDevice(g->device));
  Sim cur;
  GCK(cudaMemcpyAsync(&cur, g->sims, sizeof(Sim), cudaMemcpyDeviceToHost, g->stream));
  GCK(cudaStreamSynchronize(g->stream));
  const int64_t last = cur.iteration;
  if (first < -1 || first + (int64_t)count - 1 > last)
    return gfail(g, CABS_ERR_INVALID, "cabs_graph_stats_all: iterations not executed yet");
  if (last - first >= (int64_t)cur.hist_cap)
    return gfail(g, CABS_ERR_CAPACITY, "cabs_graph_stats_all: rows already overwritten (history_capacity)");
  const size_t row = kStats * sizeof(double), n_sims = g->cfg.n_sims;
  const size_t dst_pitch = (size_t)count * row, src_pitch = (size_t)cur.hist_cap * row;
  uint32_t k = 0;
  if (first < 0) {   // the initial rows live in their own array, one per simulation
    GCK(cudaMemcpy2DAsync(rows, dst_pitch, g->initial_rows, row, row, n_sims, cudaMemcpyDeviceToHost, g->stream));
    k = 1;
  }
  while (k < count) {
    const int64_t slot = (first + k) % cur.hist_cap;
    const uint32_t run = (uint32_t)std::min<int64_t>(count - k, cur.hist_cap - slot);
    GCK(cudaMemcpy2DAsync(rows + (size_t)k * kStats, dst_pitch, g->history + (size_t)slot * kStats, src_pitch,
                          (size_t)run * row, n_sims, cudaMemcpyDeviceToHost, g->stream));
    k += run;
  }
  GCK(cudaStreamSynchronize(g->stream));
  return CABS_OK;
}

extern "C" int cabs_graph_aggregate(cabs_graph *g, int64_t first, uint32_t count, double *out) {
  NvtxRangeG nvtx_("cabs_graph_aggregate");
  if (!g || (count && !out)) return gfail(g, CABS_ERR_INVALID, "cabs_graph_aggregate: null argument");
  for (uint32_t s = 0; s < g->cfg.n_sims; ++s)
    if (!g->uploaded[s]) return gfail(g, CABS_ERR_INVALID, "cabs_graph_aggregate: every simulation of the handle must be uploaded");
  if (count == 0) return CABS_OK;
  GCK(cudaSetDevice(g->device));
  Sim cur;
  GCK(cudaMemcpyAsync(&cur, g->sims, sizeof(Sim), cudaMemcpyDeviceToHost, g->stream));
  GCK(cudaStreamSynchronize(g->stream));
  const int64_t last = cur.iteration;  // every simulation of a handle is at the same iteration
  if (first < -1 || first + (int64_t)count - 1 > last)
    return gfail(g, CABS_ERR_INVALID, "cabs_graph_aggregate: iterations not executed yet");
  if (last - first >= (int64_t)cur.hist_cap)
    return gfail(g, CABS_ERR_CAPACITY, "cabs_graph_aggregate: rows already overwritten (history_capacity)");
  double *dev = nullptr;
  const size_t bytes = (size_t)count * kStats * 4 * sizeof(double);
  GCK(cudaMalloc(&dev, bytes));
  k_graph_aggregate<<<count, kStats * 32, 0, g->stream>>>(g->history, g->initial_rows, (int)g->cfg.n_sims, cur.hist_cap,
                                                          (long long)first, dev);
  g->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dev, bytes, cudaMemcpyDeviceToHost, g->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g->stream);
  cudaFree(dev);
  GCK(e);
  return CABS_OK;
}

extern "C" int cabs_graph_moments(cabs_graph *g, uint32_t sim_first, uint32_t sim_count, int64_t first, uint32_t count,
                                  double *out) {
  NvtxRangeG nvtx_("cabs_graph_moments");
  if (!g || (count && !out)) return gfail(g, CABS_ERR_INVALID, "cabs_graph_moments: null argument");
  if (sim_count == 0 || sim_first + (uint64_t)sim_count > g->cfg.n_sims)
    return gfail(g, CABS_ERR_INVALID, "cabs_graph_moments: simulation range outside the handle");
  for (uint32_t s = sim_first; s < sim_first + sim_count; ++s)
    if (!g->uploaded[s]) return gfail(g, CABS_ERR_INVALID, "cabs_graph_moments: every simulation of the range must be uploaded");
  if (count == 0) return CABS_OK;
  GCK(cudaSetDevice(g->device));
  Sim cur;
  GCK(cudaMemcpyAsync(&cur, g->sims, sizeof(Sim), cudaMemcpyDeviceToHost, g->stream));
  GCK(cudaStreamSynchronize(g->stream));
  const int64_t last = cur.iteration;
  if (first < 0 || first + (int64_t)count - 1 > last)
    return gfail(g, CABS_ERR_INVALID, "cabs_graph_moments: iterations not executed yet");
  if (last - first >= (int64_t)cur.hist_cap)
    return gfail(g, CABS_ERR_CAPACITY, "cabs_graph_moments: rows already overwritten (history_capacity)");
  double *dev = nullptr;
  const size_t bytes = (size_t)count * kStats * 4 * sizeof(double);
  GCK(cudaMalloc(&dev, bytes));
  k_graph_moments<<<count, kStats * 32, 0, g->stream>>>(g->history, (int)sim_first, (int)sim_count, cur.hist_cap,
                                                        (long long)first, dev);
  g->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dev, bytes, cudaMemcpyDeviceToHost, g->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g->stream);
  cudaFree(dev);
  GCK(e);
  return CABS_OK;
}

template <typename T>
static cudaError_t down(T *dst, const T *src, size_t n, cudaStream_t s) {
  return (n && dst) ? cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, s) : cudaSuccess;
}

extern "C" int cabs_graph_download(cabs_graph *g, uint32_t sim, cabs_graph_world *w) {
  NvtxRangeG nvtx_("cabs_graph_download");
  if (!g || !w) return gfail(g, CABS_ERR_INVALID, "cabs_graph_download: null argument");
  if (sim >= g->cfg.n_sims || !g->uploaded[sim]) return gfail(g, CABS_ERR_INVALID, "cabs_graph_download: bad simulation index");
  const size_t n = g->np[sim], nh = g->nh[sim], nb = g->nb[sim];
  if ((size_t)w->n_persons < n || (size_t)w->n_houses < nh || (size_t)w->n_business < nb)
    return gfail(g, CABS_ERR_CAPACITY, "cabs_graph_download: host arrays smaller than the world");
  GCK(cudaSetDevice(g->device));
  const Sim &s = g->host_sims[sim];
  Sim cur;
  GCK(cudaMemcpyAsync(&cur, g->sims + sim, sizeof(Sim), cudaMemcpyDeviceToHost, g->stream));
  std::vector<uint32_t> attr(n);
  std::vector<long long> pw(n), hf(nh), bw(nb), bf(nb), bi(nb);
  std::vector<HouseRec> hrec(nh);
  GCK(down(w->px, s.px, n, g->stream)); GCK(down(w->py, s.py, n, g->stream));
  GCK(down(w->house, s.house, n, g->stream)); GCK(down(w->employer, s.employer, n, g->stream));
  GCK(down(w->hire_seq, s.hire_seq, n, g->stream)); GCK(down(attr.data(), s.attr, n, g->stream));
  GCK(down(pw.data(), s.pwealth, n, g->stream));
  GCK(down(hrec.data(), s.hrec, nh, g->stream));
  GCK(down(w->hstratum, s.hstratum, nh, g->stream)); GCK(down(w->hsize, s.hsize, nh, g->stream));
  GCK(down(hf.data(), s.hfixed, nh, g->stream));
  GCK(down(w->bx, s.bx, nb, g->stream)); GCK(down(w->by, s.by, nb, g->stream));
  GCK(down(w->bstratum, s.bstratum, nb, g->stream)); GCK(down(w->bstocks, s.bstocks, nb, g->stream));
  GCK(down(w->bsales, s.bsales, nb, g->stream)); GCK(down(w->bnum_employees, s.bnum, nb, g->stream));
  GCK(down(w->bprice, s.bprice, nb, g->stream));
  GCK(down(bw.data(), s.bwealth, nb, g->stream)); GCK(down(bf.data(), s.bfixed, nb, g->stream));
  GCK(down(bi.data(), s.bincomes, nb, g->stream));
  GCK(cudaStreamSynchronize(g->stream));
  const double scale = s.scale;
  for (size_t i = 0; i < n; ++i) {
    const uint32_t a = attr[i];
    if (w->status) w->status[i] = (uint8_t)(a & 3u);
    if (w->severity) w->severity[i] = (uint8_t)((a >> 2) & 3u);
    if (w->stratum) w->stratum[i] = (uint8_t)((a >> 4) & 7u);
    if (w->infected_time) w->infected_time[i] = (uint8_t)((a >> 8) & 0xFFu);
    if (w->age) w->age[i] = (uint8_t)((a >> 16) & 0xFFu);
    if (w->active) w->active[i] = (a & kActive) ? 1 : 0;
    if (w->isolated) w->isolated[i] = (a & kIsolated) ? 1 : 0;
    if (w->pwealth) w->pwealth[i] = (double)pw[i] / scale;
  }
  for (size_t h = 0; h < nh; ++h) {
    if (w->hx) w->hx[h] = hrec[h].x;
    if (w->hy) w->hy[h] = hrec[h].y;
    if (w->hwealth) w->hwealth[h] = (double)hrec[h].wealth / scale;
    if (w->hfixed) w->hfixed[h] = (double)hf[h] / scale;
    if (w->hincomes) w->hincomes[h] = (double)hrec[h].incomes / scale;
    if (w->hexpenses) w->hexpenses[h] = (double)hrec[h].expenses / scale;
  }
  for (size_t b = 0; b < nb; ++b) {
    if (w->bwealth) w->bwealth[b] = (double)bw[b] / scale;
    if (w->bfixed) w->bfixed[b] = (double)bf[b] / scale;
    if (w->bincomes) w->bincomes[b] = (double)bi[b] / scale;
  }
  w->n_persons = (int32_t)n; w->n_houses = (int32_t)nh; w->n_business = (int32_t)nb;
  w->gov_x = s.gov_x; w->gov_y = s.gov_y; w->health_x = s.health_x; w->health_y = s.health_y;
  w->gov_wealth = (double)cur.gov_wealth / scale; w->gov_price = s.gov_price; w->gov_fixed = (double)cur.gov_fixed / scale;
  w->health_wealth = (double)cur.health_wealth / scale; w->health_fixed = (double)cur.health_fixed / scale;
  w->health_expenses = (double)cur.health_expenses / scale;
  w->iteration = cur.iteration;
  return CABS_OK;
}

#ifdef CABS_GRAPH_TIMERS
// timing-only build: read and reset the per-phase clocks (graph_kernels.cuh); not part of include/covidabs.h
extern "C" int cabs_graph_phase_clocks(unsigned long long *out32) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(out32, cabs::graph::g_graph_phase_clk, 32 * sizeof(unsigned long long)) != cudaSuccess) return -1;
  unsigned long long zero[32] = {};
  return cudaMemcpyToSymbol(cabs::graph::g_graph_phase_clk, zero, sizeof zero) == cudaSuccess ? 0 : -1;
}
#endif
