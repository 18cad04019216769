// Device side of the COVID-ABS agent step for sm_100a (B200).
//
// Population layout in HBM (structure of arrays, kept physically ordered by cell key):
//   hot  uint4 {x, y, attr, id}: what the cell-list build and the contact scan read; one 128-bit
//              load/store per agent, two records per 32-byte sector
//        x,y   i32  lattice position (agents.py:45-47)
//        attr  u32  status[1:0] severity[3:2] stratum[6:4] new[7] | infected_time[15:8] | age[23:16]
//        id    u32  position in the reference's population list (abs.py:15); keys all draws
//   w    f64   wealth (agents.py:60)
//   aux  uint2 {displacement (ix, iy as int16 pair), rank inside the destination cell}: written by
//              pass A, read by pass B
// Two copies of hot/w (ping/pong): every step scatters the updated records straight into their
// cell-sorted slot of the other copy, so the contact scan reads contiguous cell ranges.
// Two cell tables (ping/pong): pass B of a step zeroes the table of the step before, so the
// histogram of the next step starts clean without a separate clearing pass.
//
// One step (Simulation.execute, abs.py:232-273), four launches:
//   k_move_count       pass A: new position of every agent -> cell key -> histogram; the value the
//                      atomic returns is the agent's rank inside its cell
//   k_scan_cells       single-pass exclusive scan (decoupled look-back) of the histogram = cell starts
//   k_update_scatter   pass B: move + update + economy (abs.py:156-230), partial statistics
//                      (abs.py:291-314), record -> slot cell_start[key] + rank, list of the slots
//                      that hold Infected agents
//   k_contagion        pass C: contact() (abs.py:141-154) pushed from the Infected agents to the
//                      Susceptible ones in reach; the last block to finish closes the step:
//                      statistics record, triggers (abs.py:267-271), grid of the next step
#pragma once
#include <limits.h>
#include <stdint.h>

#include <cooperative_groups.h>

#include "abs_rng.cuh"

namespace cabs {

constexpr int kMaxTriggers = 8;
constexpr int kNumStats = 12;

// sticky device error bits
constexpr uint32_t kErrOutOfGrid = 1u;     // an agent left the planned grid (internal invariant broken)
constexpr uint32_t kErrDomain = 2u;        // bounding box wider than 2^28
constexpr uint32_t kErrWealthRange = 4u;   // fixed-point wealth sum would overflow
constexpr uint32_t kErrExchange = 8u;      // slab exchange buffer overflow (migrants / ghosts)
constexpr uint32_t kErrBadIds = 32u;       // uploaded ids are not a permutation of 0..n-1

struct DevTrigger {
  int metric, op, attribute, pad;
  double threshold;
  double value[4];
};

// Raw statistics of one step: 7 counts (abs.py:300-307) + 5 fixed-point wealth sums (abs.py:309-312)
struct StatRecord {
  unsigned long long cnt[7];
  long long q[5];
  unsigned long long new_infections;
  unsigned long long n_resident;
  unsigned long long pad[2];
};

struct Grid {
  int gx0, gy0;       // lattice coordinates of cell (1,1)'s lower corner
  int eshift;         // cell edge = 1 << eshift lattice units
  int ncx, ncy;       // cells per row / rows, including one padding cell on each side
  unsigned ncells;    // ncx * ncy; cell_start has ncells + 1 entries
};

// Everything the kernels read that can change from step to step lives here (device memory),
// so that a whole run of steps can be enqueued without the host in the loop.
struct Ctl {
  // ---- cell grids: `g` is the one the next build (k_move_count .. k_contagion) uses, planned by
  //      the previous step's closing; `built` is the one the current cell table was built with
  Grid g, built;
  int last_bb[4];     // bounding box of the current positions (xmin, xmax, ymin, ymax)
  unsigned max_cells;      // entries the cell tables hold (fine budget)
  unsigned coarse_cells;   // budget while the epidemic is sparse
  unsigned cell_budget;    // the one in force for the next grid (chosen when a step closes)
  int pull_mode;           // contact pass direction for the next step: 0 push from the Infected list, 1 pull by the Susceptible
  unsigned cells_used[2];  // entries of each cell table that may be non-zero
  int cur_table;           // table the build in flight / the last build uses
  // ---- simulation attributes (abs.py:17-49); triggers rewrite them on the device
  float amp[4];
  double length, height;
  double contagion_distance, rate, critical_limit, min_expense;
  double bi[5];
  double p_hosp[9], p_sev[9], p_death[9];
  unsigned long long pop_size;
  int T;              // largest integer d2 with sqrt(d2) <= contagion_distance, -1: none
  int recovering_time;
  int scale_bits;
  int schedule;
  unsigned k0, k1;    // Philox key (seed)
  unsigned step;
  unsigned n;         // agents resident on this device
  int crit_active;    // (Severe+Hospitalization)/N >= critical_limit at the end of the previous step
  // The probability tests compare a table entry with u = (w + 0.5) * 2^-32 (w = 32 random bits); both
  // sides are exact in binary64, so each test is an integer comparison of w with a precomputed bound:
  //   p > u   <=>  w <  ceil(p * 2^32 - 0.5)          (abs.py:209,212,220)
  //   u <= r  <=>  w <= floor(r * 2^32 - 0.5)         (abs.py:152)
  long long thr_hosp[9], thr_sev[9], thr_death[9];
  long long thr_rate;
  // ---- slab decomposition: this device owns lattice columns [slab_lo, slab_hi)
  int slab_lo, slab_hi;
  int n_ranks, rank;  // n_ranks 1: no decomposition
  unsigned capacity;  // records the population arrays hold
  unsigned mig_cap, ghost_cap;  // records per exchange buffer
  int overflow;       // residents + immigrants would not fit `capacity`: the scatter passes stand down
  // ---- accumulators of the running step
  unsigned long long cnt[7];
  long long q[5];
  unsigned long long newinf;
  int bb_xmin, bb_xmax, bb_ymin, bb_ymax;
  unsigned n_inf;          // entries of the Infected-slot list
  unsigned n_emig;         // agents that leave this slab in the running step
  unsigned pack_ticket, imm_ticket;  // blocks of k_pack_emigrants / of pass B (slab mode) that have finished
  unsigned build_seq;      // direct slab mode: builds completed; exchange flags carry build_seq + 1 of the build in flight
  unsigned pending_step;   // direct slab mode: step whose global statistics the late closing will record
  int glob_ymin, glob_ymax;  // direct slab mode: y-range of the whole population one build ago
  unsigned n_newinf;       // sequential schedule: entries of the newly-infected list
  unsigned seq_changed;    // sequential schedule: infection times lowered in the last round
  unsigned scan_ticket;    // next tile of the scan
  unsigned done_ticket;    // blocks of k_contagion that have finished
  // ---- statistics of the previous step (what triggers and the critical-limit test read)
  StatRecord prev;
  StatRecord cur;
  int ntrig;
  DevTrigger trig[kMaxTriggers];
  unsigned err;
  unsigned hist_cap;
  // ---- off-lattice mode: binary64 positions, declarative per-agent rules (abs.py:86-95)
  unsigned trig_epoch;     // bumped whenever a device-side trigger fires (the host mirrors the attributes it rewrote)
  int f64;                 // positions are binary64 (pos arrays); hot.x, hot.y hold their floors
  int rule_margin;         // lattice units a `stay` move rule can displace an agent (5.78 sigma, rounded up)
  int rule_box[4];         // bounding box of the points `point` move rules send agents to (xmin > xmax: none)
};

// ------------------------------------------------------------------------------------------
// attribute word helpers
__device__ __forceinline__ uint32_t attr_status(uint32_t a) { return a & 3u; }
__device__ __forceinline__ uint32_t attr_severity(uint32_t a) { return (a >> 2) & 3u; }
__device__ __forceinline__ uint32_t attr_stratum(uint32_t a) { return (a >> 4) & 7u; }
__device__ __forceinline__ uint32_t attr_time(uint32_t a) { return (a >> 8) & 0xFFu; }
__device__ __forceinline__ uint32_t attr_age(uint32_t a) { return (a >> 16) & 0xFFu; }
// Bit 7 marks "became Infected in the last contact phase" (set by k_contagion while the status bits
// still say Susceptible, so that agents infected in a pass do not transmit in the same pass).  Every
// other reader folds the mark into the status first.
constexpr uint32_t kNewlyInfected = 0x80u;
__device__ __forceinline__ uint32_t attr_normalize(uint32_t a) {
  return (a & kNewlyInfected) ? ((a & ~(kNewlyInfected | 3u)) | 1u) : a;
}
__host__ __device__ __forceinline__ uint32_t attr_pack(uint32_t status, uint32_t sev, uint32_t stratum, uint32_t time,
                                                       uint32_t age) {
  return (status & 3u) | ((sev & 3u) << 2) | ((stratum & 7u) << 4) | ((time & 0xFFu) << 8) | ((age & 0xFFu) << 16);
}

enum : uint32_t { ST_S = 0, ST_I = 1, ST_R = 2, ST_D = 3 };
enum : uint32_t { SEV_E = 0, SEV_A = 1, SEV_H = 2, SEV_G = 3 };

// Per-block copy of the step-invariant part of Ctl that the agent kernels need.
struct StepParams {
  int gx0, gy0, eshift, ncx, ncy;
  float amp[4];
  double min_expense;
  double me32;      // min_expense * 2^-32: folds the scaling of the income uniform (exact, a power of two)
  double bi[5];
  double mebi[5];   // fl(min_expense * basic_income[stratum]): the daily expense of abs.py:230
  long long thr_hosp[9], thr_sev[9], thr_death[9];
  long long thr_rate;
  int T, reach, recovering_time, scale_bits, crit_active;
  int f64;          // off-lattice mode: `reach` is one more than on the lattice and the distance test is binary64
  double r;         // contagion_distance (abs.py:27), for that test
  int len_i, hei_i;  // x >= length  <=>  x >= ceil(length) for integer x (abs.py:177,182)
  int slab_lo, slab_hi, overflow, pull;
  unsigned k0, k1, step, n;
};

// The part of StepParams that no kernel ever changes (only cabs_set_params / cabs_set_slab / the seed do): the two
// agent passes take it BY VALUE as a kernel parameter, i.e. from the constant bank.  A value read from shared memory
// has to be re-read after every atomic (the compiler must assume other threads may have written it), and those
// re-reads were ~40 % of the load/store-unit traffic of the passes; constant-bank operands cost nothing there.
// Same field names as StepParams, so the per-agent device functions take either.
struct StaticParams {
  double min_expense, me32;
  double bi[5], mebi[5];
  long long thr_hosp[9], thr_sev[9], thr_death[9];
  int recovering_time, scale_bits;
  int len_i, hei_i;
  int slab_lo, slab_hi;
  unsigned k0, k1;
  double length, height;   // abs.py:177,182 compare binary64 positions with these in off-lattice mode
};

__host__ __device__ inline long long prob_bound_lt(double p);

__device__ __forceinline__ void load_params(StepParams &P, const Ctl *ctl, bool use_built = false) {
  // one thread fills the shared copy; callers __syncthreads() afterwards
  const Grid &g = use_built ? ctl->built : ctl->g;
  P.gx0 = g.gx0; P.gy0 = g.gy0; P.eshift = g.eshift; P.ncx = g.ncx; P.ncy = g.ncy;
  for (int i = 0; i < 4; ++i) P.amp[i] = ctl->amp[i];
  P.min_expense = ctl->min_expense;
  P.me32 = __dmul_rn(ctl->min_expense, 0x1p-32);
  for (int i = 0; i < 5; ++i) P.mebi[i] = __dmul_rn(ctl->min_expense, ctl->bi[i]);
  P.len_i = (int)fmin(fmax(ceil(ctl->length), -2147483647.0), 2147483647.0);
  P.hei_i = (int)fmin(fmax(ceil(ctl->height), -2147483647.0), 2147483647.0);
  for (int i = 0; i < 5; ++i) P.bi[i] = ctl->bi[i];
  for (int i = 0; i < 9; ++i) {
    P.thr_hosp[i] = ctl->thr_hosp[i]; P.thr_sev[i] = ctl->thr_sev[i]; P.thr_death[i] = ctl->thr_death[i];
  }
  P.thr_rate = ctl->thr_rate;
  P.T = ctl->T;
  // reach = floor(sqrt(T)): the largest |dx| a contact can have
  int reach = P.T > 0 ? (int)__dsqrt_rn((double)P.T) : 0;
  while ((long long)reach * reach > (long long)P.T) --reach;
  while ((long long)(reach + 1) * (reach + 1) <= (long long)P.T) ++reach;
  // off the lattice |xi - xj| <= r only bounds the floors by floor(r) + 1
  P.f64 = ctl->f64;
  P.r = ctl->contagion_distance;
  P.reach = reach + (ctl->f64 ? 1 : 0);
  P.recovering_time = ctl->recovering_time; P.scale_bits = ctl->scale_bits;
  P.crit_active = ctl->crit_active;
  P.slab_lo = ctl->slab_lo; P.slab_hi = ctl->slab_hi; P.overflow = ctl->overflow; P.pull = ctl->pull_mode;
  P.k0 = ctl->k0; P.k1 = ctl->k1; P.step = ctl->step; P.n = ctl->n;
}

// Simulation.move (abs.py:156-189): displacement of this step.  `w` = the two words of the move stream.
__device__ __forceinline__ void draw_displacement(const StepParams &P, uint32_t attr, const uint2 &w, int &ix,
                                                  int &iy) {
  const uint32_t st = attr_status(attr), sev = attr_severity(attr);
  const bool mobile = !(st == ST_D || (st == ST_I && (sev == SEV_H || sev == SEV_G)));  // abs.py:164-167
  ix = 0;
  iy = 0;
  if (mobile) {
    const float a = P.amp[st];
#ifdef CABS_EXACT_NORMALS
    displacement_exact(w.x, w.y, a, ix, iy);
#else
    // special-function evaluation with a per-draw proof that truncation agrees with the defining arithmetic;
    // the few draws that land within the error bound of an integer are redone exactly (abs_rng.cuh)
    if (!displacement_fast(w.x, w.y, a, ix, iy)) displacement_exact(w.x, w.y, a, ix, iy);
#endif
  }
}
// abs.py:177-185: the increment, not the position, is reflected
template <class SP>
__device__ __forceinline__ void apply_displacement(const SP &S, int x, int y, int ix, int iy, int &nx, int &ny) {
  const int tx = x + ix, ty = y + iy;
  nx = (tx <= 0 || tx >= S.len_i) ? x - ix : tx;
  ny = (ty <= 0 || ty >= S.hei_i) ? y - iy : ty;
}

// Cell of a lattice point.  Interior cells are 1..ncx-2 / 1..ncy-2; row/column 0 and the last one
// are always-empty padding so that the neighbourhood of a cell never needs a bounds test.
struct GridRegs {   // the cell grid of the build in flight, copied to registers by the agent passes
  int gx0, gy0, eshift, ncx, ncy;
};
template <class G>
__device__ __forceinline__ unsigned cell_key(const G &P, int x, int y, unsigned *err) {
  int cx = ((x - P.gx0) >> P.eshift) + 1;
  int cy = ((y - P.gy0) >> P.eshift) + 1;
  if (cx < 1 || cx > P.ncx - 2 || cy < 1 || cy > P.ncy - 2) {
    atomicOr(err, kErrOutOfGrid);
    cx = min(max(cx, 1), P.ncx - 2);
    cy = min(max(cy, 1), P.ncy - 2);
  }
  return (unsigned)cy * (unsigned)P.ncx + (unsigned)cx;
}


// ------------------------------------------------------------------------------------------
// Slab decomposition.  A device owns the lattice columns [slab_lo, slab_hi).  Agents whose
// move takes them out of the slab travel to the neighbour as complete post-update records (exchange 1,
// before the cell-list scan, so that the receiver can count them in its histogram); Infected agents within
// reach of a slab edge are sent as ghosts (exchange 2) and push their contacts on the neighbour's side.
// Exchange buffers (caller-allocated device memory): a 32-byte header {count} followed by the records.
constexpr unsigned kEmigrant = 0xFFFFFFFFu;  // aux rank of an agent that leaves the slab
struct alignas(32) MigrantRecord {
  uint4 hot;
  double w;
  unsigned long long pad;
};
struct alignas(32) ExchangeHeader {
  unsigned count;
  unsigned pad[7];
};
constexpr int kSlabRow = 16;  // int64 values per rank in the statistics all-gather

// Direct exchange over NVLink peer memory.  Every rank owns one "exchange block" (device memory that its
// neighbours map through CUDA IPC); a producer kernel's last block copies what its rank packed straight into
// the consumer's block with plain stores, fences at system scope and raises a flag that carries the build
// sequence number; the consumer kernel's blocks spin on the flag in their own memory.  Receive buffers and
// rows are double-buffered by the parity of the sequence number: a neighbour is never more than one build
// ahead (it needs this rank's next message to go further), so two generations are enough.  No collective call,
// no host involvement: a whole step is enqueued back to back.
constexpr int kMaxRanks = 16;
struct BlockLayout {
  size_t flags, rows, mig_recv, ghost_recv, mig_send, ghost_send, total, mig_bytes, ghost_bytes;
};
__host__ __device__ inline BlockLayout block_layout(unsigned n_ranks, unsigned mig_cap, unsigned ghost_cap) {
  BlockLayout L;
  L.mig_bytes = sizeof(ExchangeHeader) + (size_t)mig_cap * sizeof(MigrantRecord);
  L.ghost_bytes = sizeof(ExchangeHeader) + (size_t)ghost_cap * sizeof(uint4);
  L.flags = 0;                                              // unsigned[8 + 2 * kMaxRanks]
  L.rows = 256;                                             // [parity][rank][kSlabRow] int64
  L.mig_recv = L.rows + 2 * (size_t)kMaxRanks * kSlabRow * sizeof(long long);   // [from dir][parity]
  L.ghost_recv = L.mig_recv + 4 * L.mig_bytes;              // [from dir][parity]
  L.mig_send = L.ghost_recv + 4 * L.ghost_bytes;            // [to dir]   (local staging)
  L.ghost_send = L.mig_send + 2 * L.mig_bytes;              // [to dir]
  L.total = L.ghost_send + 2 * L.ghost_bytes;
  (void)n_ranks;
  return L;
}
struct SlabDirect {
  unsigned char *self, *left, *right;   // exchange blocks as mapped in this process; nullptr on an outer side
  unsigned char *const *all;            // device array [n_ranks]: every rank's block (rows are broadcast)
  unsigned seq;                         // sequence number of the build, RELATIVE to Ctl::build_seq (1 = the build in
                                        // flight), so that the same launch arguments serve every build (CUDA graphs)
  int enabled;
};
enum { FLAG_MIG = 0, FLAG_GHOST = 4, FLAG_ROW = 8 };   // + from_dir * 2 + parity   /   + parity * kMaxRanks + rank

constexpr uint32_t kErrPeerTimeout = 16u;   // a neighbour's message did not arrive (its process died or fell out of step)
// `err`: sticky error word; a wait gives up after ~30 s so that a lost peer ends in an error, not in a hung GPU
__device__ __forceinline__ void wait_flag(const unsigned char *block, int index, unsigned seq, unsigned *err = nullptr) {
  const volatile unsigned *f = reinterpret_cast<const volatile unsigned *>(block) + index;
  const long long t0 = clock64();
  while (*f != seq) {
    __nanosleep(64);
    if (clock64() - t0 > 60000000000ll) {
      if (err) atomicOr(err, kErrPeerTimeout);
      break;
    }
  }
  __threadfence_system();
}
// Last block of a producer kernel (all its threads): copy `local` (header + records) into `remote`, then publish.
__device__ __forceinline__ void push_buffer(const ExchangeHeader *local, unsigned char *remote_block, size_t remote_off,
                                            unsigned rec_bytes, unsigned cap, int flag_index, unsigned seq) {
  const unsigned count = min(__ldcg(&local->count), cap);
  const uint4 *src = reinterpret_cast<const uint4 *>(local + 1);
  uint4 *dst = reinterpret_cast<uint4 *>(remote_block + remote_off + sizeof(ExchangeHeader));
  const size_t n16 = (size_t)count * rec_bytes / 16;
  for (size_t k = threadIdx.x; k < n16; k += blockDim.x) dst[k] = __ldcg(src + k);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    reinterpret_cast<ExchangeHeader *>(remote_block + remote_off)->count = count;
    __threadfence_system();
    *(reinterpret_cast<volatile unsigned *>(remote_block) + flag_index) = seq;
  }
}

// ------------------------------------------------------------------------------------------
// Bulk asynchronous copies (TMA, cp.async.bulk) global -> shared, completion on an mbarrier.  The agent
// passes stream their input columns through a ring of shared-memory stages this way: the loads cost no
// registers and run several tiles ahead of the arithmetic.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// orders this thread's generic-proxy accesses to shared memory before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void mbar_fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Programmatic dependent launch (opt-in, CABS_PDL=1; measured slower than plain graph edges, see abs_capi.cu): the
// kernels of a step can be chained with cudaLaunchAttributeProgrammaticStreamSerialization, so that the blocks of the
// next kernel are resident and parked at pdl_wait() while the tail of the previous one drains.  pdl_wait()
// returns once every block of the preceding kernel has finished and its writes are visible, so nothing produced by
// it may be touched before; pdl_trigger() lets the successor's blocks be scheduled (they park at their own wait).
// Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr int kTile = 256;    // agents per tile = consumer threads per block of the agent passes
constexpr int kPad = kTile;   // every per-agent array is allocated with this many spare elements (whole-tile loads)
constexpr int kConsumerWarps = kTile / 32;
constexpr int kAgentThreads = kTile;       // threads per block of pass B

// ------------------------------------------------------------------------------------------
// Pass A: where will every agent be after this step's move?  Writes the displacement and the rank
// inside the destination cell (for pass B) and counts the agent in the histogram of that cell.
// kAdvance=false builds the cell list of the current positions (upload / contact-pair hook).
constexpr int kScanTile = 4096;      // histogram entries per scan tile (512 threads x 8)
constexpr int kScanThreads = 512;

// Direct slab mode, steady state: an agent whose move takes it out of the slab finishes its step right where pass A
// finds it and travels at once (k_pack_emigrants does the same from a list, in a launch of its own: two launches and
// ~25 us per step at 8 GPUs for ~16 000 agents).  Out of line, so that the registers of the update do not count
// against pass A's 32.  The critical-limit flag of abs.py:214 belongs to the statistics of the step before, whose global
// half (k_slab_close_global) has not run yet when pass A does: the few threads that get here read the ranks' rows
// themselves -- the same sum, the same comparison (close_global below).
__device__ __noinline__ void pack_emigrant_inline(Ctl *ctl, const StaticParams &SP, unsigned step, uint4 h, int nx, int ny,
                                                  int ix, int iy, double w, ExchangeHeader *send_left,
                                                  ExchangeHeader *send_right, const double *sqrt_tab, const SlabDirect &D,
                                                  int late);

// Pass A reads one 16-byte column with plenty of resident warps, so plain vectorised loads (one agent
// ahead) do as well as staged ones; pass B, which streams three columns at low occupancy, stages them.
#ifndef CABS_A_MINBLOCKS
#define CABS_A_MINBLOCKS 8   // 32 registers: pass A is latency-bound and loses 15 % at 6 resident blocks (measured)
#endif
// kPack (direct slab mode, steady state): emigrants are updated and sent from here (pack_emigrant_inline), the last
// block publishes the counts and raises the neighbours' flags -- no k_pack_emigrants launch.
template <bool kAdvance, bool kPack = false>
__global__ void __launch_bounds__(256, CABS_A_MINBLOCKS) k_move_count(Ctl *__restrict__ ctl, const uint4 *__restrict__ hot,
                                                    uint2 *__restrict__ aux, unsigned *__restrict__ cells,
                                                    unsigned long long *__restrict__ tile_state,
                                                    unsigned *__restrict__ emig_list,
                                                    const __grid_constant__ StaticParams SP,
                                                    const double *__restrict__ wealth = nullptr,
                                                    ExchangeHeader *__restrict__ send_left = nullptr,
                                                    ExchangeHeader *__restrict__ send_right = nullptr,
                                                    const double *__restrict__ sqrt_tab = nullptr,
                                                    const __grid_constant__ SlabDirect D = SlabDirect(), int late = 0) {
  __shared__ StepParams P;
  __shared__ unsigned s_last_a;
  pdl_wait();
  if (threadIdx.x == 0) load_params(P, ctl);
  __syncthreads();
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  // side job: reset the look-back state of the scan that follows
  {
    const unsigned nscan = ((unsigned)P.ncx * (unsigned)P.ncy + 1u + kScanTile - 1) / kScanTile;
    for (unsigned t = tid; t < nscan; t += stride) tile_state[t] = 0ull;
  }
  // the rank an atomic returns is stored one iteration later, so that its round trip to L2 overlaps
  // the next agent's arithmetic
  bool pending = false;
  unsigned p_i = 0, p_rank = 0;
  uint32_t p_d = 0;
  uint4 h_next = make_uint4(0, 0, 0, 0);
  const unsigned n_now = P.n, step_now = P.step;   // registers: a shared-memory value is re-read after every atomic
#ifdef CABS_GRID_IN_SMEM
  const StepParams &G = P;
#else
  const GridRegs G = {P.gx0, P.gy0, P.eshift, P.ncx, P.ncy};
#endif
#ifdef CABS_A_PREFETCH2
  // two records ahead in registers: the load queues behind the atomics of the iterations before it in the load/store
  // unit, and one iteration of arithmetic does not cover that
  uint4 h_next2 = make_uint4(0, 0, 0, 0);
  if (tid < n_now) h_next = __ldg(&hot[tid]);
  if (tid + stride < n_now) h_next2 = __ldg(&hot[tid + stride]);
  for (unsigned i = tid; i < n_now; i += stride) {
    const uint4 h = h_next;
    h_next = h_next2;
    if (i + 2 * stride < n_now) h_next2 = __ldg(&hot[i + 2 * stride]);
#else
  if (tid < n_now) h_next = __ldg(&hot[tid]);
  for (unsigned i = tid; i < n_now; i += stride) {
    const uint4 h = h_next;
    if (i + stride < n_now) h_next = __ldg(&hot[i + stride]);
#endif
    int nx = (int)h.x, ny = (int)h.y;
    uint32_t d = 0;
    if (kAdvance) {
      const uint32_t a = attr_normalize(h.z);
#ifdef CABS_DIAG_A_NODRAW     // timing diagnostics only: no Philox rounds, no Box-Muller
      const uint2 w = make_uint2(h.w * 0x9E3779B9u, h.w * 0x85EBCA6Bu + step_now);
      int ix = (int)(w.x >> 29) - 4, iy = (int)(w.y >> 29) - 4;
#else
      const uint2 w = agent_draw(h.w, step_now, kStreamMove, SP.k0, SP.k1);
      int ix, iy;
      draw_displacement(P, a, w, ix, iy);
#endif
      apply_displacement(SP, nx, ny, ix, iy, nx, ny);
      d = ((uint32_t)ix & 0xFFFFu) | ((uint32_t)iy << 16);
    }
    if (nx < SP.slab_lo || nx >= SP.slab_hi) {
      if (kPack) {   // leaves the slab: finishes its step here and goes straight into the neighbour's receive buffer
        atomicAdd(&ctl->n_emig, 1u);
        pack_emigrant_inline(ctl, SP, step_now, h, nx, ny, (int)(short)(d & 0xFFFFu), (int)(short)(d >> 16),
                             __ldg(&wealth[i]), send_left, send_right, sqrt_tab, D, late);
      } else {       // k_pack_emigrants finishes its step and hands it to the neighbour
        const unsigned e = atomicAdd(&ctl->n_emig, 1u);
        if (e < 2u * ctl->mig_cap) emig_list[e] = i;
        else atomicOr(&ctl->err, kErrExchange);
      }
      aux[i] = make_uint2(d, kEmigrant);
      continue;
    }
    const unsigned key = cell_key(G, nx, ny, &ctl->err);
#ifdef CABS_DIAG_A_NOSTORE
    if (pending && p_rank == 0xFFFFFFF0u) aux[p_i] = make_uint2(p_d, p_rank);
#else
    if (pending) aux[p_i] = make_uint2(p_d, p_rank);
#endif
#if defined(CABS_DIAG_A_NOATOMIC) && defined(CABS_DIAG_A_NOGATHER)
    p_rank = key >> 31;   // 0, but the compiler cannot know
#elif defined(CABS_DIAG_A_NOATOMIC)   // timing diagnostics only: a plain load in place of the returning atomic
    p_rank = __ldcg(&cells[key]);
#else
    p_rank = atomicAdd(&cells[key], 1u);
#endif
    p_d = d;
    p_i = i;
    pending = true;
  }
  pdl_trigger();
  if (pending) aux[p_i] = make_uint2(p_d, p_rank);
  if (kPack) {   // every block has fenced its remote stores; the last one publishes counts and flags (as k_pack_emigrants)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last_a = atomicAdd(&ctl->pack_ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last_a && threadIdx.x < 2) {
      __threadfence_system();
      const BlockLayout L = block_layout(ctl->n_ranks, ctl->mig_cap, ctl->ghost_cap);
      const unsigned dseq = ctl->build_seq + D.seq;
      const unsigned par = dseq & 1u;
      const int to_left = threadIdx.x == 0;
      unsigned char *peer = to_left ? D.left : D.right;
      if (peer) {
        const int from_dir = to_left ? 1 : 0;
        const unsigned cnt = min(__ldcg(&(to_left ? send_left : send_right)->count), ctl->mig_cap);
        reinterpret_cast<ExchangeHeader *>(peer + L.mig_recv + (from_dir * 2 + par) * L.mig_bytes)->count = cnt;
        __threadfence_system();
        *(reinterpret_cast<volatile unsigned *>(peer) + FLAG_MIG + from_dir * 2 + par) = dseq;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Exclusive scan of cells[0 .. ncells] in place, one pass: every tile publishes its sum, then its
// inclusive prefix, and looks back over the tiles before it (decoupled look-back).  Tiles are taken
// in ticket order so that every tile a block waits for is already running or done.
//   tile_state: bits 63..62 = 0 nothing, 1 tile sum, 2 inclusive prefix; bits 31..0 = value
constexpr unsigned long long kTileSum = 1ull << 62, kTilePrefix = 2ull << 62;

__global__ void __launch_bounds__(kScanThreads) k_scan_cells(Ctl *__restrict__ ctl, unsigned *__restrict__ cells,
                                                             unsigned long long *__restrict__ tile_state) {
  __shared__ unsigned warp_sums[kScanThreads / 32];
  __shared__ unsigned s_tile, s_prefix;
  pdl_wait();
  pdl_trigger();   // few blocks, short: let pass B's blocks take their places right away
  const unsigned m = ctl->g.ncells + 1;
  const unsigned ntiles = (m + kScanTile - 1) / kScanTile;
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(&ctl->scan_ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    if (tile >= ntiles) break;
    const unsigned base = tile * kScanTile + threadIdx.x * 8;
    unsigned v[8];
    if (base + 8 <= m) {
      const uint4 a = *reinterpret_cast<const uint4 *>(cells + base);
      const uint4 b = *reinterpret_cast<const uint4 *>(cells + base + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (base + k < m) ? cells[base + k] : 0u;
    }
    unsigned sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += v[k];
    unsigned inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      unsigned s = (lane < kScanThreads / 32) ? warp_sums[lane] : 0u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
      }
      if (lane < kScanThreads / 32) warp_sums[lane] = s;  // inclusive over warps
      const unsigned total = __shfl_sync(0xffffffffu, s, kScanThreads / 32 - 1);
      // publish, then look back (warp 0: lane l inspects tile - 1 - l of the current window)
      unsigned prefix = 0;
      if (tile == 0) {
        if (lane == 0) atomicExch(&tile_state[0], kTilePrefix | total);
      } else {
        if (lane == 0) atomicExch(&tile_state[tile], kTileSum | total);
        int look = (int)tile - 1;
        for (;;) {
          const int idx = look - (int)lane;
          unsigned long long st = kTilePrefix;  // tiles before 0 count as a zero prefix
          if (idx >= 0) {
            do {
              st = *reinterpret_cast<volatile unsigned long long *>(&tile_state[idx]);
            } while ((st >> 62) == 0ull);
          }
          const unsigned is_prefix = __ballot_sync(0xffffffffu, (st >> 62) == 2ull);
          const unsigned val = (unsigned)(st & 0xFFFFFFFFull);
          // sum the window up to and including the nearest tile that knows its inclusive prefix
          const int stop = is_prefix ? __ffs(is_prefix) - 1 : 31;
          unsigned contrib = ((int)lane <= stop) ? val : 0u;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
          prefix += contrib;
          if (is_prefix) break;
          look -= 32;
        }
        if (lane == 0) atomicExch(&tile_state[tile], kTilePrefix | (unsigned long long)(prefix + total));
      }
      if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    unsigned ex = s_prefix + (wid ? warp_sums[wid - 1] : 0u) + inc - sum;
    unsigned o8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      o8[k] = ex;
      ex += v[k];
    }
    if (base + 8 <= m) {
      *reinterpret_cast<uint4 *>(cells + base) = make_uint4(o8[0], o8[1], o8[2], o8[3]);
      *reinterpret_cast<uint4 *>(cells + base + 4) = make_uint4(o8[4], o8[5], o8[6], o8[7]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (base + k < m) cells[base + k] = o8[k];
    }
    __syncthreads();  // s_tile / warp_sums are reused by the next tile
  }
}

// ------------------------------------------------------------------------------------------
// Pass B: Simulation.move + Simulation.update (abs.py:156-230) per agent, partial statistics of
// abs.py:291-314, and the scatter of the record into its cell-sorted slot.
struct BlockStats {
  unsigned long long cnt[7];
  long long q[5];
  int bb[4];
};

#ifndef CABS_B_MINBLOCKS
#define CABS_B_MINBLOCKS 4
#endif
constexpr int kSqrtTable = 4096;  // sqrt(d2) for d2 < 4096, exact (IEEE sqrt), filled once by k_init_ctl

// One agent's move + update + economy (abs.py:156-230).  ix, iy: this step's displacement.
// S: the step-invariant parameters (StepParams or StaticParams); step, crit_active: what changes every step.
// sqrt(ix^2 + iy^2) of the move income (abs.py:187), exact: table for the displacements that occur, IEEE sqrt beyond
__device__ __forceinline__ double move_distance(int ix, int iy, const double *__restrict__ sqrt_tab) {
  const int d2 = ix * ix + iy * iy;
  return d2 < kSqrtTable ? __ldg(&sqrt_tab[d2]) : __dsqrt_rn((double)d2);
}

template <class SP>
__device__ __forceinline__ void advance_agent(const SP &S, uint32_t step, int crit_active, uint32_t aid, double dist,
                                              uint32_t &a, double &w, bool moved_by_rule = false) {
  const uint32_t st = attr_status(a);
  if (st == ST_D) return;   // the dead neither move (abs.py:164) nor update (abs.py:193)
  const uint32_t sev = attr_severity(a);
  const uint32_t stratum = min(attr_stratum(a), 4u);
  const uint2 r0 = agent_draw(aid, step, kStreamEconomy, S.k0, S.k1);
  if (!(st == ST_I && sev >= SEV_H) && !moved_by_rule) {   // mobile (abs.py:164-167); a move trigger returns before the income (169-172)
    // abs.py:187-189: wealth += sqrt(ix^2+iy^2) * rand * minimum_expense * basic_income[stratum], rand = (r + .5) 2^-32.
    // The power of two rides on minimum_expense (me32): scaling by it commutes with every rounding, so the three
    // products below round exactly like ((dist * rand) * minimum_expense) * basic_income.
    const double x = __dmul_rn(dist, __dadd_rn((double)r0.x, 0.5));
    w = __dadd_rn(w, __dmul_rn(__dmul_rn(x, S.me32), S.bi[stratum]));
  }
  // abs.py:191-230
  if (st == ST_I) {
    uint32_t nst = ST_I, nsev = sev, time = attr_time(a) + 1u;
    const uint32_t age = attr_age(a);
    const uint32_t idx = min(age > 10u ? age / 10u - 1u : 0u, 8u);  // abs.py:204
    const long long w_sub = (long long)r0.y;
    if (sev == SEV_A) {
      if (w_sub < S.thr_hosp[idx]) nsev = SEV_H;  // age_hospitalization_probs[idx] > u (abs.py:209)
    } else if (sev == SEV_H) {
      if (w_sub < S.thr_sev[idx]) {              // age_severe_probs[idx] > u (abs.py:212)
        nsev = SEV_G;
        if (crit_active) {                     // abs.py:214-217
          nst = ST_D;
          nsev = SEV_A;
        }
      }
    }
    const uint2 r1 = agent_draw(aid, step, kStreamDeath, S.k0, S.k1);
    bool pays = true;
    if ((long long)r1.x < S.thr_death[idx]) {    // age_death_probs[idx] > u (abs.py:219-223)
      nst = ST_D;
      nsev = SEV_A;
      pays = false;
    } else if ((int)time > S.recovering_time) {  // abs.py:225-228
      time = 0;
      nst = ST_R;
      nsev = SEV_A;
    }
    a = (a & 0xFFFF00F0u) | nst | (nsev << 2) | ((time & 0xFFu) << 8);
    if (!pays) return;                           // abs.py:223: the death of this step returns before the expense
  }
  w = __dadd_rn(w, -S.mebi[stratum]);            // abs.py:230
}

// An Infected agent within reach of a slab edge also pushes its contacts on the neighbour's side.
__device__ __forceinline__ void send_ghost(Ctl *ctl, const StepParams &P, const uint4 &rec,
                                           ExchangeHeader *__restrict__ ghost_left,
                                           ExchangeHeader *__restrict__ ghost_right) {
  const int x = (int)rec.x;
  if (P.slab_lo != INT_MIN && (long long)x - P.reach < (long long)P.slab_lo && P.T >= 0) {
    const unsigned g = atomicAdd(&ghost_left->count, 1u);
    if (g < ctl->ghost_cap) reinterpret_cast<uint4 *>(ghost_left + 1)[g] = rec;
    else atomicOr(&ctl->err, kErrExchange);
  }
  if (P.slab_hi != INT_MAX && (long long)x + P.reach >= (long long)P.slab_hi && P.T >= 0) {
    const unsigned g = atomicAdd(&ghost_right->count, 1u);
    if (g < ctl->ghost_cap) reinterpret_cast<uint4 *>(ghost_right + 1)[g] = rec;
    else atomicOr(&ctl->err, kErrExchange);
  }
}

#ifndef CABS_B_STAGES
#define CABS_B_STAGES 1   // measured 1 / 2 / 3 / 4 stages: 0.143 / 0.147 / 0.153 / 0.157 ms (profiles/r2_ab_carveout_stages.log)
#endif
constexpr int kStagesB = CABS_B_STAGES;
#ifndef CABS_B_INFSTAGE
#define CABS_B_INFSTAGE 1024
#endif
constexpr int kInfStage = CABS_B_INFSTAGE;   // Infected-list entries a block of pass B collects before it appends them in one piece
#ifndef CABS_B_GROUP
#define CABS_B_GROUP 4
#endif
constexpr unsigned kGroupB = CABS_B_GROUP;   // consecutive tiles a block takes before it strides on
struct alignas(128) StageB {
  uint4 hot[kTile];
  double w[kTile];
  uint2 aux[kTile];
};
// dynamic shared memory of pass B: the ring, the Q partial sums, the block's Infected list
constexpr size_t kPassBSmem = kStagesB * sizeof(StageB) + 5 * kTile * sizeof(long long) + kInfStage * sizeof(unsigned);
// acquire-release add on a shared-memory counter: the warp that brings a stage's count to kConsumerWarps has
// observed every other warp's release of that stage
__device__ __forceinline__ unsigned stage_release(unsigned *counter) {
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(counter)) : "memory");
  return old;
}

// kExtras = false: one device, synchronous schedule -- the instance the headline configuration runs.
// kExtras = true: additionally the slab paths (immigrants, ghosts, direct exchange) and the infection-time array of
// the sequential schedule; every extra is still guarded by its pointer.
// A block owns a CONTIGUOUS run of tiles: consecutive tiles land in neighbouring cells, so the cell-start
// gather of a tile mostly hits lines the block's previous tiles brought into L1, and the scattered record
// stores of a block stay within a narrow band of the destination copy.
// Input columns stream through a ring of kStagesB shared-memory stages filled by bulk copies (TMA, cp.async.bulk).
// There is no producer warp: a warp copies its 32 records of a stage to registers, fences (the refill is an
// async-proxy write, the reads are generic-proxy: without the fence a refill can overtake reads still queued in the
// shared-memory pipe) and counts itself off the stage; the LAST warp to do so issues the refill, so nobody ever waits
// for a stage to drain.
template <bool kAdvance, bool kExtras>
__global__ void __launch_bounds__(kTile, CABS_B_MINBLOCKS)
k_update_scatter(Ctl *__restrict__ ctl, const uint4 *__restrict__ hot, const double *__restrict__ wealth,
                 const uint2 *__restrict__ aux, const unsigned *__restrict__ cell_start,
                 unsigned *__restrict__ stale_cells, uint4 *__restrict__ ohot, double *__restrict__ owealth,
                 unsigned *__restrict__ inf_list, const double *__restrict__ sqrt_tab,
                 ExchangeHeader *__restrict__ ghost_left, ExchangeHeader *__restrict__ ghost_right,
                 unsigned long long *__restrict__ tinf, const ExchangeHeader *__restrict__ recv_left,
                 const ExchangeHeader *__restrict__ recv_right, const unsigned *__restrict__ imm_rank, SlabDirect D,
                 const __grid_constant__ StaticParams SP) {
  __shared__ StepParams P;
  __shared__ BlockStats B;
  __shared__ unsigned s_last;
  __shared__ unsigned s_stale;
  __shared__ uint64_t s_full[kStagesB];
  __shared__ unsigned s_released[kStagesB];
  __shared__ unsigned s_inf_n, s_inf_base, s_inf_valid;
  // dynamic shared memory (more than the 48 KB a static allocation may take): the staging ring, the per-thread
  // Q1..Q5 partial sums (abs.py:309-312; one column per thread: no bank conflicts) and the block's Infected list
  extern __shared__ __align__(128) unsigned char s_dyn[];
  StageB *s_in = reinterpret_cast<StageB *>(s_dyn);
  long long (*s_q)[kTile] = reinterpret_cast<long long (*)[kTile]>(s_dyn + kStagesB * sizeof(StageB));
  unsigned *s_inf = reinterpret_cast<unsigned *>(s_dyn + kStagesB * sizeof(StageB) + 5 * kTile * sizeof(long long));
#pragma unroll
  for (int k = 0; k < 5; ++k) s_q[k][threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    s_inf_n = 0;
    s_inf_valid = kInfStage;
  }
  pdl_wait();
  if (threadIdx.x == 0) {
    load_params(P, ctl);
    for (int i = 0; i < 7; ++i) B.cnt[i] = 0;
    for (int i = 0; i < 5; ++i) B.q[i] = 0;
    B.bb[0] = B.bb[2] = INT_MAX;
    B.bb[1] = B.bb[3] = INT_MIN;
    s_stale = ctl->cells_used[ctl->cur_table ^ 1];
    for (int k = 0; k < kStagesB; ++k) {
      mbar_init(&s_full[k], 1);
      s_released[k] = 0;
    }
    mbar_fence_init();
  }
  __syncthreads();
  const unsigned ntiles = (P.n + kTile - 1) / kTile;
  // k-th tile of this block: groups of kGroupB consecutive tiles, groups dealt round-robin over the blocks
  auto tile_of = [&](unsigned k) -> unsigned {
    return ((k / kGroupB) * gridDim.x + blockIdx.x) * kGroupB + (k % kGroupB);
  };
  constexpr unsigned kStageBytes = kTile * (sizeof(uint4) + sizeof(double) + sizeof(uint2));
  if (threadIdx.x == 0) {   // fill the ring
    for (unsigned k = 0; k < (unsigned)kStagesB && tile_of(k) < ntiles; ++k) {
      const size_t off = (size_t)tile_of(k) * kTile;
      mbar_expect_tx(&s_full[k], kStageBytes);
      bulk_load(s_in[k].hot, hot + off, kTile * sizeof(uint4), &s_full[k]);
      bulk_load(s_in[k].w, wealth + off, kTile * sizeof(double), &s_full[k]);
      bulk_load(s_in[k].aux, aux + off, kTile * sizeof(uint2), &s_full[k]);
    }
  }
  // side job: zero the cell table of the step before (the next step's histogram)
  {
    const unsigned m = s_stale, m4 = m >> 2;
    uint4 *c4 = reinterpret_cast<uint4 *>(stale_cells);
    const unsigned chunk = (m4 + gridDim.x - 1) / gridDim.x;
    const unsigned b0 = min(m4, blockIdx.x * chunk), b1 = min(m4, b0 + chunk);
    for (unsigned i = b0 + threadIdx.x; i < b1; i += kTile) c4[i] = make_uint4(0, 0, 0, 0);
    if (blockIdx.x == 0 && threadIdx.x < (m & 3u)) stale_cells[(m4 << 2) + threadIdx.x] = 0;
  }
  // per-thread partial statistics: 16-bit fields (a thread sees far fewer than 65535 agents per launch
  // for any population the 2^31 capacity allows: n / (gridDim * 256) < 2^31 / (148*8*256))
  unsigned long long c_st = 0, c_sev = 0;  // status S,I,R,D and severity A,H,G counts
  int bxmin = INT_MAX, bxmax = INT_MIN, bymin = INT_MAX, bymax = INT_MIN;
  const double qscale = (double)(1ull << SP.scale_bits);
  const unsigned lane = threadIdx.x & 31;
  const bool stand_down = kExtras && P.overflow;
  const unsigned n_now = P.n, step_now = P.step;      // registers: shared-memory values are re-read after every atomic
  const int crit_now = P.crit_active;
  const bool pull_now = P.pull != 0;
#ifdef CABS_GRID_IN_SMEM
  const StepParams &G = P;
#else
  const GridRegs G = {P.gx0, P.gy0, P.eshift, P.ncx, P.ncy};
#endif

  for (unsigned k = 0;; ++k) {
    const unsigned tile = tile_of(k);
    if (tile >= ntiles) break;
    const unsigned s = k % kStagesB;
    mbar_wait(&s_full[s], (k / kStagesB) & 1u);
    const uint4 h = s_in[s].hot[threadIdx.x];
    double w = s_in[s].w[threadIdx.x];
    const uint2 ax = s_in[s].aux[threadIdx.x];
    mbar_fence_async_shared();
    __syncwarp();
    if (lane == 0 && stage_release(&s_released[s]) == kConsumerWarps - 1) {
      // every warp has its records of this stage in registers: refill it with the tile kStagesB ahead
      s_released[s] = 0;
      const unsigned next = tile_of(k + kStagesB);
      if (next < ntiles) {
        const size_t off = (size_t)next * kTile;
        mbar_expect_tx(&s_full[s], kStageBytes);
        bulk_load(s_in[s].hot, hot + off, kTile * sizeof(uint4), &s_full[s]);
        bulk_load(s_in[s].w, wealth + off, kTile * sizeof(double), &s_full[s]);
        bulk_load(s_in[s].aux, aux + off, kTile * sizeof(uint2), &s_full[s]);
      }
    }
    const unsigned i = tile * kTile + threadIdx.x;
    bool valid = i < n_now;
    if (kExtras) valid = valid && ax.y != kEmigrant && !stand_down;
    int nx = (int)h.x, ny = (int)h.y;
    uint32_t a = attr_normalize(h.z);
    const uint32_t aid = h.w;
    const int ix = (int)(short)(ax.x & 0xFFFFu), iy = (int)(short)(ax.x >> 16);
    if (kAdvance) apply_displacement(SP, nx, ny, ix, iy, nx, ny);
    // the destination slot only needs the new position: ask for the cell start now so that the
    // round trip to the table overlaps the arithmetic below
    unsigned start = 0;
#ifndef CABS_DIAG_B_NOGATHER
    if (valid) start = __ldg(&cell_start[cell_key(G, nx, ny, &ctl->err)]);
#endif
    // the distance of the move income is a table lookup too: asked for here, consumed after the draw of the update
    const double dist = kAdvance ? move_distance(ix, iy, sqrt_tab) : 0.0;
#ifndef CABS_DIAG_B_NOCOMPUTE   // timing diagnostics only (scripts/ab_variants.sh): results are wrong with any CABS_DIAG_*
    if (kAdvance && valid) advance_agent(SP, step_now, crit_now, aid, dist, a, w);
#endif
    // partial statistics (abs.py:300-312); infections of this step are added by k_contagion
    const uint32_t st = attr_status(a), sev = attr_severity(a);
    // The sum is tied to the updated attribute word (bit 7 is always clear here, which the compiler cannot know) so
    // that it is scheduled after the update arithmetic: the whole of it then covers the table's round trip.
#ifdef CABS_DIAG_B_NOSCATTER
    const unsigned slot = i ^ ((start + ax.y) >> 31) ^ (a & kNewlyInfected);   // coalesced stores, same dependencies
#else
    const unsigned slot = (start + ax.y) ^ (a & kNewlyInfected);
#endif
    if (valid) {
      c_st += 1ull << (16 * st);
      if (st != ST_D) {
        c_sev += (sev != SEV_E) ? (1ull << (16 * (sev - 1))) : 0ull;
        if (attr_age(a) >= 18u) {
          s_q[min(attr_stratum(a), 4u)][threadIdx.x] += __double2ll_rn(__dmul_rn(w, qscale));
        }
      }
      bxmin = min(bxmin, nx); bxmax = max(bxmax, nx); bymin = min(bymin, ny); bymax = max(bymax, ny);
      ohot[slot] = make_uint4((uint32_t)nx, (uint32_t)ny, a, aid);
      owealth[slot] = w;
      if (kExtras && tinf) tinf[slot] = st == ST_I ? 0ull : ~0ull;   // sequential schedule: time of infection inside this step
    }
    // slots of the agents that can transmit in this step's contact phase (warp-aggregated append)
    const bool is_inf = valid && st == ST_I;
    if (kExtras && is_inf && ghost_left)
      send_ghost(ctl, P, make_uint4((uint32_t)nx, (uint32_t)ny, a, aid), ghost_left, ghost_right);
    // Every warp of every block appending to ONE global counter serialises in L2 (87 000 same-address atomics per
    // launch at 0.4 % Infected: the reservation came back a whole tile late and still stalled the warp).  A block
    // collects its entries in shared memory and appends them in one piece at its end; a block that finds more than
    // kInfStage of them (a dense epidemic) spills the rest straight to the global list.
    const unsigned m_inf = pull_now ? 0u : __ballot_sync(0xffffffffu, is_inf);   // pull mode needs no source list
    if (m_inf) {
      unsigned base = 0;
      const unsigned cnt = (unsigned)__popc(m_inf);
      if (lane == 0) base = atomicAdd(&s_inf_n, cnt);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + cnt <= (unsigned)kInfStage) {
        if (is_inf) s_inf[base + __popc(m_inf & ((1u << lane) - 1u))] = slot;
      } else {   // no room.  The counter only grows, so every later reservation starts beyond the capacity too: the
                 // stored entries are exactly [0, first overflowing base) -- remember that base, append globally
        unsigned gbase = 0;
        if (lane == 0) {
          atomicMin(&s_inf_valid, base);
          gbase = atomicAdd(&ctl->n_inf, cnt);
        }
        gbase = __shfl_sync(0xffffffffu, gbase, 0);
        if (is_inf) inf_list[gbase + __popc(m_inf & ((1u << lane) - 1u))] = slot;
      }
    }
  }
  pdl_trigger();
  __syncthreads();
  {   // append the block's list: one reservation, coalesced stores
    if (threadIdx.x == 0) {
      s_inf_n = min(s_inf_n, s_inf_valid);
      s_inf_base = s_inf_n ? atomicAdd(&ctl->n_inf, s_inf_n) : 0u;
    }
    __syncthreads();
    const unsigned cnt = s_inf_n, base = s_inf_base;
    for (unsigned k = threadIdx.x; k < cnt; k += kTile) inf_list[base + k] = s_inf[k];
  }
  // Slab mode: the agents that moved in from the neighbour slabs take their slots here too (they were counted in the
  // histogram before the scan): statistics, Infected list and ghosts exactly as for the residents -- every agent is
  // accounted once, by the device that owns it after the move.  A handful of records per block, hidden behind the tiles.
  if (kExtras && recv_left) {
    const unsigned cap = ctl->mig_cap;
    const unsigned nl = min(recv_left->count, cap), nr = min(recv_right->count, cap);
    const unsigned total = stand_down ? 0u : nl + nr;
    for (unsigned t = blockIdx.x * kTile + threadIdx.x; t < total; t += gridDim.x * kTile) {
      const MigrantRecord *src = t < nl ? reinterpret_cast<const MigrantRecord *>(recv_left + 1) + t
                                        : reinterpret_cast<const MigrantRecord *>(recv_right + 1) + (t - nl);
      const uint4 h = src->hot;
      const double w = src->w;
      const int nx = (int)h.x, ny = (int)h.y;
      const uint32_t a = h.z, st = attr_status(a), sev = attr_severity(a);
      const unsigned slot = cell_start[cell_key(P, nx, ny, &ctl->err)] + imm_rank[t];
      c_st += 1ull << (16 * st);
      if (st != ST_D) {
        c_sev += (sev != SEV_E) ? (1ull << (16 * (sev - 1))) : 0ull;
        if (attr_age(a) >= 18u) s_q[min(attr_stratum(a), 4u)][threadIdx.x] += __double2ll_rn(__dmul_rn(w, qscale));
      }
      bxmin = min(bxmin, nx); bxmax = max(bxmax, nx); bymin = min(bymin, ny); bymax = max(bymax, ny);
      ohot[slot] = h;
      owealth[slot] = w;
      if (st == ST_I) {
        if (!P.pull) inf_list[atomicAdd(&ctl->n_inf, 1u)] = slot;
        send_ghost(ctl, P, h, ghost_left, ghost_right);
      }
    }
  }
  // block reduction: per-warp shared atomics, then one set of global atomics per block
  {
    unsigned long long cnt[7];
    long long q[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) q[k] = s_q[k][threadIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k) cnt[k] = (c_st >> (16 * k)) & 0xFFFFull;
#pragma unroll
    for (int k = 0; k < 3; ++k) cnt[4 + k] = (c_sev >> (16 * k)) & 0xFFFFull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 7; ++k) cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], o);
#pragma unroll
      for (int k = 0; k < 5; ++k) q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
      bxmin = min(bxmin, __shfl_xor_sync(0xffffffffu, bxmin, o));
      bxmax = max(bxmax, __shfl_xor_sync(0xffffffffu, bxmax, o));
      bymin = min(bymin, __shfl_xor_sync(0xffffffffu, bymin, o));
      bymax = max(bymax, __shfl_xor_sync(0xffffffffu, bymax, o));
    }
    if (lane == 0) {
      for (int k = 0; k < 7; ++k) atomicAdd(&B.cnt[k], cnt[k]);
      for (int k = 0; k < 5; ++k) atomicAdd((unsigned long long *)&B.q[k], (unsigned long long)q[k]);
      atomicMin(&B.bb[0], bxmin); atomicMax(&B.bb[1], bxmax); atomicMin(&B.bb[2], bymin); atomicMax(&B.bb[3], bymax);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 7; ++k)
      if (B.cnt[k]) atomicAdd(&ctl->cnt[k], B.cnt[k]);
    for (int k = 0; k < 5; ++k)
      if (B.q[k]) atomicAdd((unsigned long long *)&ctl->q[k], (unsigned long long)B.q[k]);
    if (B.bb[0] <= B.bb[1]) {
      atomicMin(&ctl->bb_xmin, B.bb[0]); atomicMax(&ctl->bb_xmax, B.bb[1]);
      atomicMin(&ctl->bb_ymin, B.bb[2]); atomicMax(&ctl->bb_ymax, B.bb[3]);
    }
  }
  if (kExtras && D.enabled) {   // direct slab mode: every ghost is listed once all blocks are done; the last one hands them over
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->imm_ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      const BlockLayout L = block_layout(ctl->n_ranks, ctl->mig_cap, ctl->ghost_cap);
      const unsigned dseq = ctl->build_seq + D.seq;
      const unsigned par = dseq & 1u;
      // both directions at once: one round of stores, one system fence, then the two flags
      const unsigned nl_g = D.left ? min(__ldcg(&ghost_left->count), ctl->ghost_cap) : 0u;
      const unsigned nr_g = D.right ? min(__ldcg(&ghost_right->count), ctl->ghost_cap) : 0u;
      uint4 *dl = D.left ? reinterpret_cast<uint4 *>(D.left + L.ghost_recv + (1 * 2 + par) * L.ghost_bytes +
                                                     sizeof(ExchangeHeader)) : nullptr;
      uint4 *dr = D.right ? reinterpret_cast<uint4 *>(D.right + L.ghost_recv + (0 * 2 + par) * L.ghost_bytes +
                                                      sizeof(ExchangeHeader)) : nullptr;
      const uint4 *sl = reinterpret_cast<const uint4 *>(ghost_left + 1), *sr = reinterpret_cast<const uint4 *>(ghost_right + 1);
      for (unsigned k = threadIdx.x; k < nl_g + nr_g; k += blockDim.x) {
        if (k < nl_g) dl[k] = __ldcg(sl + k);
        else dr[k - nl_g] = __ldcg(sr + (k - nl_g));
      }
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x < 2) {
        unsigned char *peer = threadIdx.x == 0 ? D.left : D.right;
        if (peer) {
          const int from_dir = threadIdx.x == 0 ? 1 : 0;
          reinterpret_cast<ExchangeHeader *>(peer + L.ghost_recv + (from_dir * 2 + par) * L.ghost_bytes)->count =
              threadIdx.x == 0 ? nl_g : nr_g;
          __threadfence_system();
          *(reinterpret_cast<volatile unsigned *>(peer) + FLAG_GHOST + from_dir * 2 + par) = dseq;
        }
      }
    }
  }
}

__device__ __noinline__ void pack_emigrant_inline(Ctl *ctl, const StaticParams &SP, unsigned step, uint4 h, int nx, int ny,
                                                  int ix, int iy, double w, ExchangeHeader *send_left,
                                                  ExchangeHeader *send_right, const double *sqrt_tab, const SlabDirect &D,
                                                  int late) {
  const BlockLayout L = block_layout(ctl->n_ranks, ctl->mig_cap, ctl->ghost_cap);
  int crit = ctl->crit_active;
  if (late) {   // the statistics of the step before: every rank's row of the build that was closed locally (D.seq = 0 there)
    const unsigned cseq = ctl->build_seq, cpar = cseq & 1u;
    const long long *rows = reinterpret_cast<const long long *>(D.self + L.rows) + (size_t)cpar * kMaxRanks * kSlabRow;
    unsigned long long c5 = 0, c6 = 0;
    for (int g = 0; g < ctl->n_ranks; ++g) {
      wait_flag(D.self, FLAG_ROW + cpar * kMaxRanks + g, cseq, &ctl->err);
      c5 += (unsigned long long)__ldcg(rows + (size_t)g * kSlabRow + 5);
      c6 += (unsigned long long)__ldcg(rows + (size_t)g * kSlabRow + 6);
    }
    const double n = (double)ctl->pop_size;   // abs.py:215, as close_global evaluates it
    crit = ((double)c6 / n + (double)c5 / n) >= ctl->critical_limit;
  }
  uint32_t a = attr_normalize(h.z);
  advance_agent(SP, step, crit, h.w, move_distance(ix, iy, sqrt_tab), a, w, false);
  const bool to_left = nx < SP.slab_lo;
  ExchangeHeader *dst = to_left ? send_left : send_right;   // the local header counts; the record goes remote
  const unsigned k = atomicAdd(&dst->count, 1u);
  if (k >= ctl->mig_cap) {
    atomicOr(&ctl->err, kErrExchange);
    return;
  }
  MigrantRecord r;
  r.hot = make_uint4((uint32_t)nx, (uint32_t)ny, a, h.w);
  r.w = w;
  r.pad = 0;
  const unsigned par = (ctl->build_seq + D.seq) & 1u;
  unsigned char *peer = to_left ? D.left : D.right;
  MigrantRecord *out = reinterpret_cast<MigrantRecord *>(dst + 1);
  if (peer)   // what goes left arrives there "from the right" (dir 1)
    out = reinterpret_cast<MigrantRecord *>(peer + L.mig_recv + ((to_left ? 1 : 0) * 2 + par) * L.mig_bytes +
                                            sizeof(ExchangeHeader));
  out[k] = r;
}

// ------------------------------------------------------------------------------------------
// Slab exchange 1, sender side: the agents pass A found leaving the slab finish their step here
// (abs.py:156-230) and are appended, as complete records, to the buffer of the neighbour they move to.
template <bool kAdvance>
__global__ void __launch_bounds__(256)
k_pack_emigrants(Ctl *__restrict__ ctl, const uint4 *__restrict__ hot, const double *__restrict__ wealth,
                 const uint2 *__restrict__ aux, const unsigned *__restrict__ emig_list,
                 ExchangeHeader *__restrict__ send_left, ExchangeHeader *__restrict__ send_right,
                 const double *__restrict__ sqrt_tab, SlabDirect D) {
  __shared__ StepParams P;
  __shared__ unsigned s_last;
  pdl_wait();
  if (threadIdx.x == 0) load_params(P, ctl);
  __syncthreads();
  const unsigned count = min(ctl->n_emig, 2u * ctl->mig_cap);
  const BlockLayout L = block_layout(ctl->n_ranks, ctl->mig_cap, ctl->ghost_cap);
  const size_t remote_mig = L.mig_recv, mig_bytes = L.mig_bytes;
  const unsigned dseq = ctl->build_seq + D.seq;
  const unsigned par = dseq & 1u;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
    const unsigned i = emig_list[t];
    const uint4 h = hot[i];
    double w = wealth[i];
    const uint2 ax = aux[i];
    int nx = (int)h.x, ny = (int)h.y;
    uint32_t a = attr_normalize(h.z);
    const int ix = (int)(short)(ax.x & 0xFFFFu), iy = (int)(short)(ax.x >> 16);
    if (kAdvance) {
      apply_displacement(P, nx, ny, ix, iy, nx, ny);
      advance_agent(P, P.step, P.crit_active, h.w, move_distance(ix, iy, sqrt_tab), a, w);
    }
    const bool to_left = nx < P.slab_lo;
    ExchangeHeader *dst = to_left ? send_left : send_right;   // the local header counts; records may go remote
    const unsigned k = atomicAdd(&dst->count, 1u);
    if (k < ctl->mig_cap) {
      MigrantRecord r;
      r.hot = make_uint4((uint32_t)nx, (uint32_t)ny, a, h.w);
      r.w = w;
      r.pad = 0;
      MigrantRecord *out = reinterpret_cast<MigrantRecord *>(dst + 1);
      unsigned char *peer = D.enabled ? (to_left ? D.left : D.right) : nullptr;
      if (peer) {
        // straight into the neighbour's receive buffer: what goes left arrives there "from the right" (dir 1)
        out = reinterpret_cast<MigrantRecord *>(peer + remote_mig + ((to_left ? 1 : 0) * 2 + par) * mig_bytes +
                                                sizeof(ExchangeHeader));
      }
      out[k] = r;
    } else {
      atomicOr(&ctl->err, kErrExchange);
    }
  }
  if (D.enabled) {   // every block has fenced its remote stores; the last one publishes counts and flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ctl->pack_ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x < 2) {
      __threadfence_system();
      const int to_left = threadIdx.x == 0;
      unsigned char *peer = to_left ? D.left : D.right;
      if (peer) {
        const int from_dir = to_left ? 1 : 0;
        const unsigned cnt = min(__ldcg(&(to_left ? send_left : send_right)->count), ctl->mig_cap);
        reinterpret_cast<ExchangeHeader *>(peer + remote_mig + (from_dir * 2 + par) * mig_bytes)->count = cnt;
        __threadfence_system();
        *(reinterpret_cast<volatile unsigned *>(peer) + FLAG_MIG + from_dir * 2 + par) = dseq;
      }
    }
  }
}

// Slab exchange 1, receiver side, before the scan: count the immigrants in the histogram and keep
// their ranks; check that residents + immigrants fit.
// close_first (direct mode, steady state): block 0 first runs the global half of the previous step's closing
// (k_slab_close_global's work: every rank's row has long arrived) -- it sets the critical-limit flag pass B reads, and
// nothing this kernel uses.
__device__ void close_global(Ctl *ctl, const long long *rows, StatRecord *history, int record_stats);
__global__ void __launch_bounds__(256)
k_count_immigrants(Ctl *__restrict__ ctl, const ExchangeHeader *__restrict__ recv_left,
                   const ExchangeHeader *__restrict__ recv_right, unsigned *__restrict__ cells,
                   unsigned *__restrict__ imm_rank, SlabDirect D, int close_first, SlabDirect C,
                   StatRecord *__restrict__ history) {
  __shared__ StepParams P;
  pdl_wait();
  if (close_first && blockIdx.x == 0 && threadIdx.x == 32) {   // its own warp: thread 0 waits for the neighbours meanwhile
    const BlockLayout L = block_layout(ctl->n_ranks, ctl->mig_cap, ctl->ghost_cap);
    const unsigned cseq = ctl->build_seq + C.seq;
    const unsigned cpar = cseq & 1u;
    for (int g = 0; g < ctl->n_ranks; ++g) wait_flag(C.self, FLAG_ROW + cpar * kMaxRanks + g, cseq, &ctl->err);
    close_global(ctl, reinterpret_cast<const long long *>(C.self + L.rows) + (size_t)cpar * kMaxRanks * kSlabRow, history, 1);
  }
  if (threadIdx.x == 0) {
    load_params(P, ctl);
    if (D.enabled) {   // the neighbours' migrants of this build have landed in my block
      const unsigned dseq = ctl->build_seq + D.seq;
      const unsigned par = dseq & 1u;
      if (D.left) wait_flag(D.self, FLAG_MIG + 0 * 2 + par, dseq, &ctl->err);
      if (D.right) wait_flag(D.self, FLAG_MIG + 1 * 2 + par, dseq, &ctl->err);
    }
  }
  __syncthreads();
  const unsigned cap = ctl->mig_cap;
  const unsigned nl = min(recv_left->count, cap), nr = min(recv_right->count, cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (recv_left->count > cap || recv_right->count > cap) atomicOr(&ctl->err, kErrExchange);
    const unsigned long long total = (unsigned long long)P.n - min(ctl->n_emig, P.n) + nl + nr;
    if (total > ctl->capacity) {
      ctl->overflow = 1;
      atomicOr(&ctl->err, kErrExchange);
    }
  }
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < nl + nr; t += gridDim.x * blockDim.x) {
    const MigrantRecord *src = t < nl ? reinterpret_cast<const MigrantRecord *>(recv_left + 1) + t
                                      : reinterpret_cast<const MigrantRecord *>(recv_right + 1) + (t - nl);
    const uint4 h = src->hot;
    imm_rank[t] = atomicAdd(&cells[cell_key(P, (int)h.x, (int)h.y, &ctl->err)], 1u);
  }
}

// ------------------------------------------------------------------------------------------
// Cells a contact of the agent at (x, y) can be in: columns [clo, chi] of rows [rlo, rhi].  Only the
// cells that intersect [x - reach, x + reach] x [y - reach, y + reach] (reach = floor(sqrt(T)) <= cell
// edge); each row's columns are one contiguous slot range [cell_start[row*ncx+clo], cell_start[row*ncx+chi+1]).
__device__ __forceinline__ void contact_cells(const StepParams &P, int x, int y, int &clo, int &chi, int &rlo,
                                              int &rhi) {
  clo = min(max(((x - P.reach - P.gx0) >> P.eshift) + 1, 0), P.ncx - 1);
  chi = min(max(((x + P.reach - P.gx0) >> P.eshift) + 1, 0), P.ncx - 1);
  rlo = min(max(((y - P.reach - P.gy0) >> P.eshift) + 1, 0), P.ncy - 1);
  rhi = min(max(((y + P.reach - P.gy0) >> P.eshift) + 1, 0), P.ncy - 1);
}

// The reference's own predicate on binary64 positions (abs.py:9-10,258): sqrt((xi - xj)^2 + (yi - yj)^2) <= r, every
// operation rounded once (no FMA contraction), SURVEY.md A.3.
__device__ __forceinline__ bool within_f64(const double2 a, const double2 b, double r) {
  const double dx = __dadd_rn(a.x, -b.x), dy = __dadd_rn(a.y, -b.y);
  return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) <= r;
}

__device__ void close_step(Ctl *ctl, StatRecord *history, int record_stats);
__device__ void slab_export(Ctl *ctl, long long *rows, SlabDirect D);
__device__ void close_local(Ctl *ctl, const unsigned *cell_start, int record_stats);

// Pass C: the contact loop of abs.py:249-265 pushed from the agents that were Infected before this
// contact phase to the Susceptible agents in reach; contact() = abs.py:141-154 with one draw per pair,
// keyed on the pair so that the result does not depend on who scans whom.  Sources: the slots pass B
// listed, then (slab mode) the ghost records received from the two neighbour slabs.
//   finish = 1: the last block to finish closes the step (one device);
//   finish = 2: the last block exports this slab's partial statistics row (slab mode).
#ifndef CABS_SOURCE_LANES
#define CABS_SOURCE_LANES 8
#endif
constexpr unsigned kSourceLanes = CABS_SOURCE_LANES;
__global__ void __launch_bounds__(256)
k_contagion(Ctl *__restrict__ ctl, uint4 *__restrict__ hot, const unsigned *__restrict__ cell_start,
            const unsigned *__restrict__ inf_list, const ExchangeHeader *__restrict__ ghosts_left,
            const ExchangeHeader *__restrict__ ghosts_right, StatRecord *__restrict__ history,
            long long *__restrict__ rows, int finish, SlabDirect D, ExchangeHeader *__restrict__ send_reset,
            const double2 *__restrict__ pos) {
  __shared__ StepParams P;
  __shared__ unsigned s_new, s_local, s_left, s_right;
  pdl_wait();
  const bool ghosts_late = (finish & 16) != 0;   // bit 16: local sources first, then wait for the ghosts (one launch)
  if (threadIdx.x == 0) {
    load_params(P, ctl);
    if (D.enabled && !ghosts_late) {   // the neighbours' ghosts of this build have landed in my block
      const unsigned dseq = ctl->build_seq + D.seq;
      const unsigned par = dseq & 1u;
      if (D.left) wait_flag(D.self, FLAG_GHOST + 0 * 2 + par, dseq, &ctl->err);
      if (D.right) wait_flag(D.self, FLAG_GHOST + 1 * 2 + par, dseq, &ctl->err);
    }
    s_new = 0;
    s_local = (finish & 8) ? 0u : ctl->n_inf;   // bit 8: ghost sources only (the local ones ran in their own launch);
                                                // in pull mode pass B listed nobody, so this is 0
    s_left = (ghosts_left && !ghosts_late) ? min(ghosts_left->count, ctl->ghost_cap) : 0u;
    s_right = (ghosts_right && !ghosts_late) ? min(ghosts_right->count, ctl->ghost_cap) : 0u;
  }
  __syncthreads();
  unsigned newinf = 0;
  const unsigned stride = gridDim.x * blockDim.x;
  unsigned total = s_local + s_left + s_right;
  if (P.T >= 0 && P.pull && !(finish & 8) && *(const volatile unsigned long long *)&ctl->cnt[ST_S] > 0) {
    // pull: every Susceptible agent (in cell order: neighbouring threads share cells) looks for an agent that was
    // Infected before this phase and stops at the first contact that transmits
    // residents after this step's migration (slab mode): the end of the last cell, not the count the step began with
    const unsigned n_now = cell_start[ctl->g.ncells];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_now; i += stride) {
      const uint4 h = hot[i];
      if ((h.z & (3u | kNewlyInfected)) != ST_S) continue;
      const int xi = (int)h.x, yi = (int)h.y;
      int clo, chi, rlo, rhi;
      contact_cells(P, xi, yi, clo, chi, rlo, rhi);
      bool hit = false;
      for (int row = rlo; row <= rhi && !hit; ++row) {
        const unsigned base = (unsigned)row * (unsigned)P.ncx;
        const unsigned jb = cell_start[base + clo], je = cell_start[base + chi + 1];
        for (unsigned j = jb; j < je; ++j) {
          const uint2 cp = *reinterpret_cast<const uint2 *>(&hot[j]);
          const int ddx = (int)cp.x - xi, ddy = (int)cp.y - yi;
          if ((unsigned)(ddx + P.reach) > 2u * (unsigned)P.reach || (unsigned)(ddy + P.reach) > 2u * (unsigned)P.reach)
            continue;
          if (ddx * ddx + ddy * ddy > P.T) continue;
          const uint2 cq = *(reinterpret_cast<const uint2 *>(&hot[j]) + 1);
          if ((cq.x & 3u) != ST_I) continue;   // newly infected agents still read Susceptible: they do not transmit yet
          const uint4 r = philox4x32_10(min(h.w, cq.y), P.step, kStreamContact, max(h.w, cq.y), P.k0, P.k1);
          if ((long long)r.x <= P.thr_rate) {   // contagion_test <= contagion_rate (abs.py:150-153)
            hit = true;
            break;
          }
        }
      }
      if (hit) {   // a ghost source of the same launch may have marked this agent too: count it once
        const uint32_t old = atomicOr(reinterpret_cast<uint32_t *>(&hot[i]) + 2, kNewlyInfected);
        if (!(old & kNewlyInfected)) ++newinf;
      }
    }
  }
  if (P.T >= 0) {
    // kSourceLanes adjacent lanes share one source and split its candidates: the candidate records of a source are
    // a dependent chain of scattered loads when one thread walks them alone
    // (only while cells are coarse -- several candidates per lane; on the fine grid of a dense epidemic a source has
    // a handful of candidates and one lane each is the better split)
    const unsigned lanes = (P.n >= 4u * (unsigned)P.ncx * (unsigned)P.ncy) ? kSourceLanes : 1u;
    const unsigned sub = threadIdx.x % lanes;
    unsigned first = 0;
    for (int part = 0; part < (ghosts_late ? 2 : 1); ++part) {
    if (part == 1) {   // the local sources are done: now the neighbours' ghosts of this build (they have had that long to land)
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned dseq = ctl->build_seq + D.seq;
        const unsigned par = dseq & 1u;
        if (D.left) wait_flag(D.self, FLAG_GHOST + 0 * 2 + par, dseq, &ctl->err);
        if (D.right) wait_flag(D.self, FLAG_GHOST + 1 * 2 + par, dseq, &ctl->err);
        s_left = ghosts_left ? min(*(const volatile unsigned *)&ghosts_left->count, ctl->ghost_cap) : 0u;
        s_right = ghosts_right ? min(*(const volatile unsigned *)&ghosts_right->count, ctl->ghost_cap) : 0u;
      }
      __syncthreads();
      first = s_local;
      total = s_local + s_left + s_right;
    }
    for (unsigned t = first + (blockIdx.x * blockDim.x + threadIdx.x) / lanes; t < total; t += stride / lanes) {
      uint4 h;
      double2 pi = make_double2(0.0, 0.0);
      if (t < s_local) {
        h = hot[inf_list[t]];
        if (pos) pi = pos[inf_list[t]];
      } else if (t < s_local + s_left) h = reinterpret_cast<const uint4 *>(ghosts_left + 1)[t - s_local];
      else h = reinterpret_cast<const uint4 *>(ghosts_right + 1)[t - s_local - s_left];
      const int xi = (int)h.x, yi = (int)h.y;
      int clo, chi, rlo, rhi;
      contact_cells(P, xi, yi, clo, chi, rlo, rhi);
      for (int row = rlo; row <= rhi; ++row) {
        const unsigned base = (unsigned)row * (unsigned)P.ncx;
        const unsigned jb = cell_start[base + clo], je = cell_start[base + chi + 1];
        for (unsigned j = jb + sub; j < je; j += lanes) {
          // positions first (8 of the 16 bytes): almost every candidate of a coarse cell is out of reach
          const uint2 cp = *reinterpret_cast<const uint2 *>(&hot[j]);
          const int ddx = (int)cp.x - xi, ddy = (int)cp.y - yi;
          if ((unsigned)(ddx + P.reach) > 2u * (unsigned)P.reach || (unsigned)(ddy + P.reach) > 2u * (unsigned)P.reach)
            continue;
          if (pos) {                                      // off the lattice: the binary64 predicate on the exact positions
            if (!within_f64(pi, pos[j], P.r)) continue;
          } else if (ddx * ddx + ddy * ddy > P.T) continue;  // |ddx|, |ddy| <= reach <= 16000: no overflow
          const uint2 cq = *(reinterpret_cast<const uint2 *>(&hot[j]) + 1);
          const uint4 c = make_uint4(cp.x, cp.y, cq.x, cq.y);
          if ((c.z & (3u | kNewlyInfected)) != ST_S) continue;  // only a still Susceptible agent can be infected
          const uint4 r = philox4x32_10(min(h.w, c.w), P.step, kStreamContact, max(h.w, c.w), P.k0, P.k1);
          if ((long long)r.x <= P.thr_rate) {  // contagion_test <= contagion_rate (abs.py:150-153)
            // status bits stay Susceptible during this pass (see attr_normalize)
            const uint32_t old = atomicOr(reinterpret_cast<uint32_t *>(&hot[j]) + 2, kNewlyInfected);
            if (!(old & kNewlyInfected)) ++newinf;
          }
        }
      }
    }
    }
  }
  pdl_trigger();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) newinf += __shfl_xor_sync(0xffffffffu, newinf, o);
  if ((threadIdx.x & 31) == 0 && newinf) atomicAdd(&s_new, newinf);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_new) atomicAdd(&ctl->newinf, (unsigned long long)s_new);
    finish &= 7;
    if (finish) {
      // the last block to get here has seen every other block's contribution
      __threadfence();
      const unsigned ticket = atomicAdd(&ctl->done_ticket, 1u);
      if (ticket == gridDim.x - 1) {
        __threadfence();
        if (finish == 1) {
          close_step(ctl, history, 1);
        } else {
          slab_export(ctl, rows, D);
          if (finish == 3) {   // direct slab mode: the local half of the closing, nobody is waited for
            const BlockLayout L = block_layout(ctl->n_ranks, ctl->mig_cap, ctl->ghost_cap);
            unsigned char *self = reinterpret_cast<unsigned char *>(send_reset);
            for (int d = 0; d < 2; ++d) {
              reinterpret_cast<ExchangeHeader *>(self + L.mig_send + d * L.mig_bytes)->count = 0;
              reinterpret_cast<ExchangeHeader *>(self + L.ghost_send + d * L.ghost_bytes)->count = 0;
            }
            close_local(ctl, cell_start, 1);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Sequential schedule (abs.py:261-265 as written): the reference walks the pair list in lexicographic (i, j)
// order while statuses mutate, so an agent infected by an earlier pair transmits along later ones (SURVEY.md
// A.4).  With pair-keyed draws that is an earliest-arrival problem on a temporal graph in which the edge
// (i, j) exists exactly at time rank(i, j):  T(a) = min{ rank(p) : p = (a, b) passes its draw, T(b) < rank(p) },
// T = 0 for the agents Infected before the contact phase.  Earlier infection never blocks anyone, so
// label-correcting relaxation with atomicMin converges to the exact answer; the host repeats rounds until a round
// lowers no label.  Round 0 relaxes from the Infected list, later rounds from everyone infected in this step.
__device__ __forceinline__ unsigned long long pair_rank(uint32_t a, uint32_t b) {
  return (((unsigned long long)min(a, b) << 32) | (unsigned long long)max(a, b)) + 1ull;
}

__global__ void __launch_bounds__(256)
k_contagion_sequential(Ctl *__restrict__ ctl, uint4 *__restrict__ hot, const unsigned *__restrict__ cell_start,
                       unsigned long long *__restrict__ tinf, const unsigned *__restrict__ sources, int first_round,
                       unsigned *__restrict__ newinf_list, const double2 *__restrict__ pos) {
  __shared__ StepParams P;
  __shared__ unsigned s_count;
  if (threadIdx.x == 0) {
    load_params(P, ctl);
    s_count = first_round ? ctl->n_inf : ctl->n_newinf;
  }
  __syncthreads();
  if (P.T < 0) return;
  const unsigned count = s_count;  // snapshot: agents appended during this round are picked up by the next one
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
    const unsigned s = sources[t];
    const uint4 h = hot[s];
    const unsigned long long ts = *reinterpret_cast<volatile unsigned long long *>(&tinf[s]);
    const int xi = (int)h.x, yi = (int)h.y;
    int clo, chi, rlo, rhi;
    contact_cells(P, xi, yi, clo, chi, rlo, rhi);
    for (int row = rlo; row <= rhi; ++row) {
      const unsigned base = (unsigned)row * (unsigned)P.ncx;
      const unsigned jb = cell_start[base + clo], je = cell_start[base + chi + 1];
      for (unsigned j = jb; j < je; ++j) {
        const uint4 c = hot[j];
        if ((c.z & 3u) != ST_S || j == s) continue;   // Susceptible at the start of the phase (marked or not)
        const long long dx = (int)c.x - xi, dy = (int)c.y - yi;
        if (pos) {
          if (!within_f64(pos[s], pos[j], P.r)) continue;
        } else if (dx * dx + dy * dy > (long long)P.T) continue;
        const unsigned long long rank = pair_rank(h.w, c.w);
        if (rank <= ts) continue;                      // the pair was walked before the source was infected
        const uint4 r = philox4x32_10(min(h.w, c.w), P.step, kStreamContact, max(h.w, c.w), P.k0, P.k1);
        if ((long long)r.x > P.thr_rate) continue;     // contagion_test <= contagion_rate (abs.py:150-153)
        const unsigned long long old = atomicMin(&tinf[j], rank);
        if (rank < old) {
          atomicAdd(&ctl->seq_changed, 1u);
          if (old == ~0ull) {
            atomicOr(reinterpret_cast<uint32_t *>(&hot[j]) + 2, kNewlyInfected);
            atomicAdd(&ctl->newinf, 1ull);
            newinf_list[atomicAdd(&ctl->n_newinf, 1u)] = j;
          }
        }
      }
    }
  }
}

// The same relaxation with every round on the device: ONE cooperative launch per step.  All blocks are resident
// (the host sizes the grid with the occupancy calculator), a grid-wide barrier separates the rounds, and the
// "did this round lower a label" counters rotate over three slots so that clearing one never races with a block that
// is still adding to, or reading, another.  The last round's closing (statistics record, triggers, next grid) runs
// in block 0 after the final barrier, so a sequential step is four launches and has no host round trip.
//   seq_round[3]: labels lowered in round r are counted in slot r % 3.
__global__ void __launch_bounds__(256)
k_contagion_sequential_all(Ctl *__restrict__ ctl, uint4 *__restrict__ hot, const unsigned *__restrict__ cell_start,
                           unsigned long long *__restrict__ tinf, const unsigned *__restrict__ inf_list,
                           unsigned *__restrict__ newinf_list, StatRecord *__restrict__ history,
                           unsigned *__restrict__ seq_round, const double2 *__restrict__ pos) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  __shared__ StepParams P;
  __shared__ unsigned s_count;
  if (threadIdx.x == 0) load_params(P, ctl);
  __syncthreads();
  for (unsigned round = 0;; ++round) {
    unsigned *changed = seq_round + round % 3u;
    if (P.T >= 0) {
      if (threadIdx.x == 0) s_count = round == 0 ? ctl->n_inf : *(volatile unsigned *)&ctl->n_newinf;
      __syncthreads();
      const unsigned count = s_count;  // snapshot: agents appended during this round are picked up by the next one
      const unsigned *sources = round == 0 ? inf_list : newinf_list;
      const unsigned lanes = (P.n >= 4u * (unsigned)P.ncx * (unsigned)P.ncy) ? kSourceLanes : 1u;
      const unsigned sub = threadIdx.x % lanes;
      for (unsigned t = (blockIdx.x * blockDim.x + threadIdx.x) / lanes; t < count; t += gridDim.x * blockDim.x / lanes) {
        const unsigned s = __ldcg(&sources[t]);
        const uint4 h = __ldcg(&hot[s]);
        const unsigned long long ts = *reinterpret_cast<volatile unsigned long long *>(&tinf[s]);
        const int xi = (int)h.x, yi = (int)h.y;
        int clo, chi, rlo, rhi;
        contact_cells(P, xi, yi, clo, chi, rlo, rhi);
        for (int row = rlo; row <= rhi; ++row) {
          const unsigned base = (unsigned)row * (unsigned)P.ncx;
          const unsigned jb = cell_start[base + clo], je = cell_start[base + chi + 1];
          for (unsigned j = jb + sub; j < je; j += lanes) {
            const uint4 c = __ldcg(&hot[j]);
            if ((c.z & 3u) != ST_S || j == s) continue;   // Susceptible at the start of the phase (marked or not)
            const long long dx = (int)c.x - xi, dy = (int)c.y - yi;
            if (pos) {
              if (!within_f64(pos[s], pos[j], P.r)) continue;
            } else if (dx * dx + dy * dy > (long long)P.T) continue;
            const unsigned long long rank = pair_rank(h.w, c.w);
            if (rank <= ts) continue;                      // the pair was walked before the source was infected
            const uint4 r = philox4x32_10(min(h.w, c.w), P.step, kStreamContact, max(h.w, c.w), P.k0, P.k1);
            if ((long long)r.x > P.thr_rate) continue;     // contagion_test <= contagion_rate (abs.py:150-153)
            const unsigned long long old = atomicMin(&tinf[j], rank);
            if (rank < old) {
              atomicAdd(changed, 1u);
              if (old == ~0ull) {
                atomicOr(reinterpret_cast<uint32_t *>(&hot[j]) + 2, kNewlyInfected);
                atomicAdd(&ctl->newinf, 1ull);
                newinf_list[atomicAdd(&ctl->n_newinf, 1u)] = j;
              }
            }
          }
        }
      }
    }
    __threadfence();
    grid.sync();
    const unsigned lowered = *(volatile unsigned *)changed;
    if (blockIdx.x == 0 && threadIdx.x == 0) seq_round[(round + 2u) % 3u] = 0;   // next used two rounds from now
    if (lowered == 0 || P.T < 0) break;
  }
  // the loop ends on a round that lowered nothing, so all three slots are zero again for the next launch
  if (blockIdx.x == 0 && threadIdx.x == 0) close_step(ctl, history, 1);
}

// ------------------------------------------------------------------------------------------
// Test hook for the pair loop of abs.py:253-259: every (i, j), i < j, within contagion_distance.
__global__ void __launch_bounds__(256)
k_contact_pairs(Ctl *__restrict__ ctl, const uint4 *__restrict__ hot, const unsigned *__restrict__ cell_start,
                long long *__restrict__ out, unsigned long long capacity, unsigned long long *__restrict__ count,
                const double2 *__restrict__ pos) {
  __shared__ StepParams P;
  if (threadIdx.x == 0) load_params(P, ctl, true);
  __syncthreads();
  if (P.T < 0) return;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
    const uint4 h = hot[i];
    const int xi = (int)h.x, yi = (int)h.y;
    int clo, chi, rlo, rhi;
    contact_cells(P, xi, yi, clo, chi, rlo, rhi);
    for (int row = rlo; row <= rhi; ++row) {
      const unsigned base = (unsigned)row * (unsigned)P.ncx;
      const unsigned jb = cell_start[base + clo], je = cell_start[base + chi + 1];
      for (unsigned j = jb; j < je; ++j) {
        const uint4 c = hot[j];
        if (c.w <= h.w) continue;
        const long long dx = (int)c.x - xi, dy = (int)c.y - yi;
        if (pos) {
          if (!within_f64(pos[i], pos[j], P.r)) continue;
        } else if (dx * dx + dy * dy > (long long)P.T) continue;
        const unsigned long long slot = atomicAdd(count, 1ull);
        if (slot < capacity) {
          out[2 * slot] = (long long)h.w;
          out[2 * slot + 1] = (long long)c.w;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Grid planning: conservative bounding box of where agents can be after the next move.
__device__ inline void plan_grid(Ctl *ctl, int xmin, int xmax, int ymin, int ymax, bool add_margin) {
  if (xmin > xmax) { xmin = xmax = 0; ymin = ymax = 0; }  // empty population
  int margin = 0;
  if (add_margin) {
    float amax = 0.f;
    for (int k = 0; k < 3; ++k) amax = fmaxf(amax, fabsf(ctl->amp[k]));
    margin = (int)(kMaxAbsNormal * amax) + 2;
  }
  if (add_margin) margin = max(margin, ctl->rule_margin);
  if (add_margin && ctl->rule_box[0] <= ctl->rule_box[1]) {   // agents a `point` rule may teleport
    xmin = min(xmin, ctl->rule_box[0]); xmax = max(xmax, ctl->rule_box[1]);
    ymin = min(ymin, ctl->rule_box[2]); ymax = max(ymax, ctl->rule_box[3]);
  }
  long long x0 = (long long)xmin - margin, x1 = (long long)xmax + margin;
  const long long y0 = (long long)ymin - margin, y1 = (long long)ymax + margin;
  // a slab's residents and immigrants are inside [slab_lo, slab_hi) by construction
  if (ctl->slab_lo != INT_MIN) x0 = ctl->slab_lo;
  if (ctl->slab_hi != INT_MAX) x1 = (long long)ctl->slab_hi - 1;
  if (x1 < x0) x1 = x0;
  if (x1 - x0 >= (1ll << 28) || y1 - y0 >= (1ll << 28) || x0 < -(1ll << 30) || x1 > (1ll << 30) ||
      y0 < -(1ll << 30) || y1 > (1ll << 30)) {
    ctl->err |= kErrDomain;
  }
  // cell edge: power of two, at least floor(sqrt(T)) so contacts never skip a cell, grown until
  // the grid fits max_cells (keeps the cell table L2-resident)
  int eshift = 0;
  long long T = ctl->T;
  if (ctl->f64 && T >= 0) {   // floors of two positions within r differ by at most floor(r) + 1
    long long reach = (long long)__dsqrt_rn((double)T);
    while (reach * reach > T) --reach;
    while ((reach + 1) * (reach + 1) <= T) ++reach;
    T = (reach + 1) * (reach + 1);
  }
  while (T > 0 && ((1ll << eshift) * (1ll << eshift)) < T && eshift < 28) ++eshift;
  // (1<<eshift)^2 >= T  =>  1<<eshift >= sqrt(T) >= any |dx| of a contact
  for (;;) {
    const long long ncx = ((x1 - x0) >> eshift) + 3, ncy = ((y1 - y0) >> eshift) + 3;
    if (ncx * ncy <= (long long)ctl->cell_budget || eshift >= 28) {
      ctl->g.ncx = (int)ncx;
      ctl->g.ncy = (int)ncy;
      ctl->g.ncells = (unsigned)min(ncx * ncy, (long long)ctl->cell_budget);
      break;
    }
    ++eshift;
  }
  ctl->g.eshift = eshift;
  ctl->g.gx0 = (int)x0;
  ctl->g.gy0 = (int)y0;
}

__device__ inline void reset_accumulators(Ctl *ctl) {
  for (int k = 0; k < 7; ++k) ctl->cnt[k] = 0;
  for (int k = 0; k < 5; ++k) ctl->q[k] = 0;
  ctl->newinf = 0;
  ctl->bb_xmin = INT_MAX; ctl->bb_xmax = INT_MIN; ctl->bb_ymin = INT_MAX; ctl->bb_ymax = INT_MIN;
  ctl->n_inf = 0;
  ctl->n_emig = 0;
  ctl->pack_ticket = 0;
  ctl->imm_ticket = 0;
  ctl->n_newinf = 0;
  ctl->seq_changed = 0;
  ctl->scan_ticket = 0;
  ctl->done_ticket = 0;
}

__device__ inline double stat_value(const Ctl *ctl, const StatRecord &r, int metric) {
  if (metric < 7) return (double)r.cnt[metric] / (double)ctl->pop_size;
  return (double)r.q[metric - 7] / (double)(1ull << ctl->scale_bits);
}

__host__ __device__ inline long long prob_bound_lt(double p) {   // p > u  <=>  w < bound
  const double t = ceil(p * 4294967296.0 - 0.5);
  return t <= 0.0 ? 0ll : t >= 4294967296.0 ? 4294967296ll : (long long)t;
}
__device__ inline long long prob_bound_le(double r) {   // u <= r  <=>  w <= bound
  const double t = floor(r * 4294967296.0 - 0.5);
  return !(t >= 0.0) ? -1ll : t >= 4294967295.0 ? 4294967295ll : (long long)t;
}

__device__ inline int threshold_T(double r) {
  // largest integer n with sqrt(n) <= r in binary64; -1 if r < 0 or NaN
  if (!(r >= 0.0)) return -1;
  if (r > 16000.0) r = 16000.0;  // d2 must fit int32 for the largest supported box anyway
  long long n = (long long)floor(r * r);
  while (__dsqrt_rn((double)(n + 1)) <= r) ++n;
  while (n >= 0 && __dsqrt_rn((double)n) > r) --n;
  return (int)n;
}

// accumulators other blocks updated with atomics: read them from L2, not from a stale L1 line
__device__ __forceinline__ unsigned long long ld_acc(const unsigned long long *p) { return __ldcg(p); }
__device__ __forceinline__ long long ld_acc(const long long *p) { return __ldcg(p); }
__device__ __forceinline__ int ld_acc(const int *p) { return __ldcg(p); }

// End of Simulation.execute (abs.py:267-273) + the statistics row batch_experiment appends
// (experiments.py:193-196).  One thread.
//   record_stats = 1 : after a real step
//   record_stats = 0 : after upload / static rebuild (statistics of the initial state = "previous" stats)
// Dense epidemic = many agents that can transmit AND many that can be infected: the contact pass then pays for every
// candidate in the cells it visits, and a fine grid is worth its larger table.
// Direction of the contact pass: whoever is rarer does the scanning.  While the Infected are few, their list is
// pushed to the Susceptible in reach; once they outnumber the Susceptible, every Susceptible agent looks
// for an Infected one instead (and stops at the first success) -- in a saturated epidemic almost nobody scans.  Both
// directions evaluate the same pair draws, so the result is the same bit for bit; the choice uses the counts of the
// step just closed and only affects speed.
#ifndef CABS_PULL_X4
#define CABS_PULL_X4 4    // pull once Infected > (CABS_PULL_X4 / 4) x Susceptible (measured: 1, 4, 16 within 6 %)
#endif
__device__ inline void choose_cell_budget(Ctl *ctl, const StatRecord &cur) {
  const unsigned long long lesser = min(cur.cnt[ST_S], cur.cnt[ST_I]);
  ctl->cell_budget = (lesser * 16ull > ctl->pop_size && ctl->T >= 0) ? ctl->max_cells : ctl->coarse_cells;
  ctl->pull_mode = (ctl->schedule == 0 && !ctl->f64 &&
                    4ull * cur.cnt[ST_I] > (unsigned long long)CABS_PULL_X4 * cur.cnt[ST_S]) ? 1 : 0;
}

// Statistics half of the closing: the step's record, triggers on the memo of the step before, critical-limit flag.
__device__ void close_stats(Ctl *ctl, StatRecord *history, int record_stats) {
  StatRecord cur;
  const unsigned long long newinf = ld_acc(&ctl->newinf);
  for (int k = 0; k < 7; ++k) cur.cnt[k] = ld_acc(&ctl->cnt[k]);
  cur.cnt[ST_S] -= newinf;  // infections of the contact phase (abs.py:153)
  cur.cnt[ST_I] += newinf;
  for (int k = 0; k < 5; ++k) cur.q[k] = ld_acc(&ctl->q[k]);
  cur.new_infections = newinf;
  cur.n_resident = ctl->n;
  cur.pad[0] = cur.pad[1] = 0;
  if (record_stats) {
    history[ctl->step % ctl->hist_cap] = cur;
    // simulation triggers (abs.py:267-271) see the memoised statistics of the previous step
    for (int t = 0; t < ctl->ntrig; ++t) {
      const DevTrigger &tr = ctl->trig[t];
      const double v = stat_value(ctl, ctl->prev, tr.metric);
      const bool hit = tr.op == 0 ? v >= tr.threshold : tr.op == 1 ? v <= tr.threshold
                     : tr.op == 2 ? v > tr.threshold : v < tr.threshold;
      if (!hit) continue;
      ctl->trig_epoch += 1;
      if (tr.attribute == 0) {
        for (int k = 0; k < 4; ++k) ctl->amp[k] = (float)tr.value[k];
      } else if (tr.attribute == 1) {
        ctl->rate = tr.value[0];
        ctl->thr_rate = prob_bound_le(tr.value[0]);
      } else if (tr.attribute == 2) {
        ctl->contagion_distance = tr.value[0];
        ctl->T = threshold_T(tr.value[0]);
      } else if (tr.attribute == 3) {
        ctl->critical_limit = tr.value[0];
      }
    }
    ctl->step += 1;
  }
  ctl->prev = cur;
  ctl->cur = cur;
  // abs.py:215: statistics['Severe'] + statistics['Hospitalization'] >= critical_limit
  const double n = (double)ctl->pop_size;
  ctl->crit_active = ((double)cur.cnt[6] / n + (double)cur.cnt[5] / n) >= ctl->critical_limit;
  choose_cell_budget(ctl, cur);
}

__device__ void close_step(Ctl *ctl, StatRecord *history, int record_stats) {
  ctl->overflow = 0;
  close_stats(ctl, history, record_stats);
  // the table this build used stays valid (contact-pair hook) until the next build's pass B zeroes it
  ctl->cells_used[ctl->cur_table] = ctl->g.ncells + 1;
  ctl->cur_table ^= 1;
  ctl->built = ctl->g;
  const int bx0 = ld_acc(&ctl->bb_xmin), bx1 = ld_acc(&ctl->bb_xmax);
  const int by0 = ld_acc(&ctl->bb_ymin), by1 = ld_acc(&ctl->bb_ymax);
  ctl->last_bb[0] = bx0; ctl->last_bb[1] = bx1; ctl->last_bb[2] = by0; ctl->last_bb[3] = by1;
  plan_grid(ctl, bx0, bx1, by0, by1, true);
  reset_accumulators(ctl);
  __threadfence();
}

__global__ void k_finalize(Ctl *ctl, StatRecord *history, int record_stats) {
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  close_step(ctl, history, record_stats);
}

// Slab mode: a build ends with  k_slab_export -> all-gather of the rows -> k_slab_close  instead of the
// in-kernel closing: statistics are summed and the y-range is joined over the slabs, so every device
// takes the same trigger / critical-limit decisions and plans a grid that holds its future immigrants.
__device__ void slab_export(Ctl *ctl, long long *rows, SlabDirect D) {
  long long r[kSlabRow];
  for (int k = 0; k < 7; ++k) r[k] = (long long)ld_acc(&ctl->cnt[k]);
  for (int k = 0; k < 5; ++k) r[7 + k] = ld_acc(&ctl->q[k]);
  r[12] = (long long)ld_acc(&ctl->newinf);
  r[13] = ld_acc(&ctl->bb_ymin);
  r[14] = ld_acc(&ctl->bb_ymax);
  r[15] = (long long)atomicOr(&ctl->err, 0u);
  if (!D.enabled) {
    long long *dst = rows + (size_t)ctl->rank * kSlabRow;
    for (int k = 0; k < kSlabRow; ++k) dst[k] = r[k];
    return;
  }
  // broadcast this rank's row into every rank's block (its own included), then raise the flags
  const BlockLayout L = block_layout(ctl->n_ranks, ctl->mig_cap, ctl->ghost_cap);
  const unsigned dseq = ctl->build_seq + D.seq;
  const unsigned par = dseq & 1u;
  for (int g = 0; g < ctl->n_ranks; ++g) {
    long long *dst = reinterpret_cast<long long *>(D.all[g] + L.rows) + ((size_t)par * kMaxRanks + ctl->rank) * kSlabRow;
    for (int k = 0; k < kSlabRow; ++k) dst[k] = r[k];
  }
  __threadfence_system();
  for (int g = 0; g < ctl->n_ranks; ++g)
    *(reinterpret_cast<volatile unsigned *>(D.all[g]) + FLAG_ROW + par * kMaxRanks + ctl->rank) = dseq;
}

__global__ void k_slab_export(Ctl *ctl, long long *rows, SlabDirect D) {
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  slab_export(ctl, rows, D);
}

// Direct slab mode splits the closing so that no rank waits for the others at the end of a step:
//   close_local   right after the contact pass: triggers (they read the statistics of the step before, which are
//                 already global), step counter, table flip, next grid, accumulator reset.  The next grid needs
//                 the y-range of the whole population; the range of one build ago widened by one move is a
//                 superset of it, and results do not depend on the grid.
//   close_global  during the next build, after its pass A (which needs nothing global): waits for every rank's
//                 row, sums the statistics of the step, records them, sets the critical-limit flag pass B reads.
// Everything is identical to the one-device closing except for when it happens.
__device__ void close_local(Ctl *ctl, const unsigned *cell_start, int record_stats) {
  float amax = 0.f;
  for (int k = 0; k < 3; ++k) amax = fmaxf(amax, fabsf(ctl->amp[k]));
  const int widen = (int)(kMaxAbsNormal * amax) + 2;   // one move with the amplitudes of the step just done
  ctl->overflow = 0;
  if (record_stats) {
    for (int t = 0; t < ctl->ntrig; ++t) {             // abs.py:267-271, on the memo of the previous step
      const DevTrigger &tr = ctl->trig[t];
      const double v = stat_value(ctl, ctl->prev, tr.metric);
      const bool hit = tr.op == 0 ? v >= tr.threshold : tr.op == 1 ? v <= tr.threshold
                     : tr.op == 2 ? v > tr.threshold : v < tr.threshold;
      if (!hit) continue;
      ctl->trig_epoch += 1;
      if (tr.attribute == 0) {
        for (int k = 0; k < 4; ++k) ctl->amp[k] = (float)tr.value[k];
      } else if (tr.attribute == 1) {
        ctl->rate = tr.value[0];
        ctl->thr_rate = prob_bound_le(tr.value[0]);
      } else if (tr.attribute == 2) {
        ctl->contagion_distance = tr.value[0];
        ctl->T = threshold_T(tr.value[0]);
      } else if (tr.attribute == 3) {
        ctl->critical_limit = tr.value[0];
      }
    }
    ctl->pending_step = ctl->step;
    ctl->step += 1;
  }
  const unsigned n_next = cell_start[ctl->g.ncells];   // residents that stayed + immigrants
  ctl->cells_used[ctl->cur_table] = ctl->g.ncells + 1;
  ctl->cur_table ^= 1;
  ctl->built = ctl->g;
  int bx0 = ld_acc(&ctl->bb_xmin), bx1 = ld_acc(&ctl->bb_xmax);
  int by0 = ld_acc(&ctl->bb_ymin), by1 = ld_acc(&ctl->bb_ymax);
  if (ctl->glob_ymin <= ctl->glob_ymax) {
    by0 = min(by0, ctl->glob_ymin - widen);
    by1 = max(by1, ctl->glob_ymax + widen);
  }
  if (bx0 > bx1) bx0 = bx1 = ctl->slab_lo != INT_MIN ? ctl->slab_lo : ctl->slab_hi - 1;
  ctl->last_bb[0] = bx0; ctl->last_bb[1] = bx1; ctl->last_bb[2] = by0; ctl->last_bb[3] = by1;
  plan_grid(ctl, bx0, bx1, by0, by1, true);
  reset_accumulators(ctl);
  ctl->n = n_next;
  ctl->build_seq += 1;
  __threadfence();
}

__device__ void close_global(Ctl *ctl, const long long *rows, StatRecord *history, int record_stats) {
  StatRecord cur;
  for (int k = 0; k < 7; ++k) cur.cnt[k] = 0;
  for (int k = 0; k < 5; ++k) cur.q[k] = 0;
  unsigned long long newinf = 0;
  int ymin = INT_MAX, ymax = INT_MIN;
  for (int g = 0; g < ctl->n_ranks; ++g) {
    const long long *r = rows + (size_t)g * kSlabRow;
    for (int k = 0; k < 7; ++k) cur.cnt[k] += (unsigned long long)r[k];
    for (int k = 0; k < 5; ++k) cur.q[k] += r[7 + k];
    newinf += (unsigned long long)r[12];
    if (r[13] <= r[14]) {
      ymin = min(ymin, (int)r[13]);
      ymax = max(ymax, (int)r[14]);
    }
    if (r[15]) atomicOr(&ctl->err, (unsigned)r[15]);  // an exchange overflow anywhere invalidates the run everywhere
  }                                                   // (atomic: pass A of the next build may be running beside this)
  cur.cnt[ST_S] -= newinf;  // infections of the contact phase (abs.py:153)
  cur.cnt[ST_I] += newinf;
  cur.new_infections = newinf;
  cur.n_resident = ctl->n;
  cur.pad[0] = cur.pad[1] = 0;
  if (record_stats) history[ctl->pending_step % ctl->hist_cap] = cur;
  ctl->prev = cur;
  ctl->cur = cur;
  ctl->glob_ymin = ymin;
  ctl->glob_ymax = ymax;
  const double n = (double)ctl->pop_size;   // abs.py:215
  ctl->crit_active = ((double)cur.cnt[6] / n + (double)cur.cnt[5] / n) >= ctl->critical_limit;
  choose_cell_budget(ctl, cur);
}

__global__ void k_slab_close_local(Ctl *ctl, const unsigned *cell_start, int record_stats, ExchangeHeader *s0,
                                   ExchangeHeader *s1, ExchangeHeader *s2, ExchangeHeader *s3) {
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  s0->count = 0; s1->count = 0; s2->count = 0; s3->count = 0;
  close_local(ctl, cell_start, record_stats);
}

__global__ void k_slab_close_global(Ctl *ctl, StatRecord *history, int record_stats, SlabDirect D) {
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const BlockLayout L = block_layout(ctl->n_ranks, ctl->mig_cap, ctl->ghost_cap);
  const unsigned dseq = ctl->build_seq + D.seq;   // D.seq = 0: the build that has just been closed locally
  const unsigned par = dseq & 1u;
  for (int g = 0; g < ctl->n_ranks; ++g) wait_flag(D.self, FLAG_ROW + par * kMaxRanks + g, dseq, &ctl->err);
  close_global(ctl, reinterpret_cast<const long long *>(D.self + L.rows) + (size_t)par * kMaxRanks * kSlabRow, history,
               record_stats);
}

__global__ void k_slab_close(Ctl *ctl, const long long *rows, const unsigned *cell_start, StatRecord *history,
                             int record_stats, ExchangeHeader *s0, ExchangeHeader *s1, ExchangeHeader *s2,
                             ExchangeHeader *s3, SlabDirect D) {
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (D.enabled) {   // every rank's row of this build has landed in my block
    const BlockLayout L = block_layout(ctl->n_ranks, ctl->mig_cap, ctl->ghost_cap);
    const unsigned dseq = ctl->build_seq + D.seq;
    const unsigned par = dseq & 1u;
    for (int g = 0; g < ctl->n_ranks; ++g) wait_flag(D.self, FLAG_ROW + par * kMaxRanks + g, dseq, &ctl->err);
    rows = reinterpret_cast<const long long *>(D.self + L.rows) + (size_t)par * kMaxRanks * kSlabRow;
  }
  s0->count = 0; s1->count = 0; s2->count = 0; s3->count = 0;  // this slab's send buffers start the next build empty
  for (int k = 0; k < 7; ++k) ctl->cnt[k] = 0;
  for (int k = 0; k < 5; ++k) ctl->q[k] = 0;
  ctl->newinf = 0;
  int ymin = INT_MAX, ymax = INT_MIN;
  for (int g = 0; g < ctl->n_ranks; ++g) {
    const long long *r = rows + (size_t)g * kSlabRow;
    for (int k = 0; k < 7; ++k) ctl->cnt[k] += (unsigned long long)r[k];
    for (int k = 0; k < 5; ++k) ctl->q[k] += r[7 + k];
    ctl->newinf += (unsigned long long)r[12];
    if (r[13] <= r[14]) {
      ymin = min(ymin, (int)r[13]);
      ymax = max(ymax, (int)r[14]);
    }
    ctl->err |= (unsigned)r[15];  // an exchange overflow anywhere invalidates the run everywhere
  }
  if (ymin <= ymax) {
    ctl->bb_ymin = ymin;
    ctl->bb_ymax = ymax;
    if (ctl->bb_xmin > ctl->bb_xmax) {  // no resident here at the moment: any x inside the slab will do
      ctl->bb_xmin = ctl->bb_xmax = ctl->slab_lo != INT_MIN ? ctl->slab_lo : ctl->slab_hi - 1;
    }
  }
  const unsigned n_next = cell_start[ctl->g.ncells];  // residents that stayed + immigrants
  close_step(ctl, history, record_stats);
  ctl->n = n_next;
}

// ------------------------------------------------------------------------------------------
// Small populations (the reference's own sizes: 20 .. a few hundred agents, abs.py:17) are launch-latency bound
// when every step is four kernels.  One block holds the whole population in registers / shared memory and
// stays resident for a run of steps: move + update (the same device functions as passes A and B), contacts
// by scanning the list of Infected agents (at these sizes a cell list costs more than it saves), closing.
// Results are identical to the tiled path bit for bit: same draws, same arithmetic, integer statistics.
constexpr int kSmallMax = 1024;

__global__ void __launch_bounds__(kSmallMax)
k_small_run(Ctl *__restrict__ ctl, uint4 *__restrict__ hot, double *__restrict__ wealth,
            StatRecord *__restrict__ history, const double *__restrict__ sqrt_tab, int nsteps) {
  __shared__ StepParams P;
  __shared__ int s_x[kSmallMax], s_y[kSmallMax];
  __shared__ uint32_t s_id[kSmallMax];
  __shared__ unsigned s_inf[kSmallMax];
  __shared__ unsigned s_ninf, s_newinf;
  __shared__ unsigned long long s_cnt[7];
  __shared__ long long s_q[5];
  const unsigned i = threadIdx.x, n = ctl->n;
  const bool mine = i < n;
  uint4 h = mine ? hot[i] : make_uint4(0, 0, 0, 0);
  double w = mine ? wealth[i] : 0.0;
  for (int s = 0; s < nsteps; ++s) {
    if (i == 0) {
      load_params(P, ctl);
      s_ninf = 0;
      s_newinf = 0;
      for (int k = 0; k < 7; ++k) s_cnt[k] = 0;
      for (int k = 0; k < 5; ++k) s_q[k] = 0;
    }
    __syncthreads();
    uint32_t a = attr_normalize(h.z);
    int nx = (int)h.x, ny = (int)h.y;
    if (mine) {
      // Simulation.move + Simulation.update (abs.py:156-230)
      const uint2 r = agent_draw(h.w, P.step, kStreamMove, P.k0, P.k1);
      int ix, iy;
      draw_displacement(P, a, r, ix, iy);
      apply_displacement(P, nx, ny, ix, iy, nx, ny);
      advance_agent(P, P.step, P.crit_active, h.w, move_distance(ix, iy, sqrt_tab), a, w);
      // statistics (abs.py:300-312)
      const uint32_t st = attr_status(a), sev = attr_severity(a);
      atomicAdd(&s_cnt[st], 1ull);
      if (st != ST_D) {
        if (sev != SEV_E) atomicAdd(&s_cnt[3 + sev], 1ull);
        if (attr_age(a) >= 18u)
          atomicAdd((unsigned long long *)&s_q[min(attr_stratum(a), 4u)],
                    (unsigned long long)__double2ll_rn(__dmul_rn(w, (double)(1ull << P.scale_bits))));
      }
      s_x[i] = nx;
      s_y[i] = ny;
      s_id[i] = h.w;
      if (st == ST_I) s_inf[atomicAdd(&s_ninf, 1u)] = i;
    }
    __syncthreads();
    // contact() (abs.py:141-154, 249-265): a Susceptible agent looks at every agent Infected before this phase
    if (mine && attr_status(a) == ST_S && P.T >= 0) {
      const unsigned ninf = s_ninf;
      for (unsigned k = 0; k < ninf; ++k) {
        const unsigned j = s_inf[k];
        const long long dx = s_x[j] - nx, dy = s_y[j] - ny;
        if (dx * dx + dy * dy > (long long)P.T) continue;
        const uint32_t other = s_id[j];
        const uint4 r = philox4x32_10(min(h.w, other), P.step, kStreamContact, max(h.w, other), P.k0, P.k1);
        if ((long long)r.x <= P.thr_rate) {   // contagion_test <= contagion_rate (abs.py:150-153)
          a |= kNewlyInfected;
          atomicAdd(&s_newinf, 1u);
          break;
        }
      }
    }
    h = make_uint4((uint32_t)nx, (uint32_t)ny, a, h.w);
    __syncthreads();
    if (i == 0) {
      for (int k = 0; k < 7; ++k) ctl->cnt[k] = s_cnt[k];
      for (int k = 0; k < 5; ++k) ctl->q[k] = s_q[k];
      ctl->newinf = s_newinf;
      close_stats(ctl, history, 1);
      for (int k = 0; k < 7; ++k) ctl->cnt[k] = 0;
      for (int k = 0; k < 5; ++k) ctl->q[k] = 0;
      ctl->newinf = 0;
    }
    __syncthreads();
  }
  if (mine) {
    hot[i] = h;
    wealth[i] = w;
  }
}

// Bounding box of the resident positions into ctl->bb_* (used once per upload).
__global__ void __launch_bounds__(256) k_bbox(Ctl *__restrict__ ctl, const uint4 *__restrict__ hot) {
  int bxmin = INT_MAX, bxmax = INT_MIN, bymin = INT_MAX, bymax = INT_MIN;
  const unsigned n = ctl->n, stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 h = hot[i];
    bxmin = min(bxmin, (int)h.x); bxmax = max(bxmax, (int)h.x);
    bymin = min(bymin, (int)h.y); bymax = max(bymax, (int)h.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bxmin = min(bxmin, __shfl_xor_sync(0xffffffffu, bxmin, o));
    bxmax = max(bxmax, __shfl_xor_sync(0xffffffffu, bxmax, o));
    bymin = min(bymin, __shfl_xor_sync(0xffffffffu, bymin, o));
    bymax = max(bymax, __shfl_xor_sync(0xffffffffu, bymax, o));
  }
  if ((threadIdx.x & 31) == 0 && bxmin <= bxmax) {
    atomicMin(&ctl->bb_xmin, bxmin); atomicMax(&ctl->bb_xmax, bxmax);
    atomicMin(&ctl->bb_ymin, bymin); atomicMax(&ctl->bb_ymax, bymax);
  }
}

// Plan `g` from a bounding box: from_accum = the box k_bbox just reduced (exact, for a static
// build), otherwise the saved box of the current positions plus one move of margin.
__global__ void k_replan(Ctl *ctl, int from_accum) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ctl->T = threshold_T(ctl->contagion_distance);
  if (from_accum) {
    plan_grid(ctl, ctl->bb_xmin, ctl->bb_xmax, ctl->bb_ymin, ctl->bb_ymax, false);
  } else {
    plan_grid(ctl, ctl->last_bb[0], ctl->last_bb[1], ctl->last_bb[2], ctl->last_bb[3], true);
  }
  reset_accumulators(ctl);
}

// Zero the first `used` entries of a cell table (static rebuilds; the step path clears inside pass B).
__global__ void k_clear_table(Ctl *__restrict__ ctl, unsigned *__restrict__ cells, int which) {
  const unsigned m = ctl->cells_used[which];
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) cells[i] = 0;
}

// Host-side parameter block -> Ctl (cabs_set_params / cabs_set_triggers / creation).
struct ParamBlock {
  float amp[4];
  double length, height, contagion_distance, rate, critical_limit, min_expense;
  double bi[5];
  double p_hosp[9], p_sev[9], p_death[9];
  unsigned long long pop_size;
  int recovering_time, scale_bits, schedule;
};

__global__ void k_set_params(Ctl *ctl, ParamBlock p) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int k = 0; k < 4; ++k) ctl->amp[k] = p.amp[k];
  ctl->length = p.length; ctl->height = p.height; ctl->contagion_distance = p.contagion_distance;
  ctl->rate = p.rate; ctl->critical_limit = p.critical_limit; ctl->min_expense = p.min_expense;
  for (int k = 0; k < 5; ++k) ctl->bi[k] = p.bi[k];
  for (int k = 0; k < 9; ++k) {
    ctl->p_hosp[k] = p.p_hosp[k]; ctl->p_sev[k] = p.p_sev[k]; ctl->p_death[k] = p.p_death[k];
    ctl->thr_hosp[k] = prob_bound_lt(p.p_hosp[k]);
    ctl->thr_sev[k] = prob_bound_lt(p.p_sev[k]);
    ctl->thr_death[k] = prob_bound_lt(p.p_death[k]);
  }
  ctl->thr_rate = prob_bound_le(p.rate);
  ctl->pop_size = p.pop_size;
  ctl->recovering_time = p.recovering_time; ctl->scale_bits = p.scale_bits; ctl->schedule = p.schedule;
  ctl->T = threshold_T(p.contagion_distance);
  const double n = (double)ctl->pop_size;
  ctl->crit_active = ((double)ctl->cur.cnt[6] / n + (double)ctl->cur.cnt[5] / n) >= ctl->critical_limit;
}

struct TriggerBlock {
  int ntrig;
  DevTrigger trig[kMaxTriggers];
};

__global__ void k_set_triggers(Ctl *ctl, TriggerBlock t) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ctl->ntrig = t.ntrig;
  for (int k = 0; k < kMaxTriggers; ++k) ctl->trig[k] = t.trig[k];
}

__global__ void k_set_slab(Ctl *ctl, int rank, int n_ranks, int lo, int hi, unsigned mig_cap, unsigned ghost_cap) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ctl->rank = rank; ctl->n_ranks = n_ranks; ctl->slab_lo = lo; ctl->slab_hi = hi;
  ctl->mig_cap = mig_cap; ctl->ghost_cap = ghost_cap;
  ctl->build_seq = 0;
}

__global__ void k_init_ctl(Ctl *ctl, unsigned k0, unsigned k1, unsigned max_cells, unsigned coarse_cells,
                           unsigned hist_cap, unsigned capacity, double *sqrt_tab) {
  for (int d2 = threadIdx.x; d2 < kSqrtTable; d2 += blockDim.x) sqrt_tab[d2] = __dsqrt_rn((double)d2);
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ctl->k0 = k0; ctl->k1 = k1; ctl->max_cells = max_cells; ctl->hist_cap = hist_cap; ctl->capacity = capacity;
  ctl->coarse_cells = coarse_cells; ctl->cell_budget = coarse_cells; ctl->pull_mode = 0;
  ctl->step = 0; ctl->n = 0; ctl->err = 0; ctl->ntrig = 0; ctl->pop_size = 1; ctl->T = -1;
  ctl->cells_used[0] = ctl->cells_used[1] = 0;
  ctl->cur_table = 0;
  ctl->n_ranks = 1; ctl->rank = 0;
  ctl->slab_lo = INT_MIN; ctl->slab_hi = INT_MAX;
  ctl->overflow = 0; ctl->mig_cap = 0; ctl->ghost_cap = 0;
  ctl->pending_step = 0; ctl->glob_ymin = INT_MAX; ctl->glob_ymax = INT_MIN;
  ctl->f64 = 0; ctl->rule_margin = 0;
  ctl->rule_box[0] = ctl->rule_box[2] = 1; ctl->rule_box[1] = ctl->rule_box[3] = 0;   // empty
  reset_accumulators(ctl);
  plan_grid(ctl, 0, 0, 0, 0, false);
  ctl->built = ctl->g;
}

// ------------------------------------------------------------------------------------------
// MultiPopulationSimulation (abs.py:352-366): contacts between the agents of two populations whose boxes sit at fixed
// offsets of a common plane.  Every agent of A against every agent of B (the reference's double loop; B streams through
// shared memory), distance on the shifted integer positions against the same threshold T(r) as inside a population
// (SURVEY.md A.3).  `a_i.status == Susceptible and b_j.status == Infected` -> draw <= A's contagion_rate infects a_i,
// and the mirror image with B's rate (abs.py:365-366 call each population's own contact()).  Synchronous rule as in
// k_contagion: only agents Infected when the phase starts transmit, so the marks go to byte arrays first and
// k_cross_apply sets the newly-infected bit afterwards; the draw is keyed on (id in A, id in B, step, pair of
// populations), words x / y for the two directions.
constexpr uint32_t kStreamCross = 3;
__global__ void __launch_bounds__(256)
k_cross_contact(const Ctl *__restrict__ ctl_a, const uint4 *__restrict__ hot_a, const Ctl *__restrict__ ctl_b,
                const uint4 *__restrict__ hot_b, int shift_x, int shift_y, double distance, uint32_t step,
                uint32_t pair_code, uint32_t k0, uint32_t k1, unsigned char *__restrict__ mark_a,
                unsigned char *__restrict__ mark_b) {
  __shared__ uint4 s_b[256];
  __shared__ int s_T;
  if (threadIdx.x == 0) s_T = threshold_T(distance);
  const unsigned na = ctl_a->n, nb = ctl_b->n;
  const long long thr_a = ctl_a->thr_rate, thr_b = ctl_b->thr_rate;
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  const uint4 ha = i < na ? hot_a[i] : make_uint4(0, 0, 0, 0);
  const uint32_t st_a = attr_status(attr_normalize(ha.z));
  const long long xa = (long long)(int)ha.x + shift_x, ya = (long long)(int)ha.y + shift_y;   // in B's frame
  bool hit_a = false;
  for (unsigned base = 0; base < nb; base += 256u) {
    __syncthreads();
    if (base + threadIdx.x < nb) s_b[threadIdx.x] = hot_b[base + threadIdx.x];
    __syncthreads();
    const unsigned m = min(256u, nb - base);
    const long long T = s_T;
    if (i >= na || (st_a != ST_S && st_a != ST_I)) continue;
    for (unsigned j = 0; j < m; ++j) {
      const uint4 hb = s_b[j];
      const long long dx = xa - (int)hb.x, dy = ya - (int)hb.y;
      if (dx * dx + dy * dy > T) continue;
      const uint32_t st_b = attr_status(attr_normalize(hb.z));
      const bool a_from_b = st_a == ST_S && st_b == ST_I, b_from_a = st_b == ST_S && st_a == ST_I;
      if (!a_from_b && !b_from_a) continue;
      const uint4 r = philox4x32_10(ha.w, step, kStreamCross | (pair_code << 8), hb.w, k0, k1);
      if (a_from_b && (long long)r.x <= thr_a) hit_a = true;
      if (b_from_a && (long long)r.y <= thr_b) mark_b[base + j] = 1;   // several writers, one value
    }
  }
  if (hit_a) mark_a[i] = 1;
}

// Marked agents become Infected the way k_contagion's victims do (bit folded by attr_normalize at the next read).
__global__ void k_cross_apply(const Ctl *__restrict__ ctl, uint4 *__restrict__ hot, unsigned char *__restrict__ mark) {
  const unsigned n = ctl->n;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (mark[i]) {
      hot[i].z |= kNewlyInfected;
      mark[i] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Host SoA (separate columns, agents.py:44-64) <-> device records.
__global__ void k_pack(unsigned n, const int *__restrict__ x, const int *__restrict__ y,
                       const uint8_t *__restrict__ status, const uint8_t *__restrict__ severity,
                       const uint8_t *__restrict__ time, const uint8_t *__restrict__ age,
                       const uint8_t *__restrict__ stratum, const uint32_t *__restrict__ id_in,
                       uint4 *__restrict__ hot) {
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    hot[i] = make_uint4((uint32_t)x[i], (uint32_t)y[i], attr_pack(status[i], severity[i], stratum[i], time[i], age[i]),
                        id_in ? id_in[i] : i);
  }
}

// Uploaded ids of a one-device handle must be a permutation of 0..n-1: they key the draws, and cabs_download
// scatters by id.  `mark` holds n zeroed words.
__global__ void __launch_bounds__(256) k_check_ids(Ctl *__restrict__ ctl, const uint4 *__restrict__ hot, unsigned n,
                                                   unsigned *__restrict__ mark) {
  bool bad = false;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t id = hot[i].w;
    if (id >= n) bad = true;
    else if (atomicExch(&mark[id], 1u) != 0u) bad = true;
  }
  if (bad) atomicOr(&ctl->err, kErrBadIds);
}
__global__ void k_clear_error(Ctl *ctl, unsigned keep_mask) {
  if (threadIdx.x == 0 && blockIdx.x == 0) ctl->err &= keep_mask;
}

// Scatter back to id order (ids must be a permutation of 0..n-1 unless by_slot).
__global__ void k_unpack(unsigned n, int by_slot, const uint4 *__restrict__ hot, const double *__restrict__ wealth,
                         int *__restrict__ ox, int *__restrict__ oy, uint8_t *__restrict__ status,
                         uint8_t *__restrict__ severity, uint8_t *__restrict__ time, uint8_t *__restrict__ age,
                         uint8_t *__restrict__ stratum, double *__restrict__ owealth, uint32_t *__restrict__ oid) {
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 h = hot[i];
    const uint32_t a = attr_normalize(h.z);
    const unsigned d = by_slot ? i : h.w;
    if (d >= n) continue;   // cannot happen after k_check_ids; never write outside the buffers
    ox[d] = (int)h.x;
    oy[d] = (int)h.y;
    status[d] = (uint8_t)attr_status(a);
    severity[d] = (uint8_t)attr_severity(a);
    time[d] = (uint8_t)attr_time(a);
    age[d] = (uint8_t)attr_age(a);
    stratum[d] = (uint8_t)attr_stratum(a);
    owealth[d] = wealth[i];
    oid[d] = h.w;
  }
}


// ==========================================================================================================
// Off-lattice mode: binary64 positions and declarative per-agent rules.
//
// The reference keeps agents on the integer lattice only as long as nothing assigns them a float: a population
// trigger with attribute 'move' (abs.py:86-95, applied at abs.py:169-172) replaces the random walk of an agent by
// whatever position its action returns -- the shipped idiom is "some point + np.random.normal(0, sigma)" (run_batch.py:
// 446-450) -- and from then on x, y are binary64: the walk adds integer increments to them (abs.py:177-185), and the
// contact test is sqrt((xi-xj)^2 + (yi-yj)^2) <= contagion_distance in binary64 (abs.py:9-10,258).  Other population
// triggers rewrite an attribute after update() (abs.py:244-247).
//
// A Python lambda cannot run here, so rules are declarative (DevRule): a condition tabulated over the discrete
// attributes (status x severity x age x stratum, 8 000 bits -- any predicate of those four is representable) and
//   move rules       new position = target + N(0, sigma) per axis, target = the agent's own position or a fixed point;
//                    the first matching rule wins and the agent earns no move income (abs.py:169-172 returns early);
//   attribute rules  status / severity / infected_time / age / stratum := value, or wealth := p * wealth + q,
//                    applied in order after the update of the step.
// Layout: pos[2] (double2, ping/pong with hot) holds the exact positions, hot.x / hot.y their floors -- the cell list,
// the bounding box and the candidate pre-filter stay integer work; only the final distance test reads pos.  These
// kernels are the plain grid-stride form of passes A and B (no staging ring): the mode exists for drop-in coverage
// of the reference's API, the lattice path is the one that is tuned.
constexpr int kRuleWords = 250;   // 8 000 condition bits
constexpr int kMaxRules = 8;
enum { RULE_MOVE_STAY = 1, RULE_MOVE_POINT = 2, RULE_ATTRIBUTE = 3 };
enum { RATTR_STATUS = 0, RATTR_SEVERITY = 1, RATTR_TIME = 2, RATTR_AGE = 3, RATTR_STRATUM = 4, RATTR_WEALTH = 5 };
struct DevRule {
  int kind, attribute;
  double tx, ty, sigma;    // move rules
  double p, q;             // attribute rules: discrete attributes take (int)q, wealth becomes p * wealth + q
  uint32_t cond[kRuleWords];
  uint32_t pad[2];
};
__device__ __forceinline__ bool rule_matches(const DevRule &r, uint32_t a) {
  const uint32_t idx = ((attr_status(a) * 4u + attr_severity(a)) * 100u + min(attr_age(a), 99u)) * 5u +
                       min(attr_stratum(a), 4u);
  return (__ldg(&r.cond[idx >> 5]) >> (idx & 31u)) & 1u;
}
constexpr unsigned kRuleMoved = 0x80000000u;   // aux.y bit: this step's position came from a move rule

__device__ __forceinline__ bool floor_to_lattice(double v, int &out) {
  if (!(v > -268435456.0 && v < 268435456.0)) return false;   // also NaN
  out = __double2int_rd(v);
  return true;
}

template <bool kAdvance>
__global__ void __launch_bounds__(256)
k_move_count_f64(Ctl *__restrict__ ctl, const uint4 *__restrict__ hot, const double2 *__restrict__ pos,
                 double2 *__restrict__ npos, uint2 *__restrict__ aux, unsigned *__restrict__ cells,
                 unsigned long long *__restrict__ tile_state, const DevRule *__restrict__ rules, int n_move_rules,
                 const __grid_constant__ StaticParams SP) {
  __shared__ StepParams P;
  if (threadIdx.x == 0) load_params(P, ctl);
  __syncthreads();
  const unsigned stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
  {
    const unsigned nscan = ((unsigned)P.ncx * (unsigned)P.ncy + 1u + kScanTile - 1) / kScanTile;
    for (unsigned t = tid; t < nscan; t += stride) tile_state[t] = 0ull;
  }
  const unsigned n_now = P.n, step_now = P.step;
  const GridRegs G = {P.gx0, P.gy0, P.eshift, P.ncx, P.ncy};
  for (unsigned i = tid; i < n_now; i += stride) {
    const uint4 h = hot[i];
    const double2 p = pos[i];
    double nx = p.x, ny = p.y;
    uint32_t d = 0, flag = 0;
    if (kAdvance) {
      const uint32_t a = attr_normalize(h.z);
      const uint32_t st = attr_status(a), sev = attr_severity(a);
      const bool mobile = !(st == ST_D || (st == ST_I && (sev == SEV_H || sev == SEV_G)));   // abs.py:164-167
      if (mobile) {
        const uint2 w = agent_draw(h.w, step_now, kStreamMove, SP.k0, SP.k1);
        int hit = -1;
        for (int k = 0; k < n_move_rules && hit < 0; ++k)
          if (rule_matches(rules[k], a)) hit = k;                                             // abs.py:169-172
        if (hit >= 0) {
          float z0, z1;
          normal_pair(w.x, w.y, z0, z1);
          const bool stay = rules[hit].kind == RULE_MOVE_STAY;
          const double sigma = rules[hit].sigma;
          nx = __dadd_rn(stay ? p.x : rules[hit].tx, __dmul_rn((double)z0, sigma));
          ny = __dadd_rn(stay ? p.y : rules[hit].ty, __dmul_rn((double)z1, sigma));
          flag = kRuleMoved;
        } else {
          int ix, iy;
          draw_displacement(P, a, w, ix, iy);
          // abs.py:177-185 on binary64: the increment, not the position, is reflected
          const double tx = __dadd_rn(p.x, (double)ix), ty = __dadd_rn(p.y, (double)iy);
          nx = (tx <= 0.0 || tx >= SP.length) ? __dadd_rn(p.x, -(double)ix) : tx;
          ny = (ty <= 0.0 || ty >= SP.height) ? __dadd_rn(p.y, -(double)iy) : ty;
          d = ((uint32_t)ix & 0xFFFFu) | ((uint32_t)iy << 16);
        }
      }
    }
    npos[i] = make_double2(nx, ny);
    int fx = 0, fy = 0;
    if (!floor_to_lattice(nx, fx) || !floor_to_lattice(ny, fy)) atomicOr(&ctl->err, kErrDomain);
    const unsigned key = cell_key(G, fx, fy, &ctl->err);
    aux[i] = make_uint2(d, atomicAdd(&cells[key], 1u) | flag);
  }
}

template <bool kAdvance>
__global__ void __launch_bounds__(256)
k_update_scatter_f64(Ctl *__restrict__ ctl, const uint4 *__restrict__ hot, const double *__restrict__ wealth,
                     const uint2 *__restrict__ aux, const double2 *__restrict__ npos,
                     const unsigned *__restrict__ cell_start, unsigned *__restrict__ stale_cells,
                     uint4 *__restrict__ ohot, double *__restrict__ owealth, double2 *__restrict__ opos,
                     unsigned *__restrict__ inf_list, const double *__restrict__ sqrt_tab,
                     unsigned long long *__restrict__ tinf, const DevRule *__restrict__ rules, int n_move_rules,
                     int n_attr_rules, const __grid_constant__ StaticParams SP) {
  __shared__ StepParams P;
  __shared__ BlockStats B;
  __shared__ long long s_q[5][256];
  __shared__ unsigned s_stale;
#pragma unroll
  for (int k = 0; k < 5; ++k) s_q[k][threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    load_params(P, ctl);
    for (int i = 0; i < 7; ++i) B.cnt[i] = 0;
    for (int i = 0; i < 5; ++i) B.q[i] = 0;
    B.bb[0] = B.bb[2] = INT_MAX;
    B.bb[1] = B.bb[3] = INT_MIN;
    s_stale = ctl->cells_used[ctl->cur_table ^ 1];
  }
  __syncthreads();
  {   // zero the cell table of the step before (the next step's histogram)
    const unsigned m = s_stale;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) stale_cells[i] = 0;
  }
  unsigned long long c_st = 0, c_sev = 0;
  int bxmin = INT_MAX, bxmax = INT_MIN, bymin = INT_MAX, bymax = INT_MIN;
  const double qscale = (double)(1ull << SP.scale_bits);
  const unsigned lane = threadIdx.x & 31;
  const unsigned n_now = P.n, step_now = P.step;
  const int crit_now = P.crit_active;
  const GridRegs G = {P.gx0, P.gy0, P.eshift, P.ncx, P.ncy};
  // whole warps iterate together (the Infected-list append below is warp-aggregated)
  const unsigned n_round = (n_now + 31u) & ~31u;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    const bool valid = i < n_now;
    uint4 h = make_uint4(0, 0, 0, 0);
    double w = 0.0;
    uint2 ax = make_uint2(0, 0);
    double2 np = make_double2(0.0, 0.0);
    if (valid) { h = hot[i]; w = wealth[i]; ax = aux[i]; np = npos[i]; }
    uint32_t a = attr_normalize(h.z);
    const int ix = (int)(short)(ax.x & 0xFFFFu), iy = (int)(short)(ax.x >> 16);
    int fx = 0, fy = 0;
    floor_to_lattice(np.x, fx);
    floor_to_lattice(np.y, fy);
    unsigned slot = 0;
    if (valid) slot = cell_start[cell_key(G, fx, fy, &ctl->err)] + (ax.y & ~kRuleMoved);
    if (kAdvance && valid) {
      advance_agent(SP, step_now, crit_now, h.w, move_distance(ix, iy, sqrt_tab), a, w, (ax.y & kRuleMoved) != 0);
      // population triggers on other attributes (abs.py:244-247), in order, on the updated agent
      for (int k = n_move_rules; k < n_move_rules + n_attr_rules; ++k) {
        const DevRule &r = rules[k];
        if (!rule_matches(r, a)) continue;
        const uint32_t v = (uint32_t)(int)r.q;
        switch (r.attribute) {
          case RATTR_STATUS: a = (a & ~3u) | (v & 3u); break;
          case RATTR_SEVERITY: a = (a & ~0xCu) | ((v & 3u) << 2); break;
          case RATTR_TIME: a = (a & ~0xFF00u) | ((v & 0xFFu) << 8); break;
          case RATTR_AGE: a = (a & ~0xFF0000u) | ((v & 0xFFu) << 16); break;
          case RATTR_STRATUM: a = (a & ~0x70u) | ((min(v, 4u) & 7u) << 4); break;
          default: w = __dadd_rn(__dmul_rn(r.p, w), r.q); break;
        }
      }
    }
    const uint32_t st = attr_status(a), sev = attr_severity(a);
    if (valid) {
      c_st += 1ull << (16 * st);
      if (st != ST_D) {
        c_sev += (sev != SEV_E) ? (1ull << (16 * (sev - 1))) : 0ull;
        if (attr_age(a) >= 18u) s_q[min(attr_stratum(a), 4u)][threadIdx.x] += __double2ll_rn(__dmul_rn(w, qscale));
      }
      bxmin = min(bxmin, fx); bxmax = max(bxmax, fx); bymin = min(bymin, fy); bymax = max(bymax, fy);
      ohot[slot] = make_uint4((uint32_t)fx, (uint32_t)fy, a, h.w);
      owealth[slot] = w;
      opos[slot] = np;
      if (tinf) tinf[slot] = st == ST_I ? 0ull : ~0ull;
    }
    const bool is_inf = valid && st == ST_I;
    const unsigned m_inf = __ballot_sync(0xffffffffu, is_inf);
    if (m_inf) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(&ctl->n_inf, (unsigned)__popc(m_inf));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (is_inf) inf_list[base + __popc(m_inf & ((1u << lane) - 1u))] = slot;
    }
    // a thread's 16-bit count fields hold 65 535 agents; flush long before (grid-stride loops can be long here)
    if ((c_st & 0xFFFFull) > 60000ull || ((c_st >> 16) & 0xFFFFull) > 60000ull || ((c_st >> 32) & 0xFFFFull) > 60000ull ||
        (c_st >> 48) > 60000ull) {
      for (int k = 0; k < 4; ++k) atomicAdd(&ctl->cnt[k], (c_st >> (16 * k)) & 0xFFFFull);
      for (int k = 0; k < 3; ++k) atomicAdd(&ctl->cnt[4 + k], (c_sev >> (16 * k)) & 0xFFFFull);
      c_st = c_sev = 0;
    }
  }
  {
    unsigned long long cnt[7];
    long long q[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) q[k] = s_q[k][threadIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k) cnt[k] = (c_st >> (16 * k)) & 0xFFFFull;
#pragma unroll
    for (int k = 0; k < 3; ++k) cnt[4 + k] = (c_sev >> (16 * k)) & 0xFFFFull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 7; ++k) cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], o);
#pragma unroll
      for (int k = 0; k < 5; ++k) q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
      bxmin = min(bxmin, __shfl_xor_sync(0xffffffffu, bxmin, o));
      bxmax = max(bxmax, __shfl_xor_sync(0xffffffffu, bxmax, o));
      bymin = min(bymin, __shfl_xor_sync(0xffffffffu, bymin, o));
      bymax = max(bymax, __shfl_xor_sync(0xffffffffu, bymax, o));
    }
    if (lane == 0) {
      for (int k = 0; k < 7; ++k) atomicAdd(&B.cnt[k], cnt[k]);
      for (int k = 0; k < 5; ++k) atomicAdd((unsigned long long *)&B.q[k], (unsigned long long)q[k]);
      atomicMin(&B.bb[0], bxmin); atomicMax(&B.bb[1], bxmax); atomicMin(&B.bb[2], bymin); atomicMax(&B.bb[3], bymax);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 7; ++k)
      if (B.cnt[k]) atomicAdd(&ctl->cnt[k], B.cnt[k]);
    for (int k = 0; k < 5; ++k)
      if (B.q[k]) atomicAdd((unsigned long long *)&ctl->q[k], (unsigned long long)B.q[k]);
    if (B.bb[0] <= B.bb[1]) {
      atomicMin(&ctl->bb_xmin, B.bb[0]); atomicMax(&ctl->bb_xmax, B.bb[1]);
      atomicMin(&ctl->bb_ymin, B.bb[2]); atomicMax(&ctl->bb_ymax, B.bb[3]);
    }
  }
}

// Exact positions in and out (cabs_upload_f64 / cabs_download_positions_f64): records are in storage order on upload
// (before the first build) and scattered back to id order on download.
__global__ void k_pack_pos(unsigned n, const double *__restrict__ xf, const double *__restrict__ yf,
                           uint4 *__restrict__ hot, double2 *__restrict__ pos, Ctl *__restrict__ ctl) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint4 h = hot[i];
    double2 p = xf ? make_double2(xf[i], yf[i]) : make_double2((double)(int)h.x, (double)(int)h.y);
    int fx = 0, fy = 0;
    if (!floor_to_lattice(p.x, fx) || !floor_to_lattice(p.y, fy)) atomicOr(&ctl->err, kErrDomain);
    h.x = (uint32_t)fx;
    h.y = (uint32_t)fy;
    hot[i] = h;
    pos[i] = p;
  }
}
__global__ void k_unpack_pos(unsigned n, const uint4 *__restrict__ hot, const double2 *__restrict__ pos,
                             double *__restrict__ xf, double *__restrict__ yf) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned d = hot[i].w;
    if (d >= n) continue;
    xf[d] = pos[i].x;
    yf[d] = pos[i].y;
  }
}
// Upload without a sort: statistics (abs.py:291-314) and bounding box of the uploaded records in one pass over them.
// The first step after an upload sorts anyway (its pass B scatters into cell order whatever order it reads), so
// cabs_upload no longer builds a cell list of its own -- that was a second full scatter of the population, a third of
// the device time of an upload-step-statistics cycle.  The cell list of the uploaded positions is built on demand
// (contact-pair hook).
__global__ void __launch_bounds__(256)
k_upload_stats(Ctl *__restrict__ ctl, const uint4 *__restrict__ hot, const double *__restrict__ wealth) {
  __shared__ BlockStats B;
  __shared__ long long s_q[5][256];
#pragma unroll
  for (int k = 0; k < 5; ++k) s_q[k][threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 7; ++i) B.cnt[i] = 0;
    for (int i = 0; i < 5; ++i) B.q[i] = 0;
    B.bb[0] = B.bb[2] = INT_MAX;
    B.bb[1] = B.bb[3] = INT_MIN;
  }
  __syncthreads();
  unsigned long long cnt[7] = {0, 0, 0, 0, 0, 0, 0};
  int bxmin = INT_MAX, bxmax = INT_MIN, bymin = INT_MAX, bymax = INT_MIN;
  const double qscale = (double)(1ull << ctl->scale_bits);
  const unsigned n = ctl->n, lane = threadIdx.x & 31;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 h = hot[i];
    const uint32_t a = attr_normalize(h.z), st = attr_status(a), sev = attr_severity(a);
    cnt[st] += 1;
    if (st != ST_D) {
      if (sev != SEV_E) cnt[3 + sev] += 1;
      if (attr_age(a) >= 18u) s_q[min(attr_stratum(a), 4u)][threadIdx.x] += __double2ll_rn(__dmul_rn(wealth[i], qscale));
    }
    bxmin = min(bxmin, (int)h.x); bxmax = max(bxmax, (int)h.x);
    bymin = min(bymin, (int)h.y); bymax = max(bymax, (int)h.y);
  }
  long long q[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) q[k] = s_q[k][threadIdx.x];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 7; ++k) cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], o);
#pragma unroll
    for (int k = 0; k < 5; ++k) q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
    bxmin = min(bxmin, __shfl_xor_sync(0xffffffffu, bxmin, o));
    bxmax = max(bxmax, __shfl_xor_sync(0xffffffffu, bxmax, o));
    bymin = min(bymin, __shfl_xor_sync(0xffffffffu, bymin, o));
    bymax = max(bymax, __shfl_xor_sync(0xffffffffu, bymax, o));
  }
  if (lane == 0) {
    for (int k = 0; k < 7; ++k) atomicAdd(&B.cnt[k], cnt[k]);
    for (int k = 0; k < 5; ++k) atomicAdd((unsigned long long *)&B.q[k], (unsigned long long)q[k]);
    atomicMin(&B.bb[0], bxmin); atomicMax(&B.bb[1], bxmax); atomicMin(&B.bb[2], bymin); atomicMax(&B.bb[3], bymax);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 7; ++k)
      if (B.cnt[k]) atomicAdd(&ctl->cnt[k], B.cnt[k]);
    for (int k = 0; k < 5; ++k)
      if (B.q[k]) atomicAdd((unsigned long long *)&ctl->q[k], (unsigned long long)B.q[k]);
    if (B.bb[0] <= B.bb[1]) {
      atomicMin(&ctl->bb_xmin, B.bb[0]); atomicMax(&ctl->bb_xmax, B.bb[1]);
      atomicMin(&ctl->bb_ymin, B.bb[2]); atomicMax(&ctl->bb_ymax, B.bb[3]);
    }
  }
}
// ... and its closing: the statistics become the memo the first step's triggers and critical-limit test read, the
// grid of the first build is planned around the uploaded positions.  No cell table was built: the tables keep their state.
__global__ void k_close_upload(Ctl *ctl) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ctl->overflow = 0;
  close_stats(ctl, nullptr, 0);
  const int bx0 = ld_acc(&ctl->bb_xmin), bx1 = ld_acc(&ctl->bb_xmax);
  const int by0 = ld_acc(&ctl->bb_ymin), by1 = ld_acc(&ctl->bb_ymax);
  ctl->last_bb[0] = bx0; ctl->last_bb[1] = bx1; ctl->last_bb[2] = by0; ctl->last_bb[3] = by1;
  ctl->T = threshold_T(ctl->contagion_distance);
  plan_grid(ctl, bx0, bx1, by0, by1, true);
  reset_accumulators(ctl);
}

// Live-view feed (graphics.py:113-142,241-276 read every agent's position and status per frame): every stride-th
// agent of the reference's population list, in list order.
__global__ void k_snapshot(unsigned n, unsigned stride, const uint4 *__restrict__ hot, int *__restrict__ ox,
                           int *__restrict__ oy, uint8_t *__restrict__ status, uint8_t *__restrict__ severity) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 h = hot[i];
    if (h.w % stride != 0u || h.w >= n) continue;
    const unsigned d = h.w / stride;
    const uint32_t a = attr_normalize(h.z);
    ox[d] = (int)h.x;
    oy[d] = (int)h.y;
    status[d] = (uint8_t)attr_status(a);
    severity[d] = (uint8_t)attr_severity(a);
  }
}
__global__ void k_set_rules(Ctl *ctl, int f64, int rule_margin, int bx0, int bx1, int by0, int by1) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ctl->f64 = f64;
  ctl->rule_margin = rule_margin;
  ctl->rule_box[0] = bx0; ctl->rule_box[1] = bx1; ctl->rule_box[2] = by0; ctl->rule_box[3] = by1;
}

}  // namespace cabs
