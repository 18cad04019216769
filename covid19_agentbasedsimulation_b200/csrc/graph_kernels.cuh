// Device side of the GraphSimulation step (covid_abs/network/graph_abs.py:190-324) for sm_100a.
//
// One thread block owns one simulation and stays resident for a whole run of hourly iterations
// (`k_graph_run`): the reference's configs are small populations stepped thousands of times (1 440 hourly
// iterations, run_batch.py:384), which is launch-latency bound if every hour is a sequence of kernels, and
// scenario x seed ensembles are thousands of independent simulations -- one block
// each fills the GPU with no inter-block communication at all.  Business, government and healthcare
// accounts live in shared memory for the whole run; persons and houses stream from global memory (L2).
//
// Phases of one executed hour (block-wide barriers between them):
//   1 persons     on_person_move / routine move (network/agents.py:293-342), daily Person.update (362-415),
//                 purchase requests (graph_abs.py:230-233) -> one event word per person
//   2 stocks      Business.supply / checkin stock arithmetic (network/agents.py:82-93,101-102) resolved in the
//                 reference's person order with a block-wide scan of the max-plus maps s -> max(s + a, m)
//   3 accounts    Business/House/Government daily update and monthly accounting (network/agents.py:108-168,227-245)
//   4 contacts    the contagious persons (listed by phase 1) are grouped by hashed cell (graph_abs.py:256-300)
//   5 statistics  every Susceptible person PULLS contact() from the contagious persons within reach and stops at the
//                 first success (graph_abs.py:278-300: the outcome is an OR over pair-keyed draws, so the early exit is
//                 exact); GraphSimulation.get_statistics (graph_abs.py:302-324) -> one row of the history
//
// Money that several threads add to is 64-bit fixed point (`money_bits` fractional bits, every flow rounded
// to nearest-even once), so the sums do not depend on the order of the additions; the CPU restatement the tests check against
// does the same and the two agree bit for bit.  Draws: Philox4x32-10, counter (entity, iteration, stream, aux).
#pragma once
#include <limits.h>
#include <stdint.h>

#include "abs_rng.cuh"

namespace cabs {
namespace graph {

// One block of 1 024 threads per SM (64 registers each: the whole register file).  Two blocks of 512 share the issue slots
// and the L1 of an SM less well than one block twice the size -- 296 simulations x 10^5 persons, day 2 of "do nothing":
// 50.2 ms per 48 hours with 2 x 512 (static schedule), 33.7 ms with 1 x 1024 (the same simulations drawn from the task
// queue by 148 blocks); the 896-simulation bench 2.63e10 -> 2.70e10 agent-updates/s.  -DCABS_GRAPH_THREADS=512 restores
// the old shape.
#ifndef CABS_GRAPH_THREADS
#define CABS_GRAPH_THREADS 1024
#endif
constexpr int kThreads = CABS_GRAPH_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxBusiness = 16;
constexpr int kBGrid = 32;
// Shared memory budget: what is not shared memory is L1, and L1 matters to this kernel (with two blocks of 512 threads per
// SM: 164 KB carveout 2.63e10 agent-updates/s on the bench, 196 KB 2.43e10, 228 KB 1.77e10).  One block of 1 024 threads
// needs 84 KB (100 KB carveout, 156 KB of L1); maps twice the size filter better but cost more than they save (2.63e10).
constexpr int kHotCells = 512;
#ifndef CABS_MAP_BITS
#define CABS_MAP_BITS 262144
#endif
constexpr int kSrcBits = CABS_MAP_BITS;                              // bits of each contact map (a power of two >= 512)
constexpr size_t kGraphDynSmem = 2 * (size_t)kSrcBits / 8;           // two maps of 32 KB: dynamic shared memory
constexpr int kStats = 14;

enum : uint32_t { S_ = 0, I_ = 1, R_ = 2, D_ = 3 };
enum : uint32_t { SEV_E_ = 0, SEV_A_ = 1, SEV_H_ = 2, SEV_G_ = 3 };
enum : int { MODE_NONE = 0, MODE_LOCKDOWN = 1, MODE_CONDITIONAL = 2, MODE_VERTICAL = 3 };
// draw streams: shared contract with the CPU restatement the tests check against
enum : uint32_t { ST_MOVE = 0, ST_UPDATE = 1, ST_HOSP_MOVE = 2, ST_DEATH_MOVE = 3, ST_CONTACT = 4, ST_GOV = 5,
                  ST_ACCOUNT = 6, ST_SHOP = 16 };
constexpr uint32_t kGovId = 0xFFFFFFFFu, kBusinessId0 = 0xFFFF0000u;

// person attribute word: status[1:0] severity[3:2] stratum[6:4] new[7] infected_time[15:8] age[23:16]
//                        active[24] isolated[25]
constexpr uint32_t kActive = 1u << 24, kIsolated = 1u << 25;   // bit 7 ("new") is not used by the pull design
__device__ __forceinline__ uint32_t a_status(uint32_t a) { return a & 3u; }
__device__ __forceinline__ uint32_t a_sev(uint32_t a) { return (a >> 2) & 3u; }
__device__ __forceinline__ uint32_t a_stratum(uint32_t a) { return (a >> 4) & 7u; }
__device__ __forceinline__ uint32_t a_time(uint32_t a) { return (a >> 8) & 0xFFu; }
__device__ __forceinline__ uint32_t a_age(uint32_t a) { return (a >> 16) & 0xFFu; }
__device__ __forceinline__ uint32_t a_set(uint32_t a, uint32_t st, uint32_t sev, uint32_t time) {
  return (a & 0xFFFF0070u) | (st & 3u) | ((sev & 3u) << 2) | ((time & 0xFFu) << 8);
}

struct alignas(32) HouseRec {
  int x, y;
  long long wealth, expenses, incomes;
};

// Per-simulation descriptor (device memory).  Arrays are this simulation's segments.
struct Sim {
  int n, nh, nb;
  // persons
  int *px, *py;
  uint32_t *attr;
  int *house, *employer, *hire_seq;
  long long *pwealth;
  unsigned long long *events;   // scratch: one nibble per business (0 none, 1..9 request, 15 check-in)
  // houses: position and the three accounts persons touch every hour share ONE 32-byte sector (a person's house is a
  // random index: as four separate arrays every check-in cost four scattered sectors -- half of the kernel's DRAM
  // traffic in lockdown and night hours, profiles/r2_graph_final_summary.md)
  HouseRec *hrec;
  int *hstratum, *hsize;
  long long *hfixed;
  // businesses
  int *bx, *by, *bstratum, *bstocks, *bsales, *bnum;
  double *bprice;
  long long *bwealth, *bfixed, *bincomes;
  // government / healthcare
  int gov_x, gov_y, health_x, health_y;
  long long gov_wealth, gov_fixed, health_wealth, health_fixed, health_expenses;
  double gov_price;
  // parameters (graph_abs.py:13-32, abs.py:17-49, common.py:13-27)
  double amp[4];
  double bi[5];
  double min_income, min_expense, total_wealth, critical_limit, pop_size, scale;
  long long thr_hosp[9], thr_sev[9], thr_death[9], thr_rate;
  int T, T_bus, reach, incubation, contagion_time, recovering;
  int mode, sleep;
  unsigned k0, k1;
  // run state
  int iteration;          // graph_abs.py:25 starts at -1
  int next_hire_seq;
  double prev_infected, prev_hosp_severe;   // memo of the previous executed hour (statistics fractions)
  unsigned prev_susceptible, prev_sources;  // ... its Susceptible persons and contact sources (a scheduling hint only)
  // contact scratch: the contagious persons of the hour as {x, y, person, infected_time}
  unsigned long long *cell_keys;   // cell_cap: open-addressing table of the cells that hold a source (key = cell)
  unsigned *cell_count;   // cell_cap + 2: sources per slot of that table, then the start of the slot's run in cell_items
  int4 *cell_items;       // n: the sources grouped by cell
  int4 *src_list;         // n: the sources in the order phase 1 met them
  unsigned *req_idx;      // n: the persons that asked a business for goods this hour
  int cell_cap;
  // output
  double *history;        // hist_cap rows of kStats doubles, row = iteration % hist_cap
  int hist_cap;
  unsigned err;
};

__device__ __forceinline__ long long fx(double v, double scale) { return __double2ll_rn(__dmul_rn(v, scale)); }
__device__ __forceinline__ double fl(long long v, double scale) { return __ddiv_rn((double)v, scale); }
__device__ __forceinline__ int trunc_add(int base, float z, double sigma) {
  // int(base + normal(0, sigma)) of network/agents.py:304-305,323-324,339-340,349-350
  return __double2int_rz(__dadd_rn((double)base, __dmul_rn((double)z, sigma)));
}
__device__ __forceinline__ uint32_t mulhi_pick(uint32_t w, uint32_t n) { return __umulhi(w, n); }
__device__ __forceinline__ void atom_add(long long *p, long long v) {
  atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
}

// k-th component of a 16-byte vector held in registers (k is a loop counter, not a constant: selects, not an array)
template <typename V>
__device__ __forceinline__ auto vget(const V &v, int k) -> decltype(v.x) {
  return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w;
}
template <typename V, typename T>
__device__ __forceinline__ void vset(V &v, int k, T t) {
  if (k == 0) v.x = t; else if (k == 1) v.y = t; else if (k == 2) v.z = t; else v.w = t;
}
// four consecutive elements from `p + i0` (16-byte aligned): one 16-byte load for a whole group, guarded scalar loads
// for the last, partial one
template <typename V, typename T>
__device__ __forceinline__ V load4(const T *p, int i0, int n, T fill) {
  V v;
  if (i0 + 3 < n) {
    v = *reinterpret_cast<const V *>(p + i0);
  } else {
    v.x = p[i0];
    v.y = i0 + 1 < n ? p[i0 + 1] : fill;
    v.z = i0 + 2 < n ? p[i0 + 2] : fill;
    v.w = fill;
  }
  return v;
}
// the same through L2 only (data that atomics of this block have just written)
__device__ __forceinline__ uint4 load4_cg(const unsigned *p, int i0, int n) {
  uint4 v;
  if (i0 + 3 < n) {
    v = __ldcg(reinterpret_cast<const uint4 *>(p + i0));
  } else {
    v.x = __ldcg(p + i0);
    v.y = i0 + 1 < n ? __ldcg(p + i0 + 1) : 0u;
    v.z = i0 + 2 < n ? __ldcg(p + i0 + 2) : 0u;
    v.w = 0u;
  }
  return v;
}
template <typename V, typename T>
__device__ __forceinline__ void store4(T *p, int i0, int n, const V &v) {
  if (i0 + 3 < n) {
    *reinterpret_cast<V *>(p + i0) = v;
  } else {
    p[i0] = v.x;
    if (i0 + 1 < n) p[i0 + 1] = v.y;
    if (i0 + 2 < n) p[i0 + 2] = v.z;
  }
}

// block-wide helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ long long block_sum(long long v, long long *s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  long long t = 0;
  for (int k = 0; k < kWarps; ++k) t += s_red[k];
  return t;
}

// shared state of one block / simulation
struct Shared {
  long long bwealth[kMaxBusiness], bfixed[kMaxBusiness], bincomes[kMaxBusiness];
  int bstocks[kMaxBusiness], bsales[kMaxBusiness], bnum[kMaxBusiness], bx[kMaxBusiness], by[kMaxBusiness];
  double bprice[kMaxBusiness];
  long long gov_wealth, health_wealth, health_expenses;
  long long red[kWarps];
  long long acc[8];
  int wa[kWarps][kMaxBusiness], wm[kWarps][kMaxBusiness];   // per-warp max-plus totals
  int ws[kWarps][kMaxBusiness];                              // stock before each warp of the current chunk
  unsigned present_req[3];   // businesses with a purchase request in the current chunk (the order-dependent ones)
  // which businesses can be within business_distance of a position: kBGrid x kBGrid cells over the bounding box of the
  // businesses (+- reach), one bit per business.  A person outside the box, or in a cell with no bit, has none in reach.
  unsigned bgrid[kBGrid][kBGrid];
  int bg_x0, bg_y0, bg_shift;
  // check-ins at work per (business, stratum): the amount a check-in takes from the business depends on the stratum only
  // (network/agents.py:101-103), so the hour's check-ins are COUNTED with native 32-bit shared atomics and charged once
  // (a 64-bit shared atomicAdd is a compare-and-swap loop, and 7 000 colleagues hit the same word every working hour)
  unsigned ci_cnt[kMaxBusiness][5];
  // purchase requests of the hour: quantity asked of every business, the persons who asked (Sim::req_idx), and the
  // businesses whose stock at the start of the hour does not cover what was asked (only those need the person order)
  unsigned req_q[kMaxBusiness];
  unsigned n_req, scarce;
  // the first cells phase 4 meets, counted in shared memory (see there)
  unsigned long long hot_key[kHotCells];
  unsigned hot_cnt[kHotCells];
  unsigned n_contagious;
  unsigned cnt[8];
  int sel;                   // result slot of block-wide selections
  int next_hire_seq;
};

constexpr int kNegInf = -(1 << 29);

// Timing-only build (-DCABS_GRAPH_TIMERS): clocks of thread 0 between the phases of an executed hour, summed over all
// blocks into g_graph_phase_clk[hour class][phase] (class 0 = hour 0, 1 = working hours 8..16, 2 = evening 17..23).
#ifdef CABS_GRAPH_TIMERS
__device__ unsigned long long g_graph_phase_clk[4][8];
#define GT_BEGIN() long long gt_prev = clock64(); const int gt_cls = hour == 0 ? 0 : (hour < 17 ? 1 : 2)
#define GT_MARK(k)                                                                   \
  do {                                                                               \
    if (tid == 0) {                                                                  \
      const long long gt_now = clock64();                                            \
      atomicAdd(&g_graph_phase_clk[gt_cls][k], (unsigned long long)(gt_now - gt_prev)); \
      gt_prev = gt_now;                                                              \
    }                                                                                \
  } while (0)
#define GT_PULL_PARAM , unsigned *gt_cnt
#define GT_PULL_ARG , gt_cnt
#define GT_COUNT(k) (++gt_cnt[k])
#else
#define GT_BEGIN() do { } while (0)
#define GT_MARK(k) do { } while (0)
#define GT_PULL_PARAM
#define GT_PULL_ARG
#define GT_COUNT(k) ((void)0)
#endif

// compose max-plus maps: apply (a1, m1) first, then (a2, m2):  s -> max(max(s + a1, m1) + a2, m2)
__device__ __forceinline__ void compose(int a1, int m1, int a2, int m2, int &a, int &m) {
  a = a1 + a2;
  m = max(m1 > kNegInf / 2 ? m1 + a2 : kNegInf, m2);
}

// ---------------------------------------------------------------------------------------------------
// Contacts (graph_abs.py:256-300).  The reference walks the pairs (i, j) within contagion_distance and calls contact()
// for a Susceptible / contagious pair; a Susceptible person ends the hour Infected if ANY of its pairs succeeds, and the
// draws are keyed by the pair, so neither the order of the pairs nor skipping the pairs of a person that is already
// infected changes anything.  The sources are grouped by cell of a grid whose edge is a power of two >= reach; the cells
// that hold a source live in an open-addressing hash table (key = the cell, linear probing, 2^k slots with k chosen
// every hour for a load factor <= 1/2: no bounding box, nothing to clamp).  Every Susceptible person then looks up the
// <= 9 cells around it and stops at the first contact that infects.  Persons pile up on single lattice points
// (thousands of colleagues at the four points of a workplace, the hospitalised at the hospital): pushed from the
// sources that is sources x targets pair draws per pile, pulled with the early exit it is about one draw per target;
// and because a slot is exactly one cell, nobody ever walks a pile it is not next to (with plain hashed buckets the
// few persons whose lookups collided with the hospital's bucket walked 13 000 records each while their block waited).
constexpr unsigned long long kEmptyCell = 0x8000000080000000ull;
__device__ __forceinline__ unsigned long long cell_key(int cx, int cy) {
  return ((unsigned long long)(unsigned)cx << 32) | (unsigned long long)(unsigned)cy;
}
__device__ __forceinline__ unsigned cell_hash(int cx, int cy) {
  unsigned h = (unsigned)cx * 0x9E3779B1u ^ (unsigned)cy * 0x85EBCA77u;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  return h ^ (h >> 13);
}
// Two bit maps in (dynamic) shared memory, 512 x 512 cells each, indexed by the cell coordinates modulo 512 (nothing to
// hash for the cells a Susceptible person tests every hour; an alias only costs a lookup in the table):
//   sbits  the cell holds a source of this hour's contacts
//   nbits  some cell within reach of this one holds a source (every source sets the up to nine bits around it): ONE
//          test tells most Susceptible persons that nobody is near.
// 64 KB per block on purpose: the loads of this kernel stream (L1 hit rate < 3 %), so the space is worth more as a
// filter than as cache -- with 4 KB maps a quarter of the persons of a lockdown went to the table in L2 every hour
// because of aliases (profiles/r2_graph_phase_times_pull.log).
__device__ __forceinline__ unsigned cell_bit(int cx, int cy) {
  return (((unsigned)cy << 9) | ((unsigned)cx & 511u)) & (unsigned)(kSrcBits - 1);     // 512 columns x kSrcBits / 512 rows
}
// In the hours in which the sources outnumber the Susceptible persons five to one (the peak of an unmitigated epidemic:
// 87 % of the persons are sources, 5 % can still be infected) phase 1 has every Susceptible person mark the cells within
// its reach -- in the memory of the `sbits` map, which nobody needs before phase 4 -- and a source whose cell is not
// marked (nobody it could infect is near) never enters the table.  Building the table for all of them was 45 MB of
// random sector traffic per simulation and hour, three quarters of such an hour.
// NOTE sbits and nbits must use the SAME cell -> bit mapping: a source dilates into nbits only when it is the first to
// set its sbits bit, which is right because cells that alias in one map alias, with all their neighbours, in the other
// (a build with a smaller sbits map infected fewer persons).
// contact geometry: cell edge = power of two >= reach; which of the 3 x 3 cells can be within reach at all (with cells
// of edge 1 that is a property of the offset alone)
__device__ __forceinline__ void contact_geometry(const Sim &S, int &eshift, unsigned &qmask) {
  eshift = 0;
  while ((1 << eshift) < S.reach && eshift < 30) ++eshift;
  qmask = 0x1FFu;
  if (eshift == 0) {
    qmask = 0u;
    for (int q = 0; q < 9; ++q) {
      const long long ox = q % 3 - 1, oy = q / 3 - 1;
      if (ox * ox + oy * oy <= (long long)S.T) qmask |= 1u << q;
    }
  }
}
__device__ __forceinline__ bool is_source(const Sim &S, uint32_t a) {
  // Infected persons whose infected_time can fall inside the window for some (low, up) of graph_abs.py:289-290
  return a_status(a) == I_ && (int)a_time(a) >= S.incubation - 1 && (int)a_time(a) <= S.contagion_time;
}
// Structured control flow on purpose (flags in the loop conditions, no early return): with a return from inside the
// nested loops the compiler reconverges a warp only at the exit of the CALLER's person loop, and the lanes that had a
// lookup to do ran the rest of the hour one or two at a time (measured: 1.9 active lanes per instruction, 17 x the
// instructions of the converged loop, profiles/r2_graph_final_summary.md).
__device__ __forceinline__ bool pull_contact(const Sim &S, const unsigned *sbits, const unsigned *nbits,
                                             const unsigned long long *keys,
                                             const unsigned *table, const int4 *items, unsigned mask, int eshift,
                                             unsigned qmask, int i, int x, int y, uint32_t it GT_PULL_PARAM) {
  const int cx = x >> eshift, cy = y >> eshift;
  const long long T = S.T;
  bool infected = false;
  {
    const unsigned nb0 = cell_bit(cx, cy);
    if (!((nbits[nb0 >> 5] >> (nb0 & 31u)) & 1u)) qmask = 0u;      // nobody in any cell around this one
  }
  // qmask: the cells of the 3 x 3 neighbourhood that can hold somebody within reach.  With cells of edge 1 (reach 1,
  // every shipped scenario) that is a property of the offset alone -- the diagonals are out -- and was worked out once
  // per hour; with larger cells it depends on where in its cell the person stands and is tested below.
#pragma unroll 1
  for (unsigned qm = qmask; qm && !infected; qm &= qm - 1u) {
    const int q = __ffs(qm) - 1, oy = (q * 11) >> 5, ccx = cx + (q - 3 * oy) - 1, ccy = cy + oy - 1;
    const unsigned hb = cell_bit(ccx, ccy);
    bool look = (sbits[hb >> 5] >> (hb & 31u)) & 1u;
    if (look && eshift > 0) {     // nearest lattice point of the cell: a cell out of reach as a whole is not looked up
      const int edge1 = (1 << eshift) - 1, lox = ccx << eshift, loy = ccy << eshift;
      const long long ddx = x < lox ? lox - x : (x > lox + edge1 ? x - lox - edge1 : 0);
      const long long ddy = y < loy ? loy - y : (y > loy + edge1 ? y - loy - edge1 : 0);
      look = ddx * ddx + ddy * ddy <= T;
    }
    if (look) {
      GT_COUNT(1);
      const unsigned long long key = cell_key(ccx, ccy);
      unsigned b = cell_hash(ccx, ccy) & mask;
      unsigned long long kk = __ldcg(keys + b);
      while (kk != key && kk != kEmptyCell) {      // linear probing: the cell's slot, or an empty one (no source there)
        b = (b + 1u) & mask;
        kk = __ldcg(keys + b);
      }
      if (kk == key) {
        const unsigned je = __ldcg(table + b + 1);     // table[b]: start of the run of slot b, table[b + 1]: its end
        unsigned k = __ldcg(table + b);
        for (; k < je && !infected; ++k) {
          const int4 c = items[k];
          GT_COUNT(2);
          const long long dx = c.x - x, dy = c.y - y;
          if (dx * dx + dy * dy <= T) {
            GT_COUNT(3);
            const int j = c.z, tj = c.w;
            const uint4 w = philox4x32_10((uint32_t)min(i, j), it, ST_CONTACT, (uint32_t)max(i, j), S.k0, S.k1);
            const int low = -1 + (int)(w.x >> 31), up = -1 + (int)(w.y >> 31);           // np.random.randint(-1, 1)
            infected = tj >= S.incubation + low && tj <= S.contagion_time + up && (long long)w.z <= S.thr_rate;
          }                                                                              // graph_abs.py:292-297
        }
      }
    }
  }
  return infected;
}

// ---------------------------------------------------------------------------------------------------
// Person moves (network/agents.py:293-342).  Return the business checked in at (move_to_work) or -1.
struct PersonCtx {
  const Sim *S;
  Shared *sh;
  uint32_t it;
  double scale;
};

__device__ __forceinline__ double person_expenses(const Sim &S, uint32_t a) { return __dmul_rn(S.bi[min(a_stratum(a), 4u)], S.min_expense); }
__device__ __forceinline__ double person_incomes(const Sim &S, uint32_t a) {
  return (a & kActive) ? __dmul_rn(S.bi[min(a_stratum(a), 4u)], S.min_income) : 0.0;
}

__device__ __forceinline__ void house_checkin(const Sim &S, int h, uint32_t a) {   // network/agents.py:202-207
  const long long v = fx(__ddiv_rn(person_expenses(S, a), 720.0), S.scale);
  atom_add(&S.hrec[h].wealth, -v);
  atom_add(&S.hrec[h].expenses, v);
}

__device__ __forceinline__ void move_freely(const Sim &S, uint32_t i, uint32_t it, uint32_t stream, uint32_t a, int &x,
                                            int &y) {
  if (a_sev(a) != SEV_A_) return;
  const uint4 w = philox4x32_10(i, it, stream, 0u, S.k0, S.k1);
  float z0, z1;
  normal_pair(w.x, w.y, z0, z1);
  const double amp = S.amp[a_status(a)];
  x = trunc_add(x, z0, amp);
  y = trunc_add(y, z1, amp);
}

__device__ __forceinline__ void move_to_home(const Sim &S, uint32_t i, uint32_t it, uint32_t stream, uint32_t a, int h,
                                             int &x, int &y) {
  if (a_sev(a) != SEV_A_) return;
  if (h >= 0) {
    house_checkin(S, h, a);
    const uint4 w = philox4x32_10(i, it, stream, 0u, S.k0, S.k1);
    float z0, z1;
    normal_pair(w.x, w.y, z0, z1);
    x = trunc_add(S.hrec[h].x, z0, 0.25);
    y = trunc_add(S.hrec[h].y, z1, 0.25);
  } else {
    atom_add(&S.pwealth[i], -fx(__ddiv_rn(person_incomes(S, a), 720.0), S.scale));
    move_freely(S, i, it, stream, a, x, y);
  }
}

// Person.supply (network/agents.py:286-291): income goes to the house, or to the person if homeless
__device__ __forceinline__ void person_supply(const Sim &S, uint32_t i, int h, long long v) {
  if (h >= 0) {
    atom_add(&S.hrec[h].wealth, v);
    atom_add(&S.hrec[h].incomes, v);
  } else {
    atom_add(&S.pwealth[i], v);
  }
}
__device__ __forceinline__ void person_demand(const Sim &S, uint32_t i, int h, long long v) {   // 279-284
  if (h >= 0) {
    atom_add(&S.hrec[h].wealth, -v);
    atom_add(&S.hrec[h].expenses, v);
  } else {
    atom_add(&S.pwealth[i], -v);
  }
}

// Business.fire (network/agents.py:44-53)
__device__ __forceinline__ void fire(const Sim &S, Shared &sh, int b, uint32_t i, uint32_t a) {
  S.employer[i] = -1;
  S.hire_seq[i] = -1;
  const long long v = fx(person_incomes(S, a), S.scale);
  person_supply(S, i, S.house[i], v);
  atom_add(&sh.bwealth[b], -v);
  atomicSub(&sh.bnum[b], 1);
  atom_add(&sh.bfixed[b], -fx(__dmul_rn(__ddiv_rn(person_expenses(S, a), 720.0), 24.0), S.scale));
}

// ---------------------------------------------------------------------------------------------------
// Two schedules of the same hour code:
//   kQueue = false   block b owns simulation b for the whole run (grid = n_sims): nothing is shared between blocks;
//                    right whenever the simulations fit the GPU at once (one block per SM).
//   kQueue = true    a persistent grid (one block per resident slot) draws (simulation, hour) tasks from a ticket
//                    counter, hour-major: task t = hour * n_sims + simulation.  A simulation's hours still run in order
//                    (a worker waits until `progress[simulation]` says the hour before is done -- it was handed out
//                    n_sims tickets earlier, to a worker that is running, so the wait cannot deadlock), but which
//                    block runs which hour is decided as blocks become free.  An ensemble of 896 simulations on 148
//                    slots is then 6.05 slots' worth of time instead of 7 rounds (the last round of the static
//                    schedule runs 8 blocks on an otherwise idle GPU).  Between tasks the accounts go back to global
//                    memory (a few hundred bytes); results are identical bit for bit.
template <bool kQueue>
__global__ void __launch_bounds__(kThreads, 1024 / kThreads)
k_graph_run(Sim *sims, int iterations, int n_sims, unsigned *ticket, int *progress) {
  // the descriptor (pointers, parameters) is read by every thread all the time: keep a copy in shared memory
  __shared__ Sim S;
  __shared__ Shared sh;
  extern __shared__ unsigned s_maps[];                       // kGraphDynSmem bytes: the contact maps
  unsigned *const s_nbits = s_maps, *const s_sbits = s_maps + kSrcBits / 32, *const s_want = s_sbits;
  __shared__ unsigned s_task;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (bool first_pass = true;; first_pass = false) {
  int sim_index = blockIdx.x, hours = iterations, task_hour = 0;
  if (kQueue) {
    if (tid == 0) s_task = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned task = s_task;
    if (task >= (unsigned)n_sims * (unsigned)iterations) break;
    sim_index = (int)(task % (unsigned)n_sims);
    task_hour = (int)(task / (unsigned)n_sims);
    hours = 1;
    if (tid == 0) {
      while (*(volatile int *)&progress[sim_index] < task_hour) __nanosleep(200);
      __threadfence();
    }
    __syncthreads();
    __threadfence();   // what the previous hour of this simulation wrote (on another SM) is read from L2, not a stale L1 line
  } else if (!first_pass) {
    break;
  }
  for (unsigned k = tid; k < sizeof(Sim) / 4; k += kThreads)
    reinterpret_cast<uint32_t *>(&S)[k] = kQueue ? __ldcg(reinterpret_cast<const uint32_t *>(&sims[sim_index]) + k)
                                                 : reinterpret_cast<const uint32_t *>(&sims[sim_index])[k];
  __syncthreads();
  const int n = S.n, nh = S.nh, nb = S.nb;
  const double scale = S.scale;

  // ---- load the accounts that live in shared memory for the run
  if (tid < nb) {
    sh.bwealth[tid] = S.bwealth[tid]; sh.bfixed[tid] = S.bfixed[tid]; sh.bincomes[tid] = S.bincomes[tid];
    sh.bstocks[tid] = S.bstocks[tid]; sh.bsales[tid] = S.bsales[tid]; sh.bnum[tid] = S.bnum[tid];
    sh.bx[tid] = S.bx[tid]; sh.by[tid] = S.by[tid]; sh.bprice[tid] = S.bprice[tid];
  }
  if (tid == 0) {
    sh.gov_wealth = S.gov_wealth; sh.health_wealth = S.health_wealth; sh.health_expenses = S.health_expenses;
    sh.next_hire_seq = S.next_hire_seq;
  }
  __syncthreads();

  int bus_reach = S.T_bus > 0 ? (int)__dsqrt_rn((double)S.T_bus) + 1 : 0;         // >= floor(sqrt(T_bus))
  if (S.T_bus < 0) bus_reach = -1;                                                  // nothing is in reach
  for (int c = tid; c < kBGrid * kBGrid; c += kThreads) (&sh.bgrid[0][0])[c] = 0u;
  if (tid == 0) {
    int x0 = INT_MAX, x1 = INT_MIN, y0 = INT_MAX, y1 = INT_MIN;
    for (int b = 0; b < nb; ++b) {
      x0 = min(x0, sh.bx[b]); x1 = max(x1, sh.bx[b]); y0 = min(y0, sh.by[b]); y1 = max(y1, sh.by[b]);
    }
    const int r = max(bus_reach, 0);
    x0 -= r; y0 -= r; x1 += r; y1 += r;
    int shift = 0;
    while (nb > 0 && ((((long long)x1 - x0) >> shift) >= kBGrid || (((long long)y1 - y0) >> shift) >= kBGrid)) ++shift;
    sh.bg_x0 = x0; sh.bg_y0 = y0; sh.bg_shift = shift;
  }
  __syncthreads();
  if (tid < nb && bus_reach >= 0) {
    const int sft = sh.bg_shift;
    const int cx0 = (sh.bx[tid] - bus_reach - sh.bg_x0) >> sft, cx1 = (sh.bx[tid] + bus_reach - sh.bg_x0) >> sft;
    const int cy0 = (sh.by[tid] - bus_reach - sh.bg_y0) >> sft, cy1 = (sh.by[tid] + bus_reach - sh.bg_y0) >> sft;
    for (int cy = cy0; cy <= cy1; ++cy)
      for (int cx = cx0; cx <= cx1; ++cx) atomicOr(&sh.bgrid[cy][cx], 1u << tid);
  }
  __syncthreads();

  int it = S.iteration;
  double prev_infected = S.prev_infected, prev_hs = S.prev_hosp_severe;
  unsigned prev_susceptible = S.prev_susceptible, prev_sources = S.prev_sources;

  for (int step = 0; step < hours; ++step) {
    ++it;                                                                       // graph_abs.py:192
    const int hour = it % 24, day = it / 24;
    const bool bed = hour < 8, work = hour >= 8 && hour < 16, free_t = hour >= 17, lunch = hour == 12;
    const bool new_day = hour == 0, work_day = (day % 7 + 1) != 1 && (day % 7 + 1) != 7;
    const bool new_month = (day % 30 + 1) == 1 && hour == 0;
    double *row = S.history + (size_t)(it % S.hist_cap) * kStats;
    if (S.sleep && !new_day && bed) {                                           // run_batch.py:8-13 (on_execute)
      // the reference returns before anything changes; batch_experiment re-reads the same statistics
      const double *prev_row = S.history + (size_t)((it + S.hist_cap - 1) % S.hist_cap) * kStats;
      if (tid < kStats) row[tid] = prev_row[tid];
      __syncthreads();
      continue;
    }
    GT_BEGIN();
    const bool contacts_on = S.T >= 0;
    const bool lock_all = S.mode == MODE_LOCKDOWN || (S.mode == MODE_CONDITIONAL && prev_infected > 0.05);
    const bool crit_active = prev_hs >= S.critical_limit;                        // network/agents.py:389-390
    const bool monthly = it > 1 && new_month;
    // sources that nobody can catch anything from are left out of the table when they are the many (decided from the
    // previous hour's counts: a hint, the result does not depend on it)
    const bool want_filter = contacts_on && prev_sources > 5u * prev_susceptible;
    if (want_filter)
      for (int c = tid * 4; c < kSrcBits / 32; c += kThreads * 4) *reinterpret_cast<uint4 *>(s_want + c) = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) sh.n_contagious = sh.n_req = sh.scarce = 0u;
    if (tid < kMaxBusiness) sh.req_q[tid] = 0u;
    if (tid < kMaxBusiness * 5) (&sh.ci_cnt[0][0])[tid] = 0u;
    __syncthreads();

    // ================= phase 1: persons (graph_abs.py:210-233) =================
    // A thread takes FOUR consecutive persons at a time: their five columns arrive as five 16-byte loads (four times the
    // bytes in flight per thread of a person-per-thread loop -- the hour is bound by memory latency, not by bandwidth),
    // the three columns that change and the event words leave the same way.
    // Every thread makes the same number of trips and the warp is brought back together after every person (see
    // pull_contact: lanes that leave a loop early otherwise stay apart until the next block-wide barrier).
    for (int base = 0; base < n; base += kThreads * 4) {
      const int i0 = base + tid * 4;
      uint4 A4 = make_uint4(D_, D_, D_, D_);
      int4 X4 = make_int4(0, 0, 0, 0), Y4 = X4, H4 = make_int4(-1, -1, -1, -1), E4 = H4;
      if (i0 < n) {
        A4 = load4<uint4, uint32_t>(S.attr, i0, n, (uint32_t)D_);
        X4 = load4<int4, int>(S.px, i0, n, 0); Y4 = load4<int4, int>(S.py, i0, n, 0);
        H4 = load4<int4, int>(S.house, i0, n, -1); E4 = load4<int4, int>(S.employer, i0, n, -1);
      }
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
      const int i = i0 + k;
      // the person of this trip sits in component .x; the vectors are rotated by one at the end of the trip (after four
      // trips they are back in place, updated): moves instead of selects on a loop counter
      uint32_t a = A4.x;
      int x = X4.x, y = Y4.x;
      const int h = H4.x, emp0 = E4.x;
      if (i < n) {
      unsigned long long ev = 0ull;
      if (a_status(a) != D_) {
        int checked_in = -1, emp = emp0;   // emp: -1 once the critical-limit rule below has fired the person
        // ---- on_person_move (run_batch.py:16-50) or the default routine (graph_abs.py:212-220)
        if (lock_all || (S.mode == MODE_VERTICAL && !(a & kActive))) {
          if (h >= 0) house_checkin(S, h, a);
        } else if ((a & kIsolated) || bed) {
          move_to_home(S, i, it, ST_MOVE, a, h, x, y);
        } else if (lunch || free_t || !work_day) {
          move_freely(S, i, it, ST_MOVE, a, x, y);
        } else if (work_day && work) {                                           // move_to_work, 293-310
          if (a_sev(a) == SEV_A_ && (a & kActive)) {
            const int b = emp0;
            if (b >= 0) {
              const uint4 w = philox4x32_10(i, it, ST_MOVE, 0u, S.k0, S.k1);
              float z0, z1;
              normal_pair(w.x, w.y, z0, z1);
              x = trunc_add(sh.bx[b], z0, 0.25);
              y = trunc_add(sh.by[b], z1, 0.25);
              atomicAdd(&sh.ci_cnt[b][min(a_stratum(a), 4u)], 1u);                        // checkin, 101-103
              checked_in = b;
            } else {
              move_freely(S, i, it, ST_MOVE, a, x, y);
            }
          }
        }
        // ---- daily health update (network/agents.py:362-415)
        if (new_day && a_status(a) == I_) {
          uint32_t st = I_, sev = a_sev(a), time = a_time(a) + 1;
          const uint32_t age = a_age(a);
          const uint32_t ix = min(age > 10u ? age / 10u - 1u : 0u, 8u);
          const uint4 u = philox4x32_10(i, it, ST_UPDATE, 0u, S.k0, S.k1);
          if (sev == SEV_A_) {
            if ((long long)u.x < S.thr_hosp[ix]) {
              sev = SEV_H_;
              const uint4 w = philox4x32_10(i, it, ST_HOSP_MOVE, 0u, S.k0, S.k1);   // move_to(healthcare), 344-354
              float z0, z1;
              normal_pair(w.x, w.y, z0, z1);
              x = trunc_add(S.health_x, z0, 0.25);
              y = trunc_add(S.health_y, z1, 0.25);
              atom_add(&sh.health_expenses, fx(person_expenses(S, a), scale));       // checkin, 104-105
            }
          } else if (sev == SEV_H_) {
            if ((long long)u.x < S.thr_sev[ix]) {
              sev = SEV_G_;
              if (crit_active) {                                                     // 388-401
                st = D_;
                sev = SEV_A_;
                if (h >= 0) {                                                        // remove_mate, 196-200
                  atom_add(&S.hrec[h].wealth, -fx(__ddiv_rn(fl(S.pwealth[i], scale), 2.0), scale));
                  atomicSub(&S.hsize[h], 1);
                  atom_add(&S.hfixed[h], -fx(__dmul_rn(__ddiv_rn(person_expenses(S, a), 720.0), 24.0), scale));
                } else {
                  atom_add(&sh.gov_wealth, -fx(person_expenses(S, a), scale));
                }
                if (emp >= 0) {
                  fire(S, sh, emp, i, a);
                  emp = -1;
                } else {
                  atom_add(&sh.gov_wealth, -fx(person_expenses(S, a), scale));
                }
              }
            }
          }
          bool died = false;
          if ((long long)u.y < S.thr_death[ix]) {                                    // 403-408
            st = D_;
            sev = SEV_A_;
            a = a_set(a, st, sev, time);
            move_to_home(S, i, it, ST_DEATH_MOVE, a, h, x, y);
            died = true;
          }
          if (!died && (int)time > S.recovering) {                                   // 410-413
            time = 0;
            st = R_;
            sev = SEV_A_;
          }
          a = a_set(a, st, sev, time);
        }
        // ---- purchase requests at every business in reach other than the employer (graph_abs.py:230-233)
        if (checked_in >= 0) ev |= 15ull << (4 * checked_in);
        if (a_sev(a) == SEV_A_) {
          uint4 q4 = make_uint4(0, 0, 0, 0);
          int have = -1;
          // the businesses whose box of reach covers this cell (almost always none), in ascending order
          const int gx = (x - sh.bg_x0) >> sh.bg_shift, gy = (y - sh.bg_y0) >> sh.bg_shift;
          unsigned near = ((unsigned)gx < (unsigned)kBGrid && (unsigned)gy < (unsigned)kBGrid) ? sh.bgrid[gy][gx] : 0u;
          for (; near; near &= near - 1u) {
            const int b = __ffs(near) - 1;
            if (b == emp) continue;
            const long long dx = sh.bx[b] - x, dy = sh.by[b] - y;
            if (dx * dx + dy * dy > (long long)S.T_bus) continue;
            if (have != (b >> 2)) {
              q4 = philox4x32_10(i, it, ST_SHOP + (b >> 2), 0u, S.k0, S.k1);
              have = b >> 2;
            }
            const uint32_t w = (b & 3) == 0 ? q4.x : (b & 3) == 1 ? q4.y : (b & 3) == 2 ? q4.z : q4.w;
            const unsigned qty = 1u + mulhi_pick(w, 9u);                              // np.random.randint(1, 10)
            ev |= (unsigned long long)qty << (4 * b);
            atomicAdd(&sh.req_q[b], qty);
          }
          if (have >= 0) S.req_idx[atomicAdd(&sh.n_req, 1u)] = (unsigned)i;           // asked somebody for goods
        }
        if (want_filter && a_status(a) == S_) {     // the cells a source must be in to reach this person
          int eshift;
          unsigned qmask;
          contact_geometry(S, eshift, qmask);
          const int cx = x >> eshift, cy = y >> eshift;
          for (unsigned qm = qmask; qm; qm &= qm - 1u) {
            const int q = __ffs(qm) - 1, oy = (q * 11) >> 5;
            const unsigned wb = cell_bit(cx + (q - 3 * oy) - 1, cy + oy - 1);
            if (!((s_want[wb >> 5] >> (wb & 31u)) & 1u)) atomicOr(&s_want[wb >> 5], 1u << (wb & 31u));
          }
        }
        // ---- the sources of this hour's contacts (graph_abs.py:278-290): position and infected_time are final here
        if (contacts_on && is_source(S, a)) {
          const unsigned m = __activemask();          // the lanes that are here together share one reservation
          const int leader = __ffs(m) - 1;
          unsigned at = 0;
          if (lane == leader) at = atomicAdd(&sh.n_contagious, (unsigned)__popc(m));
          at = __shfl_sync(m, at, leader) + (unsigned)__popc(m & ((1u << lane) - 1u));
          S.src_list[at] = make_int4(x, y, i, (int)a_time(a));
        }
      }
      S.events[i] = ev;
      }
      A4 = make_uint4(A4.y, A4.z, A4.w, a);
      X4 = make_int4(X4.y, X4.z, X4.w, x);
      Y4 = make_int4(Y4.y, Y4.z, Y4.w, y);
      H4 = make_int4(H4.y, H4.z, H4.w, h);
      E4 = make_int4(E4.y, E4.z, E4.w, emp0);
      __syncwarp();
      }
      if (i0 < n) {
        store4<uint4, uint32_t>(S.attr, i0, n, A4);
        store4<int4, int>(S.px, i0, n, X4);
        store4<int4, int>(S.py, i0, n, Y4);
      }
    }
    __syncthreads();
    if (tid < nb) {   // Business.checkin (network/agents.py:101-103): expenses / 720 of every employee that came in
      long long total = 0;
      unsigned came = 0;
      for (int k = 0; k < 5; ++k) {
        total += (long long)sh.ci_cnt[tid][k] * fx(__ddiv_rn(__dmul_rn(S.bi[k], S.min_expense), 720.0), scale);
        came += sh.ci_cnt[tid][k];
      }
      sh.bwealth[tid] -= total;
      // Stocks (network/agents.py:82-93,101-102): a check-in adds one unit, a request takes min(quantity, stock) -- the
      // only order-dependent rule of the person loop.  A business whose stock at the START of the hour covers everything
      // asked of it this hour serves every request in full whatever the order (check-ins only add): settled here in
      // one step.  The others are resolved in person order by phase 2.
      const unsigned asked = sh.req_q[tid];
      if (sh.bstocks[tid] >= 0 && (unsigned)sh.bstocks[tid] >= asked) {
        sh.bstocks[tid] += (int)came - (int)asked;
        sh.bsales[tid] += (int)asked;
      } else {
        atomicOr(&sh.scarce, 1u << tid);
      }
    }
    __syncthreads();
    {
      const unsigned scarce = sh.scarce, nreq = sh.n_req;
      for (unsigned t = tid; t < nreq; t += kThreads) {     // Business.supply for the requests served in full
        const int i = (int)S.req_idx[t];
        unsigned long long e = S.events[i];
        const uint32_t a = S.attr[i];
        const int h = S.house[i];
        while (e) {
          const int b = (__ffsll((long long)e) - 1) >> 2;
          const unsigned nib = (unsigned)(e >> (4 * b)) & 15u;
          e &= ~(15ull << (4 * b));
          if (nib > 9u || ((scarce >> b) & 1u)) continue;
          const long long value = fx(__dmul_rn(__dmul_rn(sh.bprice[b], (double)a_stratum(a)), (double)nib), scale);
          person_demand(S, i, h, value);
          atom_add(&sh.bwealth[b], value);
          atom_add(&sh.bincomes[b], value);
        }
      }
    }
    __syncthreads();
    GT_MARK(1);

    // ================= phase 2: stocks in person order (network/agents.py:82-93,101-102) =================
    // Only a purchase request reads the stock (min(quantity, stock)); check-ins just add one.  Per chunk of kThreads
    // persons the businesses that got a request are scanned in person order (their check-ins included); for every
    // other business the check-ins of the chunk are simply counted.  `present_req` rotates through three slots so
    // that one barrier per chunk is enough (slot k+1 is cleared while chunk k-1 may still be read and chunk k is set).
    if (sh.scarce) {
      const unsigned scarce = sh.scarce;
      if (tid < 3) sh.present_req[tid] = 0u;
      __syncthreads();
      unsigned long long ev_next = tid < n ? S.events[tid] : 0ull;
      int slot = 0;
      for (int base = 0; base < n; base += kThreads, slot = slot == 2 ? 0 : slot + 1) {
        const int i = base + tid;
        const unsigned long long ev = ev_next;
        ev_next = i + kThreads < n ? S.events[i + kThreads] : 0ull;      // the next chunk is on its way during this one
        if (tid == 0) sh.present_req[slot == 2 ? 0 : slot + 1] = 0u;
        unsigned req = 0, ci = 0;
        for (unsigned long long e = ev; e;) {
          const int b = (__ffsll((long long)e) - 1) >> 2;
          const unsigned nib = (unsigned)(e >> (4 * b)) & 15u;
          if (nib == 15u) ci |= 1u << b; else req |= 1u << b;
          e &= ~(15ull << (4 * b));
        }
        req &= scarce;      // the other businesses were settled after phase 1
        ci &= scarce;
        const unsigned wreq = __reduce_or_sync(0xffffffffu, req);
        unsigned wci = __reduce_or_sync(0xffffffffu, ci);
        if (lane == 0 && wreq) atomicOr(&sh.present_req[slot], wreq);
        __syncthreads();
        const unsigned present = sh.present_req[slot];
        for (wci &= ~present; wci; wci &= wci - 1u) {                     // check-ins nobody's request can observe
          const int b = __ffs(wci) - 1;
          const unsigned who = __ballot_sync(0xffffffffu, (ci >> b) & 1u);
          if (lane == 0) atomicAdd(&sh.bstocks[b], __popc(who));
        }
        if (!present) continue;
        // the requested businesses one after the other (almost always none or one): scan inside the warps, carry the
        // stock through the warps, apply -- everything of a business stays in registers
        for (unsigned todo = present; todo; todo &= todo - 1u) {
          const int b = __ffs(todo) - 1;
          const unsigned nib = (unsigned)((ev >> (4 * b)) & 15ull);
          int ea = nib == 15u ? 1 : -(int)nib;
          int em = (nib >= 1u && nib <= 9u) ? 0 : kNegInf;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int pa = __shfl_up_sync(0xffffffffu, ea, o), pm = __shfl_up_sync(0xffffffffu, em, o);
            if (lane >= o) compose(pa, pm, ea, em, ea, em);
          }
          if (lane == 31) {
            sh.wa[wid][0] = ea;
            sh.wm[wid][0] = em;
          }
          __syncthreads();
          // one thread carries the stock through the warps of the chunk: ws[w] = stock before warp w
          if (tid == 0) {
            int s0 = sh.bstocks[b];
            for (int w = 0; w < kWarps; ++w) {
              sh.ws[w][0] = s0;
              s0 = max(s0 + sh.wa[w][0], sh.wm[w][0]);
            }
            sh.bstocks[b] = s0;
          }
          __syncthreads();
          const int s0 = sh.ws[wid][0];
          const int after = max(s0 + ea, em);
          // exclusive prefix = inclusive prefix of the lane before
          const int pa = __shfl_up_sync(0xffffffffu, ea, 1), pm = __shfl_up_sync(0xffffffffu, em, 1);
          if (nib >= 1u && nib <= 9u) {
            const int before = lane == 0 ? s0 : max(s0 + pa, pm);
            const int taken = before - after;                                           // min(qty, stocks)
            const uint32_t a = S.attr[i];
            const long long value = fx(__dmul_rn(__dmul_rn(sh.bprice[b], (double)a_stratum(a)), (double)taken), scale);
            person_demand(S, i, S.house[i], value);
            atom_add(&sh.bwealth[b], value);
            atom_add(&sh.bincomes[b], value);
            atomicAdd(&sh.bsales[b], taken);
          }
          __syncthreads();   // ws / wa / wm are free again
        }
      }
      __syncthreads();
    }

    GT_MARK(2);
    // ================= phase 3: accounts (graph_abs.py:235-254) =================
    // ---- businesses: daily update (network/agents.py:147-153)
    if (new_day && tid < nb) {
      atom_add(&sh.gov_wealth, fx(__ddiv_rn(fl(sh.bfixed[tid], scale), 3.0), scale));
      atom_add(&sh.bwealth[tid], -sh.bfixed[tid]);
    }
    __syncthreads();
    // ---- businesses: monthly accounting, one after the other (network/agents.py:115-145)
    if (monthly) {
      for (int b = 0; b < nb; ++b) {
        long long labor = 0;
        for (int i = tid; i < n; i += kThreads) {                                     // demand(), 55-76
          if (S.employer[i] != b) continue;
          const uint32_t a = S.attr[i];
          if (a_status(a) == D_ || a_sev(a) != SEV_A_) continue;
          const long long v = fx(person_incomes(S, a), scale);
          person_supply(S, i, S.house[i], v);
          labor += v;
        }
        labor = block_sum(labor, sh.red);
        __syncthreads();
        if (tid == 0) {
          sh.bwealth[b] -= labor;
          const long long tax = fx(__dadd_rn(__dmul_rn(S.gov_price, (double)sh.bnum[b]),
                                             __ddiv_rn(fl(sh.bincomes[b], scale), 20.0)), scale);   // taxes(), 108-113
          sh.gov_wealth += tax;
          sh.bwealth[b] -= tax;
          const double lt = __dadd_rn(fl(labor, scale), fl(tax, scale)), inc = fl(sh.bincomes[b], scale);
          sh.sel = __dmul_rn(2.0, lt) < inc ? 1 : (lt > inc ? 2 : 0);                   // 1 hire, 2 fire
        }
        __syncthreads();
        const int action = sh.sel;
        __syncthreads();
        const uint4 w = philox4x32_10(kBusinessId0 + b, it, ST_ACCOUNT, 0u, S.k0, S.k1);
        if (action == 1) {
          // ---- hire the ix-th unemployed person in population order (graph_abs.py:43-45)
          long long cnt = 0;
          for (int i = tid; i < n; i += kThreads) {
            const uint32_t a = S.attr[i];
            cnt += (S.employer[i] < 0 && (a & kActive) && a_status(a) != D_ && a_sev(a) == SEV_A_) ? 1 : 0;
          }
          cnt = block_sum(cnt, sh.red);
          __syncthreads();
          if (cnt > 0) {
            const int target = (int)mulhi_pick(w.x, (uint32_t)cnt);
            // ordered search: chunks of kThreads persons, running count
            int seen = 0;
            if (tid == 0) sh.sel = -1;
            __syncthreads();
            for (int base = 0; base < n; base += kThreads) {
              const int i = base + tid;
              bool u = false;
              if (i < n) {
                const uint32_t a = S.attr[i];
                u = S.employer[i] < 0 && (a & kActive) && a_status(a) != D_ && a_sev(a) == SEV_A_;
              }
              const unsigned bal = __ballot_sync(0xffffffffu, u);
              if (lane == 0) sh.wa[wid][0] = __popc(bal);
              __syncthreads();
              int before = seen;
              for (int k = 0; k < wid; ++k) before += sh.wa[k][0];
              const int my = before + __popc(bal & ((1u << lane) - 1u));
              if (u && my == target) sh.sel = i;
              int total = 0;
              for (int k = 0; k < kWarps; ++k) total += sh.wa[k][0];
              seen += total;
              __syncthreads();
              if (sh.sel >= 0) break;
            }
            if (tid == 0 && sh.sel >= 0) {                                             // hire(), 36-42
              const int i = sh.sel;
              const uint32_t a = S.attr[i];
              S.employer[i] = b;
              S.hire_seq[i] = sh.next_hire_seq++;
              sh.bnum[b] += 1;
              sh.bfixed[b] += fx(__dmul_rn(__ddiv_rn(person_expenses(S, a), 720.0), 24.0), scale);
            }
            __syncthreads();
          }
        } else if (action == 2 && sh.bnum[b] > 0) {
          // ---- fire the ix-th employee in list (= hire) order: smallest v with #(hire_seq <= v) > ix
          const int target = (int)mulhi_pick(w.x, (uint32_t)sh.bnum[b]);
          long long members = 0;
          for (int i = tid; i < n; i += kThreads) members += S.employer[i] == b ? 1 : 0;
          members = block_sum(members, sh.red);
          __syncthreads();
          if (target < members) {
            int lo = 0, hi = sh.next_hire_seq;    // answer in [lo, hi)
            while (hi - lo > 1) {
              const int mid = lo + (hi - lo) / 2;  // count employees with hire_seq < mid
              long long c = 0;
              for (int i = tid; i < n; i += kThreads) c += (S.employer[i] == b && S.hire_seq[i] < mid) ? 1 : 0;
              c = block_sum(c, sh.red);
              __syncthreads();
              if (c > target) hi = mid; else lo = mid;
            }
            for (int i = tid; i < n; i += kThreads)
              if (S.employer[i] == b && S.hire_seq[i] == lo) fire(S, sh, b, i, S.attr[i]);
            __syncthreads();
          }
        }
        if (tid == 0) {
          sh.bincomes[b] = 0;
          sh.bsales[b] = 0;
        }
        __syncthreads();
      }
    }
    // ---- houses (network/agents.py:227-245)
    if (new_day || monthly) {
      long long gov = 0;
      for (int h = tid; h < nh; h += kThreads) {
        if (S.hsize[h] <= 0) continue;
        if (new_day) gov += fx(__ddiv_rn(fl(S.hfixed[h], scale), 10.0), scale);
        if (monthly) {
          const long long tax = fx(__dadd_rn(__dmul_rn(S.gov_price, (double)S.hsize[h]),
                                             __ddiv_rn(fl(S.hrec[h].expenses, scale), 10.0)), scale);
          gov += tax;
          S.hrec[h].wealth -= tax;
          S.hrec[h].incomes = 0;
          S.hrec[h].expenses = 0;
        }
      }
      gov = block_sum(gov, sh.red);
      if (tid == 0) sh.gov_wealth += gov;
      __syncthreads();
    }
    // ---- government and healthcare (network/agents.py:147-166, 132-139)
    if (new_day && tid == 0) {
      const uint4 w = philox4x32_10(kGovId, it, ST_GOV, 0u, S.k0, S.k1);
      const int b = (int)mulhi_pick(w.x, (uint32_t)nb);
      const int q = min(1 + (int)mulhi_pick(w.y, 9u), sh.bstocks[b]);
      const long long value = fx(__dmul_rn(__dmul_rn(sh.bprice[b], 4.0), (double)q), scale);   // buys as stratum 4
      sh.gov_wealth -= value;
      sh.bwealth[b] += value;
      sh.bincomes[b] += value;
      sh.bstocks[b] -= q;
      sh.bsales[b] += q;
      sh.gov_wealth += fx(__ddiv_rn(fl(S.health_fixed, scale), 3.0), scale);                   // healthcare.update()
      sh.health_wealth -= S.health_fixed;
    }
    __syncthreads();
    if (monthly) {
      long long paid = 0;
      for (int i = tid; i < n; i += kThreads) {
        const uint32_t a = S.attr[i];
        if (a_status(a) == D_ || a_sev(a) != SEV_A_) continue;
        const int h = S.house[i];
        const long long v = fx(person_expenses(S, a), scale);
        if (h < 0) {                                                                    // homeless
          person_supply(S, i, h, v);
          paid += v;
        }
        if (S.employer[i] < 0 && (a & kActive)) {                                       // unemployed
          person_supply(S, i, h, v);
          paid += v;
        }
      }
      paid = block_sum(paid, sh.red);
      if (tid == 0) {
        const long long labor = sh.health_expenses;                                     // demand(healthcare)
        sh.health_wealth += labor;
        sh.gov_wealth -= labor + paid;
      }
      __syncthreads();
    }

    GT_MARK(3);
    // ================= phase 4: contacts, the sources grouped by hashed cell (graph_abs.py:256-300) =================
    const unsigned nsrc = sh.n_contagious;
    int eshift;
    unsigned qmask;
    contact_geometry(S, eshift, qmask);
    unsigned cmask = 255u;                                    // slots: a power of two >= 2 x sources (<= cell_cap)
    while (cmask + 1u < 2u * nsrc && 2u * (cmask + 1u) <= (unsigned)S.cell_cap) cmask = 2u * cmask + 1u;
    if (nsrc && want_filter) {      // the marks live in the memory of sbits: settle every source before that map is reset
      for (unsigned t = tid; t < nsrc; t += kThreads) {
        const int4 r = S.src_list[t];
        const unsigned wb = cell_bit(r.x >> eshift, r.y >> eshift);
        if (!((s_want[wb >> 5] >> (wb & 31u)) & 1u)) S.src_list[t].w = -1;     // not in the table: nobody can reach it
      }
      __syncthreads();
    }
    if (nsrc) {
      const int nbuckets = (int)cmask + 1;
      for (int c = tid * 4; c <= nbuckets; c += kThreads * 4)      // nbuckets is a multiple of 256
        store4<uint4, unsigned>(S.cell_count, c, nbuckets + 1, make_uint4(0u, 0u, 0u, 0u));
      for (int c = tid * 2; c < nbuckets; c += kThreads * 2)
        *reinterpret_cast<ulonglong2 *>(S.cell_keys + c) = make_ulonglong2(kEmptyCell, kEmptyCell);
      for (int c = tid * 4; c < 2 * kSrcBits / 32; c += kThreads * 4)
        *reinterpret_cast<uint4 *>(s_maps + c) = make_uint4(0u, 0u, 0u, 0u);
      for (int c = tid; c < kHotCells; c += kThreads) {
        sh.hot_key[c] = kEmptyCell;
        sh.hot_cnt[c] = 0u;
      }
      __syncthreads();
      // ---- count: every source gets its RANK inside its cell (kept in the upper bits of the record's last word).
      // Thousands of sources share one cell at a workplace or at the hospital; their increments of one counter in
      // global memory are served one after the other by the L2 (milliseconds per hour at the peak of an epidemic).
      // The first kHotCells cells the block meets are therefore counted in shared memory -- the big piles are among the
      // first few hundred records with near certainty -- and handed to the table once, with their totals.
      for (unsigned base = 0; base < nsrc; base += kThreads) {
        const unsigned t = base + tid;
        if (t < nsrc) {
          const int4 r = S.src_list[t];
          const int ccx = r.x >> eshift, ccy = r.y >> eshift;
          if (r.w != -1) {
          const unsigned h = cell_hash(ccx, ccy), hb = cell_bit(ccx, ccy);
          const unsigned long long key = cell_key(ccx, ccy);
          unsigned rank = 0xFFFFFFFFu;
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) {            // a cell is either always found here or never (slots are not freed)
            const unsigned slot = ((h >> 12) + pr) & (kHotCells - 1);
            if (rank == 0xFFFFFFFFu) {
              const unsigned long long prev = atomicCAS(&sh.hot_key[slot], kEmptyCell, key);
              if (prev == kEmptyCell || prev == key) rank = atomicAdd(&sh.hot_cnt[slot], 1u);
            }
          }
          if (rank == 0xFFFFFFFFu) {
            unsigned b = h & cmask;
            unsigned long long prev = atomicCAS(&S.cell_keys[b], kEmptyCell, key);   // claim the cell's slot or find it
            while (prev != kEmptyCell && prev != key) {
              b = (b + 1u) & cmask;
              prev = atomicCAS(&S.cell_keys[b], kEmptyCell, key);
            }
            rank = atomicAdd(&S.cell_count[b], 1u);
          }
          S.src_list[t].w = r.w | (int)(rank << 8);
          if (!((s_sbits[hb >> 5] >> (hb & 31u)) & 1u)) {        // first source seen in this cell (or a benign race)
            atomicOr(&s_sbits[hb >> 5], 1u << (hb & 31u));
            for (unsigned qm = qmask; qm; qm &= qm - 1u) {        // qmask is symmetric: the cells that can reach this one
              const int q = __ffs(qm) - 1, oy = (q * 11) >> 5;
              const unsigned nbq = cell_bit(ccx + (q - 3 * oy) - 1, ccy + oy - 1);
              atomicOr(&s_nbits[nbq >> 5], 1u << (nbq & 31u));
            }
          }
          }
        }
        __syncwarp();
      }
      __syncthreads();
      for (int c = tid; c < kHotCells; c += kThreads) {          // the cells counted in shared memory enter the table
        const unsigned long long key = sh.hot_key[c];
        if (key != kEmptyCell) {
          unsigned b = cell_hash((int)(unsigned)(key >> 32), (int)(unsigned)key) & cmask;
          unsigned long long prev = atomicCAS(&S.cell_keys[b], kEmptyCell, key);
          while (prev != kEmptyCell && prev != key) {
            b = (b + 1u) & cmask;
            prev = atomicCAS(&S.cell_keys[b], kEmptyCell, key);
          }
          atomicAdd(&S.cell_count[b], sh.hot_cnt[c]);
        }
      }
      __syncthreads();
      // exclusive scan of cell_count[0..nbuckets]: cell_count[b] = start of the run of slot b.  Every warp owns one
      // contiguous slab: it sums it, the 16 sums are combined, then it walks the slab again with the running total in a
      // register -- two streaming passes with independent 16-byte loads and no block-wide barrier per tile (a scan that
      // synchronises the block every 512 entries spends a memory round trip per tile: 1 ms of the hour at 2^18 slots).
      {
        const int entries = nbuckets + 1;
        const int slab = ((entries + kWarps * 128 - 1) / (kWarps * 128)) * 128;
        const int w0 = wid * slab, w1 = min(w0 + slab, entries);
        unsigned sum = 0;
        for (int base = w0; base < w1; base += 128) {
          const int c = base + lane * 4;
          if (c < entries) {
            const uint4 v = load4_cg(S.cell_count, c, entries);
            sum += v.x + v.y + v.z + v.w;
          }
        }
        sum = __reduce_add_sync(0xffffffffu, sum);
        if (lane == 0) sh.wa[wid][0] = (int)sum;
        __syncthreads();
        unsigned carry = 0;
        for (int k = 0; k < wid; ++k) carry += (unsigned)sh.wa[k][0];
        for (int base = w0; base < w1; base += 128) {
          const int c = base + lane * 4;
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (c < entries) v = load4_cg(S.cell_count, c, entries);
          const unsigned mine = v.x + v.y + v.z + v.w;
          unsigned inc = mine;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
          }
          const unsigned e0 = carry + inc - mine;
          if (c < entries) store4<uint4, unsigned>(S.cell_count, c, entries, make_uint4(e0, e0 + v.x, e0 + v.x + v.y, e0 + v.x + v.y + v.z));
          carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncthreads();
      }
      // fill: start of the cell's run + the rank taken above (no atomics)
      for (unsigned base = 0; base < nsrc; base += kThreads) {
        const unsigned t = base + tid;
        const int4 r = t < nsrc ? S.src_list[t] : make_int4(0, 0, 0, -1);
        if (r.w != -1) {
          const int ccx = r.x >> eshift, ccy = r.y >> eshift;
          const unsigned long long key = cell_key(ccx, ccy);
          unsigned b = cell_hash(ccx, ccy) & cmask;
          while (__ldcg(S.cell_keys + b) != key) b = (b + 1u) & cmask;
          S.cell_items[__ldcg(S.cell_count + b) + ((unsigned)r.w >> 8)] = make_int4(r.x, r.y, r.z, r.w & 255);
        }
        __syncwarp();
      }
      __syncthreads();
    }

    GT_MARK(4);
    // ================= phase 5: statistics (graph_abs.py:302-324) =================
    {
      // per-thread counts packed 16 bits each (a thread sees n / kThreads persons): no dynamically indexed local array
      unsigned long long c_st = 0ull, c_sev = 0ull;
#ifdef CABS_GRAPH_TIMERS
      unsigned gt_cnt[5] = {0, 0, 0, 0, 0};   // pulls, bitmap hits, items visited, pairs in range (= draws), infections
#endif
      // every thread of the block makes the same number of trips (guards inside), and the warp is brought back
      // together after every person: a lane that had a lookup to do must not run the rest of the loop on its own
      for (int base = 0; base < n; base += kThreads * 4) {
        const int i0 = base + tid * 4;
        uint4 A4 = make_uint4(0u, 0u, 0u, 0u);
        int4 X4 = make_int4(0, 0, 0, 0), Y4 = X4;
        if (i0 < n) {
          A4 = load4<uint4, uint32_t>(S.attr, i0, n, 0u);
          if (nsrc) { X4 = load4<int4, int>(S.px, i0, n, 0); Y4 = load4<int4, int>(S.py, i0, n, 0); }
        }
        bool changed = false;
        if (nsrc) {
          // the Susceptible persons of the group, one after the other (the warp makes as many trips as its busiest lane)
          unsigned todo = (a_status(A4.x) == S_ ? 1u : 0u) | (a_status(A4.y) == S_ ? 2u : 0u) |
                          (a_status(A4.z) == S_ ? 4u : 0u) | (a_status(A4.w) == S_ ? 8u : 0u);
          if (i0 + 3 >= n) todo &= i0 < n ? (1u << (n - i0)) - 1u : 0u;
          while (__any_sync(0xffffffffu, todo != 0u)) {
            if (todo) {
              const int k = __ffs(todo) - 1;
              todo &= todo - 1u;
              GT_COUNT(0);
              if (pull_contact(S, s_sbits, s_nbits, S.cell_keys, S.cell_count, S.cell_items, cmask, eshift, qmask, i0 + k,
                               vget(X4, k), vget(Y4, k), (uint32_t)it GT_PULL_ARG)) {
                GT_COUNT(4);
                vset(A4, k, (vget(A4, k) & ~3u) | I_);   // contact() rewrites the status only (graph_abs.py:297-298)
                changed = true;
              }
            }
          }
        }
        if (i0 < n) {
          const uint32_t a4[4] = {A4.x, A4.y, A4.z, A4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t a = a4[k];
            if (i0 + k < n) {
              c_st += 1ull << (16 * a_status(a));
              if (a_status(a) != D_ && a_sev(a) != SEV_E_) c_sev += 1ull << (16 * (a_sev(a) - 1));
            }
          }
        }
        if (changed) store4<uint4, uint32_t>(S.attr, i0, n, A4);
      }
      __syncwarp();
#ifdef CABS_GRAPH_TIMERS
      for (int k = 0; k < 5; ++k) {
        const unsigned v = __reduce_add_sync(0xffffffffu, gt_cnt[k]);
        if (lane == 0 && v) atomicAdd(&g_graph_phase_clk[3][k], (unsigned long long)v);
      }
      if (tid == 0) atomicAdd(&g_graph_phase_clk[3][5], (unsigned long long)nsrc);
      {
        const unsigned v = __reduce_max_sync(0xffffffffu, gt_cnt[2]);
        if (lane == 0 && v) atomicMax(&g_graph_phase_clk[3][6], (unsigned long long)v);
      }
#endif
#ifdef CABS_GRAPH_TIMERS
      __syncthreads();
      GT_MARK(6);     // phase 5, person loop (contacts + counts)
#endif
      long long q[5] = {0, 0, 0, 0, 0};
      for (int h = tid; h < nh; h += kThreads) {
        const int s = S.hstratum[h];
        const long long wv = S.hrec[h].wealth;
#pragma unroll
        for (int k = 0; k < 5; ++k) q[k] += s == k ? wv : 0ll;
      }
#ifdef CABS_GRAPH_TIMERS
      __syncthreads();
      GT_MARK(7);     // phase 5, house loop
#endif
      if (tid < 8) sh.cnt[tid] = 0;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned v = __reduce_add_sync(0xffffffffu, (unsigned)(c_st >> (16 * k)) & 0xFFFFu);
        if (lane == 0 && v) atomicAdd(&sh.cnt[k], v);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const unsigned v = __reduce_add_sync(0xffffffffu, (unsigned)(c_sev >> (16 * k)) & 0xFFFFu);
        if (lane == 0 && v) atomicAdd(&sh.cnt[4 + k], v);
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const long long t = block_sum(q[k], sh.red);
        if (tid == 0) sh.acc[k] = t;
      }
      __syncthreads();
      if (tid == 0) {
        long long bsum = 0;
        for (int b = 0; b < nb; ++b) bsum += sh.bwealth[b];
        for (int k = 0; k < 5; ++k) row[k] = __ddiv_rn(fl(sh.acc[k], scale), S.total_wealth);
        row[5] = __ddiv_rn(fl(bsum, scale), S.total_wealth);
        row[6] = __ddiv_rn(fl(sh.gov_wealth, scale), S.total_wealth);
        for (int k = 0; k < 7; ++k) row[7 + k] = __ddiv_rn((double)sh.cnt[k], S.pop_size);
      }
      __syncthreads();
      prev_infected = __ddiv_rn((double)sh.cnt[1], S.pop_size);
      prev_hs = __dadd_rn(__ddiv_rn((double)sh.cnt[6], S.pop_size), __ddiv_rn((double)sh.cnt[5], S.pop_size));
      prev_susceptible = sh.cnt[0];
      prev_sources = nsrc;
      __syncthreads();
    }
    GT_MARK(5);
  }

  // ---- write the shared accounts back
  if (tid < nb) {
    S.bwealth[tid] = sh.bwealth[tid]; S.bfixed[tid] = sh.bfixed[tid]; S.bincomes[tid] = sh.bincomes[tid];
    S.bstocks[tid] = sh.bstocks[tid]; S.bsales[tid] = sh.bsales[tid]; S.bnum[tid] = sh.bnum[tid];
  }
  if (tid == 0) {
    Sim &G = sims[sim_index];
    G.gov_wealth = sh.gov_wealth; G.health_wealth = sh.health_wealth; G.health_expenses = sh.health_expenses;
    G.next_hire_seq = sh.next_hire_seq;
    G.iteration = it;
    G.prev_infected = prev_infected;
    G.prev_hosp_severe = prev_hs;
    G.prev_susceptible = prev_susceptible;
    G.prev_sources = prev_sources;
  }
  if (kQueue) {   // publish the hour: everything this block wrote, then the progress counter
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      *(volatile int *)&progress[sim_index] = task_hour + 1;
    }
  }
  }   // next task
}

// Statistics of the initial state (row -1 is never read by batch_experiment, but get_statistics() before the
// first execute() and the first hour's memo need it): same reduction as phase 5, one block per simulation.
__global__ void __launch_bounds__(kThreads) k_graph_initial_stats(Sim *sims, double *rows) {
  Sim &S = sims[blockIdx.x];
  __shared__ long long red[kWarps];
  __shared__ long long acc[12];
  const int tid = threadIdx.x;
  long long v[12] = {0};
  for (int i = tid; i < S.n; i += kThreads) {
    const uint32_t a = S.attr[i];
    v[a_status(a)]++;
    if (a_status(a) != D_ && a_sev(a) != SEV_E_) v[3 + a_sev(a)]++;
  }
  for (int h = tid; h < S.nh; h += kThreads) {
    const int s = S.hstratum[h];
    if (s >= 0 && s < 5) v[7 + s] += S.hrec[h].wealth;
  }
  for (int k = 0; k < 12; ++k) {
    const long long t = block_sum(v[k], red);
    if (tid == 0) acc[k] = t;
  }
  __syncthreads();
  if (tid == 0) {
    double *row = rows + (size_t)blockIdx.x * kStats;
    long long bsum = 0;
    for (int b = 0; b < S.nb; ++b) bsum += S.bwealth[b];
    for (int k = 0; k < 5; ++k) row[k] = __ddiv_rn(fl(acc[7 + k], S.scale), S.total_wealth);
    row[5] = __ddiv_rn(fl(bsum, S.scale), S.total_wealth);
    row[6] = __ddiv_rn(fl(S.gov_wealth, S.scale), S.total_wealth);
    for (int k = 0; k < 7; ++k) row[7 + k] = __ddiv_rn((double)acc[k], S.pop_size);
    S.prev_infected = row[8];
    S.prev_hosp_severe = __dadd_rn(row[13], row[12]);
    // the row before the first iteration: a slept first hour copies it
    double *h = S.history + (size_t)((S.iteration + S.hist_cap) % S.hist_cap) * kStats;
    for (int k = 0; k < kStats; ++k) h[k] = row[k];
  }
}

// What batch_experiment does with the rows of its experiments (experiments.py:200-208): per iteration and metric
// the minimum, mean, population standard deviation (ddof = 0) and maximum over the simulations of the handle.
// One block per iteration, one warp per metric; lanes stride over the simulations and combine in a fixed
// butterfly, so the result does not depend on scheduling.  out[iteration][metric] = {min, avg, std, max}.
__global__ void __launch_bounds__(kStats * 32)
k_graph_aggregate(const double *__restrict__ history, const double *__restrict__ initial_rows, int n_sims, int hist_cap,
                  long long first, double *__restrict__ out) {
  const long long it = first + blockIdx.x;
  const int m = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row = it < 0 ? 0 : (size_t)(it % hist_cap) * kStats;
  const double *base = it < 0 ? initial_rows : history + row;
  const size_t stride = it < 0 ? (size_t)kStats : (size_t)hist_cap * kStats;
  double mn = INFINITY, mx = -INFINITY, sum = 0.0;
  for (int s = lane; s < n_sims; s += 32) {
    const double v = base[(size_t)s * stride + m];
    mn = fmin(mn, v);
    mx = fmax(mx, v);
    sum = __dadd_rn(sum, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sum = __dadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
  }
  const double mean = __ddiv_rn(sum, (double)n_sims);
  double ss = 0.0;
  for (int s = lane; s < n_sims; s += 32) {
    const double d = __dsub_rn(base[(size_t)s * stride + m], mean);
    ss = __dadd_rn(ss, __dmul_rn(d, d));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss = __dadd_rn(ss, __shfl_xor_sync(0xffffffffu, ss, o));
  if (lane == 0) {
    double *o4 = out + ((size_t)blockIdx.x * kStats + m) * 4;
    o4[0] = mn;
    o4[1] = mean;
    o4[2] = __dsqrt_rn(__ddiv_rn(ss, (double)n_sims));
    o4[3] = mx;
  }
}

// The same reduction in mergeable form over a RANGE of the handle's simulations: out[iteration][metric] =
// {min, sum, sum of squares, max}.  A scenario x seed sweep keeps the seeds of a scenario in consecutive simulations;
// every GPU reduces its part of each scenario and the parts merge with min / + / + / max (three tiny all-reduces).
__global__ void __launch_bounds__(kStats * 32)
k_graph_moments(const double *__restrict__ history, int sim_first, int sim_count, int hist_cap, long long first,
                double *__restrict__ out) {
  const long long it = first + blockIdx.x;
  const int m = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double *base = history + (size_t)(it % hist_cap) * kStats + (size_t)sim_first * hist_cap * kStats;
  const size_t stride = (size_t)hist_cap * kStats;
  double mn = INFINITY, mx = -INFINITY, sum = 0.0, sq = 0.0;
  for (int s = lane; s < sim_count; s += 32) {
    const double v = base[(size_t)s * stride + m];
    mn = fmin(mn, v);
    mx = fmax(mx, v);
    sum = __dadd_rn(sum, v);
    sq = __dadd_rn(sq, __dmul_rn(v, v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sum = __dadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
    sq = __dadd_rn(sq, __shfl_xor_sync(0xffffffffu, sq, o));
  }
  if (lane == 0) {
    double *o4 = out + ((size_t)blockIdx.x * kStats + m) * 4;
    o4[0] = mn; o4[1] = sum; o4[2] = sq; o4[3] = mx;
  }
}

}  // namespace graph
}  // namespace cabs
