// Host-side world builder of GraphSimulation ensembles (include/covidabs.h: cabs_graph_build_world).
//
// GraphSimulation.initialize (covid_abs/network/graph_abs.py:89-188) builds a world in a Python loop over the
// persons; the mirror in network/graph_abs.py::build_world repeats it call for call so that np.random.seed(s) gives
// the reference's world bit for bit -- at 4 us per draw that is ~1 s per 10^5-person world, hours for the 7 168 worlds
// of the ensemble sweep.  This file restates the SAME construction (every rule and quirk of
// graph_abs.py:89-188 and network/agents.py:36-42,186-194, in the same order over the persons) in C++ with its own
// counter-based draws, so a world costs milliseconds and worlds are built concurrently (the function is re-entrant;
// the Python binding calls it from a thread pool with the GIL released).  Worlds from here and from the Python
// builder are draws from the same distribution, not the same draws (tests/test_world_builder.py compares them).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/covidabs.h"

namespace {

// Philox4x32-10 (Salmon et al., SC'11), counter = running index, key = seed
struct Draws {
  uint32_t k0, k1;
  uint64_t ctr = 0;
  uint32_t buf[4];
  int have = 0;
  explicit Draws(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  void refill() {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x9E3779B9u, c3 = 0x7F4A7C15u;
    uint32_t a = k0, b = k1;
    for (int r = 0; r < 10; ++r) {
      const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
      const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ b;
      c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    buf[0] = c0; buf[1] = c1; buf[2] = c2; buf[3] = c3;
    have = 4;
    ++ctr;
  }
  uint32_t word() {
    if (!have) refill();
    return buf[--have];
  }
  double uniform() {   // (0, 1), 53 bits
    const uint64_t hi = word(), lo = word();
    return ((double)(((hi << 32) | lo) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  int randint(int n) { return n <= 1 ? 0 : (int)(((uint64_t)word() * (uint64_t)n) >> 32); }   // np.random.randint(0, n)
  double normal() {    // Box-Muller, one value per call
    const double u = uniform(), v = uniform();
    return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v);
  }
  double gamma_int(int k) {   // Gamma(k, 1), integer shape: -ln of a product of k uniforms
    double p = 1.0;
    for (int i = 0; i < k; ++i) p *= uniform();
    return -log(p);
  }
  double beta25() {    // np.random.beta(2, 5)
    const double x = gamma_int(2), y = gamma_int(5);
    return x / (x + y);
  }
};

inline int clip_int(double v, int lo, int hi) {   // np.clip(int(v), lo, hi): int() truncates toward zero
  const long long t = (long long)v;
  return (int)(t < lo ? lo : t > hi ? hi : t);
}

}  // namespace

extern "C" int cabs_graph_world_sizes(const cabs_graph_world_spec *spec, int32_t *n_persons, int32_t *max_houses,
                                      int32_t *n_business) {
  if (!spec || !n_persons || !max_houses || !n_business) return CABS_ERR_INVALID;
  if (spec->population_size < 1 || spec->homemates_avg < 1 || spec->total_business < 1) return CABS_ERR_INVALID;
  *n_persons = spec->population_size;
  *max_houses = spec->population_size / spec->homemates_avg + 5;   // + one per quintile that would have none
  *n_business = spec->total_business;
  return CABS_OK;
}

extern "C" int cabs_graph_build_world(const cabs_graph_world_spec *spec, uint64_t seed, cabs_graph_world *w) {
  if (!spec || !w) return CABS_ERR_INVALID;
  int32_t n, max_h, nb;
  if (cabs_graph_world_sizes(spec, &n, &max_h, &nb) != CABS_OK) return CABS_ERR_INVALID;
  if (!w->px || !w->py || !w->status || !w->severity || !w->infected_time || !w->age || !w->stratum || !w->active ||
      !w->isolated || !w->house || !w->employer || !w->hire_seq || !w->pwealth || !w->hx || !w->hy || !w->hstratum ||
      !w->hsize || !w->hwealth || !w->hfixed || !w->hincomes || !w->hexpenses || !w->bx || !w->by || !w->bstratum ||
      !w->bstocks || !w->bsales || !w->bnum_employees || !w->bprice || !w->bwealth || !w->bfixed || !w->bincomes)
    return CABS_ERR_INVALID;
  Draws rng(seed);
  const int L = spec->length, H = spec->height;
  auto random_x = [&]() { return clip_int(L / 2.0 + rng.normal() * (L / 3.0), 0, L); };   // abs.py:97-101
  auto random_y = [&]() { return clip_int(H / 2.0 + rng.normal() * (H / 3.0), 0, H); };

  w->health_x = random_x(); w->health_y = random_y();                        // graph_abs.py:96-98
  w->health_fixed = 0.0 + spec->minimum_expense * 3; w->health_wealth = 0.0; w->health_expenses = 0.0;
  w->gov_x = random_x(); w->gov_y = random_y();                              // graph_abs.py:99-102
  w->gov_fixed = 0.0 + n * (spec->minimum_expense * 0.05);
  w->gov_price = 1.0;
  w->gov_wealth = spec->total_wealth * spec->public_gdp_share;               // graph_abs.py:126
  w->iteration = -1;

  int nh = n / spec->homemates_avg;                                          // graph_abs.py:105-106
  for (int h = 0; h < nh; ++h) {
    w->hx[h] = random_x(); w->hy[h] = random_y(); w->hstratum[h] = h % 5;
    w->hsize[h] = 0; w->hwealth[h] = 0.0; w->hfixed[h] = 0.0; w->hincomes[h] = 0.0; w->hexpenses[h] = 0.0;
  }
  for (int b = 0; b < nb; ++b) {                                             // graph_abs.py:109-110, network/agents.py:14-31
    w->bx[b] = random_x(); w->by[b] = random_y(); w->bstratum[b] = 5 - (b % 5);
    w->bstocks[b] = 10; w->bsales[b] = 0; w->bnum_employees[b] = 0;
    w->bprice[b] = (w->bstratum[b] + 1) * 12.0; w->bwealth[b] = 0.0; w->bfixed[b] = 0.0; w->bincomes[b] = 0.0;
  }
  // persons, in list order: Infected (infected_time 5), Recovered_Immune, Susceptible (graph_abs.py:113-122)
  const int n_inf = (int)(n * spec->initial_infected_perc), n_imm = (int)(n * spec->initial_immune_perc);
  for (int k = 0; k < n; ++k) {
    const int a = (int)(rng.beta25() * 100.0);                               // graph_abs.py:73
    int stratum;
    if (k < n_inf + n_imm) stratum = (int)floor(rng.uniform() * 100.0 / 20.0);   // int(rand * 100 // 20)
    else stratum = (k - n_inf - n_imm) % 5;
    w->status[k] = k < n_inf ? CABS_INFECTED : k < n_inf + n_imm ? CABS_RECOVERED_IMMUNE : CABS_SUSCEPTIBLE;
    w->severity[k] = CABS_ASYMPTOMATIC;
    w->infected_time[k] = k < n_inf ? 5 : 0;
    w->age[k] = (uint8_t)(a > 255 ? 255 : a);
    w->stratum[k] = (uint8_t)stratum;
    w->active[k] = (a > 16 && a <= 65) ? 1 : 0;                              // network/agents.py:270-271
    w->isolated[k] = 0;
    w->px[k] = 0; w->py[k] = 0; w->house[k] = -1; w->employer[k] = -1; w->hire_seq[k] = -1; w->pwealth[k] = 0.0;
  }
  std::vector<std::vector<int>> employees(nb);
  std::vector<int> houses_q;
  const int cap = spec->homemates_avg + spec->homemates_std;
  for (int q = 0; q < 5; ++q) {                                              // graph_abs.py:128-186
    houses_q.clear();
    for (int h = 0; h < nh; ++h)
      if (w->hstratum[h] == q) houses_q.push_back(h);
    if (houses_q.empty()) {
      if (nh >= max_h) return CABS_ERR_CAPACITY;
      w->hx[nh] = random_x(); w->hy[nh] = random_y(); w->hstratum[nh] = q;
      w->hsize[nh] = 0; w->hwealth[nh] = 0.0; w->hfixed[nh] = 0.0; w->hincomes[nh] = 0.0; w->hexpenses[nh] = 0.0;
      houses_q.push_back(nh++);
    }
    const int nhq = (int)houses_q.size();
    double btotal, bqty;
    if (spec->total_business > 5) {
      btotal = spec->lorenz_curve[q] * (spec->total_wealth * spec->business_gdp_share);
      int c = 0;
      for (int b = 0; b < nb; ++b) c += w->bstratum[b] == q;
      bqty = c > 1 ? c : 1;
    } else {
      btotal = spec->total_wealth * spec->business_gdp_share;
      bqty = spec->total_business;
    }
    for (int b = 0; b < nb; ++b)
      if (w->bstratum[b] == q) w->bwealth[b] = btotal / bqty;
    const double ptotal = spec->lorenz_curve[q] * spec->total_wealth *
                          (1 - (spec->public_gdp_share + spec->business_gdp_share));
    long long pq = 0;
    for (int k = 0; k < n; ++k) pq += (w->stratum[k] == q && w->active[k]);
    const double share = ptotal / (double)(pq > 1 ? pq : 1);
    for (int k = 0; k < n; ++k) {
      if (w->stratum[k] != q) continue;
      if (w->active[k]) {
        w->pwealth[k] = share;
        if (rng.uniform() >= spec->unemployment_rate) {
          // Business.hire (network/agents.py:36-42): agent.expenses is still 0 here, fixed_expenses do not grow
          const int b = rng.randint(spec->total_business);
          employees[b].push_back(k);
          w->employer[k] = b;
          w->bnum_employees[b] += 1;
        }
      }
      const double expenses = spec->basic_income[w->stratum[k] < 5 ? w->stratum[k] : 4] * spec->minimum_expense;
      const double homeless_test = rng.uniform();
      if (!(q == 0 && homeless_test <= spec->homeless_rate)) {
        auto append_mate = [&](int h) {                                      // network/agents.py:186-194
          w->hwealth[h] += w->pwealth[k];
          w->hsize[h] += 1;
          w->house[k] = h;
          const double dx = rng.normal() * 0.25, dy = rng.normal() * 0.25;
          w->px[k] = (int)((double)w->hx[h] + dx);
          w->py[k] = (int)((double)w->hy[h] + dy);
          w->hfixed[h] += (expenses / 720) * 24;
        };
        for (int kp = 0; kp < 5; ++kp) {                                     // 177-182: `continue`, not `break`
          const int h = houses_q[rng.randint(nhq)];
          if (w->hsize[h] < cap) append_mate(h);
        }
        if (w->house[k] < 0) append_mate(rng.randint(nhq));                  // 183-186: local index into the global list
      }
    }
  }
  int seq = 0;
  for (int b = 0; b < nb; ++b)
    for (int k : employees[b]) w->hire_seq[k] = seq++;
  if (spec->isolation_rate >= 0.0)                                           // run_batch.py:31-35
    for (int k = 0; k < n; ++k) w->isolated[k] = rng.uniform() <= spec->isolation_rate ? 1 : 0;
  w->n_persons = n;
  w->n_houses = nh;
  w->n_business = nb;
  return CABS_OK;
}
