// adaptdl_b200 -- native core of the Pollux genetic search (host code, no CUDA).
//
// The scheduler repeats an NSGA-II search every 60 s over candidates
// ``state[j][n]`` = replicas of job j on node n (J jobs x W = 2N node columns, the
// second half being copies of the autoscaling node template). The numpy version
// (adaptdl_b200/sched/policy/pollux.py + nsga2.py) makes ~25 passes per
// generation over the dense [population, J, W] tensor: at 200 jobs x 64 nodes
// that is 10 MB per pass and 9 s per cycle. But a candidate is almost empty -- a
// job sits on a handful of nodes -- so here a candidate is a list of
// (node, replicas) entries per job, every operator (crossover, mutation, the five
// repair rules, the objectives, duplicate detection) costs O(entries) instead of
// O(J x W), the candidates of a generation are spread over threads, and the
// bookkeeping of NSGA-II (non-dominated sorting, crowding distance, binary
// tournaments, elitist survival) never leaves C++.
//
// Same search as the numpy path, operator by operator (the comments name the
// Python function each block mirrors; capabilities of the reference's
// sched/adaptdl_sched/policy/pollux.py:144-428), but its own random streams:
// one stream per (generation, mating), so a run is reproducible from its seed
// and independent of the number of threads.
//
// Speedups come from a per-job table owned by this object. The search stops
// and reports the (job, nodes, replicas) entries it is missing; the caller
// (Python: the jobs' SpeedupFunction objects) fills them in and resumes.
//
// C ABI, driven through ctypes (adaptdl_b200/_native/host.py).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

// ---------------------------------------------------------------- random streams
struct Rng {
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  Rng(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t x = seed ^ (a * 0xD1342543DE82EF95ull) ^ (b * 0xA0761D6478BD642Full + 0x2545F4914F6CDD1Dull);
    for (int i = 0; i < 4; ++i) s[i] = splitmix(x);
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {                                   // xoshiro256**
    const uint64_t result = rotl(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
  }
  // uniform integer in [0, n), n >= 1 (multiply-shift; the bias of 2^-64 * n is irrelevant here)
  uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0, 1)
  // -ln(u) for u uniform in (0, 1]: exponent from the bit pattern, ln of the mantissa from a
  // degree-6 polynomial in (m - 1) / (m + 1) (absolute error < 1e-7: far below what a
  // sampler can resolve); std::log was the largest single cost of a child
  double neg_log_uniform() {
    const double u = 1.0 - uniform();                 // (0, 1]
    uint64_t bits;
    std::memcpy(&bits, &u, sizeof(bits));
    const int e = (int)((bits >> 52) & 0x7FF) - 1023;
    bits = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m;
    std::memcpy(&m, &bits, sizeof(m));                // [1, 2)
    const double t = (m - 1.0) / (m + 1.0), t2 = t * t;
    const double ln_m = 2.0 * t * (1.0 + t2 * (1.0 / 3 + t2 * (1.0 / 5 + t2 * (1.0 / 7 + t2 * (1.0 / 9 +
                        t2 * (1.0 / 11 + t2 * (1.0 / 13)))))));
    return -((double)e * 0.6931471805599453 + ln_m);
  }
  // failures before the first success of a Bernoulli(p) sequence; inv = -1 / ln(1 - p)
  // (0 when p >= 1: every trial succeeds)
  uint64_t skips_inv(double inv) {
    if (inv <= 0.0) return 0;
    const double k = std::floor(neg_log_uniform() * inv);
    return k > 1e18 ? (uint64_t)1e18 : (uint64_t)k;
  }
  uint64_t skips(double p) { return skips_inv(p >= 1.0 ? 0.0 : -1.0 / std::log1p(-p)); }
};

// ---------------------------------------------------------------- candidates
struct Entry { int32_t col, val; };                   // val replicas on node col

// A stored candidate: per job the entries with val > 0, ascending in col (CSR).
struct Genome {
  std::vector<Entry> e;
  std::vector<int32_t> off;                            // [J + 1]
  double f[2] = {0, 0};
  uint64_t key = 0;
  const Entry* row(int j) const { return e.data() + off[j]; }
  int len(int j) const { return off[j + 1] - off[j]; }
};

// A candidate being worked on: one growable row per job, ascending in col; values may be
// zero while the repair rules run (squeezed out before it is stored).
struct Work {
  std::vector<std::vector<Entry>> rows;
  std::vector<int64_t> used;                           // [W] scratch
  std::vector<uint8_t> seen;                           // [W] scratch
  std::vector<uint8_t> spread;                         // [J] scratch
  std::vector<int> order;
  std::vector<std::vector<std::pair<int, int>>> col_jobs;   // [W] (job, index in its row)
};

// ---------------------------------------------------------------- speedup tables
constexpr int DENSE_NODES = 65, DENSE_REPLICAS = 513;

struct JobTable {
  int tn = 0, tr = 0;                                  // dense part: [tn][tr]
  std::vector<double> dense;                           // < 0: unknown
  std::unordered_map<uint64_t, double> sparse;         // beyond the dense part
  bool get(int n, int r, double* out) const {
    if (n < tn && r < tr) {
      const double v = dense[(size_t)n * tr + r];
      if (v < 0) return false;
      *out = v;
      return true;
    }
    auto it = sparse.find(((uint64_t)(uint32_t)n << 32) | (uint32_t)r);
    if (it == sparse.end()) return false;
    *out = it->second;
    return true;
  }
  void set(int n, int r, double v) {
    if (n < tn && r < tr) dense[(size_t)n * tr + r] = v;
    else sparse[((uint64_t)(uint32_t)n << 32) | (uint32_t)r] = v;
  }
};

// ---------------------------------------------------------------- the search
struct Search {
  int J, W, R;
  std::vector<int64_t> job_res, node_res;              // [J][R], [W][R]
  Genome base;                                         // current allocation
  std::vector<uint8_t> pinned;                         // [J]
  std::vector<int32_t> min_rep, max_rep;               // [J]
  std::vector<int32_t> min_fill, max_fit;              // [J][W]
  std::vector<double> weight;                          // [J] dominant share x W
  double restart_penalty;
  int pop_size, n_gen, threads;
  uint64_t seed;

  std::vector<double> skip_inv;                        // [W + 1]: -1 / ln(1 - 1/m), 0 for m <= 1
  std::vector<std::vector<Entry>> floor_rows;          // per job: entries of min_fill > 0
  std::vector<int> res_used;                           // resources somebody requests
  std::vector<JobTable> tables;

  std::vector<Genome> pop;                             // the population
  std::vector<Genome> kids;                            // offspring waiting for their scores
  bool pending = false;
  int gen = -1;                                        // -1: the initial population is pending
  std::vector<int32_t> miss_job, miss_nodes, miss_rep;
  double t_rank = 0, t_children = 0, t_dedupe = 0, t_missing = 0, t_select = 0;   // seconds (diagnosis)
  static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }

  // ---- conversions
  void to_work(const int32_t* dense, Work& w) const {
    w.rows.resize(J);
    for (int j = 0; j < J; ++j) {
      auto& row = w.rows[j];
      row.clear();
      const int32_t* d = dense + (size_t)j * W;
      for (int c = 0; c < W; ++c)
        if (d[c] > 0) row.push_back({c, d[c]});
    }
  }
  void to_dense(const Genome& g, int32_t* dense) const {
    std::memset(dense, 0, sizeof(int32_t) * (size_t)J * W);
    for (int j = 0; j < J; ++j) {
      const Entry* r = g.row(j);
      for (int i = 0, n = g.len(j); i < n; ++i) dense[(size_t)j * W + r[i].col] = r[i].val;
    }
  }
  // squeeze zeros out, pack, hash
  void store(Work& w, Genome& g) const {
    g.e.clear();
    g.off.resize(J + 1);
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int j = 0; j < J; ++j) {
      g.off[j] = (int32_t)g.e.size();
      for (const Entry& x : w.rows[j]) {
        if (x.val <= 0) continue;
        g.e.push_back(x);
        uint64_t z = ((uint64_t)(uint32_t)j << 40) ^ ((uint64_t)(uint32_t)x.col << 20) ^ (uint64_t)(uint32_t)x.val;
        z *= 0xFF51AFD7ED558CCDull; z ^= z >> 32;
        h = (h ^ z) * 0xC2B2AE3D27D4EB4Full;
        h ^= h >> 29;
      }
    }
    g.off[J] = (int32_t)g.e.size();
    g.key = h;
  }

  static int size_of(const Work& w) {                  // ClusterProblem.cluster_sizes
    int size = 0;
    for (const auto& row : w.rows)
      for (auto it = row.rbegin(); it != row.rend(); ++it)
        if (it->val > 0) { size = std::max(size, it->col + 1); break; }
    return size;
  }
  int size_of(const Genome& g) const {
    int size = 0;
    for (int j = 0; j < J; ++j)
      if (g.len(j)) size = std::max(size, g.row(j)[g.len(j) - 1].col + 1);
    return size;
  }

  // ClusterProblem.repair, one candidate
  void repair(Work& w, Rng& rng) const {
    // 1. non-preemptible jobs that already run keep their placement
    for (int j = 0; j < J; ++j)
      if (pinned[j]) w.rows[j].assign(base.row(j), base.row(j) + base.len(j));
    // 2. a node hosts at most one multi-node job (the first in job order). "Multi-node" is
    //    decided once, before anything is removed.
    w.seen.assign(W, 0);
    w.spread.assign(J, 0);
    for (int j = 0; j < J; ++j) {
      int cnt = 0;
      for (const Entry& x : w.rows[j]) cnt += x.val > 0;
      w.spread[j] = cnt > 1;
    }
    for (int j = 0; j < J; ++j) {
      if (!w.spread[j]) continue;
      for (Entry& x : w.rows[j]) {
        if (x.val <= 0) continue;
        if (w.seen[x.col]) x.val = 0;
        else w.seen[x.col] = 1;
      }
    }
    // 3. no more than max_replicas per job: rows above their cap are trimmed in a random
    //    node order
    for (int j = 0; j < J; ++j) {
      auto& row = w.rows[j];
      int64_t total = 0;
      for (const Entry& x : row) total += x.val;
      if (total <= max_rep[j]) continue;
      w.order.clear();
      for (int i = 0; i < (int)row.size(); ++i)
        if (row[i].val > 0) w.order.push_back(i);
      for (size_t i = w.order.size(); i > 1; --i) std::swap(w.order[i - 1], w.order[rng.below(i)]);
      int64_t left = max_rep[j];
      for (int i : w.order) {
        const int64_t keep = std::min<int64_t>(row[i].val, left);
        row[i].val = (int32_t)keep;
        left -= keep;
      }
    }
    // 4. node capacities, one resource after the other: on an oversubscribed node the jobs
    //    claim the resource in priority (job) order and keep what it can still grant
    bool have_cols = false;
    for (int r : res_used) {
      w.used.assign(W, 0);
      for (int j = 0; j < J; ++j) {
        const int64_t need = job_res[(size_t)j * R + r];
        if (need == 0) continue;
        for (const Entry& x : w.rows[j]) w.used[x.col] += (int64_t)x.val * need;
      }
      for (int c = 0; c < W; ++c) {
        const int64_t cap = node_res[(size_t)c * R + r];
        if (w.used[c] <= cap) continue;
        if (!have_cols) {                              // who sits on which node, in job order
          w.col_jobs.resize(W);
          for (auto& v : w.col_jobs) v.clear();
          for (int j = 0; j < J; ++j)
            for (int i = 0; i < (int)w.rows[j].size(); ++i)
              w.col_jobs[w.rows[j][i].col].push_back({j, i});
          have_cols = true;
        }
        int64_t claimed = 0;                           // min(cumulative claim, capacity)
        for (const auto& ji : w.col_jobs[c]) {
          const int64_t need = job_res[(size_t)ji.first * R + r];
          int32_t& v = w.rows[ji.first][ji.second].val;
          if (need <= 0 || v <= 0) continue;
          const int64_t now = std::min(claimed + (int64_t)v * need, cap);
          const int64_t granted = (now - claimed) / need;
          claimed = now;
          if (granted < v) v = (int32_t)granted;
        }
      }
    }
    // 5. all-or-nothing below min_replicas
    for (int j = 0; j < J; ++j) {
      int64_t total = 0;
      for (const Entry& x : w.rows[j]) total += x.val;
      if (total < min_rep[j]) w.rows[j].clear();
    }
  }

  static void put(std::vector<Entry>& row, int col, int32_t val) {   // insert keeping col order
    auto it = std::lower_bound(row.begin(), row.end(), col, [](const Entry& x, int c) { return x.col < c; });
    if (it != row.end() && it->col == col) it->val = val;
    else row.insert(it, {col, val});
  }

  // ClusterProblem.mutate, one candidate (rows hold val > 0 only on entry)
  void mutate(Work& w, Rng& rng) const {
    const int size = size_of(w);
    // the cluster grows by a geometrically distributed number of nodes (usually one)
    const int limit = (int)std::min<uint64_t>((uint64_t)W, (uint64_t)size + 1 + rng.skips(0.5));
    for (int j = 0; j < J; ++j) {
      auto& row = w.rows[j];
      const int32_t* lo = min_fill.data() + (size_t)j * W;
      const int32_t* hi = max_fit.data() + (size_t)j * W;
      auto draw = [&](int c) {
        const int64_t span = (int64_t)hi[c] - lo[c] + 1;
        return span > 0 ? (int32_t)(lo[c] + (int64_t)rng.below((uint64_t)span)) : lo[c];
      };
      const int nz = (int)row.size();
      // every zero entry left of the limit is re-drawn with probability 1 / (W - nz): the
      // k-th zero of the row is hit, k advancing by geometric skips. Decided against the
      // row as it is now, applied after the non-zero entries had their turn.
      w.order.clear();
      if (W - nz > 0) {
        const double inv_zero = skip_inv[W - nz];
        const int zeros = limit - nz;                  // all entries sit left of size <= limit
        for (uint64_t k = rng.skips_inv(inv_zero); k < (uint64_t)std::max(zeros, 0);
             k += 1 + rng.skips_inv(inv_zero)) {
          int col = (int)k;                            // k-th zero -> its column
          for (const Entry& x : row) { if (x.col <= col) ++col; else break; }
          w.order.push_back(col);
        }
      }
      // every non-zero entry is re-drawn with probability 1 / nz
      if (nz > 0) {
        const double inv_pos = skip_inv[std::min(nz, W)];
        for (uint64_t i = rng.skips_inv(inv_pos); i < (uint64_t)nz; i += 1 + rng.skips_inv(inv_pos))
          row[i].val = draw(row[i].col);
      }
      for (int col : w.order) {
        const int32_t v = draw(col);
        if (v > 0) put(row, col, v);
      }
      apply_floor(row, j);
    }
  }
  // nothing drops below the minimum spread of the job's guaranteed replicas
  void apply_floor(std::vector<Entry>& row, int j) const {
    for (const Entry& f : floor_rows[j]) {
      auto it = std::lower_bound(row.begin(), row.end(), f.col, [](const Entry& x, int c) { return x.col < c; });
      if (it != row.end() && it->col == f.col) { if (it->val < f.val) it->val = f.val; }
      else row.insert(it, f);
    }
  }

  // (nodes, replicas) of job j in a stored candidate
  static void usage(const Genome& g, int j, int* nodes, int* replicas) {
    const Entry* r = g.row(j);
    int reps = 0;
    for (int i = 0, n = g.len(j); i < n; ++i) reps += r[i].val;
    *nodes = g.len(j);
    *replicas = reps;
  }

  // ClusterProblem.evaluate, one candidate (every table entry is known)
  void evaluate(Genome& g) const {
    double total = 0.0;
    for (int j = 0; j < J; ++j) {
      int nodes, reps;
      usage(g, j, &nodes, &reps);
      double sp = 0.0;
      tables[j].get(nodes, reps, &sp);
      double scaled = sp * weight[j];
      // (memcmp's pointers must not be null even for zero bytes: an empty row has no storage)
      const bool moved = g.len(j) != base.len(j) ||
          (g.len(j) != 0 && std::memcmp(g.row(j), base.row(j), sizeof(Entry) * g.len(j)) != 0);
      if (moved) scaled *= 1.0 - restart_penalty;
      total += scaled;
    }
    g.f[0] = -total;
    g.f[1] = (double)size_of(g);
  }

  // fn(index, worker): ``worker`` < threads identifies the calling thread (scratch space)
  template <typename Fn>
  void parallel_for(int count, Fn fn) const {
    const int workers = std::max(1, std::min(threads, count));
    if (workers == 1 || (size_t)J * (size_t)count < 4096) {
      for (int i = 0; i < count; ++i) fn(i, 0);
      return;
    }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    auto body = [&](int worker) {
      for (int i = next.fetch_add(1); i < count; i = next.fetch_add(1)) fn(i, worker);
    };
    for (int t = 1; t < workers; ++t) pool.emplace_back(body, t);
    body(0);
    for (auto& th : pool) th.join();
  }

  // ---- NSGA-II bookkeeping (nsga2.py)
  static void fronts_of(const std::vector<Genome>& P, std::vector<std::vector<int>>& fronts) {
    const int n = (int)P.size();
    fronts.clear();
    std::vector<int> counts(n, 0);
    std::vector<std::vector<int>> dominated(n);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < n; ++k) {
        if (i == k) continue;
        const bool le = P[i].f[0] <= P[k].f[0] && P[i].f[1] <= P[k].f[1];
        const bool lt = P[i].f[0] < P[k].f[0] || P[i].f[1] < P[k].f[1];
        if (le && lt) { dominated[i].push_back(k); ++counts[k]; }
      }
    std::vector<char> left(n, 1);
    int remaining = n;
    while (remaining > 0) {
      std::vector<int> front;
      for (int i = 0; i < n; ++i)
        if (left[i] && counts[i] == 0) front.push_back(i);
      if (front.empty())                                // numerical safety (NaN objectives)
        for (int i = 0; i < n; ++i)
          if (left[i]) front.push_back(i);
      for (int i : front) { left[i] = 0; --remaining; }
      for (int i : front)
        for (int k : dominated[i]) --counts[k];
      fronts.push_back(std::move(front));
    }
  }
  static void crowding(const std::vector<Genome>& P, const std::vector<int>& front, std::vector<double>& dist) {
    const int m = (int)front.size();
    dist.assign(m, 0.0);
    if (m <= 2) { std::fill(dist.begin(), dist.end(), std::numeric_limits<double>::infinity()); return; }
    std::vector<int> order(m);
    for (int k = 0; k < 2; ++k) {
      std::iota(order.begin(), order.end(), 0);
      std::stable_sort(order.begin(), order.end(),
                       [&](int a, int b) { return P[front[a]].f[k] < P[front[b]].f[k]; });
      const double lo = P[front[order[0]]].f[k], hi = P[front[order[m - 1]]].f[k];
      dist[order[0]] = dist[order[m - 1]] = std::numeric_limits<double>::infinity();
      if (hi - lo > 0)
        for (int i = 1; i + 1 < m; ++i)
          dist[order[i]] += (P[front[order[i + 1]]].f[k] - P[front[order[i - 1]]].f[k]) / (hi - lo);
    }
  }

  // one generation's offspring: tournaments, crossover, mutation, repair, duplicate removal
  void breed() {
    const int n = (int)pop.size();
    const double t0 = now();
    std::vector<std::vector<int>> fronts;
    fronts_of(pop, fronts);
    std::vector<int> rank(n, 0);
    std::vector<double> crowd(n, 0.0), dist;
    for (size_t r = 0; r < fronts.size(); ++r) {
      crowding(pop, fronts[r], dist);
      for (size_t i = 0; i < fronts[r].size(); ++i) { rank[fronts[r][i]] = (int)r; crowd[fronts[r][i]] = dist[i]; }
    }
    const int matings = (pop_size + 1) / 2;
    Rng pick(seed, (uint64_t)gen, 0xFFFFFFFFull);
    auto tournament = [&]() {
      const int a = (int)pick.below(n), b = (int)pick.below(n);
      const bool better = rank[a] < rank[b] || (rank[a] == rank[b] && crowd[a] >= crowd[b]);
      return better ? a : b;
    };
    std::vector<int> pa(matings), pb(matings);
    for (int m = 0; m < matings; ++m) pa[m] = tournament();
    for (int m = 0; m < matings; ++m) pb[m] = tournament();
    std::vector<Genome> born((size_t)2 * matings);
    std::vector<Work> works(std::max(1, threads));
    const double t1 = now();
    t_rank += t1 - t0;
    parallel_for(matings, [&](int m, int worker) {
      Work& w = works[worker];
      Rng rng(seed, (uint64_t)gen, (uint64_t)m);
      const Genome& a = pop[pa[m]];
      const Genome& b = pop[pb[m]];
      // ClusterProblem.crossover: one cut over the job axis; each child also inherits a
      // cluster size drawn between the parents' sizes (nodes beyond it emptied)
      const int cut = (int)rng.below((uint64_t)J);
      const int size_a = size_of(a), size_b = size_of(b);
      const int lo = std::min(size_a, size_b), hi = std::max(size_a, size_b);
      for (int k = 0; k < 2; ++k) {
        const Genome& head = k == 0 ? a : b;
        const Genome& tail = k == 0 ? b : a;
        const int size = lo + (int)rng.below((uint64_t)(hi - lo + 1));
        w.rows.resize(J);
        for (int j = 0; j < J; ++j) {
          const Genome& src = j < cut ? head : tail;
          auto& row = w.rows[j];
          row.clear();
          const Entry* r = src.row(j);
          for (int i = 0, len = src.len(j); i < len && r[i].col < size; ++i) row.push_back(r[i]);
        }
        mutate(w, rng);
        repair(w, rng);
        store(w, born[(size_t)2 * m + k]);
      }
    });
    // children that are new (not in the population, not repeated among themselves)
    const double t2 = now();
    t_children += t2 - t1;
    std::unordered_set<uint64_t> seen;
    for (const Genome& g : pop) seen.insert(g.key);
    kids.clear();
    for (Genome& g : born)
      if (seen.insert(g.key).second) kids.push_back(std::move(g));
    t_dedupe += now() - t2;
  }

  void find_missing() {
    miss_job.clear(); miss_nodes.clear(); miss_rep.clear();
    std::unordered_set<uint64_t> asked;
    for (const Genome& g : kids)
      for (int j = 0; j < J; ++j) {
        int nodes, reps;
        usage(g, j, &nodes, &reps);
        double v;
        if (tables[j].get(nodes, reps, &v)) continue;
        const uint64_t key = ((uint64_t)j << 44) ^ ((uint64_t)(uint32_t)nodes << 24) ^ (uint32_t)reps;
        if (!asked.insert(key).second) continue;
        miss_job.push_back(j); miss_nodes.push_back(nodes); miss_rep.push_back(reps);
      }
  }

  // score the pending children, merge, elitist survival
  void select() {
    parallel_for((int)kids.size(), [&](int i, int) { evaluate(kids[i]); });
    for (Genome& g : kids) pop.push_back(std::move(g));
    kids.clear();
    if ((int)pop.size() <= pop_size) return;
    std::vector<std::vector<int>> fronts;
    fronts_of(pop, fronts);
    std::vector<int> keep;
    std::vector<double> dist;
    for (auto& front : fronts) {
      if ((int)(keep.size() + front.size()) <= pop_size) {
        keep.insert(keep.end(), front.begin(), front.end());
      } else {
        crowding(pop, front, dist);
        std::vector<int> order(front.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return dist[a] > dist[b]; });
        for (size_t i = 0; keep.size() < (size_t)pop_size; ++i) keep.push_back(front[order[i]]);
        break;
      }
    }
    std::vector<Genome> survivors;
    survivors.reserve(keep.size());
    for (int i : keep) survivors.push_back(std::move(pop[i]));
    pop.swap(survivors);
  }

  // runs until the search is over (returns 0) or speedups are missing (returns how many)
  int run() {
    for (;;) {
      if (!pending) {
        if (gen >= n_gen) return 0;
        breed();
        pending = true;
      }
      double t0 = now();
      find_missing();
      t_missing += now() - t0;
      if (!miss_job.empty()) return (int)miss_job.size();
      t0 = now();
      select();
      t_select += now() - t0;
      pending = false;
      ++gen;
      if (gen >= n_gen) return 0;
    }
  }
};

}  // namespace

extern "C" {

// Arrays are copied. ``max_fit`` entries must already be capped to something that fits an
// int32. Returns an opaque handle (nullptr on bad sizes).
void* adl_pollux_create(int J, int W, int R, const int64_t* job_res, const int64_t* node_res,
                        const int32_t* base, const uint8_t* pinned, const int32_t* min_rep,
                        const int32_t* max_rep, const int32_t* min_fill, const int32_t* max_fit,
                        const double* weight, double restart_penalty, int pop_size, int n_gen,
                        uint64_t seed, int threads) {
  if (J <= 0 || W <= 0 || R < 0 || pop_size <= 0 || n_gen < 0) return nullptr;
  if (J >= (1 << 20) || W >= (1 << 20)) return nullptr;
  Search* s = new Search();
  s->J = J; s->W = W; s->R = R;
  const size_t G = (size_t)J * W;
  s->job_res.assign(job_res, job_res + (size_t)J * R);
  s->node_res.assign(node_res, node_res + (size_t)W * R);
  s->pinned.assign(pinned, pinned + J);
  s->min_rep.assign(min_rep, min_rep + J);
  s->max_rep.assign(max_rep, max_rep + J);
  s->min_fill.assign(min_fill, min_fill + G);
  s->max_fit.assign(max_fit, max_fit + G);
  s->weight.assign(weight, weight + J);
  s->restart_penalty = restart_penalty;
  s->pop_size = pop_size; s->n_gen = n_gen; s->seed = seed;
  if (threads <= 0) {
    threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if (threads > 16) threads = 16;
  }
  s->threads = threads;
  {
    Work w;
    s->to_work(base, w);
    s->store(w, s->base);
  }
  s->skip_inv.assign((size_t)W + 1, 0.0);
  for (int m = 2; m <= W; ++m) s->skip_inv[m] = -1.0 / std::log1p(-1.0 / m);
  s->floor_rows.resize(J);
  for (int j = 0; j < J; ++j)
    for (int c = 0; c < W; ++c)
      if (s->min_fill[(size_t)j * W + c] > 0) s->floor_rows[j].push_back({c, s->min_fill[(size_t)j * W + c]});
  for (int r = 0; r < R; ++r) {
    bool any = false;
    for (int j = 0; j < J; ++j) any |= s->job_res[(size_t)j * R + r] != 0;
    if (any) s->res_used.push_back(r);
  }
  s->tables.resize(J);
  for (int j = 0; j < J; ++j) {
    // the dense part covers what this job can be given: at most max_replicas, and at most
    // what fits on all the nodes together
    int64_t fit = 0;
    for (int c = 0; c < W; ++c) fit += s->max_fit[(size_t)j * W + c];
    const int64_t cap = std::max<int64_t>(1, std::min<int64_t>(s->max_rep[j], fit));
    JobTable& t = s->tables[j];
    t.tr = (int)std::min<int64_t>(cap + 1, DENSE_REPLICAS);
    t.tn = (int)std::min<int64_t>(std::min<int64_t>(cap, W) + 1, DENSE_NODES);
    t.dense.assign((size_t)t.tn * t.tr, -1.0);
    t.dense[0] = 0.0;                                  // nothing allocated
  }
  return s;
}

void adl_pollux_destroy(void* h) { delete static_cast<Search*>(h); }

// The starting candidates [count][J][W] (repaired and de-duplicated here).
int adl_pollux_seed(void* h, const int32_t* initial, int count) {
  Search* s = static_cast<Search*>(h);
  if (!s || count <= 0) return -1;
  s->pop.clear();
  s->kids.clear();
  std::unordered_set<uint64_t> seen;
  Work w;
  for (int i = 0; i < count; ++i) {
    s->to_work(initial + (size_t)i * s->J * s->W, w);
    Rng rng(s->seed, 0xFFFFFFFEull, (uint64_t)i);
    s->repair(w, rng);
    Genome g;
    s->store(w, g);
    if (seen.insert(g.key).second) s->kids.push_back(std::move(g));
  }
  s->pending = true;
  s->gen = -1;
  return (int)s->kids.size();
}

int adl_pollux_run(void* h) { return static_cast<Search*>(h)->run(); }

// The entries the last adl_pollux_run() stopped for.
int adl_pollux_missing(void* h, int32_t* job, int32_t* nodes, int32_t* replicas, int capacity) {
  Search* s = static_cast<Search*>(h);
  const int count = (int)std::min<size_t>(s->miss_job.size(), (size_t)std::max(capacity, 0));
  if (count == 0) return 0;                      // nothing to copy (and data() may be null)
  std::memcpy(job, s->miss_job.data(), sizeof(int32_t) * count);
  std::memcpy(nodes, s->miss_nodes.data(), sizeof(int32_t) * count);
  std::memcpy(replicas, s->miss_rep.data(), sizeof(int32_t) * count);
  return count;
}

int adl_pollux_fill(void* h, int count, const int32_t* job, const int32_t* nodes, const int32_t* replicas,
                    const double* value) {
  Search* s = static_cast<Search*>(h);
  for (int i = 0; i < count; ++i) {
    if (job[i] < 0 || job[i] >= s->J || nodes[i] < 0 || replicas[i] < 0) return -1;
    // a negative or NaN speedup would read as "unknown" for ever
    const double v = value[i] >= 0 ? value[i] : 0.0;
    s->tables[job[i]].set(nodes[i], replicas[i], v);
  }
  return 0;
}

// seconds spent in: ranking, breeding (parallel part), duplicate removal, table checks, selection
void adl_pollux_timing(void* h, double* out) {
  Search* s = static_cast<Search*>(h);
  out[0] = s->t_rank; out[1] = s->t_children; out[2] = s->t_dedupe; out[3] = s->t_missing; out[4] = s->t_select;
}

int adl_pollux_population(void* h) { return (int)static_cast<Search*>(h)->pop.size(); }
int adl_pollux_generation(void* h) { return static_cast<Search*>(h)->gen; }

// states [n][J][W] and objective values [n][2] of the current population
int adl_pollux_result(void* h, int32_t* states, double* values) {
  Search* s = static_cast<Search*>(h);
  const size_t G = (size_t)s->J * s->W;
  for (size_t i = 0; i < s->pop.size(); ++i) {
    s->to_dense(s->pop[i], states + i * G);
    values[2 * i] = s->pop[i].f[0];
    values[2 * i + 1] = s->pop[i].f[1];
  }
  return (int)s->pop.size();
}

// One dense candidate through the repair rules (tests compare it with the Python rules).
int adl_pollux_repair(void* h, int32_t* state, uint64_t stream) {
  Search* s = static_cast<Search*>(h);
  Rng rng(s->seed, 0xFFFFFFFDull, stream);
  Work w;
  Genome g;
  s->to_work(state, w);
  s->repair(w, rng);
  s->store(w, g);
  s->to_dense(g, state);
  return 0;
}

// One dense candidate through the mutation operator (tests check its statistics).
int adl_pollux_mutate(void* h, int32_t* state, uint64_t stream) {
  Search* s = static_cast<Search*>(h);
  Rng rng(s->seed, 0xFFFFFFFCull, stream);
  Work w;
  Genome g;
  s->to_work(state, w);
  s->mutate(w, rng);
  s->store(w, g);
  s->to_dense(g, state);
  return 0;
}

}  // extern "C"
