// Greedy decoder walks (reference inference.py:70-164, 231-305) -- HOST code.
// The step after the scoring pass: from each of nb_paths sampled start edges (src, dst) the reference walks greedily
// forwards from dst and "backwards" from src (a forward walk from src ^ 1 on the reverse-complement strand, mirrored
// afterwards), always moving to the not-yet-visited successor with the largest log-probability.  Sequential, branchy
// pointer chasing over a few successors per node: no GPU work.  What the reference spends its time on is Python (dict
// and set look-ups, one torch.topk per step); here the successor lists are a CSR, the visited sets a byte map per
// worker thread, and the candidates of one iteration run on a pool of threads (they only read the shared visited map).
#include <algorithm>
#include <cmath>
#include <mutex>
#include <thread>
#include <vector>

#include "gnb_common.cuh"

namespace gnb {
namespace {

// Visited maps are N bytes each and get_contigs_greedy calls the walker once per contig: allocating and zero-filling
// one per worker thread and call would cost ~1 GB of memset per decoder iteration on a 10M-node graph.  A worker
// borrows a map from this pool and returns it all-zero (it clears what it touched), so a map is zero-filled once.
class MarkPool {
 public:
  std::vector<uint8_t> take(size_t n) {
    std::vector<uint8_t> m;
    {
      std::lock_guard<std::mutex> lock(mu_);
      if (!free_.empty()) {
        m = std::move(free_.back());
        free_.pop_back();
      }
    }
    if (m.size() < n) m.assign(n, 0);   // a map of another (smaller) graph: start over
    return m;
  }
  void give(std::vector<uint8_t>&& m) {
    std::lock_guard<std::mutex> lock(mu_);
    if (free_.size() < 64) free_.push_back(std::move(m));
  }
 private:
  std::mutex mu_;
  std::vector<std::vector<uint8_t>> free_;
};
static MarkPool g_marks;

struct Walker {
  const gnb_walk_graph_t& g;
  const float* logp;
  const uint8_t* visited_old;
  std::vector<uint8_t> mark;          // this candidate's visited set (both walks): cleared through `touched`
  std::vector<int32_t> touched;

  Walker(const gnb_walk_graph_t& graph, const float* lp, const uint8_t* vo)
      : g(graph), logp(lp), visited_old(vo), mark(g_marks.take((size_t)graph.num_nodes)) {}
  ~Walker() {
    reset();
    g_marks.give(std::move(mark));
  }
  Walker(const Walker&) = delete;
  Walker& operator=(const Walker&) = delete;

  void set(int32_t v) {
    if (!mark[v]) {
      mark[v] = 1;
      touched.push_back(v);
    }
  }
  bool seen(int32_t v) const { return (visited_old && visited_old[v]) || mark[v]; }
  void reset() {
    for (int32_t v : touched) mark[v] = 0;
    touched.clear();
  }

  // greedy_forwards (inference.py:70-114) with RANDOM = False and early_stopping = False (:25-28); also the body of
  // greedy_backwards_rc (:117-161), which runs the same loop from start ^ 1.  Returns the fp32 running sum of the
  // log-probabilities, accumulated in the reference's order (sumLogProb is a float32 tensor there).
  float walk_from(int32_t start, std::vector<int32_t>& walk) {
    int32_t current = start;
    float sum = 0.f;
    for (;;) {
      walk.push_back(current);
      set(current);
      set(current ^ 1);
      const int64_t b = g.succ_ptr[current], e = g.succ_ptr[current + 1];
      if (e == b) break;
      if (e - b == 1) {
        const int32_t nb = g.succ_node[b];
        if (seen(nb)) break;
        sum += logp[g.succ_edge[b]];
        current = nb;
        continue;
      }
      // torch.topk(neighbor_p, k=1) over the unvisited successors in list order.  Exact ties go to the first one, which
      // is what torch's CPU topk returns for up to 16 candidates (beyond that its choice among equals is unspecified).
      int64_t best = -1;
      float best_p = 0.f;
      for (int64_t q = b; q < e; ++q) {
        if (seen(g.succ_node[q])) continue;
        const float p = logp[g.succ_edge[q]];
        if (best < 0 || p > best_p || (std::isnan(p) && !std::isnan(best_p))) {   // topk ranks nan above everything
          best = q;
          best_p = p;
        }
      }
      if (best < 0) break;
      sum += best_p;
      current = g.succ_node[best];
    }
    return sum;
  }
};

struct Candidate {
  std::vector<int32_t> walk;   // walk_b + walk_f
  int64_t back_len = 0;
  float sum_f = 0.f, sum_b = 0.f;
};

// run_greedy_both_ways (inference.py:164-168)
void run_candidate(Walker& w, int32_t src, int32_t dst, Candidate& out) {
  w.reset();
  for (int32_t v : {src, src ^ 1, dst, dst ^ 1}) w.set(v);        // tmp_visited = visited | {src, src^1, dst, dst^1}
  std::vector<int32_t> fwd, back;
  out.sum_f = w.walk_from(dst, fwd);                              // marks of the forward walk stay set: the backward
  out.sum_b = w.walk_from(src ^ 1, back);                         // walk sees tmp_visited | visited_f
  out.walk.clear();
  out.walk.reserve(fwd.size() + back.size());
  for (auto it = back.rbegin(); it != back.rend(); ++it) out.walk.push_back(*it ^ 1);   // reversed, complemented
  out.back_len = (int64_t)back.size();
  out.walk.insert(out.walk.end(), fwd.begin(), fwd.end());
}

bool graph_ok(const gnb_walk_graph_t* g) {
  return g != nullptr && g->num_nodes >= 0 && g->num_nodes < (1ll << 31) && g->succ_ptr != nullptr &&
         (g->succ_ptr[g->num_nodes] == 0 || (g->succ_node != nullptr && g->succ_edge != nullptr));
}

}  // namespace
}  // namespace gnb

using namespace gnb;

extern "C" int gnb_greedy_walks(const gnb_walk_graph_t* g, const float* log_probs, const uint8_t* visited,
                                int64_t n_cand, const int32_t* cand_src, const int32_t* cand_dst, int threads,
                                int32_t* walk_buf, int64_t walk_cap, int64_t* walk_off, int64_t* back_len,
                                float* sum_logp) {
  GNB_REQUIRE(graph_ok(g), "gnb_greedy_walks: bad graph");
  GNB_REQUIRE(n_cand >= 0 && walk_cap >= 0 && walk_off != nullptr, "gnb_greedy_walks: bad arguments");
  walk_off[0] = 0;
  if (n_cand == 0) return 0;
  GNB_REQUIRE(log_probs && cand_src && cand_dst && back_len && sum_logp && (walk_buf || walk_cap == 0),
              "gnb_greedy_walks: null pointer");
  GNB_REQUIRE(g->num_nodes % 2 == 0, "gnb_greedy_walks: nodes come in strand pairs (2k, 2k+1); N=%lld is odd",
              (long long)g->num_nodes);
  for (int64_t k = 0; k < n_cand; ++k)
    GNB_REQUIRE(cand_src[k] >= 0 && cand_src[k] < g->num_nodes && cand_dst[k] >= 0 && cand_dst[k] < g->num_nodes,
                "gnb_greedy_walks: candidate %lld has an endpoint out of range", (long long)k);
  std::vector<Candidate> res((size_t)n_cand);
  int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (threads <= 0 && T > 16) T = 16;   // the walks are short: more threads only add start-up cost
  if (T < 1) T = 1;
  if ((int64_t)T > n_cand) T = (int)n_cand;
  auto work = [&](int t) {
    Walker w(*g, log_probs, visited);
    for (int64_t k = t; k < n_cand; k += T) run_candidate(w, cand_src[k], cand_dst[k], res[(size_t)k]);
  };
  if (T == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  int64_t total = 0;
  for (int64_t k = 0; k < n_cand; ++k) {
    total += (int64_t)res[(size_t)k].walk.size();
    walk_off[k + 1] = total;
    back_len[k] = res[(size_t)k].back_len;
    sum_logp[2 * k] = res[(size_t)k].sum_f;
    sum_logp[2 * k + 1] = res[(size_t)k].sum_b;
  }
  if (total > walk_cap) {   // walk_off is complete: the caller sizes walk_buf from walk_off[n_cand] and calls again
    set_error("gnb_greedy_walks: walk buffer of %lld entries, %lld needed", (long long)walk_cap, (long long)total);
    return GNB_E_WORKSPACE;
  }
  for (int64_t k = 0; k < n_cand; ++k)
    std::copy(res[(size_t)k].walk.begin(), res[(size_t)k].walk.end(), walk_buf + walk_off[k]);
  return 0;
}

extern "C" int gnb_walk_contig_length(const gnb_walk_graph_t* g, const int64_t* prefix_length, const int64_t* read_length,
                                      const int32_t* walk, int64_t len, int64_t* out) {
  GNB_REQUIRE(graph_ok(g) && prefix_length && read_length && out && (walk || len == 0), "gnb_walk_contig_length: bad arguments");
  GNB_REQUIRE(len >= 1, "gnb_walk_contig_length: empty walk");
  int64_t total = 0;
  for (int64_t i = 0; i + 1 < len; ++i) {
    const int32_t u = walk[i], v = walk[i + 1];
    GNB_REQUIRE(u >= 0 && u < g->num_nodes, "gnb_walk_contig_length: node %d out of range", u);
    int64_t q = g->succ_ptr[u];
    const int64_t e = g->succ_ptr[u + 1];
    while (q < e && g->succ_node[q] != v) ++q;
    GNB_REQUIRE(q < e, "gnb_walk_contig_length: no edge %d -> %d (step %lld of the walk)", u, v, (long long)i);
    total += prefix_length[g->succ_edge[q]];
  }
  GNB_REQUIRE(walk[len - 1] >= 0 && walk[len - 1] < g->num_nodes, "gnb_walk_contig_length: node out of range");
  *out = total + read_length[walk[len - 1]];
  return 0;
}

extern "C" int gnb_walk_jumped_nodes(const gnb_walk_graph_t* succ, const gnb_walk_graph_t* pred, const int32_t* walk,
                                     int64_t len, uint8_t* mark) {
  GNB_REQUIRE(graph_ok(succ) && graph_ok(pred) && succ->num_nodes == pred->num_nodes && mark && (walk || len == 0),
              "gnb_walk_jumped_nodes: bad arguments");
  GNB_REQUIRE(succ->num_nodes % 2 == 0, "gnb_walk_jumped_nodes: nodes come in strand pairs; N is odd");
  std::vector<uint8_t> is_succ((size_t)succ->num_nodes, 0);
  for (int64_t i = 0; i + 1 < len; ++i) {
    const int32_t ss = walk[i], dd = walk[i + 1];
    GNB_REQUIRE(ss >= 0 && ss < succ->num_nodes && dd >= 0 && dd < succ->num_nodes, "gnb_walk_jumped_nodes: node out of range");
    for (int64_t q = succ->succ_ptr[ss]; q < succ->succ_ptr[ss + 1]; ++q) is_succ[succ->succ_node[q]] = 1;
    for (int64_t q = pred->succ_ptr[dd]; q < pred->succ_ptr[dd + 1]; ++q) {
      const int32_t t = pred->succ_node[q];
      if (is_succ[t]) {   // t in succs[ss] & preds[dd]: a read the walk jumps over; it and its complement are used up
        mark[t] = 1;
        mark[t ^ 1] = 1;
      }
    }
    for (int64_t q = succ->succ_ptr[ss]; q < succ->succ_ptr[ss + 1]; ++q) is_succ[succ->succ_node[q]] = 0;
  }
  return 0;
}
