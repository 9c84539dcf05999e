// The Girvan-Newman half of the reference's netclu_ng.py (:101-117, split_until_max_k), natively, with the SAME result
// as the script gives under networkx 3.x on CPython 3.12 — including the places where that result depends on iteration
// order (which of several edges of equal betweenness is removed first).  Used by netclu_cc -g.
//
// What the script does for a connected component `coco` in which two genes of one genome lack an edge:
//     snet  = pnet.subgraph(coco)                       a view of the ROOT graph, filtered by the node set
//     coms  = next(girvan_newman(snet))                 remove max-edge-betweenness edges until snet falls in two
//     for com in sorted coms: collision -> split_until_max_k(com, snet), else a family
// and what decides ties inside networkx (all restated below, each with the place it comes from):
//   * edge_betweenness_centrality returns a dict in G.edges() order; max(d, key=d.get) keeps the FIRST maximum;
//   * G.edges() order = node order of g x adjacency order of g, g = snet.copy().to_undirected();
//   * node order of a subgraph view (coreviews.FilterAtlas.__iter__): the node SET's own iteration order when
//     2*len(set) < len(root graph), else the root graph's order filtered;
//   * the node set is set(nbunch_iter(nodes)): built by adding the elements one at a time — at the top level in the
//     iteration order of the set `coco` that connected_components/_plain_bfs built in BFS discovery order, below that
//     in ascending order (the script passes sorted lists) — so CPython's set layout matters (PyIntSet);
//   * betweenness values are float sums whose order of additions follows the BFS order (Brandes), then multiplied by
//     1/(n(n-1)); they are computed here in the same order with IEEE doubles (compile with -ffp-contract=off).
// Only the resulting partition reaches the .clus, so nothing else of the script's stdout is reproduced.
#pragma once

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_set>
#include <utility>
#include <vector>

namespace pd_host {

// A CPython 3.12 `set` holding non-negative ints < 2^61-1 (hash(i) == i), insert only: enough of
// Objects/setobject.c (set_add_entry, set_table_resize, set_insert_clean) to reproduce its iteration order.
class PyIntSet {
    static constexpr size_t kLinearProbes = 9;  // LINEAR_PROBES
    static constexpr int kPerturbShift = 5;     // PERTURB_SHIFT
    std::vector<int64_t> slot_;                 // -1 = unused
    size_t mask_ = 7, used_ = 0;                // PySet_MINSIZE = 8; no deletions, so fill == used

    void insert_clean(std::vector<int64_t>& t, size_t mask, int64_t key) const {
        size_t perturb = static_cast<size_t>(key), i = static_cast<size_t>(key) & mask;
        for (;;) {
            if (t[i] < 0) {
                t[i] = key;
                return;
            }
            if (i + kLinearProbes <= mask)
                for (size_t j = 1; j <= kLinearProbes; j++)
                    if (t[i + j] < 0) {
                        t[i + j] = key;
                        return;
                    }
            perturb >>= kPerturbShift;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }
    void resize(size_t minused) {
        size_t newsize = 8;
        while (newsize <= minused) newsize <<= 1;
        std::vector<int64_t> t(newsize, -1);
        for (int64_t k : slot_)
            if (k >= 0) insert_clean(t, newsize - 1, k);  // old table order
        slot_.swap(t);
        mask_ = newsize - 1;
    }

public:
    PyIntSet() : slot_(8, -1) {}
    size_t size() const { return used_; }
    bool contains(uint32_t key) const {
        size_t perturb = key, i = key & mask_;
        for (;;) {
            const size_t probes = (i + kLinearProbes <= mask_) ? kLinearProbes : 0;
            for (size_t j = 0; j <= probes; j++) {
                if (slot_[i + j] < 0) return false;
                if (slot_[i + j] == static_cast<int64_t>(key)) return true;
            }
            perturb >>= kPerturbShift;
            i = (i * 5 + 1 + perturb) & mask_;
        }
    }
    bool add(uint32_t key) {
        size_t perturb = key, i = key & mask_;
        for (;;) {
            const size_t probes = (i + kLinearProbes <= mask_) ? kLinearProbes : 0;
            for (size_t j = 0; j <= probes; j++) {
                int64_t& e = slot_[i + j];
                if (e == static_cast<int64_t>(key)) return false;
                if (e < 0) {
                    e = key;
                    used_++;
                    if (used_ * 5 >= mask_ * 3) resize(used_ > 50000 ? used_ * 2 : used_ * 4);
                    return true;
                }
            }
            perturb >>= kPerturbShift;
            i = (i * 5 + 1 + perturb) & mask_;
        }
    }
    std::vector<uint32_t> order() const {  // iteration order: slot order
        std::vector<uint32_t> out;
        out.reserve(used_);
        for (int64_t k : slot_)
            if (k >= 0) out.push_back(static_cast<uint32_t>(k));
        return out;
    }
};

// Threads that run the same function side by side (the caller is worker 0); started once, woken per call.
class Workers {
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable wake_, done_;
    const std::function<void(unsigned)>* job_ = nullptr;
    uint64_t generation_ = 0;
    unsigned pending_ = 0;
    bool stop_ = false;

    void loop(unsigned id) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(unsigned)>* job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                wake_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                job = job_;
            }
            (*job)(id);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_one();
        }
    }

public:
    explicit Workers(unsigned n) {
        for (unsigned i = 1; i < n; i++) threads_.emplace_back([this, i] { loop(i); });
    }
    ~Workers() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        wake_.notify_all();
        for (auto& t : threads_) t.join();
    }
    unsigned size() const { return static_cast<unsigned>(threads_.size()) + 1; }
    void run(const std::function<void(unsigned)>& job) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &job;
            pending_ = static_cast<unsigned>(threads_.size());
            generation_++;
        }
        wake_.notify_all();
        job(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }
};

// The root graph `pnet` restricted to what the split needs: its node count, every node's position in its node order
// (first appearance in the .net, netclu_ng.py:43-56) and the insertion-ordered adjacency of the nodes handed to split().
struct RootGraph {
    size_t n_nodes = 0;                           // len(pnet)
    std::vector<uint32_t> pos;                    // gene id -> position in pnet's node order
    std::vector<uint32_t> dense;                  // gene id -> index into adj (UINT32_MAX: not kept)
    std::vector<std::vector<uint32_t>> adj;       // neighbours (gene ids) in the order add_edge first saw them
    const std::vector<uint32_t>& nbrs(uint32_t s) const { return adj[dense[s]]; }
};

class GirvanNewman {
    struct Nbr {
        uint32_t node, edge;
    };
    const RootGraph& root_;
    std::function<bool(const std::vector<uint32_t>&)> collision_;  // get_max_collision(com, .) > 0, com ascending

    // --- working graph g (local ids = position in g's node order)
    std::vector<std::vector<Nbr>> adj_;
    std::vector<std::pair<uint32_t, uint32_t>> edge_;
    // Brandes scratch: one per worker, and the per-source contributions of a block of sources
    struct Scratch {
        std::vector<double> sigma, delta;
        std::vector<int32_t> dist;
        std::vector<uint32_t> order;
    };
    std::vector<Scratch> scratch_;
    std::vector<double> bet_, block_;
    Workers workers_;
    int trace_ = 0;
    uint64_t par_min_ = 1u << 22;  // nodes x edges from which the blocks pay for their barriers (PD_NETCLU_PAR_MIN)

    // g = snet.copy().to_undirected(): Graph.copy / to_undirected add the nodes in the view's order, then the edges as
    // `for u in view: for v in view[u]: adj[u][v] = adj[v][u] = data` — a key already present keeps its place.
    void build_copy(const std::vector<uint32_t>& nodes, const PyIntSet& node_set) {
        const size_t n = nodes.size();
        std::vector<uint32_t> local_sorted(nodes);  // gene id -> local id by binary search over a sorted copy
        std::sort(local_sorted.begin(), local_sorted.end());
        std::vector<uint32_t> local_of(n);
        {
            std::vector<std::pair<uint32_t, uint32_t>> t(n);
            for (size_t i = 0; i < n; i++) t[i] = {nodes[i], static_cast<uint32_t>(i)};
            std::sort(t.begin(), t.end());
            for (size_t i = 0; i < n; i++) local_of[i] = t[i].second;
        }
        auto local = [&](uint32_t gene) {
            return local_of[std::lower_bound(local_sorted.begin(), local_sorted.end(), gene) - local_sorted.begin()];
        };
        std::vector<std::vector<uint32_t>> a(n), b(n);
        std::unordered_set<uint64_t> have;
        auto put = [&](std::vector<std::vector<uint32_t>>& g, uint32_t u, uint32_t v) {
            if (have.insert(static_cast<uint64_t>(u) << 32 | v).second) g[u].push_back(v);
        };
        for (uint32_t u = 0; u < n; u++)  // the view's adjacency: the root's order, filtered by the node set
            for (uint32_t w : root_.nbrs(nodes[u]))
                if (node_set.contains(w)) {
                    const uint32_t v = local(w);
                    put(a, u, v);
                    put(a, v, u);
                }
        have.clear();
        for (uint32_t u = 0; u < n; u++)  // .to_undirected() of the copy
            for (uint32_t v : a[u]) {
                put(b, u, v);
                put(b, v, u);
            }
        adj_.assign(n, {});
        edge_.clear();
        for (uint32_t u = 0; u < n; u++)
            for (uint32_t v : b[u])
                if (v > u) edge_.push_back({u, v});  // no self loops in pnet (netclu_ng.py:54)
        std::vector<std::pair<uint64_t, uint32_t>> ids(edge_.size());
        for (size_t e = 0; e < edge_.size(); e++) ids[e] = {static_cast<uint64_t>(edge_[e].first) << 32 | edge_[e].second, static_cast<uint32_t>(e)};
        std::sort(ids.begin(), ids.end());
        for (uint32_t u = 0; u < n; u++)
            for (uint32_t v : b[u]) {
                const uint64_t key = static_cast<uint64_t>(std::min(u, v)) << 32 | std::max(u, v);
                const uint32_t e = std::lower_bound(ids.begin(), ids.end(), std::make_pair(key, 0u))->second;
                adj_[u].push_back({v, e});
            }
        for (Scratch& sc : scratch_) {
            sc.sigma.assign(n, 0.0);
            sc.delta.assign(n, 0.0);
            sc.dist.assign(n, -1);
            sc.order.reserve(n);
        }
    }

    // One source of nx.edge_betweenness_centrality (betweenness.py: _single_source_shortest_path_basic, then
    // _accumulate_edges): `add(edge, c)` is called once for every edge on a shortest path from s, in the script's order
    // of w (S reversed).  P[w] is not kept: it is the neighbours one level nearer to s, and the order inside P[w] touches
    // no sum (every v of one w adds to a different edge and a different delta[v]).
    template <typename Add>
    void source(Scratch& sc, uint32_t s, Add add) {
        const uint32_t n = static_cast<uint32_t>(adj_.size());
        sc.order.clear();
        for (uint32_t v = 0; v < n; v++) {
            sc.sigma[v] = 0.0;
            sc.dist[v] = -1;
        }
        sc.sigma[s] = 1.0;
        sc.dist[s] = 0;
        sc.order.push_back(s);
        for (size_t head = 0; head < sc.order.size(); head++) {  // the deque and S hold the same sequence
            const uint32_t v = sc.order[head];
            const int32_t dv = sc.dist[v];
            const double sv = sc.sigma[v];
            for (const Nbr& nb : adj_[v]) {
                const uint32_t w = nb.node;
                if (sc.dist[w] < 0) {
                    sc.order.push_back(w);
                    sc.dist[w] = dv + 1;
                }
                if (sc.dist[w] == dv + 1) sc.sigma[w] += sv;
            }
        }
        for (uint32_t v : sc.order) sc.delta[v] = 0.0;
        for (size_t i = sc.order.size(); i-- > 0;) {
            const uint32_t w = sc.order[i];
            const int32_t dw = sc.dist[w];
            const double coeff = (1.0 + sc.delta[w]) / sc.sigma[w];
            for (const Nbr& nb : adj_[w])
                if (sc.dist[nb.node] == dw - 1) {
                    const double c = sc.sigma[nb.node] * coeff;
                    add(nb.edge, c);
                    sc.delta[nb.node] += c;
                }
        }
    }

    // nx.edge_betweenness_centrality(g), _rescale, and max(betweenness, key=betweenness.get): the edge to remove.  Every
    // betweenness[e] is the float sum of its per-source terms IN SOURCE ORDER; an edge gets at most one term per source,
    // so large graphs are done in blocks of sources — terms side by side, then each edge's terms added in order.
    uint32_t most_central_edge() {
        const uint32_t n = static_cast<uint32_t>(adj_.size());
        const size_t m = edge_.size();
        bet_.assign(m, 0.0);
        if (workers_.size() == 1 || static_cast<uint64_t>(n) * m < par_min_) {
            for (uint32_t s = 0; s < n; s++) source(scratch_[0], s, [&](uint32_t e, double c) { bet_[e] += c; });
        } else {
            const unsigned T = workers_.size();
            size_t block = std::max<size_t>(T, std::min<size_t>(512, (size_t(1) << 25) / m));  // <= 256 MB of terms
            block = (block + T - 1) / T * T;
            block_.resize(block * m);
            for (uint32_t s0 = 0; s0 < n; s0 += block) {
                const uint32_t cnt = static_cast<uint32_t>(std::min<size_t>(block, n - s0));
                std::atomic<uint32_t> next{0};
                workers_.run([&](unsigned id) {
                    for (uint32_t i; (i = next.fetch_add(1)) < cnt;) {
                        double* terms = block_.data() + static_cast<size_t>(i) * m;
                        std::fill(terms, terms + m, 0.0);
                        source(scratch_[id], s0 + i, [&](uint32_t e, double c) { terms[e] = c; });
                    }
                });
                workers_.run([&](unsigned id) {
                    const size_t lo = m * id / T, hi = m * (id + 1) / T;
                    for (uint32_t i = 0; i < cnt; i++) {
                        const double* terms = block_.data() + static_cast<size_t>(i) * m;
                        for (size_t e = lo; e < hi; e++) bet_[e] += terms[e];  // + 0.0 where the source gave no term
                    }
                });
            }
        }
        const double scale = 1.0 / (static_cast<double>(n) * static_cast<double>(n - 1));
        uint32_t best = UINT32_MAX;
        double best_v = 0.0;
        for (uint32_t u = 0; u < n; u++)  // G.edges(): (u, v) when v has not been iterated as a node yet
            for (const Nbr& nb : adj_[u])
                if (nb.node > u) {
                    const double b = bet_[nb.edge] * scale;
                    if (best == UINT32_MAX || b > best_v) {
                        best = nb.edge;
                        best_v = b;
                    }
                }
        return best;
    }

    void remove_edge(uint32_t e) {
        const auto [u, v] = edge_[e];
        auto drop = [&](uint32_t a) {
            auto& l = adj_[a];
            l.erase(std::find_if(l.begin(), l.end(), [&](const Nbr& nb) { return nb.edge == e; }));  // dict order is kept
        };
        drop(u);
        drop(v);
    }

    // nodes reachable from local node 0 (the first component connected_components yields)
    std::vector<uint8_t> reach_first() {
        std::vector<uint8_t> seen(adj_.size(), 0);
        std::vector<uint32_t> q{0};
        seen[0] = 1;
        for (size_t h = 0; h < q.size(); h++)
            for (const Nbr& nb : adj_[q[h]])
                if (!seen[nb.node]) {
                    seen[nb.node] = 1;
                    q.push_back(nb.node);
                }
        return seen;
    }

public:
    size_t removed_edges = 0, splits = 0;

    // threads: 0 = PD_NETCLU_THREADS or the hardware's (at most 32); the result does not depend on it.  PD_NETCLU_TRACE=1 prints the script's own
    // `gn ([..], [..])` line (netclu_ng.py:107) for every split on stderr, in the script's order; =2 also `rm a b` for
    // every removed edge (a, b as networkx' edge tuple has them).
    GirvanNewman(const RootGraph& root, std::function<bool(const std::vector<uint32_t>&)> collision, unsigned threads = 0)
        : root_(root), collision_(std::move(collision)), workers_(pick_threads(threads)) {
        scratch_.resize(workers_.size());
        const char* t = getenv("PD_NETCLU_TRACE");
        trace_ = t ? atoi(t) : 0;
        if (const char* pm = getenv("PD_NETCLU_PAR_MIN")) par_min_ = strtoull(pm, nullptr, 10);
    }
    static unsigned pick_threads(unsigned asked) {
        if (!asked) {
            const char* e = getenv("PD_NETCLU_THREADS");
            asked = e ? static_cast<unsigned>(atoi(e)) : std::thread::hardware_concurrency();
        }
        return std::max(1u, std::min(asked, 32u));
    }
    unsigned threads() const { return workers_.size(); }

    // connected_components' set for the component of `source`: _plain_bfs adds in discovery order (connected.py)
    PyIntSet component_set(uint32_t source) const {
        PyIntSet seen;
        seen.add(source);
        std::vector<uint32_t> level{source}, next;
        while (!level.empty()) {
            next.clear();
            for (uint32_t v : level)
                for (uint32_t w : root_.nbrs(v))
                    if (seen.add(w)) next.push_back(w);
            level.swap(next);
        }
        return seen;
    }

    // split_until_max_k(coco, pnet) for the component whose first node (pnet order) is `source`; the families (each in
    // ascending id) are appended to `families`.
    void split(uint32_t source, std::vector<std::vector<uint32_t>>* families) {
        std::vector<std::vector<uint32_t>> work;  // node lists in the order they are handed to .subgraph()
        work.push_back(component_set(source).order());
        while (!work.empty()) {
            std::vector<uint32_t> given = std::move(work.back());
            work.pop_back();
            PyIntSet node_set;  // filters.show_nodes: set(nbunch_iter(nodes))
            for (uint32_t s : given) node_set.add(s);
            std::vector<uint32_t> nodes;
            if (2 * node_set.size() < root_.n_nodes) {
                nodes = node_set.order();
            } else {
                nodes = given;
                std::sort(nodes.begin(), nodes.end(), [&](uint32_t a, uint32_t b) { return root_.pos[a] < root_.pos[b]; });
            }
            build_copy(nodes, node_set);
            std::vector<uint8_t> first;
            for (;;) {  // _without_most_central_edges: g is connected, so one more component ends it
                const uint32_t e = most_central_edge();
                if (trace_ >= 2) fprintf(stderr, "rm %u %u\n", nodes[edge_[e].first], nodes[edge_[e].second]);
                remove_edge(e);
                removed_edges++;
                first = reach_first();
                if (std::find(first.begin(), first.end(), 0) != first.end()) break;
            }
            splits++;
            std::vector<uint32_t> com[2];  // [0]: the component of g's first node, the one connected_components yields first
            for (size_t i = 0; i < nodes.size(); i++) com[first[i] ? 0 : 1].push_back(nodes[i]);
            for (auto& c : com) std::sort(c.begin(), c.end());
            if (trace_) {
                std::string line = "gn (";
                for (int c = 0; c < 2; c++) {
                    line += c ? ", [" : "[";
                    for (size_t i = 0; i < com[c].size(); i++) line += (i ? ", " : "") + std::to_string(com[c][i]);
                    line += "]";
                }
                fprintf(stderr, "%s)\n", line.c_str());
            }
            // the script recurses into the first before it looks at the second: the stack pops the back, so push the
            // second first (only the order of the trace depends on it; families are a set)
            for (int c = 1; c >= 0; c--) {
                if (collision_(com[c])) work.push_back(std::move(com[c]));
                else families->push_back(std::move(com[c]));
            }
        }
    }
};

}  // namespace pd_host
