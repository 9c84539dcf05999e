// netclu_cc: the connected-component half of the reference's netclu_ng.py, natively.
//
//   netclu_cc <in.faa> <in.net> [-r rest.net | -g]
//
// netclu_ng.py:64-72,149-172 takes the connected components of the .net graph; a component in which no two genes of one
// genome lack an edge (get_max_collision == 0, netclu_ng.py:79-96) IS a family and is printed as it stands; only the
// others go through networkx' Girvan-Newman split (:101-117).  This tool does the first part for any graph size
// (union-find over the edge lines, collision test on sorted edge pairs) and prints, on stdout, the same `F{ ` lines
// netclu_ng.py prints for
//   * every component that needs no split:   "F{ a ; b ; c}"     (print_family, :119-126; members in ascending id)
//   * every gene outside the network:        "F{ a }"            (remaining_singletons, :174-175)
// The lines of the components that DO need a split are copied verbatim, in file order, to rest.net (-r); running the
// unmodified netclu_ng.py on (in.faa, rest.net) splits exactly those (node and edge insertion order inside a component
// is what it was in the full file, so networkx' tie-breaking is unchanged).  Of that run keep the `F{ ` lines that do
// not end in " }" (its own remaining-singleton lines, which this tool already printed or which belong to families
// printed here).  INTEGRATION.md shows the pipeline; tests/test_netclu_cc.py checks it against netclu_ng.py's stdout.
// (A component split alone can fall differently from the same component split inside the full network: networkx orders
// the nodes of a subgraph view one way when the component is less than half of the graph and another way otherwise, and
// that order breaks ties between edges of equal betweenness.  The -r pipeline is exact up to such ties; -g is exact.)
//
// With -g the split is done here as well (girvan_newman.h: networkx' Girvan-Newman restated with its iteration orders
// and float arithmetic, so that ties fall as they do in the script run on the FULL network), the families it makes are
// printed as `F{ ` lines too, and the exit status is 0: no Python, no networkx, one pass.
//
// Readers follow the script: header = every even line of the .faa, `strip().split('\t')` -> genome, name
// (netclu_ng.py:17-30); .net line -> int, int, weight; a line with src == dst only adds the node (:43-56).
// Exit status: 0 = all families printed, 3 = some components were left for the split (rest.net written), 1 = error.
#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <string>
#include <string_view>
#include <unordered_map>
#include <utility>
#include <vector>

#include "faa.h"
#include "girvan_newman.h"

namespace {

// Python's str.strip() without arguments, on the ASCII whitespace set
std::string_view strip(const char* b, const char* e) {
    auto ws = [](char c) { return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f); };
    while (b < e && ws(*b)) b++;
    while (e > b && ws(e[-1])) e--;
    return std::string_view(b, static_cast<size_t>(e - b));
}

// the n-th tab-separated column, false when the line has fewer
bool column(std::string_view line, int n, std::string_view* out) {
    size_t pos = 0;
    for (int i = 0;; i++) {
        const size_t tab = line.find('\t', pos);
        if (i == n) {
            *out = line.substr(pos, tab == std::string_view::npos ? std::string_view::npos : tab - pos);
            return true;
        }
        if (tab == std::string_view::npos) return false;
        pos = tab + 1;
    }
}

struct Dsu {
    std::vector<uint32_t> p;
    explicit Dsu(size_t n) : p(n) { std::iota(p.begin(), p.end(), 0u); }
    uint32_t find(uint32_t x) {
        while (p[x] != x) {
            p[x] = p[p[x]];
            x = p[x];
        }
        return x;
    }
    void unite(uint32_t a, uint32_t b) {
        a = find(a);
        b = find(b);
        if (a != b) p[std::max(a, b)] = std::min(a, b);  // the root is the component's smallest id
    }
};

int fail(const char* what, const std::string& arg) {
    fprintf(stderr, "netclu_cc: %s%s\n", what, arg.c_str());
    return 1;
}

}  // namespace

int main(int argc, char** argv) {
    std::string faa_path, net_path, rest_path;
    bool native_split = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-r" && i + 1 < argc) rest_path = argv[++i];
        else if (a == "-g") native_split = true;
        else if (faa_path.empty()) faa_path = a;
        else if (net_path.empty()) net_path = a;
        else return fail("unexpected argument ", a);
    }
    if (net_path.empty()) {
        fprintf(stderr, "usage: netclu_cc <in.faa> <in.net> [-r rest.net | -g]\n");
        return 1;
    }

    // --- sequences: name and genome of every header line (netclu_ng.py:17-30)
    pd_host::MappedFile faa(faa_path);
    if (!faa.ok()) return fail("cannot read ", faa_path);
    std::vector<std::string_view> names;
    std::vector<uint32_t> genome;  // dense id of the genome string
    std::unordered_map<std::string_view, uint32_t> genome_ids;
    bool bad = false;
    pd_host::for_each_line(faa.data(), faa.size(), [&](size_t index, const char* b, const char* e) {
        if (index % 2) return;
        const std::string_view line = strip(b, e);
        std::string_view g, n, d;
        if (!column(line, 0, &g) || !column(line, 1, &n) || !column(line, 2, &d)) {  // the script raises IndexError
            bad = true;
            return;
        }
        names.push_back(n);
        genome.push_back(genome_ids.emplace(g, static_cast<uint32_t>(genome_ids.size())).first->second);
    });
    if (bad) return fail("header line with fewer than 3 tab-separated columns in ", faa_path);
    const uint32_t S = static_cast<uint32_t>(names.size());

    // --- network lines (netclu_ng.py:43-56)
    pd_host::MappedFile net(net_path);
    if (!net.ok()) return fail("cannot read ", net_path);
    struct Line {
        uint32_t a, b;
        const char *begin, *end;
    };
    std::vector<Line> lines;
    pd_host::for_each_line(net.data(), net.size(), [&](size_t, const char* b, const char* e) {
        const std::string_view line = strip(b, e);
        std::string_view ca, cb, cw;
        uint32_t a = 0, c = 0;
        if (!column(line, 0, &ca) || !column(line, 1, &cb) || !column(line, 2, &cw) ||
            std::from_chars(ca.data(), ca.data() + ca.size(), a).ec != std::errc() ||
            std::from_chars(cb.data(), cb.data() + cb.size(), c).ec != std::errc() || a >= S || c >= S) {
            bad = true;
            return;
        }
        lines.push_back({a, c, b, e});
    });
    if (bad) return fail("malformed line or gene id outside the .faa in ", net_path);

    // --- connected components; member lists in ascending id, components in order of their smallest member
    Dsu dsu(S);
    std::vector<uint8_t> in_net(S, 0);
    std::vector<uint64_t> edges;  // (min << 32 | max) of every edge line, sorted: has_edge by binary search
    edges.reserve(lines.size());
    std::vector<uint32_t> node_pos(S, UINT32_MAX);  // position in pnet's node order (first appearance, netclu_ng.py:47-52)
    uint32_t n_nodes = 0;
    for (const Line& l : lines) {
        in_net[l.a] = 1;
        if (node_pos[l.a] == UINT32_MAX) node_pos[l.a] = n_nodes++;
        if (l.a == l.b) continue;
        in_net[l.b] = 1;
        if (node_pos[l.b] == UINT32_MAX) node_pos[l.b] = n_nodes++;
        dsu.unite(l.a, l.b);
        edges.push_back(static_cast<uint64_t>(std::min(l.a, l.b)) << 32 | std::max(l.a, l.b));
    }
    std::sort(edges.begin(), edges.end());
    auto has_edge = [&](uint32_t a, uint32_t b) {
        return std::binary_search(edges.begin(), edges.end(), static_cast<uint64_t>(std::min(a, b)) << 32 | std::max(a, b));
    };
    std::vector<uint32_t> comp_start(S + 1, 0), members;
    for (uint32_t s = 0; s < S; s++)
        if (in_net[s]) comp_start[dsu.find(s) + 1]++;
    for (uint32_t s = 0; s < S; s++) comp_start[s + 1] += comp_start[s];
    members.resize(comp_start[S]);
    {
        std::vector<uint32_t> fill(comp_start.begin(), comp_start.end() - 1);
        for (uint32_t s = 0; s < S; s++)
            if (in_net[s]) members[fill[dsu.find(s)]++] = s;
    }

    // --- per component: does any pair of genes of one genome lack an edge? (get_max_collision > 0, netclu_ng.py:79-96)
    std::vector<uint8_t> split(S, 0);  // indexed by root
    std::string out;
    size_t n_comp = 0, n_split = 0;
    std::vector<std::pair<uint32_t, uint32_t>> by_genome;
    auto has_collision = [&](const uint32_t* first, const uint32_t* last) {
        by_genome.clear();
        for (const uint32_t* m = first; m != last; m++) by_genome.emplace_back(genome[*m], *m);
        std::sort(by_genome.begin(), by_genome.end());
        for (size_t i = 0; i < by_genome.size();) {
            size_t j = i;
            while (j < by_genome.size() && by_genome[j].first == by_genome[i].first) j++;
            for (size_t x = i; x < j; x++)
                for (size_t y = x + 1; y < j; y++)
                    if (!has_edge(by_genome[x].second, by_genome[y].second)) return true;
            i = j;
        }
        return false;
    };
    auto print_family = [&](const uint32_t* first, const uint32_t* last) {  // netclu_ng.py:119-126
        out += "F{ ";
        for (const uint32_t* m = first; m != last; m++) {
            if (m != first) out += " ; ";
            out.append(names[*m]);
        }
        out += "}\n";
    };
    for (uint32_t root = 0; root < S; root++) {
        const uint32_t lo = comp_start[root], hi = comp_start[root + 1];
        if (lo == hi) continue;
        n_comp++;
        const bool collision = has_collision(members.data() + lo, members.data() + hi);
        if (collision) {
            split[root] = 1;
            n_split++;
            continue;
        }
        print_family(members.data() + lo, members.data() + hi);
    }

    // --- -g: the Girvan-Newman split of the others (split_until_max_k, netclu_ng.py:101-117)
    size_t n_split_families = 0, n_removed = 0;
    if (native_split && n_split) {
        pd_host::RootGraph rg;
        rg.n_nodes = n_nodes;
        rg.pos = node_pos;
        rg.dense.assign(S, UINT32_MAX);
        std::vector<uint32_t> sources;  // per split component, its first node in pnet's order
        for (uint32_t root = 0; root < S; root++) {
            if (!split[root]) continue;
            uint32_t src = members[comp_start[root]];
            for (uint32_t i = comp_start[root]; i < comp_start[root + 1]; i++) {
                rg.dense[members[i]] = static_cast<uint32_t>(rg.adj.size());
                rg.adj.emplace_back();
                if (node_pos[members[i]] < node_pos[src]) src = members[i];
            }
            sources.push_back(src);
        }
        std::unordered_map<uint64_t, char> seen_edge;
        for (const Line& l : lines) {  // add_edge keeps the place of a neighbour it has seen before (:55-56)
            if (l.a == l.b || !split[dsu.find(l.a)]) continue;
            if (!seen_edge.emplace(static_cast<uint64_t>(std::min(l.a, l.b)) << 32 | std::max(l.a, l.b), 1).second) continue;
            rg.adj[rg.dense[l.a]].push_back(l.b);
            rg.adj[rg.dense[l.b]].push_back(l.a);
        }
        pd_host::GirvanNewman gn(rg, [&](const std::vector<uint32_t>& com) { return has_collision(com.data(), com.data() + com.size()); });
        std::vector<std::vector<uint32_t>> families;
        // connected_components(pnet) yields the components in the order of their first node (netclu_ng.py:149)
        std::sort(sources.begin(), sources.end(), [&](uint32_t a, uint32_t b) { return node_pos[a] < node_pos[b]; });
        for (uint32_t src : sources) gn.split(src, &families);
        for (const auto& f : families) print_family(f.data(), f.data() + f.size());
        n_split_families = families.size();
        n_removed = gn.removed_edges;
    }
    size_t n_single = 0;
    for (uint32_t s = 0; s < S; s++)
        if (!in_net[s]) {
            out += "F{ ";
            out.append(names[s]);
            out += " }\n";
            n_single++;
        }
    fwrite(out.data(), 1, out.size(), stdout);

    // --- the lines of the components left for Girvan-Newman, verbatim and in file order
    if (!rest_path.empty()) {
        FILE* f = fopen(rest_path.c_str(), "w");
        if (!f) return fail("cannot write ", rest_path);
        for (const Line& l : lines)
            if (split[dsu.find(l.a)]) {
                fwrite(l.begin, 1, static_cast<size_t>(l.end - l.begin), f);
                fputc('\n', f);
            }
        if (fclose(f) != 0) return fail("cannot write ", rest_path);
    }
    fprintf(stderr, "netclu_cc: %u sequences, %zu genomes, %zu network lines, %zu components: %zu families as they stand, "
                    "%zu left for the split%s, %zu singletons\n",
            S, genome_ids.size(), lines.size(), n_comp, n_comp - n_split, n_split, rest_path.empty() ? "" : " (written to -r)", n_single);
    if (native_split) {
        fprintf(stderr, "netclu_cc: -g split them into %zu families, removing %zu edges\n", n_split_families, n_removed);
        return 0;
    }
    return n_split ? 3 : 0;
}
