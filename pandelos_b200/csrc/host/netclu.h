// netclu_ng.py's clustering (netclu_ng.py:64-72,79-126,149-175) over an edge list in memory: connected components, the
// collision test, optionally the Girvan-Newman split (girvan_newman.h), and the `F{ ` lines the script prints.  Shared
// by netclu_cc (edges = the lines of a .net file) and pangenes --clus (edges = the network it has just computed).
#pragma once

#include <algorithm>
#include <cstdint>
#include <numeric>
#include <string>
#include <string_view>
#include <unordered_map>
#include <utility>
#include <vector>

#include "girvan_newman.h"

namespace pd_host {

struct NetEdge {
    uint32_t a, b;  // one .net line; a == b only adds the node (netclu_ng.py:47-56)
};

struct NetcluResult {
    std::string f_lines;             // "F{ a ; b ; c}\n" per family (members in ascending id), then "F{ a }\n" per gene outside the network
    std::vector<uint8_t> left;       // per edge: belongs to a component that needs the split (and native_split was off)
    size_t n_comp = 0, n_split = 0, n_single = 0, n_split_families = 0, n_removed = 0;
};

namespace netclu_detail {
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
}  // namespace netclu_detail

// names / genome: per gene (id = position in the .faa); lines: the network in .net line order (the order decides
// networkx' node and adjacency order, hence ties in the split).  Gene ids in `lines` must be < names.size().
inline NetcluResult netclu(const std::vector<std::string_view>& names, const std::vector<uint32_t>& genome,
                           const std::vector<NetEdge>& lines, bool native_split) {
    const uint32_t S = static_cast<uint32_t>(names.size());
    NetcluResult res;
    std::string& out = res.f_lines;
    // --- connected components; member lists in ascending id, components in order of their smallest member
    netclu_detail::Dsu dsu(S);
    std::vector<uint8_t> in_net(S, 0);
    std::vector<uint64_t> edges;  // (min << 32 | max) of every edge line, sorted: has_edge by binary search
    edges.reserve(lines.size());
    std::vector<uint32_t> node_pos(S, UINT32_MAX);  // position in pnet's node order (first appearance, netclu_ng.py:47-52)
    uint32_t n_nodes = 0;
    for (const NetEdge& l : lines) {
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
    size_t& n_comp = res.n_comp;
    size_t& n_split = res.n_split;
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
    if (native_split && n_split) {
        RootGraph rg;
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
        for (const NetEdge& l : lines) {  // add_edge keeps the place of a neighbour it has seen before (:55-56)
            if (l.a == l.b || !split[dsu.find(l.a)]) continue;
            if (!seen_edge.emplace(static_cast<uint64_t>(std::min(l.a, l.b)) << 32 | std::max(l.a, l.b), 1).second) continue;
            rg.adj[rg.dense[l.a]].push_back(l.b);
            rg.adj[rg.dense[l.b]].push_back(l.a);
        }
        GirvanNewman gn(rg, [&](const std::vector<uint32_t>& com) { return has_collision(com.data(), com.data() + com.size()); });
        std::vector<std::vector<uint32_t>> families;
        // connected_components(pnet) yields the components in the order of their first node (netclu_ng.py:149)
        std::sort(sources.begin(), sources.end(), [&](uint32_t a, uint32_t b) { return node_pos[a] < node_pos[b]; });
        for (uint32_t src : sources) gn.split(src, &families);
        for (const auto& f : families) print_family(f.data(), f.data() + f.size());
        res.n_split_families = families.size();
        res.n_removed = gn.removed_edges;
    }
    for (uint32_t s = 0; s < S; s++)
        if (!in_net[s]) {
            out += "F{ ";
            out.append(names[s]);
            out += " }\n";
            res.n_single++;
        }
    if (!native_split && n_split) {
        res.left.resize(lines.size());
        for (size_t i = 0; i < lines.size(); i++) res.left[i] = split[dsu.find(lines[i].a)];
    }
    return res;
}

// pandelos.sh:79 on those lines: grep "F{ " | sed s/F{\ //g | sed s/}//g | sed s/\ \;//g | sort | uniq, in byte order
inline std::string clus_text(const std::string& f_lines) {
    auto drop_all = [](std::string& s, const char* what) {
        const size_t n = std::char_traits<char>::length(what);
        for (size_t at = 0; (at = s.find(what, at)) != std::string::npos;) s.erase(at, n);
    };
    std::vector<std::string> fams;
    for (size_t at = 0; at < f_lines.size();) {
        size_t nl = f_lines.find('\n', at);
        if (nl == std::string::npos) nl = f_lines.size();
        std::string ln = f_lines.substr(at, nl - at);
        at = nl + 1;
        if (ln.find("F{ ") == std::string::npos) continue;
        drop_all(ln, "F{ ");
        drop_all(ln, "}");
        drop_all(ln, " ;");
        fams.push_back(std::move(ln));
    }
    std::sort(fams.begin(), fams.end());
    fams.erase(std::unique(fams.begin(), fams.end()), fams.end());
    std::string text;
    for (const std::string& f : fams) text += f + "\n";
    return text;
}

}  // namespace pd_host
