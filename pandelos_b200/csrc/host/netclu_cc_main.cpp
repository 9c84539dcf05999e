// netclu_cc: the connected-component half of the reference's netclu_ng.py, natively.
//
//   netclu_cc <in.faa> <in.net> [-r rest.net | -g [-o out.clus]]
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
// printed as `F{ ` lines too, and the exit status is 0: no Python, no networkx, one pass.  -o writes the .clus that
// pandelos.sh:79 makes of those lines (grep | sed | sort | uniq, byte order) as well.
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
#include "netclu.h"

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

int fail(const char* what, const std::string& arg) {
    fprintf(stderr, "netclu_cc: %s%s\n", what, arg.c_str());
    return 1;
}

}  // namespace

int main(int argc, char** argv) {
    std::string faa_path, net_path, rest_path, clus_path;
    bool native_split = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-r" && i + 1 < argc) rest_path = argv[++i];
        else if (a == "-g") native_split = true;
        else if (a == "-o" && i + 1 < argc) clus_path = argv[++i];
        else if (faa_path.empty()) faa_path = a;
        else if (net_path.empty()) net_path = a;
        else return fail("unexpected argument ", a);
    }
    if (net_path.empty()) {
        fprintf(stderr, "usage: netclu_cc <in.faa> <in.net> [-r rest.net | -g [-o out.clus]]\n");
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

    std::vector<pd_host::NetEdge> net_edges(lines.size());
    for (size_t i = 0; i < lines.size(); i++) net_edges[i] = {lines[i].a, lines[i].b};
    const pd_host::NetcluResult res = pd_host::netclu(names, genome, net_edges, native_split);
    const std::string& out = res.f_lines;
    fwrite(out.data(), 1, out.size(), stdout);
    if (!clus_path.empty()) {
        if (!native_split) return fail("-o needs -g: without the split the families are incomplete", "");
        const std::string text = pd_host::clus_text(out);
        FILE* f = fopen(clus_path.c_str(), "w");
        if (!f || fwrite(text.data(), 1, text.size(), f) != text.size() || fclose(f) != 0) return fail("cannot write ", clus_path);
    }

    // --- the lines of the components left for Girvan-Newman, verbatim and in file order
    if (!rest_path.empty()) {
        FILE* f = fopen(rest_path.c_str(), "w");
        if (!f) return fail("cannot write ", rest_path);
        for (size_t i = 0; i < lines.size(); i++)
            if (!res.left.empty() && res.left[i]) {
                fwrite(lines[i].begin, 1, static_cast<size_t>(lines[i].end - lines[i].begin), f);
                fputc('\n', f);
            }
        if (fclose(f) != 0) return fail("cannot write ", rest_path);
    }
    fprintf(stderr, "netclu_cc: %u sequences, %zu genomes, %zu network lines, %zu components: %zu families as they stand, "
                    "%zu left for the split%s, %zu singletons\n",
            S, genome_ids.size(), lines.size(), res.n_comp, res.n_comp - res.n_split, res.n_split, rest_path.empty() ? "" : " (written to -r)",
            res.n_single);
    if (native_split) {
        fprintf(stderr, "netclu_cc: -g split them into %zu families, removing %zu edges\n", res.n_split_families, res.n_removed);
        return 0;
    }
    return res.n_split ? 3 : 0;
}
