// pangenes — native host of the B200 engine with the command line of the reference's Java host
// (reference ig/infoasys/cli/pangenes/Cli.java:13-57: -i/--input, -k/--kvalue, -o/--output required; -j/--threads,
// -c/--complexity, -h/--help), for boxes without a JVM:
//
//     pangenes -i in.faa -k K -o out.net          the line of pandelos.sh:73, one word changed
//
//   .faa reader      two lines per gene, blank lines skipped, header split on TAB, genome ids in first-seen order
//                    (PangeneIData.java:30-75)
//   per genome       pd_genome_edges: computeScores + the BBH / paralog filter of Pangenes.java:98-176 on the GPU
//   network          PangeneNet.addConnection semantics (first score of a (src, dest) wins, PangeneNet.java:49-62);
//                    .net written as saveToFile(file, false) (PangeneNet.java:159-179): one line "src\tdst\tscore"
//                    per undirected edge with src <= dst, score = the float widened to double and printed the way
//                    Java's Double.toString prints it.  Line ORDER: the reference iterates a HashMap; here lines are
//                    sorted by (src, dst) — netclu_ng.py reads the file into a graph, the order carries no meaning.
//   --clus <file>    (not in the reference's CLI) also cluster that network here, as `netclu_ng.py in.faa out.net` +
//                    pandelos.sh:79 would (netclu.h: components, collision test, Girvan-Newman split), and write the
//                    gene families: the network goes from memory to the clustering without being parsed back from text.
//                    Gene names = second header field.
#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "faa.h"
#include "netclu.h"
#include "pandelos_b200.h"

namespace {

// Double.toString(d) for finite d > 0 (JDK 19+: shortest decimal that round-trips; decimal notation for
// 1e-3 <= d < 1e7, otherwise computerized scientific notation)
std::string java_double_to_string(double d) {
    if (d == 0) return "0.0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), d, std::chars_format::scientific);
    std::string s(buf, r.ptr);  // d[.ddd]e[+-]xx, shortest round trip
    const size_t e = s.find('e');
    std::string mant = s.substr(0, e);
    const int exp10 = atoi(s.c_str() + e + 1);
    std::string digits;
    for (char ch : mant)
        if (ch != '.' && ch != '-') digits.push_back(ch);
    std::string out = d < 0 ? "-" : "";
    if (exp10 >= -3 && exp10 < 7) {
        if (exp10 >= 0) {
            std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)exp10 + 1));
            while ((int)ip.size() < exp10 + 1) ip.push_back('0');
            std::string fp = digits.size() > (size_t)exp10 + 1 ? digits.substr((size_t)exp10 + 1) : "0";
            out += ip + "." + fp;
        } else {
            out += "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
        }
    } else {
        out += digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(exp10);
    }
    return out;
}

// PANGENES_TIMING=1: wall clock of the phases on stderr
struct PhaseTimer {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    const bool on = getenv("PANGENES_TIMING") != nullptr;
    void lap(const char* what) {
        const auto n = std::chrono::steady_clock::now();
        if (on) fprintf(stderr, "[pangenes] %-14s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// appends Double.toString(d) at p (at most 32 characters).  Scores lie in (0, 1]: between 1e-3 and 1e7 Java prints the
// shortest round-trip digits in plain decimal notation, which is to_chars' fixed format plus ".0" for whole numbers.
char* append_java_double(char* p, double d) {
    if (d >= 1e-3 && d < 1e7) {
        char* e = std::to_chars(p, p + 32, d, std::chars_format::fixed).ptr;
        bool dot = false;
        for (char* q = p; q < e; q++) dot |= *q == '.';
        if (!dot) {
            *e++ = '.';
            *e++ = '0';
        }
        return e;
    }
    const std::string s = java_double_to_string(d);
    memcpy(p, s.data(), s.size());
    return p + s.size();
}

void usage() {
    printf("usage: pangenes -i <arg> -k <arg> -o <arg> [-j <arg>] [-c] [-h]\n"
           " -c,--complexity     Compute the required number of operations without computing the network (fast)\n"
           " -h,--help           Print this help message\n"
           " -i,--input <arg>    Input file (.faa) to process\n"
           " -j,--threads <arg>  Upper bound on the host threads issuing per-genome calls (1 to 4 are used, by genome count)\n"
           " -k,--kvalue <arg>   Length of the kmers used by the algorithm\n"
           " -o,--output <arg>   Output file for the network\n"
           "    --clus <arg>     Also write the gene families of that network (what netclu_ng.py + pandelos.sh make of it)\n");
}

}  // namespace

int main(int argc, char** argv) {
    std::string in, out, clus;
    int k = 0, threads = 0;
    bool complexity = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto need = [&](const char* what) -> const char* {
            if (i + 1 >= argc) {
                fprintf(stderr, "Missing argument for option: %s\n", what);
                exit(2);
            }
            return argv[++i];
        };
        if (a == "-i" || a == "--input") in = need("i");
        else if (a == "-o" || a == "--output") out = need("o");
        else if (a == "-k" || a == "--kvalue") k = atoi(need("k"));
        else if (a == "-j" || a == "--threads") threads = atoi(need("j"));
        else if (a == "-c" || a == "--complexity") complexity = true;
        else if (a == "--clus") clus = need("clus");
        else if (a == "--selftest-format") {
            // the fast writer against the reference formatter on float32 patterns in (0, 1] and around the notation
            // switches (tests/test_network_filter.py; no GPU involved)
            const long n = atol(need("selftest-format"));
            uint32_t x = 0x9E3779B9u;
            long bad = 0;
            for (long i = 0; i < n; i++) {
                x = x * 1664525u + 1013904223u;
                uint32_t bits = 0x30000000u + (x % (0x3F800000u - 0x30000000u + 1u));  // 4.6e-10 .. 1.0
                if (i % 7 == 0) bits = 0x3A83126Fu + (uint32_t)(i % 5) - 2u;            // around 1e-3
                float f;
                memcpy(&f, &bits, 4);
                char buf[64];
                const std::string fast(buf, append_java_double(buf, (double)f));
                if (fast != java_double_to_string((double)f)) {
                    if (bad++ < 5) fprintf(stderr, "mismatch at %a: %s vs %s\n", (double)f, fast.c_str(), java_double_to_string((double)f).c_str());
                }
            }
            printf("selftest-format: %ld values, %ld mismatches\n", n, bad);
            return bad ? 1 : 0;
        } else if (a == "-h" || a == "--help") {
            usage();
            return 0;
        } else {
            fprintf(stderr, "Unrecognized option: %s\n", a.c_str());
            usage();
            return 2;
        }
    }
    if (in.empty() || out.empty() || k == 0) {
        fprintf(stderr, "Missing required options: i, k, o\n");
        usage();
        return 2;
    }

    PhaseTimer timer;
    // ---- PangeneIData.readFromFile (PangeneIData.java:30-75) over the mapped file: no per-line strings for the sequences
    pd_host::MappedFile f(in);
    if (!f.ok()) {
        fprintf(stderr, "cannot read %s\n", in.c_str());
        return 1;
    }
    std::vector<uint8_t> residues;
    residues.reserve(f.size());
    std::vector<uint64_t> offsets(1, 0);
    std::vector<uint32_t> genome_of;
    std::vector<std::string_view> gene_name;  // second header field (netclu_ng.py:22), views into the mapped file
    std::string_view name_field;
    std::unordered_map<std::string, uint32_t> genome_id;
    std::string genome_name, bad_header;
    bool name_line = true;
    pd_host::for_each_line(f.data(), f.size(), [&](size_t, const char* b, const char* e) {
        while (b < e && (unsigned char)*b <= 0x20) b++;  // String.trim(): code points <= U+0020 off both ends
        while (e > b && (unsigned char)e[-1] <= 0x20) e--;
        if (b == e || !bad_header.empty()) return;  // blank lines are skipped
        if (name_line) {
            const char* tab = static_cast<const char*>(memchr(b, '\t', (size_t)(e - b)));
            // the reference indexes cc[1], cc[2]: a header with fewer than three fields is an error there too
            if (!tab || !memchr(tab + 1, '\t', (size_t)(e - tab - 1))) {
                bad_header.assign(b, e);
                return;
            }
            genome_name.assign(b, tab);
            name_field = std::string_view(tab + 1, (size_t)(static_cast<const char*>(memchr(tab + 1, '\t', (size_t)(e - tab - 1))) - tab - 1));
        } else {
            residues.insert(residues.end(), reinterpret_cast<const uint8_t*>(b), reinterpret_cast<const uint8_t*>(e));
            offsets.push_back(residues.size());
            auto it = genome_id.find(genome_name);
            if (it == genome_id.end()) it = genome_id.emplace(genome_name, (uint32_t)genome_id.size()).first;
            genome_of.push_back(it->second);
            gene_name.push_back(name_field);
        }
        name_line = !name_line;
    });
    if (!bad_header.empty()) {
        fprintf(stderr, "malformed header line (need genome<TAB>gene<TAB>product): %s\n", bad_header.c_str());
        return 1;
    }
    const uint32_t S = (uint32_t)genome_of.size();
    timer.lap("read .faa");

    pd_options opt;
    memset(&opt, 0, sizeof(opt));
    opt.device = -1;
    opt.verbose = 1;
    // every visible GPU gets a replica of the index and a share of the genomes (PD_DEVICES caps it)
    opt.devices = pd_device_count();
    if (const char* e = getenv("PD_DEVICES")) opt.devices = std::max(1, std::min(opt.devices, atoi(e)));
    pd_index* ix = nullptr;
    if (pd_build(residues.data(), offsets.data(), genome_of.data(), S, k, &opt, &ix) != PD_OK) {
        if (k <= 0) {  // library.cpp:90-93
            printf("K value must be greater than 0.\n");
            return 1;
        }
        fprintf(stderr, "pangenes: %s\n", pd_last_error());
        return 1;
    }
    if (complexity) {  // Pangenes.java:33-36
        pd_free(ix);
        return 0;
    }
    pd_index_info info;
    pd_info(ix, &info);
    timer.lap("index build");

    // ---- Pangenes.java:60-183, one task per genome; PangeneNet.addConnection keeps the FIRST score of a directed
    // (src, dest) pair (PangeneNet.java:49-62) and saveToFile(file, false) prints the pairs with src <= dest.  An
    // inter-genome edge is added in both directions at once (Pangenes.java:103-104), so the printed pair (min, max) gets
    // its score from the first task that reports the edge in either direction: edges are collected once, as
    // (min << 32 | max, score) in task order, then stable-sorted and the first of every key kept.  (The reference's
    // HashMap of boxed pairs does not survive 10^8 edges; 16 bytes per reported edge do.)
    struct Edge {
        uint64_t key;
        float score;
    };
    // one task per genome from a small pool of host threads, as Pangenes.java:54-66 (-j threads there; here at most as
    // many as the engine has score contexts: while one call's edges cross PCIe the others' kernels run).  Each task keeps
    // its own edge list; lists are joined in genome order, which is the order the sequential host would report them in.
    std::vector<std::vector<Edge>> per_genome(info.G);
    std::vector<unsigned long long> filtered(info.G, 0);
    const int n_dev = std::max(1, pd_devices(ix));
    std::vector<std::vector<uint32_t>> queue((size_t)n_dev);   // genomes by the device that serves them
    for (uint32_t g = 0; g < info.G; g++) queue[(size_t)std::max(0, pd_genome_device(ix, g))].push_back(g);
    std::vector<std::atomic<uint32_t>> next((size_t)n_dev);
    for (auto& n : next) n.store(0);
    std::atomic<bool> failed(false);
    std::string failure;
    std::mutex failure_mu;
    auto worker = [&](int dev) {
        for (;;) {
            const uint32_t qi = next[(size_t)dev].fetch_add(1);
            if (qi >= queue[(size_t)dev].size() || failed.load()) return;
            const uint32_t g = queue[(size_t)dev][qi];
            pd_edges e;
            if (pd_genome_edges(ix, g, &e) != PD_OK) {
                std::lock_guard<std::mutex> lk(failure_mu);
                if (!failed.exchange(true)) failure = pd_last_error();
                return;
            }
            filtered[g] = e.cells;
            std::vector<Edge>& out_edges = per_genome[g];
            out_edges.reserve(e.count);
            for (uint64_t i = 0; i < e.count; i++) {
                const uint32_t a = e.src[i], b = e.dst[i];
                // an intra-genome edge is added as (src, dest) only (Pangenes.java:171); it is reported with src < dest
                out_edges.push_back(Edge{a <= b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a, e.score[i]});
            }
            pd_edges_release(ix, &e);
        }
    };
    {
        // every extra worker is an extra score context to warm up (pinned and device result buffers, ~0.1 s): worth it
        // only on inputs with hundreds of genomes
        const int wanted = info.G >= 400 ? 4 : (info.G >= 100 ? 2 : 1);
        const unsigned per_dev = (unsigned)std::max(1, std::min(threads > 0 ? std::min(threads, wanted) : wanted, 4));
        std::vector<std::thread> pool;
        for (int d = 0; d < n_dev; d++)
            for (unsigned t = 0; t < per_dev; t++)
                if (d || t) pool.emplace_back(worker, d);
        worker(0);
        for (std::thread& t : pool) t.join();
    }
    if (failed.load()) {
        fprintf(stderr, "pangenes: %s\n", failure.c_str());
        return 1;
    }
    std::vector<Edge> net;
    {
        size_t total = 0;
        for (const auto& v : per_genome) total += v.size();
        net.reserve(total);
    }
    for (uint32_t g = 0; g < info.G; g++) {
        printf("Working on genome %u/%u\n", g, info.G);
        printf("Filtered count: %llu\n", filtered[g]);
        net.insert(net.end(), per_genome[g].begin(), per_genome[g].end());
        std::vector<Edge>().swap(per_genome[g]);
    }
    pd_free(ix);
    timer.lap("scores + filter");
    std::stable_sort(net.begin(), net.end(), [](const Edge& x, const Edge& y) { return x.key < y.key; });
    timer.lap("sort edges");

    // ---- PangeneNet.saveToFile(file, false)
    FILE* o = fopen(out.c_str(), "w");
    if (!o) {
        fprintf(stderr, "cannot write %s\n", out.c_str());
        return 1;
    }
    // lines are formatted into a large buffer with to_chars: per-line fprintf and string building cost more than the
    // whole GPU job on networks of millions of edges
    std::vector<char> obuf((size_t)8 << 20);
    size_t used = 0;
    uint64_t lines = 0;
    for (size_t i = 0; i < net.size(); i++) {
        if (i && net[i].key == net[i - 1].key) continue;  // a later report of the same pair
        if (used + 96 > obuf.size()) {
            fwrite(obuf.data(), 1, used, o);
            used = 0;
        }
        char* p = obuf.data() + used;
        p = std::to_chars(p, p + 12, (uint32_t)(net[i].key >> 32)).ptr;
        *p++ = '\t';
        p = std::to_chars(p, p + 12, (uint32_t)net[i].key).ptr;
        *p++ = '\t';
        p = append_java_double(p, (double)net[i].score);
        *p++ = '\n';
        used = (size_t)(p - obuf.data());
        lines++;
    }
    fwrite(obuf.data(), 1, used, o);
    fclose(o);
    timer.lap("write .net");
    printf("Network: %llu undirected edges written to %s\n", (unsigned long long)lines, out.c_str());

    // ---- --clus: netclu_ng.py on that network, from memory (the .net lines in their written order)
    if (!clus.empty()) {
        std::vector<pd_host::NetEdge> net_lines;
        net_lines.reserve(lines);
        for (size_t i = 0; i < net.size(); i++)
            if (!i || net[i].key != net[i - 1].key) net_lines.push_back({(uint32_t)(net[i].key >> 32), (uint32_t)net[i].key});
        const pd_host::NetcluResult fam = pd_host::netclu(gene_name, genome_of, net_lines, true);
        const std::string text = pd_host::clus_text(fam.f_lines);
        FILE* c = fopen(clus.c_str(), "w");
        if (!c || fwrite(text.data(), 1, text.size(), c) != text.size() || fclose(c) != 0) {
            fprintf(stderr, "cannot write %s\n", clus.c_str());
            return 1;
        }
        timer.lap("families");
        printf("Families: %zu components (%zu split into %zu by Girvan-Newman, %zu edges removed), %zu genes outside the network; written to %s\n",
               fam.n_comp, fam.n_split, fam.n_split_families, fam.n_removed, fam.n_single, clus.c_str());
    }
    return 0;
}
