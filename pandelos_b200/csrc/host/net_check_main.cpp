// net_check: the reference's only codified acceptance rule for the similarity stage — the `check` command of its Rust
// benchmark harness (benchmark/test-framework/src/main.rs:129-169, verify.rs:48-86) — restated, so that a maintainer can
// compare the `.net` of this package with the `.net` of the Java/CPU path the way the reference compares its own two
// builds.  (No cargo/rustc in this image: the restatement follows the source and is not pinned on the harness' output.)
//
//   net_check <first.net> <second.net>
//
// Both files are read as sets of UNORDERED gene pairs: each line `a b w` (any whitespace) becomes (min, max, float32 w);
// the lines are stably sorted by pair and of equal pairs the first is kept (verify.rs:80-81); a file that cannot be
// opened is an empty network (verify.rs:50-54).  Printed, as the harness prints them:
//   MissingA a <-> b weight: w     pair of the first file that the second lacks
//   a <-> b = w1 ~ w2              pair of both whose weights differ by more than 0.001 (float32 arithmetic)
//   MissingB a <-> b weight: w     pair of the second file that the first lacks
//   Values <missingA>+<missingB>+(<different weights>) / <pairs of the first file found in the second>
// Exit status (not in the harness, which only prints): 0 when the three counts are zero, 2 otherwise, 1 on a malformed line.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "faa.h"

namespace {

struct Point {
    uint64_t first, second;
    float value;
};

bool pair_less(const Point& a, const Point& b) { return a.first != b.first ? a.first < b.first : a.second < b.second; }
bool pair_equal(const Point& a, const Point& b) { return a.first == b.first && a.second == b.second; }

// Rust's Display for f32: the shortest digits that read back as the same float, never in exponent notation
std::string show(float v) {
    char buf[128];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

bool load(const std::string& path, std::vector<Point>* out) {
    pd_host::MappedFile f(path);
    if (!f.ok()) return true;  // an empty network
    bool bad = false;
    pd_host::for_each_line(f.data(), f.size(), [&](size_t, const char* b, const char* e) {
        const char* col[3][2];
        int n = 0;
        auto ws = [](char c) { return c == ' ' || (c >= '\t' && c <= '\r'); };
        for (const char* p = b; p < e && n < 3;) {
            while (p < e && ws(*p)) p++;
            if (p == e) break;
            col[n][0] = p;
            while (p < e && !ws(*p)) p++;
            col[n++][1] = p;
        }
        Point pt{};
        if (n < 3 || std::from_chars(col[0][0], col[0][1], pt.first).ptr != col[0][1] ||
            std::from_chars(col[1][0], col[1][1], pt.second).ptr != col[1][1] ||
            std::from_chars(col[2][0], col[2][1], pt.value).ptr != col[2][1]) {
            bad = true;  // the harness panics (unwrap)
            return;
        }
        if (pt.first > pt.second) std::swap(pt.first, pt.second);
        out->push_back(pt);
    });
    if (bad) return false;
    std::stable_sort(out->begin(), out->end(), pair_less);
    out->erase(std::unique(out->begin(), out->end(), pair_equal), out->end());
    return true;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: net_check <first.net> <second.net>\n");
        return 1;
    }
    std::vector<Point> a, b;
    if (!load(argv[1], &a) || !load(argv[2], &b)) {
        fprintf(stderr, "net_check: malformed line (need: int int float)\n");
        return 1;
    }
    size_t missing_a = 0, missing_b = 0, present = 0, diff_weight = 0;
    std::string tail;
    size_t j = 0;
    for (const Point& x : a) {  // both are sorted by pair: one merge pass, printed in the harness' order (A's lines, then B's)
        while (j < b.size() && pair_less(b[j], x)) {
            tail += "MissingB " + std::to_string(b[j].first) + " <-> " + std::to_string(b[j].second) + " weight: " + show(b[j].value) + "\n";
            missing_b++;
            j++;
        }
        if (j < b.size() && pair_equal(b[j], x)) {
            if (std::fabs(x.value - b[j].value) > 0.001f) {
                printf("%llu <-> %llu = %s ~ %s\n", (unsigned long long)x.first, (unsigned long long)x.second, show(x.value).c_str(),
                       show(b[j].value).c_str());
                diff_weight++;
            }
            present++;
            j++;
        } else {
            printf("MissingA %llu <-> %llu weight: %s\n", (unsigned long long)x.first, (unsigned long long)x.second, show(x.value).c_str());
            missing_a++;
        }
    }
    for (; j < b.size(); j++) {
        tail += "MissingB " + std::to_string(b[j].first) + " <-> " + std::to_string(b[j].second) + " weight: " + show(b[j].value) + "\n";
        missing_b++;
    }
    fputs(tail.c_str(), stdout);
    printf("Values %zu+%zu+(%zu) / %zu\n", missing_a, missing_b, diff_weight, present);
    return (missing_a || missing_b || diff_weight) ? 2 : 0;
}
