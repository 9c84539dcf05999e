// calculate_k — native drop-in for the reference's calculate_k.py (calculate_k.py:1-30), the step of pandelos.sh:66-68
// that chooses k:   python3 calculate_k.py in.faa > tmp ; k=`grep -E "^k =" tmp`
// The reference walks every residue in a Python loop (minutes at a gigabyte of sequence); this is one pass over the
// mapped file with byte histograms.  Same output, line for line:
//
//     total length  <N>            N   = residues on the odd (0-based) lines, each line stripped     calculate_k.py:8-10
//     alphabet {'M': c, ...}       counts per letter, in order of first appearance (dict order)      calculate_k.py:11-15
//     LG  <log_a N>                a   = number of distinct letters                                  calculate_k.py:19
//     {'M': c, ...}                                                                                  calculate_k.py:21
//     uk =  <log_a N>                                                                                calculate_k.py:28
//     fk =  <log_a N / H>          H   = sum over letters of -log_a(c/size) * (c/size), dict order    calculate_k.py:24-29
//     k =  <floor(fk)>                                                                               calculate_k.py:30
//
// Arithmetic: math.log(x, a) is log(x) / log(a) in doubles (CPython mathmodule.c), c/size is the correctly rounded
// quotient, the sum runs in dict order — all repeated here, so fk carries the same bits and prints with Python's repr.
// Input must be ASCII (the reference counts code points of the decoded text; bytes >= 0x80 are refused here rather than
// counted differently).
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "faa.h"

namespace {

// str.strip() with no argument: ASCII whitespace is " \t\n\r\x0b\x0c" plus the separators \x1c-\x1f
inline bool py_space(unsigned char c) { return c == ' ' || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f); }

// repr() of a one-character str
std::string py_char_repr(unsigned char c) {
    const char quote = c == '\'' ? '"' : '\'';
    std::string s(1, quote);
    if (c == '\\') s += "\\\\";
    else if (c == '\t') s += "\\t";
    else if (c == '\n') s += "\\n";
    else if (c == '\r') s += "\\r";
    else if (c < 0x20 || c == 0x7f) {
        char b[8];
        snprintf(b, sizeof(b), "\\x%02x", c);
        s += b;
    } else s.push_back(static_cast<char>(c));
    s.push_back(quote);
    return s;
}

// repr() of a float: shortest digits that round-trip; fixed notation for 1e-4 <= |x| < 1e16, else d.ddde+XX
std::string py_float_repr(double x) {
    if (std::isnan(x)) return "nan";
    if (std::isinf(x)) return x < 0 ? "-inf" : "inf";
    if (x == 0) return std::signbit(x) ? "-0.0" : "0.0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), std::fabs(x), std::chars_format::scientific);
    const std::string s(buf, r.ptr);
    const size_t e = s.find('e');
    const int exp10 = atoi(s.c_str() + e + 1);
    std::string digits;
    for (size_t i = 0; i < e; i++)
        if (s[i] != '.') digits.push_back(s[i]);
    std::string out = x < 0 ? "-" : "";
    if (exp10 >= -4 && exp10 < 16) {
        if (exp10 >= 0) {
            std::string ip = digits.substr(0, std::min<size_t>(digits.size(), static_cast<size_t>(exp10) + 1));
            while (static_cast<int>(ip.size()) < exp10 + 1) ip.push_back('0');
            const std::string fp = digits.size() > static_cast<size_t>(exp10) + 1 ? digits.substr(static_cast<size_t>(exp10) + 1) : "0";
            out += ip + "." + fp;
        } else {
            out += "0." + std::string(static_cast<size_t>(-exp10 - 1), '0') + digits;
        }
    } else {
        out += digits.substr(0, 1);
        if (digits.size() > 1) out += "." + digits.substr(1);
        char eb[16];
        snprintf(eb, sizeof(eb), "e%c%02d", exp10 < 0 ? '-' : '+', std::abs(exp10));
        out += eb;
    }
    return out;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: calculate_k <input.faa>\n");
        return 2;
    }
    pd_host::MappedFile f(argv[1]);
    if (!f.ok()) {
        fprintf(stderr, "calculate_k: cannot read %s\n", argv[1]);
        return 1;
    }
    uint64_t hist[4][256] = {};
    bool seen[256] = {};
    std::vector<unsigned char> order;  // letters in order of first appearance
    uint64_t total = 0;
    bool non_ascii = false;
    pd_host::for_each_line(f.data(), f.size(), [&](size_t index, const char* b, const char* e) {
        if ((index & 1) == 0) return;  // calculate_k.py:9: sequences are the odd lines
        const unsigned char* p = reinterpret_cast<const unsigned char*>(b);
        const unsigned char* q = reinterpret_cast<const unsigned char*>(e);
        while (p < q && py_space(*p)) p++;
        while (q > p && py_space(q[-1])) q--;
        total += static_cast<uint64_t>(q - p);
        const unsigned char* s = p;
        unsigned unseen = 0;  // does the line hold a letter not met before?  (almost never after the first lines)
        for (; s + 4 <= q; s += 4) {  // four histograms: no store-to-load stalls on runs of one letter
            hist[0][s[0]]++;
            hist[1][s[1]]++;
            hist[2][s[2]]++;
            hist[3][s[3]]++;
            unseen |= (unsigned)!seen[s[0]] | (unsigned)!seen[s[1]] | (unsigned)!seen[s[2]] | (unsigned)!seen[s[3]];
        }
        for (; s < q; s++) {
            hist[0][*s]++;
            unseen |= (unsigned)!seen[*s];
        }
        if (unseen) {  // first appearances, in order
            for (s = p; s < q; s++) {
                if (!seen[*s]) {
                    seen[*s] = true;
                    order.push_back(*s);
                    if (*s >= 0x80) non_ascii = true;
                }
            }
        }
    });
    if (non_ascii) {
        fprintf(stderr, "calculate_k: %s holds bytes above 0x7f; the reference counts decoded characters, refusing to guess\n", argv[1]);
        return 1;
    }
    uint64_t count[256];
    for (int c = 0; c < 256; c++) count[c] = hist[0][c] + hist[1][c] + hist[2][c] + hist[3][c];

    std::string dict = "{";
    for (size_t i = 0; i < order.size(); i++) {
        if (i) dict += ", ";
        dict += py_char_repr(order[i]) + ": " + std::to_string(count[order[i]]);
    }
    dict += "}";
    printf("total length  %llu\n", static_cast<unsigned long long>(total));
    printf("alphabet %s\n", dict.c_str());
    const size_t a = order.size();
    if (total == 0 || a == 0) {  // math.log(0, 0)
        fflush(stdout);
        fprintf(stderr, "ValueError: math domain error\n");
        return 1;
    }
    if (a == 1) {  // log(a) == 0
        fflush(stdout);
        fprintf(stderr, "ZeroDivisionError: float division by zero\n");
        return 1;
    }
    const double log_a = std::log(static_cast<double>(a));
    const double lg = std::log(static_cast<double>(total)) / log_a;
    printf("LG  %s\n", py_float_repr(lg).c_str());
    printf("%s\n", dict.c_str());
    double h = 0.0;  // `k` in calculate_k.py:23-26 before it becomes the result
    const double size = static_cast<double>(total);  // sum(alphabet.values())
    for (unsigned char c : order) {
        const double frac = static_cast<double>(count[c]) / size;
        h += -(std::log(frac) / log_a) * frac;
    }
    printf("uk =  %s\n", py_float_repr(lg).c_str());
    if (h == 0.0) {
        fflush(stdout);
        fprintf(stderr, "ZeroDivisionError: float division by zero\n");
        return 1;
    }
    const double fk = lg / h;
    printf("fk =  %s\n", py_float_repr(fk).c_str());
    printf("k =  %lld\n", static_cast<long long>(std::floor(fk)));
    return 0;
}
