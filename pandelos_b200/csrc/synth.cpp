// Synthetic pan-genome workload generator (measurement/test input, not part of the scoring path).
//
// Produces family-structured protein sets of the shapes BASELINE.json names: ancestral
// protein families drawn from the UniProt amino-acid background, a core/accessory presence model,
// per-copy substitutions, occasional paralogs and N-terminal truncations.  Output is the packed form the
// C ABI consumes (residue bytes + gene offsets + genome ids), i.e. what PangeneIData holds after parsing a
// `.faa` (reference: ig/infoasys/cli/pangenes/PangeneIData.java:30-75).
//
// Deterministic for a given (params, seed) and independent of the thread count: every genome draws from
// its own counter-seeded stream.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Rng {  // xoshiro256** seeded through splitmix64
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    Rng(uint64_t seed, uint64_t stream) {
        uint64_t x = seed * 0xD1342543DE82EF95ull + stream * 0x2545F4914F6CDD1Dull + 0x1234567ull;
        for (auto& v : s) v = splitmix(x);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return static_cast<uint32_t>((static_cast<unsigned __int128>(next()) * n) >> 64); }
    double normal() {
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    double gamma(double shape) {  // Marsaglia–Tsang, shape >= 1
        double d = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
        for (;;) {
            double x = normal(), v = 1.0 + c * x;
            if (v <= 0) continue;
            v = v * v * v;
            double u = uniform();
            if (std::log(u) < 0.5 * x * x + d - d * v + d * std::log(v)) return d * v;
        }
    }
};

// UniProt background (percent), order ARNDCQEGHILKMFPSTWYV
const char kLetters[21] = "ARNDCQEGHILKMFPSTWYV";
const double kFreq[20] = {8.25, 5.53, 4.06, 5.45, 1.37, 3.93, 6.75, 7.07, 2.27, 5.96,
                          9.66, 5.84, 2.42, 3.86, 4.70, 6.56, 5.34, 1.08, 2.92, 6.87};

struct Sampler {
    uint32_t thr[20];
    Sampler() {
        double tot = 0, acc = 0;
        for (double f : kFreq) tot += f;
        for (int i = 0; i < 20; i++) {
            acc += kFreq[i];
            thr[i] = static_cast<uint32_t>(std::min(4294967295.0, acc / tot * 4294967296.0));
        }
        thr[19] = 0xFFFFFFFFu;
    }
    char draw(Rng& r) const {
        uint32_t u = static_cast<uint32_t>(r.next() >> 32);
        int i = 0;
        while (u > thr[i]) i++;
        return kLetters[i];
    }
};

struct Copy {
    uint32_t family;
    uint32_t trim;     // residues removed from the N-terminus
    uint8_t paralog;   // substitution rate doubled
    uint64_t stream;   // rng stream for the substitutions
};

}  // namespace

extern "C" {

struct pd_synth_params {
    uint32_t genomes;
    uint32_t genes_per_genome;  // target M; genomes end up with ~0.94 M genes (family pool exhausted first)
    double mean_len;            // mean ancestral protein length
    double mu;                  // per-residue substitution rate of a copy
    uint64_t seed;
    double low_complexity;      // fraction of families carrying a poly-residue run (posting-list skew stress)
    uint32_t threads;           // 0 = hardware concurrency
};

struct pd_synth_plan {
    pd_synth_params p;
    std::vector<uint64_t> fam_off;   // ancestral sequences, packed
    std::vector<char> fam_res;
    std::vector<std::vector<Copy>> per_genome;
    std::vector<uint64_t> gene_off;  // S+1
    uint32_t S = 0;
};

pd_synth_plan* pd_synth_plan_new(const pd_synth_params* params) {
    pd_synth_plan* pl = new pd_synth_plan;
    pl->p = *params;
    const pd_synth_params& p = pl->p;
    Sampler smp;
    uint32_t n_fam = std::max<uint32_t>(1, static_cast<uint32_t>(1.3 * p.genes_per_genome));
    std::vector<double> presence(n_fam);
    {
        Rng r(p.seed, 0);
        pl->fam_off.resize(n_fam + 1);
        pl->fam_off[0] = 0;
        std::vector<uint32_t> len(n_fam);
        for (uint32_t f = 0; f < n_fam; f++) {
            double l = r.gamma(2.2) * (p.mean_len / 2.2);
            len[f] = static_cast<uint32_t>(std::max(30.0, std::floor(l)));
            pl->fam_off[f + 1] = pl->fam_off[f] + len[f];
            presence[f] = (r.uniform() < 0.6) ? 0.98 : 0.6 * r.uniform();
        }
        pl->fam_res.resize(pl->fam_off[n_fam]);
        for (uint32_t f = 0; f < n_fam; f++) {
            char* dst = pl->fam_res.data() + pl->fam_off[f];
            for (uint32_t i = 0; i < len[f]; i++) dst[i] = smp.draw(r);
            if (p.low_complexity > 0 && r.uniform() < p.low_complexity) {
                // poly-residue run of 12..60 copies of one letter somewhere inside
                uint32_t run = 12 + r.below(49);
                if (run + 4 < len[f]) {
                    uint32_t at = r.below(len[f] - run);
                    char c = (r.uniform() < 0.5) ? 'A' : 'Q';
                    for (uint32_t i = 0; i < run; i++) dst[at + i] = c;
                }
            }
        }
    }
    pl->per_genome.resize(p.genomes);
    for (uint32_t g = 0; g < p.genomes; g++) {
        Rng r(p.seed, 1 + g);
        std::vector<uint32_t> order(n_fam);
        for (uint32_t f = 0; f < n_fam; f++) order[f] = f;
        for (uint32_t i = n_fam; i > 1; i--) std::swap(order[i - 1], order[r.below(i)]);
        auto& copies = pl->per_genome[g];
        uint64_t stream = (static_cast<uint64_t>(g) + 1) << 32;
        for (uint32_t f : order) {
            if (copies.size() >= p.genes_per_genome) break;
            if (r.uniform() >= presence[f]) continue;
            uint32_t flen = static_cast<uint32_t>(pl->fam_off[f + 1] - pl->fam_off[f]);
            int n_copies = (r.uniform() < 0.03) ? 2 : 1;
            for (int c = 0; c < n_copies && copies.size() < p.genes_per_genome; c++) {
                Copy cp;
                cp.family = f;
                cp.paralog = (c == 1);
                cp.trim = (r.uniform() < 0.2) ? r.below(flen / 10 + 1) : 0;
                cp.stream = stream++;
                copies.push_back(cp);
            }
        }
    }
    uint64_t S = 0;
    for (auto& v : pl->per_genome) S += v.size();
    pl->S = static_cast<uint32_t>(S);
    pl->gene_off.resize(S + 1);
    uint64_t off = 0, i = 0;
    for (auto& v : pl->per_genome)
        for (auto& cp : v) {
            pl->gene_off[i++] = off;
            off += (pl->fam_off[cp.family + 1] - pl->fam_off[cp.family]) - cp.trim;
        }
    pl->gene_off[S] = off;
    return pl;
}

uint32_t pd_synth_genes(const pd_synth_plan* pl) { return pl->S; }
uint64_t pd_synth_residues(const pd_synth_plan* pl) { return pl->gene_off[pl->S]; }

// Fills caller-allocated arrays: residues[total], offsets[S+1], genome_of[S], family_of[S] (family_of may be null).
void pd_synth_fill(const pd_synth_plan* pl, uint8_t* residues, uint64_t* offsets, uint32_t* genome_of, uint32_t* family_of) {
    const pd_synth_params& p = pl->p;
    memcpy(offsets, pl->gene_off.data(), sizeof(uint64_t) * (pl->S + 1));
    std::vector<uint32_t> first(p.genomes + 1, 0);
    for (uint32_t g = 0; g < p.genomes; g++) first[g + 1] = first[g] + static_cast<uint32_t>(pl->per_genome[g].size());
    uint32_t nt = p.threads ? p.threads : std::max(1u, std::thread::hardware_concurrency());
    nt = std::min<uint32_t>(nt, std::max(1u, p.genomes));
    Sampler smp;
    auto work = [&](uint32_t tid) {
        for (uint32_t g = tid; g < p.genomes; g += nt) {
            const auto& copies = pl->per_genome[g];
            for (size_t j = 0; j < copies.size(); j++) {
                const Copy& cp = copies[j];
                uint32_t gene = first[g] + static_cast<uint32_t>(j);
                genome_of[gene] = g;
                if (family_of) family_of[gene] = cp.family;
                const char* src = pl->fam_res.data() + pl->fam_off[cp.family] + cp.trim;
                uint64_t len = pl->gene_off[gene + 1] - pl->gene_off[gene];
                uint8_t* dst = residues + pl->gene_off[gene];
                memcpy(dst, src, len);
                double mu = p.mu * (cp.paralog ? 2.0 : 1.0);
                if (mu <= 0) continue;
                // geometric skipping between substitution sites
                Rng r(p.seed ^ 0xA5A5A5A5ull, cp.stream);
                double inv = 1.0 / std::log(1.0 - std::min(mu, 0.999999));
                double pos = std::floor(std::log(1.0 - r.uniform()) * inv);
                while (pos < static_cast<double>(len)) {
                    uint64_t at = static_cast<uint64_t>(pos);
                    char c;
                    do { c = smp.draw(r); } while (c == static_cast<char>(dst[at]));
                    dst[at] = static_cast<uint8_t>(c);
                    pos += 1.0 + std::floor(std::log(1.0 - r.uniform()) * inv);
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
}

void pd_synth_plan_free(pd_synth_plan* pl) { delete pl; }

}  // extern "C"
