"""TEST INFRASTRUCTURE — restatement of the reference's Java host for the similarity path, which cannot run here (no JVM):

* ``genome_task``      Pangenes.java:98-176 — inter-genome BBH test, per-row threshold, intra-genome (paralog) test,
                       over one ``Scores`` object (from the unmodified reference library or the C oracle)
* ``PangeneNet``       PangeneNet.java:49-62 (addConnection: first score of a (src, dest) wins) and :159-179
                       (saveToFile(file, false): "src<TAB>dst<TAB>score" for src <= dst)
* ``java_double_to_string``  Double.toString of the float widened to double (JDK 19+: shortest round-trip digits)
* ``read_faa``         PangeneIData.java:30-75

Float semantics follow Java: ``scores`` are float32 and compared with ``==`` / ``>=`` / ``<`` as float32.
Only tests/ and tools that build golden files import this.
"""
import numpy as np


def genome_task(s, genome, G):
    """Returns the addConnection calls of one pool task, in the reference's order: [(src, dst, score_float32), ...]."""
    calls = []
    n = s.scoresCount
    inter_max = np.zeros(G, np.float32)
    should = np.zeros(n, bool)
    mm = s.scoresMaxMappings
    for i in range(n):
        g1, g2 = int(s.first_seq_genome[i]), int(s.second_seq_genome[i])
        if g1 != g2:
            sc = s.scores[i]
            if sc == s.max_genome_score[mm[s.row[i]]][g2] and sc == s.max_genome_score_col[s.column[i]]:
                calls.append((int(s.row[i]), int(s.column[i]), sc))
                calls.append((int(s.column[i]), int(s.row[i]), sc))
                should[i] = True
                if sc < np.float32(1.0) and sc > inter_max[g2]:
                    inter_max[g2] = sc
    S = len(mm)
    thr = np.full(S, np.inf, np.float32)
    for i in range(n):
        if should[i]:
            r = int(s.row[i])
            thr[r] = min(thr[r], inter_max[int(s.second_seq_genome[i])])
    for i in range(n):
        r, c = int(s.row[i]), int(s.column[i])
        g1, g2 = int(s.first_seq_genome[i]), int(s.second_seq_genome[i])
        if r < c and g1 == g2:
            sc = s.scores[i]
            if sc == s.max_genome_score[mm[r]][g2] and sc == s.max_genome_score[mm[c]][g2] and sc >= thr[r]:
                calls.append((r, c, sc))
    return calls


def java_double_to_string(x):
    d = float(x)
    if d == 0.0:
        return "0.0"
    r = repr(abs(d))  # shortest round-trip digits, like JDK 19+
    if "e" in r or "E" in r:
        mant, exp = r.lower().split("e")
        exp10 = int(exp)
    else:
        mant, exp10 = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    digits = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    if ip.strip("0"):
        e10 = exp10 + len(ip.lstrip("0")) - 1
    else:
        e10 = exp10 - (len(fp) - len(fp.lstrip("0"))) - 1
    digits = digits.rstrip("0") or "0"
    sign = "-" if d < 0 else ""
    if -3 <= e10 < 7:
        if e10 >= 0:
            ipart = digits[:e10 + 1].ljust(e10 + 1, "0")
            fpart = digits[e10 + 1:] or "0"
            return sign + ipart + "." + fpart
        return sign + "0." + "0" * (-e10 - 1) + digits
    return sign + digits[0] + "." + (digits[1:] or "0") + "E" + str(e10)


class PangeneNet:
    def __init__(self):
        self.adj = {}

    def add_connection(self, src, dst, score):
        self.adj.setdefault(src, {}).setdefault(dst, score)

    def lines(self):
        out = []
        for src in sorted(self.adj):
            for dst in sorted(self.adj[src]):
                if src <= dst:
                    out.append("%d\t%d\t%s" % (src, dst, java_double_to_string(np.float32(self.adj[src][dst]))))
        return out

    def edge_set(self):
        return {(src, dst, np.float32(sc).view(np.uint32).item()) for src, d in self.adj.items() for dst, sc in d.items() if src <= dst}

    def save(self, path):
        with open(path, "w") as f:
            for ln in self.lines():
                f.write(ln + "\n")


def run(scores_of_genome, G):
    """The whole of Pangenes.main's network construction; scores_of_genome(g) -> Scores."""
    net = PangeneNet()
    for g in range(G):
        for src, dst, sc in genome_task(scores_of_genome(g), g, G):
            net.add_connection(src, dst, sc)
    return net


def read_faa(path):
    """PangeneIData.readFromFile: returns (sequences, genome ids, gene names, genome names)."""
    seqs, gids, names, gname_of = [], [], [], {}
    name_line = True
    genome = gene = None
    with open(path) as f:
        for line in f:
            t = line.strip()
            if not t:
                continue
            if name_line:
                cc = t.split("\t")
                genome, gene = cc[0], cc[1]
            else:
                seqs.append(t)
                names.append(gene)
                gids.append(gname_of.setdefault(genome, len(gname_of)))
            name_line = not name_line
    return seqs, gids, names, list(gname_of)
