"""Generates tests/golden/netclu/* in the build container (needs /root/reference and networkx): what the UNMODIFIED
netclu_ng.py prints for a few (.faa, .net) inputs, digested per connected component, as the pin of the native
connected-component tool (pandelos_b200/csrc/host/netclu_cc_main.cpp).

  <case>.faa / <case>.net   inputs (family5 is the Pangenes golden of make_net_golden.py; the others are seeded random
                            networks over header-only .faa files — netclu_ng.py reads the header lines only; the
                            ties* cases are rings, grids, ladders, barbells ...: many edges of EQUAL betweenness, so
                            the split depends on networkx' iteration orders)
  <case>.json               {"kept": [F-lines of the components printed as they stand], "split": [[ids] of every component
                            that went through girvan_newman], "split_families": [F-lines the split produced],
                            "singletons": [F-lines of genes outside the network],
                            "gn": [the script's own `gn (..)` lines, one per split, in its order],
                            "rest_pipeline_differs": true when splitting the components ALONE (netclu_cc -r, then the
                            script on rest.net) gives other families than the script on the full network}

    python tests/golden/make_netclu_golden.py
    python tests/golden/make_netclu_golden.py --fuzz LO HI     no files: tie_case(seed) for LO <= seed < HI, netclu_cc -g against
                                                               the script (families and every `gn (..)` line); run for seeds
                                                               0..500 in round 2: 10,757 splits, no difference
"""
import json
import os
import random
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "netclu")
SCRIPT = "/root/reference/netclu_ng.py"


def digest(stdout):
    lines = stdout.splitlines()
    # the per-component section starts after the second line of 40 dashes
    start = [i for i, ln in enumerate(lines) if ln == "-" * 40][1] + 1
    kept, split, split_families, singletons = [], [], [], []
    i = start
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("coco "):
            ids = json.loads(ln[5:])
            is_split = len(lines[i + 1].split()) == 2      # "max_k len" (netclu_ng.py:155) against "len" (:120)
            if is_split:
                split.append(ids)
            i += 1
            while i < len(lines) and lines[i] != "-" * 10 and not lines[i].endswith(" }"):
                if lines[i].startswith("F{ "):
                    (split_families if is_split else kept).append(lines[i])
                i += 1
            continue
        if ln.startswith("F{ ") and ln.endswith(" }"):
            singletons.append(ln)
        i += 1
    return {"kept": sorted(kept), "split": sorted(split), "split_families": sorted(split_families), "singletons": sorted(singletons),
            "gn": [ln for ln in lines if ln.startswith("gn (")]}


def tie_case(name, seed):
    """Components with many edges of equal betweenness, ids shuffled over a gene range of varying size (so that a
    component is sometimes more, sometimes less than half of the network: the two node orders of a networkx view)."""
    rng = random.Random(seed)
    S = rng.choice([40, 120, 400, 3000])
    G = rng.choice([2, 3, 5, 8])
    genome = [rng.randrange(G) for _ in range(S)]
    ids = list(range(S))
    rng.shuffle(ids)
    edges = []
    used = 0
    for _ in range(rng.choice([1, 1, 2, 5, 20])):
        size = rng.choice([4, 6, 9, 12, 16, 25, 40]) if S >= 120 else rng.choice([4, 6, 9, 12])
        if used + size > S:
            break
        nodes = ids[used:used + size]
        used += size
        kind = rng.choice(["ring", "grid", "barbell", "sparse", "dense", "path", "star2", "ladder"])
        e = []
        h = size // 2
        if kind == "ring":
            e = [(nodes[i], nodes[(i + 1) % size]) for i in range(size)]
        elif kind == "path":
            e = [(nodes[i], nodes[i + 1]) for i in range(size - 1)]
        elif kind == "grid":
            w = max(2, int(size ** 0.5))
            for i in range(size):
                if (i + 1) % w and i + 1 < size:
                    e.append((nodes[i], nodes[i + 1]))
                if i + w < size:
                    e.append((nodes[i], nodes[i + w]))
        elif kind == "barbell":
            for part in (nodes[:h], nodes[h:]):
                e += [(part[x], part[y]) for x in range(len(part)) for y in range(x + 1, len(part))]
            e.append((nodes[0], nodes[h]))
            if rng.random() < 0.5:
                e.append((nodes[1], nodes[h + 1]))
        elif kind == "ladder":
            for i in range(h - 1):
                e += [(nodes[i], nodes[i + 1]), (nodes[h + i], nodes[h + i + 1])]
            e += [(nodes[i], nodes[h + i]) for i in range(h)]
        elif kind == "star2":
            e += [(nodes[0], nodes[i]) for i in range(1, h)]
            e += [(nodes[h], nodes[i]) for i in range(h + 1, size)]
            e.append((nodes[0], nodes[h]))
        else:
            p = 0.15 if kind == "sparse" else 0.6
            e += [(nodes[i], nodes[rng.randrange(i)]) for i in range(1, size)]
            e += [(nodes[x], nodes[y]) for x in range(size) for y in range(x + 1, size) if rng.random() < p]
        edges += e
    rng.shuffle(edges)
    with open(os.path.join(OUT, name + ".faa"), "w") as f:
        for s in range(S):
            f.write("G%d\tn%d\tdesc\nA\n" % (genome[s], s))
    with open(os.path.join(OUT, name + ".net"), "w") as f:
        for a, b in edges:
            if rng.random() < 0.5:
                a, b = b, a
            f.write("%d\t%d\t%r\n" % (a, b, round(rng.random(), 3)))
            if rng.random() < 0.1:
                f.write("%d\t%d\t0.5\n" % (b, a))
            if rng.random() < 0.03:
                f.write("%d\t%d\t1.0\n" % (a, a))


def clus(f_lines):
    return sorted(set(ln.replace("F{ ", "").replace("}", "").replace(" ;", "").strip() for ln in f_lines))


def rest_pipeline_differs(name, d):
    """netclu_cc -r + the script on rest.net, against the script on the full network (digest d)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from pandelos_b200 import build
    build.build_host()
    faa, net, rest = (os.path.join(OUT, name + e) for e in (".faa", ".net", ".rest.tmp"))
    r = subprocess.run([build.NETCLU_BIN, faa, net, "-r", rest], capture_output=True, text=True)
    lines = r.stdout.splitlines()
    if r.returncode == 3:
        s = subprocess.run([sys.executable, SCRIPT, faa, rest], capture_output=True, text=True, check=True)
        lines += [ln for ln in s.stdout.splitlines() if ln.startswith("F{ ") and not ln.endswith(" }")]
    os.remove(rest)
    return clus(lines) != clus(d["kept"] + d["split_families"] + d["singletons"])


def random_case(name, genomes, genes, seed):
    """Families as near-cliques: some paralog pairs joined (no collision), some not (collision -> split), chains,
    self-loop lines (node only, netclu_ng.py:47-56), repeated lines, genes outside the network."""
    rng = random.Random(seed)
    faa, net = os.path.join(OUT, name + ".faa"), os.path.join(OUT, name + ".net")
    ids = []
    with open(faa, "w") as f:
        for g in range(genomes):
            for n in range(genes):
                ids.append((g, len(ids)))
                f.write("G%d\tg%d_%d@G%d:1\tfam%d\nA\n" % (g, g, n, g, rng.randrange(50)))
    free = [s for _, s in ids]
    rng.shuffle(free)
    genome_of = dict((s, g) for g, s in ids)
    edges = []
    while len(free) > genomes * genes // 6:
        size = rng.choice([1, 2, 2, 3, 4, 5, 6, 8])
        fam, free = free[:size], free[size:]
        if size == 1:
            if rng.random() < 0.5:
                edges.append((fam[0], fam[0]))
            continue
        mode = rng.choice(["clique", "clique", "drop_intra", "chain", "drop_any"])
        for x in range(len(fam)):
            for y in range(x + 1, len(fam)):
                a, b = fam[x], fam[y]
                same = genome_of[a] == genome_of[b]
                if mode == "chain" and y != x + 1:
                    continue
                if mode == "drop_intra" and same:
                    continue
                if mode == "drop_any" and rng.random() < 0.3:
                    continue
                edges.append((a, b) if rng.random() < 0.5 else (b, a))
    rng.shuffle(edges)
    edges += edges[:5]
    with open(net, "w") as f:
        for a, b in edges:
            f.write("%d\t%d\t%r\n" % (a, b, rng.choice([1.0, 0.5, 0.3333333432674408, rng.random()])))


def main():
    os.makedirs(OUT, exist_ok=True)
    for ext in ("faa", "net"):
        shutil.copy(os.path.join(HERE, "net", "family5." + ext), os.path.join(OUT, "family5." + ext))
    random_case("random6x30", 6, 30, 11)
    random_case("random12x25", 12, 25, 12)
    random_case("random3x80", 3, 80, 13)
    ties = {"ties3": 3, "ties21": 21, "ties40": 40, "ties77": 77, "ties105": 105, "ties118": 118}
    for name, seed in ties.items():
        tie_case(name, seed)
    for name in ["family5", "random6x30", "random12x25", "random3x80"] + list(ties):
        r = subprocess.run([sys.executable, SCRIPT, os.path.join(OUT, name + ".faa"), os.path.join(OUT, name + ".net")],
                           capture_output=True, text=True, check=True)
        d = digest(r.stdout)
        if rest_pipeline_differs(name, d):
            d["rest_pipeline_differs"] = True
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(d, f, indent=0)
        print(name, dict((k, len(v) if isinstance(v, list) else v) for k, v in d.items()))


def fuzz(lo, hi):
    import tempfile
    global OUT
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from pandelos_b200 import build
    build.build_host()
    keep, bad, splits = OUT, 0, 0
    with tempfile.TemporaryDirectory() as td:
        OUT = td
        try:
            for seed in range(lo, hi):
                tie_case("c", seed)
                faa, net = os.path.join(td, "c.faa"), os.path.join(td, "c.net")
                a = subprocess.run([sys.executable, SCRIPT, faa, net], capture_output=True, text=True, check=True).stdout.splitlines()
                b = subprocess.run([build.NETCLU_BIN, faa, net, "-g"], capture_output=True, text=True, check=True,
                                   env=dict(os.environ, PD_NETCLU_TRACE="1"))
                gn = [ln for ln in a if ln.startswith("gn (")]
                splits += len(gn)
                if gn != [ln for ln in b.stderr.splitlines() if ln.startswith("gn (")] or \
                        clus([ln for ln in a if "F{ " in ln]) != clus(b.stdout.splitlines()):
                    bad += 1
                    print("MISMATCH seed", seed, flush=True)
        finally:
            OUT = keep
    print("seeds %d..%d: %d mismatches, %d splits" % (lo, hi, bad, splits))


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--fuzz":
        fuzz(int(sys.argv[2]), int(sys.argv[3]))
    else:
        main()
