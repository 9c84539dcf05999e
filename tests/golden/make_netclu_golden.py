"""Generates tests/golden/netclu/* in the build container (needs /root/reference and networkx): what the UNMODIFIED
netclu_ng.py prints for a few (.faa, .net) inputs, digested per connected component, as the pin of the native
connected-component tool (pandelos_b200/csrc/host/netclu_cc_main.cpp).

  <case>.faa / <case>.net   inputs (family5 is the Pangenes golden of make_net_golden.py; the others are seeded random
                            networks over header-only .faa files — netclu_ng.py reads the header lines only)
  <case>.json               {"kept": [F-lines of the components printed as they stand], "split": [[ids] of every component
                            that went through girvan_newman], "split_families": [F-lines the split produced],
                            "singletons": [F-lines of genes outside the network]}

    python tests/golden/make_netclu_golden.py
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
    return {"kept": sorted(kept), "split": sorted(split), "split_families": sorted(split_families), "singletons": sorted(singletons)}


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
    for name in ("family5", "random6x30", "random12x25", "random3x80"):
        r = subprocess.run([sys.executable, SCRIPT, os.path.join(OUT, name + ".faa"), os.path.join(OUT, name + ".net")],
                           capture_output=True, text=True, check=True)
        d = digest(r.stdout)
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(d, f, indent=0)
        print(name, dict((k, len(v)) for k, v in d.items()))


if __name__ == "__main__":
    main()
