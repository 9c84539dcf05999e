"""Generates tests/golden/*.json from the UNMODIFIED reference library (oracle/_ref/libnative_ref.so, built from
/root/reference/ig/native/library.cpp where it lies) driven through the fake JNIEnv of oracle/fakejni.cpp.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
Each file holds the input and, per genome, every field of the reference's Scores object with floats as uint32
bit patterns, cells sorted by (row, column)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fixtures  # noqa: E402
from oracle import refjni  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).tolist()


def dump(name, w, k, ref):
    ref.preprocess(w.residues, w.offsets, w.genome_of, k)
    genomes = []
    for g in range(w.G):
        s = ref.compute_scores(g)
        c = s.canonical()
        genomes.append({
            "scoresCount": int(s.scoresCount),
            "row": c["row"].tolist(), "column": c["column"].tolist(),
            "first_seq_genome": c["first_seq_genome"].tolist(), "second_seq_genome": c["second_seq_genome"].tolist(),
            "scores": bits(c["scores"]), "percs": bits(c["percs"]), "tr_percs": bits(c["tr_percs"]),
            "max_genome_score": [bits(r) for r in s.max_genome_score],
            "max_genome_score_col": bits(s.max_genome_score_col),
            "scoresMaxMappings": s.scoresMaxMappings.tolist(),
        })
    doc = {"name": name, "k": k, "sequences": [w.sequence(i) for i in range(w.S)], "genomes": w.genome_of.tolist(),
           "source": "reference ig/native/library.cpp via oracle/fakejni.cpp", "scores": genomes}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    return sum(g["scoresCount"] for g in genomes)


def main():
    assert refjni.build_ref(), "needs /root/reference to build oracle/_ref"
    ref = refjni.RefJni()
    for name in fixtures.LITERAL:
        if name == "all_short":
            continue  # no k-mers at all: the reference dereferences kmers.begin() of an empty vector (library.cpp:297)
        w, k = fixtures.literal(name)
        print(name, dump(name, w, k, ref))
    for seed in (11, 12, 13):
        w, k = fixtures.random_workload(seed)
        print(w.name, dump(w.name, w, k, ref))
    from pandelos_b200 import synth
    w = synth.generate(4, 60, 90.0, 0.08, 21, name="family_small")
    print(w.name, dump(w.name, w, 4, ref))
    w = synth.generate(3, 50, 80.0, 0.10, 22, low_complexity=0.5, name="family_lowcomplexity")
    print(w.name, dump(w.name, w, 3, ref))


if __name__ == "__main__":
    main()
