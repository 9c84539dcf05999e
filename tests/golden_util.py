"""Loads tests/golden/*.json (outputs of the unmodified reference, see tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np

from pandelos_b200 import native, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json")))


def load(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        doc = json.load(f)
    w = synth.from_sequences(doc["sequences"], doc["genomes"], name=name)
    scores = []
    for g in doc["scores"]:
        def fb(x):
            return np.asarray(x, dtype=np.uint32).view(np.float32)
        rows = len(g["max_genome_score"])
        G = len(doc["scores"])
        scores.append(native.Scores(
            scoresCount=g["scoresCount"], scores=fb(g["scores"]), percs=fb(g["percs"]), tr_percs=fb(g["tr_percs"]),
            row=np.asarray(g["row"], np.int32), column=np.asarray(g["column"], np.int32),
            first_seq_genome=np.asarray(g["first_seq_genome"], np.int32), second_seq_genome=np.asarray(g["second_seq_genome"], np.int32),
            max_genome_score=np.asarray(g["max_genome_score"], dtype=np.uint32).reshape(rows, G).view(np.float32),
            max_genome_score_col=fb(g["max_genome_score_col"]), scoresMaxMappings=np.asarray(g["scoresMaxMappings"], np.int32)))
    return w, doc["k"], scores
