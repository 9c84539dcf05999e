"""Stress of the per-genome call path (GPU box): several host threads, large result arrays in flight on PCIe, every call
compared with the digest of a lone call.    python tools/stress_calls.py [GENOMES] [QUERY] [THREADS] [ROUNDS]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from pandelos_b200 import native, synth  # noqa: E402
from test_gpu_parity import _scores_digest  # noqa: E402


def main():
    genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 250
    query = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    w = synth.shape("scaleout1000", genomes=genomes)
    k = synth.calculate_k(w)
    pn = native.PangeneNative(k, native.PangeneIData(w.residues, w.offsets, w.genome_of), contexts=threads)
    t = time.time()
    want = [_scores_digest(pn.generateScoresPart(g)) for g in range(query)]
    print("lone calls: %.2f s, cells per genome ~%d" % (time.time() - t, want[0][0]), flush=True)
    bad, errs = [], []

    def work(tid):
        try:
            for rep in range(rounds):
                for g in range(query):
                    gg = (g + tid * 5) % query
                    if _scores_digest(pn.generateScoresPart(gg)) != want[gg]:
                        bad.append((tid, rep, gg))
        except Exception as e:
            errs.append(repr(e))

    t = time.time()
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [x.start() for x in th]
    [x.join() for x in th]
    print("threads=%d rounds=%d: %.2f s, mismatches %d, errors %s" % (threads, rounds, time.time() - t, len(bad), errs[:2]), flush=True)
    pn.close()
    sys.exit(1 if bad or errs else 0)


if __name__ == "__main__":
    main()
