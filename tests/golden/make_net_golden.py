"""Generates tests/golden/net/* in the build container (needs /root/reference):

  family5.faa    a small synthetic input (5 genomes x 60 genes)
  family5.net    Pangenes' network for it: scores from the UNMODIFIED reference library (oracle/_ref, fake JNIEnv),
                 filter + writer from the restated Java host (oracle/pangenes_java.py; no JVM here)
  family5.clus   the reference's own netclu_ng.py run on that .faa/.net, post-processed as pandelos.sh:79 does
                 (grep "F{ " | sed | sort | uniq)

    python tests/golden/make_net_golden.py
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pangenes_java, refjni  # noqa: E402
from pandelos_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "net")
K = 4


def main():
    os.makedirs(OUT, exist_ok=True)
    w = synth.generate(5, 60, 120.0, 0.1, 97)
    faa = os.path.join(OUT, "family5.faa")
    w.write_faa(faa)
    ref = refjni.RefJni()
    ref.preprocess(w.residues, w.offsets, w.genome_of, K)
    net = pangenes_java.run(ref.compute_scores, w.G)
    netf = os.path.join(OUT, "family5.net")
    net.save(netf)
    r = subprocess.run([sys.executable, "/root/reference/netclu_ng.py", faa, netf], capture_output=True, text=True, check=True)
    fams = sorted(set(ln.replace("F{ ", "").replace("}", "").replace(" ;", "").strip() for ln in r.stdout.splitlines() if ln.startswith("F{ ")))
    with open(os.path.join(OUT, "family5.clus"), "w") as f:
        f.write("\n".join(fams) + "\n")
    print("%d genes, %d net lines, %d families" % (w.S, len(net.lines()), len(fams)))


if __name__ == "__main__":
    main()
