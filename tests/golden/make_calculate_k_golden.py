"""Golden vectors for the native calculate_k: small .faa inputs and the stdout of the UNMODIFIED reference script on them.

    python tests/golden/make_calculate_k_golden.py          (needs /root/reference; run in the build container)

Writes tests/golden/calculate_k/<name>.faa and <name>.out.  The inputs cover the reader's corner cases: CRLF and lone-CR
line ends, no final newline, padded sequence lines, a blank line that shifts the odd/even roles, letters outside the
twenty amino acids, a quote character in the alphabet."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "calculate_k")
REF = "/root/reference/calculate_k.py"
AA = "ACDEFGHIKLMNPQRSTVWY"


def proteins(seed, n, mean_len, letters=AA):
    rng = np.random.default_rng(seed)
    p = rng.dirichlet(np.ones(len(letters)) * 3.0)
    out = []
    for i in range(n):
        ln = max(1, int(rng.poisson(mean_len)))
        out.append("".join(rng.choice(list(letters), size=ln, p=p)))
    return out


def faa(seqs, eol="\n", final_eol=True, pad=("", "")):
    lines = []
    for i, s in enumerate(seqs):
        lines.append("genome%d\tgene%d\tproduct %d" % (i % 3, i, i))
        lines.append(pad[0] + s + pad[1])
    text = eol.join(lines)
    return text + (eol if final_eol else "")


CASES = {
    "plain": faa(proteins(1, 40, 120)),
    "crlf": faa(proteins(2, 25, 80), eol="\r\n"),
    "lone_cr": faa(proteins(3, 25, 80), eol="\r"),
    "no_final_newline": faa(proteins(4, 30, 60), final_eol=False),
    "padded_lines": faa(proteins(5, 30, 60), pad=("  ", " \t")),
    "extra_letters": faa(proteins(6, 30, 90, letters=AA + "XBZU*")),
    "two_letters": faa(proteins(7, 20, 50, letters="AC")),
    "quote_in_alphabet": faa(proteins(8, 20, 40, letters="ACDE'\\")),
    "blank_line_shifts_roles": "g0\tx\tp\nMKV\n\nMKKA\ng1\ty\tp\nAAV\n",
    "long": faa(proteins(9, 400, 330)),
}


def main():
    if not os.path.exists(REF):
        sys.exit("reference script not found: " + REF)
    os.makedirs(OUT, exist_ok=True)
    for name, text in sorted(CASES.items()):
        path = os.path.join(OUT, name + ".faa")
        with open(path, "w", newline="") as f:
            f.write(text)
        r = subprocess.run([sys.executable, REF, path], capture_output=True, text=True)
        if r.returncode != 0:
            sys.exit("%s: reference failed: %s" % (name, r.stderr))
        with open(os.path.join(OUT, name + ".out"), "w") as f:
            f.write(r.stdout)
        print(name, r.stdout.strip().splitlines()[-1])


if __name__ == "__main__":
    main()
