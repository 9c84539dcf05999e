"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`), in launch order
of first appearance:  python tools/launch_summary.py profiles/r01_launches.csv"""
import csv
import sys


def main(path):
    rows = [ln for ln in open(path) if ln.startswith('"')]
    tot, cnt, order = {}, {}, []
    for r in csv.DictReader(rows):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        if name not in tot:
            tot[name], cnt[name] = 0.0, 0
            order.append(name)
        tot[name] += ns
        cnt[name] += 1
    total = sum(tot.values())
    for name in order:
        print("%-70s   n=%4d  %10.3f ms  %5.1f%%" % (name[:70], cnt[name], tot[name] / 1e6, 100.0 * tot[name] / total))
    print("total %.3f ms" % (total / 1e6))


if __name__ == "__main__":
    main(sys.argv[1])
