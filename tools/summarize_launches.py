"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for ONE step
(the window between two consecutive loss_partial_kernel launches: backward of step i + forward of step i+1)."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    marks = [i for i, n in enumerate(names) if "loss_partial" in n]
    if len(marks) < 2:
        raise SystemExit("need two loss_partial_kernel launches in the capture, found %d" % len(marks))
    return rows[marks[0]:marks[1]]


def short(name):
    n = re.sub(r"\(.*", "", name).replace("void ", "").replace("fs2::", "")
    return n


def main(path, detail=False):
    rows = load(path)
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for i, r in enumerate(rows):
        n = short(r["Kernel Name"])
        t = float(r["Metric Value"]) / 1e3
        agg[n][0] += 1
        agg[n][1] += t
        tot += t
        if detail:
            print("%4d %-60s %-14s %9.1f" % (i, n[:60], r["Grid Size"], t))
    print("one step: %d launches, %.0f us summed kernel time (ncu: serialised, cold caches)" % (len(rows), tot))
    print("%-72s %5s %10s %6s" % ("kernel", "n", "us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
        print("%-72s %5d %10.1f %5.1f%%" % (n[:72], c, t, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1], detail="--detail" in sys.argv)
