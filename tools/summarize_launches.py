"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel for ONE step
(the window between two consecutive encoder posenc_add launches)."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    marks = [i for i, n in enumerate(names) if "posenc_add" in n]
    a, b = marks[0], marks[1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows[a:b]:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("fs2::", "")
        t = float(r["Metric Value"]) / 1e3
        agg[n][0] += 1
        agg[n][1] += t
        tot += t
    print("one step: %d launches, %.0f us summed kernel time (ncu: serialised, cold caches)" % (b - a, tot))
    print("%-72s %5s %10s %6s" % ("kernel", "n", "us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
        print("%-72s %5d %10.1f %5.1f%%" % (n[:72], c, t, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1])
