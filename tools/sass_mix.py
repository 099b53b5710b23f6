"""Opcode mix of one kernel from an `ncu --set full --import-source on` report:
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > k.csv; python tools/sass_mix.py k.csv
Prints executed warp-instructions and stall samples per SASS opcode (where does the issue bandwidth go)."""
import collections
import csv
import re
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    h = rows[hi]
    ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ops, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in rows[hi + 1:]:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ia])
        op = m.group(2) if m else r[ia][:10]
        n = int(r[ie])
        ops[op] += n
        tot += n
        samp[op] += int(r[isamp]) if r[isamp].isdigit() else 0
    print("total executed warp instructions: %d" % tot)
    for k, v in ops.most_common(top):
        print("%-10s %10d %5.1f%%   stall samples %d" % (k, v, 100.0 * v / tot, samp[k]))


if __name__ == "__main__":
    main(sys.argv[1])
