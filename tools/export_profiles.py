"""profiles/ artefacts from one `ncu --set full` report of tools/prof_kernels.py (reps = 1, kernel filter
"attn|gemm|conv|wgrad|ln_"):
  python tools/export_profiles.py gpurun_out/prof_r2c.ncu-rep r2
writes profiles/<tag>_kernels_raw.csv.gz (the raw page: every metric of every captured launch),
profiles/<tag>_kernels_summary.txt (one line per launch) and profiles/<tag>_kernel_traffic.json
(label of bench.py's roofline entry -> DRAM bytes read + written per launch, what `roofline.traffic` reports)."""
import csv
import gzip
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# launch order of tools/prof_kernels.py after its two set-up launches (attn_fwd2, ln_fwd)
ORDER = [
    ("fused attention fwd (QK^T, softmax, PV)", 1),
    ("fused attention bwd (dK/dV + dQ kernels)", 3),
    ("QKV projection fwd (N=768, K=256)", 1),
    ("output projection fwd", 1),
    ("FFN Conv1d k=1 1024->256 fwd", 1),
    ("FFN Conv1d k=9 256->1024 fwd (+bias+ReLU+mask)", 1),
    ("FFN Conv1d k=1 input-gradient (+ReLU mask)", 1),
    ("QKV projection input-gradient (+residual)", 1),
    ("FFN Conv1d k=9 input-gradient (+residual)", 1),
    ("FFN Conv1d k=9 weight-gradient", 1),
    ("LayerNorm fwd (dropout + residual + LN + pad-zero)", 1),
    ("LayerNorm bwd (+dgamma/dbeta/dbias, stored keep bits)", 1),
]
COLS = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("launch__registers_per_thread", "regs"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%")]


def main(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    raw = raw[raw.index('"ID"'):]
    out = os.path.join(ROOT, "profiles")
    with gzip.open(os.path.join(out, "%s_kernels_raw.csv.gz" % tag), "wt") as f:
        f.write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, data = rows[0], rows[1], rows[2:]
    idx = {name: h.index(name) for name, _ in COLS}
    names = [r[h.index("Kernel Name")] for r in data]
    start = next(i for i, n in enumerate(names) if "ln_fwd" in n) + 1  # after the set-up launches
    labels = {}
    i = start
    for label, n in ORDER:
        for k in range(n):
            labels[i + k] = label
        i += n
    lines = ["%3s %-40s %-58s %-10s" % ("id", "kernel", "bench.py roofline entry", "grid") +
             "".join("%9s" % c for _, c in COLS)]
    traffic = {}
    for j, r in enumerate(data):
        vals = [float(r[idx[name]].replace(",", "")) for name, _ in COLS]
        # ncu prints bytes in the unit of the second header row (Mbyte here)
        lines.append("%3d %-40s %-58s %-10s" % (j, names[j].replace("void ", "").split("(")[0][:40], labels.get(j, "(set-up)")[:58],
                                                r[h.index("Grid Size")].replace(" ", "")) + "".join("%9.1f" % v for v in vals))
        if j in labels:
            scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
            rd = vals[2] * scale[units[idx["dram__bytes_read.sum"]]]
            wr = vals[3] * scale[units[idx["dram__bytes_write.sum"]]]
            traffic[labels[j]] = traffic.get(labels[j], 0.0) + rd + wr
    open(os.path.join(out, "%s_kernels_summary.txt" % tag), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(os.path.join(out, "%s_kernel_traffic.json" % tag), "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
