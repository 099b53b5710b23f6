import json,sys
for f in sys.argv[1:]:
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3))
    for k in d.get("roofline_kernels",[]):
        print("   %-52s %-18s %7.1f us  %5.3f" % (k["label"][:52], k["kernel"], k["avg_launch_ms"]*1e3, k["frac"]))
    for k in d.get("roofline_hbm",[]):
        print("   %-52s %-18s %7.1f us  %5.3f" % (k["label"][:52], "hbm", k["avg_launch_ms"]*1e3, k["frac"]))
