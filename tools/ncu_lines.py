#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line / by function.
usage: python tools/ncu_lines.py prof.cudasass.csv.gz [top_n]"""
import csv, gzip, io, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
lines = {}
cur_file = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    # the header has two "Source" columns (cuda line text, sass); dict keeps the last; line text is r[1]
    def num(k):
        try:
            return float(d.get(k, "0") or 0)
        except ValueError:
            return 0.0
    key = (cur_file, int(r[0]))
    e = lines.setdefault(key, {"src": r[1].strip()[:90], "samples": 0.0, "inst": 0.0, "st": {}, "wf": 0.0, "wfx": 0.0})
    e["samples"] += num("# Samples")
    e["inst"] += num("Instructions Executed")
    e["wf"] += num("L1 Wavefronts Shared")
    e["wfx"] += num("L1 Wavefronts Shared Excessive")
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            e["st"][k] = e["st"].get(k, 0.0) + num(k)
tot = sum(e["samples"] for e in lines.values()) or 1.0
print("total samples %.0f over %d source lines" % (tot, len(lines)))
# function ranges of tt_static.cuh (by line number) -> phase names
PH = [(179, 250, "fwd_stage"), (296, 361, "final_partial"), (365, 419, "final_reduce"), (420, 442, "gate funcs"),
      (777, 909, "bwd_data_stage"), (944, 997, "bwd_weight_stage"), (1000, 1040, "flush_dw"),
      (1235, 1300, "fetch/prefetch"), (1300, 1335, "step prologue"), (1336, 1412, "gates+delta"), (34, 58, "helpers(cp.async/ffma2/ld4)")]
ph = {}
for (f, ln), e in lines.items():
    name = "other"
    if f == "tt_static.cuh":
        for a, b, n in PH:
            if a <= ln <= b:
                name = n
                break
    else:
        name = f
    p = ph.setdefault(name, {"samples": 0.0, "inst": 0.0, "wf": 0.0, "wfx": 0.0, "st": {}})
    p["samples"] += e["samples"]; p["inst"] += e["inst"]; p["wf"] += e["wf"]; p["wfx"] += e["wfx"]
    for k, v in e["st"].items():
        p["st"][k] = p["st"].get(k, 0.0) + v
print("\nby phase:")
for n, p in sorted(ph.items(), key=lambda kv: -kv[1]["samples"]):
    st = sorted(p["st"].items(), key=lambda kv: -kv[1])[:4]
    print("  %-28s %5.1f%%  inst %.3e  smem wf %.3e (excess %.2e)  %s" % (
        n, 100 * p["samples"] / tot, p["inst"], p["wf"], p["wfx"], " ".join("%s=%.0f%%" % (k[6:], 100 * v / max(p["samples"], 1)) for k, v in st)))
print("\ntop lines:")
for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(e["st"].items(), key=lambda kv: -kv[1])[:3]
    print("  %s:%-5d %5.1f%% inst %.2e wf %.2e/%.2e  %-60s %s" % (f, ln, 100 * e["samples"] / tot, e["inst"], e["wf"], e["wfx"], e["src"][:60],
                                                      " ".join("%s=%.0f%%" % (k[6:], 100 * v / max(e["samples"], 1)) for k, v in st)))
