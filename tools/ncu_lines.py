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
# function ranges of tt_static.cuh -> phase names, located by their signatures in the current source
import os
SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tensorized_rnn_b200", "csrc", "tt_static.cuh")
MARK = [("TTS_DEV void cp_async16", "helpers(cp.async/ffma2/ld4)"), ("template <class S, int k>\nstruct St {", "shape structs"),
        ("TTS_DEV void fwd_stage(", "fwd_stage"), ("struct FinMap {", "final map"), ("TTS_DEV void final_partial(", "final_partial"),
        ("TTS_DEV void final_reduce(", "final_reduce"), ("TTS_DEV float sigmoidf_acc", "gate funcs"), ("struct Tune {", "tuning structs"),
        ("__global__ void __launch_bounds__(NTHR, MINB) k_rnn_fwd_s", "k_rnn_fwd_s body"), ("TTS_DEV void stage_weights_t(", "stage_weights_t"),
        ("TTS_DEV void bwd_data_stage(", "bwd_data_stage"), ("struct BwMap {", "bw map"), ("TTS_DEV void bwd_weight_stage(", "bwd_weight_stage"),
        ("TTS_DEV void flush_dw(", "flush_dw"), ("struct TuneB {", "bwd structs"),
        ("__global__ void __launch_bounds__(NTHR, 1) k_rnn_bwd_s", "k_rnn_bwd_s prologue"),
        ("auto fetch_h = [&]", "fetch/prefetch"), ("for (int t = a.steps - 1; t >= 0; --t) {", "step prologue"),
        ("// ---- gates and their gradients", "gates+delta"), ("// ---- backward chain: core gradients", "bwd chain call"),
        ("// Batched TT matvec (ih projection", "batched kernels")]
PH = []
try:
    text = open(SRC).read()
    pos = []
    for pat, name in MARK:
        i = text.find(pat.replace("\\n", "\n"))
        if i >= 0:
            pos.append((text.count("\n", 0, i) + 1, name))
    pos.sort()
    for q, (ln, name) in enumerate(pos):
        end = pos[q + 1][0] - 1 if q + 1 < len(pos) else 10 ** 9
        PH.append((ln, end, name))
except Exception:
    pass
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
