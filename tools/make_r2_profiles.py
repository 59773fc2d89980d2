#!/usr/bin/env python
"""Assemble profiles/r2_*.md from what tools/r2_profile.sh and the bench runs brought back in gpurun_out/:
launch lists (ncu gpu__time_duration), ncu --set full summaries, per-source-line stall aggregation, bench JSON lines.
usage: python tools/make_r2_profiles.py"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def bench_line(path):
    txt = open(path).read()
    lines = [l for l in txt.split("\n") if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def launch_table(csv_path):
    rows = [l for l in open(csv_path) if l.startswith('"')]
    rd = list(csv.reader(io.StringIO("".join(rows))))
    hdr, body = rd[0], rd[1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in body:
        name = r[ik]
        short = name.split("(")[0].replace("void ", "")
        if "<" in short:
            short = short.split("<")[0]
        short = short.split("::")[-2] + "::" + short.split("::")[-1] if short.count("::") >= 2 else short
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.1f | %.1f %% |" % (k, n, us, 100 * us / tot))
    return "\n".join(out)


def summary_block(tag):
    p = os.path.join(G, "r2_prof_%s.summary.txt" % tag)
    return open(p).read().strip() if os.path.exists(p) else "(no capture)"


def lines_block(tag, top=12):
    p = os.path.join(G, "r2_prof_%s.cudasass.csv.gz" % tag)
    if not os.path.exists(p):
        return "(no source page)"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), p, str(top)], capture_output=True, text=True)
    return r.stdout.strip()


def main():
    os.makedirs(P, exist_ok=True)
    for cfg in (3, 2):
        lf = os.path.join(G, "r2_launches_cfg%d.csv" % cfg)
        if not os.path.exists(lf):
            continue
        shutil.copy(lf, os.path.join(P, "r2_launches_cfg%d.csv" % cfg))
        with open(os.path.join(P, "r2_cfg%d_summary.md" % cfg), "w") as f:
            f.write("# Round 2 - cfg%d on 1 x B200: launch list and ncu captures\n\n" % cfg)
            f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv python bench.py --config %d "
                    "--steps 2 --warmup 1 --no-cpu-baseline` (tools/r2_profile.sh; full list: `r2_launches_cfg%d.csv`).  Durations under "
                    "ncu are serialised and cold-cache: compare SHARES with the CUDA-event figures of the bench line, not absolutes; "
                    "`k_ffma_probe` is bench.py's FP32 peak probe, `at::*` are torch's fills / loss.\n\n" % (cfg, cfg))
            f.write("## Launch list\n\n" + launch_table(lf) + "\n\n")
            tags = {3: [("cfg3_bwd", "k_rnn_bwd_s (dX-only, kept gates)"), ("cfg3_fwd", "k_rnn_fwd_s")], 2: [("cfg2_bwd", "k_rnn_bwd_s (fused, kept gates)")]}[cfg]
            for tag, title in tags:
                f.write("## %s: ncu --set full --clock-control none\n\n```\n%s\n```\n\nStall samples by phase / source line "
                        "(`tools/ncu_lines.py`):\n\n```\n%s\n```\n\n" % (title, summary_block(tag), lines_block(tag)))
    with open(os.path.join(P, "r2_tc_gemm.md"), "w") as f:
        f.write("# Round 2 - the tcgen05 3xTF32 GEMMs (csrc/tt_tc.cuh) on 1 x B200\n\n")
        f.write("SASS evidence (`cuobjdump -sass tensorized_rnn_b200/csrc/libttrnn_b200.so | grep -o ...`): see `r2_sass_mnemonics.txt`.\n\n")
        for tag, title in (("cfg3_k_tc_red", "k_tc_red (cfg3: dW^T = X^T delta, rows 102 400, M 256 / 40, N 1024)"),
                           ("cfg3_k_tc_rows", "k_tc_rows (cfg3: xg = X W^T and dX = delta W)")):
            f.write("## %s\n\n```\n%s\n```\n\n```\n%s\n```\n\n" % (title, summary_block(tag), lines_block(tag, 10)))
        for name in ("r2_tc_test1.log", "r2_tc_test2.log", "r2_tc_test3.log", "r2_tc_test4.log"):
            p = os.path.join(G, name)
            if os.path.exists(p):
                f.write("## tools/tc_gemm_test: %s\n\n```\n%s\n```\n\n" % (name, open(p).read().strip()))
    # all-config table from the bench lines
    rows = []
    for name, label in (("r2_bench_final.json", "N = 1"), ("r2_scale_n2.json", "N = 2"), ("r2_scale_n4.json", "N = 4"), ("r2_scale_n8.json", "N = 8")):
        p = os.path.join(G, name)
        if os.path.exists(p):
            d = bench_line(p)
            if d:
                rows.append((label, d))
                shutil.copy(p, os.path.join(P, name))
    with open(os.path.join(P, "r2_all_configs.md"), "w") as f:
        f.write("# Round 2 - every config, 1 / 2 / 8 x B200 (bench.py lines copied next to this file)\n\n")
        f.write("`value` = cell-steps/s of the whole job, device-timed (CUDA events, L2 flushed between steps, max over ranks); strong = the "
                "config's global batch split evenly over the ranks, weak = every rank runs the full batch.\n\n")
        for label, d in rows:
            f.write("## %s  (clocks: %s)\n\n" % (label, json.dumps(d.get("clocks"))))
            f.write("| cfg | batch / GPU | ms/step | M cell-steps/s (strong) | e2e M/s | weak: ms / M/s | dominant kernel: frac of FFMA peak | whole-step frac | kernel ms per step |\n|---|---:|---:|---:|---:|---|---|---:|---|\n")
            for r in sorted(d["all_configs"], key=lambda r: r["id"]):
                w = r.get("weak")
                km = ", ".join("%s %.2f" % (k, v["ms_per_step"]) for k, v in r["roofline"]["kernels"].items() if v["ms_per_step"] > 0.05)
                f.write("| %d %s | %d | %.3f | %.3f | %s | %s | %s %.3f | %.3f | %s |\n" % (
                    r["id"], r["config"]["workload"].split(" ", 1)[1][:48], r["config"]["batch_per_gpu"], r["ms_per_step"], r["value"] / 1e6,
                    "%.3f" % (r["e2e"]["value"] / 1e6) if r.get("e2e") else "-",
                    "%.3f / %.3f" % (w["ms_per_step"], w["value"] / 1e6) if w else "-",
                    r["roofline"]["kernel"], r["roofline"]["frac"] or 0, r["roofline"]["whole_step_frac"] or 0, km))
            cb = [(r["id"], r["cpu_baseline"]) for r in d["all_configs"] if r.get("cpu_baseline")]
            if cb:
                f.write("\nCPU oracle on the box's host cores (bounded samples): " + "; ".join(
                    "cfg%d %.0f cell-steps/s (batch %d x T %d, %d threads)" % (i, c["value"], c["batch"], c["seq_len"], c["cores"]) for i, c in cb) + "\n")
            f.write("\n")
    print("wrote", sorted(p for p in os.listdir(P) if p.startswith("r2_")))


if __name__ == "__main__":
    main()
