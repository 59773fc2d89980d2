#!/usr/bin/env python
"""Assemble profiles/r2b_*.md (second half of round 2: row groups, dX-only d2 kernels, fused log_grads) from what
tools/r2b_profile.sh and the A/B scripts brought back in gpurun_out/.   usage: python tools/make_r2b_profiles.py"""
import json
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_r2_profiles import G, P, bench_line, launch_table  # noqa: E402
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def summary_block(tag):
    p = os.path.join(G, "r2b_prof_%s.summary.txt" % tag)
    return open(p).read().strip() if os.path.exists(p) else "(no capture)"


def lines_block(tag, top=12):
    p = os.path.join(G, "r2b_prof_%s.cudasass.csv.gz" % tag)
    if not os.path.exists(p):
        return "(no source page)"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), p, str(top)], capture_output=True, text=True)
    return r.stdout.strip()


def dram(tag):
    rd = wr = None
    for l in summary_block(tag).split("\n"):
        t = l.split()
        if len(t) >= 3 and t[0] == "dram__bytes_read.sum":
            rd = float(t[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[t[2]]
        if len(t) >= 3 and t[0] == "dram__bytes_write.sum":
            wr = float(t[1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[t[2]]
    return rd, wr


def main():
    os.makedirs(P, exist_ok=True)
    titles = {3: [("cfg3_bwd", "k_rnn_bwd_s (dX-only, kept gates), group 0 launch: 3 rows per CTA, 148 CTAs, 444 rows"),
                  ("cfg3_fwd", "k_rnn_fwd_s, group 0 launch: 3 rows per CTA, 148 CTAs, 444 rows")],
              2: [("cfg2_bwd", "k_rnn_bwd_s (dX-only, kept gates, rank-one input, GRU): 7 rows per CTA, 147 CTAs")]}
    for cfg in (3, 2):
        lf = os.path.join(G, "r2b_launches_cfg%d.csv" % cfg)
        if not os.path.exists(lf):
            continue
        shutil.copy(lf, os.path.join(P, "r2b_launches_cfg%d.csv" % cfg))
        with open(os.path.join(P, "r2b_cfg%d_summary.md" % cfg), "w") as f:
            f.write("# Round 2, final build - cfg%d on 1 x B200: launch list and ncu captures\n\n" % cfg)
            f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv python bench.py --config %d "
                    "--steps 2 --warmup 1 --no-cpu-baseline` (tools/r2b_profile.sh; full list: `r2b_launches_cfg%d.csv`).  Durations under "
                    "ncu are serialised (the two row groups of cfg3 run concurrently outside ncu) and cold-cache: compare SHARES with the "
                    "CUDA-event figures of the bench line's kernel-timing pass, not absolutes; `k_ffma_probe` is bench.py's FP32 peak "
                    "probe, `at::*` are torch's fills / loss.\n\n" % (cfg, cfg))
            f.write("## Launch list\n\n" + launch_table(lf) + "\n\n")
            for tag, title in titles[cfg]:
                f.write("## %s: ncu --set full --clock-control none\n\n```\n%s\n```\n\nStall samples by phase / source line "
                        "(`tools/ncu_lines.py`):\n\n```\n%s\n```\n\n" % (title, summary_block(tag), lines_block(tag)))
    # traffic.json: per-launch DRAM bytes of the dominant kernels
    tp = os.path.join(P, "traffic.json")
    tj = json.load(open(tp)) if os.path.exists(tp) else {}
    for key, tag, what in (("cfg3", "cfg3_bwd", "one k_rnn_bwd_s launch of cfg3: group 0 of layer 2, 444 rows x T 160, 3 rows per CTA, 148 CTAs"),
                           ("cfg2", "cfg2_bwd", "k_rnn_bwd_s (dX-only, rank-one input) at B 1024, T 784")):
        rd, wr = dram(tag)
        if rd is not None and wr is not None:
            tj[key] = {"kernel": "k_rnn_bwd", "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
                       "source": "profiles/r2b_%s_summary.md (ncu --set full, %s)" % (key, what)}
    json.dump(tj, open(tp, "w"), indent=1)
    for name in ("r2b_bench_final.json", "r2b_scale_n2.json"):
        p = os.path.join(G, name)
        if os.path.exists(p) and bench_line(p):
            shutil.copy(p, os.path.join(P, name))
    d = bench_line(os.path.join(G, "r2b_bench_final.json")) if os.path.exists(os.path.join(G, "r2b_bench_final.json")) else None
    if d:
        with open(os.path.join(P, "r2b_all_configs.md"), "w") as f:
            f.write("# Round 2, final build - every config on 1 x B200 (`r2b_bench_final.json`)\n\n")
            f.write("`ms/step` = the concurrent run (CUDA events around the whole step, L2 flushed between steps); kernel columns = the "
                    "separate kernel-timing pass (row groups serialised, each kernel's own duration).  clocks: %s\n\n" % json.dumps(d.get("clocks")))
            f.write("| cfg | ms/step | M cell-steps/s | e2e M/s | row groups | dominant kernel | frac of FFMA peak | kind-group frac | whole-step frac | kernel ms per step (timing pass) |\n|---|---:|---:|---:|---|---|---:|---:|---:|---|\n")
            for r in sorted(d["all_configs"], key=lambda r: r["id"]):
                ro = r["roofline"]
                h = ro["plan"][0]
                km = ", ".join("%s %.2f" % (k, v["ms_per_step"]) for k, v in ro["kernels"].items() if v["ms_per_step"] > 0.05)
                grp = "%s + %s" % (h.get("group_rows0"), h.get("group_rows1")) if h.get("row_groups") == 2 else "-"
                f.write("| %d %s | %.3f | %.3f | %.3f | %s | %s | %.3f | %.3f | %.3f | %s |\n" % (
                    r["id"], r["config"]["workload"].split(" ", 1)[1][:44], r["ms_per_step"], r["value"] / 1e6, r["e2e"]["value"] / 1e6, grp,
                    ro["kernel"], ro["frac"] or 0, ro.get("kind_group_frac") or 0, ro["whole_step_frac"] or 0, km))
            f.write("\n## Recurrent kernels per variant (rows per CTA, grid)\n\n| cfg | kernel | rows/CTA | CTAs | rows | launches/step | ms/launch | frac of FFMA peak | on the occupied SMs |\n|---|---|---:|---:|---:|---:|---:|---:|---:|\n")
            for r in sorted(d["all_configs"], key=lambda r: r["id"]):
                for v in r["roofline"].get("recurrent_variants", []):
                    f.write("| %d | %s | %d | %d | %d | %.1f | %.3f | %.3f | %.3f |\n" % (r["id"], v["kernel"], v["rows_per_cta"], v["ctas"], v["rows"],
                            v["launches_per_step"], v["ms_per_launch"], v["frac_of_ffma_peak"] or 0, v["frac_on_occupied_sms"] or 0))
            cb = [(r["id"], r["cpu_baseline"]) for r in d["all_configs"] if r.get("cpu_baseline")]
            if cb:
                f.write("\nCPU oracle on the box's host cores (bounded samples): " + "; ".join(
                    "cfg%d %.0f cell-steps/s (batch %d x T %d, %d threads)" % (i, c["value"], c["batch"], c["seq_len"], c["cores"]) for i, c in cb) + "\n")
    print("wrote", sorted(p for p in os.listdir(P) if p.startswith("r2b_")))


if __name__ == "__main__":
    main()
