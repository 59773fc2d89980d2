#!/usr/bin/env python
"""Build profiles/r1_all_configs.md and profiles/r1_cfg2_summary.md from the files tools/final_evidence.sh
brought back in gpurun_out/ (bench JSON lines, ncu launch list, ncu summaries)."""
import csv, json, os, sys, collections, gzip, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

def line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])

rows, raw = [], []
for c in (1, 2, 3, 4, 5):
    f = os.path.join(G, "final_cfg%d.json" % c)
    if not os.path.exists(f):
        continue
    l = line(f)
    raw.append(l)
    r = l["roofline"]
    cb = l.get("cpu_baseline") or {}
    rows.append("| %d | %s | %.2f | %.3e | %.3e | %s %.3f | %.3f | %s | %s |" % (
        c, l["config"]["workload"], l["ms_per_step"], l["value"], l["e2e"]["value"], r["kernel"], r["frac"],
        r["whole_step_frac"], ", ".join("L%d %s" % (x["layer"], x["route"]) for x in r.get("ih_projection", [])),
        ("%.0f (%s, %d cores) -> %.0fx" % (cb["value"], cb["sample"].split(",")[0], cb["cores"], l["e2e"]["value"] / cb["value"])) if cb else "-"))
with open(os.path.join(P, "r1_all_configs.md"), "w") as f:
    f.write("# Round 1 (final build) - all five BASELINE.json configs on 1 x B200\n\n")
    f.write("`python bench.py --config N --steps 3 --warmup 3` (cfg2: defaults, 10 steps).  Fractions use ALGORITHMIC FLOPs of the\n"
            "reference's core-by-core sweep (backward credited at 2x forward; recomputed stages are not credited; where the\n"
            "ih projection runs in the cheaper dense order the kernels execute fewer FLOPs than credited - see `ih route`).\n"
            "FP32 peak = `ttrnn_ffma_probe` in the same run.\n\n")
    f.write("| cfg | workload | ms/step | cell-steps/s | e2e cell-steps/s | dominant kernel: frac of FP32 peak | whole-step frac | ih route | CPU oracle baseline -> e2e speed-up |\n")
    f.write("|---|---|---:|---:|---:|---|---:|---|---|\n")
    f.write("\n".join(rows) + "\n\nRaw bench lines:\n\n")
    for l in raw:
        f.write("```json\n" + json.dumps(l) + "\n```\n")
    sc = []
    nraw = 0
    for c in (2, 3):
        f1 = os.path.join(G, "final_cfg%d.json" % c)
        if not os.path.exists(f1):
            continue
        a = line(f1)
        cells = ["%d (%s)" % (c, a["config"]["mode"]), "%.2f ms, %.3e/s" % (a["ms_per_step"], a["value"])]
        for n in (2, 4, 8):
            fn = os.path.join(G, "scale%d_cfg%d.json" % (n, c))
            if os.path.exists(fn):
                b = line(fn)
                cells.append("%.2f ms, %.3e/s (%.2fx)" % (b["ms_per_step"], b["value"], b["value"] / a["value"]))
                raw.append(b); nraw += 1
            else:
                cells.append("-")
        sc.append("| " + " | ".join(cells) + " |")
    if sc:
        f.write("\nWeak scaling (fixed per-GPU batch; `gpurun --gpus N`, torchrun, NCCL, one flat all-reduce of the TT-core gradients per\n"
                "step; max over ranks of device time):\n\n"
                "| cfg | 1 GPU | 2 GPUs | 4 GPUs | 8 GPUs |\n|---|---|---|---|---|\n" + "\n".join(sc) + "\n\n```json\n"
                + "\n".join(json.dumps(x) for x in raw[-nraw:]) + "\n```\n")
    rf = os.path.join(G, "final_reference_cfg2.json")
    if os.path.exists(rf):
        f.write("\nReference arm (`bench.py --impl reference`, oracle port of the reference's PyTorch path on the host cores):\n\n```json\n"
                + open(rf).read().strip().splitlines()[-1] + "\n```\n")
print("wrote r1_all_configs.md with", len(rows), "configs")

# launch list
lf = os.path.join(G, "final_launches_cfg2.csv")
if os.path.exists(lf):
    txt = [ln for ln in open(lf) if ln.startswith('"')]
    rd = list(csv.reader(txt))
    hdr = rd[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rd[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("<")[0].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    unit = rd[1][hdr.index("Metric Unit")] if len(rd) > 1 else "ns"
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, "r1_cfg2_summary.md"), "w") as f:
        f.write("# Round 1 (final build) - cfg2 (permuted-MNIST TT-GRU d2 r4, B=1024, T=784, fwd+bwd) on 1 x B200\n\n")
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline`\n"
                "(full launch list: `r1_launches_cfg2.csv`; durations under ncu are serialised/cold-cache: compare SHARES)\n\n")
        f.write("## Launch list (all launches of the command, ncu durations)\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f %% |\n" % (name, n, v * scale, 100 * v / tot))
        l = line(os.path.join(G, "final_cfg2.json"))
        f.write("\n## bench.py line of the same build (CUDA events, not under ncu)\n\n```json\n" + json.dumps(l) + "\n```\n\n")
        k = l["roofline"]["kernels"]
        f.write("Event-timed kernel shares of the step: " + ", ".join("%s %.1f %%" % (n, 100 * v["ms_per_step"] / l["ms_per_step"]) for n, v in k.items())
                + " (rest: torch glue, reductions).\n\n")
        for tag, title in (("final_cfg2_bwd", "k_rnn_bwd_s (dominant kernel)"), ("final_cfg2_fwd", "k_rnn_fwd_s")):
            sf = os.path.join(G, "prof_%s.summary.txt" % tag)
            if os.path.exists(sf):
                f.write("## %s, ncu --set full --clock-control none, T=784\n\n```\n%s```\n(durations in ms, dram bytes in GB)\n\n" % (title, open(sf).read()))
            cf = os.path.join(G, "prof_%s.cudasass.csv.gz" % tag)
            if os.path.exists(cf):
                out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), cf, "14"], capture_output=True, text=True).stdout
                f.write("Stall samples by phase / source line (`tools/ncu_lines.py`):\n\n```\n" + out + "```\n\n")
    import shutil
    shutil.copy(lf, os.path.join(P, "r1_launches_cfg2.csv"))
    # measured DRAM traffic of the dominant kernel -> profiles/traffic.json (read by bench.py)
    sf = os.path.join(G, "prof_final_cfg2_bwd.summary.txt")
    if os.path.exists(sf):
        vals = {}
        for ln in open(sf):
            parts = ln.split()
            if len(parts) >= 2 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = float(parts[1])
                unit = parts[2].lower() if len(parts) > 2 else "gbyte"      # ncu prints these two in Gbyte at this size
                vals[parts[0]] = v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1e9)
        if len(vals) == 2:
            json.dump({l["config"]["workload"]: {"kernel": "k_rnn_bwd", "dram_bytes_per_launch": sum(vals.values()),
                                                 "dram_read_bytes": vals["dram__bytes_read.sum"],
                                                 "dram_write_bytes": vals["dram__bytes_write.sum"],
                                                 "source": "profiles/r1_cfg2_summary.md (ncu --set full, dram__bytes_read.sum + "
                                                           "dram__bytes_write.sum of k_rnn_bwd_s at T=784)"}},
                      open(os.path.join(P, "traffic.json"), "w"), indent=1)
            print("wrote traffic.json", vals)
    print("wrote r1_cfg2_summary.md")


def config_summary(cfg_no, title, kernels):
    """profiles/r1_cfg<N>_summary.md from final_launches_cfg<N>.csv + prof_final_cfg<N>_*.summary.txt"""
    lf = os.path.join(G, "final_launches_cfg%d.csv" % cfg_no)
    if not os.path.exists(lf):
        return
    txt = [ln for ln in open(lf) if ln.startswith('"')]
    rd = list(csv.reader(txt))
    hdr = rd[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rd[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("<")[0].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    unit = rd[1][hdr.index("Metric Unit")] if len(rd) > 1 else "ns"
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, "r1_cfg%d_summary.md" % cfg_no), "w") as f:
        f.write("# Round 1 (final build) - %s on 1 x B200\n\n" % title)
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --config %d --steps 2 --warmup 1 --no-cpu-baseline`\n"
                "(full launch list: `r1_launches_cfg%d.csv`; durations under ncu are serialised/cold-cache: compare SHARES; `k_ffma_probe` is the peak probe of bench.py)\n\n" % (cfg_no, cfg_no))
        f.write("## Launch list (all launches of the command, ncu durations)\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f %% |\n" % (name, n, v * scale, 100 * v / tot))
        bf = os.path.join(G, "final_cfg%d.json" % cfg_no)
        if os.path.exists(bf):
            l = line(bf)
            f.write("\n## bench.py line of the same build (CUDA events, not under ncu)\n\n```json\n" + json.dumps(l) + "\n```\n\n")
        for tag, ktitle in kernels:
            sf = os.path.join(G, "prof_%s.summary.txt" % tag)
            if os.path.exists(sf):
                f.write("## %s, ncu --set full --clock-control none\n\n```\n%s```\n\n" % (ktitle, open(sf).read()))
            cf = os.path.join(G, "prof_%s.cudasass.csv.gz" % tag)
            if os.path.exists(cf):
                out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), cf, "10"], capture_output=True, text=True).stdout
                f.write("Stall samples by phase / source line (`tools/ncu_lines.py`):\n\n```\n" + out + "```\n\n")
    import shutil
    shutil.copy(lf, os.path.join(P, "r1_launches_cfg%d.csv" % cfg_no))
    print("wrote r1_cfg%d_summary.md" % cfg_no)


config_summary(3, "cfg3 (GE2E speaker encoder, 3 x TT-LSTM d3 r8, B=640, T=160, fwd+bwd)",
               [("final_cfg3_bwd", "k_rnn_bwd_s (dX-only, kept gates)"), ("final_cfg3_fwd", "k_rnn_fwd_s"),
                ("final_cfg3_red", "k_gemm_red (dense core-gradient accumulation)"), ("final_cfg3_rows", "k_gemm_rows (dense ih projection / dX)")])
