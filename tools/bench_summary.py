#!/usr/bin/env python
"""Print one compact line per bench JSON file given on the command line."""
import json, sys
for f in sys.argv[1:]:
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    k = l["roofline"]["kernels"]
    print("%s: %.2f ms/step, %.3e steps/s, e2e %.3e | %s | dom %s frac %.3f whole %.3f" % (
        f.split("/")[-1], l["ms_per_step"], l["value"], l["e2e"]["value"],
        " ".join("%s %.2f(%d)" % (n.replace("k_", ""), v["ms_per_step"], v["launches_per_step"]) for n, v in k.items()),
        l["roofline"]["kernel"], l["roofline"]["frac"], l["roofline"]["whole_step_frac"]))
