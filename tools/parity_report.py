"""Print forward / gradient relative errors of the CUDA path against every reference fixture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch
from helpers import golden_index, load_golden, rel_err
from test_gpu_parity import _run_ours
dev = torch.device("cuda:0")
for case in golden_index():
    g = load_golden(case["name"])
    m, x, init, out, h, c = _run_ours(case, g, dev)
    full = case["full_outputs"]
    fo = rel_err(out, g["f32:out"]) if full else rel_err(out[:, -1], g["f32:out_last"])
    fh = rel_err(h, g["f32:hT"])
    ge = max(rel_err(p.grad, g["f32:grad:" + n]) for n, p in m.named_parameters())
    g64 = max(rel_err(p.grad, g["f64:grad:" + n]) for n, p in m.named_parameters())
    r64 = max(rel_err(g["f32:grad:" + n], g["f64:grad:" + n]) for n, p in m.named_parameters())
    print("%-24s fwd out %.2e hT %.2e | max grad err vs ref32 %.2e vs ref64 %.2e (ref32 vs ref64 %.2e)" % (case["name"], fo, fh, ge, g64, r64))
