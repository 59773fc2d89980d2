"""Forward error of the ih-projection routes on an amplifying case (scale 1.5 weights, T = 100): tensor cores vs FFMA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from tensorized_rnn_b200 import _lib
from helpers import oracle, rel_err
from test_gpu_round2 import make_pair, options

for (cell, I, H, L, d, r, B, T, chunk, scale) in [("lstm", 256, 1024, 1, 4, 8, 6, 100, 48, 1.5), ("lstm", 256, 1024, 1, 4, 8, 6, 100, 48, 1.0),
                                                   ("lstm", 40, 256, 3, 3, 8, 32, 100, 0, 1.5), ("lstm", 40, 256, 3, 3, 8, 32, 160, 0, 1.0)]:
    layers, m = make_pair(cell, I, H, L, d, r, scale=scale)
    x = torch.rand(B, T, I, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        o_ref, (h_ref, _) = oracle.lstm_forward(layers, x)
        dl = [{k: (v.double() if torch.is_tensor(v) else ([c.double() for c in v] if isinstance(v, list) else v)) for k, v in p.items()} for p in layers]
        try:
            o64, _ = oracle.lstm_forward(dl, x.double(), (torch.zeros(B, H, dtype=torch.float64), torch.zeros(B, H, dtype=torch.float64)))
        except Exception as e:
            o64 = None
    for tc in (1, 0):
        with options(chunk_steps=chunk, tc_gemm=tc):
            with torch.no_grad():
                out = m(x.to("cuda:0"))[0]
        print("I=%d H=%d L=%d T=%d scale=%.1f tc=%d : rel err vs fp32 oracle %.3e%s" % (
            I, H, L, T, scale, tc, rel_err(out, o_ref), "" if o64 is None else "  vs fp64 oracle %.3e (oracle32 vs 64: %.3e)" % (rel_err(out, o64), rel_err(o_ref, o64))))
