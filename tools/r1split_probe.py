"""Small rank-one-input stacks through the split (dX-only) BPTT variants; run under compute-sanitizer when debugging."""
import io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contextlib import redirect_stdout
import torch
import tensorized_rnn_b200 as tr
from tensorized_rnn_b200 import _lib
sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 12
for cell, cls in (("lstm", tr.TTLSTM), ("gru", tr.TTGRU)):
    torch.manual_seed(1)
    with redirect_stdout(io.StringIO()):
        m = cls(1, 256, 1, torch.device("cpu"), n_cores=2, tt_rank=4).to("cuda:0")
    print(_lib.describe_plan(m.spec().desc(B, T)), flush=True)
    x = torch.rand(B, T, 1, device="cuda:0")
    res = m(x)
    out = res[0]
    out[:, -1].sum().backward()
    torch.cuda.synchronize()
    print(cell, "ok", float(out.abs().sum()), [float(p.grad.abs().sum()) for p in m.parameters()][:4], flush=True)
