"""Step time of a log_grads=True module: fused logging (sequence kernels + per-step statistics) vs the reference's
per-step hooks (cell-step mode).  usage: python tools/log_grads_timing.py [B] [T]"""
import io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contextlib import redirect_stdout
import torch
import tensorized_rnn_b200 as tr
from tensorized_rnn_b200 import _lib
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 784)
lib = _lib.load()
for cell, cls in (("lstm", tr.TTLSTM), ("gru", tr.TTGRU)):
    tr.ActivGradLogger.reset()
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        m = cls(1, 256, 1, torch.device("cpu"), n_cores=2, tt_rank=4, log_grads=True).to("cuda:0")
    x = torch.rand(B, T, 1, device="cuda:0")
    for fused in (True, False):
        m.fused_logging = fused
        times = []
        for it in range(3):
            for p in m.parameters():
                p.grad = None
            lib.ttrnn_launch_count(1)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = m(x)[0]
            out[:, -1].sum().backward()
            torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
            tr.ActivGradLogger.end_minibatch()
        print("%s B %d T %d fused_logging=%s: %.1f ms per step, %d library launches" % (cell, B, T, fused, min(times) * 1e3, lib.ttrnn_launch_count(0)), flush=True)
