"""Shared helpers for the test-suite (oracle access, fixtures, error metrics)."""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ttrnn_oracle as oracle  # noqa: E402  (test infrastructure)

FWD_TOL = 1e-5     # north_star: outputs / final states, relative
GRAD_TOL = 1e-4    # north_star: gradients, relative


def rel_err(a, b):
    """||a - b|| / ||b|| (norm-wise relative error; the bar SURVEY.md section 8c states)."""
    a = (a.detach() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))).double().cpu().reshape(-1)
    b = (b.detach() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b))).double().cpu().reshape(-1)
    den = float(b.norm())
    num = float((a - b).norm())
    if den == 0.0:
        return num
    return num / den


def golden_index():
    with open(os.path.join(GOLDEN, "index.json")) as f:
        return json.load(f)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def state_dict_from_golden(g):
    return {k[len("param:"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param:")}


def build_module(case, device="cpu"):
    """Our module with the parameters of a golden case loaded."""
    import tensorized_rnn_b200 as tr
    cls = tr.TTLSTM if case["cell"] == "lstm" else tr.TTGRU
    with redirect_stdout(io.StringIO()):
        m = cls(case["input_size"], case["hidden_size"], case["num_layers"], torch.device("cpu"),
                n_cores=case["n_cores"], tt_rank=case["tt_rank"], bias=case["bias"])
    return m


def quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)
