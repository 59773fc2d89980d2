"""GPU parity of the cell-step mode (is_naive / new_core / log_grads) against fixtures generated from the
reference (tests/golden/make_golden_variants.py).  Bars: 1e-5 forward, 1e-4 gradients, 1e-4 on the logged
norms (they are sums of squares of gradients)."""
import json
import os

import pytest
import torch

import tensorized_rnn_b200 as tr
from helpers import FWD_TOL, GOLDEN, GRAD_TOL, load_golden, rel_err, state_dict_from_golden
from test_variants_cpu import VARIANTS, build_variant

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", VARIANTS, ids=[c["name"] for c in VARIANTS])
def test_variant_matches_reference_fixture(case):
    """log_grads cases run the FUSED logging path here (sequence kernels + per-step statistics from the BPTT kernel)."""
    _run_variant(case, fused_logging=True)


@pytest.mark.parametrize("case", [c for c in VARIANTS if c["log_grads"]], ids=[c["name"] for c in VARIANTS if c["log_grads"]])
def test_log_grads_cell_step_hooks_match_reference_fixture(case):
    """The reference's own mechanism (one cell call per step, forward hooks + tensor hooks) against the same fixtures."""
    _run_variant(case, fused_logging=False)


def _run_variant(case, fused_logging):
    dev = torch.device("cuda:0")
    tr.ActivGradLogger.reset()
    g = load_golden("variant_" + case["name"])
    m = build_variant(case)
    m.load_state_dict(state_dict_from_golden(g), strict=True)
    m = m.to(dev)
    m.fused_logging = fused_logging
    from tensorized_rnn_b200 import _lib
    _lib.load().ttrnn_launch_count(1)
    x = torch.from_numpy(g["x"]).to(dev).requires_grad_(True)
    lstm = case["cell"] == "lstm"
    init = None
    if case["init_states"]:
        h0 = torch.from_numpy(g["h0"]).to(dev).requires_grad_(True)
        init = (h0, torch.from_numpy(g["c0"]).to(dev).requires_grad_(True)) if lstm else h0
    if lstm:
        out, (h, c) = m(x, init)
        loss = (out * torch.from_numpy(g["w_out"]).to(dev)).sum() + (h * torch.from_numpy(g["w_h"]).to(dev)).sum() \
            + (c * torch.from_numpy(g["w_c"]).to(dev)).sum()
    else:
        out, h = m(x, init)
        loss = (out * torch.from_numpy(g["w_out"]).to(dev)).sum() + (h * torch.from_numpy(g["w_h"]).to(dev)).sum()
    loss.backward()
    torch.cuda.synchronize()
    assert rel_err(out, g["f32:out"]) <= FWD_TOL
    assert rel_err(h, g["f32:hT"]) <= FWD_TOL
    if lstm:
        assert rel_err(c, g["f32:cT"]) <= FWD_TOL
    errs = {"dx": rel_err(x.grad, g["f32:dx"])}
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        errs[name] = rel_err(p.grad, g["f32:grad:" + name])
    if case["init_states"]:
        errs["dh0"] = rel_err((init[0] if lstm else init).grad, g["f32:dh0"])
        if lstm:
            errs["dc0"] = rel_err(init[1].grad, g["f32:dc0"])
    bad = {k: v for k, v in errs.items() if not v <= GRAD_TOL}
    assert not bad, "gradient rel err above %.0e: %s" % (GRAD_TOL, bad)
    if case["log_grads"] and not case.get("is_naive"):
        # fused: a handful of launches per layer; cell-step: several per (step, layer)
        n = int(_lib.load().ttrnn_launch_count(0))
        T, L = x.shape[1], case["num_layers"]
        assert (n < 40 * L) if fused_logging else (n >= 4 * T * L), (n, T, L)
    if case["log_grads"]:
        tr.ActivGradLogger.end_minibatch()
        tr.ActivGradLogger.end_epoch()
        logs = tr.ActivGradLogger.get_logs()
        ref_keys = sorted(k for k in g if k.startswith("log:"))
        assert sorted("log:%s:%s" % k for k in logs) == ref_keys
        for (var, qnt), mat in logs.items():
            assert rel_err(mat, g["log:%s:%s" % (var, qnt)]) <= 1e-4, (var, qnt)
    tr.ActivGradLogger.reset()


def test_new_core_runs_fused_and_naive_runs_stepwise():
    """new_core only changes the TT shapes (fused kernels); is_naive / log_grads switch to cell-step mode.
    Both modes of the same concat-gates module must agree."""
    dev = torch.device("cuda:0")
    tr.ActivGradLogger.reset()
    torch.manual_seed(3)
    import io
    from contextlib import redirect_stdout
    with redirect_stdout(io.StringIO()):
        fused = tr.TTLSTM(28, 64, 2, torch.device("cpu"), n_cores=2, tt_rank=4).to(dev)
        hooked = tr.TTLSTM(28, 64, 2, torch.device("cpu"), n_cores=2, tt_rank=4, log_grads=True).to(dev)
    hooked.load_state_dict(fused.state_dict())
    assert not fused.cell_step_mode and hooked.cell_step_mode
    x = torch.rand(5, 9, 28, device=dev)
    res = []
    for m in (fused, hooked):
        out, (h, c) = m(x)
        (out.sum() + 2 * h.sum() + c.sum()).backward()
        res.append((out.detach(), [p.grad.clone() for p in m.parameters()]))
    assert rel_err(res[1][0], res[0][0]) <= FWD_TOL
    for a, b in zip(res[1][1], res[0][1]):
        assert rel_err(a, b) <= GRAD_TOL
    tr.ActivGradLogger.reset()
