"""GPU parity: CUDA path vs fixtures generated from the reference and vs the CPU oracle.

Tolerances are north_star's: 1e-5 relative on outputs / final states, 1e-4 relative on gradients
(norm-wise, ||a-b||/||b||).  Everything goes through the public module API, which calls the C ABI.
"""
import numpy as np
import pytest
import torch

from helpers import (FWD_TOL, GRAD_TOL, build_module, golden_index, load_golden, oracle, rel_err,
                     state_dict_from_golden)

pytestmark = pytest.mark.gpu

CASES = golden_index()


def _run_ours(case, g, dev):
    m = build_module(case)
    m.load_state_dict(state_dict_from_golden(g))
    m = m.to(dev)
    x = torch.from_numpy(g["x"]).to(dev).requires_grad_(True)
    lstm = case["cell"] == "lstm"
    init = None
    if case["init_states"]:
        h0 = torch.from_numpy(g["h0"]).to(dev).requires_grad_(True)
        init = (h0, torch.from_numpy(g["c0"]).to(dev).requires_grad_(True)) if lstm else h0
    if lstm:
        out, (h, c) = m(x, init)
    else:
        out, h = m(x, init)
        c = None
    full = case["full_outputs"]
    if full:
        loss = (out * torch.from_numpy(g["w_out"]).to(dev)).sum()
    else:
        loss = (out[:, -1] * torch.from_numpy(g["w_out_last"]).to(dev)).sum()
    loss = loss + (h * torch.from_numpy(g["w_h"]).to(dev)).sum()
    if lstm:
        loss = loss + (c * torch.from_numpy(g["w_c"]).to(dev)).sum()
    loss.backward()
    return m, x, init, out, h, c


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_matches_reference_fixture(case):
    dev = torch.device("cuda:0")
    g = load_golden(case["name"])
    m, x, init, out, h, c = _run_ours(case, g, dev)
    full = case["full_outputs"]
    errs = {}
    if full:
        errs["out"] = rel_err(out, g["f32:out"])
    else:
        errs["out_last"] = rel_err(out[:, -1], g["f32:out_last"])
    errs["hT"] = rel_err(h, g["f32:hT"])
    if c is not None:
        errs["cT"] = rel_err(c, g["f32:cT"])
    for k, v in errs.items():
        assert v <= FWD_TOL, "forward %s rel err %.3e > %.0e (%s)" % (k, v, FWD_TOL, errs)
    gerrs = {}
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        gerrs[name] = rel_err(p.grad, g["f32:grad:" + name])
    if full:
        gerrs["dx"] = rel_err(x.grad, g["f32:dx"])
    if case["init_states"]:
        if case["cell"] == "lstm":
            gerrs["dh0"] = rel_err(init[0].grad, g["f32:dh0"])
            gerrs["dc0"] = rel_err(init[1].grad, g["f32:dc0"])
        else:
            gerrs["dh0"] = rel_err(init.grad, g["f32:dh0"])
    bad = {k: v for k, v in gerrs.items() if not v <= GRAD_TOL}
    assert not bad, "gradient rel err above %.0e: %s" % (GRAD_TOL, bad)
    # tie-breaker: we must not be further from the FP64 reference than the bar either
    if full:
        assert rel_err(out, g["f64:out"]) <= FWD_TOL
