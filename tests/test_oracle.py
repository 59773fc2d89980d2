"""Pin the CPU oracle against fixtures produced by the reference itself (tests/golden/make_golden.py).

The oracle restates the reference op by op, so on the fixture inputs it must agree with the stored
reference outputs to FP32 round-off (observed: bit-exact); the bar asserted here is 1e-6 relative,
ten times tighter than the forward parity bar the CUDA path is held to.
"""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, golden_index, load_golden, oracle, rel_err, state_dict_from_golden

CASES = golden_index()
SHORT = [c for c in CASES if c["seq_len"] <= 16]
LONG = [c for c in CASES if c["seq_len"] > 16]


def _oracle_run(case, g, dtype=torch.float32, grads=True):
    sd = state_dict_from_golden(g)
    layers = oracle.layers_from_state_dict(sd, case["num_layers"], dtype=dtype, requires_grad=grads)
    x = torch.from_numpy(g["x"]).to(dtype).requires_grad_(grads)
    lstm = case["cell"] == "lstm"
    init = None
    if case["init_states"]:
        h0 = torch.from_numpy(g["h0"]).to(dtype).requires_grad_(grads)
        init = (h0, torch.from_numpy(g["c0"]).to(dtype).requires_grad_(grads)) if lstm else h0
    if lstm:
        out, (h, c) = oracle.lstm_forward(layers, x, init)
    else:
        out, h = oracle.gru_forward(layers, x, init)
        c = None
    return layers, x, init, out, h, c


@pytest.mark.parametrize("case", SHORT, ids=[c["name"] for c in SHORT])
def test_oracle_forward_and_grads_match_reference(case):
    g = load_golden(case["name"])
    layers, x, init, out, h, c = _oracle_run(case, g)
    assert rel_err(out, g["f32:out"]) <= 1e-6
    assert rel_err(h, g["f32:hT"]) <= 1e-6
    loss = (out * torch.from_numpy(g["w_out"])).sum() + (h * torch.from_numpy(g["w_h"])).sum()
    if c is not None:
        assert rel_err(c, g["f32:cT"]) <= 1e-6
        loss = loss + (c * torch.from_numpy(g["w_c"])).sum()
    loss.backward()
    assert rel_err(x.grad, g["f32:dx"]) <= 1e-6
    # oracle parameter order == state_dict order per layer: ih cores, ih bias, hh cores, hh bias
    for li, p in enumerate(layers):
        for short, long in (("ih", "input_weights"), ("hh", "hidden_weights")):
            for k, core in enumerate(p[short + "_cores"]):
                key = "f32:grad:cell%d.%s.parameters.%d" % (li, long, k)
                assert rel_err(core.grad, g[key]) <= 1e-6, key
            if p[short + "_bias"] is not None:
                key = "f32:grad:cell%d.%s.bias" % (li, long)
                assert rel_err(p[short + "_bias"].grad, g[key]) <= 1e-6, key
    if case["init_states"]:
        h0 = init[0] if case["cell"] == "lstm" else init
        assert rel_err(h0.grad, g["f32:dh0"]) <= 1e-6


@pytest.mark.parametrize("case", LONG, ids=[c["name"] for c in LONG])
def test_oracle_long_sequence_forward(case):
    g = load_golden(case["name"])
    with torch.no_grad():
        _, _, _, out, h, c = _oracle_run(case, g, grads=False)
    assert rel_err(out[:, -1], g["f32:out_last"]) <= 1e-6
    assert rel_err(h, g["f32:hT"]) <= 1e-6


@pytest.mark.parametrize("case", SHORT[:4], ids=[c["name"] for c in SHORT[:4]])
def test_oracle_fp64_agrees_with_reference_fp64(case):
    g = load_golden(case["name"])
    with torch.no_grad():
        _, _, _, out, h, c = _oracle_run(case, g, dtype=torch.float64, grads=False)
    # the reference writes its FP64 steps into a default-dtype (FP32) `outputs` buffer (lstm.py:117),
    # so only the returned final state is true FP64
    assert rel_err(h, g["f64:hT"]) <= 1e-12
    assert rel_err(out, g["f64:out"]) <= 2e-7


def test_tt_matvec_equals_dense_matrix():
    torch.manual_seed(3)
    layers = oracle.random_layers("lstm", 40, 64, 1, 3, 4, seed=5, dtype=torch.float64)
    cores = layers[0]["ih_cores"]
    w = oracle.tt_dense(cores)
    x = torch.randn(7, 40, dtype=torch.float64)
    assert rel_err(oracle.tt_matvec(cores, x), x @ w.t()) <= 1e-12


def test_ttlinear_fixtures():
    with open(os.path.join(GOLDEN, "ttlinear_index.json")) as f:
        idx = json.load(f)
    for e in idx:
        g = dict(np.load(os.path.join(GOLDEN, e["name"] + ".npz")))
        cores = [torch.from_numpy(g["param:parameters.%d" % k]) for k in range(e["d"])]
        bias = torch.from_numpy(g["param:bias"]) if e["bias"] else None
        y = oracle.ttlinear(cores, bias, torch.from_numpy(g["x"]))
        assert rel_err(y, g["y"]) <= 1e-6


def test_zero_length_sequence_raises_like_reference():
    layers = oracle.random_layers("gru", 4, 8, 1, 2, 2)
    with pytest.raises(NameError):
        oracle.gru_forward(layers, torch.zeros(2, 0, 4))
