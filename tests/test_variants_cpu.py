"""Variants that run in cell-step mode (is_naive, new_core, log_grads): pin the oracle against fixtures
generated from the reference (tests/golden/make_golden_variants.py) and check the module mirror's
construction-time contract (state_dict keys / shapes, param_count) on CPU.  The CUDA parity of the same
fixtures is in test_gpu_variants.py."""
import json
import os

import numpy as np
import pytest
import torch

import tensorized_rnn_b200 as tr
from helpers import GOLDEN, load_golden, oracle, quiet, rel_err, state_dict_from_golden

with open(os.path.join(GOLDEN, "variants_index.json")) as f:
    VARIANTS = json.load(f)


def build_variant(case):
    tr.ActivGradLogger.reset()
    cls = tr.TTLSTM if case["cell"] == "lstm" else tr.TTGRU
    return quiet(cls, case["input_size"], case["hidden_size"], case["num_layers"], torch.device("cpu"),
                 n_cores=case["n_cores"], tt_rank=case["tt_rank"], bias=case["bias"], is_naive=case["is_naive"],
                 log_grads=case["log_grads"], new_core=case["new_core"])


@pytest.mark.parametrize("case", VARIANTS, ids=[c["name"] for c in VARIANTS])
def test_oracle_matches_reference_variant(case):
    g = load_golden("variant_" + case["name"])
    layers = oracle.layers_from_state_dict(state_dict_from_golden(g), case["num_layers"], requires_grad=True)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    lstm = case["cell"] == "lstm"
    init = None
    if case["init_states"]:
        h0 = torch.from_numpy(g["h0"]).requires_grad_(True)
        init = (h0, torch.from_numpy(g["c0"]).requires_grad_(True)) if lstm else h0
    logs = {} if case["log_grads"] else None
    if lstm:
        out, (h, c) = oracle.lstm_forward(layers, x, init, logs=logs)
        loss = (out * torch.from_numpy(g["w_out"])).sum() + (h * torch.from_numpy(g["w_h"])).sum() \
            + (c * torch.from_numpy(g["w_c"])).sum()
        assert rel_err(c, g["f32:cT"]) <= 1e-6
    else:
        out, h = oracle.gru_forward(layers, x, init, logs=logs)
        loss = (out * torch.from_numpy(g["w_out"])).sum() + (h * torch.from_numpy(g["w_h"])).sum()
    loss.backward()
    assert rel_err(out, g["f32:out"]) <= 1e-6
    assert rel_err(h, g["f32:hT"]) <= 1e-6
    assert rel_err(x.grad, g["f32:dx"]) <= 1e-6
    # parameters in oracle order vs the reference's named gradients
    names = [k[len("f32:grad:"):] for k in g if k.startswith("f32:grad:")]
    by_name = {}
    for li, p in enumerate(layers):
        for short, long in (("ih", "input_weights"), ("hh", "hidden_weights")):
            if short + "_gates" in p:
                for gi, (cores, bias) in enumerate(p[short + "_gates"]):
                    for k, cten in enumerate(cores):
                        by_name["cell%d.%s.gates.%d.parameters.%d" % (li, long, gi, k)] = cten
                    if bias is not None:
                        by_name["cell%d.%s.gates.%d.bias" % (li, long, gi)] = bias
            else:
                for k, cten in enumerate(p[short + "_cores"]):
                    by_name["cell%d.%s.parameters.%d" % (li, long, k)] = cten
                if p[short + "_bias"] is not None:
                    by_name["cell%d.%s.bias" % (li, long)] = p[short + "_bias"]
    checked = 0
    for n in names:
        key = n.replace(".gate0.", ".gates.0.").replace(".gate1.", ".gates.1.").replace(".gate2.", ".gates.2.") \
               .replace(".gate3.", ".gates.3.")
        assert rel_err(by_name[key].grad, g["f32:grad:" + n]) <= 1e-6, n
        checked += 1
    assert checked >= len(by_name)
    if case["log_grads"]:
        for var, lg in logs.items():
            for qnt, vec in lg.stacked().items():
                ref = g["log:%s:%s" % (var, qnt)]
                assert ref.shape == (1, case["seq_len"])
                assert rel_err(vec, ref[0]) <= 1e-5, (var, qnt)


@pytest.mark.parametrize("case", VARIANTS, ids=[c["name"] for c in VARIANTS])
def test_variant_module_contract(case):
    """Same state_dict keys (in order) and shapes as the reference module; reference parameters load strictly."""
    g = load_golden("variant_" + case["name"])
    m = build_variant(case)
    sd_ref = state_dict_from_golden(g)
    assert list(m.state_dict().keys()) == case["state_dict_keys"]
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd_ref[k].shape), k
    m.load_state_dict(sd_ref, strict=True)
    n_ref = sum(int(np.prod(v.shape)) for k, v in sd_ref.items() if ".gate" not in k or ".gates." in k)
    assert m.param_count() == n_ref
    assert m.cell_step_mode == bool(case["is_naive"] or case["log_grads"])
    tr.ActivGradLogger.reset()


def test_variant_cpu_tensors_raise():
    """No CPU fallback in cell-step mode either."""
    tr.ActivGradLogger.reset()
    m = quiet(tr.TTGRU, 12, 24, 1, torch.device("cpu"), n_cores=2, tt_rank=2, is_naive=True)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 3, 12))


def test_logger_registry_and_aggregation():
    """ActivGradLogger bookkeeping (reference rnn_utils.py:174-215) on hand-fed values."""
    tr.ActivGradLogger.reset()
    lg = tr.ActivGradLogger("hidden_0")
    with pytest.raises(AssertionError):
        tr.ActivGradLogger("hidden_0")
    fwd, bwd = lg.create_hooks(0)
    for mb in range(2):
        for t in range(3):
            fwd(None, None, torch.full((2, 4), float(t + 1 + mb)))
        for t in reversed(range(3)):
            bwd(torch.full((2, 4), float(t + 1)))
        tr.ActivGradLogger.end_minibatch()
    tr.ActivGradLogger.end_epoch()
    logs = tr.ActivGradLogger.get_logs()
    assert logs[("hidden_0", "act")].shape == (1, 3)
    # ||t||^2 of a (4,) vector of value v is 4 v^2; minibatch mean of v = (t+1), (t+2)
    want = torch.tensor([[(4 * 1 + 4 * 4) / 2, (4 * 4 + 4 * 9) / 2, (4 * 9 + 4 * 16) / 2]])
    assert torch.allclose(logs[("hidden_0", "act")], want)
    assert torch.allclose(logs[("hidden_0", "grad")], torch.tensor([[4.0, 16.0, 36.0]]))
    assert torch.allclose(logs[("hidden_0", "log_grad")], torch.log(torch.tensor([[4.0, 16.0, 36.0]])))
    tr.ActivGradLogger.reset()
