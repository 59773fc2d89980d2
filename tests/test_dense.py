"""Dense LSTM / GRU baselines: the CPU oracle against fixtures generated from the reference's `LSTM` / `GRU`
(tests/golden/make_golden_dense.py), the module mirror's state_dict contract, and the CUDA path against both.
Tolerances: 1e-5 forward, 1e-4 gradients, norm-wise relative."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

from helpers import FWD_TOL, GRAD_TOL, GOLDEN, ROOT, rel_err

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dense_oracle  # noqa: E402  (test infrastructure)

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "dense_*.npz")))


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    I, H, L, bias, B, T, with_init, want_dx = (int(v) for v in g["meta"])
    sd = {k[len("param:"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param:")}
    return g, str(g["cell"]), I, H, L, bool(bias), B, T, bool(with_init), bool(want_dx), sd


def run(cell, fwd, x, init, w_out, w_h):
    if cell == "lstm":
        out, (h, c) = fwd(x, init)
        loss = (out * w_out).sum() + (h * w_h).sum() + (c * w_h).sum() * 0.5
        return out, h, c, loss
    out, h = fwd(x, init)
    return out, h, None, (out * w_out).sum() + (h * w_h).sum()


def test_fixtures_exist():
    assert len(CASES) >= 6


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference(name):
    g, cell, I, H, L, bias, B, T, with_init, want_dx, sd = load(name)
    layers = dense_oracle.layers_from_state_dict(sd, L, requires_grad=True)
    x = torch.from_numpy(g["x"]).requires_grad_(want_dx)
    init = None
    if with_init:
        h0 = torch.from_numpy(g["h0"]).requires_grad_(True)
        init = (h0, torch.from_numpy(g["c0"]).requires_grad_(True)) if cell == "lstm" else h0
    fwd = (lambda a, b: dense_oracle.lstm_forward(layers, a, b)) if cell == "lstm" else (lambda a, b: dense_oracle.gru_forward(layers, a, b))
    out, h, c, loss = run(cell, fwd, x, init, torch.from_numpy(g["w_out"]), torch.from_numpy(g["w_h"]))
    loss.backward()
    assert rel_err(out, g["out"]) <= 1e-6 and rel_err(h, g["hT"]) <= 1e-6
    names = [k for k in g if k.startswith("grad:")]
    key_of = {"w_ih": "input_weights.weight", "b_ih": "input_weights.bias", "w_hh": "hidden_weights.weight", "b_hh": "hidden_weights.bias"}
    for l, p in enumerate(layers):
        for short, long in key_of.items():
            if p[short] is not None:
                assert rel_err(p[short].grad, g["grad:cell%d.%s" % (l, long)]) <= 1e-5, (l, short)
    assert len(names) == len(dense_oracle.flat_params(layers))
    if want_dx:
        assert rel_err(x.grad, g["dx"]) <= 1e-5


@pytest.mark.parametrize("name", CASES)
def test_module_mirror_state_dict_contract(name):
    import tensorized_rnn_b200 as tr
    g, cell, I, H, L, bias, B, T, with_init, want_dx, sd = load(name)
    m = (tr.LSTM if cell == "lstm" else tr.GRU)(I, H, L, torch.device("cpu"), bias=bias)
    own = m.state_dict()
    assert list(own.keys()) == list(sd.keys())
    for k in sd:
        assert tuple(own[k].shape) == tuple(sd[k].shape)
    m.load_state_dict(sd)
    assert m.param_count() == sum(v.numel() for v in sd.values())
    with pytest.raises(RuntimeError):                       # no CPU path
        m(torch.rand(2, 3, I))


def test_dense_rejects_log_grads():
    import tensorized_rnn_b200 as tr
    with pytest.raises(NotImplementedError):
        tr.LSTM(8, 128, 1, torch.device("cpu"), log_grads=True)


def _gpu_case(cell, I, H, L, bias, x, init, w_out, w_h, sd, want_dx):
    import tensorized_rnn_b200 as tr
    dev = "cuda:0"
    m = (tr.LSTM if cell == "lstm" else tr.GRU)(I, H, L, torch.device("cpu"), bias=bias)
    m.load_state_dict(sd)
    m = m.to(dev)
    xd = x.detach().to(dev).requires_grad_(want_dx)
    initd = None
    if init is not None:
        initd = tuple(t.detach().to(dev).requires_grad_(True) for t in init) if cell == "lstm" else init.detach().to(dev).requires_grad_(True)
    out, h, c, loss = run(cell, m, xd, initd, w_out.to(dev), w_h.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    return m, xd, initd, out, h, c


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_dense_matches_reference_fixture(name):
    g, cell, I, H, L, bias, B, T, with_init, want_dx, sd = load(name)
    x = torch.from_numpy(g["x"])
    init = None
    if with_init:
        init = (torch.from_numpy(g["h0"]), torch.from_numpy(g["c0"])) if cell == "lstm" else torch.from_numpy(g["h0"])
    m, xd, initd, out, h, c = _gpu_case(cell, I, H, L, bias, x, init, torch.from_numpy(g["w_out"]), torch.from_numpy(g["w_h"]), sd, want_dx)
    assert rel_err(out, g["out"]) <= FWD_TOL and rel_err(h, g["hT"]) <= FWD_TOL
    if cell == "lstm":
        assert rel_err(c, g["cT"]) <= FWD_TOL
    bad = {}
    for k, p in m.named_parameters():
        e = rel_err(p.grad, g["grad:" + k])
        if not e <= GRAD_TOL:
            bad[k] = e
    if want_dx:
        e = rel_err(xd.grad, g["dx"])
        if not e <= GRAD_TOL:
            bad["dx"] = e
    if with_init:
        e = rel_err((initd[0] if cell == "lstm" else initd).grad, g["dh0"])
        if not e <= GRAD_TOL:
            bad["dh0"] = e
        if cell == "lstm":
            e = rel_err(initd[1].grad, g["dc0"])
            if not e <= GRAD_TOL:
                bad["dc0"] = e
    assert not bad, bad


@pytest.mark.gpu
@pytest.mark.parametrize("cell,I,H,L,B,T", [("lstm", 40, 256, 3, 160, 12), ("gru", 40, 256, 2, 130, 10), ("lstm", 128, 256, 1, 300, 8)])
def test_cuda_dense_tensor_core_sizes_match_oracle(cell, I, H, L, B, T):
    """>= 128 rows per step and >= 256 rows per reduction: the per-step GEMMs and the batched projections / weight
    gradients run on the tcgen05 kernels."""
    import tensorized_rnn_b200 as tr
    from tensorized_rnn_b200 import _lib
    torch.manual_seed(5)
    ref = (tr.LSTM if cell == "lstm" else tr.GRU)(I, H, L, torch.device("cpu"))
    sd = ref.state_dict()
    layers = dense_oracle.layers_from_state_dict(sd, L, requires_grad=True)
    g = torch.Generator().manual_seed(6)
    x = torch.rand(B, T, I, generator=g)
    w_out, w_h = torch.randn(B, T, H, generator=g), torch.randn(B, H, generator=g)
    want_dx = I % 128 == 0
    xr = x.clone().requires_grad_(want_dx)
    fwd = (lambda a, b: dense_oracle.lstm_forward(layers, a, b)) if cell == "lstm" else (lambda a, b: dense_oracle.gru_forward(layers, a, b))
    o_ref, h_ref, c_ref, loss = run(cell, fwd, xr, None, w_out, w_h)
    loss.backward()
    lib = _lib.load()
    lib.ttrnn_tc_launch_count(1)
    m, xd, _, out, h, c = _gpu_case(cell, I, H, L, True, x, None, w_out, w_h, sd, want_dx)
    assert int(lib.ttrnn_tc_launch_count(0)) > 0
    assert rel_err(out, o_ref) <= FWD_TOL and rel_err(h, h_ref) <= FWD_TOL
    bad = {}
    for p, r in zip(m.flat_parameters(), dense_oracle.flat_params(layers)):
        e = rel_err(p.grad, r.grad)
        if not e <= GRAD_TOL:
            bad[tuple(p.shape)] = e
    if want_dx and not rel_err(xd.grad, xr.grad) <= GRAD_TOL:
        bad["dx"] = rel_err(xd.grad, xr.grad)
    assert not bad, bad
