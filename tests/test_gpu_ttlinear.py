"""GPU parity of the stand-alone TTLinear and of the two caller patterns of the reference
(MNIST classifier head, GE2E speaker-encoder head) built from the B200 modules."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import tensorized_rnn_b200 as tr
from helpers import FWD_TOL, GOLDEN, GRAD_TOL, oracle, quiet, rel_err

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN, "ttlinear_index.json")) as f:
    LIN = json.load(f)


@pytest.mark.parametrize("e", LIN, ids=[e["name"] for e in LIN])
def test_ttlinear_matches_reference_fixture(e):
    dev = torch.device("cuda:0")
    g = dict(np.load(os.path.join(GOLDEN, e["name"] + ".npz")))
    lin = quiet(tr.TTLinear, in_features=e["in_features"], out_features=e["out_features"], bias=e["bias"],
                auto_shapes=True, d=e["d"], tt_rank=e["tt_rank"])
    assert lin.shape == e["shape"]
    lin.load_state_dict({k[len("param:"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param:")})
    lin = lin.to(dev)
    x = torch.from_numpy(g["x"]).to(dev).requires_grad_(True)
    y = lin(x)
    (y * torch.from_numpy(g["w"]).to(dev)).sum().backward()
    assert rel_err(y, g["y"]) <= FWD_TOL
    assert rel_err(x.grad, g["dx"]) <= GRAD_TOL
    for name, p in lin.named_parameters():
        assert rel_err(p.grad, g["grad:" + name]) <= GRAD_TOL, name


def test_ttlinear_many_rows_and_leading_dims():
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    lin = quiet(tr.TTLinear, in_features=256, out_features=1024, bias=True, d=3, tt_rank=8).to(dev)
    cores = [c.detach().cpu() for c in lin.weight_t.tt_cores]
    x = torch.randn(5, 37, 256)
    y = lin(x.to(dev))
    ref = oracle.ttlinear(cores, lin.bias.detach().cpu(), x.reshape(-1, 256)).reshape(5, 37, 1024)
    assert y.shape == (5, 37, 1024)
    assert rel_err(y, ref) <= FWD_TOL


def test_mnist_classifier_pattern():
    """rnn -> out[:, -1] -> TTLinear -> log_softmax  (reference mnist_classifier.py:50-57)."""
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    rnn = quiet(tr.TTGRU, 28, 64, 1, torch.device("cpu"), n_cores=2, tt_rank=4)
    head = quiet(tr.TTLinear, in_features=64, out_features=10, bias=True, auto_shapes=True, d=2, tt_rank=4)
    layers = oracle.layers_from_state_dict(rnn.state_dict(), 1, requires_grad=True)
    hc = [c.detach().clone().requires_grad_(True) for c in head.weight_t.tt_cores]
    hb = head.bias.detach().clone().requires_grad_(True)
    x = torch.rand(9, 28, 28)
    tgt = torch.randint(0, 10, (9,))
    o_ref, _ = oracle.gru_forward(layers, x)
    loss_ref = F.nll_loss(F.log_softmax(oracle.ttlinear(hc, hb, o_ref[:, -1, :]), dim=1), tgt)
    loss_ref.backward()
    rnn, head = rnn.to(dev), head.to(dev)
    out, _ = rnn(x.to(dev))
    logp = F.log_softmax(head(out[:, -1, :]), dim=1)
    loss = F.nll_loss(logp, tgt.to(dev))
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 1e-5 * abs(float(loss_ref.detach()))
    for p, ref in zip(rnn.flat_parameters(), oracle.flat_params(layers)):
        assert rel_err(p.grad, ref.grad) <= GRAD_TOL
    for p, ref in zip(head.weight_t.tt_cores, hc):
        assert rel_err(p.grad, ref.grad) <= GRAD_TOL
    assert rel_err(head.bias.grad, hb.grad) <= GRAD_TOL


def test_speaker_encoder_pattern():
    """rnn -> last_hidden -> TTLinear -> ReLU -> L2 norm  (reference speaker_encoder.py:80-89)."""
    dev = torch.device("cuda:0")
    torch.manual_seed(12)
    rnn = quiet(tr.TTLSTM, 40, 64, 3, torch.device("cpu"), n_cores=3, tt_rank=4)
    head = quiet(tr.TTLinear, in_features=64, out_features=64, bias=True, auto_shapes=True, d=3, tt_rank=4)
    layers = oracle.layers_from_state_dict(rnn.state_dict(), 3)
    hc = [c.detach().clone() for c in head.weight_t.tt_cores]
    x = torch.rand(12, 20, 40)
    with torch.no_grad():
        _, (h_ref, _) = oracle.lstm_forward(layers, x)
        e_ref = torch.relu(oracle.ttlinear(hc, head.bias.detach(), h_ref))
        e_ref = e_ref / torch.norm(e_ref, dim=1, keepdim=True)
        rnn, head = rnn.to(dev), head.to(dev)
        _, (h, _) = rnn(x.to(dev))
        e = torch.relu(head(h))
        e = e / torch.norm(e, dim=1, keepdim=True)
    assert rel_err(e, e_ref) <= FWD_TOL


def test_cell_single_step_api():
    dev = torch.device("cuda:0")
    torch.manual_seed(13)
    rnn = quiet(tr.TTLSTM, 12, 24, 1, torch.device("cpu"), n_cores=2, tt_rank=2)
    layers = oracle.layers_from_state_dict(rnn.state_dict(), 1)
    x, h0, c0 = torch.rand(5, 12), 0.2 * torch.randn(5, 24), 0.2 * torch.randn(5, 24)
    with torch.no_grad():
        h_ref, c_ref = oracle.lstm_cell(layers[0], x, h0, c0)
        rnn = rnn.to(dev)
        h, c = rnn.cell0(x.to(dev), h0.to(dev), c0.to(dev))
    assert rel_err(h, h_ref) <= FWD_TOL and rel_err(c, c_ref) <= FWD_TOL
