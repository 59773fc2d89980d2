"""GE2E head: the CPU oracle against fixtures generated from the reference's SpeakerEncoder
(tests/golden/make_golden_ge2e.py), and the CUDA head against both.  Tolerances: 1e-5 forward, 1e-4 gradients."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

from helpers import FWD_TOL, GRAD_TOL, GOLDEN, ROOT, rel_err

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ge2e_oracle  # noqa: E402  (test infrastructure)

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "ge2e_*.npz")))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def test_fixtures_exist():
    assert len(CASES) >= 4


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference(name):
    g = load(name)
    S, U, E = (int(v) for v in g["shape"])
    raw = torch.from_numpy(g["raw"]).requires_grad_(True)
    w = torch.from_numpy(g["w"]).requires_grad_(True)
    b = torch.from_numpy(g["b"]).requires_grad_(True)
    emb = ge2e_oracle.embed_normalize(raw)
    assert rel_err(emb, g["embeds"]) <= 1e-6
    sim = ge2e_oracle.similarity_matrix(emb.view(S, U, E), w, b)
    assert rel_err(sim, g["sim"]) <= 1e-6
    loss = ge2e_oracle.loss(emb.view(S, U, E), w, b)
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    loss.backward()
    assert rel_err(raw.grad, g["d_raw"]) <= 1e-5
    assert rel_err(w.grad, g["d_w"]) <= 1e-5 and abs(float(b.grad)) < 1e-6


def test_head_refuses_cpu_tensors():
    from tensorized_rnn_b200 import ge2e
    with pytest.raises(RuntimeError):
        ge2e.embed_normalize(torch.rand(4, 8))
    with pytest.raises(RuntimeError):
        ge2e.ge2e_loss(torch.rand(3, 2, 8), torch.tensor([10.]), torch.tensor([-5.]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_head_matches_reference_fixture(name):
    from tensorized_rnn_b200 import _lib, ge2e
    g = load(name)
    S, U, E = (int(v) for v in g["shape"])
    dev = "cuda:0"
    lib = _lib.load()
    lib.ttrnn_launch_count(1)
    raw = torch.from_numpy(g["raw"]).to(dev).requires_grad_(True)
    head = ge2e.GE2EHead(device=dev)
    with torch.no_grad():
        head.similarity_weight.copy_(torch.from_numpy(g["w"]))
        head.similarity_bias.copy_(torch.from_numpy(g["b"]))
    emb = head(raw)
    assert rel_err(emb, g["embeds"]) <= FWD_TOL
    sim = head.similarity_matrix(emb.view(S, U, E))
    assert rel_err(sim, g["sim"]) <= FWD_TOL
    loss = head.loss(emb.view(S, U, E))
    assert abs(float(loss) - float(g["loss"])) <= FWD_TOL * abs(float(g["loss"]))
    (3.0 * loss).backward()                                # a non-unit upstream gradient
    torch.cuda.synchronize()
    assert rel_err(raw.grad / 3.0, g["d_raw"]) <= GRAD_TOL
    assert rel_err(head.similarity_weight.grad / 3.0, g["d_w"]) <= GRAD_TOL
    # a bias added to every logit does not change a softmax: d loss / d bias is exactly 0 and the reference's stored value
    # is its own rounding noise (~1e-9), so this one is an absolute bound
    assert abs(float(g["d_b"])) < 1e-6 and abs(float(head.similarity_bias.grad)) < 1e-6
    assert int(lib.ttrnn_launch_count(0)) >= 9            # embed fwd/bwd + 3 loss fwd (x2 calls) + 4 loss bwd kernels


@pytest.mark.gpu
def test_cuda_head_after_ttlstm_matches_oracle_end_to_end():
    """TTLSTM -> TTLinear -> ReLU/L2 -> GE2E loss on the GPU against the oracle chain on the CPU, gradients down to the
    TT cores of the first layer (the training step of encoder/main.py without the CPU round trip)."""
    import tensorized_rnn_b200 as tr
    from tensorized_rnn_b200 import ge2e
    from helpers import oracle, quiet
    from test_gpu_static_paths import _sd_from_layers
    S, U, I, H, E, T = 6, 4, 40, 64, 32, 12
    layers = oracle.random_layers("lstm", I, H, 2, 3, 4, bias=True, seed=21, requires_grad=True)
    rnn = quiet(tr.TTLSTM, I, H, 2, torch.device("cpu"), n_cores=3, tt_rank=4)
    rnn.load_state_dict(_sd_from_layers(layers))
    lin = torch.nn.Linear(H, E)
    gen = torch.Generator().manual_seed(4)
    x = torch.rand(S * U, T, I, generator=gen)
    w, b = torch.tensor([10.], requires_grad=True), torch.tensor([-5.], requires_grad=True)
    # oracle chain
    _, (h_ref, _) = oracle.lstm_forward(layers, x)
    emb_ref = ge2e_oracle.embed_normalize(lin(h_ref))
    loss_ref = ge2e_oracle.loss(emb_ref.view(S, U, E), w, b)
    loss_ref.backward()
    g_ref = [p.grad.clone() for p in oracle.flat_params(layers)]
    glin_ref = lin.weight.grad.clone()
    lin.weight.grad = None
    lin.bias.grad = None
    # CUDA chain
    dev = "cuda:0"
    rnn = rnn.to(dev)
    lin_d = torch.nn.Linear(H, E).to(dev)
    lin_d.load_state_dict(lin.state_dict())
    head = ge2e.GE2EHead(device=dev)
    _, (h, _) = rnn(x.to(dev))
    emb = head(lin_d(h))
    loss = head.loss(emb.view(S, U, E))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_ref)) <= FWD_TOL * abs(float(loss_ref))
    for p, gr in zip(rnn.flat_parameters(), g_ref):
        assert rel_err(p.grad, gr) <= GRAD_TOL
    assert rel_err(lin_d.weight.grad, glin_ref) <= GRAD_TOL
    assert rel_err(head.similarity_weight.grad, w.grad) <= GRAD_TOL
