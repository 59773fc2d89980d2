"""Parity properties at BASELINE.json's FULL sizes (the CPU oracle cannot run these in test time).

Size-independent properties of the recurrence that any correct implementation must satisfy, checked on the
statically specialised kernels at the benchmark shapes:
  * batch independence: running two halves of the batch separately gives the same outputs, and the parameter
    gradients of the halves add up to those of the full batch (this is also what the DP all-reduce relies on);
  * determinism: two identical runs are bit-identical (gradient partials are reduced in a fixed order);
  * linearity of BPTT in the upstream gradient;
  * causality / state consistency: outputs of the first T' steps do not depend on later inputs, h_T == out[:, -1],
    and feeding (h, c) of a prefix as `init_states` of the remainder reproduces the single run (single layer);
  * invariance to the time-chunking of the ih projection.
Tolerances: 1e-5 forward, 1e-4 gradients (north_star), norm-wise relative.
"""
import pytest
import torch

import tensorized_rnn_b200 as tr
from tensorized_rnn_b200 import _lib
from helpers import FWD_TOL, GRAD_TOL, quiet, rel_err

pytestmark = pytest.mark.gpu


def _model(cell, I, H, L, d, r, seed=0):
    torch.manual_seed(seed)
    cls = tr.TTLSTM if cell == "lstm" else tr.TTGRU
    return quiet(cls, I, H, L, torch.device("cpu"), n_cores=d, tt_rank=r).to("cuda:0")


def _run(m, x, w_last=None, w_h=None, scale=1.0):
    for p in m.parameters():
        p.grad = None
    res = m(x)
    out = res[0]
    h = res[1][0] if isinstance(res[1], tuple) else res[1]
    loss = 0.0
    if w_last is not None:
        loss = loss + scale * (out[:, -1] * w_last).sum()
    if w_h is not None:
        loss = loss + scale * (h * w_h).sum()
    loss.backward()
    return out.detach(), h.detach(), [p.grad.clone() for p in m.parameters()]


@pytest.mark.parametrize("cfg", [("cfg1", "lstm", 1, 256, 1, 2, 4, 256, 784), ("cfg2", "gru", 1, 256, 1, 2, 4, 1024, 784),
                                 ("cfg3", "lstm", 40, 256, 3, 3, 8, 640, 160)], ids=lambda c: c[0])
def test_full_size_batch_independence_determinism_linearity(cfg):
    name, cell, I, H, L, d, r, B, T = cfg
    m = _model(cell, I, H, L, d, r, seed=3)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(B, T, I, device="cuda", generator=g)
    w_last = torch.randn(B, H, device="cuda", generator=g)
    w_h = torch.randn(B, H, device="cuda", generator=g)
    out, h, grads = _run(m, x, w_last, w_h)
    assert torch.isfinite(out).all() and all(torch.isfinite(gp).all() for gp in grads)
    assert torch.equal(h, out[:, -1])
    # determinism
    out2, h2, grads2 = _run(m, x, w_last, w_h)
    assert torch.equal(out, out2)
    for a, b in zip(grads, grads2):
        assert torch.equal(a, b)
    # linearity of backward in the upstream gradient
    _, _, grads3 = _run(m, x, w_last, w_h, scale=2.0)
    for a, b in zip(grads, grads3):
        assert rel_err(b, 2.0 * a) <= 1e-6
    # batch independence (different rows-per-CTA variants are selected for the halves)
    half = B // 2
    acc = [torch.zeros_like(gp) for gp in grads]
    outs = []
    for lo, hi in ((0, half), (half, B)):
        o, _, gs = _run(m, x[lo:hi], w_last[lo:hi], w_h[lo:hi])
        outs.append(o)
        for a, gp in zip(acc, gs):
            a += gp
    assert rel_err(torch.cat(outs), out) <= FWD_TOL
    for a, b in zip(acc, grads):
        assert rel_err(a, b) <= GRAD_TOL


def test_full_size_inference_causality_and_chunk_invariance():
    """cfg4 (speaker-encoder inference, 3 x TT-LSTM d4 r16, one GPU's 2048-utterance shard, T = 160)."""
    lib = _lib.load()
    m = _model("lstm", 40, 256, 3, 4, 16, seed=5)
    g = torch.Generator(device="cuda").manual_seed(12)
    x = torch.rand(2048, 160, 40, device="cuda", generator=g)
    with torch.no_grad():
        out, (h, c) = m(x)
        assert torch.isfinite(out).all()
        assert torch.equal(h, out[:, -1])
        out_p, (h_p, c_p) = m(x[:, :64])                    # causality: a prefix does not see the future
        assert rel_err(out_p, out[:, :64]) <= FWD_TOL
        lib.ttrnn_set_option(b"chunk_steps", 48)            # 160 = 48 + 48 + 48 + 16
        try:
            out_c, (h_c, c_c) = m(x)
        finally:
            lib.ttrnn_set_option(b"chunk_steps", 0)
        assert rel_err(out_c, out) <= FWD_TOL and rel_err(c_c, c) <= FWD_TOL


def test_full_batch_state_handoff_and_chunked_backward():
    """cfg5 shape (H = 1024, d4 r8, B = 4096) at a reduced T: prefix state handed over as `init_states` reproduces
    the single run; gradients are invariant to the time-chunking (split backward + dense ih route)."""
    lib = _lib.load()
    m = _model("lstm", 256, 1024, 1, 4, 8, seed=7)
    g = torch.Generator(device="cuda").manual_seed(13)
    B, T = 4096, 24
    x = torch.rand(B, T, 256, device="cuda", generator=g)
    w_last = torch.randn(B, 1024, device="cuda", generator=g)
    out, h, grads = _run(m, x, w_last, None)
    with torch.no_grad():
        o1, (h1, c1) = m(x[:, :10])
        o2, (h2, c2) = m(x[:, 10:], (h1, c1))
    assert rel_err(torch.cat([o1, o2], dim=1), out) <= FWD_TOL
    lib.ttrnn_set_option(b"chunk_steps", 7)
    try:
        out_c, _, grads_c = _run(m, x, w_last, None)
    finally:
        lib.ttrnn_set_option(b"chunk_steps", 0)
    assert rel_err(out_c, out) <= FWD_TOL
    for a, b in zip(grads_c, grads):
        assert rel_err(a, b) <= GRAD_TOL
