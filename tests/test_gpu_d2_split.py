"""dX-only ("split") BPTT variants of the two-core chains: the recurrent kernel writes delta_hh of every step
and keeps the W_ih-column / bias gradients in its own slot; the hh core gradients come from the dense accumulation
dW_hh^T = H_prev^T delta and one H-row projection onto the cores.  Reference semantics: autograd through
tensorized_rnn/lstm.py:23-41,101-135 and gru.py:25-50,104-136; bars 1e-5 forward / 1e-4 gradients.
Edge cases: T = 1 with and without an initial state (no (h_{t-1}, delta_t) rows at all / only the h0 rows), ragged last
row tile, non-zero initial states, and the fused variant (`split_kept = 0`) as the cross-check on the same inputs."""
import pytest
import torch

from tensorized_rnn_b200 import _lib
from helpers import FWD_TOL, GRAD_TOL, oracle, rel_err
from test_gpu_round2 import DEV, assert_grads, make_pair, options

pytestmark = pytest.mark.gpu

CASES = [
    # name, cell, I, B, T, with initial state
    ("lstm_rank_one_T1_no_state", "lstm", 1, 6, 1, False),
    ("lstm_rank_one_T1_state", "lstm", 1, 6, 1, True),
    ("gru_rank_one_T1_no_state", "gru", 1, 5, 1, False),
    ("gru_rank_one_state_ragged", "gru", 1, 9, 7, True),
    ("lstm_rank_one_state_ragged", "lstm", 1, 11, 6, True),
    ("gru_rank_one_long", "gru", 1, 16, 40, False),
    ("lstm_projected_input_state", "lstm", 40, 10, 6, True),
]


def run(cell, fwd, x, init, w_out, w_h, dev):
    x = x.to(dev)
    init = None if init is None else tuple(s.clone().to(dev).requires_grad_(True) for s in init)
    if cell == "lstm":
        out, (h, _) = fwd(x, init)
    else:
        out, h = fwd(x, None if init is None else init[0])
    ((out * w_out.to(dev)).sum() + (h * w_h.to(dev)).sum()).backward()
    return out.detach(), h.detach(), (None if init is None else [s.grad for s in init[: (2 if cell == "lstm" else 1)]])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_split_variants_match_oracle_and_fused_variant(case):
    name, cell, I, B, T, with_state = case
    H = 256
    layers, m = make_pair(cell, I, H, 1, 2, 4)
    g = torch.Generator().manual_seed(17)
    x = torch.rand(B, T, I, generator=g)
    init = (0.5 * torch.randn(B, H, generator=g), 0.5 * torch.randn(B, H, generator=g)) if with_state else None
    w_out, w_h = torch.randn(B, T, H, generator=g), torch.randn(B, H, generator=g)
    for p in oracle.flat_params(layers):
        p.grad = None
    ref_fwd = (lambda xs, st: oracle.lstm_forward(layers, xs, st)) if cell == "lstm" else (lambda xs, st: oracle.gru_forward(layers, xs, st))
    o_ref, h_ref, ds_ref = run(cell, ref_fwd, x, init, w_out, w_h, "cpu")
    g_ref = [p.grad.clone() for p in oracle.flat_params(layers)]
    results = {}
    for split in (1, 0):
        for p in m.parameters():
            p.grad = None
        with options(split_kept=split):
            plan = _lib.describe_plan(m.spec().desc(B, T))
            assert ("split" in plan[1]["bwd_kernel"]) == bool(split), plan[1]
            out, h, ds = run(cell, lambda xs, st: m(xs, st), x, init, w_out, w_h, DEV)
            torch.cuda.synchronize()
        assert rel_err(out, o_ref) <= FWD_TOL and rel_err(h, h_ref) <= FWD_TOL
        grads = [p.grad.clone() for p in m.flat_parameters()]
        # T = 1 without an initial state: h_{t-1} = 0, so the hh core gradients are exactly zero on both sides
        assert_grads([a for a, b in zip(grads, g_ref) if float(b.abs().sum()) > 0], [b for b in g_ref if float(b.abs().sum()) > 0])
        for a, b in zip(grads, g_ref):
            if float(b.abs().sum()) == 0:
                assert float(a.abs().max()) <= 1e-12
        if ds_ref is not None:
            for a, b in zip(ds, ds_ref):
                assert rel_err(a, b) <= GRAD_TOL
        results[split] = grads
    for a, b in zip(results[1], results[0]):
        if float(b.abs().sum()) > 0:
            assert rel_err(a, b) <= GRAD_TOL
