"""CPU oracle for the TT-LSTM / TT-GRU hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement, in plain torch-on-CPU ops, of the
algorithm the reference (onucharles/tensorized-rnn) runs for the path that
`tensorized_rnn_b200` replaces.  It is *never* imported by the product
package: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it, and there only as the
checker / the timed CPU baseline.

Parity pinning: the reference ships no golden vectors, so
this restatement is pinned against outputs of the reference itself, generated
in the build container by `tests/golden/make_golden.py` (which imports
/root/reference) and committed under `tests/golden/`.  `tests/test_oracle.py`
replays every fixture through this file.

Every function cites the reference file:line it follows.  The op sequence per
timestep deliberately mirrors the reference (one einsum per TT core, a
`.contiguous()` reshuffle between cores, a Python loop over steps and layers,
an in-place `outputs[:, t] = h` write) so that timing this file on host cores
is a fair stand-in for timing the reference where the reference itself cannot
travel (the GPU box has no /root/reference).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# Shape selection (construction time only)
# --------------------------------------------------------------------------
def auto_shape(n: int, d: int = 3) -> List[int]:
    """Max-entropy ascending d-factorisation of n.

    Follows t3nsor/utils.py:39-64 (`_get_all_factors`: prime factors padded
    with ones, every multiset partition into d blocks, block products sorted
    ascending, de-duplicated through a set) and t3nsor/utils.py:71-81
    (`auto_shape`: scipy entropy of each candidate, argmax).
    """
    from scipy.stats import entropy
    from sympy.ntheory import factorint
    from sympy.utilities.iterables import multiset_partitions

    primes: List[int] = []
    for p, mult in factorint(n).items():
        primes += [p] * mult
    if len(primes) < d:
        primes = primes + [1] * (d - len(primes))
    cands = [tuple(sorted(np.prod(block) for block in part))
             for part in multiset_partitions(primes, d)]
    cands = list(set(cands))
    scores = [entropy(c) for c in cands]
    return [int(v) for v in cands[int(np.argmax(scores))]]


def tt_shape(in_features: int, out_features: int, n_cores: int, n_gates: int,
             new_core: Optional[str] = None) -> List[List[int]]:
    """[in_quant, out_quant] of the concat-gates TT matrix.

    Follows tensorized_rnn/rnn_utils.py:20-36: without `new_core` the gates
    are folded into the output dimension before factorising; with
    'first'/'last' an extra (1 x n_gates) core is prepended / appended.
    """
    assert new_core in (None, "first", "last")
    if new_core is None:
        out_features = out_features * n_gates
    in_q = auto_shape(in_features, n_cores)
    out_q = auto_shape(out_features, n_cores)
    if new_core == "first":
        in_q, out_q = [1] + in_q, [n_gates] + out_q
    elif new_core == "last":
        in_q, out_q = in_q + [1], out_q + [n_gates]
    return [in_q, out_q]


def glorot_core_std(in_quant: Sequence[int], out_quant: Sequence[int], tt_rank: int) -> float:
    """Std-dev of every core entry under the reference's Glorot TT init.

    Follows t3nsor/initializers.py:286-300 (lambda = 2/(n_in+n_out)) and
    :218-282 (core_std = stddev**(1/d) * prod(rank_k ** (-1/(2d)))).
    """
    d = len(in_quant)
    ranks = np.array([1] + [tt_rank] * (d - 1) + [1], dtype=np.float64)
    lamb = 2.0 / (float(np.prod(in_quant)) + float(np.prod(out_quant)))
    return float(np.sqrt(lamb) ** (1.0 / d) * np.prod(ranks ** (-1.0 / (2 * d))))


# --------------------------------------------------------------------------
# TT-matrix x dense
# --------------------------------------------------------------------------
def tt_matvec(cores: Sequence[Tensor], x: Tensor) -> Tensor:
    """y (B, M) = x (B, N) @ W^T with W (M x N) held as TT cores.

    `cores[k]` has the *stored* layout of the reference's `weight_t`
    parameters, (r_k, i_k, j_k, r_{k+1}) -- output mode before input mode
    (t3nsor/layers.py:113 + t3nsor/ops.py:47-51).

    Follows t3nsor/ops.py:54-93 as called from t3nsor/layers.py:121-127: the
    batch is moved next to the leading input modes, then cores are absorbed
    right-to-left, each by one einsum over (j_k, r_{k+1}) followed by a
    contiguous reshuffle that exposes j_{k-1}.
    """
    d = len(cores)
    in_q = [int(c.shape[2]) for c in cores]
    ranks = [int(c.shape[0]) for c in cores] + [1]
    n_in = int(np.prod(in_q))
    if x.shape[1] != n_in:
        raise ValueError("input width %d does not match TT column size %d" % (x.shape[1], n_in))
    batch = x.shape[0]
    # ops.py:78-79 -- (B*j0..j_{d-2}, j_{d-1}, 1)
    data = x.contiguous().view(-1, in_q[-1], 1)
    for k in reversed(range(d)):
        # ops.py:85 -- contract (j_k, r_{k+1}); result (i_k, rho, r_k)
        data = torch.einsum("aijb,rjb->ira", cores[k], data)
        if k > 0:
            # ops.py:89-90
            data = data.contiguous().view(-1, in_q[k - 1], ranks[k])
    m_out = int(np.prod([int(c.shape[1]) for c in cores]))
    # ops.py:93 gives (M, B); layers.py:125-127 transposes back
    return data.view(m_out, batch).transpose(0, 1)


def ttlinear(cores: Sequence[Tensor], bias: Optional[Tensor], x: Tensor) -> Tensor:
    """t3nsor/layers.py:121-127: TT matvec, then the optional bias add."""
    y = tt_matvec(cores, x)
    return y if bias is None else y + bias


def ttlinear_set(gates: Sequence[Tuple[Sequence[Tensor], Optional[Tensor]]], x: Tensor) -> Tensor:
    """tensorized_rnn/tt_linearset.py:27-38: n_gates independent TTLinear maps ("naive TT"), outputs
    concatenated column-wise (gate g fills columns [g*H, (g+1)*H))."""
    return torch.cat([ttlinear(cores, bias, x) for cores, bias in gates], dim=1)


def _proj(p: "LayerParams", short: str, x: Tensor) -> Tensor:
    """ih / hh projection of one layer: concat-gates TTLinear, or the naive per-gate set."""
    if short + "_gates" in p:
        return ttlinear_set(p[short + "_gates"], x)
    return ttlinear(p[short + "_cores"], p[short + "_bias"], x)


def av_norm(t: Tensor, average_logs: bool = False) -> Tensor:
    """tensorized_rnn/rnn_utils.py:217-226: batch mean of ||t_b||^2 (or of its log)."""
    norms = (t ** 2).sum(list(range(1, t.dim())))
    if average_logs:
        norms = torch.log(norms)
    return norms.mean()


class StepLog(object):
    """What `log_grads=True` records for one state variable of one layer over one minibatch
    (rnn_utils.py:127-172): per-step activation norms in time order, gradient norms prepended as
    backward walks the sequence in reverse."""

    def __init__(self):
        self.act, self.log_act, self.grad, self.log_grad = [], [], [], []

    def forward(self, t: Tensor) -> None:
        with torch.no_grad():
            self.act.append(av_norm(t.detach()))
            self.log_act.append(av_norm(t.detach(), True))

    def backward(self, g: Tensor) -> None:
        with torch.no_grad():
            self.grad.insert(0, av_norm(g.detach()))
            self.log_grad.insert(0, av_norm(g.detach(), True))

    def stacked(self) -> Dict[str, Tensor]:
        return {k: torch.stack(getattr(self, k)) for k in ("act", "log_act", "grad", "log_grad")}


def tt_dense(cores: Sequence[Tensor]) -> Tensor:
    """Densify the stored TT cores into the (M x N) matrix W (test helper).

    Equivalent to t3nsor/tensor_train.py:116-146 applied to contiguous copies
    of the cores (the reference's own `.full()` raises on the transposed
    views).
    """
    res = cores[0].reshape(-1, cores[0].shape[-1])            # (i0*j0, r1)
    modes = [(int(cores[0].shape[1]), int(cores[0].shape[2]))]
    for c in cores[1:]:
        res = res @ c.reshape(c.shape[0], -1)
        res = res.reshape(-1, c.shape[-1])
        modes.append((int(c.shape[1]), int(c.shape[2])))
    d = len(cores)
    res = res.reshape([v for ij in modes for v in ij])
    perm = list(range(0, 2 * d, 2)) + list(range(1, 2 * d, 2))
    m = int(np.prod([ij[0] for ij in modes]))
    n = int(np.prod([ij[1] for ij in modes]))
    return res.permute(perm).reshape(m, n)


# --------------------------------------------------------------------------
# Cells
# --------------------------------------------------------------------------
LayerParams = Dict[str, object]   # {"ih_cores": [...], "ih_bias": T|None, "hh_cores": [...], "hh_bias": T|None}


def lstm_cell(p: LayerParams, x: Tensor, h: Tensor, c: Tensor) -> Tuple[Tensor, Tensor]:
    """tensorized_rnn/lstm.py:23-41 with TT weights (tt_lstm.py:16-40).

    Gate order i, f, g, o; both the ih and the hh TTLinear carry a bias.
    """
    hid = h.shape[1]
    gates = _proj(p, "ih", x) + _proj(p, "hh", h)
    i = torch.sigmoid(gates[:, :hid])
    f = torch.sigmoid(gates[:, hid:2 * hid])
    g = torch.tanh(gates[:, 2 * hid:3 * hid])
    o = torch.sigmoid(gates[:, 3 * hid:])
    c_new = f * c + i * g
    h_new = o * torch.tanh(c_new)
    return h_new, c_new


def gru_cell(p: LayerParams, x: Tensor, h: Tensor) -> Tensor:
    """tensorized_rnn/gru.py:25-50 with TT weights (gru.py:148-172).

    Gate order r, z, n; the hidden bias of the n block sits inside the reset
    product.
    """
    hid = h.shape[1]
    a = _proj(p, "ih", x)
    u = _proj(p, "hh", h)
    r = torch.sigmoid(a[:, :hid] + u[:, :hid])
    z = torch.sigmoid(a[:, hid:2 * hid] + u[:, hid:2 * hid])
    n = torch.tanh(a[:, 2 * hid:] + r * u[:, 2 * hid:])
    return (1 - z) * n + z * h


# --------------------------------------------------------------------------
# Sequence loops
# --------------------------------------------------------------------------
def lstm_forward(layers: Sequence[LayerParams], x: Tensor,
                 init_states: Optional[Tuple[Tensor, Tensor]] = None,
                 logs: Optional[Dict[str, "StepLog"]] = None
                 ) -> Tuple[Tensor, Tuple[Tensor, Tensor]]:
    """tensorized_rnn/lstm.py:101-135.

    x is batch-first (B, T, I).  One (h, c) pair seeds *every* layer
    (lstm.py:120-121); the loop is step-major / layer-minor (lstm.py:123-133);
    the return is (outputs (B, T, H), (h_T, c_T)) of the last layer only.
    """
    batch, seq_len, _ = x.shape
    hid = _hidden_size(layers[0])
    outputs = torch.zeros(batch, seq_len, hid, dtype=x.dtype)
    if init_states is None:
        h0 = torch.zeros(batch, hid, dtype=x.dtype)
        c0 = torch.zeros(batch, hid, dtype=x.dtype)
    else:
        h0, c0 = init_states
    state = [(h0, c0)] * len(layers)
    if seq_len == 0:
        raise NameError("reference raises NameError for T == 0 (lstm.py:135)")
    for t in range(seq_len):
        inp = x[:, t, :]
        for li, p in enumerate(layers):
            h, c = state[li]
            inp, c_new = lstm_cell(p, inp, h, c)
            if logs is not None:          # lstm.py:35-39, 66-80 (log_grads hooks)
                hl = logs.setdefault("hidden_%d" % li, StepLog())
                cl = logs.setdefault("cell_%d" % li, StepLog())
                hl.forward(inp)
                cl.forward(c_new)
                if inp.requires_grad:
                    inp.register_hook(hl.backward)
                    c_new.register_hook(cl.backward)
            state[li] = (inp, c_new)
        outputs[:, t, :] = inp                      # lstm.py:133 (in-place write)
    return outputs, (inp, c_new)


def gru_forward(layers: Sequence[LayerParams], x: Tensor,
                init_states: Optional[Tensor] = None,
                logs: Optional[Dict[str, "StepLog"]] = None) -> Tuple[Tensor, Tensor]:
    """tensorized_rnn/gru.py:104-136 (same loop shape as the LSTM, no cell state)."""
    batch, seq_len, _ = x.shape
    hid = _hidden_size(layers[0])
    outputs = torch.zeros(batch, seq_len, hid, dtype=x.dtype)
    h0 = torch.zeros(batch, hid, dtype=x.dtype) if init_states is None else init_states
    state = [h0] * len(layers)
    if seq_len == 0:
        raise NameError("reference raises NameError for T == 0 (gru.py:136)")
    for t in range(seq_len):
        inp = x[:, t, :]
        for li, p in enumerate(layers):
            inp = gru_cell(p, inp, state[li])
            if logs is not None:          # gru.py:47-48, 76-85
                hl = logs.setdefault("hidden_%d" % li, StepLog())
                hl.forward(inp)
                if inp.requires_grad:
                    inp.register_hook(hl.backward)
            state[li] = inp
        outputs[:, t, :] = inp                      # gru.py:134
    return outputs, inp


def _hidden_size(p: LayerParams) -> int:
    cores = p["hh_gates"][0][0] if "hh_gates" in p else p["hh_cores"]
    return int(np.prod([int(c.shape[2]) for c in cores]))


# --------------------------------------------------------------------------
# state_dict <-> oracle parameter lists
# --------------------------------------------------------------------------
def layers_from_state_dict(sd: Dict[str, Tensor], num_layers: int, dtype=None,
                           requires_grad: bool = False) -> List[LayerParams]:
    """Rebuild oracle parameter lists from a reference-format state_dict.

    Key names are the reference's:
    `cell{l}.input_weights.parameters.{k}`, `cell{l}.input_weights.bias`, and
    the same under `hidden_weights`.
    """
    def grab(prefix):
        cores = []
        k = 0
        while "%s.parameters.%d" % (prefix, k) in sd:
            t = sd["%s.parameters.%d" % (prefix, k)].detach().clone().contiguous()
            if dtype is not None:
                t = t.to(dtype)
            cores.append(t.requires_grad_(requires_grad))
            k += 1
        bias = None
        if prefix + ".bias" in sd:
            bias = sd[prefix + ".bias"].detach().clone()
            if dtype is not None:
                bias = bias.to(dtype)
            bias = bias.requires_grad_(requires_grad)
        return cores, bias

    layers: List[LayerParams] = []
    for li in range(num_layers):
        p: LayerParams = {}
        for short, long in (("ih", "input_weights"), ("hh", "hidden_weights")):
            base = "cell%d.%s" % (li, long)
            if base + ".gates.0.parameters.0" in sd:
                # naive form (tt_linearset.py): `gates.{g}` (the same tensors are also listed as `gate{g}`)
                gates = []
                g = 0
                while "%s.gates.%d.parameters.0" % (base, g) in sd:
                    gates.append(grab("%s.gates.%d" % (base, g)))
                    g += 1
                p[short + "_gates"] = gates
            else:
                p[short + "_cores"], p[short + "_bias"] = grab(base)
        layers.append(p)
    return layers


def flat_params(layers: Sequence[LayerParams]) -> List[Tensor]:
    """Parameters in blob order: per layer ih cores, ih bias, hh cores, hh bias."""
    out: List[Tensor] = []
    for p in layers:
        for short in ("ih", "hh"):
            pairs = p[short + "_gates"] if short + "_gates" in p else [(p[short + "_cores"], p[short + "_bias"])]
            for cores, bias in pairs:
                out += list(cores)
                if bias is not None:
                    out.append(bias)
    return out


def random_layers(cell: str, input_size: int, hidden_size: int, num_layers: int, n_cores: int,
                  tt_rank: int, bias: bool = True, seed: int = 0, dtype=torch.float32,
                  requires_grad: bool = False, scale: float = 1.0) -> List[LayerParams]:
    """Random cores with the reference's init distribution (section 8a-2); for tests and
    benches that cannot import the reference.  `scale` multiplies the Glorot core std."""
    g = torch.Generator().manual_seed(seed)
    n_gates = 4 if cell == "lstm" else 3
    layers: List[LayerParams] = []
    for li in range(num_layers):
        p: LayerParams = {}
        for short, n_in in (("ih", input_size if li == 0 else hidden_size), ("hh", hidden_size)):
            in_q, out_q = tt_shape(n_in, hidden_size, n_cores, n_gates)
            std = glorot_core_std(in_q, out_q, tt_rank) * scale
            d = len(in_q)
            ranks = [1] + [tt_rank] * (d - 1) + [1]
            cores = []
            for k in range(d):
                t = torch.randn(ranks[k], out_q[k], in_q[k], ranks[k + 1], generator=g, dtype=torch.float64) * std
                cores.append(t.to(dtype).requires_grad_(requires_grad))
            p[short + "_cores"] = cores
            if bias:
                b = (1e-3 * torch.ones(n_gates * hidden_size, dtype=dtype))
                # break the all-equal bias so that bias-gradient bugs are visible
                b = b + 1e-2 * torch.randn(n_gates * hidden_size, generator=g, dtype=torch.float64).to(dtype)
                p[short + "_bias"] = b.requires_grad_(requires_grad)
            else:
                p[short + "_bias"] = None
        layers.append(p)
    return layers
