"""CPU oracle of the dense LSTM / GRU baselines.  TEST INFRASTRUCTURE ONLY (never imported by the product package).

Restates the reference's dense cells and sequence loops in plain torch-on-CPU ops:
LSTMCell.forward tensorized_rnn/lstm.py:23-41, LSTM.forward lstm.py:101-135, GRUCell.forward gru.py:25-50,
GRU.forward gru.py:104-136.  Pinned by tests/golden/dense_*.npz (tests/golden/make_golden_dense.py, generated from the
reference's own classes)."""
import torch


def layers_from_state_dict(sd, num_layers, requires_grad=False):
    layers = []
    for l in range(num_layers):
        p = {}
        for short, long in (("ih", "input_weights"), ("hh", "hidden_weights")):
            p["w_" + short] = sd["cell%d.%s.weight" % (l, long)].clone().requires_grad_(requires_grad)
            b = sd.get("cell%d.%s.bias" % (l, long))
            p["b_" + short] = None if b is None else b.clone().requires_grad_(requires_grad)
        layers.append(p)
    return layers


def flat_params(layers):
    out = []
    for p in layers:
        for k in ("w_ih", "b_ih", "w_hh", "b_hh"):
            if p[k] is not None:
                out.append(p[k])
    return out


def _lin(x, w, b):
    y = x @ w.t()
    return y if b is None else y + b


def lstm_forward(layers, x, init_states=None):
    """lstm.py:101-135: batch-first input, one (h, c) shared by all layers, step-major / layer-minor loop."""
    B, T, _ = x.shape
    H = layers[0]["w_hh"].shape[1]
    if init_states is None:
        init_states = (torch.zeros(B, H, dtype=x.dtype), torch.zeros(B, H, dtype=x.dtype))
    states = [init_states] * len(layers)
    outputs = []
    for t in range(T):
        inp = x[:, t, :]
        for l, p in enumerate(layers):
            h, c = states[l]
            g = _lin(inp, p["w_ih"], p["b_ih"]) + _lin(h, p["w_hh"], p["b_hh"])         # lstm.py:24
            i, f = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H])
            gg, o = torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
            c = f * c + i * gg
            h = o * torch.tanh(c)
            states[l] = (h, c)
            inp = h
        outputs.append(inp)
    return torch.stack(outputs, dim=1), states[-1]


def gru_forward(layers, x, init_states=None):
    """gru.py:104-136 with the cell of gru.py:25-50 (b_hn inside the reset product)."""
    B, T, _ = x.shape
    H = layers[0]["w_hh"].shape[1]
    if init_states is None:
        init_states = torch.zeros(B, H, dtype=x.dtype)
    states = [init_states] * len(layers)
    outputs = []
    for t in range(T):
        inp = x[:, t, :]
        for l, p in enumerate(layers):
            h = states[l]
            a, u = _lin(inp, p["w_ih"], p["b_ih"]), _lin(h, p["w_hh"], p["b_hh"])
            r = torch.sigmoid(a[:, :H] + u[:, :H])
            z = torch.sigmoid(a[:, H:2 * H] + u[:, H:2 * H])
            n = torch.tanh(a[:, 2 * H:] + r * u[:, 2 * H:])
            h = (1 - z) * n + z * h
            states[l] = h
            inp = h
        outputs.append(inp)
    return torch.stack(outputs, dim=1), states[-1]
