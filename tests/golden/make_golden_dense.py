"""Generate dense LSTM / GRU fixtures from the reference (run in the build container, where /root/reference exists):

    python tests/golden/make_golden_dense.py

Writes tests/golden/dense_*.npz: parameters (state_dict), input, optional initial states, outputs, final states and every
gradient of a fixed scalar loss, all computed by the reference's own `LSTM` / `GRU` classes (tensorized_rnn/lstm.py,
gru.py) on the CPU."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from tensorized_rnn.gru import GRU      # noqa: E402
from tensorized_rnn.lstm import LSTM    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [
    # name, cell, I, H, L, bias, B, T, with_init, want_dx   (engine limits: hidden_size % 128 == 0; the gradient wrt the
    # input needs input_size % 128 == 0)
    ("dense_lstm_i40_h128_L2", "lstm", 40, 128, 2, True, 5, 9, False, False),
    ("dense_lstm_i128_h128_L1_init_dx", "lstm", 128, 128, 1, True, 4, 7, True, True),
    ("dense_lstm_nobias_i28_L2", "lstm", 28, 128, 2, False, 3, 6, True, False),
    ("dense_gru_i40_h128_L1", "gru", 40, 128, 1, True, 5, 9, False, False),
    ("dense_gru_i128_h128_L2_init_dx", "gru", 128, 128, 2, True, 4, 8, True, True),
    ("dense_gru_nobias", "gru", 1, 128, 1, False, 6, 20, False, False),
]


def main():
    cpu = torch.device("cpu")
    for name, cell, I, H, L, bias, B, T, with_init, want_dx in CASES:
        torch.manual_seed(abs(hash(name)) % 1000 + 1 if False else len(name) * 7 + I + H)
        m = (LSTM if cell == "lstm" else GRU)(I, H, L, cpu, bias=bias)
        g = torch.Generator().manual_seed(3 + B)
        x = torch.rand(B, T, I, generator=g).requires_grad_(want_dx)
        init = None
        extra = {}
        if with_init:
            h0 = (0.3 * torch.randn(B, H, generator=g)).requires_grad_(True)
            if cell == "lstm":
                c0 = (0.3 * torch.randn(B, H, generator=g)).requires_grad_(True)
                init = (h0, c0)
            else:
                init = h0
        w_out, w_h = torch.randn(B, T, H, generator=g), torch.randn(B, H, generator=g)
        if cell == "lstm":
            out, (h, c) = m(x, init)
            loss = (out * w_out).sum() + (h * w_h).sum() + (c * w_h).sum() * 0.5
            extra["cT"] = c.detach().numpy()
        else:
            out, h = m(x, init)
            loss = (out * w_out).sum() + (h * w_h).sum()
        loss.backward()
        d = {"x": x.detach().numpy(), "out": out.detach().numpy(), "hT": h.detach().numpy(), "w_out": w_out.numpy(),
             "w_h": w_h.numpy(), "meta": np.array([I, H, L, int(bias), B, T, int(with_init), int(want_dx)]),
             "cell": np.array(cell)}
        d.update(extra)
        for k, v in m.state_dict().items():
            d["param:" + k] = v.detach().numpy()
        for k, p in m.named_parameters():
            d["grad:" + k] = p.grad.numpy()
        if want_dx:
            d["dx"] = x.grad.numpy()
        if with_init:
            d["h0"] = init[0].detach().numpy() if cell == "lstm" else init.detach().numpy()
            d["dh0"] = (init[0] if cell == "lstm" else init).grad.numpy()
            if cell == "lstm":
                d["c0"] = init[1].detach().numpy()
                d["dc0"] = init[1].grad.numpy()
        np.savez(os.path.join(HERE, name + ".npz"), **d)
        print(name, "loss %.5f" % loss.item())


if __name__ == "__main__":
    main()
