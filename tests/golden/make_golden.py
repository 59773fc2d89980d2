"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports onucharles/tensorized-rnn read-only from /root/reference, builds the
reference TTLSTM / TTGRU modules under fixed seeds, runs forward + backward on
CPU in FP32 (and FP64 as a tie-breaker) and stores parameters, inputs, outputs,
final states and every parameter / input gradient as .npz files, plus a table of
`auto_shape` / `tt_shape` results.  The fixtures pin both `oracle/` and the CUDA
path; nothing in the test-suite reads /root/reference at run time.
"""
import copy
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

sys.path.insert(0, REF)
from tensorized_rnn.tt_lstm import TTLSTM          # noqa: E402
from tensorized_rnn.gru import TTGRU                # noqa: E402
from tensorized_rnn.rnn_utils import tt_shape       # noqa: E402
from t3nsor.utils import auto_shape                 # noqa: E402
from t3nsor.layers import TTLinear                  # noqa: E402

# name, cell, I, H, L, d, r, bias, B, T, init_states?, store_full_outputs?
CASES = [
    ("cfg1_lstm_d2r4",      "lstm", 1,   256,  1, 2, 4,  True,  3, 7,   False, True),
    ("cfg2_gru_d2r4",       "gru",  1,   256,  1, 2, 4,  True,  3, 7,   False, True),
    ("cfg3_lstm_d3r8_L3",   "lstm", 40,  256,  3, 3, 8,  True,  2, 5,   False, True),
    ("cfg3alt_lstm_d2r2_L3", "lstm", 40, 256,  3, 2, 2,  True,  2, 5,   True,  True),
    ("cfg4_lstm_d4r16_L3",  "lstm", 40,  256,  3, 4, 16, True,  2, 3,   False, True),
    ("cfg5_lstm_d4r8_H1024", "lstm", 256, 1024, 1, 4, 8,  True,  2, 3,   False, True),
    ("gru_d3r3_L2_nobias",  "gru",  28,  64,   2, 3, 3,  False, 3, 6,   True,  True),
    ("lstm_d2r2_L2_small",  "lstm", 12,  24,   2, 2, 2,  True,  5, 4,   True,  True),
    ("gru_d3r8_L3_ge2e",    "gru",  40,  256,  3, 3, 8,  True,  2, 5,   False, True),
    ("gru_d4r5_odd",        "gru",  30,  100,  1, 4, 5,  True,  3, 4,   True,  True),
    ("lstm_d3r4_H50_odd",   "lstm", 7,   50,   2, 3, 4,  True,  4, 3,   False, True),
    ("cfg1_lstm_T784",      "lstm", 1,   256,  1, 2, 4,  True,  2, 784, False, False),
    ("cfg2_gru_T784",       "gru",  1,   256,  1, 2, 4,  True,  2, 784, False, False),
]


def build(cell, I, H, L, d, r, bias, seed):
    torch.manual_seed(seed)
    with redirect_stdout(io.StringIO()):
        if cell == "lstm":
            return TTLSTM(I, H, L, torch.device("cpu"), n_cores=d, tt_rank=r, bias=bias)
        return TTGRU(I, H, L, torch.device("cpu"), n_cores=d, tt_rank=r, bias=bias)


def run(model, cell, x, init, w_out, w_h, w_c):
    model.zero_grad()
    x = x.clone().requires_grad_(True)
    if init is not None:
        init = tuple(t.clone().requires_grad_(True) for t in init) if cell == "lstm" \
            else init.clone().requires_grad_(True)
    if cell == "lstm":
        out, (h, c) = model(x, init)
        loss = (out * w_out).sum() + (h * w_h).sum() + (c * w_c).sum()
    else:
        out, h = model(x, init)
        c = None
        loss = (out * w_out).sum() + (h * w_h).sum()
    loss.backward()
    res = {"out": out.detach(), "hT": h.detach(), "dx": x.grad.detach()}
    if c is not None:
        res["cT"] = c.detach()
    if init is not None:
        if cell == "lstm":
            res["dh0"], res["dc0"] = init[0].grad.detach(), init[1].grad.detach()
        else:
            res["dh0"] = init.grad.detach()
    for name, p in model.named_parameters():
        res["grad:" + name] = p.grad.detach().clone()
    return res


def main():
    index = []
    for ci, (name, cell, I, H, L, d, r, bias, B, T, with_init, full) in enumerate(CASES):
        seed = 1000 + ci
        model = build(cell, I, H, L, d, r, bias, seed)
        g = torch.Generator().manual_seed(seed + 7)
        x = torch.rand(B, T, I, generator=g)
        if I == 1:   # "synthetic digits" normalisation, digit_classification/utils.py:9-10
            x = (x - 0.1307) / 0.3081
        init = None
        if with_init:
            h0 = 0.3 * torch.randn(B, H, generator=g)
            init = (h0, 0.3 * torch.randn(B, H, generator=g)) if cell == "lstm" else h0
        w_out = torch.randn(B, T, H, generator=g)
        if not full:          # long-run cases: gradient enters at the last step only (mnist_classifier.py:55)
            w_out[:, :-1, :] = 0
        w_h = torch.randn(B, H, generator=g)
        w_c = torch.randn(B, H, generator=g)

        r32 = run(model, cell, x, init, w_out, w_h, w_c)
        m64 = copy.deepcopy(model).double()
        init64 = None
        if cell == "lstm":
            init64 = tuple(t.double() for t in init) if init is not None else \
                (torch.zeros(B, H, dtype=torch.float64), torch.zeros(B, H, dtype=torch.float64))
        else:
            init64 = init.double() if init is not None else torch.zeros(B, H, dtype=torch.float64)
        r64 = run(m64, cell, x.double(), init64, w_out.double(), w_h.double(), w_c.double())

        blob = {"x": x.numpy(), "w_out_last": w_out[:, -1].numpy() if not full else np.zeros(0, np.float32),
                "w_out": w_out.numpy() if full else np.zeros(0, np.float32),
                "w_h": w_h.numpy(), "w_c": w_c.numpy()}
        if init is not None:
            blob["h0"] = (init[0] if cell == "lstm" else init).numpy()
            if cell == "lstm":
                blob["c0"] = init[1].numpy()
        for k, v in model.state_dict().items():
            blob["param:" + k] = v.contiguous().numpy()
        for tag, res in (("f32", r32), ("f64", r64)):
            for k, v in res.items():
                if k == "out" and not full:
                    v = v[:, -1]
                    k = "out_last"
                if k == "dx" and not full:
                    continue
                if tag == "f64" and (k.startswith("grad:") or k in ("dx", "out")) and not full:
                    pass
                blob[tag + ":" + k] = v.numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **blob)
        index.append({"name": name, "cell": cell, "input_size": I, "hidden_size": H, "num_layers": L,
                      "n_cores": d, "tt_rank": r, "bias": bias, "batch": B, "seq_len": T,
                      "init_states": with_init, "full_outputs": full, "seed": seed})
        print(name, "%.1f KiB" % (os.path.getsize(path) / 1024))
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1)

    # ---- stand-alone TTLinear cases (t3nsor/layers.py:83-127) -----------------
    lin = []
    for li, (fin, fout, d, r, bias, B) in enumerate([(256, 10, 2, 4, True, 5), (256, 256, 3, 8, True, 4),
                                                     (40, 1024, 4, 16, False, 3), (28, 1024, 2, 4, True, 6)]):
        torch.manual_seed(2000 + li)
        with redirect_stdout(io.StringIO()):
            m = TTLinear(in_features=fin, out_features=fout, bias=bias, auto_shapes=True, d=d, tt_rank=r)
        g = torch.Generator().manual_seed(3000 + li)
        x = torch.randn(B, fin, generator=g).requires_grad_(True)
        w = torch.randn(B, fout, generator=g)
        y = m(x)
        (y * w).sum().backward()
        blob = {"x": x.detach().numpy(), "w": w.numpy(), "y": y.detach().numpy(), "dx": x.grad.numpy()}
        for k, v in m.state_dict().items():
            blob["param:" + k] = v.contiguous().numpy()
        for k, p in m.named_parameters():
            blob["grad:" + k] = p.grad.numpy()
        np.savez_compressed(os.path.join(HERE, "ttlinear_%d.npz" % li), **blob)
        lin.append({"name": "ttlinear_%d" % li, "in_features": fin, "out_features": fout, "d": d,
                    "tt_rank": r, "bias": bias, "batch": B, "shape": [list(map(int, s)) for s in m.shape]})
    with open(os.path.join(HERE, "ttlinear_index.json"), "w") as f:
        json.dump(lin, f, indent=1)

    # ---- shape tables ---------------------------------------------------------
    table = {"auto_shape": {}, "tt_shape": []}
    ns = sorted(set(list(range(1, 401)) + [512, 640, 768, 784, 1000, 1024, 1536, 2048, 3072, 4096, 8192]))
    for d in (1, 2, 3, 4, 5):
        for n in ns:
            table["auto_shape"]["%d,%d" % (n, d)] = [int(v) for v in auto_shape(n, d=d)]
    for (fin, H, d, G) in [(1, 256, 2, 4), (256, 256, 2, 4), (1, 256, 2, 3), (256, 256, 2, 3), (40, 256, 3, 4),
                           (256, 256, 3, 4), (40, 256, 4, 4), (256, 256, 4, 4), (256, 1024, 4, 4),
                           (1024, 1024, 4, 4), (28, 256, 2, 4), (40, 256, 2, 4), (768, 768, 4, 4),
                           (256, 256, 3, 3), (256, 256, 4, 3), (40, 768, 3, 4), (80, 512, 3, 3)]:
        for nc in (None, "first", "last"):
            s = tt_shape(fin, H, d, G, new_core=nc)
            table["tt_shape"].append({"in": fin, "hidden": H, "n_cores": d, "n_gates": G, "new_core": nc,
                                      "shape": [[int(v) for v in s[0]], [int(v) for v in s[1]]]})
    with open(os.path.join(HERE, "shapes.json"), "w") as f:
        json.dump(table, f)
    print("shapes:", len(table["auto_shape"]), len(table["tt_shape"]))


if __name__ == "__main__":
    main()
