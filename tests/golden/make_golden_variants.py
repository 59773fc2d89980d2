"""Golden fixtures FROM THE REFERENCE for the module variants that run in cell-step mode
(SURVEY.md section 8f-2/3): `is_naive=True` (TTLinearSet, tensorized_rnn/tt_linearset.py),
`new_core='first'/'last'` (tensorized_rnn/rnn_utils.py:29-34) and `log_grads=True`
(ActivGradLogger, tensorized_rnn/rnn_utils.py:42-215).

    python tests/golden/make_golden_variants.py        (build container only: needs /root/reference)

Same layout as make_golden.py; `log:*` arrays hold what `ActivGradLogger.get_logs()` returns after one
minibatch + `end_minibatch()` + `end_epoch()`.
"""
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
from tensorized_rnn.rnn_utils import ActivGradLogger  # noqa: E402

# name, cell, I, H, L, d, r, bias, B, T, init_states, is_naive, new_core, log_grads
CASES = [
    ("naive_lstm_d2r3_L2",       "lstm", 12, 24,  2, 2, 3, True,  4, 5, True,  True,  None,    False),
    ("naive_gru_d3r2_L2",        "gru",  28, 64,  2, 3, 2, True,  3, 4, False, True,  None,    False),
    ("naive_gru_nobias",         "gru",  1,  36,  1, 2, 4, False, 3, 6, True,  True,  None,    False),
    ("newcore_first_lstm",       "lstm", 40, 64,  2, 2, 4, True,  3, 4, False, False, "first", False),
    ("newcore_last_lstm",        "lstm", 40, 64,  1, 3, 2, True,  3, 4, True,  False, "last",  False),
    ("newcore_last_gru",         "gru",  28, 48,  2, 2, 3, True,  2, 5, False, False, "last",  False),
    ("newcore_first_gru_nobias", "gru",  12, 24,  1, 2, 2, False, 4, 3, True,  False, "first", False),
    ("loggrads_lstm_d2r4_L2",    "lstm", 28, 64,  2, 2, 4, True,  3, 6, False, False, None,    True),
    ("loggrads_gru_d2r4",        "gru",  1,  256, 1, 2, 4, True,  2, 7, False, False, None,    True),
    ("loggrads_naive_lstm",      "lstm", 12, 24,  1, 2, 2, True,  3, 5, True,  True,  None,    True),
]


def main():
    index = []
    for ci, (name, cell, I, H, L, d, r, bias, B, T, with_init, naive, new_core, log_grads) in enumerate(CASES):
        seed = 5000 + ci
        ActivGradLogger.all_loggers.clear()
        torch.manual_seed(seed)
        cls = TTLSTM if cell == "lstm" else TTGRU
        with redirect_stdout(io.StringIO()):
            model = cls(I, H, L, torch.device("cpu"), n_cores=d, tt_rank=r, bias=bias, is_naive=naive,
                        log_grads=log_grads, new_core=new_core)
        g = torch.Generator().manual_seed(seed + 7)
        x = torch.rand(B, T, I, generator=g).requires_grad_(True)
        init = None
        if with_init:
            h0 = (0.3 * torch.randn(B, H, generator=g)).requires_grad_(True)
            init = (h0, (0.3 * torch.randn(B, H, generator=g)).requires_grad_(True)) if cell == "lstm" else h0
        w_out = torch.randn(B, T, H, generator=g)
        w_h = torch.randn(B, H, generator=g)
        w_c = torch.randn(B, H, generator=g)
        if cell == "lstm":
            out, (h, c) = model(x, init)
            loss = (out * w_out).sum() + (h * w_h).sum() + (c * w_c).sum()
        else:
            out, h = model(x, init)
            c = None
            loss = (out * w_out).sum() + (h * w_h).sum()
        loss.backward()
        blob = {"x": x.detach().numpy(), "w_out": w_out.numpy(), "w_h": w_h.numpy(), "w_c": w_c.numpy(),
                "f32:out": out.detach().numpy(), "f32:hT": h.detach().numpy(), "f32:dx": x.grad.numpy()}
        if c is not None:
            blob["f32:cT"] = c.detach().numpy()
        if init is not None:
            if cell == "lstm":
                blob["h0"], blob["c0"] = init[0].detach().numpy(), init[1].detach().numpy()
                blob["f32:dh0"], blob["f32:dc0"] = init[0].grad.numpy(), init[1].grad.numpy()
            else:
                blob["h0"] = init.detach().numpy()
                blob["f32:dh0"] = init.grad.numpy()
        for k, v in model.state_dict().items():
            blob["param:" + k] = v.detach().contiguous().numpy()
        for k, p in model.named_parameters():
            blob["f32:grad:" + k] = p.grad.detach().contiguous().numpy()
        if log_grads:
            ActivGradLogger.end_minibatch()
            ActivGradLogger.end_epoch()
            for (var, qnt), mat in ActivGradLogger.get_logs().items():
                blob["log:%s:%s" % (var, qnt)] = mat.numpy()
        path = os.path.join(HERE, "variant_" + name + ".npz")
        np.savez_compressed(path, **blob)
        index.append({"name": name, "cell": cell, "input_size": I, "hidden_size": H, "num_layers": L, "n_cores": d,
                      "tt_rank": r, "bias": bias, "batch": B, "seq_len": T, "init_states": with_init,
                      "is_naive": naive, "new_core": new_core, "log_grads": log_grads, "seed": seed,
                      "state_dict_keys": list(model.state_dict().keys())})
        print(name, "%.1f KiB" % (os.path.getsize(path) / 1024))
    ActivGradLogger.all_loggers.clear()
    with open(os.path.join(HERE, "variants_index.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
