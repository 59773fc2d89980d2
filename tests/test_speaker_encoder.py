"""The GE2E training step end to end (SURVEY.md 8f-4 + 8f-3): our `SpeakerEncoder` (TT or dense recurrent stack, projection,
ReLU + L2 norm, GE2E loss -- all on the device) against fixtures generated from the reference's own SpeakerEncoder
(tests/golden/make_golden_speaker_encoder.py): embeddings, loss, EER and the gradient of every parameter."""
import glob
import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from helpers import FWD_TOL, GRAD_TOL, GOLDEN, rel_err

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "spkenc_*.npz")))
KW = {"spkenc_tt_lstm_h64_L2": dict(compression="tt", n_cores=3, rank=4),
      "spkenc_dense_lstm_h128_L1": dict(compression=None),
      "spkenc_tt_gru_h64_L1": dict(compression="tt", n_cores=2, rank=2, use_gru=True)}


def build(name, g):
    import tensorized_rnn_b200 as tr
    S, U, T, H, L = (int(v) for v in g["shape"])
    with redirect_stdout(io.StringIO()):
        enc = tr.SpeakerEncoder(40, H, L, 32, torch.device("cpu"), None, **KW[name])
    sd = {k[len("param:"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param:")}
    return enc, sd, (S, U, T, H, L)


@pytest.mark.parametrize("name", CASES)
def test_state_dict_matches_reference(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    enc, sd, _ = build(name, g)
    own = enc.state_dict()
    assert sorted(own.keys()) == sorted(sd.keys())
    for k in sd:
        assert tuple(own[k].shape) == tuple(sd[k].shape), k
    enc.load_state_dict(sd)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_training_step_matches_reference(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    enc, sd, (S, U, T, H, L) = build(name, g)
    enc.load_state_dict(sd)
    enc = enc.to("cuda:0")
    x = torch.from_numpy(g["x"]).to("cuda:0")
    embeds = enc(x)
    assert rel_err(embeds, g["embeds"]) <= FWD_TOL
    loss, eer = enc.loss(embeds.view(S, U, -1), compute_eer=True)
    assert abs(float(loss.detach()) - float(g["loss"])) <= FWD_TOL * abs(float(g["loss"]))
    assert abs(float(eer) - float(g["eer"])) <= 1e-3
    loss.backward()
    torch.cuda.synchronize()
    bad = {}
    for k, p in enc.named_parameters():
        ref = g.get("grad:" + k)
        if ref is None:
            continue
        if k == "similarity_bias":
            assert abs(float(p.grad)) < 1e-6            # exactly zero mathematically (softmax shift invariance)
            continue
        e = rel_err(p.grad, ref)
        if not e <= GRAD_TOL:
            bad[k] = e
    assert not bad, bad
    enc.do_gradient_ops()                               # gradient scale + clipping run on the device tensors
