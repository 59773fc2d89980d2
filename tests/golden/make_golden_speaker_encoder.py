"""End-to-end fixtures of the reference's GE2E training step (run in the build container, /root/reference present):

    python tests/golden/make_golden_speaker_encoder.py

For a TT and a dense `SpeakerEncoder` (experiments/speaker_verification/encoder/speaker_encoder.py) on the CPU: state_dict,
utterances (S*U, T, mel), embeddings, loss, EER and the gradient of the loss wrt every parameter.  `np.int` is restored first
(SURVEY.md 8c)."""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

np.int = int                                           # noqa
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_ge2e import load_speaker_encoder      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    SpeakerEncoder = load_speaker_encoder()
    cpu = torch.device("cpu")
    for name, kw, S, U, T in [("spkenc_tt_lstm_h64_L2", dict(compression="tt", n_cores=3, rank=4), 4, 3, 10),
                              ("spkenc_dense_lstm_h128_L1", dict(compression=None), 3, 4, 8),
                              ("spkenc_tt_gru_h64_L1", dict(compression="tt", n_cores=2, rank=2, use_gru=True), 5, 2, 9)]:
        H = 128 if kw["compression"] is None else 64
        L = 1 if "L1" in name else 2
        torch.manual_seed(17)
        with redirect_stdout(io.StringIO()):
            enc = SpeakerEncoder(40, H, L, 32, cpu, cpu, **kw)
        x = torch.rand(S * U, T, 40, generator=torch.Generator().manual_seed(5))
        embeds = enc(x)
        loss, eer = enc.loss(embeds.view(S, U, -1))
        enc.zero_grad()
        loss.backward()
        d = {"x": x.numpy(), "embeds": embeds.detach().numpy(), "loss": np.float32(loss.item()), "eer": np.float32(eer),
             "shape": np.array([S, U, T, H, L]), "d_similarity_weight": enc.similarity_weight.grad.numpy(),
             "d_similarity_bias": enc.similarity_bias.grad.numpy(), "similarity_weight": enc.similarity_weight.detach().numpy(),
             "similarity_bias": enc.similarity_bias.detach().numpy()}
        for k, v in enc.state_dict().items():
            d["param:" + k] = v.detach().numpy()
        for k, p in enc.named_parameters():
            if p.grad is not None:
                d["grad:" + k] = p.grad.numpy()
        np.savez(os.path.join(HERE, name + ".npz"), **d)
        print(name, "loss %.6f eer %.4f" % (loss.item(), eer), sorted(k for k in d if k.startswith("param:"))[:3])


if __name__ == "__main__":
    main()
