"""Generate GE2E-head fixtures from the reference (run in the build container, where /root/reference exists):

    python tests/golden/make_golden_ge2e.py

Writes tests/golden/ge2e_*.npz: raw embeddings, normalised embeddings, scaled similarity matrix, loss and the gradients of
the loss wrt the raw embeddings and the similarity weight / bias, all computed by the reference's own SpeakerEncoder code
(experiments/speaker_verification/encoder/speaker_encoder.py) on the CPU.  `np.int` is restored first: the reference uses
the alias that numpy >= 1.24 removed."""
import os
import sys
import types

import numpy as np
import torch

np.int = int                                           # noqa: the reference's speaker_encoder.py:120,163
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "experiments", "speaker_verification"))
HERE = os.path.dirname(os.path.abspath(__file__))


def load_speaker_encoder():
    """Import the reference's SpeakerEncoder class without its training-script dependencies."""
    for name in ("utils", "utils.modelutils"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["utils.modelutils"].count_model_params = lambda m: sum(p.numel() for p in m.parameters())
    import importlib
    pkg = importlib.import_module("encoder.speaker_encoder")
    return pkg.SpeakerEncoder


def main():
    SpeakerEncoder = load_speaker_encoder()
    cpu = torch.device("cpu")
    for name, S, U, E, seed in [("ge2e_s6_u4_e32", 6, 4, 32, 1), ("ge2e_s8_u10_e256", 8, 10, 256, 2),
                                ("ge2e_s64_u10_e256", 64, 10, 256, 3), ("ge2e_s5_u3_e40", 5, 3, 40, 4)]:
        torch.manual_seed(seed)
        import io
        from contextlib import redirect_stdout
        with redirect_stdout(io.StringIO()):
            enc = SpeakerEncoder(40, 64, 1, E, cpu, cpu, compression="tt", n_cores=3, rank=4)
        w = torch.nn.Parameter(torch.tensor([10.0 + 0.5 * seed]))
        b = torch.nn.Parameter(torch.tensor([-5.0 + 0.25 * seed]))
        enc.similarity_weight, enc.similarity_bias = w, b
        raw = torch.randn(S * U, E, requires_grad=True)
        # forward tail of SpeakerEncoder.forward (speaker_encoder.py:86-89)
        embeds_raw = enc.relu(raw)
        embeds = embeds_raw / torch.norm(embeds_raw, dim=1, keepdim=True)
        ev = embeds.view(S, U, E)
        sim = enc.similarity_matrix(ev)
        loss, eer = enc.loss(ev)
        loss.backward()
        np.savez(os.path.join(HERE, name + ".npz"), raw=raw.detach().numpy(), embeds=embeds.detach().numpy(),
                 sim=sim.detach().numpy(), loss=np.float32(loss.item()), eer=np.float32(eer), w=w.detach().numpy(),
                 b=b.detach().numpy(), d_raw=raw.grad.numpy(), d_w=w.grad.numpy(), d_b=b.grad.numpy(),
                 shape=np.array([S, U, E]))
        print(name, "loss %.6f eer %.4f" % (loss.item(), eer))


if __name__ == "__main__":
    main()
