"""CPU oracle of the GE2E head.  TEST INFRASTRUCTURE ONLY (same rules as ttrnn_oracle.py: imported by tests/,
smoke() and bench.py's CPU legs, never by the product package).

Restates, in plain torch-on-CPU ops, the training path of the reference's SpeakerEncoder
(experiments/speaker_verification/encoder/speaker_encoder.py).  Pinned by tests/golden/ge2e_*.npz, generated from the
reference itself by tests/golden/make_golden_ge2e.py (which sets `np.int = int` before importing it: the reference
uses the alias numpy removed).
"""
import torch


def embed_normalize(x):
    """speaker_encoder.py:86-89: embeds_raw = relu(linear(h)); embeds = embeds_raw / ||embeds_raw||_2 per row."""
    r = torch.relu(x)
    return r / torch.norm(r, dim=1, keepdim=True)


def similarity_matrix(embeds, w, b):
    """speaker_encoder.py:93-141 with enrollment_embeds = None.  embeds (S, U, E) -> (S, U, S)."""
    S, U = embeds.shape[:2]
    # :109-110 inclusive centroids, L2-normalised
    c_incl = torch.mean(embeds, dim=1, keepdim=True)
    c_incl = c_incl / torch.norm(c_incl, dim=2, keepdim=True)
    # :114-116 exclusive centroids (one per utterance), L2-normalised
    c_excl = (torch.sum(embeds, dim=1, keepdim=True) - embeds) / (U - 1)
    c_excl = c_excl / torch.norm(c_excl, dim=2, keepdim=True)
    # :121-125 the Python loop over speakers: column j = dot with c_incl[j] for other speakers' utterances,
    # dot with the utterance's own exclusive centroid on the block diagonal
    sim = torch.zeros(S, U, S, dtype=embeds.dtype)
    cols = []
    for j in range(S):
        col = (embeds * c_incl[j]).sum(dim=2)                       # (S, U)
        own = (embeds[j] * c_excl[j]).sum(dim=1)                    # (U,)
        col = torch.cat([col[:j], own[None], col[j + 1:]], dim=0)
        cols.append(col)
    sim = torch.stack(cols, dim=2)
    return sim * w + b                                               # :139


def loss(embeds, w, b):
    """speaker_encoder.py:143-156: mean cross-entropy of the (S*U, S) scaled similarities against the speaker index."""
    S, U = embeds.shape[:2]
    sm = similarity_matrix(embeds, w, b).reshape(S * U, S)
    target = torch.arange(S).repeat_interleave(U)
    return torch.nn.functional.cross_entropy(sm, target)
