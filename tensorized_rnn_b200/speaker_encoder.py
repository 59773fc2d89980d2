"""`SpeakerEncoder`: the GE2E speaker encoder of the reference with the recurrent stack, the projection and the GE2E head all
on the device.

Mirror of experiments/speaker_verification/encoder/speaker_encoder.py:17-170 -- same constructor signature, attribute and
sub-module names (`rnn`, `linear`, `similarity_weight`, `similarity_bias`: reference checkpoints load with `load_state_dict`),
`forward(utterances)` -> L2-normalised embeddings, `similarity_matrix`, `loss`, `do_gradient_ops`.  Differences: `loss_device` is
accepted and ignored (the reference moves the embeddings to the CPU for the loss, encoder/main.py:279-280; here the loss runs
in the CUDA library, ge2e.py); the EER, which the reference computes with scikit-learn on the CPU and does not backpropagate
(speaker_encoder.py:158-169), is computed the same way from a host copy of the similarity matrix only when `compute_eer=True`;
the evaluation form of `similarity_matrix` (`enrollment_embeds` given) is not provided by the CUDA head and raises.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .dense import GRU, LSTM
from .ge2e import embed_normalize, ge2e_loss
from .layers import TTLinear
from .rnn import TTGRU, TTLSTM


class SpeakerEncoder(nn.Module):
    def __init__(self, mel_n_channels, model_hidden_size, model_num_layers, model_embedding_size, device, loss_device=None,
                 compression=None, n_cores=3, rank=8, clip=3, use_gru=False):
        super().__init__()
        self.loss_device = loss_device
        self.clip = clip
        self.use_gru = use_gru
        if compression is None:
            cls = GRU if use_gru else LSTM
            self.rnn = cls(mel_n_channels, model_hidden_size, model_num_layers, device)
            self.linear = nn.Linear(in_features=model_hidden_size, out_features=model_embedding_size).to(device)
        elif compression == 'tt':
            print("Encoding linear layer as a tensor-train...")
            if not use_gru:
                self.rnn = TTLSTM(mel_n_channels, model_hidden_size, model_num_layers, device, n_cores=n_cores, tt_rank=rank)
            else:
                self.rnn = TTGRU(mel_n_channels, model_hidden_size, model_num_layers, device, bias=True, n_cores=n_cores,
                                 tt_rank=rank)
            self.linear = TTLinear(in_features=model_hidden_size, out_features=model_embedding_size, bias=True,
                                   auto_shapes=True, d=n_cores, tt_rank=rank).to(device)
        else:
            raise ValueError("Unknown compression type: '{}'".format(compression))
        # cosine similarity scaling, fixed initial values (speaker_encoder.py:55-57)
        self.similarity_weight = nn.Parameter(torch.tensor([10.], device=device))
        self.similarity_bias = nn.Parameter(torch.tensor([-5.], device=device))

    def do_gradient_ops(self):
        self.similarity_weight.grad *= 0.01
        self.similarity_bias.grad *= 0.01
        torch.nn.utils.clip_grad_norm_(self.parameters(), self.clip, norm_type=2)

    def forward(self, utterances, hidden_init=None):
        if not self.use_gru:
            out, (last_hidden, _) = self.rnn(utterances)
        else:
            out, last_hidden = self.rnn(utterances)
        return embed_normalize(self.linear(last_hidden))            # relu + L2 norm (speaker_encoder.py:86-89)

    def similarity_matrix(self, verification_embeds, enrollment_embeds=None):
        if enrollment_embeds is not None:
            raise NotImplementedError("the device head implements the training form (enrollment = verification)")
        return ge2e_loss(verification_embeds, self.similarity_weight, self.similarity_bias, return_similarity=True)[1]

    def loss(self, verification_embeds, enrollment_embeds=None, compute_eer=False, group=None):
        """Returns (loss, eer): eer is None unless compute_eer (it needs a device-to-host copy and scikit-learn)."""
        if enrollment_embeds is not None:
            raise NotImplementedError("the device head implements the training form (enrollment = verification)")
        if not compute_eer:
            return ge2e_loss(verification_embeds, self.similarity_weight, self.similarity_bias, group=group), None
        loss, sim = ge2e_loss(verification_embeds, self.similarity_weight, self.similarity_bias, group=group,
                              return_similarity=True)
        from scipy.interpolate import interp1d
        from scipy.optimize import brentq
        from sklearn.metrics import roc_curve
        S, U = sim.shape[0], sim.shape[1]
        labels = np.eye(S, dtype=np.int64)[np.repeat(np.arange(S), U)]
        preds = sim.detach().reshape(S * U, S).cpu().numpy()
        fpr, tpr, _ = roc_curve(labels.flatten(), preds.flatten())
        eer = brentq(lambda x: 1. - x - interp1d(fpr, tpr)(x), 0., 1.)
        return loss, eer
