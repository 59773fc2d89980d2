"""TT-matrix container: the parameter layout contract of the reference.

Mirrors the public surface of reference t3nsor/tensor_train.py:6-163 that the
recurrent path touches (`tt_cores`, `raw_shape`, `shape`, `ranks`, `ndims`, `dof`,
`total`, `to_parameter()`, `parameter`, `full()`).  Cores are held in the layout the
reference's `weight_t` parameters have after `t3.transpose` (t3nsor/ops.py:47-51):
(r_k, i_k, j_k, r_{k+1}), i = output mode, j = input mode -- but CONTIGUOUS here
(the reference keeps transposed views; shapes and state_dict keys are identical,
strides are not part of any contract).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.nn as nn


class TensorTrain(object):
    def __init__(self, tt_cores: Sequence[torch.Tensor], convert_to_tensors: bool = True):
        cores = list(tt_cores)
        if convert_to_tensors:
            cores = [c if isinstance(c, torch.Tensor) else torch.as_tensor(c, dtype=torch.float32) for c in cores]
        if any(c.dim() != 4 for c in cores):
            raise ValueError("TensorTrain here holds TT-matrices only: every core must be 4-d (r, m1, m2, r')")
        self._tt_cores = cores
        self._is_parameter = False
        self._parameter = None

    # -- shape bookkeeping (tensor_train.py:14-35) ---------------------------------------
    @property
    def tt_cores(self) -> List[torch.Tensor]:
        return self._tt_cores

    @property
    def is_tt_matrix(self) -> bool:
        return True

    @property
    def raw_shape(self):
        return [[int(c.shape[1]) for c in self._tt_cores], [int(c.shape[2]) for c in self._tt_cores]]

    @property
    def shape(self):
        rs = self.raw_shape
        return [int(np.prod(rs[0])), int(np.prod(rs[1]))]

    @property
    def ndims(self) -> int:
        return len(self._tt_cores)

    @property
    def ranks(self):
        return [int(c.shape[0]) for c in self._tt_cores] + [1]

    @property
    def dof(self) -> int:
        return int(sum(c.numel() for c in self._tt_cores))

    @property
    def total(self) -> int:
        s = self.shape
        return int(s[0] * s[1])

    @property
    def is_parameter(self) -> bool:
        return self._is_parameter

    @property
    def parameter(self) -> nn.ParameterList:
        if not self._is_parameter:
            raise ValueError("Not a parameter, run .to_parameter() first")
        return self._parameter

    # -- conversions ------------------------------------------------------------------------
    def to(self, device):
        return TensorTrain([c.to(device) for c in self._tt_cores], convert_to_tensors=False)

    def detach(self):
        return TensorTrain([c.detach() for c in self._tt_cores], convert_to_tensors=False)

    def to_parameter(self) -> "TensorTrain":
        """Wrap every core in an nn.Parameter tagged `is_tt` and collect them in a ParameterList
        (tensor_train.py:104-114).  The list entries and `tt_cores` are the same objects."""
        cores = []
        for c in self._tt_cores:
            p = nn.Parameter(c.detach().clone().contiguous())
            p.is_tt = True
            cores.append(p)
        out = TensorTrain(cores, convert_to_tensors=False)
        out._parameter = nn.ParameterList(cores)
        out._is_parameter = True
        return out

    def full(self) -> torch.Tensor:
        """Dense (shape[0] x shape[1]) matrix (tensor_train.py:116-146); works on any strides."""
        cores = self._tt_cores
        res = cores[0].reshape(-1, cores[0].shape[-1])
        for c in cores[1:]:
            res = (res @ c.reshape(c.shape[0], -1)).reshape(-1, c.shape[-1])
        d = len(cores)
        inter = []
        for c in cores:
            inter += [int(c.shape[1]), int(c.shape[2])]
        res = res.reshape(inter)
        perm = list(range(0, 2 * d, 2)) + list(range(1, 2 * d, 2))
        return res.permute(perm).reshape(self.shape)

    def __str__(self) -> str:
        rs = self.raw_shape
        return ("A TT-Matrix of size %d x %d, underlying tensor shape: %s x %s, TT-ranks: %s on device '%s' "
                "with compression rate %.2f" % (self.shape[0], self.shape[1], rs[0], rs[1], self.ranks,
                                                self._tt_cores[0].device, self.total / self.dof))


def transpose(tt_matrix: TensorTrain) -> TensorTrain:
    """Swap the two mode dimensions of every core (t3nsor/ops.py:47-51); returns contiguous cores."""
    return TensorTrain([c.transpose(1, 2).contiguous() for c in tt_matrix.tt_cores], convert_to_tensors=False)
