"""Core initialisers of the TT layers.

Distribution contract (reference t3nsor/initializers.py:166-300): every entry of core k is
N(0, s^2) with s = (sqrt(2 / (n_in + n_out)))^(1/d) * prod_k r_k^(-1/(2d)), inner ranks all equal
to `tt_rank`.  Cores are drawn with torch.randn in the order k = 0..d-1 with shape
(r_k, shape[0][k], shape[1][k], r_{k+1}), so a module built here under `torch.manual_seed(s)` holds
bit-identical parameters to the reference built under the same seed (tested against the fixtures).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from .tensor_train import TensorTrain


def _ranks(tt_rank, d: int) -> np.ndarray:
    r = np.asarray(tt_rank)
    if r.size == 1:
        r = np.concatenate([[1], int(r) * np.ones(d - 1, dtype=np.int64), [1]])
    r = r.astype(np.int64)
    if r.size != d + 1 or r[0] != 1 or r[-1] != 1:
        raise ValueError("tt_rank must be a scalar or a (d+1)-vector with unit boundary ranks")
    return r


def matrix_with_random_cores(shape: Sequence[Sequence[int]], tt_rank=2, mean: float = 0., stddev: float = 1.,
                             dtype=torch.float32) -> TensorTrain:
    s0, s1 = [int(v) for v in shape[0]], [int(v) for v in shape[1]]
    if len(s0) != len(s1):
        raise ValueError("shape[0] and shape[1] must have the same number of modes")
    d = len(s0)
    r = _ranks(tt_rank, d)
    cores = [torch.randn((int(r[k]), s0[k], s1[k], int(r[k + 1])), dtype=dtype) * stddev + mean for k in range(d)]
    return TensorTrain(cores, convert_to_tensors=False)


def random_matrix(shape, tt_rank=2, mean: float = 0., stddev: float = 1., dtype=torch.float32) -> TensorTrain:
    if abs(mean) >= 1e-8:
        raise NotImplementedError("non-zero mean is not supported yet")
    d = len(shape[0])
    r = _ranks(tt_rank, d).astype(np.float64)
    core_std = stddev ** (1.0 / d) * float(np.prod(r ** (-1.0 / (2 * d))))
    return matrix_with_random_cores(shape, tt_rank=tt_rank, stddev=core_std, dtype=dtype)


def glorot_initializer(shape, tt_rank=2, dtype=torch.float32) -> TensorTrain:
    n_in = float(np.prod(shape[0]))
    n_out = float(np.prod(shape[1]))
    return random_matrix(shape, tt_rank=tt_rank, stddev=float(np.sqrt(2.0 / (n_in + n_out))), dtype=dtype)
