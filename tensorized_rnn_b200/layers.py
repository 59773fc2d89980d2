"""TTLinear: y = x W^T + b with W (out x in) held as a tensor train.

Mirror of reference t3nsor/layers.py:83-127 (constructor signature, attributes `shape`,
`weight_t`, `bias`, the `parameters.{k}` state_dict keys); forward/backward run the batched
TT-matvec CUDA kernels instead of d einsums (t3nsor/ops.py:54-93).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .functional import ttlinear
from .initializers import glorot_initializer
from .shapes import auto_shape
from .tensor_train import transpose


class TTLinear(nn.Module):
    def __init__(self, in_features=None, out_features=None, bias=True, init=None, shape=None,
                 auto_shapes=True, d=3, tt_rank=8, auto_shape_mode='ascending',
                 auto_shape_criterion='entropy'):
        super(TTLinear, self).__init__()
        if auto_shapes:
            if in_features is None or out_features is None:
                raise ValueError("Shape is not specified")
            shape = [auto_shape(in_features, d=d, criterion=auto_shape_criterion, mode=auto_shape_mode),
                     auto_shape(out_features, d=d, criterion=auto_shape_criterion, mode=auto_shape_mode)]
        if init is None:
            if shape is None:
                raise ValueError("if init is not provided, please specify shape, or set auto_shapes=True")
            init = glorot_initializer(shape, tt_rank=tt_rank)
        else:
            shape = init.raw_shape
        self.shape = [[int(v) for v in shape[0]], [int(v) for v in shape[1]]]
        self.weight_t = transpose(init).to_parameter()
        # The reference assigns the ParameterList to an attribute called `parameters`
        # (layers.py:114), which registers a sub-module of that name: state_dict keys are
        # `parameters.{k}` while `.parameters()` stays the nn.Module method.  Same here.
        setattr(self, "parameters", self.weight_t.parameter)
        n_out = 1
        for v in self.shape[1]:
            n_out *= v
        if out_features is None:
            out_features = n_out
        if bias:
            self.bias = torch.nn.Parameter(1e-3 * torch.ones(out_features))
        else:
            self.register_parameter('bias', None)
        print('Created TTLinear layer with input shape: {}. output shape: {}'.format(self.shape[0], self.shape[1]))

    # -- geometry used by the kernels ------------------------------------------------------
    @property
    def in_features_total(self) -> int:
        return self.weight_t.shape[1]

    @property
    def out_features_total(self) -> int:
        return self.weight_t.shape[0]

    def tt_modes(self):
        """(in_modes, out_modes, ranks) of the stored cores (r_k, i_k, j_k, r_{k+1})."""
        cores = self.weight_t.tt_cores
        return ([int(c.shape[2]) for c in cores], [int(c.shape[1]) for c in cores],
                [int(c.shape[0]) for c in cores] + [1])

    def forward(self, x):
        modes = self.tt_modes()
        shape = _lib.make_tt_shape(*modes)
        lead = x.shape[:-1]
        y = ttlinear(shape, self.in_features_total, self.out_features_total, x.reshape(-1, x.shape[-1]),
                     self.bias, list(self.weight_t.tt_cores))
        return y.reshape(*lead, self.out_features_total)


class TTLinearSet(nn.Module):
    """n_gates independent TTLinear maps whose outputs are concatenated column-wise: the "naive TT"
    weight of reference tensorized_rnn/tt_linearset.py:5-38 (constructor signature, sub-module names
    `gate{i}` / `gates.{i}` and so the state_dict keys are the reference's).  Each gate runs the
    batched TT-matvec CUDA kernels; the concatenation is a plain copy."""

    def __init__(self, in_features=None, out_features=None, n_gates=4, bias=True, init=None, shape=None,
                 auto_shapes=True, d=3, tt_rank=8, auto_shape_mode='ascending',
                 auto_shape_criterion='entropy'):
        super(TTLinearSet, self).__init__()
        self.n_gates = n_gates
        self.in_features = in_features
        self.out_features = out_features
        gates = []
        for i in range(n_gates):
            cur_gate = TTLinear(in_features=in_features, out_features=out_features,
                                bias=bias, auto_shapes=auto_shapes, d=d, tt_rank=tt_rank,
                                init=init, shape=shape, auto_shape_mode=auto_shape_mode,
                                auto_shape_criterion=auto_shape_criterion)
            setattr(self, 'gate{}'.format(i), cur_gate)
            gates.append(cur_gate)
        self.gates = nn.ModuleList(gates)

    def forward(self, x):
        batch_size, in_size = x.size()
        assert in_size == self.in_features
        return torch.cat([gate(x) for gate in self.gates], dim=1)
