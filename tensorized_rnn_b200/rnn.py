"""TTLSTM / TTGRU: the nn.Module mirror of the reference's recurrent modules.

Constructor signatures, attribute names, sub-module names (`cell{i}.input_weights`,
`cell{i}.hidden_weights`), state_dict keys/shapes, `init_hidden`, `param_count` and the
forward contract (batch-first input, one shared initial state, outputs of the last layer +
final state of the last layer) follow reference tensorized_rnn/tt_lstm.py:6-62,
tensorized_rnn/lstm.py:44-135 and tensorized_rnn/gru.py:52-194.  The time loop, the TT
contractions and the gate math run in the CUDA library (one call per sequence).

Not in this round (raise NotImplementedError instead of silently diverging):
  * is_naive=True  (TTLinearSet, reference tt_linearset.py)   -- SURVEY.md section 8f-3
  * log_grads=True (per-step hooks, reference rnn_utils.py:42-215) -- SURVEY.md section 8f-2
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from .functional import RnnSpec, rnn_sequence
from .layers import TTLinear
from .shapes import tt_shape


def param_count(matrix: nn.Module) -> int:
    """Number of weights in a module (reference tensorized_rnn/rnn_utils.py:300-310)."""
    assert isinstance(matrix, torch.nn.Module)
    return int(sum(p.shape.numel() for p in matrix.parameters()))


class _TTCellBase(nn.Module):
    n_gate = 0
    cell_kind = ""

    def __init__(self, input_size, hidden_size, bias, device, n_cores, tt_rank, is_naive=False, new_core=None):
        super().__init__()
        if is_naive:
            raise NotImplementedError("is_naive=True (one TT matrix per gate) is not implemented in "
                                      "tensorized_rnn_b200 yet; use the concat-gates form (is_naive=False)")
        assert new_core in [None, 'first', 'last']
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.bias = bias
        self.device = device
        self.n_cores = n_cores
        self.tt_rank = tt_rank
        self.is_naive = is_naive
        self.new_core = new_core
        # creation order (ih before hh) fixes the RNG stream, as in reference lstm.py:14-15
        self.input_weights = self._create_input_hidden_weights()
        self.hidden_weights = self._create_hidden_hidden_weights()

    def _tt_linear(self, in_features):
        shape = tt_shape(in_features, self.hidden_size, self.n_cores, self.n_gate, new_core=self.new_core)
        return TTLinear(out_features=self.n_gate * self.hidden_size, shape=shape, bias=self.bias,
                        auto_shapes=False, d=self.n_cores, tt_rank=self.tt_rank).to(self.device)

    def _create_input_hidden_weights(self):
        return self._tt_linear(self.input_size)

    def _create_hidden_hidden_weights(self):
        return self._tt_linear(self.hidden_size)

    def flat_parameters(self) -> List[torch.Tensor]:
        """Parameters in C-ABI blob order: ih cores, ih bias, hh cores, hh bias."""
        out: List[torch.Tensor] = []
        for lin in (self.input_weights, self.hidden_weights):
            out += list(lin.weight_t.tt_cores)
            if lin.bias is not None:
                out.append(lin.bias)
        return out

    def _single_step_spec(self) -> RnnSpec:
        return RnnSpec(self.cell_kind, self.input_size, self.hidden_size, self.input_weights.bias is not None,
                       [self.input_weights.tt_modes()], [self.hidden_weights.tt_modes()])


class TTLSTMCell(_TTCellBase):
    """One TT-LSTM cell (reference tt_lstm.py:6-40; step semantics lstm.py:23-41)."""
    n_gate = 4
    cell_kind = "lstm"

    def forward(self, input, hx, cx):
        out, h, c = rnn_sequence(self._single_step_spec(), input.unsqueeze(1), hx, cx, self.flat_parameters())
        return h, c


class TTGRUCell(_TTCellBase):
    """One TT-GRU cell (reference gru.py:139-172; step semantics gru.py:25-50)."""
    n_gate = 3
    cell_kind = "gru"

    def forward(self, input, hx):
        out, h = rnn_sequence(self._single_step_spec(), input.unsqueeze(1), hx, None, self.flat_parameters())
        return h


class _TTRNNBase(nn.Module):
    cell_cls = None
    cell_kind = ""

    def __init__(self, input_size, hidden_size, num_layers, device, n_cores, tt_rank, bias=True,
                 is_naive=False, log_grads=False, new_core=None):
        assert new_core in [None, 'first', 'last']
        super().__init__()
        if log_grads:
            raise NotImplementedError("log_grads=True (per-timestep activation/gradient hooks) is not implemented "
                                      "in tensorized_rnn_b200 yet: the fused sequence kernel has no per-step modules")
        self.n_cores = n_cores
        self.tt_rank = tt_rank
        self.is_naive = is_naive
        self.new_core = new_core
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.bias = bias
        self.device = device
        self.log_grads = log_grads
        self._all_layers = []
        for i in range(self.num_layers):
            cell = self._create_first_layer_cell() if i == 0 else self._create_other_layer_cell()
            setattr(self, 'cell{}'.format(i), cell)
            self._all_layers.append(cell)
        self._spec: Optional[RnnSpec] = None

    def _make_cell(self, in_size):
        return self.cell_cls(in_size, self.hidden_size, self.bias, self.device, n_cores=self.n_cores,
                             tt_rank=self.tt_rank, is_naive=self.is_naive, new_core=self.new_core)

    def _create_first_layer_cell(self):
        return self._make_cell(self.input_size)

    def _create_other_layer_cell(self):
        return self._make_cell(self.hidden_size)

    def param_count(self):
        total = 0
        for cell in self._all_layers:
            for attr in ('input_weights', 'hidden_weights'):
                total += param_count(getattr(cell, attr))
        return total

    def spec(self) -> RnnSpec:
        if self._spec is None:
            self._spec = RnnSpec(self.cell_kind, self.input_size, self.hidden_size, bool(self.bias),
                                 [c.input_weights.tt_modes() for c in self._all_layers],
                                 [c.hidden_weights.tt_modes() for c in self._all_layers])
        return self._spec

    def flat_parameters(self) -> List[torch.Tensor]:
        out: List[torch.Tensor] = []
        for cell in self._all_layers:
            out += cell.flat_parameters()
        return out

    def _check_input(self, input):
        if input.dim() != 3:
            raise ValueError("input must be (batch_size, seq_len, input_size)")
        if input.size(1) == 0:
            # the reference falls off its time loop with unbound locals (lstm.py:135 / gru.py:136)
            raise NameError("zero-length sequence: the reference raises NameError here and so does this module")


class TTLSTM(_TTRNNBase):
    """Drop-in for reference tensorized_rnn.tt_lstm.TTLSTM (tt_lstm.py:43-62)."""
    cell_cls = TTLSTMCell
    cell_kind = "lstm"

    def init_hidden(self, batch_size):
        h = torch.zeros(batch_size, self.hidden_size).to(self.device)
        c = torch.zeros(batch_size, self.hidden_size).to(self.device)
        return h, c

    def forward(self, input, init_states=None):
        """input (batch, seq_len, input_size) -> outputs (batch, seq_len, hidden), (h_T, c_T) of the last layer.
        `init_states` = (h0, c0), each (batch, hidden), shared by every layer; None = zeros."""
        self._check_input(input)
        h0, c0 = (None, None) if init_states is None else init_states
        out, h, c = rnn_sequence(self.spec(), input, h0, c0, self.flat_parameters())
        return out, (h, c)


class TTGRU(_TTRNNBase):
    """Drop-in for reference tensorized_rnn.gru.TTGRU (gru.py:175-194)."""
    cell_cls = TTGRUCell
    cell_kind = "gru"

    def init_hidden(self, batch_size):
        return torch.zeros(batch_size, self.hidden_size).to(self.device)

    def forward(self, input, init_states=None):
        """input (batch, seq_len, input_size) -> outputs (batch, seq_len, hidden), h_T of the last layer."""
        self._check_input(input)
        out, h = rnn_sequence(self.spec(), input, init_states, None, self.flat_parameters())
        return out, h
