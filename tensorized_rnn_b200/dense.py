"""Dense LSTM / GRU baselines through the same engine.

Mirrors the reference's `LSTM` / `LSTMCell` (tensorized_rnn/lstm.py:7-41, 44-135) and `GRU` / `GRUCell`
(tensorized_rnn/gru.py:11-50, 52-136): constructor signatures, attribute and sub-module names (`cell{i}.input_weights`,
`cell{i}.hidden_weights` are `nn.Linear`s, so the state_dict keys are `cell{i}.input_weights.weight` ...), the bias
convention (the dense LSTM cell gives its input map NO bias, lstm.py:17-18; the GRU cell gives both maps `bias`,
gru.py:19-23), `init_hidden`, `param_count`, and the forward contract of lstm.py:101-135 / gru.py:104-136.  These are the
modules `pmnist_test.py` builds without `--tt` and `SpeakerEncoder` builds with `compression=None`: with them the
paper's dense-vs-TT comparison runs on the same device path (C ABI `ttrnn_dense_rnn_*`, csrc/tt_dense.cuh).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
from torch import nn

from . import _lib
from .functional import _bytes_to_floats, _dense16, _ptr, _require_cuda_f32


class _DenseRnnFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cell: str, hidden_size: int, num_layers: int, has_bias: bool, x, h0, c0, *params):
        lib = _lib.load()
        _require_cuda_f32("input", x)
        for i, p in enumerate(params):
            _require_cuda_f32("parameter %d" % i, p)
        if x.dim() != 3:
            raise ValueError("input must be (batch, seq_len, input_size), got shape %s" % (tuple(x.shape),))
        B, T, I = x.shape
        H = hidden_size
        for name, s in (("h0", h0), ("c0", c0)):
            if s is not None:
                _require_cuda_f32(name, s)
                if tuple(s.shape) != (B, H):
                    raise ValueError("%s must have shape (%d, %d), got %s" % (name, B, H, tuple(s.shape)))
        x, h0c, c0c = _dense16(x), _dense16(h0), _dense16(c0)
        desc = _lib.DenseDesc(_lib.CELL_LSTM if cell == "lstm" else _lib.CELL_GRU, num_layers, I, H, int(has_bias), T, B)
        with torch.cuda.device(x.device):
            blob = torch.cat([p.detach().reshape(-1) for p in params])
            want = lib.ttrnn_dense_rnn_param_count(C.byref(desc))
            if want < 0:
                raise ValueError("ttrnn_dense_rnn_param_count failed: " + _lib.last_error())
            if want != blob.numel():
                raise RuntimeError("parameter blob has %d floats, descriptor expects %d" % (blob.numel(), want))
            sb, wb = C.c_int64(0), C.c_int64(0)
            _lib.check(lib.ttrnn_dense_rnn_workspace_bytes(C.byref(desc), C.byref(sb), C.byref(wb)), "ttrnn_dense_rnn_workspace_bytes")
            out = torch.empty((B, T, H), device=x.device, dtype=torch.float32)
            hT = torch.empty((B, H), device=x.device, dtype=torch.float32)
            cT = torch.empty((B, H), device=x.device, dtype=torch.float32) if cell == "lstm" else None
            saved = torch.empty(_bytes_to_floats(sb.value), device=x.device, dtype=torch.float32)
            scratch = torch.empty(_bytes_to_floats(wb.value), device=x.device, dtype=torch.float32)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.ttrnn_dense_rnn_forward(C.byref(desc), _ptr(x), _ptr(h0c), _ptr(c0c), _ptr(blob), _ptr(out), _ptr(hT),
                                                   _ptr(cT), _ptr(saved), _ptr(scratch), stream), "ttrnn_dense_rnn_forward")
        ctx.desc, ctx.cell = desc, cell
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.scratch_floats = _bytes_to_floats(wb.value)
        ctx.has_h0, ctx.has_c0 = h0 is not None, c0 is not None
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, h0c, c0c, blob, out, saved)
        return (out, hT, cT) if cell == "lstm" else (out, hT)

    @staticmethod
    def backward(ctx, *grads):
        lib = _lib.load()
        x, h0, c0, blob, out, saved = ctx.saved_tensors
        if getattr(ctx, "consumed", False):
            raise RuntimeError("dense LSTM/GRU: the saved activations were consumed by a previous backward "
                               "(retain_graph is not supported by this path)")
        ctx.consumed = True
        lstm = ctx.cell == "lstm"
        d_out, d_hT = _dense16(grads[0]), _dense16(grads[1])
        d_cT = _dense16(grads[2]) if lstm else None
        B, H = x.shape[0], out.shape[2]
        with torch.cuda.device(x.device):
            d_blob = torch.empty_like(blob)
            d_x = torch.empty_like(x) if ctx.needs_input_grad[4] else None
            d_h0 = torch.empty((B, H), device=x.device, dtype=torch.float32) if (ctx.has_h0 and ctx.needs_input_grad[5]) else None
            d_c0 = torch.empty((B, H), device=x.device, dtype=torch.float32) \
                if (lstm and ctx.has_c0 and ctx.needs_input_grad[6]) else None
            scratch = torch.empty(ctx.scratch_floats, device=x.device, dtype=torch.float32)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.ttrnn_dense_rnn_backward(C.byref(ctx.desc), _ptr(x), _ptr(h0), _ptr(c0), _ptr(blob), _ptr(out), _ptr(saved),
                                                    _ptr(d_out), _ptr(d_hT), _ptr(d_cT), _ptr(d_blob), _ptr(d_x), _ptr(d_h0),
                                                    _ptr(d_c0), _ptr(scratch), stream), "ttrnn_dense_rnn_backward")
        d_params, off = [], 0
        for i, shp in enumerate(ctx.shapes):
            n = 1
            for v in shp:
                n *= v
            d_params.append(d_blob[off:off + n].view(shp) if ctx.needs_input_grad[7 + i] else None)
            off += n
        return (None, None, None, None, d_x, d_h0, d_c0) + tuple(d_params)


class _DenseCell(nn.Module):
    n_gate = 0

    def __init__(self, input_size, hidden_size, bias, device):
        super().__init__()
        self.input_size, self.hidden_size, self.bias, self.device = input_size, hidden_size, bias, device
        self.input_weights = self._create_input_hidden_weights()
        self.hidden_weights = self._create_hidden_hidden_weights()

    def flat_parameters(self) -> List[torch.Tensor]:
        out = [self.input_weights.weight]
        if self.input_weights.bias is not None:
            out.append(self.input_weights.bias)
        out.append(self.hidden_weights.weight)
        if self.hidden_weights.bias is not None:
            out.append(self.hidden_weights.bias)
        return out


class LSTMCell(_DenseCell):
    """Weights of one dense LSTM layer (reference lstm.py:7-21).  The step itself runs inside the sequence call."""
    n_gate = 4

    def _create_input_hidden_weights(self):
        return nn.Linear(self.input_size, 4 * self.hidden_size, False).to(self.device)      # lstm.py:18: no bias

    def _create_hidden_hidden_weights(self):
        return nn.Linear(self.hidden_size, 4 * self.hidden_size, self.bias).to(self.device)


class GRUCell(_DenseCell):
    """Weights of one dense GRU layer (reference gru.py:11-23)."""
    n_gate = 3

    def _create_input_hidden_weights(self):
        return nn.Linear(self.input_size, 3 * self.hidden_size, self.bias).to(self.device)

    def _create_hidden_hidden_weights(self):
        return nn.Linear(self.hidden_size, 3 * self.hidden_size, self.bias).to(self.device)


class _DenseRnnBase(nn.Module):
    cell_kind = ""
    cell_cls = None

    def __init__(self, input_size, hidden_size, num_layers, device, bias=True, log_grads=False):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.bias, self.device, self.log_grads = bias, device, log_grads
        if log_grads:
            raise NotImplementedError("log_grads=True is implemented for the TT modules (TTLSTM / TTGRU); the dense baselines "
                                      "run as one fused sequence call and expose no per-step hooks")
        self._all_layers = []
        for i in range(num_layers):
            cell = self.cell_cls(input_size if i == 0 else hidden_size, hidden_size, bias, device)
            setattr(self, "cell{}".format(i), cell)
            self._all_layers.append(cell)

    def param_count(self):
        return int(sum(p.numel() for c in self._all_layers for p in c.parameters()))

    def flat_parameters(self):
        return [p for c in self._all_layers for p in c.flat_parameters()]

    def _run(self, input, h0, c0):
        if input.shape[1] == 0:
            raise NameError("name 'x' is not defined")       # the reference fails the same way on T = 0 (lstm.py:135)
        return _DenseRnnFunction.apply(self.cell_kind, self.hidden_size, self.num_layers, bool(self.bias), input, h0, c0,
                                       *self.flat_parameters())


class LSTM(_DenseRnnBase):
    """Dense LSTM stack, drop-in for reference tensorized_rnn.lstm.LSTM (lstm.py:44-135)."""
    cell_kind, cell_cls = "lstm", LSTMCell

    def init_hidden(self, batch_size):
        h = torch.zeros(batch_size, self.hidden_size).to(self.device)
        return h, torch.zeros(batch_size, self.hidden_size).to(self.device)

    def forward(self, input, init_states=None):
        h0, c0 = init_states if init_states is not None else (None, None)
        out, h, c = self._run(input, h0, c0)
        return out, (h, c)


class GRU(_DenseRnnBase):
    """Dense GRU stack, drop-in for reference tensorized_rnn.gru.GRU (gru.py:52-136)."""
    cell_kind, cell_cls = "gru", GRUCell

    def init_hidden(self, batch_size):
        return torch.zeros(batch_size, self.hidden_size).to(self.device)

    def forward(self, input, init_states=None):
        out, h = self._run(input, init_states, None)
        return out, h
