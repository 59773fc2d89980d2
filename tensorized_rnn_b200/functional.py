"""autograd glue between the nn.Module mirror and the C ABI.

One `torch.autograd.Function` per whole-sequence call: forward packs every core and bias into one
contiguous FP32 blob (the layout of include/ttrnn_b200.h), lets torch own every buffer, and calls
`ttrnn_rnn_forward`; backward calls `ttrnn_rnn_backward` and hands each parameter its slice of the
gradient blob.  PyTorch is plumbing here (memory, streams, autograd bookkeeping); all arithmetic of
the path runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda_f32(name: str, t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError("tensorized_rnn_b200: `%s` is on %s; the engine runs on CUDA (sm_100a) only and has "
                           "no CPU fallback" % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError("tensorized_rnn_b200: `%s` must be float32 (the path is FP32 end to end), got %s"
                        % (name, t.dtype))


class RnnSpec(object):
    """Static description of a TT-RNN stack: cell kind, sizes and the TT shape of every weight."""

    def __init__(self, cell: str, input_size: int, hidden_size: int, has_bias: bool,
                 ih_shapes: Sequence[Tuple[Sequence[int], Sequence[int], Sequence[int]]],
                 hh_shapes: Sequence[Tuple[Sequence[int], Sequence[int], Sequence[int]]]):
        assert cell in ("lstm", "gru")
        self.cell = cell
        self.n_gates = 4 if cell == "lstm" else 3
        self.input_size, self.hidden_size, self.has_bias = int(input_size), int(hidden_size), bool(has_bias)
        self.num_layers = len(ih_shapes)
        if self.num_layers > _lib.MAX_LAYERS:
            raise ValueError("at most %d layers are supported" % _lib.MAX_LAYERS)
        self.ih_shapes = [tuple(list(map(int, v)) for v in s) for s in ih_shapes]   # (in_modes, out_modes, ranks)
        self.hh_shapes = [tuple(list(map(int, v)) for v in s) for s in hh_shapes]
        self._cache = {}

    def desc(self, batch: int, seq_len: int) -> _lib.RnnDesc:
        key = (batch, seq_len)
        d = self._cache.get(key)
        if d is None:
            d = _lib.RnnDesc()
            d.cell = _lib.CELL_LSTM if self.cell == "lstm" else _lib.CELL_GRU
            d.num_layers, d.input_size, d.hidden_size = self.num_layers, self.input_size, self.hidden_size
            d.has_bias, d.seq_len, d.batch = int(self.has_bias), seq_len, batch
            for l in range(self.num_layers):
                d.ih[l] = _lib.make_tt_shape(*self.ih_shapes[l])
                d.hh[l] = _lib.make_tt_shape(*self.hh_shapes[l])
            if len(self._cache) > 64:
                self._cache.clear()
            self._cache[key] = d
        return d


def _bytes_to_floats(nbytes: int) -> int:
    return max(1, (int(nbytes) + 3) // 4)


def _dense16(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Contiguous and 16-byte aligned: the kernels read these buffers with float4 loads, 16-byte cp.async and TMA.
    `.contiguous()` keeps a contiguous view whose storage offset is not a multiple of 4 floats (a slice of a flat
    buffer); such a view is copied."""
    if t is None:
        return None
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


def _step_stats(lib, v: torch.Tensor, stream) -> torch.Tensor:
    """(2, T): mean_b ||v[b,t]||^2 and mean_b log ||v[b,t]||^2 of a contiguous (B,T,H) block (ttrnn_step_norms)."""
    B, T, H = v.shape
    out = torch.empty((2, T), device=v.device, dtype=torch.float32)
    _lib.check(lib.ttrnn_step_norms(_ptr(v), B, T, H, _ptr(out), stream), "ttrnn_step_norms")
    return out


class _RnnFunction(torch.autograd.Function):
    """`step_log` (None or an object with .activations(layer, var, stats) / .gradients(layer, var, stats)): fused
    log_grads.  The training forward reports the per-step statistics of every layer's h_t / c_t sequence, the
    backward those of the total gradients reaching them, (2, T) tensors each (reference rnn_utils.py:127-172)."""

    @staticmethod
    def forward(ctx, spec: RnnSpec, grad_enabled: bool, step_log, x, h0, c0, *params):
        lib = _lib.load()
        _require_cuda_f32("input", x)
        for i, p in enumerate(params):
            _require_cuda_f32("parameter %d" % i, p)
            if p.device != x.device:
                raise RuntimeError("parameter %d is on %s but the input is on %s" % (i, p.device, x.device))
        B, T, I = x.shape
        H = spec.hidden_size
        if I != spec.input_size:
            raise ValueError("input has %d features, the module was built for %d" % (I, spec.input_size))
        for name, s in (("h0", h0), ("c0", c0)):
            if s is not None:
                _require_cuda_f32(name, s)
                if tuple(s.shape) != (B, H):
                    raise ValueError("%s must have shape (%d, %d), got %s" % (name, B, H, tuple(s.shape)))
        x = _dense16(x)
        h0c = _dense16(h0)
        c0c = _dense16(c0)
        desc = spec.desc(B, T)
        with torch.cuda.device(x.device):
            blob = torch.cat([p.detach().reshape(-1) for p in params])
            want = lib.ttrnn_rnn_param_count(C.byref(desc))
            if want < 0:
                raise RuntimeError("ttrnn_rnn_param_count failed: " + _lib.last_error())
            if want != blob.numel():
                raise RuntimeError("parameter blob has %d floats, descriptor expects %d" % (blob.numel(), want))
            ws = _lib.RnnWorkspace()
            _lib.check(lib.ttrnn_rnn_workspace_bytes_ex(C.byref(desc), _lib.WS_WHOLE_BATCH if step_log is not None else 0,
                                                        C.byref(ws)), "ttrnn_rnn_workspace_bytes")
            # grad mode is always off inside Function.forward and needs_input_grad stays True for nn.Parameters under
            # torch.no_grad(): the caller captures torch.is_grad_enabled() before .apply
            training = bool(grad_enabled) and any(ctx.needs_input_grad[3:])
            if step_log is not None and not training:
                raise RuntimeError("fused per-step logging needs a training forward (it reads the kept h_t / c_t sequences)")
            out = torch.empty((B, T, H), device=x.device, dtype=torch.float32)
            hT = torch.empty((B, H), device=x.device, dtype=torch.float32)
            cT = torch.empty((B, H), device=x.device, dtype=torch.float32) if spec.cell == "lstm" else None
            saved = torch.empty(_bytes_to_floats(ws.saved_bytes), device=x.device, dtype=torch.float32) \
                if training else None
            scratch = torch.empty(_bytes_to_floats(ws.fwd_scratch_bytes), device=x.device, dtype=torch.float32)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.ttrnn_rnn_forward(C.byref(desc), C.byref(ws), _ptr(x), _ptr(h0c), _ptr(c0c), _ptr(blob), _ptr(out),
                                             _ptr(hT), _ptr(cT), _ptr(saved), _ptr(scratch), stream),
                       "ttrnn_rnn_forward")
            if step_log is not None:
                hs_off, cs_off = C.c_int64(), C.c_int64()
                for l in range(spec.num_layers):
                    _lib.check(lib.ttrnn_rnn_saved_layout(C.byref(desc), C.byref(ws), l, C.byref(hs_off), C.byref(cs_off)),
                               "ttrnn_rnn_saved_layout")
                    hs = out if hs_off.value < 0 else saved[hs_off.value:hs_off.value + B * T * H].view(B, T, H)
                    step_log.activations(l, "h", _step_stats(lib, hs, stream))
                    if cs_off.value >= 0:
                        step_log.activations(l, "c", _step_stats(lib, saved[cs_off.value:cs_off.value + B * T * H].view(B, T, H), stream))
        if training:
            ctx.spec = spec
            ctx.step_log = step_log
            ctx.shapes = [tuple(p.shape) for p in params]
            ctx.ws = ws                    # sizes + execution plan: backward runs from the same plan as this forward
            ctx.bwd_floats = _bytes_to_floats(ws.bwd_scratch_bytes)
            ctx.has_h0, ctx.has_c0 = h0 is not None, c0 is not None
            ctx.set_materialize_grads(False)
            ctx.save_for_backward(x, h0c, c0c, blob, out, saved)
        if spec.cell == "lstm":
            return out, hT, cT
        return out, hT

    @staticmethod
    def backward(ctx, *grads):
        lib = _lib.load()
        spec: RnnSpec = ctx.spec
        x, h0, c0, blob, out, saved = ctx.saved_tensors
        d_out = grads[0]
        d_hT = grads[1]
        d_cT = grads[2] if spec.cell == "lstm" else None
        B, T, _ = x.shape
        H = spec.hidden_size
        desc = spec.desc(B, T)
        d_out = _dense16(d_out)            # the BPTT kernels stage dOut tiles with 16-byte cp.async
        d_hT = _dense16(d_hT)
        d_cT = _dense16(d_cT)
        with torch.cuda.device(x.device):
            d_blob = torch.empty_like(blob)
            d_x = torch.empty_like(x) if ctx.needs_input_grad[3] else None
            d_h0 = torch.empty((B, H), device=x.device, dtype=torch.float32) \
                if (ctx.has_h0 and ctx.needs_input_grad[4]) else None
            d_c0 = torch.empty((B, H), device=x.device, dtype=torch.float32) \
                if (spec.cell == "lstm" and ctx.has_c0 and ctx.needs_input_grad[5]) else None
            scratch = torch.empty(ctx.bwd_floats, device=x.device, dtype=torch.float32)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            args = (C.byref(desc), C.byref(ctx.ws), _ptr(x), _ptr(h0), _ptr(c0), _ptr(blob), _ptr(out),
                    _ptr(saved), _ptr(d_out), _ptr(d_hT), _ptr(d_cT), _ptr(d_blob),
                    _ptr(d_x), _ptr(d_h0), _ptr(d_c0), _ptr(scratch), stream)
            if ctx.step_log is None:
                _lib.check(lib.ttrnn_rnn_backward(*args), "ttrnn_rnn_backward")
            else:
                L = spec.num_layers
                dh_log = torch.empty((L, B, T, H), device=x.device, dtype=torch.float32)
                dc_log = torch.empty((L, B, T, H), device=x.device, dtype=torch.float32) if spec.cell == "lstm" else None
                _lib.check(lib.ttrnn_rnn_backward_logged(*(args + (_ptr(dh_log), _ptr(dc_log)))), "ttrnn_rnn_backward_logged")
                for l in range(L):
                    ctx.step_log.gradients(l, "h", _step_stats(lib, dh_log[l], stream))
                    if dc_log is not None:
                        ctx.step_log.gradients(l, "c", _step_stats(lib, dc_log[l], stream))
        d_params: List[Optional[torch.Tensor]] = []
        off = 0
        for i, shp in enumerate(ctx.shapes):
            n = 1
            for v in shp:
                n *= v
            d_params.append(d_blob[off:off + n].view(shp) if ctx.needs_input_grad[6 + i] else None)
            off += n
        return (None, None, None, d_x, d_h0, d_c0) + tuple(d_params)


def rnn_sequence(spec: RnnSpec, x: torch.Tensor, h0: Optional[torch.Tensor], c0: Optional[torch.Tensor],
                 params: Sequence[torch.Tensor], step_log=None):
    """Run the whole stack over the whole sequence.  Returns (out, hT[, cT]).  `step_log`: see _RnnFunction."""
    if x.dim() != 3:
        raise ValueError("input must be (batch, seq_len, input_size), got shape %s" % (tuple(x.shape),))
    return _RnnFunction.apply(spec, torch.is_grad_enabled(), step_log, x, h0, c0, *params)


class _TTLinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, shape: _lib.TTShape, n_in: int, n_out: int, x, bias, *cores):
        lib = _lib.load()
        _require_cuda_f32("input", x)
        for i, p in enumerate(cores):
            _require_cuda_f32("core %d" % i, p)
        if bias is not None:
            _require_cuda_f32("bias", bias)
        if x.dim() != 2 or x.shape[1] != n_in:
            raise ValueError("input must be (rows, %d), got %s" % (n_in, tuple(x.shape)))
        x = _dense16(x)
        rows = x.shape[0]
        with torch.cuda.device(x.device):
            blob = torch.cat([p.detach().reshape(-1) for p in cores])
            b = None if bias is None else bias.detach().contiguous()
            y = torch.empty((rows, n_out), device=x.device, dtype=torch.float32)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.ttrnn_ttlinear_forward(C.byref(shape), rows, _ptr(x), _ptr(blob), _ptr(b), _ptr(y),
                                                  None, stream), "ttrnn_ttlinear_forward")
        if any(ctx.needs_input_grad[3:]):
            ctx.shape, ctx.core_shapes, ctx.has_bias = shape, [tuple(p.shape) for p in cores], bias is not None
            ctx.save_for_backward(x, blob)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, blob = ctx.saved_tensors
        rows = x.shape[0]
        dy = _dense16(dy)
        with torch.cuda.device(x.device):
            nbytes = lib.ttrnn_ttlinear_workspace_bytes(C.byref(ctx.shape), rows)
            if nbytes < 0:
                raise RuntimeError("ttrnn_ttlinear_workspace_bytes failed: " + _lib.last_error())
            scratch = torch.empty(_bytes_to_floats(nbytes), device=x.device, dtype=torch.float32)
            d_x = torch.empty_like(x) if ctx.needs_input_grad[3] else None
            d_blob = torch.empty_like(blob)
            d_bias = torch.empty(dy.shape[1], device=x.device, dtype=torch.float32) \
                if (ctx.has_bias and ctx.needs_input_grad[4]) else None
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(lib.ttrnn_ttlinear_backward(C.byref(ctx.shape), rows, _ptr(x), _ptr(blob), _ptr(dy), _ptr(d_x),
                                                   _ptr(d_blob), _ptr(d_bias), _ptr(scratch), stream),
                       "ttrnn_ttlinear_backward")
        d_cores, off = [], 0
        for i, shp in enumerate(ctx.core_shapes):
            n = 1
            for v in shp:
                n *= v
            d_cores.append(d_blob[off:off + n].view(shp) if ctx.needs_input_grad[5 + i] else None)
            off += n
        return (None, None, None, d_x, d_bias) + tuple(d_cores)


def ttlinear(shape: _lib.TTShape, n_in: int, n_out: int, x: torch.Tensor, bias: Optional[torch.Tensor],
             cores: Sequence[torch.Tensor]) -> torch.Tensor:
    return _TTLinearFunction.apply(shape, n_in, n_out, x, bias, *cores)


class _CellStepFunction(torch.autograd.Function):
    """Gate math + state update of ONE timestep from the two pre-activation blocks (cell-step mode:
    is_naive / log_grads).  Reference: lstm.py:26-32, gru.py:33-44; backward = SURVEY.md 8a-10."""

    @staticmethod
    def forward(ctx, cell: str, c_grad_hook, a, u, h_prev, c_prev):
        lib = _lib.load()
        for name, t in (("input projection", a), ("hidden projection", u), ("hx", h_prev)):
            _require_cuda_f32(name, t)
        lstm = cell == "lstm"
        if lstm:
            _require_cuda_f32("cx", c_prev)
        B, H = h_prev.shape
        G = 4 if lstm else 3
        if tuple(a.shape) != (B, G * H) or tuple(u.shape) != (B, G * H):
            raise ValueError("gate blocks must be (%d, %d), got %s and %s" % (B, G * H, tuple(a.shape), tuple(u.shape)))
        a, u, h_prev = a.contiguous(), u.contiguous(), h_prev.contiguous()
        c_prev = c_prev.contiguous() if lstm else None
        with torch.cuda.device(a.device):
            h = torch.empty_like(h_prev)
            c = torch.empty_like(h_prev) if lstm else None
            stream = torch.cuda.current_stream(a.device).cuda_stream
            _lib.check(lib.ttrnn_cell_forward(_lib.CELL_LSTM if lstm else _lib.CELL_GRU, B, H, _ptr(a), _ptr(u),
                                              _ptr(h_prev), _ptr(c_prev), _ptr(h), _ptr(c), stream),
                       "ttrnn_cell_forward")
        ctx.cell, ctx.c_grad_hook = cell, c_grad_hook
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(a, u, h_prev, c_prev)
        return (h, c) if lstm else h

    @staticmethod
    def backward(ctx, *grads):
        lib = _lib.load()
        a, u, h_prev, c_prev = ctx.saved_tensors
        lstm = ctx.cell == "lstm"
        dh = grads[0]
        dc = grads[1] if lstm else None
        B, H = h_prev.shape
        dh = None if dh is None else dh.contiguous()
        dc = None if dc is None else dc.contiguous()
        with torch.cuda.device(a.device):
            da, du = torch.empty_like(a), torch.empty_like(u)
            dh_prev = torch.empty_like(h_prev)
            dc_prev = torch.empty_like(h_prev) if lstm else None
            dc_total = torch.empty_like(h_prev) if (lstm and ctx.c_grad_hook is not None) else None
            stream = torch.cuda.current_stream(a.device).cuda_stream
            _lib.check(lib.ttrnn_cell_backward(_lib.CELL_LSTM if lstm else _lib.CELL_GRU, B, H, _ptr(a), _ptr(u),
                                               _ptr(h_prev), _ptr(c_prev), _ptr(dh), _ptr(dc), _ptr(da), _ptr(du),
                                               _ptr(dh_prev), _ptr(dc_prev), _ptr(dc_total), stream),
                       "ttrnn_cell_backward")
        if dc_total is not None:
            ctx.c_grad_hook(dc_total)      # total dL/dc_t, as a tensor hook on the reference's `cy` sees it
        return None, None, da, du, (None if lstm else dh_prev), dc_prev


def cell_step(cell: str, a: torch.Tensor, u: torch.Tensor, h_prev: torch.Tensor,
              c_prev: Optional[torch.Tensor] = None, c_grad_hook=None):
    """One LSTM / GRU step from a = W_ih x + b_ih and u = W_hh h + b_hh.  LSTM -> (h, c); GRU -> h.
    `c_grad_hook(grad)` (LSTM) is called during backward with the total gradient of c_t."""
    return _CellStepFunction.apply(cell, c_grad_hook, a, u, h_prev, c_prev)
