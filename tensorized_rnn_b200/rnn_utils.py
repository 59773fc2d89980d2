"""Per-timestep activation / gradient logging for `log_grads=True`.

Mirror of the logging half of reference tensorized_rnn/rnn_utils.py:42-226 (`ActivGradLogger`,
`av_norm`): the same class-level registry, hook factory and minibatch / epoch aggregation, so the
reference's training script calls (`ActivGradLogger.end_minibatch()`, `.end_epoch()`,
`.get_logs()`, pmnist_test.py:162,180,231) work unchanged against the B200 modules.  The hooks
fire because `log_grads=True` switches the modules to cell-step mode (one cell call per step).
The logged quantities are diagnostics, not part of the compute path.
"""
from __future__ import annotations

from collections import deque

import torch


def av_norm(tensor, average_logs=False):
    """mean over the batch of ||t_b||^2 (or of log ||t_b||^2) -- reference rnn_utils.py:217-226."""
    norms = (tensor ** 2).sum(list(range(1, tensor.dim())))
    if average_logs:
        norms = torch.log(norms)
    assert norms.dim() == 1
    return norms.mean()


class ActivGradLogger(object):
    """Reference rnn_utils.py:42-215.  One logger per (variable, layer), registered by name."""
    all_loggers = dict()
    QUANTITIES = ('act', 'log_act', 'grad', 'log_grad')

    @staticmethod
    def get_logs():
        """{(variable, quantity): (num_epochs, seq_len) tensor} -- reference rnn_utils.py:46-70."""
        log_dict = dict()
        for var, logger in ActivGradLogger.all_loggers.items():
            for qnt in ActivGradLogger.QUANTITIES:
                log_dict[(var, qnt)] = torch.stack(getattr(logger, qnt + '_epoch'))
        return log_dict

    def __init__(self, name):
        assert name not in ActivGradLogger.all_loggers
        ActivGradLogger.all_loggers[name] = self
        self.name = name
        self.act_epoch, self.grad_epoch, self.log_act_epoch, self.log_grad_epoch = [], [], [], []
        self.act_mini, self.grad_mini, self.log_act_mini, self.log_grad_mini = [], [], [], []
        self.act, self.log_act = [], []
        self.grad, self.log_grad = deque(), deque()

    @staticmethod
    def get_logger(name):
        if name in ActivGradLogger.all_loggers:
            return ActivGradLogger.all_loggers[name]
        print("Logger '{}' not yet initialized".format(name))

    @staticmethod
    def end_epoch():
        for logger in ActivGradLogger.all_loggers.values():
            logger._end_epoch()

    @staticmethod
    def end_minibatch():
        for logger in ActivGradLogger.all_loggers.values():
            logger._end_minibatch()

    @staticmethod
    def del_record():
        for logger in ActivGradLogger.all_loggers.values():
            logger._del_record()

    @staticmethod
    def reset():
        """Forget every registered logger (the reference keeps them for the life of the process;
        tests and repeated model construction need a way to start over)."""
        ActivGradLogger.all_loggers.clear()

    def create_hooks(self, output_ind):
        """(forward hook for the cell module, backward hook for its output tensor) -- rnn_utils.py:127-172."""
        @torch.no_grad()
        def forward_hook(rnn_cell, inputs, outputs):
            if not isinstance(outputs, tuple):
                assert isinstance(outputs, torch.Tensor) and output_ind == 0
                outputs = (outputs,)
            target = outputs[output_ind].detach()
            self.act.append(av_norm(target))
            self.log_act.append(av_norm(target, average_logs=True))

        @torch.no_grad()
        def backward_hook(grad_out):
            target = grad_out.detach()
            # gradients arrive in reverse time order
            self.grad.appendleft(av_norm(target))
            self.log_grad.appendleft(av_norm(target, average_logs=True))

        return forward_hook, backward_hook

    def _end_epoch(self):
        self.act_epoch.append(torch.mean(torch.stack(self.act_mini), 0))
        self.grad_epoch.append(torch.mean(torch.stack(self.grad_mini), 0))
        self.log_act_epoch.append(torch.mean(torch.stack(self.log_act_mini), 0))
        self.log_grad_epoch.append(torch.mean(torch.stack(self.log_grad_mini), 0))
        self.act_mini, self.grad_mini, self.log_act_mini, self.log_grad_mini = [], [], [], []

    def _end_minibatch(self):
        if self.act_mini:
            assert len(self.act_mini[0]) == len(self.act) and len(self.grad_mini[0]) == len(self.grad)
        self.act_mini.append(torch.stack(self.act))
        self.log_act_mini.append(torch.stack(self.log_act))
        self.grad_mini.append(torch.stack(list(self.grad)))
        self.log_grad_mini.append(torch.stack(list(self.log_grad)))
        self._del_record()

    def _del_record(self):
        del self.act[:]
        del self.log_act[:]
        self.grad.clear()
        self.log_grad.clear()
