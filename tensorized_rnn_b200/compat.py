"""Switch an importable copy of onucharles/tensorized-rnn over to the B200 modules.

    import tensorized_rnn_b200.compat as compat
    compat.patch_reference()          # after this, the reference's callers build the B200 modules

The reference's callers pick the recurrent classes by name at construction time
(`experiments/digit_classification/mnist_classifier.py:19-35`,
`experiments/speaker_verification/encoder/speaker_encoder.py:41-48`), so replacing the names in the
modules they import from is enough: `tensorized_rnn.tt_lstm.TTLSTM`, `tensorized_rnn.gru.TTGRU`,
`t3nsor.layers.TTLinear` (also re-exported as `t3nsor.TTLinear`), `tensorized_rnn.tt_linearset.TTLinearSet`, the dense
baselines `tensorized_rnn.lstm.LSTM` / `tensorized_rnn.gru.GRU` (with their cells)
and the logging registry `tensorized_rnn.rnn_utils.ActivGradLogger` (so that the training script's
`ActivGradLogger.end_minibatch()` / `.get_logs()` calls see the loggers the B200 modules register).
`unpatch_reference()` restores the originals.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, Tuple

from . import dense, layers, rnn, rnn_utils

_TARGETS = (
    ("tensorized_rnn.tt_lstm", "TTLSTM", rnn.TTLSTM),
    ("tensorized_rnn.tt_lstm", "TTLSTMCell", rnn.TTLSTMCell),
    ("tensorized_rnn.gru", "TTGRU", rnn.TTGRU),
    ("tensorized_rnn.gru", "TTGRUCell", rnn.TTGRUCell),
    ("tensorized_rnn.tt_linearset", "TTLinearSet", layers.TTLinearSet),
    # dense baselines (pmnist_test.py without --tt, SpeakerEncoder(compression=None))
    ("tensorized_rnn.lstm", "LSTM", dense.LSTM),
    ("tensorized_rnn.lstm", "LSTMCell", dense.LSTMCell),
    ("tensorized_rnn.gru", "GRU", dense.GRU),
    ("tensorized_rnn.gru", "GRUCell", dense.GRUCell),
    ("tensorized_rnn.rnn_utils", "ActivGradLogger", rnn_utils.ActivGradLogger),
    ("tensorized_rnn.lstm", "ActivGradLogger", rnn_utils.ActivGradLogger),
    ("tensorized_rnn.gru", "ActivGradLogger", rnn_utils.ActivGradLogger),
    ("t3nsor.layers", "TTLinear", layers.TTLinear),
    ("t3nsor", "TTLinear", layers.TTLinear),
)
_saved: Dict[Tuple[str, str], object] = {}


def patch_reference(reference_path: str = None) -> int:
    """Replace the reference's TT classes by the B200 ones.  `reference_path` is prepended to sys.path if given.
    Also rebinds the names in already-imported caller modules that did `from ... import TTLSTM`.
    Returns the number of names replaced."""
    if reference_path and reference_path not in sys.path:
        sys.path.insert(0, reference_path)
    n = 0
    for mod_name, attr, new in _TARGETS:
        try:
            mod = importlib.import_module(mod_name)
        except Exception:
            continue
        if hasattr(mod, attr) and getattr(mod, attr) is not new:
            _saved.setdefault((mod_name, attr), getattr(mod, attr))
            setattr(mod, attr, new)
            n += 1
    # callers that imported the names directly before the patch
    originals = {id(v): new for (m, a), v in _saved.items() for (mm, aa, new) in _TARGETS if (mm, aa) == (m, a)}
    for mod in list(sys.modules.values()):
        name = getattr(mod, "__name__", "")
        if mod is None or name.startswith("tensorized_rnn_b200") or not hasattr(mod, "__dict__"):
            continue
        for k, v in list(vars(mod).items()):
            new = originals.get(id(v))
            if new is not None and v is not new:
                _saved.setdefault((name, k), v)
                setattr(mod, k, new)
                n += 1
    return n


def unpatch_reference() -> None:
    for (mod_name, attr), old in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, attr, old)
    _saved.clear()
