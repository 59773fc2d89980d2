"""Row-group planner (csrc/ttrnn_capi.cu, plan_row_groups) on the host: no device is touched when the SM count is
given.  The planner is host logic of the path's launch sequence; the reference has no counterpart (lstm.py:123-133 walks
the whole batch layer by layer)."""
import ctypes
import io
from contextlib import redirect_stdout

import torch

import tensorized_rnn_b200 as tr
from tensorized_rnn_b200 import _lib


def groups(cls, I, H, L, d, r, B, T, sms=148):
    with redirect_stdout(io.StringIO()):
        m = cls(I, H, L, torch.device("cpu"), n_cores=d, tt_rank=r)
    rows = (ctypes.c_int64 * 2)()
    n = _lib.load().ttrnn_rnn_row_groups(ctypes.byref(m.spec().desc(B, T)), sms, rows)
    assert n >= 1, _lib.last_error()
    return n, rows[0], rows[1]


def test_cfg3_batch_is_cut_into_a_full_wave_and_a_remainder():
    # 640 rows on 148 SMs: 148 three-row CTAs + 98 two-row CTAs instead of 128 five-row CTAs (forward) /
    # 148 + 98 CTAs back to back (backward)
    assert groups(tr.TTLSTM, 40, 256, 3, 3, 8, 640, 160) == (2, 444, 196)
    n, r0, r1 = groups(tr.TTLSTM, 40, 256, 3, 3, 8, 600, 160)
    assert n == 2 and r0 % 148 == 0 and r0 + r1 == 600 and r0 % 4 == 0


def test_no_split_where_it_cannot_help():
    assert groups(tr.TTLSTM, 1, 256, 1, 2, 4, 256, 784)[0] == 1          # one layer: nothing to pipeline
    assert groups(tr.TTLSTM, 40, 256, 3, 3, 8, 80, 160)[0] == 1          # fewer rows than SMs
    assert groups(tr.TTLSTM, 40, 256, 3, 3, 8, 444, 160)[0] == 1         # already a whole wave
    assert groups(tr.TTLSTM, 40, 256, 3, 4, 16, 16384, 160)[0] == 1      # time-chunked / many waves per launch
    assert groups(tr.TTLSTM, 12, 60, 2, 3, 3, 640, 20)[0] == 1           # no static kernel for this shape


def test_option_switches():
    lib = _lib.load()
    try:
        lib.ttrnn_set_option(b"row_groups", 0)
        assert groups(tr.TTLSTM, 40, 256, 3, 3, 8, 640, 160) == (1, 640, 0)
        lib.ttrnn_set_option(b"row_groups", 26)                          # forced: rounded down to a multiple of 4
        assert groups(tr.TTGRU, 1, 256, 1, 2, 4, 40, 30) == (2, 24, 16)
        assert groups(tr.TTGRU, 1, 256, 1, 2, 4, 20, 30) == (1, 20, 0)   # not enough rows for the forced split
    finally:
        lib.ttrnn_set_option(b"row_groups", 1)


def test_planner_invariants_over_batches_and_sm_counts():
    """Whatever the planner decides: one or two groups, row counts that add up to the batch, and a cut at a multiple of
    4 rows (row-indexed slices of every buffer stay 16-byte aligned)."""
    for sms in (148, 132, 64):
        for B in (8, 60, 148, 150, 200, 296, 300, 320, 444, 480, 512, 600, 640, 700, 1000, 1500, 2368, 2400):
            n, r0, r1 = groups(tr.TTLSTM, 40, 256, 3, 3, 8, B, 160, sms=sms)
            assert n in (1, 2)
            if n == 1:
                assert (r0, r1) == (B, 0)
            else:
                assert r0 + r1 == B and r0 % 4 == 0 and r0 >= 4 and r1 >= 1, (sms, B, r0, r1)
    # the decision is cached per (descriptor, options, SM count): asking again gives the same answer
    assert groups(tr.TTLSTM, 40, 256, 3, 3, 8, 640, 160) == groups(tr.TTLSTM, 40, 256, 3, 3, 8, 640, 160)
