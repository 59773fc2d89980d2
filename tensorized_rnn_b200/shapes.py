"""TT mode-size selection: which factorisation of (in, gates*hidden) the cores use.

Behavioural contract (checked against a table generated from the reference,
tests/golden/shapes.json): identical results to `t3nsor.utils.auto_shape`
(reference t3nsor/utils.py:39-81, criterion 'entropy', mode 'ascending') and
`tensorized_rnn.rnn_utils.tt_shape` (reference tensorized_rnn/rnn_utils.py:20-36),
because the factorisation fixes every core's shape and therefore whether a
reference checkpoint loads.

The implementation is independent of sympy/scipy: candidates are enumerated
directly as non-decreasing d-tuples of integers >= 2 whose product is n, which
is what "all multiset partitions of the prime factors into d blocks" amounts to.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List, Optional, Tuple


def _prime_factors(n: int) -> List[int]:
    out, p = [], 2
    while p * p <= n:
        while n % p == 0:
            out.append(p)
            n //= p
        p += 1 if p == 2 else 2
    if n > 1:
        out.append(n)
    return out


def _ascending_factorisations(n: int, d: int, lo: int = 2):
    """Yield every non-decreasing d-tuple of integers >= lo with product n."""
    if d == 1:
        if n >= lo:
            yield (n,)
        return
    f = lo
    while f ** d <= n:
        if n % f == 0:
            for rest in _ascending_factorisations(n // f, d - 1, f):
                yield (f,) + rest
        f += 1


def _entropy(factors: Tuple[int, ...]) -> float:
    total = float(sum(factors))
    return -sum((f / total) * math.log(f / total) for f in factors)


@lru_cache(maxsize=None)
def _auto_shape_cached(n: int, d: int) -> Tuple[int, ...]:
    primes = _prime_factors(n)
    if len(primes) <= d:
        # fewer prime factors than cores: the only partition is one prime (or a 1) per core
        return tuple(sorted(primes + [1] * (d - len(primes))))
    best, best_score = None, -1.0
    for cand in _ascending_factorisations(n, d):
        s = _entropy(cand)
        if s > best_score + 1e-12:
            best, best_score = cand, s
    assert best is not None
    return best


def auto_shape(n: int, d: int = 3, criterion: str = "entropy", mode: str = "ascending") -> List[int]:
    """The d-factorisation of n, ascending, whose factor tuple has maximal entropy."""
    if criterion != "entropy" or mode != "ascending":
        raise NotImplementedError("only criterion='entropy', mode='ascending' (the reference's defaults, "
                                  "the only ones its recurrent modules use) are supported")
    if n < 1 or d < 1:
        raise ValueError("auto_shape needs n >= 1 and d >= 1")
    return list(_auto_shape_cached(int(n), int(d)))


def tt_shape(in_features: int, out_features: int, n_cores: int, n_gates: int,
             new_core: Optional[str] = None) -> List[List[int]]:
    """[in_quant, out_quant] for the concat-gates TT matrix of one RNN weight.

    Without `new_core` the gate count is folded into the output dimension before
    factorising; 'first' / 'last' keep `out_features` and add a (1 x n_gates) core.
    """
    assert new_core in [None, "first", "last"]
    if new_core is None:
        out_features = out_features * n_gates
    in_quant = auto_shape(in_features, d=n_cores)
    out_quant = auto_shape(out_features, d=n_cores)
    if new_core == "first":
        in_quant, out_quant = [1] + in_quant, [n_gates] + out_quant
    elif new_core == "last":
        in_quant, out_quant = in_quant + [1], out_quant + [n_gates]
    return [in_quant, out_quant]
