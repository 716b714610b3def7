"""IUPAC ambiguity with the reference's API (TrueConsense/Ambig.py).

``IsAmbiguous(one, two, three, four, cov)`` takes the four top-ranked ``(letter, count)`` tuples
and the coverage and returns ``(True, char)`` or ``(False, None)`` (Ambig.py:179-228).  The
decision is taken by the device function the call kernel uses for whole tables
(csrc/cuda/call.cu ``is_ambiguous``: IEEE-double ``(count / cov) * 100`` and ``abs(pi - pj) <= 10``),
reached through ``tc_is_ambiguous``.
"""
from __future__ import annotations

import numpy as np

from . import gpu


def GetPercentages(c1, c2, c3, c4, cov):
    """Ambig.py:102-127."""
    return (c1 / cov) * 100, (c2 / cov) * 100, (c3 / cov) * 100, (c4 / cov) * 100


def IsAmbiguousBatch(letters: np.ndarray, counts: np.ndarray, cov: np.ndarray, maxdist: float = 10.0) -> np.ndarray:
    """Vector form: letters uint8[4][n] (ASCII), counts int32[4][n], cov int32[n] -> uint8[n]
    ambiguity characters (0 = not ambiguous)."""
    return gpu.default_context().is_ambiguous(letters, counts, cov, maxdist)


def IsAmbiguous(one, two, three, four, cov):
    """Ambig.py:179-228."""
    if cov == 0:
        return False, None
    letters = np.array([[ord(t[0])] for t in (one, two, three, four)], dtype=np.uint8)
    counts = np.array([[t[1]] for t in (one, two, three, four)], dtype=np.int32)
    ch = int(IsAmbiguousBatch(letters, counts, np.array([cov], dtype=np.int32))[0])
    if ch == 0:
        return False, None
    return True, chr(ch)
