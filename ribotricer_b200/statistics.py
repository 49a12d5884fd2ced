"""``phasescore`` on the GPU (mirror of ribotricer/statistics.py:48-115).

``phasescore(values)`` accepts any sequence of ints or floats, like the reference (which the
metagene step calls on float profiles, metagene.py:243-244).  The arithmetic runs in
``phasescore_values_kernel`` through ``rt_phasescore_values``; bulk per-ORF scoring uses the fused
gather+score kernels instead (``Engine.score_host``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def phasescore(original_values, engine=None):
    """Return ``(np.float64 score, int valid_codons)`` -- statistics.py:48-115.

    The score is sqrt of the largest magnitude-squared coherence (at f = 1/3) over the three
    frames between the per-codon normalised signal and the ideal 1-0-0 signal; ``valid_codons``
    is the number of non-all-zero codons of the winning frame.
    """
    from .detect_orfs import get_engine

    eng = engine or get_engine()
    vals = np.ascontiguousarray(list(original_values) if not isinstance(original_values, np.ndarray)
                                else original_values, dtype=np.float64)
    score = C.c_double()
    valid = C.c_int32()
    _lib.check(eng.lib.rt_phasescore_values(eng.ctx, vals.ctypes.data_as(C.c_void_p), len(vals),
                                            C.byref(score), C.byref(valid)), eng.ctx)
    return np.float64(score.value), int(valid.value)
