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


def frame_tie_mask(frame_K, frame_s, tol: float = 1e-12):
    """Rows whose ``valid_codons`` the reference decides by rounding noise (SURVEY.md hazard H1).

    ``statistics.py:109`` keeps a running maximum over the three frames with a strict ``>``.  When two frames reach
    the same coherence in exact arithmetic (sparse, perfectly periodic ORFs: both exactly 1, or both exactly 0) but
    hold different numbers of codons, SciPy's last-bit noise picks the winner -- and with it ``valid_codons``,
    ``valid_codons_ratio`` and, around ``min_valid_codons``, the status.  This package resolves such ties
    deterministically (sums of unit vectors are exact integers on a 2^-42 grid; the earliest frame wins), so for
    these rows -- 0.2 to 1 % of the ORFs of a sparse library -- the TSV may differ from ribotricer's in those three
    columns and nowhere else.  Pass the ``frame_K`` / ``frame_s`` columns of ``Engine.score_host(...,
    diagnostics=True)``; returns a boolean array marking the rows.
    """
    frame_K = np.asarray(frame_K)
    frame_s = np.asarray(frame_s)
    n = len(frame_K)
    best = np.zeros(n)
    valid = np.full(n, -1, np.int64)
    tie = np.zeros(n, bool)
    for f in range(3):
        k = frame_K[:, f].astype(np.int64)
        s = frame_s[:, f]
        zero = k == 0
        ok = ~zero & ~np.isnan(s)
        with np.errstate(invalid="ignore"):
            win = ok & (s > best + tol)
            close = ok & ~win & (np.abs(s - best) <= tol)
        tie = np.where(zero | win, False, tie | (close & (valid != -1) & (valid != k)))
        best = np.where(zero, 0.0, np.where(win | close, np.fmax(best, np.where(ok, s, best)), best))
        valid = np.where(zero, 0, np.where(win, k, valid))
        valid = np.where(~zero & (valid == -1), k, valid)
    return tie
