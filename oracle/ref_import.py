"""Import the UNMODIFIED reference (ribotricer 1.5.0) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Only usable in the authoring container, where
``/root/reference`` exists; it is used by ``tests/golden/make_golden.py`` to
generate the committed golden vectors and by a few optional ``not gpu`` tests
(skipped when the tree is absent).  Nothing on the product path imports this.

The reference's I/O dependencies (pysam, quicksect, matplotlib, pyfaidx,
click_help_colors) are absent from this image; none of them carries hot-path
arithmetic.  pysam is replaced by ``oracle/pysam_restated.py`` (a pure-Python
restatement of the few calls ``split_bam`` makes, so that bam.py:33-153 runs
unmodified on real BAM bytes); the others by empty ``types.ModuleType`` stubs
(recipe: SURVEY.md 8(c)).  After that ``ribotricer.detect_orfs`` imports
unmodified and ``merge_read_lengths`` (detect_orfs.py:54), ``orf_coverage``
(:134), ``export_orf_coverages`` (:206), ``export_wig`` (:327) and
``statistics.phasescore`` (statistics.py:48) run as shipped.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RIBOTRICER_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ribotricer"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    for key, val in attrs.items():
        setattr(mod, key, val)
    sys.modules[name] = mod
    return mod


def load():
    """Return the reference ``ribotricer`` package (with I/O deps stubbed)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the tree is read-only
    try:
        import pysam  # noqa: F401
    except ImportError:
        # restated slice of pysam (see oracle/pysam_restated.py): lets the unmodified split_bam run here
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import pysam_restated

        sys.modules["pysam"] = pysam_restated
    try:
        import quicksect  # noqa: F401
    except ImportError:
        class _Interval:  # minimal stand-ins; never exercised on the scoring path
            def __init__(self, start, end, data=None):
                self.start, self.end, self.data = start, end, data

        class _IntervalTree:
            def __init__(self):
                self.items = []

            def insert(self, iv):
                self.items.append(iv)

            def find(self, iv):
                return [x for x in self.items if x.start <= iv.end and iv.start <= x.end]

        _stub("quicksect", Interval=_Interval, IntervalTree=_IntervalTree)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = _stub("matplotlib", use=lambda *a, **k: None, rcParams={})
        plt = _stub("matplotlib.pyplot")
        backends = _stub("matplotlib.backends")
        pdf = _stub("matplotlib.backends.backend_pdf", PdfPages=object)
        mpl.pyplot, mpl.backends, backends.backend_pdf = plt, backends, pdf
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import ribotricer  # noqa: E402
    import ribotricer.detect_orfs  # noqa: E402,F401
    import ribotricer.statistics  # noqa: E402,F401
    return ribotricer
