"""B200-native implementation of ribotricer's detect-orfs scoring path.

Drop-in for ``ribotricer.detect_orfs`` (detect_orfs.py:354), its inner seams
(``merge_read_lengths``, ``orf_coverage``, ``export_orf_coverages``,
``phasescore``) and the ``ribotricer detect-orfs`` CLI.  All arithmetic on the
path runs in hand-written sm_100a CUDA kernels behind a C ABI
(``include/ribotricer_b200.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"
# version of the reference whose detect-orfs behaviour is mirrored
__reference_version__ = "1.5.0"
