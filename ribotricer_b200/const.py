"""Defaults of the detect-orfs path (mirror of ribotricer/const.py:20-42)."""
from __future__ import annotations

from typing import Final

# ribotricer/const.py:20
CUTOFF: Final[float] = 0.428571428571
# ribotricer/const.py:23
TYPICAL_OFFSET: Final[int] = 12
# ribotricer/const.py:27
MINIMUM_VALID_CODONS: Final[int] = 5
# ribotricer/const.py:32
MINIMUM_READS_PER_CODON: Final[int] = 0
# ribotricer/const.py:35
MINIMUM_VALID_CODONS_RATIO: Final[float] = 0
# ribotricer/const.py:39
MINIMUM_DENSITY_OVER_ORF: Final[float] = 0.0
# ribotricer/const.py:42
META_MIN_READS: Final[int] = 100000

# Slack slots kept on both sides of every contig in the dense coverage planes, so that a
# P-site shifted off a contig end keeps a private slot.  Must be >= the
# largest P-site offset; offsets are < read length (cli.py:268), so 256 covers every
# Ribo-seq and short-read RNA-seq library.
DEFAULT_PAD: Final[int] = 256
