"""TEST INFRASTRUCTURE: what the Python host layer (detect_orfs.py, bam.py, metagene.py) asks of an ``Engine``, answered by the
C oracle on host arrays.  It lets ``-m "not gpu"`` tests drive the WHOLE host flow of ``detect_orfs()`` -- index parsing,
protocol inference, the metagene / offset step, the row selection, the native TSV / WIG / summary writers -- against the files
the unmodified reference wrote, without a GPU.  The kernels themselves are not involved and nothing here ships: the product
has no CPU path (``tests/test_abi.py::test_no_cpu_fallback``); the GPU twins of these tests are in ``tests/test_gpu_api.py``.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import c_oracle as CO
from ribotricer_b200 import _lib
from ribotricer_b200.const import DEFAULT_PAD
from ribotricer_b200.engine import ScoreParams, protocol_code


class OracleEngine:
    def __init__(self, device: int = 0):
        self.torch, self.device, self.device_index, self.ctx = torch, "cpu", device, object()
        self.lib = _lib.load()                      # the host-only entry points (rt_tsv_*, rt_wig_*, rt_metagene_sums, ...)
        self.contig_names, self.contig_len, self.pad = [], np.zeros(0, np.int64), DEFAULT_PAD
        self.contig_base, self.plane = np.zeros(0, np.int64), 0
        self.layout, self.index, self.n_orf, self.len_table = "dense", None, 0, None
        self._aux = None
        self.calls: list = []                       # names of the "device" calls, for the tests to look at

    # ---- genome / layout
    def set_genome(self, contig_names, contig_len, pad: int = DEFAULT_PAD):
        self.contig_names, self.contig_len, self.pad = list(contig_names), np.asarray(contig_len, np.int64), int(pad)
        self.contig_base, self.plane = CO.genome_layout(self.contig_len, self.pad)
        self.index, self.n_orf, self.layout = None, 0, "dense"
        self.__dict__.pop("_resident_index", None)

    def contig_id(self, name: str) -> int:
        return self.contig_names.index(name) if name in self.contig_names else -1

    def ensure_dense(self):
        self.layout = "dense"

    def set_layout(self, layout: str):
        assert layout in ("dense", "compact")
        self.layout = layout                        # the twin keeps the genome-wide planes either way

    def coverage_elems(self) -> int:
        return 2 * self.plane + (64 if self.layout == "compact" else 0)

    def new_coverage(self):
        return torch.zeros(self.coverage_elems(), dtype=torch.int32)

    def compact_from_dense(self, dense_cov, compact_cov):
        self.calls.append("compact_from_dense")
        compact_cov[:2 * self.plane] = dense_cov[:2 * self.plane]

    # ---- reads
    READ_COLUMNS = ("ref_id", "first", "last", "mlen", "flag", "mapq", "nh")

    def stream_reads(self, cols: dict, pinned: bool = False, n_threads: int = 0) -> dict:
        return {k: np.ascontiguousarray(cols[k]) for k in self.READ_COLUMNS}

    def upload_stream(self, stream: dict) -> dict:
        return stream

    def upload_reads(self, cols: dict) -> dict:
        return {k: np.ascontiguousarray(cols[k]) for k in self.READ_COLUMNS}

    def set_length_table(self, psite_offsets=None, read_lengths=None, table=None):
        self.len_table = CO.make_len_table(psite_offsets, read_lengths) if table is None else table

    def new_bin_accumulators(self):
        return torch.zeros(_lib.RT_N_STATS, dtype=torch.int64), torch.zeros(_lib.RT_LEN_TABLE, dtype=torch.int64)

    def _bin(self, cov, cols, protocol, stats, len_counts, weight):
        self.calls.append("bin")
        scratch = np.zeros(2 * self.plane, np.int32)
        _, st, lc = CO.bin_reads(cols, protocol_code(protocol), self.len_table, self.contig_base, self.contig_len, self.pad,
                                 self.plane, cov=scratch)
        if cov.numel() >= 2 * self.plane:           # Alignments.count() passes a one-element buffer: nothing is stored
            cov[:2 * self.plane] += int(weight) * torch.from_numpy(scratch)
        stats += torch.tensor([st[k] for k in _lib.ST_NAMES], dtype=torch.int64)
        len_counts += torch.from_numpy(lc)

    def bin_stream_device(self, cov, dstream, protocol, stats, len_counts, weight: int = 1, fresh: bool = False):
        if fresh:
            cov.zero_()
        self._bin(cov, dstream, protocol, stats, len_counts, weight)

    def bin_reads_device(self, cov, dcols, protocol, stats, len_counts, sorted_hint: bool = False, n=None, weight: int = 1):
        self._bin(cov, dcols, protocol, stats, len_counts, weight)

    # ---- index, scoring, profiles
    def set_index(self, exon_ptr, exon_start, exon_end, orf_contig, orf_strand):
        self.index = dict(exon_ptr=np.ascontiguousarray(exon_ptr, np.int64), exon_start=np.ascontiguousarray(exon_start, np.int32),
                          exon_end=np.ascontiguousarray(exon_end, np.int32), orf_contig=np.ascontiguousarray(orf_contig, np.int32),
                          orf_strand=np.ascontiguousarray(orf_strand, np.uint8))
        self.n_orf = len(self.index["orf_contig"])

    def _planes(self, cov):
        return np.ascontiguousarray(cov[:2 * self.plane].numpy())

    def score_host(self, cov, lo: int = 0, hi=None, params: ScoreParams | None = None, diagnostics: bool = False,
                   min_codon=None, out=None) -> dict:
        self.calls.append("score_host")
        p = params or ScoreParams()
        hi = self.n_orf if hi is None else hi
        return CO.score(self.index, self._planes(cov), self.contig_base, self.contig_len, self.pad, self.plane,
                        [p.phase_score_cutoff, p.min_valid_codons, p.min_reads_per_codon, p.min_valid_codons_ratio,
                         p.min_density_over_orf], lo, hi, diagnostics=diagnostics)

    def gather_profiles(self, cov, orf_ids, lengths):
        self.calls.append("gather_profiles")
        ptr, flat = CO.gather_profiles(self.index, np.asarray(orf_ids, np.int64), self._planes(cov), self.contig_base,
                                       self.contig_len, self.pad, self.plane)
        assert (np.diff(ptr) == np.asarray(lengths, np.int64)).all()
        return ptr, flat

    def nonzero_slots(self, cov, lo: int, n: int):
        self.calls.append("nonzero_slots")
        part = cov[lo:lo + n].numpy()
        slots = np.flatnonzero(part).astype(np.int64)
        return slots, part[slots].astype(np.int32)


def install(monkeypatch) -> OracleEngine:
    """An OracleEngine (with its auxiliary twin) as the process-wide engine of ribotricer_b200.detect_orfs, and the
    oracle's SciPy phasescore for float profiles."""
    from oracle import oracle_py as O
    from ribotricer_b200 import detect_orfs as D
    from ribotricer_b200 import statistics as S

    eng = OracleEngine()
    eng._aux = OracleEngine()
    monkeypatch.setattr(D, "_ENGINE", eng)
    monkeypatch.setattr(D, "_INDEX_CACHE", {})
    monkeypatch.setattr(S, "phasescore", lambda values, engine=None: O.phasescore_scipy(list(values)))
    return eng
