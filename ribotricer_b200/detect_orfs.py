"""``detect-orfs`` on the GPU -- same function surface as ribotricer/detect_orfs.py.

    merge_read_lengths   detect_orfs.py:54     -> K1 (bin_stream_kernel on a coordinate-sorted library, else bin_psites_kernel)
    orf_coverage         detect_orfs.py:134    -> K4 (gather_profiles_kernel)
    export_orf_coverages detect_orfs.py:206    -> K2+K3 (atom_pass_kernel + compose_refs_kernel; score_orfs_kernel only for
                                                  the ORFs the plan hands to the fallback launch) + K4 + native TSV writer
    export_wig           detect_orfs.py:327    -> wig_count_kernel + wig_fill_kernel + native WIG writer
    detect_orfs          detect_orfs.py:354    (same positional signature; learn_cutoff.py:231 calls it)

The dict-of-Counter values of the reference become device-resident objects:
``Alignments`` (read columns in HBM) and ``MergedAlignments`` (dense P-site
coverage planes in HBM).
"""
from __future__ import annotations

import datetime
import os
import pathlib
import sys
from collections import Counter, defaultdict

import numpy as np

from .bam import Alignments, split_bam
from .const import (CUTOFF, MINIMUM_DENSITY_OVER_ORF, MINIMUM_READS_PER_CODON, MINIMUM_VALID_CODONS,
                    MINIMUM_VALID_CODONS_RATIO)
from .engine import Engine, ScoreParams
from .index import ORF, NativeIndex, PackedIndex, parse_index

_ENGINE: Engine | None = None
_INDEX_CACHE: dict = {}

TSV_COLUMNS = [
    "ORF_ID", "ORF_type", "status", "phase_score", "read_count", "length", "valid_codons",
    "valid_codons_ratio", "read_density", "transcript_id", "transcript_type", "gene_id", "gene_name",
    "gene_type", "chrom", "strand", "start_codon", "profile",
]   # detect_orfs.py:241-260


def get_engine(device: int | None = None) -> Engine:
    """Process-wide engine (one rt_ctx per process and GPU)."""
    global _ENGINE
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _ENGINE is None or _ENGINE.device_index != device or _ENGINE.ctx is None:
        _ENGINE = Engine(device)
    return _ENGINE


def load_index(path: str) -> PackedIndex:
    """Parse (once per process and file version) the candidate-ORF index."""
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
    if key not in _INDEX_CACHE:
        _INDEX_CACHE.clear()
        _INDEX_CACHE[key] = NativeIndex(path)     # native loader; parse_index() is the Python statement of the rules
    return _INDEX_CACHE[key]


class MergedAlignments:
    """P-site coverage of one library after offset shifting and merging over read lengths
    (what merge_read_lengths returns, detect_orfs.py:49,72-83), as dense planes in HBM."""

    def __init__(self, engine: Engine, cov):
        self.engine = engine
        self.cov = cov

    def nonzero(self):
        """(strand_code, contig_id, pos, count) arrays of every covered position, ordered by
        strand, contig and position: one ordered compaction of each strand plane on the device
        (``Engine.nonzero_slots``), split by contig on the host.  Dense layout only."""
        eng = self.engine
        out = []
        base = np.asarray(eng.contig_base, np.int64)
        for strand in (0, 1):
            slots, vals = eng.nonzero_slots(self.cov, strand * eng.plane, eng.plane)
            if len(slots) == 0:
                continue
            c = np.searchsorted(base, slots, side="right") - 1          # contig of every covered slot
            out.append((np.full(len(slots), strand, np.int8), c.astype(np.int32), slots - base[c] - eng.pad, vals))
        if not out:
            z = np.zeros(0, np.int64)
            return z.astype(np.int8), z.astype(np.int32), z, z.astype(np.int32)
        return tuple(np.concatenate(x) for x in zip(*out))

    def covered_by_contig(self):
        """Yields ``(strand_code, contig_id, positions int64, counts int32)`` for every contig that carries coverage on
        a strand, positions ascending: the ordered compaction of a strand plane cut at the contig bases (no per-position
        host work: a human library has 10^8 covered positions)."""
        eng = self.engine
        base = np.asarray(eng.contig_base, np.int64)
        for strand in (0, 1):
            slots, vals = eng.nonzero_slots(self.cov, strand * eng.plane, eng.plane)
            if len(slots) == 0:
                continue
            cuts = np.searchsorted(slots, np.append(base, np.int64(eng.plane)))
            for ci in range(len(base)):
                a, b = int(cuts[ci]), int(cuts[ci + 1])
                if a < b:
                    yield strand, ci, slots[a:b] - (base[ci] + eng.pad), vals[a:b]

    def to_dict(self):
        """The reference's ``merged[strand][(chrom, pos)] -> count`` view (small libraries only)."""
        merged = defaultdict(Counter)
        s, c, p, n = self.nonzero()
        names = self.engine.contig_names
        for i in range(len(s)):
            merged["+" if s[i] == 0 else "-"][(names[c[i]], int(p[i]))] = int(n[i])
        return merged


def merge_read_lengths(alignments: Alignments, psite_offsets: dict) -> MergedAlignments:
    """detect_orfs.py:54-83: shift every read of a length in ``psite_offsets`` by its offset
    ('+': pos + offset, '-': pos - offset) and sum over lengths -- one K1 launch."""
    eng = alignments.engine
    eng.ensure_dense()                   # a MergedAlignments holds the genome-wide planes
    cov = eng.new_coverage()
    alignments.bin_into(cov, psite_offsets)
    return MergedAlignments(eng, cov)


def parse_ribotricer_index(ribotricer_index: str):
    """detect_orfs.py:86-131: the leading 'annotated' rows as ORF objects plus, per chromosome,
    the (start, end, strand) spans that infer_protocol needs."""
    idx = load_index(ribotricer_index)
    n = idx.n_annotated_prefix
    annotated, refseq = [], defaultdict(list)
    if isinstance(idx, NativeIndex) and n:
        # the prefix rows are the first n lines behind the header: their text fields are split here in one pass (a
        # human index has 60 k annotated rows; one native call per field costs more than the split), the intervals come
        # from the loader's columns (already sorted by start, orf.py:100)
        with open(ribotricer_index, encoding="utf-8", newline="\n") as fh:
            fh.readline()
            rows = [fh.readline().rstrip("\n").split("\t") for _ in range(n)]
        ptr = idx.exon_ptr[:n + 1].tolist()
        starts, ends = idx.exon_start[:ptr[-1]].tolist(), idx.exon_end[:ptr[-1]].tolist()
        import gc
        gc_was_on = gc.isenabled()
        gc.disable()        # hundreds of thousands of small objects are made here and none is garbage: the collector's
        try:                # generation scans would take more time than the loop itself
            for o, f in enumerate(rows):
                if f[1] != "annotated":       # detect_orfs.py:120
                    continue
                ivs = list(zip(starts[ptr[o]:ptr[o + 1]], ends[ptr[o]:ptr[o + 1]]))
                orf = ORF(f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], ivs, f[9])
                orf.row = o
                refseq[orf.chrom].append((ivs[0][0], ivs[-1][1], 1 if orf.strand == "+" else -1))
                annotated.append(orf)
        finally:
            if gc_was_on:
                gc.enable()
        return annotated, refseq
    for o in range(n):
        f = idx.fields[o]
        if f[0] != "annotated":       # detect_orfs.py:120
            continue
        a, b = idx.exon_ptr[o], idx.exon_ptr[o + 1]
        ivs = list(zip(idx.exon_start[a:b].tolist(), idx.exon_end[a:b].tolist()))
        orf = ORF(f[0], f[1], f[2], f[3], f[4], f[5], idx.chrom[o], idx.strand[o], ivs, f[6])
        orf.row = o
        refseq[orf.chrom].append((ivs[0][0], ivs[-1][1], 1 if orf.strand == "+" else -1))
        annotated.append(orf)
    return annotated, refseq


def _aux_engine(eng: Engine) -> Engine:
    """Second ctx on the same GPU for small auxiliary indexes (single-ORF queries, metagene
    windows); shares the caller-owned coverage buffer, leaves the resident index alone."""
    aux = getattr(eng, "_aux", None)
    if aux is None or aux.ctx is None:
        aux = Engine(eng.device_index)
        eng._aux = aux
    if aux.plane != eng.plane or aux.pad != eng.pad or list(aux.contig_names) != list(eng.contig_names):
        aux.set_genome(eng.contig_names, eng.contig_len, eng.pad)
    return aux


def orf_coverage(orf: ORF, alignments: MergedAlignments, offset_5p: int = 0, offset_3p: int = 0) -> list:
    """detect_orfs.py:134-203 for one ORF: coverage over the leader, every interval and the
    trailer, reversed on the '-' strand."""
    eng = alignments.engine
    if orf.strand == "-":
        offset_5p, offset_3p = offset_3p, offset_5p
    ivs = [list(iv) for iv in orf.intervals]
    ivs[0][0] -= offset_5p
    ivs[-1][1] += offset_3p
    aux = _aux_engine(eng)
    aux.set_index(np.array([0, len(ivs)], np.int64), np.array([iv[0] for iv in ivs], np.int32),
                  np.array([iv[1] for iv in ivs], np.int32),
                  np.array([eng.contig_id(orf.chrom)], np.int32),
                  np.array([0 if orf.strand == "+" else 1 if orf.strand == "-" else 2], np.uint8))
    length = sum(iv[1] - iv[0] + 1 for iv in ivs)
    _, prof = aux.gather_profiles(alignments.cov, np.array([0], np.int64), np.array([length], np.int64))
    return prof.tolist()


def export_orf_coverages(
    ribotricer_index: str,
    merged_alignments: MergedAlignments,
    prefix: str,
    phase_score_cutoff: float = CUTOFF,
    min_valid_codons: int = MINIMUM_VALID_CODONS,
    min_reads_per_codon: float = MINIMUM_READS_PER_CODON,
    min_valid_codons_ratio: float = MINIMUM_VALID_CODONS_RATIO,
    min_density_over_orf: float = MINIMUM_DENSITY_OVER_ORF,
    report_all: bool = False,
    orf_range: tuple | None = None,
    write_header: bool = True,
    path: str | None = None,
) -> dict:
    """detect_orfs.py:206-324: score every ORF of the index in index order and write
    ``{prefix}_translating_ORFs.tsv`` (non-translating rows only with ``report_all``).

    ``orf_range`` / ``write_header`` / ``path`` are extensions used by the multi-GPU driver
    (each rank writes the rows of its ORF shard).  Returns the per-ORF result columns.
    """
    eng = merged_alignments.engine
    idx = load_index(ribotricer_index)
    if getattr(eng, "_resident_index", None) is not idx:
        lut = {n: i for i, n in enumerate(eng.contig_names)}
        eng.set_index(**idx.device_columns(lut))
        eng._resident_index = idx
    lo, hi = orf_range if orf_range is not None else (0, idx.n_orf)
    params = ScoreParams(phase_score_cutoff, min_valid_codons, min_reads_per_codon, min_valid_codons_ratio,
                         min_density_over_orf)
    merged = merged_alignments
    dense_in = eng.layout == "dense" and idx.n_orf > 0
    if dense_in:
        # the library arrives in the genome-wide planes (export_wig and the metagene step need those); scoring and
        # the profile gather only read the exon union of the index: copy that part into a compact-layout buffer and
        # run the compact-layout kernels on it (same results, see tests/test_gpu_parity.py)
        eng.set_layout("compact")
        ccov = eng.torch.empty(eng.coverage_elems(), dtype=eng.torch.int32, device=eng.device)
        ccov[-64:] = 0                                              # the guard slots behind the last atom
        eng.compact_from_dense(merged_alignments.cov, ccov)
        merged = MergedAlignments(eng, ccov)
    try:
        res = eng.score_host(merged.cov, lo, hi, params)
        write_tsv(path or f"{prefix}_translating_ORFs.tsv", idx, res, merged, lo, hi, report_all, write_header)
    finally:
        if dense_in:
            eng.set_layout("dense")
    return res


def write_tsv(path, idx: PackedIndex, res: dict, merged: MergedAlignments, lo: int, hi: int, report_all: bool,
              write_header: bool = True, chunk_nt: int = 1 << 25, res_offset: int = 0):
    """Rows exactly as detect_orfs.py:304-323 formats them: np.float64 phase score and read
    density, Python-float ratio, ``str(list)`` profile.  Rows [lo, hi) of the index; their results sit at
    positions ``res_offset ...`` of the result columns, which is also their ORF number in the engine's resident
    index (0 unless the engine holds a sub-index, see multi_gpu.py)."""
    eng = merged.engine
    n = hi - lo
    view = {k: v[res_offset:res_offset + n] for k, v in res.items()}
    keep = np.arange(lo, hi) if report_all else lo + np.flatnonzero(view["status"])
    length = view["length"].astype(np.int64)
    n_codons = np.maximum(1, length // 3)                         # detect_orfs.py:281
    ratio = view["valid"].astype(np.float64) / n_codons          # :285
    density = view["count"].astype(np.float64) / n_codons        # :287
    native = isinstance(idx, NativeIndex)
    if native:
        import ctypes as C

        lib = eng.lib
        handle = C.c_void_p()
        if lib.rt_tsv_open(str(path).encode(), int(write_header), C.byref(handle)) != 0:
            raise OSError(lib.rt_io_last_error().decode())
        cols = {k: np.ascontiguousarray(view[k]) for k in ("score", "valid", "count", "length", "status")}
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        # the rows of a chunk are formatted (rt_tsv_write: all host cores, the GIL released) on a helper thread while this
        # thread gathers the profiles of the next chunk on the GPU; one chunk in flight, rows stay in order
        from concurrent.futures import ThreadPoolExecutor

        formatter, in_flight = ThreadPoolExecutor(1), None

        def emit(sel, ptr, prof):
            rc = lib.rt_tsv_write(handle, idx.handle, len(sel), p(sel), int(lo), p(cols["score"]), p(cols["valid"]),
                                  p(cols["count"]), p(cols["length"]), p(cols["status"]), p(ptr), p(prof))
            if rc != 0:       # the message is thread-local in the library: read it on this thread
                raise OSError(f"rt_tsv_write failed ({rc}): {lib.rt_io_last_error().decode()}")
    else:
        out = open(path, "w")
        if write_header:
            out.write("\t".join(TSV_COLUMNS) + "\n")
    try:
        at = 0
        csum = np.cumsum(length[keep - lo])          # profile values up to and including every reported row
        while at < len(keep):
            # bounded chunks of reported ORFs (128 MB of profile values): the buffers stay small and the gather of one chunk
            # hides behind the text of the previous one
            before = int(csum[at - 1]) if at else 0
            n_take = max(1, int(np.searchsorted(csum, before + chunk_nt, side="right")) - at)
            sel = np.ascontiguousarray(keep[at:at + n_take], np.int64)
            ptr, prof = eng.gather_profiles(merged.cov, sel - lo + res_offset, length[sel - lo])
            if native:
                if in_flight is not None:
                    in_flight.result()
                in_flight = formatter.submit(emit, sel, np.ascontiguousarray(ptr, np.int64), np.ascontiguousarray(prof, np.int32))
            else:
                rows = []
                for j, o in enumerate(sel.tolist()):
                    k = o - lo
                    f = idx.fields[o]
                    rows.append("\t".join((
                        idx.oid(o), f[0], "translating" if view["status"][k] else "nontranslating",
                        str(view["score"][k]), str(int(view["count"][k])), str(int(length[k])),
                        str(int(view["valid"][k])), repr(float(ratio[k])), str(density[k]),
                        f[1], f[2], f[3], f[4], f[5], idx.chrom[o], idx.strand[o],
                        f[6][:3] if len(f[6]) >= 3 else "None",          # ORF.start_codon, orf.py:108-119
                        str(prof[ptr[j]:ptr[j + 1]].tolist()))))
                out.write("\n".join(rows) + "\n")
            at += n_take
        if native and in_flight is not None:
            in_flight.result()
    except BaseException:
        if native:
            formatter.shutdown(wait=True)       # never close the file under a chunk that is still being formatted
            lib.rt_tsv_close(handle)
        else:
            out.close()
        raise
    if native:
        formatter.shutdown(wait=True)
        if lib.rt_tsv_close(handle) != 0:       # the file has a write-behind thread: a failed write shows up here
            raise OSError(f"cannot write {path}: {lib.rt_io_last_error().decode()}")
    else:
        out.close()


def export_wig(merged_alignments: MergedAlignments, prefix: str) -> None:
    """detect_orfs.py:327-351: variableStep WIG per strand, chromosomes in lexicographic order,
    only strands that carry coverage."""
    import ctypes as C

    eng = merged_alignments.engine
    names = eng.contig_names
    lib = eng.lib
    blocks = {0: {}, 1: {}}
    for strand, ci, pos, cnt in merged_alignments.covered_by_contig():
        blocks[strand][ci] = (pos, cnt)
    for strand, tag in ((0, "pos"), (1, "neg")):
        if not blocks[strand]:
            continue
        handle = C.c_void_p()
        if lib.rt_wig_open(f"{prefix}_{tag}.wig".encode(), C.byref(handle)) != 0:
            raise OSError(lib.rt_io_last_error().decode())
        try:
            for ci in sorted(blocks[strand], key=lambda i: names[i]):        # chromosomes in lexicographic order
                pos, cnt = blocks[strand][ci]
                pos = np.ascontiguousarray(pos, np.int64)
                cnt = np.ascontiguousarray(cnt, np.int32)
                rc = lib.rt_wig_block(handle, names[ci].encode(), len(pos), pos.ctypes.data_as(C.c_void_p),
                                      cnt.ctypes.data_as(C.c_void_p))
                if rc != 0:
                    raise OSError("rt_wig_block failed")
        except BaseException:
            lib.rt_wig_close(handle)
            raise
        if lib.rt_wig_close(handle) != 0:       # the file has a write-behind thread: a failed write shows up here
            raise OSError(f"cannot write {prefix}_{tag}.wig: {lib.rt_io_last_error().decode()}")


def _stamp(msg: str) -> None:
    print("{} ... {}".format(datetime.datetime.now().strftime("%b %d %H:%M:%S"), msg))


def detect_orfs(
    bam,
    ribotricer_index: str,
    prefix: str,
    protocol: str | None,
    read_lengths: list | None,
    psite_offsets: dict | None,
    phase_score_cutoff: float,
    min_valid_codons: int,
    min_reads_per_codon: float,
    min_valid_codons_ratio: float,
    min_density_over_orf: float,
    report_all: bool,
    meta_min_reads: int = 100000,
) -> None:
    """Same positional signature and side files as detect_orfs.py:354-526.

    ``bam`` may be a BAM (decoded on the host by the native decoder, csrc/rt_bam.cpp), a ``.npz`` of decoded read columns
    or a ``ReadColumns`` object.
    """
    from . import metagene as mg
    from .bam import load_reads

    print(datetime.datetime.now().strftime("%b %d %H:%M:%S ..... started ribotricer detect-orfs"))
    _stamp("started parsing ribotricer index file")
    annotated, refseq = parse_ribotricer_index(ribotricer_index)
    pathlib.Path(os.path.dirname(prefix) or ".").mkdir(parents=True, exist_ok=True)   # common.py:103,131

    reads = load_reads(bam)
    if protocol is None:
        _stamp("started inferring experimental design")
        protocol = mg.infer_protocol(reads, refseq, prefix)
    del refseq

    _stamp("started reading bam file")
    alignments, read_length_counts = split_bam(reads, protocol, prefix, read_lengths)

    _stamp("started plotting read length distribution")
    mg.plot_read_lengths(read_length_counts, prefix)

    _stamp("started calculating metagene profiles. This may take a long time...")
    metagenes = mg.metagene_coverage(annotated, alignments, read_length_counts, prefix,
                                     meta_min_reads=meta_min_reads)
    _stamp("started plotting metagene profiles")
    mg.plot_metagene(metagenes, read_length_counts, prefix)

    if psite_offsets is None:
        _stamp("started inferring P-site offsets")
        psite_offsets = mg.align_metagenes(metagenes, read_length_counts, prefix, phase_score_cutoff,
                                           read_lengths is None)

    _stamp("started shifting according to P-site offsets")
    merged_alignments = merge_read_lengths(alignments, psite_offsets)

    _stamp("started exporting wig file of alignments after shifting")
    export_wig(merged_alignments, prefix)

    _stamp("started calculating phase scores for each ORF")
    export_orf_coverages(ribotricer_index, merged_alignments, prefix, phase_score_cutoff, min_valid_codons,
                         min_reads_per_codon, min_valid_codons_ratio, min_density_over_orf, report_all)
    _stamp("finished ribotricer detect-orfs")


__all__ = ["detect_orfs", "merge_read_lengths", "orf_coverage", "export_orf_coverages", "export_wig",
           "parse_ribotricer_index", "MergedAlignments", "get_engine", "load_index", "sys"]
