"""ORF-sharded detect-orfs over the GPUs of one box (north star item 4).

One process per GPU (``torchrun``).  The rows of the index are cut into byte-balanced blocks ALONG THE GENOME
(``shard_plan``): a rank scores the ORFs of one block, so the coverage it needs is the exon union of that block
(its own compact layout) and the reads it needs are those of the block's genomic span -- a contiguous slice of a
coordinate-sorted library.  K1, the host-to-device copy and the scoring all shrink with the number of ranks;
nothing is replicated but the index file.  A prepare-orfs index is in genome order inside its annotated block and
after it (prepare_orfs.py:322-365 walks a sorted GTF), so a rank's rows are two runs of consecutive rows and the
output is the concatenation of the runs' part files in row order.  The data path has no collective:
``torch.distributed`` only carries control messages (the inferred P-site offsets and a barrier before the parts
are joined).
"""
from __future__ import annotations

import os
import shutil

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Shard:
    """What one rank works on: ``rows`` (ascending row numbers of the index), their maximal ``runs`` of consecutive
    rows as (first_row, end_row, position of first_row in ``rows``), and the genomic ``spans`` (contig, lo, hi;
    1-based closed) that hold every exon of those rows."""
    rows: np.ndarray
    runs: list = field(default_factory=list)
    spans: list = field(default_factory=list)


def shard_plan(exon_ptr, exon_start, exon_end, orf_contig, n_shards: int, max_runs_per_shard: int = 64) -> list:
    """Cut the rows into ``n_shards`` blocks of equal algorithmic bytes (4 L + 8 E + 42 per ORF)
    in GENOME order (contig, start of the first interval; ties in row order).  If that scatters a rank's rows over
    more than ``max_runs_per_shard`` runs (an index that is not in genome order), the blocks are contiguous row
    ranges instead (``shard_bounds``) -- still correct, but the rows of a rank then span the whole genome."""
    exon_ptr = np.asarray(exon_ptr, np.int64)
    n = len(exon_ptr) - 1
    exlen = np.asarray(exon_end, np.int64) - np.asarray(exon_start, np.int64) + 1
    cs = np.concatenate([[0], np.cumsum(exlen)])
    orf_len = cs[exon_ptr[1:]] - cs[exon_ptr[:-1]]
    n_ex = np.diff(exon_ptr)
    cost = 4 * orf_len + 8 * n_ex + 42
    first = np.where(n_ex > 0, np.asarray(exon_start, np.int64)[np.minimum(exon_ptr[:-1], max(len(exlen) - 1, 0))], 0) if n else np.zeros(0, np.int64)
    contig = np.asarray(orf_contig, np.int64)
    order = np.lexsort((np.arange(n), first, np.where(contig < 0, np.iinfo(np.int64).max, contig)))
    prefix = np.concatenate([[0], np.cumsum(cost[order])])
    total = int(prefix[-1])
    cuts = [0]
    for s in range(1, n_shards):
        cuts.append(max(cuts[-1], min(n, int(np.searchsorted(prefix, total * s // n_shards, side="left")))))
    cuts.append(n)
    blocks = [np.sort(order[cuts[s]:cuts[s + 1]]) for s in range(n_shards)]

    def runs_of(rows):
        if len(rows) == 0:
            return []
        brk = np.flatnonzero(np.diff(rows) != 1) + 1
        starts = np.concatenate([[0], brk])
        ends = np.concatenate([brk, [len(rows)]])
        return [(int(rows[a]), int(rows[b - 1]) + 1, int(a)) for a, b in zip(starts, ends)]

    plan = [Shard(rows, runs_of(rows)) for rows in blocks]
    if sum(len(sh.runs) for sh in plan) > max_runs_per_shard * n_shards:
        b = shard_bounds(orf_len, n_ex, n_shards)
        plan = [Shard(np.arange(b[s], b[s + 1]), [(int(b[s]), int(b[s + 1]), 0)] if b[s + 1] > b[s] else [])
                for s in range(n_shards)]
    ex_start, ex_end = np.asarray(exon_start, np.int64), np.asarray(exon_end, np.int64)
    for sh in plan:
        spans = []
        if len(sh.rows):
            # exons of the shard's rows, per contig
            lo_e, hi_e = exon_ptr[sh.rows], exon_ptr[sh.rows + 1]
            sel = np.repeat(lo_e, hi_e - lo_e) + (np.arange(int((hi_e - lo_e).sum())) - np.repeat(np.cumsum(hi_e - lo_e) - (hi_e - lo_e), hi_e - lo_e))
            ec = np.repeat(contig[sh.rows], hi_e - lo_e)
            for c in np.unique(ec):
                if c < 0:
                    continue
                m = ec == c
                spans.append((int(c), int(ex_start[sel[m]].min()), int(ex_end[sel[m]].max())))
        sh.spans = spans
    return plan


def sub_index(columns: dict, rows) -> dict:
    """The CSR index columns (``Engine.set_index`` arguments) restricted to ``rows``, in that order."""
    exon_ptr = np.asarray(columns["exon_ptr"], np.int64)
    rows = np.asarray(rows, np.int64)
    lo, hi = exon_ptr[rows], exon_ptr[rows + 1]
    cnt = hi - lo
    ptr = np.zeros(len(rows) + 1, np.int64)
    np.cumsum(cnt, out=ptr[1:])
    sel = np.repeat(lo, cnt) + (np.arange(int(ptr[-1])) - np.repeat(ptr[:-1], cnt))
    return dict(exon_ptr=ptr, exon_start=np.ascontiguousarray(np.asarray(columns["exon_start"])[sel]),
                exon_end=np.ascontiguousarray(np.asarray(columns["exon_end"])[sel]),
                orf_contig=np.ascontiguousarray(np.asarray(columns["orf_contig"])[rows]),
                orf_strand=np.ascontiguousarray(np.asarray(columns["orf_strand"])[rows]))


def read_slices(ref_id, first, last, spans, max_offset: int = 65535) -> list:
    """Index ranges [a, b) of a COORDINATE-SORTED library (by ref_id, then first) that hold every read whose
    P-site can fall into ``spans``: a read's 5' end is its first or its last matched position and the P-site lies
    within ``max_offset`` of it, so a read matters when first <= hi + max_offset and last >= lo - max_offset; with
    the reads sorted by ``first`` that is the range first in [lo - max_offset - longest read span, hi + max_offset]."""
    ref_id, first, last = np.asarray(ref_id), np.asarray(first), np.asarray(last)
    if len(ref_id) == 0:
        return []
    reach = int(max_offset) + int((last.astype(np.int64) - first).max())
    key = ref_id.astype(np.int64) * (1 << 32) + first.astype(np.int64)
    out = []
    for c, lo, hi in sorted(spans):
        a = int(np.searchsorted(key, c * (1 << 32) + max(lo - 1 - reach, 0), side="left"))
        b = int(np.searchsorted(key, c * (1 << 32) + hi + max_offset, side="right"))
        if out and a <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], b))
        elif b > a:
            out.append((a, b))
    return out


def shard_bounds(orf_len, exons_per_orf, n_shards: int) -> np.ndarray:
    """Cut [0, n) into ``n_shards`` contiguous ranges of equal algorithmic bytes
    (4 L + 8 E + 42 per ORF) -- the rule of ``rt_shard_bounds``."""
    orf_len = np.asarray(orf_len, np.int64)
    cost = 4 * orf_len + 8 * np.asarray(exons_per_orf, np.int64) + 42
    prefix = np.concatenate([[0], np.cumsum(cost)])
    n = len(orf_len)
    bounds = np.zeros(n_shards + 1, np.int64)
    total = int(prefix[-1])
    for s in range(1, n_shards):
        target = total * s // n_shards
        bounds[s] = max(bounds[s - 1], min(n, int(np.searchsorted(prefix, target, side="left"))))
    bounds[n_shards] = n
    return bounds


def world():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def join_runs(prefix: str, plan: list, suffix: str = "_translating_ORFs.tsv") -> str:
    """Concatenate the part files of all runs of ``plan`` in row order: ``{prefix}{suffix}.rows{first_row:012d}``."""
    final = f"{prefix}{suffix}"
    starts = sorted(run[0] for sh in plan for run in sh.runs)
    with open(final, "wb") as out:
        with open(f"{final}.header", "rb") as fh:
            shutil.copyfileobj(fh, out)
        os.remove(f"{final}.header")
        for g_lo in starts:
            part = f"{final}.rows{g_lo:012d}"
            with open(part, "rb") as fh:
                shutil.copyfileobj(fh, out)
            os.remove(part)
    return final


def join_parts(prefix: str, n_parts: int, suffix: str = "_translating_ORFs.tsv") -> str:
    """Concatenate ``{prefix}{suffix}.part{r}`` (r = 0..n_parts-1) in rank order."""
    final = f"{prefix}{suffix}"
    with open(final, "wb") as out:
        for r in range(n_parts):
            part = f"{final}.part{r}"
            with open(part, "rb") as fh:
                shutil.copyfileobj(fh, out)
            os.remove(part)
    return final


def gather_columns(local: dict, dist=None) -> dict | None:
    """Rank-ordered concatenation of per-shard result columns on rank 0 (control-plane gather of
    host arrays; used by tests and small runs -- large runs write TSV parts instead)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, parts, dst=0)
    if dist.get_rank() != 0:
        return None
    return {k: np.concatenate([p[k] for p in parts]) for k in local}


def shard_reads(reads, shard: Shard, psite_offsets: dict):
    """The reads a rank has to look at: the slices of a coordinate-sorted library that can put a P-site into the
    shard's genomic spans (``read_slices``); every read when the library is not sorted."""
    from .bam import ReadColumns

    if not reads.sorted_by_coordinate or len(reads) == 0:
        return reads
    reach = max([abs(int(v)) for v in psite_offsets.values()] + [0])
    slices = read_slices(reads.cols["ref_id"], reads.cols["first"], reads.cols["last"], shard.spans, reach)
    cols = {k: (np.concatenate([v[a:b] for a, b in slices]) if slices else v[:0]) for k, v in reads.cols.items()}
    return ReadColumns(reads.contig_names, reads.contig_len, cols, True)


def score_shard(eng, idx, shard: Shard, reads, protocol, read_lengths, psite_offsets, params, prefix, report_all):
    """One rank's part of export_orf_coverages: its rows as the engine's resident (sub-)index in the compact
    layout, its slice of the library binned into the exon union of those rows, one TSV part per run of rows."""
    from .bam import Alignments
    from .detect_orfs import MergedAlignments, write_tsv

    if len(shard.rows) == 0:
        return {}, 0
    lut = {n: i for i, n in enumerate(eng.contig_names)}
    eng.set_index(**sub_index(idx.device_columns(lut), shard.rows))
    eng.set_layout("compact")
    try:
        mine = shard_reads(reads, shard, psite_offsets)
        aln = Alignments(eng, mine, protocol, read_lengths)
        cov = eng.new_coverage()
        aln.bin_into(cov, psite_offsets)
        res = eng.score_host(cov, 0, len(shard.rows), params)
        merged = MergedAlignments(eng, cov)
        for g_lo, g_hi, r_lo in shard.runs:
            write_tsv(f"{prefix}_translating_ORFs.tsv.rows{g_lo:012d}", idx, res, merged, g_lo, g_hi, report_all,
                      write_header=False, res_offset=r_lo)
    finally:
        eng.set_layout("dense")          # what every other entry point of the package expects to find
    eng._resident_index = None
    return res, len(mine)


def detect_orfs_sharded(bam, ribotricer_index, prefix, protocol, read_lengths, psite_offsets, phase_score_cutoff,
                        min_valid_codons, min_reads_per_codon, min_valid_codons_ratio, min_density_over_orf,
                        report_all, meta_min_reads: int = 100000):
    """detect_orfs() across the ranks of a torchrun job (same arguments as detect_orfs.py:354).  Rank 0 does what
    needs the whole library (protocol / offset inference, the BAM summary, the WIG files); every rank then scores its
    genomic block of the index from its slice of the reads.  Returns this rank's ``Shard``."""
    import torch
    import torch.distributed as dist

    from . import metagene as mg
    from .bam import load_reads, split_bam
    from .const import DEFAULT_PAD
    from .detect_orfs import TSV_COLUMNS, export_wig, get_engine, load_index, merge_read_lengths, parse_ribotricer_index
    from .engine import ScoreParams

    rank, size, local = world()
    if size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = get_engine(local)
    reads = load_reads(bam)
    annotated, refseq = parse_ribotricer_index(ribotricer_index)
    if rank == 0:
        os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    # control plane: rank 0 infers protocol / offsets once, everybody gets the answer
    decided = [protocol, psite_offsets]
    if rank == 0:
        if protocol is None:
            decided[0] = mg.infer_protocol(reads, refseq, prefix)
        alignments, rlc = split_bam(reads, decided[0], prefix, read_lengths, engine=eng)
        if psite_offsets is None:
            metagenes = mg.metagene_coverage(annotated, alignments, rlc, prefix, meta_min_reads=meta_min_reads)
            decided[1] = dict(mg.align_metagenes(metagenes, rlc, prefix, phase_score_cutoff, read_lengths is None))
    if size > 1:
        dist.broadcast_object_list(decided, src=0)
    protocol, psite_offsets = decided
    if rank == 0:
        merged = merge_read_lengths(alignments, psite_offsets)     # the genome-wide planes, for the WIG files only
        export_wig(merged, prefix)
        del merged, alignments
        with open(f"{prefix}_translating_ORFs.tsv.header", "w") as fh:
            fh.write("\t".join(TSV_COLUMNS) + "\n")
    elif (list(eng.contig_names) != list(reads.contig_names) or eng.pad != DEFAULT_PAD
            or not np.array_equal(eng.contig_len, reads.contig_len)):
        eng.set_genome(reads.contig_names, reads.contig_len, DEFAULT_PAD)
    idx = load_index(ribotricer_index)
    lut = {n: i for i, n in enumerate(eng.contig_names)}
    cols = idx.device_columns(lut)
    plan = shard_plan(cols["exon_ptr"], cols["exon_start"], cols["exon_end"], cols["orf_contig"], size)
    params = ScoreParams(phase_score_cutoff, min_valid_codons, min_reads_per_codon, min_valid_codons_ratio,
                         min_density_over_orf)
    score_shard(eng, idx, plan[rank], reads, protocol, read_lengths, psite_offsets, params, prefix, report_all)
    if size > 1:
        dist.barrier()
    if rank == 0:
        join_runs(prefix, plan)
    return plan[rank]
