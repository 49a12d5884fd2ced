"""ORF-sharded detect-orfs over the GPUs of one box (north star item 4).

One process per GPU (``torchrun``).  Every rank bins the whole library into its own replica of
the coverage planes and scores one contiguous, byte-balanced range of the index; concatenating
the rank outputs in rank order gives the rows in index order.  The data path has no
collective: ``torch.distributed`` only carries control messages (the inferred P-site offsets
and a barrier before the parts are joined).
"""
from __future__ import annotations

import os
import shutil

import numpy as np


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


def detect_orfs_sharded(bam, ribotricer_index, prefix, protocol, read_lengths, psite_offsets, phase_score_cutoff,
                        min_valid_codons, min_reads_per_codon, min_valid_codons_ratio, min_density_over_orf,
                        report_all, meta_min_reads: int = 100000):
    """detect_orfs() across the ranks of a torchrun job (same arguments as detect_orfs.py:354)."""
    import torch
    import torch.distributed as dist

    from . import metagene as mg
    from .bam import load_reads, split_bam
    from .detect_orfs import (export_orf_coverages, export_wig, get_engine, load_index, merge_read_lengths,
                              parse_ribotricer_index)

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
    if rank != 0:
        alignments, _ = split_bam(reads, protocol, f"{prefix}.rank{rank}", read_lengths, engine=eng)
        os.remove(f"{prefix}.rank{rank}_bam_summary.txt")
    merged = merge_read_lengths(alignments, psite_offsets)         # every rank: full coverage replica
    if rank == 0:
        export_wig(merged, prefix)
    idx = load_index(ribotricer_index)
    exons = np.diff(idx.exon_ptr)
    bounds = shard_bounds(idx.lengths(), exons, size)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    export_orf_coverages(ribotricer_index, merged, prefix, phase_score_cutoff, min_valid_codons, min_reads_per_codon,
                         min_valid_codons_ratio, min_density_over_orf, report_all, orf_range=(lo, hi),
                         write_header=(rank == 0), path=f"{prefix}_translating_ORFs.tsv.part{rank}")
    if size > 1:
        dist.barrier()
    if rank == 0:
        join_parts(prefix, size)
    return lo, hi
