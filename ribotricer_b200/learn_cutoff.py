"""learn-cutoff on the GPU path.

``determine_cutoff_tsv`` / ``determine_cutoff_bam`` keep the reference's signatures and printed report
(learn_cutoff.py:35-144, 147-270).  The replicate index matrices are drawn exactly as the reference
draws them (``np.random.seed(42)`` + ``np.random.choice``, NumPy's legacy generator), so the replicates are
the same; the 2 x ``reps`` medians over ``sampling_ratio * n`` gathered phase scores -- the part that
costs time -- are computed by ``rt_bootstrap_medians`` (one CUDA block per replicate, order statistics by
bisection on the key bits).  ``determine_cutoff_scores`` takes the phase-score columns directly, without
the TSV round trip.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

from .const import (CUTOFF, MINIMUM_DENSITY_OVER_ORF, MINIMUM_READS_PER_CODON, MINIMUM_VALID_CODONS,
                    MINIMUM_VALID_CODONS_RATIO)


def _bootstrap_medians(engine, scores: np.ndarray, indices: np.ndarray) -> np.ndarray:
    scores = np.ascontiguousarray(scores, np.float64)
    indices = np.ascontiguousarray(indices, np.int64)
    n_sel, reps = indices.shape
    out = np.empty(reps, np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
    engine._check(engine.lib.rt_bootstrap_medians(engine.ctx, p(scores), len(scores), p(indices), n_sel, reps, p(out)))
    return out


def determine_cutoff_scores(ribo_scores, rna_scores, sampling_ratio: float = 0.33, reps: int = 10000, engine=None) -> None:
    """learn_cutoff.py:81-144 on the already filtered phase-score columns (annotated ORFs of the wanted
    transcript types, all Ribo-seq samples concatenated, likewise RNA-seq)."""
    from .detect_orfs import get_engine

    engine = engine or get_engine()
    ribo_all = np.asarray(ribo_scores, np.float64)
    rna_all = np.asarray(rna_scores, np.float64)
    n_select_ribo = int(sampling_ratio * len(ribo_all))
    n_select_rna = int(sampling_ratio * len(rna_all))
    np.random.seed(42)                                                           # learn_cutoff.py:88-90
    ribo_indices = np.random.choice(range(len(ribo_all)), (n_select_ribo, reps))
    rna_indices = np.random.choice(range(len(rna_all)), (n_select_rna, reps))
    ribo_medians = _bootstrap_medians(engine, ribo_all, ribo_indices)
    rna_medians = _bootstrap_medians(engine, rna_all, rna_indices)
    diff_medians = ribo_medians - rna_medians
    diff_all = ribo_all - rna_all                                                # learn_cutoff.py:115 (equal lengths)
    print(f"sampling_ratio: {sampling_ratio}")
    print(f"n_samples: {reps}")
    for name, arr in (("ribo_phase_score", ribo_medians), ("rna_phase_score", rna_medians)):
        print(f"{name}_mean: {np.mean(arr):.3f}")
        print(f"{name}_median: {np.median(arr):.3f}")
        print(f"{name}_sd: {np.std(arr):.3f}")
    print(f"diff_phase_score_sampled_mean: {np.mean(diff_medians):.3f}")
    print(f"diff_phase_score_sampled_median: {np.median(diff_medians):.3f}")
    print(f"diff_phase_score_sampled_sd: {np.std(diff_medians):.3f}")
    print(f"diff_phase_score_all_mean: {np.mean(diff_all):.3f}")
    print(f"diff_phase_score_all_median: {np.median(diff_all):.3f}")
    print(f"diff_phase_score_all_sd: {np.std(diff_all):.3f}")
    print(f"recommended_cutoff: {np.median(diff_medians):.3f}")


def _filtered_scores(tsvs, filter_by) -> np.ndarray:
    """Phase scores of the annotated ORFs whose transcript type is wanted (learn_cutoff.py:57-79)."""
    keep = []
    wanted = {x.lower() for x in filter_by}
    for tsv in tsvs:
        with open(tsv) as fh:
            header = fh.readline().rstrip("\n").split("\t")
            c_type, c_score, c_tt = header.index("ORF_type"), header.index("phase_score"), header.index("transcript_type")
            for line in fh:
                f = line.rstrip("\n").split("\t")
                if f[c_type] == "annotated" and f[c_tt].lower() in wanted:
                    keep.append(float(f[c_score]))
    return np.array(keep, np.float64)


def determine_cutoff_tsv(ribo_tsvs, rna_tsvs, filter_by=None, sampling_ratio: float = 0.33, reps: int = 10000,
                         engine=None) -> None:
    """Same signature and printed report as learn_cutoff.py:35-41."""
    if filter_by is None:
        filter_by = ["protein_coding"]
    determine_cutoff_scores(_filtered_scores(ribo_tsvs, filter_by), _filtered_scores(rna_tsvs, filter_by),
                            sampling_ratio, reps, engine)


def determine_cutoff_bam(ribo_bams, rna_bams, ribotricer_index, prefix, ribo_stranded_protocols=None,
                         rna_stranded_protocols=None, filter_by=None, sampling_ratio: float = 0.33, reps: int = 10000,
                         phase_score_cutoff: float = CUTOFF, min_valid_codons: int = MINIMUM_VALID_CODONS,
                         report_all: bool = True) -> None:
    """learn_cutoff.py:147-270: detect-orfs with cutoff 0 on every BAM (the GPU path), then the bootstrap."""
    from .detect_orfs import detect_orfs

    ribo_stranded_protocols = list(ribo_stranded_protocols or [])
    rna_stranded_protocols = list(rna_stranded_protocols or [])
    if len(ribo_stranded_protocols) > 1:
        if len(ribo_stranded_protocols) != len(ribo_bams):
            sys.exit("Error: Ribo-seq protocol and bam file length mismatch")
    else:
        ribo_stranded_protocols = [None] * len(ribo_bams)
    if len(rna_stranded_protocols) > 1:
        if len(ribo_stranded_protocols) != len(ribo_bams):      # the reference re-checks the Ribo-seq lists here
            sys.exit("Error: Ribo-seq protocol and bam file length mismatch")
    else:
        rna_stranded_protocols = [None] * len(rna_bams)
        ribo_stranded_protocols = [None] * len(ribo_bams)
    sample_str = "sample" if len(rna_bams) == 1 else "samples"
    tsvs = {"ribo": [], "rna": []}
    for kind, label, bams, protocols in (("ribo", "Ribo-seq", ribo_bams, ribo_stranded_protocols),
                                         ("rna", "RNA-seq", rna_bams, rna_stranded_protocols)):
        print(f"Running ribotricer on {len(rna_bams)} {label} {sample_str} ..... \n")
        for i, (bam, stranded) in enumerate(zip(bams, protocols)):
            bam_prefix = f"{prefix}__{kind}_bam_{i + 1}"
            os.makedirs(os.path.dirname(os.path.abspath(bam_prefix)), exist_ok=True)
            detect_orfs(bam, ribotricer_index, bam_prefix, stranded, read_lengths=None, psite_offsets=None,
                        phase_score_cutoff=0.0, min_valid_codons=MINIMUM_VALID_CODONS,
                        min_reads_per_codon=MINIMUM_READS_PER_CODON, min_valid_codons_ratio=MINIMUM_VALID_CODONS_RATIO,
                        min_density_over_orf=MINIMUM_DENSITY_OVER_ORF, report_all=report_all)
            tsvs[kind].append(f"{bam_prefix}_translating_ORFs.tsv")
    determine_cutoff_tsv(tsvs["ribo"], tsvs["rna"], filter_by, sampling_ratio, reps)
