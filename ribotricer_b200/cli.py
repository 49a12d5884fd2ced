"""``ribotricer detect-orfs`` command line (mirror of ribotricer/cli.py:127-289).

Same flags, defaults, validation messages and flag->argument renames as the reference
(``--min_read_density`` -> ``min_density_over_orf``, ``--stranded yes`` -> ``forward``,
``--meta-min-reads``).  Only the detect-orfs sub-command is provided: the other sub-commands are
off the path this package accelerates.  ``--bam`` also accepts a ``.npz`` of decoded read columns.
"""
from __future__ import annotations

import os
import sys

import click

from . import __version__
from .const import (CUTOFF, META_MIN_READS, MINIMUM_DENSITY_OVER_ORF, MINIMUM_READS_PER_CODON,
                    MINIMUM_VALID_CODONS, MINIMUM_VALID_CODONS_RATIO)

CONTEXT_SETTINGS = {"help_option_names": ["-h", "--help"]}

try:   # cli.py:24,45-47 -- colours are cosmetic; the package is absent in some environments
    from click_help_colors import HelpColorsGroup
    _GROUP_KW = dict(cls=HelpColorsGroup, help_headers_color="yellow", help_options_color="green")
except ImportError:
    _GROUP_KW = {}


@click.group(**_GROUP_KW)
@click.version_option(version=__version__)
def cli() -> None:
    """ribotricer: Tool for detecting translating ORF from Ribo-seq data (B200-native detect-orfs)"""


@cli.command("detect-orfs", context_settings=CONTEXT_SETTINGS, help="Detect translating ORFs from BAM file")
@click.option("--bam", help="Path to BAM file", required=True)
@click.option("--ribotricer_index",
              help=("Path to the index file of ribotricer\n"
                    "This file should be generated using ribotricer prepare-orfs"), required=True)
@click.option("--prefix", help="Prefix to output file", required=True)
@click.option("--stranded", type=click.Choice(["yes", "no", "reverse"]), default=None, show_default=True,
              help=("whether the data is from a strand-specific assay"
                    " If not provided, the experimental protocol will be automatically inferred"))
@click.option("--read_lengths", default=None, show_default=True,
              help=("Comma separated read lengths to be used, such as 28,29,30\n"
                    "If not provided, it will be automatically determined by assessing"
                    " the metagene periodicity"))
@click.option("--psite_offsets", default=None, show_default=True,
              help=("Comma separated P-site offsets for each read length "
                    "matching the read lengths provided.\n"
                    "If not provided, reads from different read lengths will be "
                    "automatically aligned using cross-correlation"))
@click.option("--phase_score_cutoff", type=float, default=CUTOFF, show_default=True,
              help="Phase score cutoff for determining active translation")
@click.option("--min_valid_codons", type=int, default=MINIMUM_VALID_CODONS, show_default=True,
              help="Minimum number of codons with non-zero reads for determining active translation")
@click.option("--min_reads_per_codon", type=int, default=MINIMUM_READS_PER_CODON, show_default=True,
              help="Minimum number of reads per codon for determining active translation")
@click.option("--min_valid_codons_ratio", type=float, default=MINIMUM_VALID_CODONS_RATIO, show_default=True,
              help="Minimum ratio of codons with non-zero reads to total codons for determining active translation")
@click.option("--min_read_density", type=float, default=MINIMUM_DENSITY_OVER_ORF, show_default=True,
              help="Minimum read density (total_reads/length) over an ORF total codons for determining active translation")
@click.option("--report_all", help=("Whether output all ORFs including those non-translating ones"), is_flag=True)
@click.option("--meta-min-reads", type=int, default=META_MIN_READS, show_default=True,
              help="Minimum number of reads for a read length to be considered")
def detect_orfs_cmd(bam, ribotricer_index, prefix, stranded, read_lengths, psite_offsets, phase_score_cutoff,
                    min_valid_codons, min_reads_per_codon, min_valid_codons_ratio, min_read_density, report_all,
                    meta_min_reads) -> None:
    """Argument validation exactly as cli.py:236-273, then detect_orfs()."""
    from .detect_orfs import detect_orfs

    if not os.path.isfile(bam):
        sys.exit("Error: BAM file not found")
    if not os.path.isfile(ribotricer_index):
        sys.exit("Error: ribotricer index file not found")
    read_lengths_list = None
    psite_offsets_dict = None
    if read_lengths is not None:
        try:
            read_lengths_list = [int(x.strip()) for x in read_lengths.strip().split(",")]
        except Exception:
            sys.exit("Error: cannot convert read_lengths into integers")
        if not all(x > 0 for x in read_lengths_list):
            sys.exit("Error: read length must be positive")
    if read_lengths_list is None and psite_offsets is not None:
        sys.exit("Error: psite_offsets only allowed when read_lengths is provided")
    if read_lengths_list is not None and psite_offsets is not None:
        try:
            psite_offsets_list = [int(x.strip()) for x in psite_offsets.strip().split(",")]
        except Exception:
            sys.exit("Error: cannot convert psite_offsets into integers")
        if len(read_lengths_list) != len(psite_offsets_list):
            sys.exit("Error: psite_offsets must match read_lengths")
        if not all(x >= 0 for x in psite_offsets_list):
            sys.exit("Error: P-site offset must be >= 0")
        if not all(x > y for (x, y) in zip(read_lengths_list, psite_offsets_list)):
            sys.exit("Error: P-site offset must be smaller than read length")
        psite_offsets_dict = dict(list(zip(read_lengths_list, psite_offsets_list)))
    if stranded == "yes":
        stranded = "forward"
    detect_orfs(bam, ribotricer_index, prefix, stranded, read_lengths_list, psite_offsets_dict,
                phase_score_cutoff, min_valid_codons, min_reads_per_codon, min_valid_codons_ratio,
                min_read_density, report_all, meta_min_reads)



@cli.command("count-orfs", context_settings=CONTEXT_SETTINGS, help="Count reads for detected ORFs at gene level")
@click.option("--ribotricer_index",
              help="Path to the index file of ribotricer\nThis file should be generated using ribotricer prepare-orfs",
              required=True)
@click.option("--detected_orfs",
              help="Path to the detected orfs file\nThis file should be generated using ribotricer detect-orfs",
              required=True)
@click.option("--features", help="ORF types separated with comma", required=True)
@click.option("--out", help="Path to output file", required=True)
@click.option("--report_all", help=("Whether output all ORFs including those non-translating ones"), is_flag=True)
def count_orfs_cmd(ribotricer_index, detected_orfs, features, out, report_all) -> None:
    """cli.py:292-339 of the reference (same flags and checks); the table is count_orfs.py:28-89."""
    from .count_orfs import count_orfs

    if not os.path.isfile(ribotricer_index):
        sys.exit("Error: ribotricer index file not found")
    if not os.path.isfile(detected_orfs):
        sys.exit("Error: detected orfs file not found")
    count_orfs(ribotricer_index, detected_orfs, set(x.strip() for x in features.split(",")), out, report_all)


def _split_list(text: str) -> list:
    """Comma separated option value -> items with surrounding blanks removed (common.py:148-161)."""
    return [item.strip(" ") for item in text.split(",")]


@cli.command("learn-cutoff", context_settings=CONTEXT_SETTINGS, help="Learn phase score cutoff from BAM/TSV file")
@click.option("--ribo_bams", help="Path(s) to Ribo-seq BAM file separated by comma")
@click.option("--rna_bams", help="Path(s) to RNA-seq BAM file separated by comma")
@click.option("--ribo_tsvs", help="Path(s) to Ribo-seq *_translating_ORFs.tsv file separated by comma")
@click.option("--rna_tsvs", help="Path(s) to RNA-seq *_translating_ORFs.tsv file separated by comma")
@click.option("--ribotricer_index",
              help=("Path to the index file of ribotricer\n"
                    "This file should be generated using ribotricer prepare-orfs (required for BAM input)"))
@click.option("--prefix", help="Prefix to output file")
@click.option("--filter_by_tx_annotation", help="transcript_type to filter regions by", type=str,
              default="protein_coding", show_default=True)
@click.option("--phase_score_cutoff", type=float, default=CUTOFF, show_default=True,
              help="Phase score cutoff for determining active translation (required for BAM input)")
@click.option("--min_valid_codons", type=int, default=MINIMUM_VALID_CODONS, show_default=True,
              help="Minimum number of codons with non-zero reads for determining active translation (required for BAM input)")
@click.option("--sampling_ratio", type=float, default=0.33, show_default=True,
              help="Number of protein coding regions to sample per bootstrap")
@click.option("--n_bootstraps", type=int, default=20000, show_default=True, help="Number of bootstraps")
def determine_cutoff_cmd(ribo_bams, rna_bams, ribo_tsvs, rna_tsvs, ribotricer_index, prefix, filter_by_tx_annotation,
                         phase_score_cutoff, min_valid_codons, sampling_ratio, n_bootstraps) -> None:
    """cli.py:435-560 of the reference: same flags, checks and messages."""
    from .learn_cutoff import determine_cutoff_bam, determine_cutoff_tsv

    filter_by = _split_list(filter_by_tx_annotation)
    if ribo_bams and ribo_tsvs:
        sys.exit("Error: --ribo-bams and --rna_bams cannot be specified together")
    if rna_bams and rna_tsvs:
        sys.exit("Error: --rna-bams and --rna_tsvs cannot be specified together")
    if (ribo_bams and rna_tsvs) or (rna_bams and ribo_tsvs):
        sys.exit("Error: BAM and TSV inputs cannot be specified together")
    if ribotricer_index and not os.path.isfile(ribotricer_index):
        sys.exit("Error: ribotricer index file not found")
    if ribo_bams:
        ribo_list, rna_list = _split_list(ribo_bams), (_split_list(rna_bams) if rna_bams else [])
        if ribo_list and rna_list:
            if not prefix:
                sys.exit("Error: --prefix required with BAM inputs")
            if not ribotricer_index:
                sys.exit("Error: --ribotricer_index required with BAM inputs")
            determine_cutoff_bam(ribo_list, rna_list, ribotricer_index, prefix, [], [], filter_by, sampling_ratio,
                                 n_bootstraps, phase_score_cutoff, min_valid_codons, report_all=True)
            return
        determine_cutoff_tsv([], [], filter_by, sampling_ratio, n_bootstraps)
        return
    determine_cutoff_tsv(_split_list(ribo_tsvs) if ribo_tsvs else [], _split_list(rna_tsvs) if rna_tsvs else [],
                         filter_by, sampling_ratio, n_bootstraps)


if __name__ == "__main__":
    cli()
