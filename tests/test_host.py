"""Host-side logic that needs no GPU: CLI validation, index parsing, sharding arithmetic."""
import numpy as np


def test_cli_validation(tmp_path):
    """Flag validation messages of cli.py:236-270 (no GPU work happens on these paths)."""
    from click.testing import CliRunner

    from ribotricer_b200.cli import cli

    f = tmp_path / "x.npz"
    f.write_bytes(b"")
    i = tmp_path / "i.tsv"
    i.write_text("h\n")
    run = CliRunner().invoke
    base = ["detect-orfs", "--bam", str(f), "--ribotricer_index", str(i), "--prefix", str(tmp_path / "p")]
    r = run(cli, ["detect-orfs", "--bam", "nope", "--ribotricer_index", str(i), "--prefix", "p"])
    assert "Error: BAM file not found" in str(r.exception) or "Error: BAM file not found" in r.output
    r = run(cli, base + ["--psite_offsets", "12"])
    assert "psite_offsets only allowed when read_lengths is provided" in str(r.exception) + r.output
    r = run(cli, base + ["--read_lengths", "28,29", "--psite_offsets", "12"])
    assert "psite_offsets must match read_lengths" in str(r.exception) + r.output
    r = run(cli, base + ["--read_lengths", "28", "--psite_offsets", "28"])
    assert "P-site offset must be smaller than read length" in str(r.exception) + r.output
    r = run(cli, base + ["--read_lengths", "a"])
    assert "cannot convert read_lengths into integers" in str(r.exception) + r.output


def test_index_parser_matches_reference_rules(tmp_path):
    """orf.py:121-182: 11 columns, intervals sorted by start, oid re-derived (orf.py:100-103)."""
    import pytest

    from ribotricer_b200.index import ORF, parse_index

    p = tmp_path / "idx.tsv"
    p.write_text("ORF_ID\tORF_type\ttranscript_id\ttranscript_type\tgene_id\tgene_name\tgene_type\tchrom\tstrand\tstart_codon\tcoordinate\n"
                 "IGNORED\tannotated\ttxA\tprotein_coding\tgA\tGA\tprotein_coding\tchrI\t+\tATG\t201-212,101-109\n"
                 "IGNORED\tuORF\ttxB\tprotein_coding\tgB\tGB\tprotein_coding\tchrI\t-\tCTG\t301-318\n")
    idx = parse_index(str(p))
    assert idx.n_orf == 2 and idx.n_annotated_prefix == 1
    assert idx.exon_start.tolist() == [101, 201, 301] and idx.exon_end.tolist() == [109, 212, 318]
    assert idx.oid(0) == "txA_101_212_21" and idx.oid(1) == "txB_301_318_18"
    assert idx.lengths().tolist() == [21, 18]
    orf = ORF.from_string("x\tannotated\ttxA\tpc\tgA\tGA\tpc\tchrI\t+\tATG\t201-212,101-109\n")
    assert orf.oid == "txA_101_212_21" and orf.intervals == [(101, 109), (201, 212)]
    with pytest.raises(SystemExit):
        ORF.from_string("too\tfew\tcolumns")
    cols = idx.device_columns({"chrI": 0})
    assert cols["orf_contig"].tolist() == [0, 0] and cols["orf_strand"].tolist() == [0, 1]
    assert idx.device_columns({})["orf_contig"].tolist() == [-1, -1]
