"""Generate the committed golden vectors by RUNNING THE UNMODIFIED REFERENCE.

Run in the authoring container only (needs /root/reference; scipy 1.18.1 /
numpy 2.3.5 at generation time -- both versions are recorded in the files):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs (small, committed):
  tests/golden/phasescore_cases.json.gz   statistics.phasescore (statistics.py:48)
  tests/golden/metagene_case.json.gz      metagene_coverage / align_metagenes (metagene.py:160,268)
  tests/golden/pipeline_cases.json.gz     merge_read_lengths (detect_orfs.py:54),
                                          orf_coverage (:134), export_orf_coverages
                                          (:206) and export_wig (:327) end to end

  tests/golden/split_bam_case.json.gz     split_bam (bam.py:33-153) on a real BAM (bytes included)
  tests/golden/count_orfs_cases.json.gz   count_orfs (count_orfs.py:28-89) on the pipeline cases' TSVs
  tests/golden/learn_cutoff_cases.json.gz determine_cutoff_tsv (learn_cutoff.py:35-144) stdout
  tests/golden/infer_protocol_case.json.gz infer_protocol (infer_protocol.py:34-124) + parse_ribotricer_index
                                          (detect_orfs.py:86-131) on a real BAM (bytes included)

``split_bam`` needs pysam, which is absent from this image: it runs UNMODIFIED on top of
oracle/pysam_restated.py, a pure-Python restatement of the handful of pysam calls it makes
(``python tests/golden/make_golden.py split_bam`` regenerates only that file).
"""
from __future__ import annotations

import gzip
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402


def versions():
    import scipy
    return {"scipy": scipy.__version__, "numpy": np.__version__, "reference": "ribotricer 1.5.0"}


# ---------------------------------------------------------------- phasescore
KAT_PROFILES = [
    [], [0], [1], [0, 0], [1, 0, 0], [0, 0, 0], [1, 1, 1], [1, 0, 0, 0, 0],
    [1] + [0] * 29, [0, 0, 0, 1] + [0] * 26, [2, 2, 2, 1, 0, 0, 1, 0, 0],
    [1, 1, 1] * 5, [0] * 30, [0, 1, 0] * 10, [0, 0, 1] * 10, [1, 0, 0] * 10,
    [3, 1, 2, 0, 0, 0, 5, 0, 1, 2, 2, 2, 0, 4, 0, 1],
    [0, 0, 0, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0], [0, 1, 1, 4, 1, 1, 2],
    [5, 0, 0, 3, 1, 0, 4, 0, 0, 2, 0, 0, 6, 0, 0, 0, 1, 0, 2, 0, 0],
    [2, 0, 0, 2, 0, 0, 1, 1, 0, 0, 0, 0, 3, 0, 0, 0, 0, 0],
    [1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 0, 7],
    [1000000, 3, 2] * 7, [500000000, 0, 1, 0, 500000000, 2, 7, 7, 7, 1, 2, 3],
    [0, 0, 5], [0, 5, 0, 0], [0, 0, 0, 0, 0, 9],
]


def make_phasescore_cases(phasescore):
    rng = np.random.default_rng(20261017)
    profiles = [list(p) for p in KAT_PROFILES]
    for _ in range(2400):
        length = int(rng.integers(1, 601))
        lam = float(rng.choice([0.005, 0.02, 0.1, 0.5, 2.0, 20.0, 200.0]))
        kind = rng.random()
        if kind < 0.5:      # 3-nt periodic Poisson
            frame = int(rng.integers(0, 3))
            w = np.full(length, 0.15 * lam)
            w[frame::3] = 0.7 * lam
            cov = rng.poisson(w * 3)
        else:               # aperiodic
            cov = rng.poisson(lam, length)
        if rng.random() < 0.25 and length > 6:   # inject uniform codons
            for _ in range(int(rng.integers(1, 4))):
                p = int(rng.integers(0, length - 3))
                cov[p:p + 3] = int(rng.integers(1, 6))
        profiles.append([int(x) for x in cov])
    out = []
    for p in profiles:
        score, valid = phasescore(p)
        out.append({"cov": p, "score": float(score).hex(), "valid": int(valid)})
    return out


# ------------------------------------------------------------------ pipeline
def appendix_a_case():
    """SURVEY.md Appendix A, regenerated (not transcribed) from the reference."""
    index = [
        "ORF_ID\tORF_type\ttranscript_id\ttranscript_type\tgene_id\tgene_name\tgene_type\tchrom\tstrand\tstart_codon\tcoordinate",
        "IGNORED\tannotated\ttxA\tprotein_coding\tgA\tGA\tprotein_coding\tchrI\t+\tATG\t101-109,201-212",
        "IGNORED\tuORF\ttxB\tprotein_coding\tgB\tGB\tprotein_coding\tchrI\t-\tCTG\t301-318",
        "IGNORED\tnovel\ttxC\tlncRNA\tgC\tGC\tlncRNA\tchrII\t+\tATG\t11-30",
    ]
    aln = [
        (28, "+", "chrI", 89, 5), (28, "+", "chrI", 92, 3), (28, "+", "chrI", 95, 4),
        (28, "+", "chrI", 192, 6), (28, "+", "chrII", -1, 1), (28, "+", "chrII", 2, 1),
        (28, "+", "chrII", 5, 1), (28, "+", "chrII", 8, 1), (28, "+", "chrII", 11, 1),
        (28, "+", "chrII", 14, 1),
        (29, "+", "chrI", 92, 1), (29, "+", "chrI", 188, 2), (29, "+", "chrI", 195, 1),
        (29, "+", "chrI", 197, 2), (29, "+", "chrII", 17, 7),
        (28, "-", "chrI", 330, 2), (28, "-", "chrI", 327, 2), (28, "-", "chrI", 318, 3),
        (29, "-", "chrI", 325, 1), (29, "-", "chrI", 324, 1),
        (30, "+", "chrI", 150, 99),
    ]
    return dict(name="appendix_a", contigs=[["chrI", 400], ["chrII", 100]], index=index,
                alignments=aln, psite_offsets={"28": 12, "29": 13})


def random_case(name, seed, n_tx, contigs, unknown_chrom=False, odd_lengths=False):
    rng = np.random.default_rng(seed)
    header = ("ORF_ID\tORF_type\ttranscript_id\ttranscript_type\tgene_id\tgene_name\t"
              "gene_type\tchrom\tstrand\tstart_codon\tcoordinate")
    lines, orfs = [], []
    cats = ["annotated", "uORF", "dORF", "novel", "super_uORF", "overlap_dORF"]
    for t in range(n_tx):
        cname, clen = contigs[int(rng.integers(0, len(contigs)))]
        strand = "+" if rng.random() < 0.5 else "-"
        n_ex = int(min(6, rng.geometric(0.45)))
        pos = int(rng.integers(1, max(2, clen - 1500)))
        exons = []
        for _ in range(n_ex):
            elen = int(rng.integers(1, 160)) if rng.random() < 0.9 else int(rng.integers(1, 4))
            exons.append((pos, pos + elen - 1))
            pos += elen + int(rng.integers(1, 120))
        exons = [(s, e) for s, e in exons if e <= clen]
        if not exons:
            continue
        total = sum(e - s + 1 for s, e in exons)
        # nested ORFs sharing the transcript 3' end: trim k nt from the 5' end
        n_nested = int(rng.integers(1, 4))
        for k in range(n_nested):
            trim = 0 if k == 0 else int(rng.integers(1, max(2, total - 3)))
            if not odd_lengths:
                trim -= trim % 3
            ivs = list(exons)
            left = trim
            if strand == "+":
                while left > 0 and ivs:
                    s, e = ivs[0]
                    if e - s + 1 <= left:
                        left -= e - s + 1
                        ivs.pop(0)
                    else:
                        ivs[0] = (s + left, e)
                        left = 0
            else:
                while left > 0 and ivs:
                    s, e = ivs[-1]
                    if e - s + 1 <= left:
                        left -= e - s + 1
                        ivs.pop()
                    else:
                        ivs[-1] = (s, e - left)
                        left = 0
            if not ivs:
                continue
            chrom = cname
            if unknown_chrom and rng.random() < 0.05:
                chrom = "chrUn_absent"
            cat = "annotated" if (k == 0 and t < n_tx // 4) else cats[int(rng.integers(1, len(cats)))]
            # exercise orf.py:100 (intervals are sorted on load): shuffle the text order
            txt_ivs = list(ivs)
            if rng.random() < 0.2:
                rng.shuffle(txt_ivs)
            coord = ",".join(f"{s}-{e}" for s, e in txt_ivs)
            lines.append(f"id{t}_{k}\t{cat}\ttx{t}\tprotein_coding\tg{t}\tG{t}\tprotein_coding\t"
                         f"{chrom}\t{strand}\t{['ATG','CTG','GTG','TTG'][int(rng.integers(0,4))]}\t{coord}")
            orfs.append((chrom, strand, ivs))
    # annotated rows first, like prepare-orfs writes them (prepare_orfs.py:322-329)
    order = sorted(range(len(lines)), key=lambda i: 0 if "\tannotated\t" in lines[i] else 1)
    lines = [lines[i] for i in order]
    orfs = [orfs[i] for i in order]
    # reads: P-sites in frame in a subset of "expressed" ORFs + background
    lens = [26, 27, 28, 29, 30, 31, 32]
    offsets = {26: 12, 27: 12, 28: 12, 29: 12, 30: 13, 31: 13}   # 32 has no offset -> dropped
    aln = {}

    def add(length, strand, chrom, psite, n=1):
        off = offsets.get(length, 12)
        pos5 = psite - off if strand == "+" else psite + off
        key = (length, strand, chrom, pos5)
        aln[key] = aln.get(key, 0) + n

    for chrom, strand, ivs in orfs:
        if chrom == "chrUn_absent" or rng.random() < 0.35:
            continue
        expr = float(rng.gamma(0.5) * 0.6)
        flat = [p for s, e in ivs for p in range(s, e + 1)]
        if strand == "-":
            flat.reverse()
        n_reads = int(rng.poisson(expr * len(flat) / 3))
        for _ in range(n_reads):
            codon = int(rng.integers(0, max(1, len(flat) // 3)))
            fr = int(rng.choice([0, 1, 2], p=[0.7, 0.15, 0.15]))
            idx = min(len(flat) - 1, 3 * codon + fr)
            add(int(rng.choice(lens)), strand, chrom, flat[idx])
    for cname, clen in contigs:
        for _ in range(int(clen * 0.02)):
            add(int(rng.choice(lens)), "+" if rng.random() < 0.5 else "-", cname,
                int(rng.integers(-5, clen + 20)), int(rng.integers(1, 4)))
    aln_list = [(k[0], k[1], k[2], k[3], v) for k, v in aln.items()]
    return dict(name=name, contigs=[list(c) for c in contigs], index=[header] + lines,
                alignments=aln_list, psite_offsets={str(k): v for k, v in offsets.items()})


# ------------------------------------------------------------------ metagene / offsets
def metagene_case(seed=1003):
    """Annotated CDSs with planted P-site offsets; the reference's metagene_coverage
    (metagene.py:160) and align_metagenes (metagene.py:268) produce the expected values."""
    from collections import Counter, defaultdict

    from ribotricer.metagene import align_metagenes, metagene_coverage
    from ribotricer.orf import ORF

    rng = np.random.default_rng(seed)
    contigs = [("mA", 60000), ("mB", 60000)]
    header = ("ORF_ID\tORF_type\ttranscript_id\ttranscript_type\tgene_id\tgene_name\t"
              "gene_type\tchrom\tstrand\tstart_codon\tcoordinate")
    lines, orfs = [], []
    planted = {27: 11, 28: 12, 29: 13, 30: 13, 31: 14}
    probs = [0.08, 0.40, 0.27, 0.20, 0.05]
    for t in range(90):
        cname, clen = contigs[t % 2]
        strand = "+" if rng.random() < 0.5 else "-"
        n_ex = int(rng.integers(1, 4))
        pos = 200 + (t // 2) * 1200
        total = 3 * int(rng.integers(60, 330))
        cuts = sorted(rng.choice(np.arange(1, total), n_ex - 1, replace=False).tolist()) if n_ex > 1 else []
        parts = [b - a for a, b in zip([0] + cuts, cuts + [total])]
        ivs = []
        for plen in parts:
            ivs.append((pos, pos + plen - 1))
            pos += plen + int(rng.integers(30, 90))
        coord = ",".join(f"{s}-{e}" for s, e in ivs)
        lines.append(f"id{t}\tannotated\ttx{t}\tprotein_coding\tg{t}\tG{t}\tprotein_coding\t{cname}\t{strand}\tATG\t{coord}")
        orfs.append((cname, strand, ivs))
    aln = {}
    for chrom, strand, ivs in orfs:
        flat = [p for s, e in ivs for p in range(s, e + 1)]
        if strand == "-":
            flat.reverse()
        n_reads = int(rng.poisson(len(flat) * float(rng.gamma(2.0) * 0.8)))
        for _ in range(n_reads):
            codon = int(rng.integers(0, len(flat) // 3))
            fr = int(rng.choice([0, 1, 2], p=[0.75, 0.15, 0.10]))
            psite = flat[3 * codon + fr]
            length = int(rng.choice(list(planted), p=probs))
            pos5 = psite - planted[length] if strand == "+" else psite + planted[length]
            key = (length, strand, chrom, pos5)
            aln[key] = aln.get(key, 0) + 1
    case = dict(name="metagene", contigs=[list(c) for c in contigs], index=[header] + lines,
                alignments=[(k[0], k[1], k[2], k[3], v) for k, v in aln.items()], planted=planted,
                meta_min_reads=1500)
    alignments = defaultdict(lambda: defaultdict(Counter))
    rlc = {}
    for length, strand, chrom, pos, n in case["alignments"]:
        alignments[length][strand][(chrom, pos)] += n
    # insertion order of read_length_counts = order of first appearance (bam.py:136); the test
    # feeds the reads grouped by ascending length, so mirror that here
    for length in sorted(alignments):
        rlc[length] = sum(sum(t.values()) for t in alignments[length].values())
    case["read_length_counts_in"] = {str(k): v for k, v in rlc.items()}
    cds = [ORF.from_string(line + "\n") for line in lines]
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, "mg")
        metagenes = metagene_coverage(cds, alignments, rlc, prefix, meta_min_reads=case["meta_min_reads"])
        case["kept_lengths"] = list(rlc)
        case["metagenes"] = {
            str(length): dict(idx5=[int(i) for i in m[0].index], prof5=[float(x) for x in m[0].tolist()],
                              idx3=[int(i) for i in m[1].index], prof3=[float(x) for x in m[1].tolist()],
                              ps5=float(m[2]), v5=int(m[3]), ps3=float(m[4]), v3=int(m[5]))
            for length, m in metagenes.items()}
        case["profiles_5p_tsv"] = open(f"{prefix}_metagene_profiles_5p.tsv").read()
        offsets = align_metagenes(metagenes, rlc, prefix, 0.428571428571, True)
        case["psite_offsets"] = {str(k): int(v) for k, v in offsets.items()}
        case["psite_offsets_txt"] = open(f"{prefix}_psite_offsets.txt").read()
    return case


PARAM_SETS = [
    dict(phase_score_cutoff=0.428571428571, min_valid_codons=5, min_reads_per_codon=0,
         min_valid_codons_ratio=0, min_density_over_orf=0.0, report_all=True),
    dict(phase_score_cutoff=0.428571428571, min_valid_codons=5, min_reads_per_codon=0,
         min_valid_codons_ratio=0, min_density_over_orf=0.0, report_all=False),
    dict(phase_score_cutoff=0.3, min_valid_codons=3, min_reads_per_codon=1,
         min_valid_codons_ratio=0.1, min_density_over_orf=0.25, report_all=True),
    dict(phase_score_cutoff=0.0, min_valid_codons=0, min_reads_per_codon=0,
         min_valid_codons_ratio=0.5, min_density_over_orf=1.0, report_all=False),
]


def split_bam_case(seed=1004):
    """A BAM with every branch of bam.py:71-137 / common.py:33-69 in it, and what the unmodified
    reference ``split_bam`` returns for it under several (protocol, read_lengths) settings."""
    import base64
    import io
    from contextlib import redirect_stdout

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bam_writer as W

    from ribotricer.bam import split_bam

    rng = np.random.default_rng(seed)
    refs = [("chrA", 5000), ("chrB", 3000), ("chrUn_1", 800)]
    cigars = [
        lambda L: [("M", L)],
        lambda L: [("M", L // 2), ("N", int(rng.integers(50, 400))), ("M", L - L // 2)],
        lambda L: [("S", 2), ("M", L), ("S", 1)],
        lambda L: [("M", 10), ("I", 2), ("M", L - 10)],
        lambda L: [("M", 8), ("D", 3), ("M", L - 8)],
        lambda L: [("H", 5), ("=", L - 4), ("X", 1), ("M", 3)],
        lambda L: [("M", 5), ("N", 100), ("M", 5), ("N", 80), ("M", L - 10)],
    ]
    flags = [0, 0, 0, 16, 16, 16, 256, 272, 512, 1024, 1040, 4, 20, 2048, 2064, 1, 83, 99, 147, 163, 528]
    recs = []
    hot = [(0, int(p)) for p in rng.integers(100, 4000, 12)] + [(1, int(p)) for p in rng.integers(100, 2500, 8)]
    for k in range(1400):
        if rng.random() < 0.55:                      # stacked 5' ends, the rule in Ribo-seq
            ref_id, pos = hot[int(rng.integers(len(hot)))]
            pos += int(rng.integers(0, 3))
        else:
            ref_id = int(rng.choice([0, 0, 0, 1, 1, 2]))
            pos = int(rng.integers(0, refs[ref_id][1] - 700))
        L = int(rng.integers(24, 36))
        flag = int(rng.choice(flags))
        mapq = int(rng.choice([255, 255, 255, 60, 3, 1, 0]))
        tag_kind = int(rng.integers(0, 8))
        aux = b""
        if tag_kind == 0:
            aux = W.aux_field("NH", "C", 1)
        elif tag_kind == 1:
            aux = W.aux_field("NH", "c", int(rng.choice([1, 2, 7])))
        elif tag_kind == 2:
            aux = W.aux_field("XS", "Z", "x") + W.aux_field("NH", "S", int(rng.choice([1, 300])))
        elif tag_kind == 3:
            aux = W.aux_field("NM", "i", 1) + W.aux_field("NH", "i", int(rng.choice([1, 4])))
        elif tag_kind == 4:
            aux = W.aux_field("AS", "C", 30) + W.aux_field("ZB", "Bs", [1, -2, 3])
        if flag & 4:
            ref_out, pos_out = (-1, -1) if rng.random() < 0.5 else (ref_id, pos)
        else:
            ref_out, pos_out = ref_id, pos
        if k % 97 == 0 and not flag & 4:
            ref_out = -1                             # mapped flag but no reference: chrom is None (bam.py:133)
        cigar = cigars[int(rng.integers(len(cigars)))](L)
        recs.append((ref_out if ref_out >= 0 else 1 << 30, pos_out,
                     W.record(ref_out, pos_out, mapq, flag, cigar, name=b"r%d" % k, aux=aux)))
    recs.sort(key=lambda r: (r[0], r[1]))
    case = {"name": "split_bam", "refs": refs, "n_records": len(recs), "runs": []}
    with tempfile.TemporaryDirectory() as tmp:
        bam = os.path.join(tmp, "lib.bam")
        W.write_bam(bam, refs, [r[2] for r in recs], sorted_header=True, block_payload=7001)
        case["bam_b64"] = base64.b64encode(open(bam, "rb").read()).decode()
        for protocol, read_lengths in (("forward", None), ("reverse", None), ("forward", [28, 29, 30]),
                                       ("reverse", [26, 31, 33]), ("no", None)):
            prefix = os.path.join(tmp, "out")
            sink = io.StringIO()
            with redirect_stdout(sink):              # is_read_uniq_mapping prints its warning per read
                alignments, rlc = split_bam(bam, protocol, prefix, read_lengths)
            flat = sorted([int(length), strand, chrom, int(pos), int(n)]
                          for length, by_strand in alignments.items()
                          for strand, ctr in by_strand.items() for (chrom, pos), n in ctr.items())
            case["runs"].append({"protocol": protocol, "read_lengths": read_lengths, "alignments": flat,
                                 "read_length_counts": {str(k): int(v) for k, v in rlc.items()},
                                 "summary": open(f"{prefix}_bam_summary.txt").read(),
                                 "warnings": sink.getvalue().count("WARNING")})
    return case


def infer_protocol_case(seed=1006):
    """A BAM + index on which the UNMODIFIED ``parse_ribotricer_index`` (detect_orfs.py:86-131) and
    ``infer_protocol`` (infer_protocol.py:34-124) run (pysam and quicksect restated, see oracle/ref_import.py):
    reads whose mapping strand follows their gene under a reverse-stranded protocol plus noise, every branch of
    is_read_uniq_mapping as infer_protocol uses it (bare truthiness), unmapped-flag reads that still carry
    coordinates, reads without CIGAR, spliced / clipped / indel CIGARs (reference_end counts D and N), overlapping
    annotated spans on opposite strands (len(interval) != 1), a read name-sorted after the annotated block."""
    import base64
    import io
    from contextlib import redirect_stdout

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bam_writer as W

    from ribotricer.detect_orfs import parse_ribotricer_index
    from ribotricer.infer_protocol import infer_protocol

    rng = np.random.default_rng(seed)
    refs = [("chrA", 30000), ("chrB", 16000), ("chrUn_2", 900)]
    header = ("ORF_ID\tORF_type\ttranscript_id\ttranscript_type\tgene_id\tgene_name\t"
              "gene_type\tchrom\tstrand\tstart_codon\tcoordinate")
    lines, genes = [], []
    pos = {"chrA": 300, "chrB": 200}
    for t in range(44):
        chrom = "chrA" if t % 3 else "chrB"
        strand = "+" if rng.random() < 0.5 else "-"
        n_ex = int(rng.integers(1, 4))
        start = pos[chrom]
        ivs, at = [], start
        for _ in range(n_ex):
            ln = 3 * int(rng.integers(20, 70))
            ivs.append((at, at + ln - 1))
            at += ln + int(rng.integers(40, 160))
        pos[chrom] = ivs[-1][1] + int(rng.integers(-150, 500))     # negative: the next span overlaps this one
        cat = "annotated"
        name = f"G{t}"
        if t == 30:            # still inside the leading block ("annotated" occurs in the line) but another category
            cat, name = "uORF", "annotated_like"
        coord = ",".join(f"{a}-{b}" for a, b in ivs)
        lines.append(f"id{t}\t{cat}\ttx{t}\tprotein_coding\tg{t}\t{name}\tprotein_coding\t{chrom}\t{strand}\tATG\t{coord}")
        if cat == "annotated":
            genes.append((chrom, ivs[0][0], ivs[-1][1], strand))
    # rows after the leading block are never read (detect_orfs.py:104-118), even an annotated one
    lines.append("idX\tnovel\ttxX\tlncRNA\tgX\tGX\tlncRNA\tchrA\t+\tCTG\t25000-25299")
    lines.append("idY\tannotated\ttxY\tprotein_coding\tgY\tGY\tprotein_coding\tchrA\t-\tATG\t26000-26299")
    cigars = [
        lambda L: [("M", L)],
        lambda L: [("M", L // 2), ("N", int(rng.integers(50, 300))), ("M", L - L // 2)],
        lambda L: [("S", 2), ("M", L), ("S", 1)],
        lambda L: [("M", 10), ("I", 2), ("M", L - 10)],
        lambda L: [("M", 8), ("D", 3), ("M", L - 8)],
        lambda L: [("M", L - 4), ("D", 2)],            # ends in a deletion: reference_end > last matched + 1
        lambda L: [("N", 7), ("M", L)],                # starts in a skip: reference_start < first matched
        lambda L: [("S", L)],                          # no reference-consuming operation: reference_end = pos + 1
        lambda L: [],                                  # no CIGAR: reference_end is None
    ]
    cig_p = [0.55, 0.12, 0.08, 0.05, 0.05, 0.04, 0.04, 0.04, 0.03]
    recs = []
    for k in range(3200):
        L = int(rng.integers(24, 36))
        if rng.random() < 0.8:
            chrom, g0, g1, gs = genes[int(rng.integers(len(genes)))]
            ref_id = 0 if chrom == "chrA" else 1
            p0 = int(rng.integers(max(0, g0 - 40), g1 + 10))
            # reverse-stranded library: reads map opposite to their gene, with 15 % noise
            minus = (gs == "+") != (rng.random() < 0.15)
        else:
            ref_id = int(rng.choice([0, 1, 2]))
            p0 = int(rng.integers(0, refs[ref_id][1] - 400))
            minus = bool(rng.random() < 0.5)
        flag = 16 if minus else 0
        r = rng.random()
        if r < 0.04: flag |= 256
        elif r < 0.07: flag |= 4
        elif r < 0.10: flag |= 512
        elif r < 0.13: flag |= 1024
        elif r < 0.15: flag |= 2048
        mapq = int(rng.choice([255, 255, 255, 60, 3, 1, 0]))
        kind = int(rng.integers(0, 12))
        aux = b""
        if kind in (0, 1, 2): aux = W.aux_field("NH", "C", 1)
        elif kind == 3: aux = W.aux_field("NH", "c", int(rng.choice([1, 2, 0, -1])))
        elif kind == 4: aux = W.aux_field("XS", "Z", "x") + W.aux_field("NH", "S", int(rng.choice([1, 300])))
        elif kind == 5: aux = W.aux_field("NH", "i", int(rng.choice([1, 4])))
        elif kind == 6: aux = W.aux_field("NH", "A", "1")
        elif kind == 7: aux = W.aux_field("NH", "f", float(rng.choice([1.0, 2.0])))
        elif kind == 8: aux = W.aux_field("NH", "C", 2) + W.aux_field("NH", "C", 1)     # dict(): the last one counts
        elif kind == 9: aux = W.aux_field("NH", "Z", "1")
        ref_out = ref_id if k % 131 else -1
        cigar = cigars[int(rng.choice(len(cigars), p=cig_p))](L)
        recs.append((ref_out if ref_out >= 0 else 1 << 30, p0,
                     W.record(ref_out, p0, mapq, flag, cigar, name=b"q%d" % k, aux=aux, l_seq=L)))
    recs.sort(key=lambda x: (x[0], x[1]))
    case = {"name": "infer_protocol", "refs": refs, "index": [header] + lines, "n_records": len(recs), "runs": []}
    with tempfile.TemporaryDirectory() as tmp:
        bam = os.path.join(tmp, "lib.bam")
        W.write_bam(bam, refs, [r[2] for r in recs], sorted_header=True, block_payload=9001)
        case["bam_b64"] = base64.b64encode(open(bam, "rb").read()).decode()
        idx = os.path.join(tmp, "idx.tsv")
        with open(idx, "w") as fh:
            fh.write("\n".join(case["index"]) + "\n")
        annotated, refseq = parse_ribotricer_index(idx)
        case["annotated_oids"] = [o.oid for o in annotated]
        case["refseq"] = {c: sorted([iv.start, iv.end, iv.data] for iv in tree.items) for c, tree in refseq.items()}
        for n_reads in (20000, 150, 0):
            prefix = os.path.join(tmp, f"out{n_reads}")
            sink = io.StringIO()
            with redirect_stdout(sink):
                protocol = infer_protocol(bam, refseq, prefix, n_reads)
            case["runs"].append({"n_reads": n_reads, "protocol": protocol,
                                 "text": open(f"{prefix}_protocol.txt").read(),
                                 "warnings": sink.getvalue().count("WARNING")})
        # the same library with every strand flipped must read as the other protocol
    return case


def start_codon_case():
    """Index rows whose start_codon field is not three characters long: the TSV prints ORF.start_codon
    (orf.py:108-119), i.e. seq[:3] or the string None, not the raw field."""
    header = ("ORF_ID\tORF_type\ttranscript_id\ttranscript_type\tgene_id\tgene_name\t"
              "gene_type\tchrom\tstrand\tstart_codon\tcoordinate")
    codons = ["ATG", "AT", "", "ATGCC", "N", "CTGA"]
    lines = [f"x{k}\t{'annotated' if k < 2 else 'uORF'}\tt{k}\tprotein_coding\tg{k}\tG{k}\tprotein_coding\tc1\t"
             f"{'+' if k % 2 == 0 else '-'}\t{c}\t{100 + 60 * k}-{100 + 60 * k + 29}" for k, c in enumerate(codons)]
    aln = [(28, "+" if k % 2 == 0 else "-", "c1", 100 + 60 * k + 3 * j - (12 if k % 2 == 0 else -12 - 2), 1 + (j % 3))
           for k in range(len(codons)) for j in range(8)]
    case = dict(name="start_codon", contigs=[["c1", 600]], index=[header] + lines, alignments=aln,
                psite_offsets={"28": 12})
    return case


def count_orfs_cases():
    """count_orfs (count_orfs.py:28-89) of the unmodified reference on the index + TSV text of the committed
    pipeline cases, for several feature sets and both report_all settings."""
    from ribotricer.count_orfs import count_orfs

    pipe = json.load(gzip.open(os.path.join(HERE, "pipeline_cases.json.gz"), "rt"))["cases"]
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for case in pipe:
            idx = os.path.join(tmp, "idx.tsv")
            with open(idx, "w") as fh:
                fh.write("\n".join(case["index"]) + "\n")
            types = sorted({line.split("\t")[1] for line in case["index"][1:]})
            feature_sets = [["annotated"], [t for t in types if t != "annotated"], types]
            for k, tsv in enumerate(case["tsv"]):
                tsv_path = os.path.join(tmp, "in.tsv")
                with open(tsv_path, "w") as fh:
                    fh.write(tsv["text"])
                for features in feature_sets:
                    for report_all in (False, True):
                        dst = os.path.join(tmp, "counts.tsv")
                        count_orfs(idx, tsv_path, set(features), dst, report_all)
                        out.append({"case": case["name"], "tsv": k, "features": features, "report_all": report_all,
                                    "text": open(dst).read()})
    return out


def learn_cutoff_cases():
    """determine_cutoff_tsv (learn_cutoff.py:35-144) of the unmodified reference: its stdout for Ribo-seq /
    RNA-seq TSV pairs built from the committed pipeline cases (the RNA-seq TSV is the Ribo-seq one with the
    phase_score column permuted, so that the two differ while holding the same ORFs)."""
    import io
    from contextlib import redirect_stdout

    from ribotricer.learn_cutoff import determine_cutoff_tsv

    pipe = {c["name"]: c for c in json.load(gzip.open(os.path.join(HERE, "pipeline_cases.json.gz"), "rt"))["cases"]}
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, reps, ratio, filter_by, n_files in (("yeastlike", 300, 0.33, None, 1), ("ragged", 257, 0.5, ["protein_coding", "lncRNA"], 2),
                                                     ("yeastlike", 64, 0.9, ["Protein_Coding"], 2)):
            text = pipe[name]["tsv"][0]["text"]                 # report_all=True: every ORF of the index
            lines = text.rstrip("\n").split("\n")
            rows = [ln.split("\t") for ln in lines[1:]]
            rng = np.random.default_rng(len(rows) + reps)
            ribo_files, rna_files, rna_texts = [], [], []
            for k in range(n_files):
                ribo = os.path.join(tmp, f"ribo{k}.tsv")
                with open(ribo, "w") as fh:
                    fh.write(text)
                perm = rng.permutation(len(rows))
                rna_rows = [r[:3] + [rows[perm[i]][3]] + r[4:] for i, r in enumerate(rows)]
                rna_text = "\n".join([lines[0]] + ["\t".join(r) for r in rna_rows]) + "\n"
                rna = os.path.join(tmp, f"rna{k}.tsv")
                with open(rna, "w") as fh:
                    fh.write(rna_text)
                ribo_files.append(ribo)
                rna_files.append(rna)
                rna_texts.append(rna_text)
            sink = io.StringIO()
            with redirect_stdout(sink):
                determine_cutoff_tsv(ribo_files, rna_files, filter_by, ratio, reps)
            out.append({"case": name, "reps": reps, "sampling_ratio": ratio, "filter_by": filter_by, "n_files": n_files,
                        "rna_texts": rna_texts, "stdout": sink.getvalue()})
    return out


def run_reference_pipeline(case, ref):
    from collections import Counter, defaultdict

    from ribotricer.detect_orfs import (export_orf_coverages, export_wig,
                                        merge_read_lengths, orf_coverage)
    from ribotricer.orf import ORF

    alignments = defaultdict(lambda: defaultdict(Counter))
    for length, strand, chrom, pos, n in case["alignments"]:
        alignments[length][strand][(chrom, pos)] += n
    offsets = {int(k): v for k, v in case["psite_offsets"].items()}
    merged = merge_read_lengths(alignments, offsets)
    case["merged"] = sorted([s, c, p, n] for s in merged for (c, p), n in merged[s].items())
    with tempfile.TemporaryDirectory() as tmp:
        idx = os.path.join(tmp, "idx.tsv")
        with open(idx, "w") as fh:
            fh.write("\n".join(case["index"]) + "\n")
        profiles = []
        for line in case["index"][1:]:
            orf = ORF.from_string(line + "\n")
            profiles.append([orf.oid, orf_coverage(orf, merged)])
        case["profiles"] = profiles
        prefix = os.path.join(tmp, "out")
        export_wig(merged, prefix)
        case["wig"] = {}
        for tag in ("pos", "neg"):
            path = f"{prefix}_{tag}.wig"
            if os.path.exists(path):
                case["wig"][tag] = open(path).read()
        case["tsv"] = []
        for params in PARAM_SETS:
            export_orf_coverages(idx, merged, prefix, **params)
            case["tsv"].append({"params": params,
                                "text": open(f"{prefix}_translating_ORFs.tsv").read()})
    return case


def main():
    ref = ref_import.load()
    if "split_bam" in sys.argv[1:] or len(sys.argv) == 1:
        sb = split_bam_case()
        with gzip.open(os.path.join(HERE, "split_bam_case.json.gz"), "wt") as fh:
            json.dump({"versions": versions(), "case": sb}, fh, separators=(",", ":"))
        for r in sb["runs"]:
            print("split_bam", r["protocol"], r["read_lengths"], "keys:", len(r["alignments"]),
                  r["summary"].split("\n\nlength")[0].replace("\n\t", " "))
    if "start_codon" in sys.argv[1:] or len(sys.argv) == 1:
        sc = run_reference_pipeline(start_codon_case(), ref)
        with gzip.open(os.path.join(HERE, "start_codon_case.json.gz"), "wt") as fh:
            json.dump({"versions": versions(), "cases": [sc]}, fh, separators=(",", ":"))
        print("start_codon:", [r.split("\t")[16] for r in sc["tsv"][0]["text"].split("\n")[1:] if r])
    if "infer_protocol" in sys.argv[1:] or len(sys.argv) == 1:
        ip = infer_protocol_case()
        with gzip.open(os.path.join(HERE, "infer_protocol_case.json.gz"), "wt") as fh:
            json.dump({"versions": versions(), "case": ip}, fh, separators=(",", ":"))
        for r in ip["runs"]:
            print("infer_protocol", r["n_reads"], r["protocol"], r["text"].replace("\n", " | "), "warnings:", r["warnings"])
    if "count_orfs" in sys.argv[1:]:      # (a full run does this last, after the pipeline cases it reads)
        cc = count_orfs_cases()
        with gzip.open(os.path.join(HERE, "count_orfs_cases.json.gz"), "wt") as fh:
            json.dump({"versions": versions(), "cases": cc}, fh, separators=(",", ":"))
        print("count_orfs cases:", len(cc), "non-trivial:", sum(c["text"].count("\n") > 1 for c in cc))
    if "learn_cutoff" in sys.argv[1:]:
        lc = learn_cutoff_cases()
        with gzip.open(os.path.join(HERE, "learn_cutoff_cases.json.gz"), "wt") as fh:
            json.dump({"versions": versions(), "cases": lc}, fh, separators=(",", ":"))
        print("learn_cutoff cases:", len(lc), lc[0]["stdout"].replace("\n", " | ")[:400])
    if len(sys.argv) > 1:
        return
    from ribotricer.statistics import phasescore

    ps = {"versions": versions(), "cases": make_phasescore_cases(phasescore)}
    with gzip.open(os.path.join(HERE, "phasescore_cases.json.gz"), "wt") as fh:
        json.dump(ps, fh, separators=(",", ":"))
    print("phasescore cases:", len(ps["cases"]))

    cases = [
        appendix_a_case(),
        random_case("yeastlike", 1001, 120, [("chrI", 6000), ("chrII", 4000), ("chrM", 900)]),
        random_case("ragged", 1002, 90, [("c1", 5000), ("c2", 2500)], unknown_chrom=True,
                    odd_lengths=True),
    ]
    done = [run_reference_pipeline(c, ref) for c in cases]
    with gzip.open(os.path.join(HERE, "pipeline_cases.json.gz"), "wt") as fh:
        json.dump({"versions": versions(), "cases": done}, fh, separators=(",", ":"))
    for c in done:
        print(c["name"], "orfs:", len(c["index"]) - 1, "alignment keys:", len(c["alignments"]))
    mg = metagene_case()
    with gzip.open(os.path.join(HERE, "metagene_case.json.gz"), "wt") as fh:
        json.dump({"versions": versions(), "case": mg}, fh, separators=(",", ":"))
    print("metagene: kept", mg["kept_lengths"], "offsets", mg["psite_offsets"], "planted", mg["planted"])
    cc = count_orfs_cases()
    with gzip.open(os.path.join(HERE, "count_orfs_cases.json.gz"), "wt") as fh:
        json.dump({"versions": versions(), "cases": cc}, fh, separators=(",", ":"))
    print("count_orfs cases:", len(cc))
    lc = learn_cutoff_cases()
    with gzip.open(os.path.join(HERE, "learn_cutoff_cases.json.gz"), "wt") as fh:
        json.dump({"versions": versions(), "cases": lc}, fh, separators=(",", ":"))
    print("learn_cutoff cases:", len(lc))


if __name__ == "__main__":
    main()
