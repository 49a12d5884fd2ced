"""CPU restatement of ribotricer's detect-orfs scoring path (pure Python / numpy).

TEST INFRASTRUCTURE ONLY -- this module is the *checker*.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  The product (``ribotricer_b200``) never does
and has no CPU fallback.

Parity status: the reference's own tests pin nothing on this path (SURVEY.md
section 4 / 8(c)), so this oracle is pinned instead against outputs of the
reference itself: ``tests/golden/*.json.gz`` were generated in the authoring
container by running the unmodified reference (``oracle/ref_import.py``) with
scipy 1.18.1 / numpy 2.3.5; ``tests/test_oracle_golden.py`` replays them.

Every function cites the reference lines it restates (paths relative to the
reference tree root).  Data structures deliberately mirror the reference's
(dict-of-dict-of-Counter), so these functions are for small inputs; the
scalable dense-array restatement is ``oracle/rt_oracle.c``.
"""
from __future__ import annotations

import math
from collections import Counter, defaultdict

import numpy as np

# ribotricer/common.py:30
SAM_NOT_UNIQ_FLAGS = (4, 20, 256, 272, 2048)

FLAG_UNMAPPED = 0x4
FLAG_REVERSE = 0x10
FLAG_SECONDARY = 0x100
FLAG_QCFAIL = 0x200
FLAG_DUPLICATE = 0x400


# ----------------------------------------------------------------------------
# A1  split_bam on decoded read columns            ribotricer/bam.py:33-153
# ----------------------------------------------------------------------------
def is_read_uniq_mapping(flag: int, mapq: int, nh: int):
    """ribotricer/common.py:33-69.  ``nh`` = value of the NH tag, 0 if absent."""
    if flag & FLAG_SECONDARY:          # common.py:51-52
        return False
    if nh != 0:                        # common.py:54-56 (tag present)
        return nh == 1
    if mapq == 255:                    # common.py:59-60
        return True
    if mapq < 1 or flag in SAM_NOT_UNIQ_FLAGS:   # common.py:61-62
        return False
    return None                        # common.py:63-69 (falsy -> read dropped)


def split_reads(cols: dict, protocol: str, read_lengths=None, contig_names=None):
    """Restates the per-read loop of ``split_bam`` (bam.py:71-137) on columns.

    cols: dict of equal-length sequences ``ref_id, first, last, mlen, flag,
    mapq, nh`` where ``first``/``last`` are the 0-based first/last matched
    reference positions (``get_reference_positions()[0]/[-1]``, bam.py:95) and
    ``mlen`` their count (bam.py:99).  Returns ``(alignments,
    read_length_counts, stats)`` with the reference's nesting
    ``alignments[length][strand][(chrom, pos1)]`` (bam.py:135).
    """
    alignments = defaultdict(lambda: defaultdict(Counter))
    read_length_counts = defaultdict(int)
    st = dict(total=0, qcfail=0, duplicate=0, secondary=0, unmapped=0, multi=0, valid=0)
    n = len(cols["ref_id"])
    for i in range(n):
        flag = int(cols["flag"][i])
        st["total"] += 1
        # filter cascade, bam.py:77-91
        if flag & FLAG_QCFAIL:
            st["qcfail"] += 1
            continue
        if flag & FLAG_DUPLICATE:
            st["duplicate"] += 1
            continue
        if flag & FLAG_SECONDARY:
            st["secondary"] += 1
            continue
        if flag & FLAG_UNMAPPED:
            st["unmapped"] += 1
            continue
        if not is_read_uniq_mapping(flag, int(cols["mapq"][i]), int(cols["nh"][i])):
            st["multi"] += 1
            continue
        map_strand = "-" if flag & FLAG_REVERSE else "+"     # bam.py:94
        length = int(cols["mlen"][i])                        # bam.py:99
        if read_lengths is not None and length not in read_lengths:   # bam.py:101
            continue
        strand = pos = None
        if protocol == "forward":                            # bam.py:105-117
            if map_strand == "+":
                strand, pos = "+", int(cols["first"][i])
            else:
                strand, pos = "-", int(cols["last"][i])
        elif protocol == "reverse":                          # bam.py:118-131
            if map_strand == "+":
                strand, pos = "-", int(cols["last"][i])
            else:
                strand, pos = "+", int(cols["first"][i])
        ref_id = int(cols["ref_id"][i])
        chrom = None
        if ref_id >= 0:
            chrom = contig_names[ref_id] if contig_names is not None else ref_id
        if strand is not None and pos is not None and chrom is not None:   # bam.py:133
            alignments[length][strand][(chrom, pos + 1)] += 1             # bam.py:135
            read_length_counts[length] += 1                              # bam.py:136
            st["valid"] += 1
    return alignments, read_length_counts, st


def bam_summary_text(st: dict, read_length_counts: dict) -> str:
    """bam.py:141-148 (note the missing blank after 'unmapped:' and 'multi:')."""
    s = (
        f"summary:\n\ttotal_reads: {st['total']}\n\tunique_mapped: {st['valid']}\n"
        f"\tqcfail: {st['qcfail']}\n\tduplicate: {st['duplicate']}\n\tsecondary: {st['secondary']}\n"
        f"\tunmapped:{st['unmapped']}\n\tmulti:{st['multi']}\n\nlength dist:\n"
    )
    for length in sorted(read_length_counts):
        s += f"\t{length}: {read_length_counts[length]}\n"
    return s


# ----------------------------------------------------------------------------
# A2  merge_read_lengths                       ribotricer/detect_orfs.py:54-83
# ----------------------------------------------------------------------------
def merge_read_lengths(alignments, psite_offsets: dict):
    merged = defaultdict(Counter)
    for length, offset in psite_offsets.items():       # :74 (other lengths dropped)
        if length not in alignments:
            continue
        for strand, table in alignments[length].items():
            for (chrom, pos), count in table.items():
                shifted = pos + offset if strand == "+" else pos - offset   # :78-81
                merged[strand][(chrom, shifted)] += count                   # :82
    return merged


# ----------------------------------------------------------------------------
# A3  orf_coverage                            ribotricer/detect_orfs.py:134-203
# ----------------------------------------------------------------------------
def orf_profile(chrom, strand, intervals, merged) -> list:
    """``intervals``: list of (start, end), 1-based closed (interval.py:20-71).

    Sorted by start (orf.py:100), concatenated in ascending genomic order
    (detect_orfs.py:176-187), reversed on '-' (:201-202); missing keys and
    unknown strands read as 0.
    """
    cov = []
    table = merged.get(strand) if hasattr(merged, "get") else None
    for start, end in sorted(intervals, key=lambda iv: iv[0]):
        for pos in range(start, end + 1):
            cov.append(table.get((chrom, pos), 0) if table is not None else 0)
    if strand == "-":
        cov.reverse()
    return cov


# ----------------------------------------------------------------------------
# A4  phasescore                                ribotricer/statistics.py:48-115
# ----------------------------------------------------------------------------
_C1, _C2 = math.cos(2 * math.pi / 3), math.cos(4 * math.pi / 3)
_S1, _S2 = math.sin(2 * math.pi / 3), math.sin(4 * math.pi / 3)


def _normalised_codons(values):
    """statistics.py:69-91 -- keep non-all-zero complete triplets, each divided
    by the modulus of its projection on the 0/120/240-degree unit vectors."""
    out = []
    i = 0
    while i + 2 < len(values):
        a, b, c = values[i], values[i + 1], values[i + 2]
        if not (a == b == c == 0):
            real = a + b * _C1 + c * _C2
            image = b * _S1 + c * _S2
            norm = math.sqrt(real**2 + image**2)
            if norm == 0:
                norm = 1
            out += [a / norm, b / norm, c / norm]
        i += 3
    return out


def phasescore_scipy(original_values):
    """Faithful restatement: same SciPy call as statistics.py:101-107.

    This is what the reference *is* on a box that has SciPy; it is slow
    (about 50 ms per 800-nt profile) and is the CPU baseline of bench.py.
    """
    import warnings

    from scipy import signal

    values_all = [float(v) if not isinstance(v, (int, np.integer)) else int(v) for v in original_values]
    coh, valid = 0.0, -1
    for frame in (0, 1, 2):
        norm_vals = _normalised_codons(values_all[frame:])
        k = len(norm_vals) // 3
        if k == 0:
            coh, valid = 0.0, 0                                # :94-95 (reset quirk)
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f, cxy = signal.coherence(
                np.array(norm_vals), np.array([1, 0, 0] * k),
                window=np.array([1.0, 1.0, 1.0]), nperseg=3, noverlap=0,
            )
        score = cxy[np.argwhere(np.isclose(f, 1 / 3.0))[0]][0]  # :108
        if score > coh:                                        # :109-111 (strict)
            coh, valid = score, k
        if valid == -1:                                        # :112-113
            valid = k
    return float(np.sqrt(coh)), valid


def frame_spectra(values):
    """Per-frame (K_f, s_f) where s_f is the magnitude-squared coherence at
    f = 1/3 that ``scipy.signal.coherence`` computes at statistics.py:101-108.

    Restated from SciPy's published definition (Welch / CSD with nperseg=3,
    noverlap=0, boxcar window, detrend='constant'): Cxy = |Pxy|^2/(Pxx*Pyy),
    each P the mean over segments of the bin-1 DFT products.  The template
    segment [1,0,0] has bin-1 value 1 after detrending, so Pyy = 1 * scale and
    the scale factors cancel.  s_f is NaN when Pxx == 0 (all codons uniform).
    """
    w1 = complex(-0.5, -math.sqrt(3.0) / 2.0)   # exp(-2*pi*i/3)
    w2 = complex(-0.5, +math.sqrt(3.0) / 2.0)   # exp(-4*pi*i/3)
    res = []
    for frame in (0, 1, 2):
        nv = _normalised_codons(list(values)[frame:])
        k = len(nv) // 3
        if k == 0:
            res.append((0, float("nan")))
            continue
        pxx = 0.0
        pxy = 0j
        for j in range(k):
            x0, x1, x2 = nv[3 * j], nv[3 * j + 1], nv[3 * j + 2]
            m = (x0 + x1 + x2) / 3.0                       # detrend='constant'
            X = (x0 - m) + (x1 - m) * w1 + (x2 - m) * w2   # rfft bin 1
            pxx += (X * X.conjugate()).real
            pxy += X.conjugate()                           # * Y, Y = 1
        pxx /= k
        pxy /= k
        with np.errstate(all="ignore"):
            s = float(np.float64(abs(pxy) ** 2) / np.float64(pxx))
        res.append((k, s))
    return res


def select_frame(frames):
    """The running-maximum logic of statistics.py:64-115 on per-frame (K, s)."""
    coh, valid = 0.0, -1
    for k, s in frames:
        if k == 0:
            coh, valid = 0.0, 0          # :94-95
            continue
        if s > coh:                      # :109 strict; NaN never wins
            coh, valid = s, k
        if valid == -1:                  # :112-113
            valid = k
    return math.sqrt(coh), valid


def phasescore(values):
    """Closed-form phasescore: (score, valid_codons).  statistics.py:48-115."""
    return select_frame(frame_spectra(values))


def is_frame_tie(frames, tol: float = 1e-12) -> bool:
    """Hazard H1: ``valid_codons`` of the reference is decided
    by ~1e-17 SciPy rounding noise when the running maximum at
    statistics.py:109 meets a score equal to it in exact arithmetic while the
    K differ.  Such ORFs are excluded from bit-exact valid_codons comparison.
    A frame only wins clearly when it beats the running maximum by more than
    ``tol``; anything closer with a different K is a tie until a clear win.
    """
    best, valid = 0.0, -1
    tie = False
    for k, s in frames:
        if k == 0:
            best, valid, tie = 0.0, 0, False
            continue
        if not math.isnan(s):
            if s > best + tol:
                best, valid, tie = s, k, False
            elif abs(s - best) <= tol:
                if valid not in (-1, k):
                    tie = True
                best = max(best, s)
        if valid == -1:
            valid = k
    return tie


# ----------------------------------------------------------------------------
# A5 / A6  derived statistics and status     ribotricer/detect_orfs.py:278-299
# ----------------------------------------------------------------------------
def collapse_coverage_to_codon(cov):
    """common.py:164-180 (chunks of 3 from index 0, trailing partial included)."""
    return [sum(cov[i:i + 3]) for i in range(0, len(cov), 3)]


def score_profile(cov, phase_score_cutoff=0.428571428571, min_valid_codons=5,
                  min_reads_per_codon=0, min_valid_codons_ratio=0,
                  min_density_over_orf=0.0, scorer=None):
    scorer = scorer or phasescore
    count = sum(cov)                                   # :278
    length = len(cov)                                  # :279
    coh, valid = scorer(cov)                           # :280
    n_codons = max(1, length // 3)                     # :281
    codon_cov = np.array(collapse_coverage_to_codon(cov))   # :284
    ratio = valid / n_codons                           # :285
    density = np.sum(codon_cov) / n_codons             # :287
    ok = (coh >= phase_score_cutoff and valid >= min_valid_codons
          and bool(np.all(codon_cov >= min_reads_per_codon))
          and ratio >= min_valid_codons_ratio and density >= min_density_over_orf)   # :289-299
    return dict(score=np.float64(coh), valid=valid, count=count, length=length,
                ratio=ratio, density=density,
                min_codon=int(codon_cov.min()) if len(codon_cov) else 0,
                status="translating" if ok else "nontranslating")


# ----------------------------------------------------------------------------
# A7  TSV emit                        ribotricer/detect_orfs.py:241-269,300-324
# ----------------------------------------------------------------------------
TSV_COLUMNS = [
    "ORF_ID", "ORF_type", "status", "phase_score", "read_count", "length",
    "valid_codons", "valid_codons_ratio", "read_density", "transcript_id",
    "transcript_type", "gene_id", "gene_name", "gene_type", "chrom", "strand",
    "start_codon", "profile",
]


def parse_index_line(line: str):
    """orf.py:121-182 (+ oid derivation orf.py:100-103)."""
    fields = line.split("\t")
    if len(fields) != 11:
        raise SystemExit("Error: unexpected number of columns found for index file\n"
                         "please run ribotricer prepare-orfs to regenerate")
    ivs = []
    for group in fields[10].split(","):
        s, e = group.split("-")
        ivs.append((int(s), int(e)))
    ivs.sort(key=lambda iv: iv[0])
    length = sum(e - s + 1 for s, e in ivs)
    return dict(category=fields[1], tid=fields[2], ttype=fields[3], gid=fields[4],
                gname=fields[5], gtype=fields[6], chrom=fields[7], strand=fields[8],
                start_codon=fields[9], intervals=ivs,
                oid=f"{fields[2]}_{ivs[0][0]}_{ivs[-1][1]}_{length}")


def export_orf_coverages(index_path, merged, prefix, phase_score_cutoff=0.428571428571,
                         min_valid_codons=5, min_reads_per_codon=0,
                         min_valid_codons_ratio=0, min_density_over_orf=0.0,
                         report_all=False, scorer=None):
    """detect_orfs.py:206-324: one row per index row in index order."""
    with open(index_path) as anno, open(f"{prefix}_translating_ORFs.tsv", "w") as out:
        out.write("\t".join(TSV_COLUMNS) + "\n")
        anno.readline()
        for line in anno:
            orf = parse_index_line(line)
            cov = orf_profile(orf["chrom"], orf["strand"], orf["intervals"], merged)
            r = score_profile(cov, phase_score_cutoff, min_valid_codons, min_reads_per_codon,
                              min_valid_codons_ratio, min_density_over_orf, scorer)
            if not report_all and r["status"] == "nontranslating":
                continue
            out.write("\t".join(str(x) for x in (
                orf["oid"], orf["category"], r["status"], r["score"], r["count"],
                r["length"], r["valid"], r["ratio"], r["density"], orf["tid"],
                orf["ttype"], orf["gid"], orf["gname"], orf["gtype"], orf["chrom"],
                orf["strand"], orf["start_codon"], cov)))
            # start_codon carries the line's trailing newline only if it is the
            # last field; it is not (coordinate is), so terminate the row here.
            out.write("\n")


def export_wig(merged, prefix):
    """detect_orfs.py:327-351."""
    for strand in merged:
        text, cur = "", ""
        for chrom, pos in sorted(merged[strand]):
            if chrom != cur:
                cur = chrom
                text += f"variableStep chrom={chrom}\n"
            text += f"{pos}\t{merged[strand][(chrom, pos)]}\n"
        with open(f"{prefix}_pos.wig" if strand == "+" else f"{prefix}_neg.wig", "w") as fh:
            fh.write(text)
