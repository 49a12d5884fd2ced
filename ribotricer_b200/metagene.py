"""Metagene profiles and P-site offset inference (mirror of ribotricer/metagene.py), plus the
protocol inference of ribotricer/infer_protocol.py on decoded read columns.

SURVEY.md 8(f) "next #1".  The heavy part -- coverage of every read length over the first
600 positions of every annotated CDS -- reuses the K1 and K4 kernels: one K1 pass per length
(that length only, offset 0, i.e. raw 5' ends as in ``alignments[length]``, bam.py:135), one K4
gather over a truncated copy of the annotated ORFs, then K1 with weight -1 to recycle the
scratch coverage.  The per-position normalisation, the pandas-free bookkeeping and the
cross-correlation (np.correlate, metagene.py:319) are cheap host code.
"""
from __future__ import annotations

import sys
from collections import OrderedDict

import numpy as np

from .const import CUTOFF, TYPICAL_OFFSET


# ------------------------------------------------------------------ infer_protocol.py:34-124
def infer_protocol(reads, refseq: dict, prefix: str, n_reads: int = 20000) -> str:
    """'forward' or 'reverse' from the first ``n_reads`` (+1, infer_protocol.py:79) uniquely mapped reads that
    overlap exactly one annotated ORF span (infer_protocol.py:75-105); writes ``{prefix}_protocol.txt``.

    The reference looks overlaps up in a quicksect tree, whose ``find`` is inclusive on both ends: a span
    (start, end, strand) is returned for the query (reference_start, reference_end) when start <= reference_end and
    end >= reference_start.  ``reads.cols`` may carry the host-only columns "pos" / "ref_end" (both BAM decoders
    fill them); without them (synthetic columns) reference_start / reference_end are first / last + 1, which is
    what they are for every CIGAR that neither starts nor ends in a deletion or skip.
    """
    cols = reads.cols
    by_contig = {}
    for chrom, spans in refseq.items():
        arr = np.array(spans, dtype=np.int64).reshape(-1, 3)
        arr = arr[np.argsort(arr[:, 0], kind="stable")]
        # spans sorted by start; their ends sorted; and, for every prefix of the start order, which span ends last
        last_end = np.zeros(len(arr), np.int64)
        if len(arr):
            running = np.maximum.accumulate(arr[:, 1])
            is_new = np.concatenate([[True], arr[1:, 1] > running[:-1]])          # a span that ends later than all before it
            last_end = np.maximum.accumulate(np.where(is_new, np.arange(len(arr)), 0))
        by_contig[chrom] = (arr, np.sort(arr[:, 1]), last_end, bool((arr[:, 0] > arr[:, 1]).any()))
    names = reads.contig_names
    counts = {"++": 0, "--": 0, "+-": 0, "-+": 0}
    n = len(cols["ref_id"])
    flag_all, mapq_all, nh_all, ref_all = cols["flag"], cols["mapq"], cols["nh"], cols["ref_id"]
    if "pos" in cols and "ref_end" in cols:
        start_all, end_all = cols["pos"], cols["ref_end"]
    else:
        start_all = cols["first"]
        end_all = np.where(flag_all & 0x4, -1, cols["last"].astype(np.int64) + 1)
    iteration = 0
    chunk = 1 << 16
    for lo in range(0, n, chunk):
        if iteration > n_reads:           # infer_protocol.py:79
            break
        hi = min(n, lo + chunk)
        fl = flag_all[lo:hi].astype(np.int64)
        nh = nh_all[lo:hi]
        # is_read_uniq_mapping is used bare here (infer_protocol.py:80): truthy only for True, i.e. not secondary
        # and (NH == 1, or no NH tag and MAPQ 255); None (no tag, other MAPQ) is falsy
        uniq = ((fl & 0x100) == 0) & np.where(nh != 0, nh == 1, mapq_all[lo:hi] == 255)
        ref = ref_all[lo:hi]
        ok = uniq & (ref >= 0) & (ref < len(names)) & (end_all[lo:hi] >= 0)     # chrom / mapped_end not None (:87)
        cand = lo + np.flatnonzero(ok)
        if len(cand) == 0:
            continue
        # for every candidate read: does it overlap exactly one annotated span, and on which strand is that span
        single = np.zeros(len(cand), bool)
        gene_plus = np.zeros(len(cand), bool)
        c_ref = np.asarray(ref_all[cand], np.int64)
        c_start, c_end = np.asarray(start_all[cand], np.int64), np.asarray(end_all[cand], np.int64)
        for r in np.unique(c_ref).tolist():
            entry = by_contig.get(names[r])
            if entry is None:
                continue
            spans, ends_sorted, last_end, odd = entry
            sel = np.flatnonzero(c_ref == r)
            st, en = c_start[sel], c_end[sel]
            if odd or (en < st).any():
                # a span that ends before it starts (or such a read): the inclusive test of the tree, span by span
                for j, s0, e0 in zip(sel.tolist(), st.tolist(), en.tolist()):
                    k = int(np.searchsorted(spans[:, 0], e0, side="right"))
                    hit = spans[:k][spans[:k, 1] >= s0]
                    single[j] = len(hit) == 1
                    gene_plus[j] = len(hit) == 1 and hit[0, 2] == 1
                continue
            # spans that start at or before the read's end, minus those that end before its start (each of these also
            # starts before the read's end): the inclusive overlap count of the reference's tree (`find`)
            k = np.searchsorted(spans[:, 0], en, side="right")
            n_hit = k - np.searchsorted(ends_sorted, st, side="left")
            one = n_hit == 1
            single[sel] = one
            # the one span that overlaps is the one that ends last among the first k
            gene_plus[sel[one]] = spans[last_end[k[one] - 1], 2] == 1
        hits = np.flatnonzero(single)[:max(0, n_reads + 1 - iteration)]           # in read order, until :79 stops the loop
        minus = (np.asarray(flag_all[cand[hits]], np.int64) & 0x10) != 0
        gp = gene_plus[hits]
        counts["++"] += int((~minus & gp).sum())
        counts["+-"] += int((~minus & ~gp).sum())
        counts["-+"] += int((minus & gp).sum())
        counts["--"] += int((minus & ~gp).sum())
        iteration += len(hits)
    for k in counts:                      # pseudocounts, infer_protocol.py:107-110
        counts[k] += 1
    total = sum(counts.values())
    fwd = counts["++"] + counts["--"]
    rev = counts["-+"] + counts["+-"]
    with open(f"{prefix}_protocol.txt", "w") as output:
        output.write(
            f"In total {total} reads checked:\n"
            f'\tNumber of reads explained by "++, --": {fwd} ({fwd / total:.4f})\n'
            f'\tNumber of reads explained by "+-, -+": {rev} ({rev / total:.4f})\n')
    return "reverse" if rev > fwd else "forward"


# ------------------------------------------------------------------ plotting.py (optional)
def plot_read_lengths(read_length_counts, prefix):
    try:
        import matplotlib
    except ImportError:
        print("matplotlib not available: skipping the read length distribution plot")
        return
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    fig, ax = plt.subplots()
    lengths = sorted(read_length_counts)
    ax.bar(lengths, [read_length_counts[x] for x in lengths])
    ax.set_xlabel("Read length")
    ax.set_ylabel("Number of reads")
    fig.savefig(f"{prefix}_read_length_dist.pdf")
    plt.close(fig)


def plot_metagene(metagenes, read_length_counts, prefix):
    try:
        import matplotlib
    except ImportError:
        print("matplotlib not available: skipping the metagene plots")
        return
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    from matplotlib.backends.backend_pdf import PdfPages
    with PdfPages(f"{prefix}_metagene_plots.pdf") as pdf:
        for length in sorted(metagenes):
            idx, prof = metagenes[length][0]
            fig, ax = plt.subplots()
            ax.vlines(idx, 0, prof)
            ax.set_title(f"{length} nt reads ({read_length_counts.get(length, 0)})")
            pdf.savefig(fig)
            plt.close(fig)


# ------------------------------------------------------------------ metagene.py:95-265
def _metagene_windows(cds, engine, max_positions, offset_5p, offset_3p):
    """Truncated copies of the annotated ORFs covering what next_genome_pos walks
    (metagene.py:42-92): leader + intervals + trailer in transcript direction, cut after
    ``max_positions`` positions.  Returned as CSR arrays for the auxiliary index."""
    ptr, st, en, contig, strand, lens = [0], [], [], [], [], []
    for orf in cds:
        minus = orf.strand == "-"
        o5, o3 = (offset_3p, offset_5p) if minus else (offset_5p, offset_3p)   # metagene.py:129-130
        ivs = [list(iv) for iv in orf.intervals]
        ivs[0][0] -= o5                      # leader_iv joins the first interval
        ivs[-1][1] += o3                     # trailer_iv joins the last one
        left = max_positions
        kept = []
        for s, e in (reversed(ivs) if minus else ivs):
            if left <= 0:
                break
            n = min(left, e - s + 1)
            kept.append((e - n + 1, e) if minus else (s, s + n - 1))
            left -= n
        if minus:
            kept.reverse()
        for s, e in kept:
            st.append(s)
            en.append(e)
        ptr.append(len(st))
        contig.append(engine.contig_id(orf.chrom))
        strand.append(0 if orf.strand == "+" else 1 if minus else 2)
        lens.append(max_positions - max(left, 0))
    return (np.array(ptr, np.int64), np.array(st, np.int32), np.array(en, np.int32),
            np.array(contig, np.int32), np.array(strand, np.uint8), np.array(lens, np.int64))


def metagene_sums(lib, flat, out_ptr, width: int):
    """``(start_sum, start_cnt, stop_sum, stop_cnt)`` over the rows ``flat[out_ptr[i]:out_ptr[i+1]]`` (metagene.py:204-252)."""
    import ctypes as C

    flat = np.ascontiguousarray(flat, np.int32)
    out_ptr = np.ascontiguousarray(out_ptr, np.int64)
    sums = [np.zeros(width, np.float64), np.zeros(width, np.int64), np.zeros(width, np.float64), np.zeros(width, np.int64)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
    rc = lib.rt_metagene_sums(p(flat), p(out_ptr), len(out_ptr) - 1, int(width), *[p(a) for a in sums])
    if rc != 0:
        raise ValueError(f"rt_metagene_sums failed ({rc}): {lib.rt_io_last_error().decode()}")
    return tuple(sums)


def metagene_coverage(cds, alignments, read_lengths: dict, prefix: str, max_positions: int = 600,
                      offset_5p: int = 20, offset_3p: int = 0, meta_min_reads: int = 100000):
    """metagene.py:160-265.  Returns ``{length: ((index_5p, profile_5p), (index_3p, profile_3p),
    phasescore_5p, valid_5p, phasescore_3p, valid_3p)}`` and writes the two profile TSVs.
    Like the reference, lengths with fewer than ``meta_min_reads`` reads are deleted from the
    caller's ``read_lengths`` dict (metagene.py:200-202)."""
    from .detect_orfs import _aux_engine
    from .statistics import phasescore

    for length, reads in list(read_lengths.items()):
        if reads < meta_min_reads:
            del read_lengths[length]
    eng = alignments.engine
    metagenes = {}
    if cds and read_lengths:
        eng.ensure_dense()                                           # the windows address the genome-wide planes
        aux = _aux_engine(eng)
        ptr, st, en, contig, strand, lens = _metagene_windows(cds, eng, max_positions, offset_5p, offset_3p)
        aux.set_index(ptr, st, en, contig, strand)
        cov = eng.new_coverage()
        sel = np.arange(len(cds), dtype=np.int64)
        width = int(lens.max()) if len(lens) else 0
    for length in read_lengths:
        if not cds:
            # no annotated ORF: the reference's per-ORF loop never runs and every length gets empty profiles
            # with phasescore([]) = (0.0, 0) (metagene.py:204-252); align_metagenes then ends in its sys.exit
            metagenes[length] = (([], []), ([], []), 0.0, 0, 0.0, 0)
            continue
        alignments.bin_into(cov, {length: 0})                       # alignments[length], bam.py:135
        out_ptr, flat = aux.gather_profiles(cov, sel, lens)
        alignments.bin_into(cov, {length: 0}, weight=-1)            # scratch back to zero
        # rows normalised by their mean, summed aligned at the start (column k <-> index k - offset_5p) and aligned at
        # the end (metagene.py:140-155): rt_metagene_sums, all host cores
        start_sum, start_cnt, stop_sum, stop_cnt = metagene_sums(eng.lib, flat, out_ptr, width)
        ks, ke = start_cnt > 0, stop_cnt > 0
        prof5 = (start_sum[ks] / start_cnt[ks]).tolist()
        prof3 = (stop_sum[ke] / stop_cnt[ke]).tolist()
        idx5 = (np.flatnonzero(ks) - offset_5p).tolist()
        idx3 = (np.flatnonzero(ke) - (width - 1) + offset_3p).tolist()
        ps5, v5 = phasescore(prof5, engine=eng)
        ps3, v3 = phasescore(prof3, engine=eng)
        metagenes[length] = ((idx5, prof5), (idx3, prof3), ps5, v5, ps3, v3)
    to_write_5p = "fragment_length\toffset_5p\tprofile\tphase_score\tvalid_codons\n"
    to_write_3p = "fragment_length\toffset_3p\tprofile\tphase_score\tvalid_codons\n"
    for length in sorted(metagenes):
        m = metagenes[length]
        to_write_5p += f"{length}\t{offset_5p}\t{m[0][1]}\t{m[2]}\t{m[3]}\n"
        to_write_3p += f"{length}\t{offset_3p}\t{m[1][1]}\t{m[4]}\t{m[5]}\n"
    with open(f"{prefix}_metagene_profiles_5p.tsv", "w") as output:
        output.write(to_write_5p)
    with open(f"{prefix}_metagene_profiles_3p.tsv", "w") as output:
        output.write(to_write_3p)
    return metagenes


# ------------------------------------------------------------------ metagene.py:268-328
def align_metagenes(metagenes, read_lengths: dict, prefix: str, phase_score_cutoff: float = CUTOFF,
                    remove_nonperiodic: bool = False):
    """P-site offset of every read length = lag of its 5' metagene against the most abundant
    length's + TYPICAL_OFFSET (metagene.py:309-324)."""
    if remove_nonperiodic:
        for length in list(metagenes):
            if metagenes[length][2] < phase_score_cutoff:       # metagene.py:298-302
                del read_lengths[length]
                del metagenes[length]
    if len(read_lengths) == 0:
        sys.exit(f"WARNING: no periodic read length found... using cutoff {phase_score_cutoff}")
    psite_offsets = OrderedDict()
    base = n_reads = 0
    for length, reads in list(read_lengths.items()):
        if reads > n_reads:
            base, n_reads = length, reads
    reference = np.asarray(metagenes[base][0][1], np.float64)
    to_write = f"relative lag to base: {base}\n"
    for length in list(metagenes):
        cov = np.asarray(metagenes[length][0][1], np.float64)
        xcorr = np.correlate(reference, cov, "full")
        origin = len(xcorr) // 2
        bound = min(base, length)
        xcorr = xcorr[(origin - bound):(origin + bound)]
        lag = int(np.argmax(xcorr) - len(xcorr) // 2)
        psite_offsets[length] = lag + TYPICAL_OFFSET
        to_write += f"\tlag of {length}: {lag}\n"
    with open(f"{prefix}_psite_offsets.txt", "w") as output:
        output.write(to_write)
    return psite_offsets
