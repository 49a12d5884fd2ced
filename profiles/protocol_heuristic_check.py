#!/usr/bin/env python
"""What does the UNMODIFIED reference infer_protocol (infer_protocol.py:34-124) say about the synthetic C1 library?

Authoring container only (needs /root/reference).  Writes the first N reads of synthetic C1 as a real BAM
(tests/bam_writer.py), the index as the 11-column TSV, runs the reference (pysam / quicksect restated, see
oracle/ref_import.py) and the product's infer_protocol on the decoded columns, and prints both verdicts.

    python profiles/protocol_heuristic_check.py [n_reads_in_bam]
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bam_writer as W  # noqa: E402
from oracle import ref_import  # noqa: E402
from ribotricer_b200 import metagene, synth  # noqa: E402
from ribotricer_b200.bam import read_bam_columns_native  # noqa: E402
from ribotricer_b200.detect_orfs import parse_ribotricer_index  # noqa: E402

n_bam = int(sys.argv[1]) if len(sys.argv) > 1 else 80_000
ref_import.load()
from ribotricer.detect_orfs import parse_ribotricer_index as ref_parse  # noqa: E402
from ribotricer.infer_protocol import infer_protocol as ref_infer  # noqa: E402

cfg = synth.config("C1")
idx = synth.make_index(cfg)
reads = synth.reads_to_numpy(synth.make_reads(cfg, idx, device="cpu"))
with tempfile.TemporaryDirectory() as tmp:
    index_path = os.path.join(tmp, "idx.tsv")
    idx.write_tsv(index_path, 0, idx.annotated_rows + 10)
    recs = []
    for i in range(n_bam):
        aux = W.aux_field("NH", "C", int(reads["nh"][i])) if reads["nh"][i] else b""
        L = int(reads["mlen"][i])
        recs.append(W.record(int(reads["ref_id"][i]), int(reads["first"][i]), int(reads["mapq"][i]), int(reads["flag"][i]),
                             [("M", L)], name=b"r", aux=aux))
    bam = os.path.join(tmp, "lib.bam")
    W.write_bam(bam, list(zip(idx.contig_names, idx.contig_len.tolist())), recs)
    _, tree = ref_parse(index_path)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        verdict = ref_infer(bam, tree, os.path.join(tmp, "ref"))
    print("reference :", verdict, "|", open(os.path.join(tmp, "ref_protocol.txt")).read().replace("\n", " | "))
    _, refseq = parse_ribotricer_index(index_path)
    ours = metagene.infer_protocol(read_bam_columns_native(bam), refseq, os.path.join(tmp, "ours"))
    print("product   :", ours, "|", open(os.path.join(tmp, "ours_protocol.txt")).read().replace("\n", " | "))
